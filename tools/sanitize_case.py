#!/usr/bin/env python3
"""The smallest program that launches every hand-synchronised kernel of the library once, for compute-sanitizer
(racecheck / memcheck / synccheck).  ONE tool per gpurun call (B200_PROFILING.md):

    compute-sanitizer --tool memcheck  python tools/sanitize_case.py
    compute-sanitizer --tool racecheck python tools/sanitize_case.py
    compute-sanitizer --tool synccheck python tools/sanitize_case.py

Kernels covered: ndt_align_kernel (latency and throughput CTA shapes, single rank; 2 emulated ranks = the mailbox
exchange path), small_build_kernel (fused scan-sized build and VoxelGrid), the staged build (minmax3d, voxel_key with the
digit histograms, digit_base, onesweep passes with decoupled look-back, head_count / head_write, voxel_build), the
fitness and score kernels.  Grids are capped (NDTB200_MAX_CTAS) so that the instrumented cooperative launches stay
co-resident."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("NDTB200_MAX_CTAS", "8")


def main():
    import toyslam_b200 as nb
    from util import load_pair
    tgt, src = load_pair("pair_ds0p3.npz")
    src = src[:2048]
    big = np.concatenate([tgt + np.float32(0.011 * k) for k in range(16)]).astype(np.float32)   # 80 k points: staged build, 2 sort passes
    out = {}
    for shape in (0, 1):
        g = nb.NormalDistributionsTransform()
        g.setMaximumIterations(3)
        g.set_throughput_mode(bool(shape))
        g.setInputTarget(tgt)                      # fused small build (latency shape) / staged build (throughput shape)
        g.setInputSource(src)
        g.align()
        out["align_shape%d" % shape] = g.result()["iterations"]
        out["fitness%d" % shape] = round(g.getFitnessScore(), 6)
        out["score%d" % shape] = round(g.calculateScore(src), 6)
    g = nb.NormalDistributionsTransform()
    g.setMaximumIterations(2)
    g.setInputTarget(big)                          # staged path with the one-sweep sort
    g.setInputSource(src)
    out["voxels_big"] = int(g.map_info()["n_voxels"])
    res = g.align_emulated_ranks(2)
    out["emulated_equal"] = bool(np.array_equal(res[0]["final"], res[1]["final"]))
    out["voxelgrid"] = len(g.voxelgrid_filter(big, 0.5))
    out["voxelgrid_small"] = len(g.voxelgrid_filter(tgt, 0.5))
    g.setNeighborhoodSearchMethod(nb.KDTREE)
    g.align()
    out["kdtree_iterations"] = g.result()["iterations"]
    print("SANITIZE_CASE", out)


if __name__ == "__main__":
    main()
