#!/usr/bin/env python3
"""Print the CTA-0 timeline of one align() of the c2 workload: where each evaluation's time goes
(local derivative work / grid reduction wait / Newton+line-search step), cold L2 vs warm L2."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    import toyslam_b200 as nb
    ap = argparse.ArgumentParser()
    ap.add_argument("--method", default="DIRECT7")
    ap.add_argument("--map-points", type=int, default=1_000_000)
    ap.add_argument("--map-scans", type=int, default=31)
    ap.add_argument("--azimuth-steps", type=int, default=1875)
    ap.add_argument("--cache", default=None)
    ap.add_argument("--eps", type=float, default=0.1)
    args = ap.parse_args()
    w = bench.make_workload(args, 0)
    ndt = nb.NormalDistributionsTransform()
    ndt.setNeighborhoodSearchMethod(bench.METHODS[args.method])
    ndt.setTransformationEpsilon(args.eps)
    ndt.setInputTarget(w["target"])
    ndt.setInputSource(w["source"])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for label in ("warm", "warm", "cold", "cold"):
        if label == "cold":
            flush.fill_(1)
            torch.cuda.synchronize()
        ndt.align_async()
        ndt.sync()
        t = ndt.timeline() / 1e3
        r = ndt.result()
        print("%s L2: kernel %.1f us, %d evals + %d hessian passes, %d Newton iterations" %
              (label, ndt.last_align_ms() * 1e3, r["n_evaluations"], r["n_hessian_passes"], r["iterations"]))
        for i, row in enumerate(t):
            print("   eval %2d: start %8.1f | local %7.1f us | reduce wait %6.1f us | step %6.1f us" %
                  (i, row[0], row[1] - row[0], row[2] - row[1], row[3] - row[2]))


if __name__ == "__main__":
    main()
