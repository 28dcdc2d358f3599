#!/usr/bin/env python3
"""Generate the committed golden fixtures from the reference's bundled scans.

Run HERE (in the build container, where /root/reference exists):
    python tests/golden/make_fixtures.py

Inputs  : /root/reference/ndt_omp/data/251370668.pcd (target), 251371071.pcd (source)
          — PCD v0.7, DATA binary, fields x y z intensity (4 x f32).
Outputs : tests/golden/pair_ds0p1.npz   0.1 m pcl::VoxelGrid downsample of both scans, exactly what
                                        ndt_omp/apps/align.cpp:57-69 feeds to NDT (config 1)
          tests/golden/pair_ds0p3.npz   0.3 m downsample (ndt_rosbag_mapping_node.cpp:88 default;
                                        line-search fixture B of SURVEY Appendix A)
          tests/golden/pair_raw.npz     the two RAW scans (no downsample): line-search fixture A of SURVEY Appendix A
                                        (2 Newton iterations, 23 evaluations = 2 x 10 extra More-Thuente trials,
                                        2 Hessian-only passes) — the heaviest exercise of trialValueSelectionMT /
                                        updateIntervalMT / computeHessian the bundled data offers
          tests/golden/golden.json      the reference's published known answers (ndt_omp/README.md:13-46) and the
                                        counts / final poses of SURVEY Appendix A (independent numpy restatement)

The downsample here is an independent numpy restatement of pcl::VoxelGrid<PointXYZ>::applyFilter
(fp32 key arithmetic identical to voxel_grid_covariance_omp_impl.hpp:218-223, fp32 centroid summed
in input order inside each voxel, output ordered by voxel index).  tests/test_oracle_golden.py
checks the C++ oracle's own downsample against these arrays bit-for-bit.
"""
import json
import os
import sys

import numpy as np

REF = "/root/reference/ndt_omp"
HERE = os.path.dirname(os.path.abspath(__file__))


def read_pcd_xyz(path):
    with open(path, "rb") as f:
        raw = f.read()
    marker = b"DATA binary\n"
    off = raw.index(marker) + len(marker)
    header = raw[:off].decode("ascii")
    fields = sizes = None
    npts = None
    for line in header.splitlines():
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "FIELDS":
            fields = tok[1:]
        elif tok[0] == "SIZE":
            sizes = [int(t) for t in tok[1:]]
        elif tok[0] == "POINTS":
            npts = int(tok[1])
    assert fields[:3] == ["x", "y", "z"] and all(s == 4 for s in sizes)
    arr = np.frombuffer(raw, dtype=np.float32, count=npts * len(fields), offset=off).reshape(npts, len(fields))
    return np.ascontiguousarray(arr[:, :3])


def voxelgrid_downsample(xyz, leaf):
    """pcl::VoxelGrid<PointXYZ> centroid downsample, fp32 throughout.  leaf: scalar or (lx, ly, lz)."""
    xyz = np.asarray(xyz, dtype=np.float32)
    leaf = np.asarray(leaf, dtype=np.float32)
    inv = np.float32(1.0) / leaf
    mn = xyz.min(axis=0)
    mx = xyz.max(axis=0)
    min_b = np.floor(mn * inv).astype(np.int32)
    max_b = np.floor(mx * inv).astype(np.int32)
    div_b = max_b - min_b + 1
    ijk = (np.floor(xyz * inv) - min_b.astype(np.float32)).astype(np.int32)
    idx = ijk[:, 0] + ijk[:, 1] * div_b[0] + ijk[:, 2] * div_b[0] * div_b[1]
    order = np.argsort(idx, kind="stable")
    sidx = idx[order]
    spts = xyz[order]
    starts = np.flatnonzero(np.concatenate(([True], sidx[1:] != sidx[:-1])))
    counts = np.diff(np.concatenate((starts, [len(sidx)])))
    acc = np.zeros((len(starts), 3), dtype=np.float32)
    for k in range(int(counts.max())):  # sequential fp32 sum in input order inside each voxel
        m = counts > k
        acc[m] = acc[m] + spts[starts[m] + k]
    return acc / counts.astype(np.float32)[:, None]


def main():
    tgt = read_pcd_xyz(os.path.join(REF, "data/251370668.pcd"))
    src = read_pcd_xyz(os.path.join(REF, "data/251371071.pcd"))
    print("raw", tgt.shape, src.shape)
    for leaf, name in ((0.1, "pair_ds0p1.npz"), (0.3, "pair_ds0p3.npz")):
        t = voxelgrid_downsample(tgt, leaf)
        s = voxelgrid_downsample(src, leaf)
        print(leaf, t.shape, s.shape)
        np.savez_compressed(os.path.join(HERE, name), target=t, source=s)
    np.savez_compressed(os.path.join(HERE, "pair_raw.npz"), target=tgt, source=src)
    # a small slice of the raw scans so the downsample itself can be checked where /root/reference is absent
    np.savez_compressed(os.path.join(HERE, "raw_head.npz"), target=tgt[:8192], source=src[:8192],
                        target_ds0p1=voxelgrid_downsample(tgt[:8192], 0.1),
                        source_ds0p1=voxelgrid_downsample(src[:8192], 0.1))
    golden = {
        "source": "ndt_omp/README.md:13-46 (rosrun ndt_omp align 251370668.pcd 251371071.pcd, Core i7-6700K)",
        "fitness": {"DIRECT7": 0.214205, "DIRECT1": 0.208511, "KDTREE": 0.213937, "pcl_ndt": 0.213937},
        "raw_points": {"target": int(tgt.shape[0]), "source": int(src.shape[0])},
        # SURVEY Appendix A (values of the survey's independent numpy restatement; class defaults, DIRECT7, identity guess)
        "appendix_a": {
            "config1_DIRECT7": {"iterations": 5, "evaluations": 6, "hessian_passes": 0,
                                "p": [0.471692, 0.111211, -0.023818, 0.005899, -0.001002, -0.010327]},
            "config1_DIRECT26": {"iterations": 2, "evaluations": 4, "hessian_passes": 1,
                                 "p": [0.083691, 0.022196, 0.002907, 0.007005, -0.008102, -0.004304], "fitness": 0.265174},
            "fixture_A_raw_pair": {"valid_voxels": 690, "iterations": 2, "evaluations": 23, "hessian_passes": 2,
                                   "p": [-0.010123, -0.020243, -0.003812, 0.059494, 0.000758, -0.024049]},
            "fixture_B_ds0p3_node_params": {"iterations": 7, "evaluations": 18, "hessian_passes": 1,
                                            "p": [0.461994, 0.134161, -0.032969, 0.006620, -0.002877, -0.010950]},
        },
        "published_ms": {
            "KDTREE_1thr": {"single": 207.697, "10times": 2059.19},
            "KDTREE_8thr": {"single": 54.9903, "10times": 500.51},
            "pcl_ndt": {"single": 282.222, "10times": 2921.92},
            "DIRECT7_1thr": {"single": 139.433, "10times": 1356.79},
            "DIRECT1_1thr": {"single": 34.6418, "10times": 317.03},
            "DIRECT7_8thr": {"single": 63.1442, "10times": 343.336},
            "DIRECT1_8thr": {"single": 17.2353, "10times": 100.025},
        },
    }
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(golden, f, indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
