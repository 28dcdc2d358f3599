"""Shared helpers for the parity tests."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_pair(name="pair_ds0p1.npz"):
    d = np.load(os.path.join(GOLDEN, name))
    return d["target"], d["source"]


def golden():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return json.load(f)


def rel_err(a, b):
    """max |a-b| relative to the max-magnitude entry of the oracle value b."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.abs(b).max()
    if scale == 0:
        return float(np.abs(a - b).max())
    return float(np.abs(a - b).max() / scale)


def pose_matrix(p):
    """fp64 reference pose matrix (Translation * Rx * Ry * Rz) for building test guesses."""
    x, y, z, r, pi, ya = p
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(pi), np.sin(pi), np.cos(ya), np.sin(ya)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rx @ Ry @ Rz
    T[:3, 3] = [x, y, z]
    return T.astype(np.float32)


def transform_delta(Ta, Tb):
    """(translation distance [m], rotation angle [rad]) between two 4x4 transforms."""
    Ta = np.asarray(Ta, dtype=np.float64)
    Tb = np.asarray(Tb, dtype=np.float64)
    dt = float(np.linalg.norm(Ta[:3, 3] - Tb[:3, 3]))
    R = Ta[:3, :3].T @ Tb[:3, :3]
    # angle from the skew part (sin) and the trace (cos): arccos alone has only sqrt(eps) resolution near 0
    w = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    c = (np.trace(R) - 1.0) / 2.0
    return dt, float(np.arctan2(np.linalg.norm(w), c))


def synthetic_scene(n_target=40000, n_source=8000, seed=0, offset=(0.0, 0.0, 0.0), extent=40.0):
    """Small structured cloud pair for parity tests: ground plane + walls + pillars with noise; the source is a
    perturbed sub-sample of the same surfaces.  `offset` shifts the TARGET only (map-scale coordinates, scan-to-map
    use): align with guess = translation(offset)."""
    rng = np.random.default_rng(seed)

    def sample(n):
        kind = rng.integers(0, 4, size=n)
        p = np.empty((n, 3))
        u = rng.uniform(-extent, extent, size=n)
        v = rng.uniform(-extent, extent, size=n)
        h = rng.uniform(0.0, 6.0, size=n)
        # 0: ground, 1: wall x = +-extent/2, 2: wall y = +-extent/3, 3: pillars
        p[:, 0] = np.where(kind == 1, np.sign(u) * extent / 2, u)
        p[:, 1] = np.where(kind == 2, np.sign(v) * extent / 3, v)
        p[:, 2] = np.where(kind == 0, 0.0, h)
        pil = kind == 3
        cx = np.round(u[pil] / 8.0) * 8.0
        cy = np.round(v[pil] / 8.0) * 8.0
        ang = rng.uniform(0, 2 * np.pi, size=pil.sum())
        p[pil, 0] = cx + 0.3 * np.cos(ang)
        p[pil, 1] = cy + 0.3 * np.sin(ang)
        p += rng.normal(0, 0.02, size=p.shape)
        return p

    tgt = sample(n_target)
    src = sample(n_source)
    true = pose_matrix([0.35, -0.22, 0.05, 0.004, -0.006, 0.015]).astype(np.float64)
    inv = np.linalg.inv(true)
    src = src @ inv[:3, :3].T + inv[:3, 3]
    off = np.asarray(offset, dtype=np.float64)
    return (tgt + off).astype(np.float32), src.astype(np.float32)


def oracle_mapping_loop(scans, voxel_leaf=0.3, map_voxel=0.5, eps=0.01, max_iter=64, step=0.1, res=1.0):
    """The reference's mapping loop (lidar_subscriber/src/ndt_rosbag_mapping_node.cpp:42-161) on the oracle:
    downsample_cloud -> perform_registration(prev, cur, guess = previous transform) -> pose = pose * transform ->
    update_global_map (transform, append, VoxelGrid map_voxel).  Returns the per-step records and the global map."""
    import oracle
    ndt = oracle.NormalDistributionsTransform()
    ndt.setResolution(res); ndt.setStepSize(step); ndt.setTransformationEpsilon(eps); ndt.setMaximumIterations(max_iter)
    ndt.setNeighborhoodSearchMethod(oracle.DIRECT7)
    pose = np.eye(4, dtype=np.float32)
    pres = np.eye(4, dtype=np.float32)
    prev = None
    gmap = np.zeros((0, 3), dtype=np.float32)
    steps = []
    for cloud in scans:
        filtered = oracle.voxelgrid_downsample(cloud, voxel_leaf)
        rec = {"n_filtered": len(filtered), "transform": np.eye(4, dtype=np.float32), "converged": False, "iterations": 0,
               "n_evaluations": 0, "fitness": 0.0}
        if prev is not None:
            ndt.setInputTarget(prev)
            ndt.setInputSource(filtered)
            ndt.align(pres)
            r = ndt.result()
            rec.update(converged=bool(r["converged"]), iterations=r["iterations"], n_evaluations=r["n_evaluations"],
                       fitness=ndt.getFitnessScore())
            if r["converged"]:
                rec["transform"] = r["final"].astype(np.float32)
            pres = rec["transform"]
            pose = (pose @ rec["transform"]).astype(np.float32)
        rec["pose"] = pose.copy()
        moved = oracle.transform_points(pose, filtered)[:, :3]
        gmap = oracle.voxelgrid_downsample(np.concatenate([gmap, moved]).astype(np.float32), map_voxel)
        rec["n_map"] = len(gmap)
        steps.append(rec)
        prev = filtered
    return steps, gmap


def as_xyzw_host(points):
    """(n,3) -> contiguous (n,4) float32 host array with w = 1 (the 16-byte PointXYZ record the C ABI takes by pointer)."""
    p = np.ones((len(points), 4), dtype=np.float32)
    p[:, :3] = np.asarray(points, dtype=np.float32)[:, :3]
    return np.ascontiguousarray(p)
