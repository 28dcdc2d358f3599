"""Deterministic synthetic LiDAR workloads for bench.py and the large-size tests.

Not part of the product: the generator only manufactures inputs (torch is used as a vectorised
ray-caster; it runs on the GPU when one is present, on the CPU otherwise).

Scene (seed 20260101): ground plane z = 0, axis-aligned building boxes (footprints 10-40 m, heights
5-30 m, ~1 per 1500 m^2 over a 400 m x 400 m tile, a street corridor along x kept free) and vertical
cylinders (poles / trunks, r = 0.15-0.4 m).  A scan is a 64-beam spinning LiDAR at z = 1.8 m (elevation
-24.8 .. +2 deg, 1875 azimuth steps, 100 m range, N(0, 0.02 m) range noise).
"""
import math

import numpy as np
import torch

SEED_C2 = 20260101
SEED_C3 = 20260102


def _device():
    return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


class Scene:
    def __init__(self, seed=SEED_C2, tile=400.0, street_half_width=6.0):
        rng = np.random.default_rng(seed)
        half = tile / 2
        n_box = rng.poisson(tile * tile / 1500.0)
        boxes = []
        while len(boxes) < n_box:
            w, d = rng.uniform(10, 40, size=2)
            cx, cy = rng.uniform(-half, half, size=2)
            if abs(cy) - d / 2 < street_half_width:
                continue
            boxes.append([cx - w / 2, cx + w / 2, cy - d / 2, cy + d / 2, rng.uniform(5, 30)])
        n_cyl = rng.poisson(tile * tile / 800.0)
        cyls = []
        while len(cyls) < n_cyl:
            cx, cy = rng.uniform(-half, half, size=2)
            if abs(cy) < 2.5:
                continue
            cyls.append([cx, cy, rng.uniform(0.15, 0.4), rng.uniform(3, 12)])
        self.boxes = np.asarray(boxes, dtype=np.float64)
        self.cyls = np.asarray(cyls, dtype=np.float64)
        self.tile = tile

    def scan(self, pose_xy_yaw, seed, beams=64, azimuth_steps=1875, max_range=100.0, noise=0.02, device=None):
        """Ray-cast one revolution.  Returns (points in the SENSOR frame, points in the WORLD frame), float32 (n,3)."""
        dev = device or _device()
        f = torch.float64
        px, py, yaw = pose_xy_yaw
        elev = torch.deg2rad(torch.linspace(-24.8, 2.0, beams, dtype=f, device=dev))
        azim = torch.arange(azimuth_steps, dtype=f, device=dev) * (2 * math.pi / azimuth_steps)
        ce, se = torch.cos(elev)[:, None], torch.sin(elev)[:, None]
        ds = torch.stack([(ce * torch.cos(azim)[None, :]).reshape(-1), (ce * torch.sin(azim)[None, :]).reshape(-1),
                          (se * torch.ones_like(azim)[None, :]).reshape(-1)], dim=1)  # sensor-frame directions
        c, s = math.cos(yaw), math.sin(yaw)
        Rz = torch.tensor([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=f, device=dev)
        dw = ds @ Rz.T
        o = torch.tensor([px, py, 1.8], dtype=f, device=dev)
        inf = torch.tensor(float("inf"), dtype=f, device=dev)
        # ground
        t = torch.where(dw[:, 2] < -1e-9, -o[2] / dw[:, 2], inf)
        # boxes (slab method), z in [0, h]
        if len(self.boxes):
            B = torch.as_tensor(self.boxes, dtype=f, device=dev)
            lo = torch.stack([B[:, 0], B[:, 2], torch.zeros_like(B[:, 0])], dim=1)
            hi = torch.stack([B[:, 1], B[:, 3], B[:, 4]], dim=1)
            inv = 1.0 / torch.where(dw.abs() < 1e-12, torch.full_like(dw, 1e-12), dw)
            t0 = (lo[None, :, :] - o[None, None, :]) * inv[:, None, :]
            t1 = (hi[None, :, :] - o[None, None, :]) * inv[:, None, :]
            tn = torch.minimum(t0, t1).amax(dim=2)
            tf = torch.maximum(t0, t1).amin(dim=2)
            hit = (tn <= tf) & (tf > 0) & (tn > 0)
            t = torch.minimum(t, torch.where(hit, tn, inf).amin(dim=1))
        # vertical cylinders
        if len(self.cyls):
            Cc = torch.as_tensor(self.cyls, dtype=f, device=dev)
            ox, oy = o[0] - Cc[:, 0], o[1] - Cc[:, 1]
            a = (dw[:, 0] ** 2 + dw[:, 1] ** 2)[:, None]
            b = 2 * (dw[:, 0:1] * ox[None, :] + dw[:, 1:2] * oy[None, :])
            cc = (ox ** 2 + oy ** 2 - Cc[:, 2] ** 2)[None, :]
            disc = b * b - 4 * a * cc
            tc = (-b - torch.sqrt(disc.clamp_min(0))) / (2 * a)
            zc = o[2] + tc * dw[:, 2:3]
            ok = (disc > 0) & (tc > 0) & (zc >= 0) & (zc <= Cc[None, :, 3])
            t = torch.minimum(t, torch.where(ok, tc, inf).amin(dim=1))
        g = torch.Generator(device="cpu").manual_seed(int(seed))
        t = t + (noise * torch.randn(t.shape, generator=g, dtype=f)).to(dev)
        keep = torch.isfinite(t) & (t < max_range) & (t > 0.5)
        ps = (ds * t[:, None])[keep]
        pw = (o[None, :] + dw * t[:, None])[keep]
        return ps.to(torch.float32).cpu().numpy(), pw.to(torch.float32).cpu().numpy()


def voxel_thin(points, leaf, max_points, seed):
    """One point per `leaf`-sized cell (first in input order), then a seeded random subset of max_points."""
    q = np.floor(points / leaf).astype(np.int64)
    q -= q.min(axis=0)
    dims = q.max(axis=0) + 1
    key = (q[:, 2] * dims[1] + q[:, 1]) * dims[0] + q[:, 0]
    _, first = np.unique(key, return_index=True)
    first.sort()
    pts = points[first]
    if len(pts) > max_points:
        sel = np.random.default_rng(seed).choice(len(pts), size=max_points, replace=False)
        sel.sort()
        pts = pts[sel]
    return np.ascontiguousarray(pts, dtype=np.float32)


def pose_matrix(p):
    x, y, z, r, pi, ya = p
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(pi), np.sin(pi), np.cos(ya), np.sin(ya)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rx @ Ry @ Rz
    T[:3, 3] = [x, y, z]
    return T


def config2_scan(scene, scan_seed, seed=SEED_C2, azimuth_steps=1875, perturb_seed=None):
    """One perturbed scan (source cloud) of the c2 scene and the transform that maps it onto the map.
    perturb_seed: use one common perturbation (about the origin) for several scans that are merged into one source."""
    rng = np.random.default_rng(seed + 7919 * (scan_seed + 1))
    sx = float(rng.uniform(-25.0, 25.0))
    _, pw = scene.scan((sx, float(rng.uniform(-1.0, 1.0)), float(rng.uniform(-0.05, 0.05))), seed=seed + 5000 + scan_seed,
                       azimuth_steps=azimuth_steps)
    if perturb_seed is not None:
        rng = np.random.default_rng(seed + 104729 * (perturb_seed + 1))
        sx = 0.0
    d2r = np.pi / 180.0
    pert = [rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5),
            rng.uniform(-0.5, 0.5) * d2r, rng.uniform(-0.5, 0.5) * d2r, rng.uniform(-2.0, 2.0) * d2r]
    centre = np.array([sx, 0.0, 0.0])  # perturb about the scan's own position so the correction stays a small pose
    T = pose_matrix(pert)
    inv = np.linalg.inv(T)
    local = pw.astype(np.float64) - centre
    src = local @ inv[:3, :3].T + inv[:3, 3] + centre
    truth = np.eye(4)
    truth[:3, :3] = T[:3, :3]
    truth[:3, 3] = T[:3, 3] + centre - T[:3, :3] @ centre
    return src.astype(np.float32), truth


def config2_map(map_points=1_000_000, n_map_scans=31, seed=SEED_C2, azimuth_steps=1875, thin_leaf=0.1):
    """The c2 target map: union of scans taken every 2 m along the street, voxel-thinned to ~map_points."""
    scene = Scene(seed)
    xs = np.linspace(-30.0, 30.0, n_map_scans)
    world = []
    for i, x in enumerate(xs):
        _, pw = scene.scan((float(x), 0.0, 0.0), seed=seed + 1000 + i, azimuth_steps=azimuth_steps)
        world.append(pw)
    return scene, voxel_thin(np.concatenate(world), thin_leaf, map_points, seed + 1)


def config2(map_points=1_000_000, n_map_scans=31, scan_seed=0, seed=SEED_C2, azimuth_steps=1875, offset=(0.0, 0.0, 0.0)):
    """BASELINE.json configs[1]: one 64-beam scan (~120 k points) against a ~1 M-point target map.

    Returns dict(target (M,3) f32, source (N,3) f32, truth 4x4 f64) where `truth` maps the source onto the
    map (what align() with an identity guess should recover: translation U(-0.5,0.5) m per axis,
    yaw U(-2,2) deg, roll/pitch U(-0.5,0.5) deg).  `scan_seed` selects which scan / perturbation
    (different seeds = independent scan pairs for the multi-GPU replicas)."""
    scene = Scene(seed)
    xs = np.linspace(-30.0, 30.0, n_map_scans)
    world = []
    for i, x in enumerate(xs):
        _, pw = scene.scan((float(x), 0.0, 0.0), seed=seed + 1000 + i, azimuth_steps=azimuth_steps)
        world.append(pw)
    world = np.concatenate(world)
    target = voxel_thin(world, 0.1, map_points, seed + 1)
    rng = np.random.default_rng(seed + 7919 * (scan_seed + 1))
    sx = float(rng.uniform(-25.0, 25.0))
    _, pw = scene.scan((sx, float(rng.uniform(-1.0, 1.0)), float(rng.uniform(-0.05, 0.05))), seed=seed + 5000 + scan_seed,
                       azimuth_steps=azimuth_steps)
    d2r = np.pi / 180.0
    pert = [rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5),
            rng.uniform(-0.5, 0.5) * d2r, rng.uniform(-0.5, 0.5) * d2r, rng.uniform(-2.0, 2.0) * d2r]
    # perturb about the scan's own position so the needed correction stays a small pose
    centre = np.array([sx, 0.0, 0.0])
    T = pose_matrix(pert)
    inv = np.linalg.inv(T)
    local = pw.astype(np.float64) - centre
    src = local @ inv[:3, :3].T + inv[:3, 3] + centre
    truth = np.eye(4)
    truth[:3, :3] = T[:3, :3]
    truth[:3, 3] = T[:3, 3] + centre - T[:3, :3] @ centre
    off = np.asarray(offset, dtype=np.float64)
    return {"target": (target.astype(np.float64) + off).astype(np.float32), "source": src.astype(np.float32),
            "truth": truth, "offset": off}


def voxel_centroid_downsample(points, leaf):
    """pcl::VoxelGrid-style centroid downsample (one fp32 centroid per occupied leaf-sized cell, cells in ascending
    index order).  Input preparation only: the generator does not claim bit parity with PCL here."""
    p = np.asarray(points, dtype=np.float32)
    inv = np.float32(1.0) / np.float32(leaf)
    q = np.floor(p * inv).astype(np.int64)
    q -= q.min(axis=0)
    dims = q.max(axis=0) + 1
    key = (q[:, 2] * dims[1] + q[:, 1]) * dims[0] + q[:, 0]
    uniq, inv_idx, cnt = np.unique(key, return_inverse=True, return_counts=True)
    out = np.zeros((len(uniq), 3), dtype=np.float64)
    for a in range(3):
        out[:, a] = np.bincount(inv_idx, weights=p[:, a].astype(np.float64), minlength=len(uniq))
    out /= cnt[:, None]
    return np.ascontiguousarray(out, dtype=np.float32)


def config3_sequence(n_scans, seed=SEED_C3, azimuth_steps=1875, leaf=0.3):
    """BASELINE.json configs[2]: consecutive scans of a simulated drive through the c2 scene generator (seed 20260102):
    10 Hz, 5-15 m/s, yaw rate <= 20 deg/s, each scan in the SENSOR frame, downsampled with a 0.3 m voxel grid (the
    mapping node's default, ndt_rosbag_mapping_node.cpp:88).  Returns (scans, poses) with poses[k] = (x, y, yaw)."""
    scene = Scene(seed)
    rng = np.random.default_rng(seed + 17)
    half = scene.tile / 2 - 40.0
    x, y, yaw, v, direction = -half, 0.0, 0.0, 10.0, 1.0
    scans, poses = [], []
    for k in range(n_scans):
        ps, _ = scene.scan((x, y, yaw), seed=seed + 9000 + k, azimuth_steps=azimuth_steps)
        scans.append(voxel_centroid_downsample(ps, leaf))
        poses.append((x, y, yaw))
        v = float(np.clip(v + rng.uniform(-0.5, 0.5), 5.0, 15.0))
        yaw_rate = float(np.deg2rad(rng.uniform(-20.0, 20.0)))
        # keep the vehicle on the street corridor: steer back towards y = 0 and the street axis
        heading = 0.0 if direction > 0 else math.pi
        err = ((heading - yaw + math.pi) % (2 * math.pi)) - math.pi
        yaw += 0.1 * float(np.clip(0.5 * yaw_rate + 2.0 * err - 0.3 * y * direction, -np.deg2rad(20.0), np.deg2rad(20.0)))
        x += 0.1 * v * math.cos(yaw)
        y += 0.1 * v * math.sin(yaw)
        if abs(x) > half:
            direction = -direction
    return scans, poses


def relative_pose_matrix(pose_a, pose_b):
    """4x4 transform taking points of scan b's sensor frame into scan a's sensor frame (planar poses x, y, yaw)."""
    def mat(p):
        c, s = math.cos(p[2]), math.sin(p[2])
        T = np.eye(4)
        T[:2, :2] = [[c, -s], [s, c]]
        T[0, 3], T[1, 3] = p[0], p[1]
        return T
    return np.linalg.inv(mat(pose_a)) @ mat(pose_b)
