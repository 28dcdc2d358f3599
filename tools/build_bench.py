#!/usr/bin/env python3
"""Target-map build throughput (VoxelGridCovariance::applyFilter on the device), BASELINE configs[4] style sweep:
M synthetic surface points, resolution 0.5 / 1.0 / 2.0.  Prints one JSON line per (M, resolution)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def surface_points(m, seed, extent=600.0, height=30.0, device="cuda"):
    """City-like surface samples: ground, vertical walls of a building grid, roofs; float32 (m,4) on the device."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    kind = torch.randint(0, 4, (m,), generator=g, device=device)
    u = (torch.rand(m, generator=g, device=device) - 0.5) * extent
    v = (torch.rand(m, generator=g, device=device) - 0.5) * extent
    h = torch.rand(m, generator=g, device=device) * height
    block = 40.0
    x = torch.where(kind == 1, torch.round(u / block) * block, u)
    y = torch.where(kind == 2, torch.round(v / block) * block, v)
    z = torch.where(kind == 0, torch.zeros_like(h), torch.where(kind == 3, torch.floor(h / 10.0) * 10.0, h))
    pts = torch.stack([x, y, z, torch.ones_like(x)], dim=1)
    pts[:, :3] += 0.02 * torch.randn(m, 3, generator=g, device=device)
    return pts.contiguous()


def main():
    import torch
    import toyslam_b200 as nb
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, nargs="+", default=[10_000_000, 50_000_000, 100_000_000])
    ap.add_argument("--res", type=float, nargs="+", default=[0.5, 1.0, 2.0])
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    peak = 6454.9
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    for m in args.points:
        pts = surface_points(m, 20260104)
        torch.cuda.synchronize()
        for res in args.res:
            ndt = nb.NormalDistributionsTransform()
            ndt.setResolution(res)
            ndt.set_target_device_view(pts.data_ptr(), m)  # warm-up (allocations); the map references the caller's cloud, as the reference does
            times = []
            for _ in range(args.reps):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                st = ndt.set_target_device_view(pts.data_ptr(), m)
                torch.cuda.synchronize()
                times.append(time.perf_counter() - t0)
            info = ndt.map_info()
            t = min(times)
            alg = 16.0 * m + 72.0 * info["n_voxels"]
            print(json.dumps({"metric": "map_build_points_per_s", "points": m, "resolution": res, "status": st, "ms": t * 1e3,
                              "value": m / t, "voxels": info["n_voxels"], "valid": info["n_valid"], "launches": None,
                              "roofline": {"bound": "hbm", "achieved": alg / t / 1e9, "peak": peak, "frac": alg / t / 1e9 / peak,
                                           "algorithmic_bytes": alg, "formula": "16*M + 72*V"}}))
            del ndt
        del pts
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
