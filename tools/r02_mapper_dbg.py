import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
import numpy as np, torch
import toyslam_b200 as nb, workloads
scans, _ = workloads.config3_sequence(60, seed=workloads.SEED_C3, azimuth_steps=1875, leaf=0.02)
hosts = []
for sc in scans:
    hb = torch.ones((len(sc), 4), dtype=torch.float32).pin_memory(); hb[:, :3] = torch.from_numpy(sc); hosts.append(hb)
for rep in range(2):
    m = nb.Mapper(device=0)
    ts = []
    for k in range(60):
        t0 = time.perf_counter(); m.push_scan_raw(hosts[k].data_ptr(), len(scans[k]), 16); ts.append((time.perf_counter() - t0) * 1e3)
    print("rep", rep, "slow steps:", [(i, round(t, 2)) for i, t in enumerate(ts) if t > 2.0], "median", round(float(np.median(ts)), 3))
    del m
