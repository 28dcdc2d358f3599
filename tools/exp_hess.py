#!/usr/bin/env python3
"""Experiment: kernel time of one fp32 derivative evaluation vs one fp64 Hessian-only pass on the c2 pair."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import toyslam_b200 as nb
class A: pass
a = A(); a.map_points = 1000000; a.map_scans = 31; a.azimuth_steps = 1875; a.cache = '/tmp/wl'; a.method = 'DIRECT7'
w = bench.make_workload(a, 0, 1)
for mode in (False, True):
    ndt = nb.NormalDistributionsTransform()
    ndt.set_throughput_mode(mode)
    ndt.setInputTarget(w['target']); ndt.setInputSource(w['source'])
    p = np.array([0.1, 0.05, 0.0, 0.001, 0.0, 0.01])
    for name, fn in (("eval+hessian", lambda: ndt.eval_derivatives(p, compute_hessian=True)), ("eval no hessian", lambda: ndt.eval_derivatives(p, compute_hessian=False)),
                     ("hessian-only fp64", lambda: ndt.eval_hessian(p))):
        fn(); ts = []
        for _ in range(20):
            fn(); ts.append(ndt.last_align_ms() * 1e3)
        print("throughput_shape=%s %-18s kernel %.1f us (median of 20, events)" % (mode, name, float(np.median(ts))))
