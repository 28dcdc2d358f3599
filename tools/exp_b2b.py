#!/usr/bin/env python3
"""Experiment: steady-state cost of back-to-back align() launches on ONE handle (no events / waits in between)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
import toyslam_b200 as nb

class A: pass
a = A(); a.map_points = 1000000; a.map_scans = 31; a.azimuth_steps = 1875; a.cache = '/tmp/wl'; a.method = 'DIRECT7'
w = bench.make_workload(a, 0, 1)
ndt = nb.NormalDistributionsTransform()
ndt.setInputTarget(w['target']); ndt.setInputSource(w['source'])
ndt.align_async(); ndt.sync()
r = ndt.result()
print("evals", r["n_evaluations"], "hess", r["n_hessian_passes"], "single kernel ms (events)", ndt.last_align_ms())
st = torch.cuda.ExternalStream(ndt.stream_ptr())
for n in (1, 10, 100):
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(st)
    for _ in range(n):
        ndt.align_async()
    e1.record(st)
    torch.cuda.synchronize()
    print("n=%d: %.1f us per align (events), %.1f us per align (wall)" % (n, e0.elapsed_time(e1) * 1e3 / n, (time.perf_counter() - t0) * 1e6 / n))
# blocking align (launch + sync + result copy), the apps/align.cpp 10times protocol
t0 = time.perf_counter()
for _ in range(100):
    ndt.align_async(); ndt.sync()
print("blocking align_async+sync: %.1f us per align (wall)" % ((time.perf_counter() - t0) * 1e4))
p = np.zeros(6)
t0 = time.perf_counter()
for _ in range(100):
    ndt.eval_derivatives(p, compute_hessian=True)
print("eval_derivatives (1 evaluation, blocking): %.1f us (wall); kernel ms (events) %.4f" % ((time.perf_counter() - t0) * 1e4, ndt.last_align_ms()))
