#!/usr/bin/env python3
"""Experiment: what bounds the batched end-to-end path (H2D source + solve + D2H aligned cloud + result)?"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
import toyslam_b200 as nb

class A: pass
a = A(); a.map_points = 1000000; a.map_scans = 31; a.azimuth_steps = 1875; a.cache = '/tmp/wl'; a.method = 'DIRECT7'
R = 64
w = bench.make_workload(a, 0, R)
hs, src_h, out_h, ns = [], [], [], []
for r in range(R):
    n = nb.NormalDistributionsTransform()
    n.setInputTarget(w['target'])
    s = w['sources'][r]
    sh = torch.ones((len(s), 4), dtype=torch.float32).pin_memory(); sh[:, :3] = torch.from_numpy(s)
    n.set_source_raw(sh.data_ptr(), len(s), 16)
    hs.append(n); src_h.append(sh); ns.append(len(s)); out_h.append(torch.empty((len(s), 4), dtype=torch.float32).pin_memory())
batch = nb.Batch(hs)
outs = [o.data_ptr() for o in out_h]
def run(label, fn, reps=4):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%-60s %8.1f aligns/s" % (label, reps * R / dt), flush=True)
def full():
    for i in range(R): hs[i].set_source_raw(src_h[i].data_ptr(), ns[i], 16)
    batch.align(None, outs, 16)
def no_out():
    for i in range(R): hs[i].set_source_raw(src_h[i].data_ptr(), ns[i], 16)
    batch.align(None, None, 16)
def no_h2d():
    batch.align(None, outs, 16)
def compute_only():
    batch.align(None, None, 16)
def full_async():   # enqueue only; one sync at the very end of the timed region
    for i in range(R): hs[i].set_source_raw(src_h[i].data_ptr(), ns[i], 16)
    batch.align_async(None, outs, 16)
def interleaved():  # upload + launch handle by handle instead of all uploads first
    for i in range(R):
        hs[i].set_source_raw(src_h[i].data_ptr(), ns[i], 16)
        hs[i].set_throughput_mode(True)
        hs[i].align_raw_async_out(outs[i]) if hasattr(hs[i], "align_raw_async_out") else None
run("H2D + solve + D2H cloud + result (bench e2e)", full)
run("H2D + solve + result (no output cloud)", no_out)
run("solve + D2H cloud + result (no H2D)", no_h2d)
run("solve + result only", compute_only)
run("H2D + solve + D2H cloud, batches enqueued back to back (no wait)", full_async)
for h in hs: h.sync()
