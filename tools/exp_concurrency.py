#!/usr/bin/env python3
"""Experiment: throughput of C independent align chains in flight at once (each handle has its own stream)."""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def main():
    import torch, bench
    import toyslam_b200 as nb
    ap = argparse.ArgumentParser()
    ap.add_argument("--method", default="DIRECT7")
    ap.add_argument("--map-points", type=int, default=1_000_000)
    ap.add_argument("--map-scans", type=int, default=31)
    ap.add_argument("--azimuth-steps", type=int, default=1875)
    ap.add_argument("--cache", default=None)
    ap.add_argument("--replicas", type=int, default=64)
    ap.add_argument("--steps", type=int, default=512)
    ap.add_argument("--conc", type=int, nargs="+", default=[1, 2, 4, 8])
    args = ap.parse_args()
    R = args.replicas
    w = bench.make_workload(args, 0, R)
    hs = []
    for r in range(R):
        ndt = nb.NormalDistributionsTransform()
        ndt.setNeighborhoodSearchMethod(bench.METHODS[args.method])
        ndt.setInputTarget(w["target"])
        ndt.setInputSource(w["sources"][r])
        ndt.align_async(); ndt.sync()
        hs.append(ndt)
    dev = torch.device("cuda", 0)
    streams = [torch.cuda.ExternalStream(h.stream_ptr(), device=dev) for h in hs]
    for C in args.conc:
        torch.cuda.synchronize()
        last = [None] * C
        t0 = time.perf_counter()
        for i in range(args.steps):
            h, st = hs[i % R], streams[i % R]
            c = i % C
            if last[c] is not None:
                st.wait_event(last[c])
            h.align_async()
            e = torch.cuda.Event()
            e.record(st)
            last[c] = e
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("concurrency %d: %.1f aligns/s (%.1f us per align, wall clock over %d aligns)" % (C, args.steps / dt, dt / args.steps * 1e6, args.steps), flush=True)

if __name__ == "__main__":
    main()
