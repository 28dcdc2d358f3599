#!/usr/bin/env python3
"""bench.py — NDT registration hot path on B200 (BASELINE.json metric: aligns/s and source-point·iterations/s,
achieved HBM GB/s vs peak), next to the reference's OpenMP CPU path (the oracle port) on the same box.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference CPU path (oracle port, all host threads)

Workload (config.workload): BASELINE.json configs[1] — a synthetic 64-beam LiDAR scan (~120 k points) against a
1 M-point target map, resolution 1.0, DIRECT7, class defaults, identity guess.  A "step" is one align() of that
pair (the `10times` loop of ndt_omp/apps/align.cpp:25-27: target map built once outside the timer).
N > 1: one process per GPU, each rank aligns its own independent scan against its own copy of the map
(independent scan pairs are the unit that shards; no data-path collective) -> weak scaling.

`value`  : aligns/s with inputs resident in HBM: 64 independent (scan, map) pairs per GPU (their touched bytes exceed the
           L2: inputs larger than L2) aligned through ndtb200_align_batch_async — every step is one launch of the
           persistent solve kernel, up to four of them co-resident per SM; CUDA events on a master stream bracket the
           whole timed region.  `latency` is the same align with ONE solve in flight at a time.
`e2e`    : the same through the C ABI with HOST buffers: per step H2D of the source cloud from pinned memory, the
           solve, D2H of the transformed cloud and of the result block (batched; `e2e.single_call` = one blocking
           ndtb200_align at a time).
Other workloads: --workload c3 (batched odometry incl. map builds), c4 (source-sharded scan-to-map), c5 (map-build
sweep), mapper (the mapping-node loop as a device-resident pipeline).
"""
import argparse
import os as _os
_os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")  # no lazy kernel-load stalls inside timed regions
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # one hardware queue per stream (up to 32): independent solves do not serialise
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METHODS = {"DIRECT1": 3, "DIRECT7": 2, "DIRECT26": 1}
KPROBE = {"DIRECT1": 1, "DIRECT7": 7, "DIRECT26": 26}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(names, f[4:8]):
                if val.lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus, bind=True):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl" if torch.cuda.is_available() else "gloo",
                                device_id=torch.device("cuda", local) if torch.cuda.is_available() else None)
        if bind and torch.cuda.is_available() and os.environ.get("NDTB200_NO_BIND") is None:
            bind_rank_to_gpu(local, world)  # before any pinned allocation
    return rank, world, local


TOPOLOGY = {}


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_rank_to_gpu(local, world):
    """Rank -> core binding for the multi-GPU runs (end-to-end efficiency at 8 GPUs).  One process per GPU:
    the process (and with it the first-touch placement of its pinned staging buffers) is bound to the cores of the NUMA
    node its GPU hangs off (sysfs numa_node of the GPU's PCI function); the cores of a node are divided evenly between
    the ranks that share it.  Without NUMA information the allowed cores are simply divided between the local ranks.
    Best effort: any failure leaves the affinity as it was.  The outcome is reported as `topology` in the JSON line."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        TOPOLOGY.update({"cores_allowed": len(allowed), "local_world": local_world, "numa_node": None, "bound_cores": len(allowed)})
        if local_world <= 1:
            return
        nodes = {}
        try:
            import pynvml
            pynvml.nvmlInit()
            for g in range(local_world):
                bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(g)).busId
                bus = bus.decode() if isinstance(bus, bytes) else bus
                path = "/sys/bus/pci/devices/%s/numa_node" % bus.lower()[-12:]
                nodes[g] = int(open(path).read().strip()) if os.path.exists(path) else -1
        except Exception:
            nodes = {}
        mine = nodes.get(local, -1)
        pool, sharers = allowed, list(range(local_world))
        if mine >= 0:
            node_cpus = _parse_cpulist(open("/sys/devices/system/node/node%d/cpulist" % mine).read()) & set(allowed)
            if node_cpus:
                pool = sorted(node_cpus)
                sharers = [g for g in range(local_world) if nodes.get(g, -1) == mine]
                TOPOLOGY["numa_node"] = mine
        k = sharers.index(local) if local in sharers else 0
        per = max(1, len(pool) // max(1, len(sharers)))
        cores = pool[k * per:(k + 1) * per] or pool
        os.sched_setaffinity(0, cores)
        TOPOLOGY.update({"bound_cores": len(cores), "gpu_numa_nodes": [nodes.get(g, -1) for g in range(local_world)] if nodes else None})
    except Exception as e:  # pragma: no cover - best effort
        TOPOLOGY["bind_error"] = str(e)[:120]


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def max_over_ranks(x, world, device):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, world, device):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def host_threads():
    """Host cores this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers, so the CPU arms set the
    oracle's thread count explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def hbm_peak():
    peak, src = 6650.0, "fallback (B200_PROFILING.md)"
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        try:
            peak = float(json.load(open(pk))["hbm_gbs"])
            src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return peak, src


def make_workload(args, rank, n_scans=1):
    """The c2 map and `n_scans` independent scans of this rank (scan seeds rank*1000 + i).  --cache DIR keeps the
    generated arrays so a profiler run (ncu) of the same command does not see the generator's kernels."""
    import workloads
    path = None
    if args.cache:
        os.makedirs(args.cache, exist_ok=True)
        path = os.path.join(args.cache, "c2_%d_%d_%d_r%d_n%d.npz" % (args.map_points, args.map_scans, args.azimuth_steps, rank, n_scans))
        if os.path.exists(path):
            d = np.load(path)
            srcs = [d["source_%d" % i] for i in range(n_scans)]
            return {"target": d["target"], "source": srcs[0], "sources": srcs}
    scene, target = workloads.config2_map(map_points=args.map_points, n_map_scans=args.map_scans, azimuth_steps=args.azimuth_steps)
    srcs = [workloads.config2_scan(scene, rank * 1000 + i, azimuth_steps=args.azimuth_steps)[0] for i in range(n_scans)]
    if path:
        np.savez(path, target=target, **{"source_%d" % i: s for i, s in enumerate(srcs)})
    return {"target": target, "source": srcs[0], "sources": srcs}


def workload_name(args, n_src, n_tgt):
    return ("c2: synthetic 64-beam LiDAR scan (%d pts) vs %d-pt target map, resolution 1.0, %s, class defaults, "
            "identity guess; every align() with the target map prebuilt (as the apps/align.cpp 10times loop)" %
            (n_src, n_tgt, args.method))


def cpu_baseline(w, args, max_seconds=25.0, max_pairs=16):
    """The oracle port (C++/OpenMP restatement of the reference) on the host cores: the first pairs of the same
    workload, same parameters, all host threads.  Returns (summary, result of pair 0)."""
    import oracle
    ref = oracle.NormalDistributionsTransform()
    nthr = host_threads()
    ref.setNumThreads(nthr)
    ref.setNeighborhoodSearchMethod({"DIRECT1": oracle.DIRECT1, "DIRECT7": oracle.DIRECT7, "DIRECT26": oracle.DIRECT26}[args.method])
    t0 = time.perf_counter()
    ref.setInputTarget(w["target"])
    t_build = time.perf_counter() - t0
    ref.setInputSource(w["sources"][0])
    ref.align()  # warm-up
    first = ref.result()
    times, pt_evals = [], 0
    t_start = time.perf_counter()
    for s in w["sources"][:max_pairs]:
        if (time.perf_counter() - t_start) > max_seconds:
            break
        ref.setInputSource(s)
        t0 = time.perf_counter()
        ref.align()
        times.append(time.perf_counter() - t0)
        pt_evals += len(s) * ref.result()["n_evaluations"]
    tot = float(np.sum(times))
    return {"value": len(times) / tot, "unit": "aligns/s", "cores": nthr, "kind": "port",
            "sample": "one full align of each of the first %d pairs of the same workload (oracle C++/OpenMP port of ndt_omp; "
                      "the reference itself needs PCL/Eigen and cannot be built here)" % len(times),
            "ms_per_align": tot / len(times) * 1e3, "map_build_ms": t_build * 1e3,
            "src_pt_iters_per_s": pt_evals / tot}, first


C1_METHODS = [("KDTREE", 0), ("DIRECT7", 2), ("DIRECT1", 3)]       # the order of ndt_omp/apps/align.cpp:89-93


def c1_clouds():
    """BASELINE configs[0]: the bundled scan pair after the app's 0.1 m pcl::VoxelGrid (apps/align.cpp:57-69) — the committed
    fixture tests/golden/pair_ds0p1.npz (made from ndt_omp/data/*.pcd by tests/golden/make_fixtures.py), plus the raw pair."""
    g = os.path.join(ROOT, "tests", "golden")
    d = np.load(os.path.join(g, "pair_ds0p1.npz"))
    raw = np.load(os.path.join(g, "pair_raw.npz"))
    with open(os.path.join(g, "golden.json")) as f:
        gold = json.load(f)
    return d["target"], d["source"], raw["target"], raw["source"], gold


def c1_protocol(make, tgt, src, k):
    """apps/align.cpp:15-33 for one registration object: setInputTarget / setInputSource, one timed align (`single`), k more
    (`10times` for k = 10), getFitnessScore.  Wall clock around blocking calls, like the app's ros::WallTime."""
    reg = make()
    t0 = time.perf_counter()
    reg.setInputTarget(tgt)
    t_build = time.perf_counter() - t0
    reg.setInputSource(src)
    t1 = time.perf_counter()
    reg.align()
    t2 = time.perf_counter()
    for _ in range(k):
        reg.align()
    t3 = time.perf_counter()
    fit = reg.getFitnessScore()
    r = reg.result()
    return {"single_ms": (t2 - t1) * 1e3, "ktimes_ms": (t3 - t2) * 1e3, "k": k, "fitness": fit, "set_input_target_ms": t_build * 1e3,
            "iterations": r["iterations"], "evaluations": r["n_evaluations"]}


def c1_oracle(k, threads):
    import oracle
    tgt, src, _, _, gold = c1_clouds()
    out = {}
    for name, method in C1_METHODS:
        def make():
            o = oracle.NormalDistributionsTransform()
            o.setNumThreads(threads)
            o.setResolution(1.0)
            o.setNeighborhoodSearchMethod(method)
            return o
        out[name] = c1_protocol(make, tgt, src, k)
    return out


def run_c1_reference(args):
    k = max(1, min(args.steps, 10))
    nthr = host_threads()
    res = c1_oracle(k, nthr)
    d7 = res["DIRECT7"]
    val = k / (d7["ktimes_ms"] * 1e-3)
    print(json.dumps({"impl": "reference", "metric": "ndt_aligns_per_s", "workload": "c1", "value": val, "unit": "aligns/s", "n_gpus": args.gpus,
                      "steps": k, "warmup": 1, "ms_per_step": d7["ktimes_ms"] / k, "higher_is_better": True, "scaling": "weak", "vs_baseline": val / 29.1,
                      "dtype": "f32", "data": "bundled ndt_omp/data scan pair (committed fixture)",
                      "config": {"workload": "c1: apps/align.cpp protocol on the bundled pair, DIRECT7 (headline) + KDTREE + DIRECT1, oracle port, %d threads" % nthr},
                      "cpu_baseline": {"value": val, "unit": "aligns/s", "cores": nthr, "kind": "port", "sample": "the whole c1 protocol"},
                      "e2e": {"value": val, "unit": "aligns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "methods": res}))
    return 0


def run_c1(args):
    """--workload c1: BASELINE configs[0], the one configuration the reference publishes timings for (ndt_omp/README.md:9-47,
    Core i7-6700K): `single`, `10times`, `fitness` per search method, through the blocking reference-style calls with HOST
    clouds (every align uploads nothing new but downloads the aligned cloud, like align(*aligned) fills its output)."""
    import torch
    import toyslam_b200 as nb
    rank, world, local = dist_setup(args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU path")
    torch.cuda.set_device(local)
    tgt, src, raw_t, raw_s, gold = c1_clouds()
    k = max(1, min(args.steps, 1000)) if args.steps_given else 10
    warm = nb.NormalDistributionsTransform(device=local)      # context / module load outside the protocol (the app's process start)
    # the app's own front end on the device: raw scans -> 0.1 m VoxelGrid (bit-identical to the fixture)
    t0 = time.perf_counter()
    ds_t, ds_s = warm.voxelgrid_filter(raw_t, 0.1), warm.voxelgrid_filter(raw_s, 0.1)
    vg_ms = (time.perf_counter() - t0) * 1e3
    same = bool(np.array_equal(ds_t, tgt) and np.array_equal(ds_s, src))
    # module load of every search method's kernel (CUDA loads a kernel lazily at its first launch: ~0.3 s once per process,
    # the GPU counterpart of the app's process start) — on a throw-away object and a fraction of the clouds
    for _, method in C1_METHODS:
        warm.setNeighborhoodSearchMethod(method)
        warm.setInputTarget(tgt[:4000])
        warm.setInputSource(src[:1000])
        warm.align()
    res = {}
    sampler = ClockSampler(local)
    sampler.start()
    launches = 0
    for name, method in C1_METHODS:
        holder = {}

        def make():
            o = nb.NormalDistributionsTransform(device=local)
            o.setResolution(1.0)
            o.setNeighborhoodSearchMethod(method)
            holder["o"] = o
            return o
        res[name] = c1_protocol(make, tgt, src, k)
        launches += holder["o"].launch_count()
        res[name]["fitness_matches_readme"] = bool("%.6f" % res[name]["fitness"] == "%.6f" % gold["fitness"][name])
        pub = gold["published_ms"]
        res[name]["readme_i7_6700K_ms"] = {"1thr": pub.get(name + "_1thr"), "8thr": pub.get(name + "_8thr")}
    clocks = sampler.stop()
    d7 = res["DIRECT7"]
    val = k / (d7["ktimes_ms"] * 1e-3)
    line = {"metric": "ndt_aligns_per_s", "workload": "c1", "value": val, "unit": "aligns/s", "n_gpus": 1, "steps": k, "warmup": 1,
            "ms_per_step": d7["ktimes_ms"] / k, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": val / 29.1, "vs_baseline_note": "README DIRECT7, 8 threads, i7-6700K: 343.336 ms per 10 aligns = 29.1 aligns/s (BASELINE.md)",
            "dtype": "f32", "data": "bundled ndt_omp/data scan pair (committed fixture tests/golden/pair_raw.npz / pair_ds0p1.npz)",
            "config": {"workload": "c1: ndt_omp/apps/align.cpp protocol (:15-33) on the bundled pair 251370668 -> 251371071 after the 0.1 m VoxelGrid "
                                   "(15772 / 15950 pts), resolution 1.0, class defaults; headline = DIRECT7 `%dtimes`" % k,
                       "timing": "host wall clock around the blocking calls (ndtb200_align with a host output cloud), one align in flight",
                       "device_voxelgrid_0p1_both_clouds_ms": vg_ms, "device_voxelgrid_equals_fixture": same},
            "methods": res, "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": val, "unit": "aligns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(len(src) * 16 + 416),
                    "note": "the protocol IS end to end: host clouds in (once), aligned host cloud + result out per align"}}
    if not args.no_cpu_baseline:
        nthr = host_threads()
        line["cpu_baseline"] = {"kind": "port", "unit": "aligns/s", "cores": nthr, "sample": "the same protocol on the oracle port, 1 thread and all threads",
                                "value": None}
        one, allt = c1_oracle(10, 1), c1_oracle(10, nthr)
        line["cpu_baseline"]["value"] = 10.0 / (allt["DIRECT7"]["ktimes_ms"] * 1e-3)
        line["cpu_baseline"]["methods_1_thread"] = one
        line["cpu_baseline"]["methods_all_threads"] = allt
    print(json.dumps(line))
    return 0


def run_reference(args):
    """The reference's CPU path (oracle port: the reference itself needs PCL / Eigen / FLANN, not installed) on this box's
    host cores, same workload / metric / unit.  A step of the GPU arm is a batch of `--replicas` aligns; here a step is a
    bounded sample of that batch — ONE full align of one pair of the batch — so the run ends within minutes; the value is
    aligns/s either way."""
    rank, world, local = dist_setup(args.gpus, bind=False)  # the CPU arm keeps all host cores
    if rank != 0:
        return 0
    import oracle
    if args.workload == "c1":
        return run_c1_reference(args)
    steps = min(args.steps, 24)          # bounded sample
    warm = min(args.warmup, 2)
    w = make_workload(args, 0, min(args.replicas, steps))
    srcs = w["sources"]
    nthr = host_threads()
    ref = oracle.NormalDistributionsTransform()
    ref.setNumThreads(nthr)
    ref.setNeighborhoodSearchMethod({"DIRECT1": oracle.DIRECT1, "DIRECT7": oracle.DIRECT7, "DIRECT26": oracle.DIRECT26}[args.method])
    ref.setInputTarget(w["target"])
    for i in range(warm):
        ref.setInputSource(srcs[i % len(srcs)])
        ref.align()
    dt, pt_evals, evals = 0.0, 0, []
    for i in range(steps):
        ref.setInputSource(srcs[i % len(srcs)])      # outside the timer, like the HBM-resident inputs of the GPU arm
        t0 = time.perf_counter()
        ref.align()
        dt += time.perf_counter() - t0
        r = ref.result()
        pt_evals += len(srcs[i % len(srcs)]) * r["n_evaluations"]
        evals.append(r["n_evaluations"])
    val = steps / dt
    line = {"impl": "reference", "metric": "ndt_aligns_per_s", "value": val, "unit": "aligns/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args, len(w["source"]), len(w["target"])),
                       "note": "reference CPU path = oracle port (C++/OpenMP restatement), %d host threads (set explicitly); "
                               "a step here = one align (a bounded sample of the GPU arm's %d-align batch step)" % (nthr, args.replicas)},
            "cpu_baseline": {"value": val, "unit": "aligns/s", "cores": nthr, "kind": "port",
                             "sample": "one full align of each of %d pairs of the c2 workload" % steps},
            "e2e": {"value": val, "unit": "aligns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "src_pt_iters_per_s": pt_evals / dt, "evaluations_per_align": float(np.mean(evals))}
    print(json.dumps(line))
    return 0


def run_b200(args):
    import torch
    import toyslam_b200 as nb
    rank, world, local = dist_setup(args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    # R independent (scan, map) pairs per GPU, cycled: every step aligns a different pair whose buffers were
    # last touched R steps ago.  With ~4 MB touched per align (ncu dram bytes) R = 64 pairs = 256 MB > the 126 MB L2,
    # so inputs are L2-cold without evicting the kernel code (which a device-wide L2 flush would also do).
    R = args.replicas
    w = make_workload(args, rank, R)
    n_tgt = len(w["target"])
    tgt_host = torch.ones((n_tgt, 4), dtype=torch.float32).pin_memory()
    tgt_host[:, :3] = torch.from_numpy(w["target"])
    handles, src_hosts, out_hosts, n_srcs = [], [], [], []
    t0 = time.perf_counter()
    for r in range(R):
        ndt = nb.NormalDistributionsTransform(device=local)
        ndt.setNeighborhoodSearchMethod(METHODS[args.method])
        ndt.set_target_raw(tgt_host.data_ptr(), n_tgt, 16)   # each handle owns its own copy of the map in HBM
        s = w["sources"][r]
        sh = torch.ones((len(s), 4), dtype=torch.float32).pin_memory()
        sh[:, :3] = torch.from_numpy(s)
        ndt.set_source_raw(sh.data_ptr(), len(s), 16)          # inputs resident in HBM before the timed region
        handles.append(ndt); src_hosts.append(sh); n_srcs.append(len(s))
        out_hosts.append(torch.empty((len(s), 4), dtype=torch.float32).pin_memory())
    map_build_ms = (time.perf_counter() - t0) * 1e3 / R
    info = handles[0].map_info()
    n_src = int(np.mean(n_srcs))
    streams = [torch.cuda.ExternalStream(hd.stream_ptr(), device=dev) for hd in handles]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev) if args.l2 == "flush" else None

    batch = nb.Batch(handles)
    master = torch.cuda.Stream(device=dev)

    # every pair once in the default (latency) shape: per-pair evaluation counts, first-use allocations
    per_pair = []
    for r in range(R):
        handles[r].align_async()
        handles[r].sync()
        per_pair.append(handles[r].result())

    # ---- latency arm: ONE align in flight at a time (the reference callers' pattern), default CTA shape -------------
    lat_steps = max(8, min(args.steps, args.latency_steps))
    prev = None
    lat_events = []
    for i in range(3 + lat_steps):
        hd, st = handles[i % R], streams[i % R]
        if flush is not None:
            with torch.cuda.stream(st):
                flush.fill_(1)  # evict L2 (not timed)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        if prev is not None:
            st.wait_event(prev)   # strictly sequential on the GPU
        e0.record(st)
        hd.align_async()
        e1.record(st)
        prev = e1
        if i >= 3:
            lat_events.append((e0, e1))
    torch.cuda.synchronize()
    lat_all = [a.elapsed_time(b) for a, b in lat_events]
    lat_ms = float(np.median(lat_all))

    # ---- throughput arm (the headline `value`): independent pairs in flight together, ndtb200_align_batch ------------
    # a STEP = one ndtb200_align_batch_async call over the R pairs of this GPU (R aligns, each one launch of the persistent
    # solve kernel on its own stream, throughput CTA shape, up to four solves co-resident per SM); steps follow each other
    # without a host wait.  A pair is touched again only R aligns (~R x 4 MB of other traffic) later: L2-cold inputs.
    def run_batches(n_steps):
        for _ in range(n_steps):
            batch.align_async()

    run_batches(max(args.warmup, 3))
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    barrier(world)
    torch.cuda.synchronize()
    sampler.start()
    for hd in handles:
        hd.reset_launch_count()
    t_wall0 = time.perf_counter()
    e_start = torch.cuda.Event(enable_timing=True)
    e_end = torch.cuda.Event(enable_timing=True)
    e_start.record(master)
    for st in streams:
        st.wait_event(e_start)
    run_batches(args.steps)
    for st in streams:
        e = torch.cuda.Event()
        e.record(st)
        master.wait_event(e)
    e_end.record(master)
    torch.cuda.synchronize()
    barrier(world)
    wall = time.perf_counter() - t_wall0
    launches = sum(hd.launch_count() for hd in handles)
    clocks = sampler.stop()
    total_ms = float(e_start.elapsed_time(e_end))
    co_resident_ms = float(np.mean([hd.last_align_ms() for hd in handles]))
    total_ms_max = max_over_ranks(total_ms, world, dev)
    n_aligns = args.steps * R                       # aligns of this rank inside the timed region
    value = n_aligns * world / (total_ms_max * 1e-3)
    ms_per_step = total_ms_max / args.steps
    used = per_pair                                  # every pair is aligned `steps` times
    evals = float(np.mean([u["n_evaluations"] for u in used]))
    hess = float(np.mean([u["n_hessian_passes"] for u in used]))
    pt_evals_local = float(args.steps * sum(n_srcs[i] * per_pair[i]["n_evaluations"] for i in range(R)))
    pt_iters = sum_over_ranks(pt_evals_local, world, dev) / (total_ms_max * 1e-3)

    # roofline of the dominant (only) kernel in the timed region: ndt_align_kernel
    kprobe = KPROBE[args.method]
    # SURVEY §8d, per launch.  The reference's Hessian-only passes are fused into the preceding line-search trial on the
    # device (no pass over the source of their own), so only the derivative evaluations count
    alg_bytes = float(np.mean([per_pair[i]["n_evaluations"] * n_srcs[i] * (16 + 4 * kprobe) + 64 * per_pair[i]["n_hits"]
                               for i in range(R)]))
    hits_total = float(np.mean([u["n_hits"] for u in used]))
    # up to four launches of the kernel are co-resident on every SM, so the GPU-level rate is what the roofline is
    # compared with: algorithmic bytes of all launches / device time of the timed region (= bytes per launch / the
    # launch's share of the device time).  The duration of one launch while sharing the SMs is reported beside it.
    kernel_ms = total_ms / n_aligns
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    achieved_latency = alg_bytes / (lat_ms * 1e-3) / 1e9
    peak, peak_src = hbm_peak()
    prof = {}
    prof_path = os.path.join(ROOT, "profiles", "align_kernel_ncu.json")
    if os.path.exists(prof_path):
        try:
            prof = json.load(open(prof_path)).get(args.method, {})
        except Exception:
            prof = {}
    traffic = prof.get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "ndt_align_kernel<%s>" % args.method, "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kernel_ms,
                "kernel_ms_note": "device time of the timed region / launches (launches overlap: up to 4 co-resident per SM)",
                "launch_duration_ms_while_sharing": co_resident_ms,
                "single_launch": {"kernel_ms": lat_ms, "achieved": achieved_latency, "frac": achieved_latency / peak,
                                  "note": "one align in flight, 148 x 1024-thread CTAs (the latency arm)"},
                "formula": "evals*N*(16+4K) + 64*hits",
                "reading": "frac is the CONTRACT figure (SURVEY 8d algorithmic bytes / time / measured copy bandwidth). The c2 working "
                           "set is L2-resident, so the DRAM the kernel really moves is ~100x smaller: see roofline_dram; the kernel "
                           "is bound by issue rate and by the serial reduction / Newton step behind its one barrier per evaluation. "
                           "HBM is the real bound only for map-sized problems: see `sharded` (100 M-point map)."}
    # what the kernel really moves through DRAM (ncu dram__bytes of the committed capture) over the same time, and how busy
    # the issue slots are (ncu smsp__issue_active of the same capture): the two numbers that say what limits it
    roofline_dram = None
    if traffic:
        roofline_dram = {"achieved": traffic / (kernel_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": traffic / (kernel_ms * 1e-3) / 1e9 / peak, "dram_bytes_per_launch": traffic,
                         "issue_active_frac_of_peak": prof.get("issue_active_pct_throughput_shape", prof.get("issue_active_pct")),
                         "warps_active_frac_of_peak": prof.get("warps_active_pct_throughput_shape", prof.get("warps_active_pct")),
                         "source": prof.get("source", "profiles/align_kernel_ncu.json"),
                         "note": "ncu numbers are per launch running ALONE (serialised replay); in the timed region up to four launches share each SM"}

    # end-to-end through the C ABI with host buffers: per pair H2D of the source (pinned), solve, D2H of the aligned
    # cloud and of the result block; pairs of a batch in flight together (ndtb200_set_source + ndtb200_align_batch)
    e2e_steps = max(1, min(args.steps, args.e2e_steps)) * R     # aligns: whole batches
    out_ptrs = [o.data_ptr() for o in out_hosts]

    def e2e_batch(wait):
        for i in range(R):
            handles[i].set_source_raw(src_hosts[i].data_ptr(), n_srcs[i], 16)
        if wait:
            return batch.align(None, out_ptrs, 16)
        batch.align_async(None, out_ptrs, 16)   # copies, solves, output + result copies enqueued; the next batch follows

    e2e_batch(True)   # first-use allocations stay outside the timed region
    barrier(world)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(e2e_steps // R):
        e2e_batch(False)                          # consecutive batches are pipelined: no host wait in between
        h2d += sum(n_srcs) * 16
        d2h += sum(n_srcs) * 16 + 416 * R
    batch.sync()                                  # every solve finished, every cloud and result block on the host
    torch.cuda.synchronize()
    barrier(world)
    e2e_dt = max_over_ranks(time.perf_counter() - t0, world, dev)
    e2e = {"value": e2e_steps * world / e2e_dt, "unit": "aligns/s", "h2d_bytes_per_step": h2d // e2e_steps,
           "d2h_bytes_per_step": d2h // e2e_steps, "steps": e2e_steps, "ms_per_step": e2e_dt / e2e_steps * 1e3,
           "api": "ndtb200_set_source (pinned host cloud) + ndtb200_align_batch_async (host output clouds + result blocks), %d pairs per call, "
                  "batches pipelined, ndtb200_sync per handle at the end" % R}
    # results only: the same pipeline without downloading the aligned cloud (out_points = NULL; the mapping node never reads
    # `aligned`, ndt_rosbag_mapping_node.cpp:120-144): D2H = the 416-byte result block per align
    def e2e_batch_pose_only():
        for i in range(R):
            handles[i].set_source_raw(src_hosts[i].data_ptr(), n_srcs[i], 16)
        batch.align_async(None, None, 16)

    e2e_batch_pose_only()
    batch.sync()
    barrier(world)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps // R):
        e2e_batch_pose_only()
    batch.sync()
    torch.cuda.synchronize()
    barrier(world)
    e2e_po_dt = max_over_ranks(time.perf_counter() - t0, world, dev)
    e2e["results_only"] = {"value": e2e_steps * world / e2e_po_dt, "unit": "aligns/s", "h2d_bytes_per_step": h2d // e2e_steps,
                           "d2h_bytes_per_step": 416, "ms_per_step": e2e_po_dt / e2e_steps * 1e3,
                           "api": "same pipeline with out_points = NULL: poses and counters come back, the aligned cloud stays on the device"}
    # the same through one blocking ndtb200_align at a time (what an unmodified caller of the reference's class does)
    n_single = max(8, min(64, e2e_steps))
    t0 = time.perf_counter()
    for i in range(n_single):
        hd = handles[i % R]
        hd.set_source_raw(src_hosts[i % R].data_ptr(), n_srcs[i % R], 16)
        hd.align_raw(None, out_hosts[i % R].data_ptr(), 16)
        hd.result()
    torch.cuda.synchronize()
    e2e_single_dt = time.perf_counter() - t0
    e2e["single_call"] = {"value": n_single / e2e_single_dt, "unit": "aligns/s", "ms_per_step": e2e_single_dt / n_single * 1e3,
                          "api": "ndtb200_set_source + ndtb200_align (blocking), one pair at a time"}

    res = per_pair[0]
    l2_note = ("inputs larger than L2: %d independent (scan, map) pairs per GPU cycled, ~4 MB touched per align" % R
               if flush is None else "flushed between timed iterations (256 MiB device write, untimed)")
    line = {"metric": "ndt_aligns_per_s", "value": value, "unit": "aligns/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args, n_src, n_tgt), "l2": l2_note,
                       "step": "one ndtb200_align_batch_async call = %d aligns (one per resident pair) per GPU; value = steps x %d x GPUs / time" % (R, R),
                       "aligns_per_step": R,
                       "parallelism": "independent scan pairs per GPU (%d in a batch, up to 4 solves co-resident per SM), no collective" % R,
                       "arith": "fp32 per-hit math, fp64 accumulation of the sums",
                       "map": {"voxels": info["n_voxels"], "valid": info["n_valid"], "build_ms_incl_h2d": map_build_ms}},
            "src_pt_iters_per_s": pt_iters, "evaluations_per_align": evals, "hessian_passes_per_align": hess,
            "hits_per_point_eval": hits_total / float(max(1.0, evals * n_src)),
            "latency": {"ms_per_align": lat_ms, "aligns_per_s": 1e3 / lat_ms,
                        "ms_max": float(np.max(lat_all)), "ms_mean": float(np.mean(lat_all)), "steps": len(lat_all),
                        "note": "one align in flight at a time, default CTA shape, inputs resident, CUDA events per launch (median)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "roofline_dram": roofline_dram,
            "topology": dict(TOPOLOGY) or None,
            "wall_s_timed_region": wall}
    # ---- the path with a real exchange step: one large source sharded over the GPUs against a map larger than L2 (BASELINE
    # configs[3]), the fused derivative-pass + 29-value NVLink exchange kernel; strong scaling (N = 1: the same solve on one GPU)
    if not args.no_sharded:
        del handles, batch, streams, src_hosts, out_hosts
        torch.cuda.empty_cache()
        try:
            if world > 1:
                import torch.distributed as dist_mod
            else:
                dist_mod = _SoloDist()
            line["sharded"] = c4_measure(args, rank, world, local, dev, dist_mod, steps=min(20, max(3, args.steps)),
                                         city_points=args.sharded_map_points, source_points=args.c4_source_points)
        except Exception as e:  # the headline must not be lost to a failure of the sub-benchmark
            line["sharded"] = {"error": repr(e)}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            cb, rr = cpu_baseline(w, args)
            line["cpu_baseline"] = cb
            dT = float(np.abs(rr["final"] - res["final"]).max())
            line["parity_vs_oracle"] = {"max_abs_dT": dT, "iterations_equal": rr["iterations"] == res["iterations"],
                                        "evaluations_equal": rr["n_evaluations"] == res["n_evaluations"]}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


class _SoloDist:
    """torch.distributed stand-in for one rank (ShardedNdt only asks for rank / world when world == 1)."""
    def get_rank(self):
        return 0

    def get_world_size(self):
        return 1

    def barrier(self):
        pass


def c4_measure(args, rank, world, local, dev, dist, steps, city_points, source_points):
    """BASELINE configs[3]: one large source cloud against a map much larger than L2, the source sharded over the N GPUs
    by contiguous ranges (every rank holds the full map), one 29-value in-kernel exchange per evaluation over NVLink
    (P2P mailboxes).  Strong scaling: total work is fixed.  Returns the record (rank 0) — timing = CUDA events around the
    solve kernel of every rank, sum over steps, max over ranks."""
    import torch
    import toyslam_b200 as nb
    import workloads
    from toyslam_b200.sharding import ShardedNdt, source_range
    ndt = nb.NormalDistributionsTransform(device=local)
    ndt.setNeighborhoodSearchMethod(METHODS[args.method])
    sh = ShardedNdt(ndt, dist)
    if city_points > 0:
        # a synthetic city map of `city_points` surface samples (ground, walls of a 40 m block grid, roofs; every rank
        # generates the same cloud on its GPU) and a source sampled from the same surfaces within 100 m of the origin,
        # moved by a small pose
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import build_bench
        tgt_dev = build_bench.surface_points(city_points, 20260103, extent=args.c4_city_extent, device=dev)
        src_dev = build_bench.surface_points(source_points, 20260105, extent=200.0, device=dev)
        T = torch.as_tensor(np.linalg.inv(workloads.pose_matrix([0.3, -0.2, 0.1, 0.004, -0.003, 0.015])), dtype=torch.float64, device=dev)
        xyz = src_dev[:, :3].to(torch.float64) @ T[:3, :3].T + T[:3, 3]
        source = np.ascontiguousarray(xyz.to(torch.float32).cpu().numpy())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ndt.set_target_device_view(tgt_dev.data_ptr(), city_points)
        keep_alive = tgt_dev  # noqa: F841  (the view's buffer)
        torch.cuda.synchronize()
        build_ms = (time.perf_counter() - t0) * 1e3
        n_target = city_points
        del src_dev, xyz
    else:
        box = [None]
        if rank == 0:
            scene, target = workloads.config2_map(map_points=args.map_points, n_map_scans=args.map_scans, azimuth_steps=args.azimuth_steps,
                                                  thin_leaf=args.thin_leaf)
            srcs = [workloads.config2_scan(scene, 100 + i, azimuth_steps=args.azimuth_steps, perturb_seed=100)[0] for i in range(args.c4_scans)]
            box[0] = (target, np.concatenate(srcs))
        if world > 1:
            dist.broadcast_object_list(box, src=0)
        target, source = box[0]
        t0 = time.perf_counter()
        sh.setInputTarget(target)
        build_ms = (time.perf_counter() - t0) * 1e3
        n_target = len(target)
    sh.setInputSource(source)
    lo, hi = source_range(len(source), rank, world)
    for _ in range(3):
        sh.align()
    res = sh.result()
    times = []
    for _ in range(steps):
        dist.barrier()
        torch.cuda.synchronize()
        ndt.align_async()
        ndt.sync()
        times.append(ndt.last_align_ms())
    total_ms = max_over_ranks(float(np.sum(times)), world, dev)
    evals, hess = res["n_evaluations"], res["n_hessian_passes"]
    kprobe = KPROBE[args.method]
    alg_bytes = evals * len(source) * (16 + 4 * kprobe) + 64 * res["n_hits"]
    peak, peak_src = hbm_peak()
    info = ndt.map_info()
    ms = total_ms / steps
    rec = {"metric": "ndt_aligns_per_s", "workload": "c4", "value": steps / (total_ms * 1e-3), "unit": "aligns/s", "n_gpus": world,
           "steps": steps, "ms_per_step": ms, "ms_per_align": ms, "scaling": "strong", "dtype": "f32", "data": "synthetic",
           "src_pt_iters_per_s": len(source) * evals * steps / (total_ms * 1e-3), "evaluations_per_align": evals,
           "hessian_passes_per_align": hess, "hits_per_point_eval": res["n_hits"] / float(evals * len(source)),
           "config": {"workload": "c4: %d-pt source (%s) sharded by contiguous ranges over %d GPU(s) vs %d-pt map "
                                  "(%d voxels, %d valid; cell table + records %.1f GB > L2), res 1.0, %s; one in-kernel 29-value P2P exchange "
                                  "per evaluation; every rank's slice sorted by voxel key once at setInputSource" %
                                  (len(source), ("%d merged scans" % args.c4_scans) if city_points == 0 else "synthetic city surfaces",
                                   world, n_target, info["n_voxels"], info["n_valid"],
                                   (4.0 * float(np.prod(info["div_b"].astype(np.float64))) + 64.0 * info["n_voxels"]) / 1e9, args.method),
                      "map_build_ms_incl_h2d": build_ms, "points_this_rank": hi - lo,
                      "timing": "CUDA events around each rank's solve kernel, sum over steps, max over ranks"},
           "roofline": {"bound": "hbm", "achieved": alg_bytes / (ms * 1e-3) / 1e9 / world, "peak": peak, "peak_source": peak_src,
                        "unit": "GB/s per GPU", "frac": alg_bytes / (ms * 1e-3) / 1e9 / world / peak,
                        "algorithmic_bytes_per_launch": alg_bytes, "formula": "evals*N*(16+4K) + 64*hits, all GPUs together"},
           "converged": res["converged"], "iterations": res["iterations"]}
    ndt.comm_detach()
    del sh, ndt
    torch.cuda.empty_cache()
    return rec


def run_c4(args):
    """--workload c4: the source-sharded scan-to-map solve on its own (see c4_measure)."""
    import torch
    rank, world, local = dist_setup(args.gpus)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
    else:
        dist = _SoloDist()
    line = c4_measure(args, rank, world, local, dev, dist, steps=min(args.steps, 200), city_points=args.c4_city_points,
                      source_points=args.c4_source_points)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def c3_workload(args, rank):
    """Consecutive scan pairs of the simulated drive (BASELINE configs[2]); `--c3-distinct` distinct pairs per rank,
    cycled to the requested number of pairs.  Cached like the c2 workload."""
    import workloads
    n = args.c3_distinct
    path = None
    if args.cache:
        os.makedirs(args.cache, exist_ok=True)
        path = os.path.join(args.cache, "c3_%d_%d_r%d.npz" % (n, args.azimuth_steps, rank))
        if os.path.exists(path):
            d = np.load(path)
            return [d["scan_%d" % i] for i in range(n + 1)], [tuple(p) for p in d["poses"]]
    scans, poses = workloads.config3_sequence(n + 1, seed=workloads.SEED_C3 + 1000 * rank, azimuth_steps=args.azimuth_steps)
    if path:
        np.savez(path, poses=np.asarray(poses), **{"scan_%d" % i: s for i, s in enumerate(scans)})
    return scans, poses


C3_PARAMS = dict(eps=0.01, max_iter=64, step=0.1, res=1.0)  # ndt_rosbag_mapping_node.cpp:83-88


def run_c3(args):
    """BASELINE configs[2]: batched scan-to-scan odometry.  A step = one pair: build the target map of scan k
    (setInputTarget, inside the timed region as the node does), setInputSource(scan k+1), align(guess = true motion of
    the previous pair).  Pairs are independent units: round-robin over the GPUs (ranks), and inside a GPU over
    `--c3-lanes` handles driven by one host thread each, so builds, copies and solves of different pairs overlap."""
    import threading
    import torch
    import toyslam_b200 as nb
    import workloads
    rank, world, local = dist_setup(args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    scans, poses = c3_workload(args, rank)
    nd = args.c3_distinct
    pairs_total = args.steps
    pairs_rank = pairs_total // world
    hosts = []
    for sc in scans:
        hb = torch.ones((len(sc), 4), dtype=torch.float32).pin_memory()
        hb[:, :3] = torch.from_numpy(sc)
        hosts.append(hb)
    guesses = [np.eye(4)] + [workloads.relative_pose_matrix(poses[k - 1], poses[k]) for k in range(1, nd)]
    truths = [workloads.relative_pose_matrix(poses[k], poses[k + 1]) for k in range(nd)]
    # lanes = handles + host threads per GPU (their waits yield the core: throughput-mode handles poll with sched_yield).
    # Measured on three boxes, staged builds: 8 lanes 9.8 k / 10.8 k / 10.7 k pairs/s, 16 lanes 7.3 k / 9.5 k / 12.6 k — 8 is
    # the steadier choice; C++ and Python lane threads measure the same (10.8 k vs 10.75 k): the lanes are bound by the
    # ~12 launches + 2 synchronisations of a staged scan-sized build, not by the host language
    # With several ranks on one host the lane threads must not oversubscribe the rank's cores (their waits poll): 8 lanes
    # where a rank has >= 8 cores, fewer otherwise (8 GPUs on a 32-core host: 4) — measured 42.4 k pairs/s at 8 GPUs
    # with 64 polling threads on 32 cores against 38.0 k at 4 GPUs
    L = args.c3_lanes if args.c3_lanes > 0 else max(2, min(8, host_threads()))
    lanes = []
    for _ in range(L):
        ndt = nb.NormalDistributionsTransform(device=local)
        ndt.setNeighborhoodSearchMethod(METHODS[args.method])
        ndt.setTransformationEpsilon(C3_PARAMS["eps"])
        ndt.setMaximumIterations(C3_PARAMS["max_iter"])
        ndt.setStepSize(C3_PARAMS["step"])
        ndt.setResolution(C3_PARAMS["res"])
        ndt.set_throughput_mode(True)
        lanes.append(ndt)
    results = [None] * nd

    def lane_loop(li, first, count):
        torch.cuda.set_device(local)
        hd = lanes[li]
        for j in range(first + li, first + count, L):
            k = j % nd
            hd.set_target_raw(hosts[k].data_ptr(), len(scans[k]), 16)          # H2D + voxel-map build
            hd.set_source_raw(hosts[k + 1].data_ptr(), len(scans[k + 1]), 16)  # H2D
            hd.align_async(guesses[k])
            hd.sync()                                                            # the node reads every pose
            results[k] = hd.result()

    def run_python(first, count):
        th = [threading.Thread(target=lane_loop, args=(li, first, count)) for li in range(L)]
        for t in th:
            t.start()
        for t in th:
            t.join()

    # native lane driver (default): ndtb200_run_pairs drives every lane from its own C++ host thread — no Python
    # threads, no GIL hand-offs between the ~5 calls of a pair
    pipes = {}

    def prepare_native(first, count):  # pointer / size / guess arrays of the pair list, built outside the timed region
        if (first, count) not in pipes:
            ks = [j % nd for j in range(first, first + count)]
            pipes[(first, count)] = (ks, nb.PairPipeline(lanes, [(hosts[k].data_ptr(), len(scans[k])) for k in ks],
                                                         [(hosts[k + 1].data_ptr(), len(scans[k + 1])) for k in ks], [guesses[k] for k in ks]))
        return pipes[(first, count)]

    def run_native(first, count):
        ks, pipe = prepare_native(first, count)
        pipe.run()
        for k, r in zip(ks, pipe.results()):
            results[k] = r

    run = run_python if args.c3_driver == "python" else run_native

    run(0, max(args.warmup, L))  # warm-up: allocations, first launches
    if run is run_native:
        prepare_native(0, pairs_rank)
    torch.cuda.synchronize()
    for hd in lanes:
        hd.reset_launch_count()
    sampler = ClockSampler(local)
    barrier(world)
    torch.cuda.synchronize()
    sampler.start()
    master = torch.cuda.Stream(device=dev)
    streams = [torch.cuda.ExternalStream(hd.stream_ptr(), device=dev) for hd in lanes]
    e_start = torch.cuda.Event(enable_timing=True)
    e_end = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e_start.record(master)
    for st in streams:
        st.wait_event(e_start)
    run(0, pairs_rank)
    for st in streams:
        e = torch.cuda.Event()
        e.record(st)
        master.wait_event(e)
    e_end.record(master)
    torch.cuda.synchronize()
    barrier(world)
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    total_ms = max_over_ranks(float(e_start.elapsed_time(e_end)), world, dev)
    launches = sum(hd.launch_count() for hd in lanes)
    done = [r for r in results if r is not None]
    err_t = []
    for k, r in enumerate(results):
        if r is not None:
            err_t.append(float(np.linalg.norm(r["final"][:3, 3] - truths[k][:3, 3])))
    evals = float(np.mean([r["n_evaluations"] for r in done]))
    pts = float(np.mean([len(s) for s in scans]))
    value = pairs_rank * world / (total_ms * 1e-3)
    line = {"metric": "ndt_aligns_per_s", "workload": "c3", "value": value, "unit": "aligns/s", "n_gpus": world, "steps": pairs_rank * world,
            "warmup": max(args.warmup, L), "ms_per_step": total_ms / pairs_rank, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "c3: %d consecutive scan pairs of a simulated drive (%d distinct pairs per GPU cycled), scans downsampled "
                                   "0.3 m (%.0f pts mean), per pair: target-map build + align, node parameters (eps 0.01, max_iter 64, "
                                   "step 0.1, res 1.0, %s), guess = true motion of the previous pair; %d lanes (handles + %s host threads) per GPU"
                                   % (pairs_rank * world, nd, pts, args.method, L, "C++" if args.c3_driver == "native" else "Python"),
                       "parallelism": "pairs round-robin over GPUs and lanes; no collective"},
            "src_pt_iters_per_s": value * pts * evals, "evaluations_per_align": evals,
            "hessian_passes_per_align": float(np.mean([r["n_hessian_passes"] for r in done])),
            "iterations_per_align": float(np.mean([r["iterations"] for r in done])),
            "mean_translation_error_vs_truth_m": float(np.mean(err_t)), "max_translation_error_vs_truth_m": float(np.max(err_t)),
            "clocks": clocks, "gpu_launches": int(launches), "wall_s_timed_region": wall, "topology": dict(TOPOLOGY) or None,
            "e2e": {"value": pairs_rank * world / max_over_ranks(wall, world, dev), "unit": "aligns/s",
                    "h2d_bytes_per_step": int(2 * pts * 16), "d2h_bytes_per_step": 416,
                    "note": "the timed region IS end to end here: host scans (pinned) in, poses out, wall clock"}}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], line["parity_vs_oracle"] = c3_cpu(args, scans, guesses, results, max_pairs=24)
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


def c3_cpu(args, scans, guesses, gpu_results, max_pairs=24, max_seconds=25.0):
    import oracle
    ref = oracle.NormalDistributionsTransform()
    ref.setNeighborhoodSearchMethod({"DIRECT1": oracle.DIRECT1, "DIRECT7": oracle.DIRECT7, "DIRECT26": oracle.DIRECT26}[args.method])
    ref.setTransformationEpsilon(C3_PARAMS["eps"])
    ref.setMaximumIterations(C3_PARAMS["max_iter"])
    ref.setStepSize(C3_PARAMS["step"])
    ref.setResolution(C3_PARAMS["res"])
    t0 = time.perf_counter()
    n = 0
    max_dT, same_counts = 0.0, True
    for k in range(min(max_pairs, len(scans) - 1)):
        ref.setInputTarget(scans[k])
        ref.setInputSource(scans[k + 1])
        ref.align(guesses[k])
        n += 1
        rr = ref.result()
        if gpu_results[k] is not None:
            max_dT = max(max_dT, float(np.abs(rr["final"] - gpu_results[k]["final"]).max()))
            same_counts = same_counts and rr["iterations"] == gpu_results[k]["iterations"] and rr["n_evaluations"] == gpu_results[k]["n_evaluations"]
        if time.perf_counter() - t0 > max_seconds:
            break
    dt = time.perf_counter() - t0
    return ({"value": n / dt, "unit": "aligns/s", "cores": oracle.max_threads(), "kind": "port",
             "sample": "the first %d pairs of the same sequence, map build + align per pair, oracle C++/OpenMP port" % n},
            {"pairs_checked": n, "max_abs_dT": max_dT, "iterations_and_evaluations_equal": bool(same_counts)})


def run_mapper(args):
    """the ndt_rosbag_mapping_node loop (downsample 0.3 m, scan-to-scan NDT with the node parameters, pose
    chaining, global map re-voxelised at 0.5 m) as a device-resident pipeline.  A step = one raw scan pushed.  Ranks
    run independent drives (replicas)."""
    import torch
    import toyslam_b200 as nb
    import workloads
    rank, world, local = dist_setup(args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    n_scans = min(args.steps, args.mapper_scans)
    path = None
    scans = None
    if args.cache:
        os.makedirs(args.cache, exist_ok=True)
        path = os.path.join(args.cache, "mapper_%d_%d_r%d.npz" % (n_scans, args.azimuth_steps, rank))
        if os.path.exists(path):
            d = np.load(path)
            scans = [d["scan_%d" % i] for i in range(n_scans)]
    if scans is None:
        scans, _ = workloads.config3_sequence(n_scans, seed=workloads.SEED_C3 + 1000 * rank, azimuth_steps=args.azimuth_steps, leaf=0.02)
        if path:
            np.savez(path, **{"scan_%d" % i: s for i, s in enumerate(scans)})
    hosts = []
    for sc in scans:
        hb = torch.ones((len(sc), 4), dtype=torch.float32).pin_memory()
        hb[:, :3] = torch.from_numpy(sc)
        hosts.append(hb)
    warm = nb.Mapper(device=local)
    for k in range(min(4, n_scans)):
        warm.push_scan_raw(hosts[k].data_ptr(), len(scans[k]), 16)
    del warm
    mapper = nb.Mapper(device=local)
    sampler = ClockSampler(local)
    barrier(world)
    torch.cuda.synchronize()
    sampler.start()
    t0 = time.perf_counter()
    steps = []
    step_ms = []
    for k in range(n_scans):
        tk = time.perf_counter()
        steps.append(mapper.push_scan_raw(hosts[k].data_ptr(), len(scans[k]), 16))
        step_ms.append((time.perf_counter() - tk) * 1e3)
    torch.cuda.synchronize()
    barrier(world)
    wall = max_over_ranks(time.perf_counter() - t0, world, dev)
    clocks = sampler.stop()
    raw_pts = float(np.mean([len(s) for s in scans]))
    line = {"metric": "mapper_scans_per_s", "workload": "mapper", "value": n_scans * world / wall, "unit": "scans/s", "n_gpus": world,
            "steps": n_scans * world, "warmup": 4, "ms_per_step": wall / n_scans * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "mapping-node loop: %d raw scans (%.0f pts mean) of a simulated drive per GPU; per scan VoxelGrid 0.3 m (%.0f pts), "
                                   "NDT vs the previous scan (eps 0.01, 64 iterations, DIRECT7, guess = previous transform), getFitnessScore, "
                                   "global map += transformed scan, VoxelGrid 0.5 m (final map %d pts)" %
                                   (n_scans, raw_pts, float(np.mean([s["n_filtered"] for s in steps])), steps[-1]["n_map"]),
                       "timing": "host wall clock over the whole drive, pinned host scans in, step records out (this IS end to end)"},
            "iterations_per_scan": float(np.mean([s["iterations"] for s in steps[1:]])),
            "evaluations_per_scan": float(np.mean([s["n_evaluations"] for s in steps[1:]])),
            "step_ms_percentiles": {"p10": float(np.percentile(step_ms, 10)), "p50": float(np.percentile(step_ms, 50)),
                                    "p90": float(np.percentile(step_ms, 90)), "max": float(np.max(step_ms)), "first_half_mean": float(np.mean(step_ms[:len(step_ms) // 2])),
                                    "second_half_mean": float(np.mean(step_ms[len(step_ms) // 2:]))},
            "gpu_launches": mapper.launch_count(), "clocks": clocks,
            "e2e": {"value": n_scans * world / wall, "unit": "scans/s", "h2d_bytes_per_step": int(raw_pts * 16), "d2h_bytes_per_step": 168}}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from util import oracle_mapping_loop
            import oracle
            n_cpu = min(n_scans, 12)
            t0 = time.perf_counter()
            ref_steps, _ = oracle_mapping_loop(scans[:n_cpu])
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": n_cpu / dt, "unit": "scans/s", "cores": oracle.max_threads(), "kind": "port",
                                    "sample": "the first %d scans of the same drive through the oracle's restatement of the node loop" % n_cpu}
            line["parity_vs_oracle"] = {"scans_checked": n_cpu,
                                        "max_abs_dT": float(max(np.abs(a["transform"] - b["transform"]).max() for a, b in zip(steps[:n_cpu], ref_steps))),
                                        "counts_equal": bool(all(a["iterations"] == b["iterations"] and a["n_filtered"] == b["n_filtered"]
                                                                 for a, b in zip(steps[:n_cpu], ref_steps)))}
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


def run_c5(args):
    """BASELINE configs[4]: target-map rebuild sweep.  A step = one VoxelGridCovariance build (ndtb200_set_target_device:
    device-resident cloud in, voxel map out).  One JSON line per (points, resolution)."""
    import torch
    import toyslam_b200 as nb
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import build_bench
    rank, world, local = dist_setup(args.gpus)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    peak = 6454.9
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    import torch.distributed as dist
    from toyslam_b200.sharding import ShardedNdt
    for m_total in args.c5_points:
        m = m_total // world          # points by contiguous range: every rank reduces its slice, partials are exchanged
        pts = build_bench.surface_points(m, 20260104 + rank, device=dev)
        for res in args.c5_res:
            ndt = nb.NormalDistributionsTransform(device=local)
            ndt.setResolution(res)
            if world > 1:
                sh = ShardedNdt(ndt, dist)
                build = lambda: sh.setInputTargetShardedDevice(pts)
            else:
                build = lambda: ndt.set_target_device_view(pts.data_ptr(), m)   # no copy: the map references the caller's device cloud
            build()
            build()
            reps = max(3, min(args.steps, 10))
            times = []
            for _ in range(reps):
                barrier(world)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                build()
                torch.cuda.synchronize()
                times.append(time.perf_counter() - t0)
            ms = max_over_ranks(float(np.min(times)) * 1e3, world, dev)
            info = ndt.map_info()
            alg = 16.0 * m * world + 72.0 * info["n_voxels"]
            if rank == 0:
                print(json.dumps({"metric": "map_build_points_per_s", "workload": "c5", "value": m * world / (ms * 1e-3), "unit": "points/s",
                                  "n_gpus": world, "steps": reps, "ms_per_step": ms, "scaling": "strong", "dtype": "f64 moments / int32 keys",
                                  "data": "synthetic",
                                  "config": {"workload": "c5: VoxelGridCovariance build, %d points, resolution %.1f%s" %
                                             (m * world, res, "" if world == 1 else " (sharded: %d points per GPU; partials sent to the owner of their key range (all-to-all), owners merge + finalise, records all-gathered, NCCL)" % m),
                                             "voxels": info["n_voxels"], "valid": info["n_valid"],
                                             "timing": "host wall clock around the call, device synchronised on both sides, best of %d, max over ranks" % reps},
                                  "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9 / world, "peak": peak, "unit": "GB/s per GPU",
                                               "frac": alg / (ms * 1e-3) / 1e9 / world / peak,
                                               "algorithmic_bytes": alg, "formula": "16*M + 72*V (SURVEY 8d)"}}), flush=True)
            del ndt
        del pts
        torch.cuda.empty_cache()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None,
                    help="timed steps; c2 (default 64): one step = one ndtb200_align_batch_async call over the --replicas pairs of a GPU; "
                         "c3 (default 4096): pairs; mapper (200): scans; c4 (20): aligns; c5 (10): builds; c1 (10): aligns of the `10times` loop")
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--method", default="DIRECT7", choices=list(METHODS))
    ap.add_argument("--map-points", type=int, default=1_000_000)
    ap.add_argument("--map-scans", type=int, default=31)
    ap.add_argument("--azimuth-steps", type=int, default=1875)
    ap.add_argument("--e2e-steps", type=int, default=8, help="batches of the end-to-end (host buffers) arm")
    ap.add_argument("--no-sharded", action="store_true", help="skip the source-sharded scan-to-map sub-benchmark (`sharded` record)")
    ap.add_argument("--sharded-map-points", type=int, default=100_000_000,
                    help="points of the synthetic city map of the `sharded` record (BASELINE configs[3]: 100 M)")
    ap.add_argument("--latency-steps", type=int, default=200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--replicas", type=int, default=64, help="independent (scan, map) pairs per GPU cycled by the timed loop")
    ap.add_argument("--l2", default="inputs", choices=["inputs", "flush"],
                    help="how timed steps see a cold L2: inputs larger than L2 (default) or a 256 MiB flush write")
    ap.add_argument("--cache", default=None, help="directory for cached workload arrays")
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5", "mapper"])
    ap.add_argument("--mapper-scans", type=int, default=200)
    ap.add_argument("--c3-distinct", type=int, default=128, help="distinct consecutive pairs generated per GPU (cycled)")
    ap.add_argument("--c3-driver", choices=["native", "python"], default="native",
                    help="c3: lanes driven by C++ host threads inside ndtb200_run_pairs (default) or by Python threads")
    ap.add_argument("--c3-lanes", type=int, default=0,
                    help="handles (+ host threads) per GPU for the c3 pipeline; 0 = auto (16)")
    ap.add_argument("--c5-points", type=int, nargs="+", default=[10_000_000, 100_000_000])
    ap.add_argument("--c5-res", type=float, nargs="+", default=[0.5, 1.0, 2.0])
    ap.add_argument("--c4-scans", type=int, default=16)
    ap.add_argument("--c4-city-points", type=int, default=0, help="c4 at full size: synthetic city map of this many points (e.g. 100000000)")
    ap.add_argument("--c4-source-points", type=int, default=2_000_000)
    ap.add_argument("--c4-city-extent", type=float, default=2000.0)
    ap.add_argument("--thin-leaf", type=float, default=0.1)
    args = ap.parse_args()
    args.steps_given = args.steps is not None
    if args.steps is None:
        args.steps = {"c1": 10, "c2": 64, "c3": 4096, "c4": 20, "c5": 10, "mapper": 200}[args.workload]
    if args.workload == "c4" and args.impl == "b200":
        return run_c4(args)
    if args.workload == "c3" and args.impl == "b200":
        return run_c3(args)
    if args.workload == "c5" and args.impl == "b200":
        return run_c5(args)
    if args.workload == "mapper" and args.impl == "b200":
        return run_mapper(args)
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "c1":
        return run_c1(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
