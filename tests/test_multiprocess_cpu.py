"""world_size-2 tests of the N>1 host logic on CPU (gloo): rank layout, max/sum-over-ranks helpers of bench.py,
the reference arm under torchrun (rank 0 prints one JSON line, the other ranks exit 0 without work), and the
source-range partition used by the sharded align."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(args, nproc=2, port=29611, timeout=300):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port)] + args
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = ""
    env["OMP_NUM_THREADS"] = "2"
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)


def test_reference_arm_under_torchrun_world2():
    out = _torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1", "--map-points", "30000",
                     "--map-scans", "3", "--azimuth-steps", "200", "--replicas", "2"])
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1            # rank 0 alone prints
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] == "port"


def test_rank_reductions_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys, json\n"
        "sys.path.insert(0, %r)\n"
        "import torch, torch.distributed as dist\n"
        "import bench\n"
        "rank, world, local = bench.dist_setup(2)\n"
        "dev = torch.device('cpu')\n"
        "mx = bench.max_over_ranks(float(rank + 1), world, dev)\n"
        "sm = bench.sum_over_ranks(float(rank + 1), world, dev)\n"
        "bench.barrier(world)\n"
        "from toyslam_b200.sharding import source_range\n"
        "lo, hi = source_range(1000003, rank, world)\n"
        "os.write(1, (json.dumps({'rank': rank, 'world': world, 'max': mx, 'sum': sm, 'lo': lo, 'hi': hi}) + '\\n').encode())  # one atomic write per rank\n"
        "dist.destroy_process_group()\n" % ROOT)
    out = _torchrun([str(script)], port=29612)
    assert out.returncode == 0, out.stderr[-2000:]
    recs = sorted((json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")), key=lambda r: r["rank"])
    assert [r["rank"] for r in recs] == [0, 1] and all(r["world"] == 2 for r in recs)
    assert all(r["max"] == 2.0 and r["sum"] == 3.0 for r in recs)
    assert recs[0]["lo"] == 0 and recs[0]["hi"] == recs[1]["lo"] and recs[1]["hi"] == 1000003


def test_source_range_partition_properties():
    from toyslam_b200.sharding import source_range
    for n in (0, 1, 31, 32, 33, 117472, 2_000_000):
        for world in (1, 2, 4, 8):
            edges = [source_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for a, b in zip(edges[:-1], edges[1:]):
                assert a[1] == b[0]
            sizes = np.array([hi - lo for lo, hi in edges])
            assert sizes.min() >= 0 and (sizes.max() - sizes[sizes > 0].min() <= 32 if (sizes > 0).any() else True)
            # boundaries fall on 32-point groups so every rank's warps stay fully coalesced
            assert all(lo % 32 == 0 for lo, hi in edges if hi > lo)


def test_point_range_partition_properties():
    """Contiguous target-point ranges of the sharded map build: cover [0, n) exactly, sizes differ by at most one."""
    from toyslam_b200.sharding import point_range
    for n in (0, 1, 7, 8, 9, 1_000_003, 500_000_000):
        for world in (1, 2, 3, 4, 8):
            edges = [point_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges[:-1], edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert min(sizes) >= 0 and max(sizes) - min(sizes) <= 1


def test_c_abi_header_symbols_are_exported():
    """Every function declared in include/ndt_b200.h is exported by libndt_b200.so (no compute call: no GPU needed)."""
    import toyslam_b200 as nb
    L = nb.load_library()
    names = nb.exported_symbols()
    assert len(names) >= 48
    for name in names:
        assert hasattr(L, name), name
    for must in ("ndtb200_align", "ndtb200_align_batch", "ndtb200_set_target", "ndtb200_voxelgrid_filter", "ndtb200_mapper_push_scan",
                 "ndtb200_build_from_partials", "ndtb200_fitness_score"):
        assert must in names


def test_rank_core_binding_helpers(monkeypatch):
    """bench.py binds a rank to the cores of its GPU's NUMA node (or to an even share of the allowed cores): the cpulist
    parser, the single-rank no-op and the even split between two local ranks (no GPU / no NVML here: the fallback)."""
    sys.path.insert(0, ROOT)
    import bench
    assert bench._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert bench._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    try:
        bench.TOPOLOGY.clear()
        bench.bind_rank_to_gpu(0, 1)
        assert os.sched_getaffinity(0) == before and bench.TOPOLOGY["bound_cores"] == len(before)
        if len(before) >= 2:
            monkeypatch.setenv("LOCAL_WORLD_SIZE", "2")
            shares = []
            for local in (0, 1):
                os.sched_setaffinity(0, before)
                bench.TOPOLOGY.clear()
                bench.bind_rank_to_gpu(local, 2)
                shares.append(os.sched_getaffinity(0))
            assert shares[0] and shares[1] and not (shares[0] & shares[1]) and (shares[0] | shares[1]) <= before
            assert len(shares[0]) == len(before) // 2
    finally:
        os.sched_setaffinity(0, before)
