"""bench.py contract on a GPU box, at toy sizes (seconds): every workload prints ONE well-formed JSON line with the
keys the driver reads (metric / value / unit / n_gpus / steps / ms_per_step / clocks / e2e / gpu_launches / roofline /
cpu_baseline), and the CPU checker inside it agrees with the GPU result."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--map-points", "30000", "--map-scans", "3", "--azimuth-steps", "300"]


def _run(args, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-3000:])
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    return [json.loads(l) for l in lines]


def test_bench_default_contract():
    (d,) = _run(["--steps", "24", "--warmup", "3", "--replicas", "4", "--e2e-steps", "8", "--latency-steps", "8",
                 "--sharded-map-points", "3000000", "--c4-source-points", "300000"] + SMALL)
    assert d["metric"] == "ndt_aligns_per_s" and d["unit"] == "aligns/s" and d["n_gpus"] == 1 and d["steps"] == 24 and d["warmup"] == 3
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["gpu_launches"] == 24 * 4 and d["config"]["aligns_per_step"] == 4     # a step = one batch call = one solve kernel per resident pair
    assert abs(d["value"] - 24 * 4 / (d["ms_per_step"] * 24 * 1e-3)) <= 1e-6 * d["value"]
    sh = d["sharded"]                                                  # the source-sharded scan-to-map sub-record (N = 1 here)
    assert sh["workload"] == "c4" and sh["scaling"] == "strong" and sh["n_gpus"] == 1 and sh["ms_per_align"] > 0 and sh["converged"]
    assert 0 < sh["roofline"]["frac"] < 2 and sh["roofline"]["unit"] == "GB/s per GPU"
    rd = d["roofline_dram"]
    assert rd is None or (rd["frac"] > 0 and rd["dram_bytes_per_launch"] > 0)
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["e2e"]["single_call"]["value"] > 0 and d["latency"]["ms_per_align"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0
    p = d["parity_vs_oracle"]
    assert p["iterations_equal"] and p["evaluations_equal"] and p["max_abs_dT"] < 1e-4
    assert "clocks" in d and "config" in d and "workload" in d["config"]


def test_bench_reference_arm_contract():
    (d,) = _run(["--impl", "reference", "--steps", "3", "--warmup", "1", "--replicas", "2"] + SMALL)
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))      # the thread count is set explicitly (torchrun exports OMP_NUM_THREADS=1)


def test_bench_c1_protocol():
    """--workload c1: the apps/align.cpp protocol on the bundled pair — single / 10times / fitness per search method, the
    README's fitness values reproduced, the device VoxelGrid front end equal to the committed fixture."""
    (d,) = _run(["--workload", "c1"])
    assert d["workload"] == "c1" and d["steps"] == 10 and d["value"] > 0 and d["vs_baseline"] > 1
    assert d["config"]["device_voxelgrid_equals_fixture"] is True
    for name in ("KDTREE", "DIRECT7", "DIRECT1"):
        m = d["methods"][name]
        assert m["fitness_matches_readme"] and m["single_ms"] > 0 and m["ktimes_ms"] > 0 and m["k"] == 10
        assert m["readme_i7_6700K_ms"]["8thr"]["10times"] > 0
        c = d["cpu_baseline"]["methods_all_threads"][name]
        assert "%.6f" % c["fitness"] == "%.6f" % m["fitness"] and c["iterations"] == m["iterations"]


def test_bench_other_workloads():
    (d,) = _run(["--workload", "c3", "--steps", "32", "--c3-distinct", "8", "--c3-lanes", "4", "--azimuth-steps", "300"])
    assert d["workload"] == "c3" and d["value"] > 0 and d["gpu_launches"] > 32
    assert d["parity_vs_oracle"]["iterations_and_evaluations_equal"] and d["parity_vs_oracle"]["max_abs_dT"] < 1e-4
    lines = _run(["--workload", "c5", "--c5-points", "200000", "--c5-res", "1.0", "2.0", "--steps", "3"])
    assert len(lines) == 2 and all(l["metric"] == "map_build_points_per_s" and l["value"] > 0 and 0 < l["roofline"]["frac"] < 1 for l in lines)
    (d,) = _run(["--workload", "mapper", "--steps", "8", "--azimuth-steps", "300"])
    assert d["workload"] == "mapper" and d["value"] > 0 and d["parity_vs_oracle"]["counts_equal"] and d["parity_vs_oracle"]["max_abs_dT"] < 1e-4
