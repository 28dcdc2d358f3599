"""BASELINE.json configs[1] at its full size (117 k-point scan vs 0.93 M-point map) and a 4 M-point build: parity through
size-independent properties of the domain, plus the oracle on the full-size pair (it finishes in seconds)."""
import os
import sys

import numpy as np
import pytest

import oracle
from util import rel_err, transform_delta

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def nb():
    import toyslam_b200
    return toyslam_b200


@pytest.fixture(scope="module")
def c2():
    sys.path.insert(0, ROOT)
    import bench

    class A:
        pass
    a = A()
    a.map_points, a.map_scans, a.azimuth_steps, a.cache, a.method = 1_000_000, 31, 1875, "/tmp/wl", "DIRECT7"
    w = bench.make_workload(a, 0, 2)
    return w["target"], w["sources"][0], w["sources"][1]


def test_full_size_map_properties(nb, c2):
    tgt, src, _ = c2
    g = nb.NormalDistributionsTransform()
    assert g.setInputTarget(tgt) == 0
    d = g.dump_voxels()
    info = g.map_info()
    assert np.all(np.diff(d["keys"]) > 0)                                   # one record per cell, ascending
    assert int(np.where(d["counts"] < 0, 0, d["counts"]).sum()) + 0 <= len(tgt)
    keys = g.point_keys()
    cnt = np.bincount(keys, minlength=int(d["keys"].max()) + 1)[d["keys"]]
    ok = d["counts"] >= 0
    assert np.array_equal(cnt[ok], d["counts"][ok]) and cnt.sum() == len(tgt)  # checksum of checksums: every point in one voxel
    assert info["n_valid"] == int((d["counts"] >= 6).sum())
    # per-voxel mean lies inside its cell (up to rounding) and the inverse covariance is symmetric positive
    valid = d["counts"] >= 6
    ijk = np.stack([d["keys"] % info["div_b"][0], (d["keys"] // info["div_b"][0]) % info["div_b"][1],
                    d["keys"] // (info["div_b"][0] * info["div_b"][1])], axis=1) + np.asarray(info["min_b"])
    assert np.all(d["mean"][valid] >= ijk[valid] - 1e-4) and np.all(d["mean"][valid] <= ijk[valid] + 1 + 1e-4)
    assert np.all(np.linalg.eigvalsh(d["icov"][valid]) > 0)
    # rebuilding gives identical bits; the staged and the sharded (3 slices on one GPU) builds agree with it
    g2 = nb.NormalDistributionsTransform()
    g2.setInputTarget(tgt)
    d2 = g2.dump_voxels()
    for k in ("keys", "counts", "mean", "icov"):
        assert np.array_equal(d[k], d2[k])
    # the oracle on the full-size map: exact keys / counts, moments within the bar
    ref = oracle.NormalDistributionsTransform()
    ref.setInputTarget(tgt)
    rl = ref.dump_leaves()
    assert np.array_equal(rl["keys"], d["keys"]) and np.array_equal(rl["counts"], d["counts"])
    assert rel_err(d["mean"], rl["mean"]) < 1e-5
    scale = np.abs(rl["icov"][valid]).reshape(valid.sum(), -1).max(axis=1)
    assert (np.abs(d["icov"][valid] - rl["icov"][valid]).reshape(valid.sum(), -1).max(axis=1) / scale).max() < 1e-5


def test_full_size_derivative_linearity_and_align(nb, c2):
    tgt, src, src2 = c2
    g = nb.NormalDistributionsTransform()
    g.setInputTarget(tgt)
    p = np.array([0.2, -0.1, 0.05, 0.003, -0.002, 0.01])
    # linearity over the source points: the sums of two halves add up to the sums of the whole cloud
    parts = []
    for part in (src, src[: len(src) // 2], src[len(src) // 2:]):
        g.setInputSource(part)
        parts.append(g.eval_derivatives(p, compute_hessian=True))
    whole, a, b = parts
    assert whole["hits"] == a["hits"] + b["hits"]
    # (the point -> thread assignment changes with the split, so the fp32 partial sums round differently: 2e-6)
    assert abs(whole["score"] - (a["score"] + b["score"])) <= 2e-6 * abs(whole["score"])
    assert rel_err(a["gradient"] + b["gradient"], whole["gradient"]) < 2e-6
    assert rel_err(a["hessian"] + b["hessian"], whole["hessian"]) < 2e-6
    # the oracle on the full-size pair
    ref = oracle.NormalDistributionsTransform()
    ref.setInputTarget(tgt); ref.setInputSource(src)
    e = ref.eval_derivatives(p, compute_hessian=True)
    assert whole["hits"] == e["hits"]
    assert rel_err(whole["gradient"], e["gradient"]) < 1e-5 and rel_err(whole["hessian"], e["hessian"]) < 1e-5
    g.setInputSource(src)
    g.align(); ref.align()
    r, rr = g.result(), ref.result()
    assert r["iterations"] == rr["iterations"] and r["n_evaluations"] == rr["n_evaluations"]
    dt, dr = transform_delta(r["final"], rr["final"])
    assert dt < 1e-4 and dr < 1e-4
    assert abs(g.getFitnessScore() - ref.getFitnessScore()) <= 1e-6 * ref.getFitnessScore()
    # idempotence: aligning again from the solution takes the minimum number of Newton iterations and stays put
    first = r["final"].copy()
    g.align(first)
    r2 = g.result()
    dt, dr = transform_delta(r2["final"], first)
    assert r2["iterations"] <= 3 and dt < 0.1 and dr < 0.01
    # repeated align() returns identical bits (apps/align.cpp:25-27), also in the throughput shape and in a batch
    g.align(); b1 = g.result()["final"].copy()
    g.align(); assert np.array_equal(g.result()["final"], b1)
    h2 = nb.NormalDistributionsTransform()
    h2.setInputTarget(tgt); h2.setInputSource(src2)
    res = nb.align_batch([g, h2])
    dt, dr = transform_delta(res[0]["final"], b1)
    assert dt < 1e-6 and dr < 1e-6 and res[1]["converged"]


def test_build_4m_points_properties(nb):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import build_bench
    m = 4_000_000
    pts = build_bench.surface_points(m, 20260104)
    g = nb.NormalDistributionsTransform()
    assert g.set_target_device(pts.data_ptr(), m) == 0
    d = g.dump_voxels()
    assert np.all(np.diff(d["keys"]) > 0)
    assert int(d["counts"][d["counts"] > 0].sum()) == m                     # no rejected leaves here: counts add up
    keys = g.point_keys()
    sel = np.random.default_rng(0).choice(m, size=20000, replace=False)     # spot-check keys against the host formula
    info = g.map_info()
    host = pts[torch.as_tensor(sel, device=pts.device)].cpu().numpy()[:, :3]
    ijk = (np.floor(host * np.float32(1.0)) - np.asarray(info["min_b"], dtype=np.float32)).astype(np.int64)
    exp = ijk[:, 0] + ijk[:, 1] * info["div_b"][0] + ijk[:, 2] * info["div_b"][0] * info["div_b"][1]
    assert np.array_equal(keys[sel], exp)


def test_build_10m_points_against_oracle(nb):
    """BASELINE configs[4] (C5) at 10 M points against the oracle: keys, counts, rejected-leaf and inflation flags exact,
    means / covariances / inverse covariances within the parity bar.  10 M points is above the payload-sort threshold,
    so this is also the at-size check of that path (onesweep_payload_kernel + the sequential moments)."""
    import torch
    import oracle
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import build_bench
    from test_gpu_parity import check_voxels
    m = 10_000_000
    pts = build_bench.surface_points(m, 20260104)
    host = np.ascontiguousarray(pts[:, :3].cpu().numpy())
    g = nb.NormalDistributionsTransform()
    assert g.set_target_device(pts.data_ptr(), m) == 0
    ref = oracle.NormalDistributionsTransform()
    assert ref.setInputTarget(host) == 0
    check_voxels(ref, g)
    # the per-point keys of a 1 M-point slice, input order (the whole array would be 40 MB through the dump path twice)
    sel = slice(3_000_000, 4_000_000)
    assert np.array_equal(ref.point_keys()[sel], g.point_keys()[sel])
