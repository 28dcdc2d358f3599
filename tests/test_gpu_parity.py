"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Tolerances (BASELINE.json north_star): voxel keys / point-to-voxel assignment / per-voxel counts bit-exact;
means, covariances, score, gradient, Hessian within 1e-5 relative (of the max-magnitude entry);
final transform within 1e-4 m translation and 1e-4 rad rotation.
"""
import numpy as np
import pytest

import oracle
from util import golden, load_pair, pose_matrix, rel_err, synthetic_scene, transform_delta

pytestmark = pytest.mark.gpu

REL = 1e-5
TRANS_TOL = 1e-4
ROT_TOL = 1e-4


@pytest.fixture(scope="module")
def nb():
    import toyslam_b200
    return toyslam_b200


def make_pair(nb, tgt, src, method=oracle.DIRECT7, res=1.0, **kw):
    ref = oracle.NormalDistributionsTransform()
    gpu = nb.NormalDistributionsTransform()
    for o in (ref, gpu):
        o.setResolution(res)
        o.setNeighborhoodSearchMethod(method)
        if "eps" in kw:
            o.setTransformationEpsilon(kw["eps"])
        if "max_iter" in kw:
            o.setMaximumIterations(kw["max_iter"])
        if "step" in kw:
            o.setStepSize(kw["step"])
    st_ref = ref.setInputTarget(tgt)
    st_gpu = gpu.setInputTarget(tgt)
    assert st_ref == st_gpu
    ref.setInputSource(src)
    gpu.setInputSource(src)
    return ref, gpu


def check_map(ref, gpu):
    assert np.array_equal(ref.point_keys(), gpu.point_keys())  # bit-exact keys, input order
    check_voxels(ref, gpu)


def check_voxels(ref, gpu):
    ri, gi = ref.map_info(), gpu.map_info()
    for k in ("min_b", "max_b", "div_b"):
        assert np.array_equal(ri[k], gi[k]), k
    assert ri["n_voxels"] == gi["n_voxels"]
    assert ri["n_valid"] == gi["n_valid"]
    rl, gl = ref.dump_leaves(), gpu.dump_voxels()
    assert np.array_equal(rl["keys"], gl["keys"])
    assert np.array_equal(rl["counts"], gl["counts"])      # incl. -1 flags
    assert np.array_equal(rl["inflated"], gl["inflated"])
    valid = rl["counts"] >= 6
    assert rel_err(gl["mean"], rl["mean"]) < REL
    # covariances / inverse covariances per voxel, relative to that voxel's largest entry
    for name in ("cov", "icov"):
        a, b = gl[name][valid], rl[name][valid]
        scale = np.abs(b).reshape(len(b), -1).max(axis=1)
        err = np.abs(a - b).reshape(len(b), -1).max(axis=1) / scale
        assert err.max() < REL, (name, err.max())


@pytest.mark.parametrize("res", [1.0, 0.5, 2.0, 0.3])
def test_map_build_bundled(nb, res):
    tgt, src = load_pair()
    ref, gpu = make_pair(nb, tgt, src, res=res)
    check_map(ref, gpu)


def test_map_build_c1_known_values(nb):
    tgt, src = load_pair()
    gpu = nb.NormalDistributionsTransform()
    gpu.setInputTarget(tgt)
    info = gpu.map_info()
    assert list(info["min_b"]) == [-24, -75, -3] and list(info["max_b"]) == [19, 8, 10]
    assert info["n_voxels"] == 1098 and info["n_valid"] == 599


@pytest.mark.parametrize("offset", [(0, 0, 0), (2000.0, -1500.0, 50.0), (20000.0, -15000.0, 50.0)])
def test_map_build_synthetic_offsets(nb, offset):
    tgt, src = synthetic_scene(offset=offset, seed=3)
    ref, gpu = make_pair(nb, tgt, src)
    check_map(ref, gpu)


def test_map_build_degenerate_voxels(nb):
    rng = np.random.default_rng(5)
    pts = []
    for n, c in ((5, (0.5, 0.5, 0.5)), (6, (1.5, 0.5, 0.5)), (7, (2.5, 0.5, 0.5))):
        pts.append(np.asarray(c) + rng.uniform(-0.3, 0.3, size=(n, 3)))
    t = rng.uniform(-0.4, 0.4, size=(20, 1))
    pts.append(np.array([4.5, 0.5, 0.5]) + t * np.array([1.0, 0.5, 0.2]))             # collinear
    uv = rng.uniform(-0.4, 0.4, size=(30, 2))
    pts.append(np.array([6.5, 0.5, 0.5]) + np.c_[uv, np.zeros(30)])                   # coplanar
    pts.append(np.repeat(np.array([[8.5, 0.5, 0.5]]), 12, axis=0))                    # duplicated point
    pts.append(np.array([[-3.0, -2.0, -1.0], [12.0, 3.0, 2.0]]))                      # bbox corners
    tgt = np.concatenate(pts).astype(np.float32)
    ref, gpu = make_pair(nb, tgt, tgt[:10])
    check_map(ref, gpu)
    # Q1 adds I/n to every covariance, so at 1 m no voxel needs inflation; at a 10 m leaf thin structures do
    # (eigenvalue inflation branch, voxel_grid_covariance_omp_impl.hpp:345-356)
    t = rng.uniform(-4, 4, size=(60, 1))
    line = np.array([5.0, 5.0, 5.0]) + t * np.array([1.0, 0.5, 0.2])
    uv = rng.uniform(-4, 4, size=(80, 2))
    plane = np.array([15.0, 5.0, 5.0]) + np.c_[uv, 0.01 * rng.normal(size=80)]
    blob = np.array([25.0, 5.0, 5.0]) + rng.uniform(-4, 4, size=(40, 3))
    tgt = np.concatenate([line, plane, blob]).astype(np.float32)
    ref, gpu = make_pair(nb, tgt, tgt[:10], res=10.0)
    check_map(ref, gpu)
    assert gpu.dump_voxels()["inflated"].sum() >= 2


def test_map_build_non_dense_and_faces(nb):
    rng = np.random.default_rng(7)
    tgt = rng.uniform(-8, 8, size=(20000, 3)).astype(np.float32)
    tgt[::7] = np.round(tgt[::7])          # points exactly on voxel faces
    tgt[5] = [np.nan, 0, 0]
    tgt[100] = [0, np.inf, 0]
    ref = oracle.NormalDistributionsTransform()
    gpu = nb.NormalDistributionsTransform()
    assert ref.setInputTarget(tgt, is_dense=False) == gpu.setInputTarget(tgt, is_dense=False) == 0
    ref.setInputSource(tgt[:10]); gpu.setInputSource(tgt[:10])
    check_map(ref, gpu)


@pytest.mark.parametrize("slices", [1, 3])
def test_map_build_from_partials(nb, slices):
    """The sharded build's pieces on one GPU: the cloud cut into contiguous slices, per-slice partials with the
    common grid, merged in slice order — keys / counts exact, moments within the bar, same lookups and derivatives."""
    import torch
    from toyslam_b200.sharding import point_range
    tgt, src = load_pair()
    ref, single = make_pair(nb, tgt, src)
    dev = torch.device("cuda", 0)
    parts, boxes, nf_total = [], [], 0
    workers = []
    for r in range(slices):
        lo, hi = point_range(len(tgt), r, slices)
        loc = torch.ones((hi - lo, 4), dtype=torch.float32, device=dev)
        loc[:, :3] = torch.as_tensor(tgt[lo:hi]).to(dev)
        w = nb.NormalDistributionsTransform()
        mn, mx, nf = w.cloud_bounds(loc.data_ptr(), hi - lo)
        boxes.append((mn, mx)); nf_total += nf
        workers.append((w, loc))
    gmin = np.min([b[0] for b in boxes], axis=0)
    gmax = np.max([b[1] for b in boxes], axis=0)
    for w, loc in workers:
        st, nv = w.build_partials(gmin, gmax)
        assert st == 0 and nv > 0
        k = torch.zeros(nv, dtype=torch.int32, device=dev)
        c = torch.zeros(nv, dtype=torch.int32, device=dev)
        m = torch.zeros((nv, 9), dtype=torch.float64, device=dev)
        w.copy_partials(k.data_ptr(), c.data_ptr(), m.data_ptr())
        parts.append((k, c, m))
    keys = torch.cat([p[0] for p in parts]).contiguous()
    cnts = torch.cat([p[1] for p in parts]).contiguous()
    moms = torch.cat([p[2] for p in parts]).contiguous()
    assert int(cnts.sum().item()) == len(tgt)
    merged = nb.NormalDistributionsTransform()
    assert merged.build_from_partials(gmin, gmax, nf_total, keys.data_ptr(), cnts.data_ptr(), moms.data_ptr(), len(keys)) == 0
    check_voxels(ref, merged)
    merged.setInputSource(src)
    p = np.array([0.4, 0.1, -0.02, 0.005, -0.001, -0.01])
    a, b = merged.eval_derivatives(p), ref.eval_derivatives(p)
    assert a["hits"] == b["hits"]
    assert rel_err(a["gradient"], b["gradient"]) < REL and rel_err(a["hessian"], b["hessian"]) < REL
    with pytest.raises(nb.NdtError):
        merged.getFitnessScore()       # the merged handle holds no raw target


@pytest.mark.parametrize("leaf", [0.1, 0.3, 0.5, 1.0])
def test_voxelgrid_filter_bit_exact(nb, leaf):
    """pcl::VoxelGrid centroid downsample on the device (apps/align.cpp:57-69): bit-identical to the oracle's
    restatement (which is pinned bit-for-bit to the committed 0.1 m fixtures), same cell order."""
    tgt, src = load_pair()
    rng = np.random.default_rng(3)
    dense = (np.repeat(src[:6000], 5, axis=0) + rng.normal(0, 0.04, size=(30000, 3))).astype(np.float32)
    scene, _ = synthetic_scene(n_target=60000, seed=2, offset=(2000.0, -1500.0, 50.0))
    gpu = nb.NormalDistributionsTransform()
    for cloud in (dense, scene, tgt[:1], np.zeros((0, 3), np.float32)):
        exp = oracle.voxelgrid_downsample(cloud, leaf) if len(cloud) else np.zeros((0, 3), np.float32)
        got = gpu.voxelgrid_filter(cloud, leaf)
        assert got.shape == exp[:, :3].shape
        assert np.array_equal(got, exp[:, :3])
    withnan = dense.copy()
    withnan[::97, 1] = np.nan
    assert np.array_equal(gpu.voxelgrid_filter(withnan, leaf), oracle.voxelgrid_downsample(withnan, leaf)[:, :3])
    # the filter leaves the object's own map alone
    gpu.setInputTarget(tgt); gpu.setInputSource(src)
    before = gpu.dump_voxels()["keys"].copy()
    gpu.voxelgrid_filter(dense, leaf)
    assert np.array_equal(gpu.dump_voxels()["keys"], before)


def test_voxelgrid_filter_reproduces_config1_fixture(nb):
    """Config 1 end to end on the device: raw head of the bundled scan -> 0.1 m VoxelGrid equals the fixture pipeline."""
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "raw_head.npz")
    d = np.load(path)
    gpu = nb.NormalDistributionsTransform()
    for k in ("target", "source"):
        raw = np.ascontiguousarray(d[k][:, :3])
        got = gpu.voxelgrid_filter(raw, 0.1)
        assert np.array_equal(got, d[k + "_ds0p1"])      # committed fixture (independent numpy VoxelGrid)
        assert np.array_equal(got, oracle.voxelgrid_downsample(raw, 0.1)[:, :3])


def test_fused_and_staged_builds_identical(nb, monkeypatch):
    """Scan-sized clouds are built by ONE cooperative kernel (small_build.cuh), large ones by the staged streaming
    kernels: same arithmetic in the same order, so the two paths must agree bit for bit — map, downsample and what
    the solver computes from them."""
    rng = np.random.default_rng(12)
    tgt, src = load_pair()
    scene, _ = synthetic_scene(n_target=50000, seed=3, offset=(2000.0, -1500.0, 50.0))
    nan_cloud = tgt.copy()
    nan_cloud[::53, 2] = np.nan
    tiny = (np.array([0.5, 0.5, 0.5]) + rng.uniform(-0.3, 0.3, size=(7, 3))).astype(np.float32)
    cases = [(tgt, 1.0, True), (tgt, 0.5, True), (tgt, 2.0, True), (scene, 1.0, True), (nan_cloud, 1.0, False), (tiny, 1.0, True)]
    for cloud, res, dense in cases:
        out = {}
        for path in ("staged", "fused", "cluster"):
            # "cluster": the fused kernel launched as ONE thread-block cluster with hardware cluster barriers (the flavour
            # handles in throughput mode use) instead of a cooperative launch with a global-memory barrier
            monkeypatch.setenv("NDTB200_BUILD_PATH", "fused" if path == "cluster" else path)
            monkeypatch.setenv("NDTB200_FUSED_LAUNCH", "cluster" if path == "cluster" else "coop")
            g = nb.NormalDistributionsTransform()
            g.setResolution(res)
            st = g.setInputTarget(cloud, dense)
            g.setInputSource(src)
            d = g.dump_voxels()
            e = g.eval_derivatives(np.array([0.3, 0.1, -0.02, 0.004, -0.002, -0.01]))
            out[path] = (st, g.map_info(), g.point_keys(), d, e, g.voxelgrid_filter(cloud, 0.3), g.voxelgrid_filter(cloud, 1.7))
        monkeypatch.delenv("NDTB200_BUILD_PATH")
        monkeypatch.delenv("NDTB200_FUSED_LAUNCH")
        a = out["staged"]
        for other in ("fused", "cluster"):
            b = out[other]
            assert a[0] == b[0]
            for k in ("min_b", "max_b", "div_b"):
                assert np.array_equal(a[1][k], b[1][k])
            assert a[1]["n_voxels"] == b[1]["n_voxels"] and a[1]["n_valid"] == b[1]["n_valid"]
            assert np.array_equal(a[2], b[2])
            for k in ("keys", "counts", "mean", "cov", "icov", "inflated"):
                assert np.array_equal(a[3][k], b[3][k], equal_nan=True), (other, k)
            assert a[4]["score"] == b[4]["score"] and np.array_equal(a[4]["gradient"], b[4]["gradient"]) and np.array_equal(a[4]["hessian"], b[4]["hessian"])
            assert np.array_equal(a[5], b[5], equal_nan=True) and np.array_equal(a[6], b[6], equal_nan=True)
    # the guard and the empty cloud behave the same on both paths
    far = np.array([[0, 0, 0], [4000, 4000, 4000], [1, 1, 1], [2, 2, 2]], dtype=np.float32)
    for path in ("staged", "fused"):
        monkeypatch.setenv("NDTB200_BUILD_PATH", path)
        g = nb.NormalDistributionsTransform()
        g.setResolution(0.5)
        assert g.setInputTarget(far) == 2 and g.map_info()["n_voxels"] == 0      # NDTB200_ERR_GRID_OVERFLOW
    monkeypatch.delenv("NDTB200_BUILD_PATH")


def test_payload_sort_build_identical(nb, monkeypatch):
    """Clouds above 4 M points are sorted with the POINTS as the payload of the radix passes (map_build.cuh:
    onesweep_payload_kernel) instead of (key, index) pairs + a gather: the same stable order and the same additions, so
    the map, the solver's sums and the consumers of the lazily built index array (getFitnessScore, KDTREE) must agree
    bit for bit with the (key, index) path.  NDTB200_PAYLOAD_SORT_MIN forces the path at test sizes."""
    tgt, src = load_pair()
    scene, _ = synthetic_scene(n_target=120000, seed=5, offset=(2000.0, -1500.0, 50.0))
    nan_cloud = tgt.copy()
    nan_cloud[::53, 2] = np.nan
    nan_cloud[7::211, 0] = np.inf
    one_pass = (tgt[:5000] * 0.1).astype(np.float32)                # < 256 cells at res 1.0: a single radix pass
    rng = np.random.default_rng(77)
    centres = np.concatenate([rng.uniform([-150, -150, -100], [150, 150, 100], size=(40, 3)), [[-150.5, -150.5, -100.5], [150.5, 150.5, 100.5]]])
    wide = (rng.uniform(-0.4, 0.4, size=(42, 9, 3)) + centres[:, None, :]).reshape(-1, 3).astype(np.float32)
    assert np.prod(np.floor(wide.max(0)) - np.floor(wide.min(0)) + 1) > 2 ** 24   # 302 x 302 x 202 cells at res 1.0: four radix passes
    cases = [(tgt, 1.0, True), (tgt, 0.5, True), (scene, 1.0, True), (scene, 0.3, True), (nan_cloud, 1.0, False), (one_pass, 1.0, True),
             (wide, 1.0, True)]
    monkeypatch.setenv("NDTB200_BUILD_PATH", "staged")
    for cloud, res, dense in cases:
        out = {}
        for path, lo in (("pairs", str(1 << 40)), ("payload", "1")):
            monkeypatch.setenv("NDTB200_PAYLOAD_SORT_MIN", lo)
            g = nb.NormalDistributionsTransform()
            g.setResolution(res)
            st = g.setInputTarget(cloud, dense)
            g.setInputSource(src)
            e = g.eval_derivatives(np.array([0.3, 0.1, -0.02, 0.004, -0.002, -0.01]))
            d = g.dump_voxels()                 # payload path: moments recomputed from the sorted cloud (sequential kernel)
            g.align()
            fit = g.getFitnessScore()           # payload path: builds the sorted index array on first use
            d2 = g.dump_voxels()                # ... after which the moments come from the gather again
            g.setNeighborhoodSearchMethod(oracle.KDTREE)
            ek = g.eval_derivatives(np.array([0.1, 0.0, 0.02, 0.0, 0.001, 0.01]), compute_hessian=False)
            out[path] = (st, g.map_info(), d, e, fit, d2, ek)
        a, b = out["pairs"], out["payload"]
        assert a[0] == b[0]
        for k in ("min_b", "max_b", "div_b"):
            assert np.array_equal(a[1][k], b[1][k])
        assert a[1]["n_voxels"] == b[1]["n_voxels"] and a[1]["n_valid"] == b[1]["n_valid"]
        for dd in (2, 5):
            for k in ("keys", "counts", "mean", "cov", "icov", "inflated"):
                assert np.array_equal(a[dd][k], b[dd][k], equal_nan=True), k
        for ee in (3, 6):
            assert a[ee]["score"] == b[ee]["score"] and np.array_equal(a[ee]["gradient"], b[ee]["gradient"])
        assert np.array_equal(a[3]["hessian"], b[3]["hessian"])
        assert a[4] == b[4]
    monkeypatch.delenv("NDTB200_PAYLOAD_SORT_MIN")
    monkeypatch.delenv("NDTB200_BUILD_PATH")


def test_grid_overflow_guard(nb):
    tgt = np.array([[0, 0, 0], [3000, 3000, 3000], [1, 1, 1]], dtype=np.float32)
    ref = oracle.NormalDistributionsTransform()
    gpu = nb.NormalDistributionsTransform()
    for o in (ref, gpu):
        o.setResolution(0.5)
    assert ref.setInputTarget(tgt) == oracle.BUILD_GRID_OVERFLOW
    assert gpu.setInputTarget(tgt) == 2  # NDTB200_ERR_GRID_OVERFLOW
    assert gpu.map_info()["n_voxels"] == 0


@pytest.mark.parametrize("method", [oracle.DIRECT1, oracle.DIRECT7, oracle.DIRECT26])
def test_lookup(nb, method):
    tgt, src = load_pair()
    ref, gpu = make_pair(nb, tgt, src, method=method)
    rng = np.random.default_rng(11)
    q = np.concatenate([src[:4000], rng.uniform(-40, 40, size=(2000, 3)).astype(np.float32),
                        np.round(src[:500])]).astype(np.float32)
    a, b = gpu.lookup(q, method), ref.lookup(q, method)
    assert np.array_equal(a, b)   # same keys in the same (reference offset) order


@pytest.mark.parametrize("method", [oracle.DIRECT1, oracle.DIRECT7, oracle.DIRECT26])
def test_hash_index_fallback(nb, method, monkeypatch):
    """The voxel index has two interchangeable forms (direct-mapped cell table / open-addressing hash for grids whose
    table would not fit): both must return the same voxels and the same derivatives."""
    tgt, src = load_pair()
    monkeypatch.setenv("NDTB200_FORCE_HASH", "1")
    ref, gpu_hash = make_pair(nb, tgt, src, method=method)
    monkeypatch.delenv("NDTB200_FORCE_HASH")
    _, gpu_dense = make_pair(nb, tgt, src, method=method)
    q = np.concatenate([src[:4000], np.round(src[:500])]).astype(np.float32)
    assert np.array_equal(gpu_hash.lookup(q, method), ref.lookup(q, method))
    assert np.array_equal(gpu_dense.lookup(q, method), ref.lookup(q, method))
    p = np.array([0.4, 0.1, -0.02, 0.005, -0.001, -0.01])
    a, b = gpu_hash.eval_derivatives(p, compute_hessian=True), gpu_dense.eval_derivatives(p, compute_hessian=True)
    assert a["score"] == b["score"] and np.array_equal(a["gradient"], b["gradient"]) and np.array_equal(a["hessian"], b["hessian"])
    gpu_hash.align()
    gpu_dense.align()
    assert np.array_equal(gpu_hash.result()["final"], gpu_dense.result()["final"])


POSES = [np.zeros(6),
         np.array([0.4, 0.1, -0.02, 0.005, -0.001, -0.01]),
         np.array([0.1, -0.3, 0.05, 0.3, -0.25, 0.6]),        # large angles: exposes Q2 (+sy / -sy)
         np.array([-0.2, 0.2, 0.0, 5e-5, -5e-5, 2e-5]),       # small-angle snap (|a| < 1e-4)
         np.array([1.0, 0.5, 0.1, -0.05, 0.08, -1.2])]


@pytest.mark.parametrize("method", [oracle.DIRECT1, oracle.DIRECT7, oracle.DIRECT26])
@pytest.mark.parametrize("hess", [True, False])
def test_derivatives(nb, method, hess):
    tgt, src = load_pair()
    ref, gpu = make_pair(nb, tgt, src, method=method)
    for p in POSES:
        a = gpu.eval_derivatives(p, compute_hessian=hess)
        b = ref.eval_derivatives(p, compute_hessian=hess)
        assert a["hits"] == b["hits"], (p, a["hits"], b["hits"])
        assert abs(a["score"] - b["score"]) <= REL * abs(b["score"])
        assert rel_err(a["gradient"], b["gradient"]) < REL
        if hess:
            assert rel_err(a["hessian"], b["hessian"]) < REL
        else:
            assert np.all(a["hessian"] == 0) and np.all(b["hessian"] == 0)


@pytest.mark.parametrize("method", [oracle.DIRECT1, oracle.DIRECT7, oracle.DIRECT26])
def test_hessian_only(nb, method):
    tgt, src = load_pair()
    ref, gpu = make_pair(nb, tgt, src, method=method)
    for p in POSES:
        assert rel_err(gpu.eval_hessian(p), ref.eval_hessian(p)) < REL


def test_derivatives_far_from_origin(nb):
    """fp64 voxel means: parity must hold at map-scale coordinates (SURVEY §7 hard part 2)."""
    for off in ((2000.0, -1500.0, 50.0), (20000.0, -15000.0, 50.0)):
        tgt, src = synthetic_scene(offset=off, seed=9)
        ref, gpu = make_pair(nb, tgt, src)
        p = np.array([off[0] + 0.2, off[1] - 0.1, off[2] + 0.02, 0.003, -0.002, 0.01])
        a, b = gpu.eval_derivatives(p), ref.eval_derivatives(p)
        assert a["hits"] == b["hits"] and b["hits"] > 1000
        assert abs(a["score"] - b["score"]) <= REL * abs(b["score"])
        assert rel_err(a["gradient"], b["gradient"]) < REL
        assert rel_err(a["hessian"], b["hessian"]) < REL


def check_align(ref, gpu, guess=None):
    ref.align(guess)
    out = gpu.align(guess)
    rr, rg = ref.result(), gpu.result()
    tr, tg = ref.trace(), gpu.trace()
    assert rg["iterations"] == rr["iterations"]
    assert rg["n_evaluations"] == rr["n_evaluations"]
    assert rg["n_hessian_passes"] == rr["n_hessian_passes"]
    assert rg["converged"] == rr["converged"]
    assert np.array_equal(tg["kind"], tr["kind"])
    assert np.abs(tg["a_t"] - tr["a_t"]).max() < 1e-6        # trial step-length sequence
    assert np.abs(tg["x"] - tr["x"]).max() < 1e-6            # evaluated poses
    dt, dr = transform_delta(rg["final"], rr["final"])
    assert dt < TRANS_TOL and dr < ROT_TOL, (dt, dr)
    assert abs(rg["trans_probability"] - rr["trans_probability"]) <= 1e-5 * abs(rr["trans_probability"])
    return out, rr, rg


@pytest.mark.parametrize("method,name", [(oracle.DIRECT7, "DIRECT7"), (oracle.DIRECT1, "DIRECT1")])
def test_align_config1_golden_fitness(nb, method, name):
    """Config 1: the reference's own published known answers (ndt_omp/README.md:26,31)."""
    tgt, src = load_pair()
    ref, gpu = make_pair(nb, tgt, src, method=method)
    out, rr, rg = check_align(ref, gpu)
    fit = gpu.getFitnessScore()
    assert abs(fit - golden()["fitness"][name]) < 5e-7, fit           # 6 printed digits
    assert abs(fit - ref.getFitnessScore()) <= 1e-6 * fit
    # output cloud = source transformed by the final pose
    exp = oracle.transform_points(rg["final"], src)
    assert np.array_equal(out[:, :3], exp[:, :3])
    # repeated align() on the same pair returns identical results (apps/align.cpp:25-27)
    gpu.align()
    assert np.array_equal(gpu.result()["final"], rg["final"])


@pytest.mark.parametrize("method", [oracle.DIRECT7, oracle.DIRECT1, oracle.DIRECT26])
def test_align_throughput_shape(nb, method):
    """The small-CTA (throughput) shape of the solve kernel must meet the same parity bar as the default shape."""
    tgt, src = load_pair("pair_ds0p3.npz")
    ref, gpu = make_pair(nb, tgt, src, method=method, eps=0.01, max_iter=64)
    gpu.set_throughput_mode(True)
    check_align(ref, gpu)
    for p in POSES[:3]:
        a, b = gpu.eval_derivatives(p, compute_hessian=True), ref.eval_derivatives(p, compute_hessian=True)
        assert rel_err(a["gradient"], b["gradient"]) < REL and rel_err(a["hessian"], b["hessian"]) < REL
        assert rel_err(gpu.eval_hessian(p), ref.eval_hessian(p)) < REL


def test_align_batch_independent_pairs(nb):
    """ndtb200_align_batch: several independent scan pairs in flight together give each pair the result its own
    align() gives (oracle parity per pair; pairs differ in source, perturbation and guess)."""
    tgt, src = load_pair("pair_ds0p3.npz")
    rng = np.random.default_rng(5)
    refs, gpus, guesses = [], [], []
    for i in range(6):
        p = np.concatenate([rng.uniform(-0.3, 0.3, 3), rng.uniform(-0.02, 0.02, 3)])
        s_i = oracle.transform_points(pose_matrix(p), src)[:, :3].astype(np.float32)
        t_i = tgt if i % 2 == 0 else (tgt + np.float32(0.37 * i)).astype(np.float32)
        s_i = s_i if i % 2 == 0 else (s_i + np.float32(0.37 * i)).astype(np.float32)
        ref, gpu = make_pair(nb, t_i, s_i, eps=0.01, max_iter=64)
        refs.append(ref); gpus.append(gpu)
        guesses.append(pose_matrix(np.array([0.05 * i, 0, 0, 0, 0, 0.002 * i])) if i >= 3 else np.eye(4))
    results = nb.align_batch(gpus, guesses)
    assert len(results) == 6
    for ref, gpu, g, rg in zip(refs, gpus, guesses, results):
        ref.align(g)
        rr = ref.result()
        assert rg["iterations"] == rr["iterations"] and rg["n_evaluations"] == rr["n_evaluations"]
        assert rg["n_hessian_passes"] == rr["n_hessian_passes"] and rg["converged"] == rr["converged"]
        dt, dr = transform_delta(rg["final"], rr["final"])
        assert dt < TRANS_TOL and dr < ROT_TOL, (dt, dr)
    # the default (latency) shape is back after the batch call
    gpus[0].align(guesses[0])
    dt, dr = transform_delta(gpus[0].result()["final"], results[0]["final"])
    assert dt < 1e-6 and dr < 1e-6


def test_align_direct26(nb):
    tgt, src = load_pair()
    ref, gpu = make_pair(nb, tgt, src, method=oracle.DIRECT26)
    _, rr, rg = check_align(ref, gpu)
    assert rr["n_hessian_passes"] >= 1   # exercises the Hessian-only pass inside the solver


def test_align_line_search_fixture_b(nb):
    """0.3 m downsample + mapping-node parameters: 10 extra More-Thuente trials."""
    tgt, src = load_pair("pair_ds0p3.npz")
    ref, gpu = make_pair(nb, tgt, src, eps=0.01, max_iter=64)
    _, rr, rg = check_align(ref, gpu)
    assert rr["n_evaluations"] > rr["iterations"] + 1


@pytest.mark.parametrize("guess_p", [[0.3, 0.1, 0.0, 0.0, 0.0, 0.0],
                                     [0.2, 0.05, -0.01, 0.01, -0.02, 0.03],
                                     [0.2, 0.05, -0.01, -0.01, 0.02, -0.03]])   # negative roll: Q5
def test_align_with_guess(nb, guess_p):
    tgt, src = load_pair()
    ref, gpu = make_pair(nb, tgt, src)
    check_align(ref, gpu, guess=pose_matrix(guess_p))


def test_align_scan_to_map_offset(nb):
    off = (2000.0, -1500.0, 50.0)
    tgt, src = synthetic_scene(offset=off, seed=21)
    ref, gpu = make_pair(nb, tgt, src)
    check_align(ref, gpu, guess=pose_matrix([off[0], off[1], off[2], 0, 0, 0]))


def test_fitness_grid_search_equals_brute_force(nb, monkeypatch):
    """getFitnessScore's grid-accelerated nearest-neighbour search is exact: same value as the O(N*M) scan and as the
    oracle, also when queries fall outside the target's grid (far guess) and at map-scale coordinates."""
    cases = [load_pair(), synthetic_scene(offset=(2000.0, -1500.0, 50.0), seed=6)]
    for ci, (tgt, src) in enumerate(cases):
        ref, gpu = make_pair(nb, tgt, src)
        guesses = [None, pose_matrix(np.array([3.0, -2.0, 0.5, 0.02, -0.01, 0.3])), pose_matrix(np.array([60.0, 0, 0, 0, 0, 0]))]
        if ci == 1:
            guesses = [pose_matrix(np.array([2000.0, -1500.0, 50.0, 0, 0, 0]))]
        for g in guesses:
            gpu.setMaximumIterations(1); ref.setMaximumIterations(1)
            gpu.align(g); ref.align(g)
            monkeypatch.delenv("NDTB200_FITNESS_BRUTE", raising=False)
            fast = gpu.getFitnessScore()
            fast_r = gpu.getFitnessScore(2.0)
            monkeypatch.setenv("NDTB200_FITNESS_BRUTE", "1")
            slow = gpu.getFitnessScore()
            slow_r = gpu.getFitnessScore(2.0)
            monkeypatch.delenv("NDTB200_FITNESS_BRUTE")
            assert fast == slow and fast_r == slow_r          # identical fp32 minima, identical sums
            assert abs(fast - ref.getFitnessScore()) <= 1e-6 * abs(fast)
            assert abs(fast_r - ref.getFitnessScore(2.0)) <= 1e-6 * abs(fast_r)


def test_kdtree_mode(nb):
    """The fourth search method (ndt_omp.h:52-57, radius search over the voxel centroids): derivatives, Hessian-only
    pass, calculateScore and the whole align against the oracle, and the README's KDTREE fitness 0.213937."""
    tgt, src = load_pair()
    ref, gpu = make_pair(nb, tgt, src, method=oracle.KDTREE)
    for ip, p in enumerate(POSES):
        a, b = gpu.eval_derivatives(p, compute_hessian=True), ref.eval_derivatives(p, compute_hessian=True)
        assert a["hits"] == b["hits"] and b["hits"] > 0, (ip, p.tolist(), a["hits"], b["hits"], gpu.map_info(), ref.map_info(),
                                                          float(np.nanmin(tgt)), float(np.nanmax(src)))
        assert abs(a["score"] - b["score"]) <= REL * abs(b["score"])
        assert rel_err(a["gradient"], b["gradient"]) < REL and rel_err(a["hessian"], b["hessian"]) < REL
        assert rel_err(gpu.eval_hessian(p), ref.eval_hessian(p)) < REL
    out, rr, rg = check_align(ref, gpu)
    fit = gpu.getFitnessScore()
    assert abs(fit - golden()["fitness"]["KDTREE"]) < 5e-7, fit
    a, b = gpu.calculateScore(src), ref.calculateScore(src)
    assert abs(a - b) <= 1e-9 * abs(b)
    # node parameters + the throughput shape
    tgt, src = load_pair("pair_ds0p3.npz")
    ref, gpu = make_pair(nb, tgt, src, method=oracle.KDTREE, eps=0.01, max_iter=64)
    gpu.set_throughput_mode(True)
    check_align(ref, gpu)
    # at map-scale coordinates
    tgt, src = synthetic_scene(offset=(2000.0, -1500.0, 50.0), seed=9)
    ref, gpu = make_pair(nb, tgt, src, method=oracle.KDTREE)
    p = np.array([2000.2, -1500.1, 50.02, 0.003, -0.002, 0.01])
    a, b = gpu.eval_derivatives(p), ref.eval_derivatives(p)
    assert a["hits"] == b["hits"] and b["hits"] > 1000
    assert rel_err(a["gradient"], b["gradient"]) < REL and rel_err(a["hessian"], b["hessian"]) < REL


def test_mapping_pipeline_matches_reference_loop(nb):
    """ndt_rosbag_mapping_node.cpp:42-161 as a device-resident pipeline (ndtb200_mapper_*): per-step downsample sizes
    exact, registrations equal to the oracle's loop (same counts, transforms within the bar), accumulated pose and
    global map consistent."""
    import workloads
    from util import oracle_mapping_loop
    scans, poses = workloads.config3_sequence(6, azimuth_steps=600, leaf=0.05)   # dense-ish raw scans (sensor frame)
    ref_steps, ref_map = oracle_mapping_loop(scans)
    mapper = nb.Mapper()
    for k, cloud in enumerate(scans):
        s, r = mapper.push_scan(cloud), ref_steps[k]
        assert s["n_filtered"] == r["n_filtered"]                       # VoxelGrid is bit-exact
        assert s["converged"] == r["converged"] and s["iterations"] == r["iterations"] and s["n_evaluations"] == r["n_evaluations"]
        dt, dr = transform_delta(s["transform"], r["transform"])
        assert dt < TRANS_TOL and dr < ROT_TOL, (k, dt, dr)
        dt, dr = transform_delta(s["pose"], r["pose"])
        assert dt < 10 * TRANS_TOL and dr < 10 * ROT_TOL, (k, dt, dr)   # products of k step transforms
        if k > 0:
            assert abs(s["fitness"] - r["fitness"]) <= 1e-5 * abs(r["fitness"])
        # poses differ by ~1e-7, so a few points may fall into the neighbouring map cell: sizes agree to 0.2 %
        assert abs(s["n_map"] - r["n_map"]) <= max(2, 0.002 * r["n_map"]), (k, s["n_map"], r["n_map"])
    gm = mapper.global_map()
    assert len(gm) == s["n_map"]
    # same map up to the cells touched by the ~1e-7 pose differences: compare the occupied 0.5 m cells
    cells = lambda c: set(map(tuple, np.floor(c / 0.5).astype(np.int64)))
    a, b = cells(gm), cells(ref_map)
    assert len(a ^ b) <= max(4, 0.004 * len(b))
    assert mapper.launch_count() > 0


def test_fitness_sums(nb):
    """ndtb200_fitness_sums returns the two sums behind getFitnessScore (what the ranks of a sharded source all-reduce;
    the 2-GPU test checks the all-reduced value against the oracle)."""
    tgt, src = load_pair()
    ref, gpu = make_pair(nb, tgt, src)
    gpu.align(); ref.align()
    s, c = gpu.fitness_sums()
    assert c == len(src)
    assert abs(gpu.getFitnessScore() - s / c) <= 1e-12 * s / c
    assert abs(s / c - ref.getFitnessScore()) <= 1e-6 * s / c
    s2, c2 = gpu.fitness_sums(0.05)                      # max_range is compared with the squared distance (pcl::Registration)
    assert 0 < c2 < c and abs(s2 / c2 - ref.getFitnessScore(0.05)) <= 1e-6 * s2 / c2


def test_calculate_score(nb):
    tgt, src = load_pair()
    ref, gpu = make_pair(nb, tgt, src)
    a, b = gpu.calculateScore(src), ref.calculateScore(src)
    assert abs(a - b) <= 1e-9 * abs(b)


def test_clone_is_independent(nb):
    tgt, src = load_pair("pair_ds0p3.npz")
    gpu = nb.NormalDistributionsTransform()
    gpu.setInputTarget(tgt); gpu.setInputSource(src)
    gpu.align()
    r0 = gpu.result()
    other = gpu.clone()
    other.align()
    assert np.array_equal(other.result()["final"], r0["final"])
    other.setInputSource(src[:100])
    gpu.align()
    assert np.array_equal(gpu.result()["final"], r0["final"])


def test_point_strides(nb):
    """PointXYZI / PointXYZRGB are 32-byte records (ndt_omp.cpp:4-6): same answer as 16-byte PointXYZ."""
    tgt, src = load_pair("pair_ds0p3.npz")
    a = nb.NormalDistributionsTransform()
    a.setInputTarget(tgt); a.setInputSource(src); a.align()
    t32 = np.zeros((len(tgt), 8), dtype=np.float32); t32[:, :3] = tgt; t32[:, 4] = 7.0
    s32 = np.zeros((len(src), 8), dtype=np.float32); s32[:, :3] = src
    b = nb.NormalDistributionsTransform()
    b.set_target_raw(t32.ctypes.data, len(tgt), 32)
    b.set_source_raw(s32.ctypes.data, len(src), 32)
    out = np.zeros((len(src), 8), dtype=np.float32)
    b.align_raw(None, out.ctypes.data, 32)
    assert np.array_equal(a.result()["final"], b.result()["final"])
    assert np.all(out[:, 3] == 1.0) and np.all(out[:, 4:] == 0.0)


def test_cpp_shim_app_reproduces_golden_fitness(nb, tmp_path):
    """apps/align_b200 (the reference's apps/align.cpp NDT section over the header-only C++ shim) prints the
    reference's published fitness values (ndt_omp/README.md:26,31)."""
    import os
    import re
    import subprocess
    from toyslam_b200 import _build
    app = _build.build_apps()
    tgt, src = load_pair()
    tp, sp = str(tmp_path / "t.bin"), str(tmp_path / "s.bin")
    np.ascontiguousarray(tgt, dtype=np.float32).tofile(tp)
    np.ascontiguousarray(src, dtype=np.float32).tofile(sp)
    out = subprocess.run([os.path.abspath(app), tp, sp], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    fit = re.findall(r"fitness: ([0-9.]+)", out.stdout)
    assert len(fit) == 3, out.stdout
    assert abs(float(fit[0]) - golden()["fitness"]["DIRECT7"]) < 1e-6
    assert abs(float(fit[1]) - golden()["fitness"]["DIRECT1"]) < 1e-6
    assert "copy converged: 1, iterations 5" in out.stdout
    assert re.search(r"batch of 4: .*iterations 5 5 5 5", out.stdout), out.stdout      # alignBatch through the shim
    mm = re.search(r"mapper: 2 scans, converged (\d), iterations (\d+), filtered (\d+) pts, map (\d+) pts", out.stdout)
    assert mm and mm.group(1) == "1" and int(mm.group(3)) == len(oracle.voxelgrid_downsample(src, 0.3)) and int(mm.group(4)) > 0


def test_cpp_shim_pcd_voxelgrid_pipeline(nb, tmp_path):
    """The reference app's front end through the C++ shim: binary PCD files in (pcl::io::loadPCDFile), 0.1 m
    pcl::VoxelGrid (apps/align.cpp:57-69) on the device, align, aligned cloud written back as a binary PCD."""
    import os
    import re
    import subprocess
    from toyslam_b200 import _build
    app = _build.build_apps()
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "raw_head.npz"))

    def write_pcd(path, xyz):   # the layout of ndt_omp/data/*.pcd: x y z intensity, float32, DATA binary
        rec = np.zeros((len(xyz), 4), dtype=np.float32)
        rec[:, :3] = xyz
        with open(path, "wb") as f:
            f.write(("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\n"
                     "COUNT 1 1 1 1\nWIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n" % (len(xyz), len(xyz))).encode())
            f.write(rec.tobytes())

    tp, sp, op = str(tmp_path / "t.pcd"), str(tmp_path / "s.pcd"), str(tmp_path / "aligned.pcd")
    write_pcd(tp, d["target"]); write_pcd(sp, d["source"])
    out = subprocess.run([os.path.abspath(app), tp, sp, "--leaf", "0.1", "--save-aligned", op], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "target 8192 pts, source 8192 pts" in out.stdout
    assert "downsampled (0.1 m): target %d pts, source %d pts" % (len(d["target_ds0p1"]), len(d["source_ds0p1"])) in out.stdout
    fit = [float(x) for x in re.findall(r"fitness: ([0-9.]+)", out.stdout)]
    ref = oracle.NormalDistributionsTransform()
    ref.setInputTarget(d["target_ds0p1"]); ref.setInputSource(d["source_ds0p1"]); ref.align()
    assert abs(fit[0] - ref.getFitnessScore()) <= 2e-6 * max(1.0, fit[0])       # printed with 6 digits
    assert ("saved %d aligned points" % len(d["source_ds0p1"])) in out.stdout
    assert os.path.getsize(op) > 12 * len(d["source_ds0p1"])


# ---- round 2: independent checks, fixture A, non-finite sources, the setResolution rule, emulated ranks --------------

def test_newton_solve_against_numpy_pinv(nb):
    """The on-device Newton solve (definite elimination / pivoted elimination / SVD pseudo-inverse, ndt_omp_impl.hpp:127-129
    = JacobiSVD(H).solve(-g)) against numpy.linalg.pinv — an independent implementation, not the oracle's twin."""
    from test_independent_checks import _solve_cases, numpy_jacobisvd_solve
    gpu = nb.NormalDistributionsTransform()
    paths = set()
    for H, g, kind in _solve_cases():
        x, path = gpu.newton_solve(H, g)
        paths.add((kind.split()[0][:4], path))
        ref = numpy_jacobisvd_solve(H, -g)
        cond = np.linalg.cond(H) if kind in ("definite", "indefinite") else 1.0
        tol = max(1e-9, 50 * cond * np.finfo(np.float64).eps) if kind in ("definite", "indefinite") else 1e-8
        assert np.abs(x - ref).max() <= tol * max(np.abs(ref).max(), 1e-12) + 1e-14, (kind, path, x, ref)
    assert ("defi", 0) in paths and ("inde", 1) in paths and ("rank", 1) in paths   # every solver path was exercised


@pytest.mark.parametrize("method,kind", [(oracle.DIRECT26, 26), (oracle.DIRECT7, 7), (oracle.DIRECT1, 1)])
def test_neighbourhoods_against_enumeration(nb, method, kind):
    """Q7: DIRECT26 = 3x3x3 minus the centre, DIRECT7 = centre + faces — as SETS against a brute enumeration built from
    the dumped voxel list (independent of the offset tables the oracle and the kernel both carry)."""
    from test_independent_checks import enumerated_neighbours, lookup_queries
    tgt, q = lookup_queries()
    gpu = nb.NormalDistributionsTransform()
    gpu.setInputTarget(tgt)
    info, vox = gpu.map_info(), gpu.dump_voxels()
    valid = set(int(k) for k, c in zip(vox["keys"], vox["counts"]) if c >= 6)
    got = gpu.lookup(q, method)
    for i in range(len(q)):
        exp = enumerated_neighbours(valid, (info["min_b"], info["max_b"], info["div_b"]), q[i], 1.0, kind)
        row = [int(k) for k in got[i] if k >= 0]
        assert len(row) == len(set(row)) and set(row) == exp, (i, q[i], sorted(row), sorted(exp))


def test_align_line_search_fixture_a(nb):
    """SURVEY Appendix A fixture A — the RAW bundled pair: 2 Newton iterations, 23 evaluations (2 x 10 extra More-Thuente
    trials), 2 Hessian-only passes: the heaviest exercise of trialValueSelectionMT / updateIntervalMT / computeHessian."""
    import os
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "pair_raw.npz"))
    ref, gpu = make_pair(nb, d["target"], d["source"])
    assert gpu.map_info()["n_valid"] == golden()["appendix_a"]["fixture_A_raw_pair"]["valid_voxels"] == 690
    _, rr, rg = check_align(ref, gpu)
    ap = golden()["appendix_a"]["fixture_A_raw_pair"]
    assert (rg["iterations"], rg["n_evaluations"], rg["n_hessian_passes"]) == (ap["iterations"], ap["evaluations"], ap["hessian_passes"])
    tr = gpu.trace()
    assert list(tr["kind"]) == [0] + ([0] + [1] * 10 + [2]) * 2
    assert np.abs(tr["x"][-1] - np.array(ap["p"])).max() < 1e-6
    gpu.set_throughput_mode(True)         # same bar for the small-CTA shape
    check_align(ref, gpu)


@pytest.mark.parametrize("method", [oracle.DIRECT7, oracle.DIRECT1, oracle.DIRECT26, oracle.KDTREE])
def test_non_finite_source_points_contribute_nothing(nb, method):
    """a NaN / inf source point has no neighbourhood in the reference (int(floor(NaN)) is far outside
    every grid) and adds exactly 0.  CUDA converts NaN to cell 0 — which IS occupied in a sensor-centred map — so the
    kernel must skip such points: score / gradient / Hessian / hit count equal the finite subset's, bit for bit, and
    equal the oracle's on the polluted cloud."""
    tgt, src = load_pair()
    gpu0 = nb.NormalDistributionsTransform()
    gpu0.setInputTarget(tgt)
    vox = gpu0.dump_voxels()
    info = gpu0.map_info()
    # the map must have a valid voxel at or next to cell (0,0,0): that is where a NaN lands on the device
    k0 = (0 - info["min_b"][0]) + (0 - info["min_b"][1]) * info["div_b"][0] + (0 - info["min_b"][2]) * info["div_b"][0] * info["div_b"][1]
    near = {k0, k0 + 1, k0 - 1, k0 + info["div_b"][0], k0 - info["div_b"][0]}
    assert any(int(k) in near and c >= 6 for k, c in zip(vox["keys"], vox["counts"]))
    bad = src.copy()
    bad[10] = [np.nan, 0.0, 0.0]
    bad[500] = [0.0, np.inf, 0.0]
    bad[501] = [np.nan, np.nan, np.nan]
    bad[9000] = [1.0, 2.0, -np.inf]
    bad[15000] = [3e38, 3e38, 3e38]       # finite, overflows in the transform
    good_mask = np.isfinite(bad).all(axis=1) & (np.abs(bad).max(axis=1) < 1e30)
    ref, gpu = make_pair(nb, tgt, bad, method=method)
    _, gpu_clean = make_pair(nb, tgt, bad[good_mask], method=method)
    for p in POSES[:3]:
        a, b, c = gpu.eval_derivatives(p), ref.eval_derivatives(p), gpu_clean.eval_derivatives(p)
        assert np.isfinite(a["score"]) and np.isfinite(a["gradient"]).all() and np.isfinite(a["hessian"]).all()
        assert a["hits"] == b["hits"] == c["hits"]
        assert abs(a["score"] - b["score"]) <= REL * abs(b["score"])
        assert rel_err(a["gradient"], b["gradient"]) < REL and rel_err(a["hessian"], b["hessian"]) < REL
        assert rel_err(a["gradient"], c["gradient"]) < 1e-6 and rel_err(a["hessian"], c["hessian"]) < 1e-6   # other point-to-lane grouping: fp32 sum order
        ha, hb = gpu.eval_hessian(p), ref.eval_hessian(p)
        assert np.isfinite(ha).all() and rel_err(ha, hb) < REL
    check_align(ref, gpu)
    sa, sb = gpu.calculateScore(bad[good_mask]), ref.calculateScore(bad[good_mask])
    assert abs(sa - sb) <= 1e-9 * abs(sb)


def test_set_resolution_rebuild_rule(nb):
    """setResolution (ndt_omp.h:132-142) rebuilds the voxel map only if the value CHANGED and an input SOURCE is set
    (`if (input_) init();`, sic): with a target but no source the old map stays until the next setInputTarget."""
    tgt, src = load_pair()
    gpu = nb.NormalDistributionsTransform()
    gpu.setInputTarget(tgt)                       # built at 1.0
    v1 = gpu.map_info()["n_voxels"]
    gpu.setResolution(2.0)                        # no source yet: NOT rebuilt
    assert gpu.map_info()["n_voxels"] == v1
    gpu.setInputSource(src)
    gpu.setResolution(2.0)                        # unchanged value: still not rebuilt
    assert gpu.map_info()["n_voxels"] == v1
    gpu.setResolution(0.5)                        # changed + source set: rebuilt
    v05 = gpu.map_info()["n_voxels"]
    ref = oracle.NormalDistributionsTransform()
    ref.setResolution(0.5)
    ref.setInputTarget(tgt)
    assert v05 == ref.map_info()["n_voxels"] and v05 > v1
    gpu.setInputTarget(tgt)                       # setInputTarget always rebuilds with the current value
    assert gpu.map_info()["n_voxels"] == v05


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("pair,kw", [("pair_ds0p1.npz", {}), ("pair_ds0p3.npz", {"eps": 0.01, "max_iter": 64})])
def test_sharded_align_emulated_ranks_one_gpu(nb, world, pair, kw):
    """The source-sharded solve on ONE GPU: `world` ranks emulated inside one cooperative launch, each
    on the contiguous source range it would own, exchanging the 29 sums through the same tagged mailbox stores / polls
    as the multi-GPU path.  Same bar as tests/test_gpu_multi.py: every rank reports identical bits, counts equal the
    oracle's, final transform within tolerance, and within 1e-9 of the single-rank solve (different summation grouping)."""
    tgt, src = load_pair(pair)
    ref, gpu = make_pair(nb, tgt, src, **kw)
    ref.align()
    rr = ref.result()
    res = gpu.align_emulated_ranks(world)
    assert len(res) == world
    for r in res[1:]:
        assert np.array_equal(r["final"], res[0]["final"]) and r["iterations"] == res[0]["iterations"]
        assert r["n_evaluations"] == res[0]["n_evaluations"] and r["final_score"] == res[0]["final_score"]
        assert r["trans_probability"] == res[0]["trans_probability"] and r["n_hits"] == res[0]["n_hits"]
    rg = res[0]
    assert (rg["iterations"], rg["n_evaluations"], rg["n_hessian_passes"], rg["converged"]) == \
           (rr["iterations"], rr["n_evaluations"], rr["n_hessian_passes"], rr["converged"])
    dt, dr = transform_delta(rg["final"], rr["final"])
    assert dt < TRANS_TOL and dr < ROT_TOL, (dt, dr)
    assert abs(rg["trans_probability"] - rr["trans_probability"]) <= 1e-5 * abs(rr["trans_probability"])
    gpu.align()
    single = gpu.result()
    assert single["n_hits"] == rg["n_hits"]
    assert abs(single["final_score"] - rg["final_score"]) <= 1e-9 * abs(single["final_score"])
    assert np.abs(single["final_pose"] - rg["final_pose"]).max() < 1e-9
    # run to run: identical bits
    again = gpu.align_emulated_ranks(world)
    assert np.array_equal(again[0]["final"], rg["final"]) and again[0]["final_score"] == rg["final_score"]
    # with a guess (the tag sequence continues across launches)
    g = pose_matrix([0.2, 0.05, -0.01, 0.01, -0.02, 0.03])
    ref.align(g)
    rg2 = gpu.align_emulated_ranks(world, g)[world - 1]
    dt, dr = transform_delta(rg2["final"], ref.result()["final"])
    assert dt < TRANS_TOL and dr < ROT_TOL and rg2["iterations"] == ref.result()["iterations"]


def test_voxelgrid_filter_non_cubic_leaf(nb):
    """pcl::VoxelGrid::setLeafSize(lx, ly, lz) (the shim used to drop ly, lz): per-axis leaves on the device,
    bit-identical to the oracle (which is pinned to the independent numpy restatement on the CPU side)."""
    tgt, src = load_pair()
    gpu = nb.NormalDistributionsTransform()
    big = np.concatenate([tgt + np.float32(0.013 * k) for k in range(6)]).astype(np.float32)   # > 64 k points: staged path
    for cloud in (tgt, big):
        for leaf in ((0.3, 0.5, 0.2), (1.0, 0.25, 2.0)):
            assert np.array_equal(gpu.voxelgrid_filter(cloud, leaf), oracle.voxelgrid_downsample(cloud, leaf))
    # the NDT grid of the same handle stays cubic after a non-cubic filter call
    gpu.setInputTarget(tgt)
    ref = oracle.NormalDistributionsTransform()
    ref.setInputTarget(tgt)
    assert gpu.map_info()["n_voxels"] == ref.map_info()["n_voxels"]


def test_large_source_sorted_by_voxel_key(nb, monkeypatch):
    """Sources of >= 262 144 points are aligned from a copy sorted by voxel key (coalesced probes, shared Gaussian
    records): only the summation order changes.  Oracle parity on a 300 k-point source in random order, equality
    (to summation-order noise) with the unsorted path, the caller's point order in the output cloud, emulated ranks."""
    tgt, src = synthetic_scene(n_target=120000, n_source=300000, seed=31)
    ref, gpu = make_pair(nb, tgt, src)
    monkeypatch.setenv("NDTB200_SORT_SOURCE_MIN", "0")
    _, plain = make_pair(nb, tgt, src)
    p = np.array([0.3, -0.2, 0.04, 0.004, -0.005, 0.012])
    b = ref.eval_derivatives(p)
    c = plain.eval_derivatives(p)
    monkeypatch.delenv("NDTB200_SORT_SOURCE_MIN")
    a = gpu.eval_derivatives(p)
    assert a["hits"] == b["hits"] == c["hits"] and b["hits"] > 500000
    assert abs(a["score"] - b["score"]) <= REL * abs(b["score"])
    assert rel_err(a["gradient"], b["gradient"]) < REL and rel_err(a["hessian"], b["hessian"]) < REL
    assert rel_err(a["gradient"], c["gradient"]) < 1e-6 and rel_err(a["hessian"], c["hessian"]) < 1e-6
    assert rel_err(gpu.eval_hessian(p), ref.eval_hessian(p)) < REL
    out, rr, rg = check_align(ref, gpu)
    assert np.array_equal(out[:, :3], oracle.transform_points(rg["final"], src)[:, :3])      # output keeps the caller's order
    again = gpu.align()
    assert np.array_equal(again, out)                                                        # reproducible
    res = gpu.align_emulated_ranks(4)
    assert all(np.array_equal(r["final"], res[0]["final"]) for r in res)
    dt, dr = transform_delta(res[0]["final"], rr["final"])
    assert dt < TRANS_TOL and dr < ROT_TOL and res[0]["n_evaluations"] == rr["n_evaluations"]


@pytest.mark.parametrize("leaf", [0.3, 0.7, 1.0, 2.5])
def test_lookup_cells_near_faces_are_bit_exact(nb, leaf):
    """The lookup's cell index is int(floor(x / leaf)) with an fp32 DIVISION (Q8); the kernel decides most points from
    x * fl(1/leaf) and falls back to the exact division near cell faces (lookup_cell).  Queries ON and within a few ulps
    of the faces (both the fp32 product k * leaf and the fp32 nearest to the real face), for leaves whose reciprocal is
    inexact, must land in the oracle's cells — checked through the returned neighbour keys and through the hit count
    of a derivative evaluation over a source made of such points."""
    tgt, src = load_pair()
    ref, gpu = make_pair(nb, tgt, src, res=leaf)
    info = gpu.map_info()
    rng = np.random.default_rng(17)
    n = 4000
    clouds = []
    for axis in range(3):
        k = rng.integers(info["min_b"][axis] - 2, info["max_b"][axis] + 3, size=n)
        f1 = (k.astype(np.float32) * np.float32(leaf)).astype(np.float32)          # fp32 product
        f2 = (k.astype(np.float64) * leaf).astype(np.float32)                      # nearest fp32 to the real face
        variants = []
        for f in (f1, f2):
            up = dn = f
            variants.append(f)
            for _ in range(3):
                up = np.nextafter(up, np.float32(np.inf)); dn = np.nextafter(dn, np.float32(-np.inf))
                variants += [up, dn]
        for v in variants:
            pts = src[rng.integers(0, len(src), size=n)].copy()
            pts[:, axis] = v
            clouds.append(pts)
    q = np.concatenate(clouds).astype(np.float32)
    assert np.array_equal(gpu.lookup(q, oracle.DIRECT7), ref.lookup(q, oracle.DIRECT7))
    assert np.array_equal(gpu.lookup(q, oracle.DIRECT26), ref.lookup(q, oracle.DIRECT26))
    ref.setInputSource(q); gpu.setInputSource(q)
    a, b = gpu.eval_derivatives(np.zeros(6)), ref.eval_derivatives(np.zeros(6))
    assert a["hits"] == b["hits"] and b["hits"] > 10000
    assert rel_err(a["gradient"], b["gradient"]) < REL


@pytest.mark.parametrize("method", [oracle.DIRECT7, oracle.DIRECT1, oracle.DIRECT26, oracle.KDTREE])
def test_calculate_score_batch_and_pose_screening(nb, method):
    """calculateScore (ndt_omp_impl.hpp:935-983) for many clouds / many candidate poses in one call (loop-closure
    screening): every entry equals the oracle's calculateScore of that cloud to 1e-9."""
    tgt, src = load_pair()
    ref, gpu = make_pair(nb, tgt, src, method=method)
    rng = np.random.default_rng(23)
    poses = [np.eye(4, dtype=np.float32)] + [pose_matrix(np.concatenate([rng.uniform(-1.0, 1.0, 3), rng.uniform(-0.1, 0.1, 3)])) for _ in range(9)]
    poses.append(pose_matrix([500.0, 0, 0, 0, 0, 0]))                       # far away: no neighbourhood at all -> 0
    clouds = [oracle.transform_points(T, src)[:, :3] for T in poses]
    clouds.append(src[:7])                                                   # ragged sizes
    clouds.append(np.zeros((0, 3), np.float32))                              # an empty cloud scores 0
    exp = np.array([ref.calculateScore(c) if len(c) else 0.0 for c in clouds])
    got = gpu.calculateScoreBatch(clouds)
    assert got.shape == exp.shape
    assert np.abs(got - exp).max() <= 1e-9 * np.abs(exp).max()
    assert got[len(poses) - 1] == 0.0
    got_p = gpu.scorePoses(poses)
    assert np.abs(got_p - exp[:len(poses)]).max() <= 1e-9 * np.abs(exp).max()
    assert abs(gpu.calculateScore(src) - exp[0]) <= 1e-12 * abs(exp[0])     # the single-cloud entry point: same kernel
    assert int(np.argmin(got_p)) == int(np.argmin(exp[:len(poses)]))        # lower is better: the screening picks the same pose


@pytest.mark.parametrize("slices,owners", [(3, 2), (4, 4), (2, 8)])
def test_map_build_owner_partitioned(nb, slices, owners):
    """The sharded build as designed with its pieces on one GPU: the cloud cut into contiguous slices,
    per-slice partials with the common grid, every partial sent to the OWNER of its key range, owners merge (slice order)
    and finalise, the finished records concatenated in owner (= key) order and installed: keys / counts exact, means /
    inverse covariances within the bar, same lookups and derivatives as the single build."""
    import torch
    from toyslam_b200.sharding import point_range
    tgt, src = load_pair()
    ref, single = make_pair(nb, tgt, src)
    dev = torch.device("cuda", 0)
    workers, boxes, nf_total = [], [], 0
    for r in range(slices):
        lo, hi = point_range(len(tgt), r, slices)
        loc = torch.ones((hi - lo, 4), dtype=torch.float32, device=dev)
        loc[:, :3] = torch.as_tensor(tgt[lo:hi]).to(dev)
        w = nb.NormalDistributionsTransform()
        mn, mx, nf = w.cloud_bounds(loc.data_ptr(), hi - lo)
        boxes.append((mn, mx)); nf_total += nf
        workers.append((w, loc))
    gmin = np.min([b[0] for b in boxes], axis=0)
    gmax = np.max([b[1] for b in boxes], axis=0)
    parts = []
    for w, loc in workers:
        st, nv = w.build_partials(gmin, gmax)
        assert st == 0 and nv > 0
        k = torch.zeros(nv, dtype=torch.int32, device=dev)
        c = torch.zeros(nv, dtype=torch.int32, device=dev)
        m = torch.zeros((nv, 9), dtype=torch.float64, device=dev)
        w.copy_partials(k.data_ptr(), c.data_ptr(), m.data_ptr())
        parts.append((k, c, m))
    all_keys = torch.sort(torch.cat([p[0] for p in parts])).values
    upper = [int(all_keys[(o + 1) * len(all_keys) // owners].item()) for o in range(owners - 1)]
    offs = [w.partials_split(upper, owners) for w, _ in workers]
    for o_, (k, _, _) in zip(offs, parts):
        assert o_[0] == 0 and o_[-1] == len(k) and np.all(np.diff(o_) >= 0)
    recs, ics, n_all = [], [], 0
    for o in range(owners):
        k = torch.cat([p[0][offs[i][o]:offs[i][o + 1]] for i, p in enumerate(parts)]).contiguous()
        c = torch.cat([p[1][offs[i][o]:offs[i][o + 1]] for i, p in enumerate(parts)]).contiguous()
        m = torch.cat([p[2][offs[i][o]:offs[i][o + 1]] for i, p in enumerate(parts)]).contiguous()
        owner = nb.NormalDistributionsTransform()
        st, n_own = owner.merge_partials(gmin, gmax, nf_total, k.data_ptr(), c.data_ptr(), m.data_ptr(), len(k))
        assert st == 0
        rec = torch.zeros((max(1, n_own), 16), dtype=torch.int32, device=dev)
        ic = torch.zeros((max(1, n_own), 6), dtype=torch.float64, device=dev)
        if n_own:
            owner.copy_records(rec.data_ptr(), ic.data_ptr())
        recs.append(rec[:n_own]); ics.append(ic[:n_own]); n_all += n_own
    g_rec, g_ic = torch.cat(recs).contiguous(), torch.cat(ics).contiguous()
    full = nb.NormalDistributionsTransform()
    assert full.set_map_from_records(gmin, gmax, nf_total, g_rec.data_ptr(), g_ic.data_ptr(), n_all) == 0
    ri, gi = ref.map_info(), full.map_info()
    for key in ("min_b", "max_b", "div_b"):
        assert np.array_equal(ri[key], gi[key])
    assert ri["n_voxels"] == gi["n_voxels"] and ri["n_valid"] == gi["n_valid"]
    rl, gl = ref.dump_leaves(), full.dump_voxels()
    assert np.array_equal(rl["keys"], gl["keys"]) and np.array_equal(rl["counts"], gl["counts"])
    valid = rl["counts"] >= 6
    assert rel_err(gl["mean"], rl["mean"]) < REL
    a, b = gl["icov"][valid], rl["icov"][valid]
    assert (np.abs(a - b).reshape(len(b), -1).max(axis=1) / np.abs(b).reshape(len(b), -1).max(axis=1)).max() < REL
    full.setInputSource(src)
    p = np.array([0.4, 0.1, -0.02, 0.005, -0.001, -0.01])
    x, y = full.eval_derivatives(p), ref.eval_derivatives(p)
    assert x["hits"] == y["hits"]
    assert rel_err(x["gradient"], y["gradient"]) < REL and rel_err(x["hessian"], y["hessian"]) < REL
    q = np.concatenate([src[:2000], np.round(src[:200])]).astype(np.float32)
    assert np.array_equal(full.lookup(q, oracle.DIRECT7), single.lookup(q, oracle.DIRECT7))
    check_align(ref, full)
    with pytest.raises(nb.NdtError):
        full.getFitnessScore()       # a records-only map holds no raw target


def test_set_target_device_view_builds_without_copy(nb):
    """ndtb200_set_target_device_view: the map references the caller's device cloud (as the reference keeps the caller's
    cloud by shared pointer): same map, same align, same fitness as the copying entry points."""
    import torch
    tgt, src = load_pair()
    dev = torch.device("cuda", 0)
    cloud = torch.ones((len(tgt), 4), dtype=torch.float32, device=dev)
    cloud[:, :3] = torch.as_tensor(tgt).to(dev)
    a, b = nb.NormalDistributionsTransform(), nb.NormalDistributionsTransform()
    assert a.set_target_device_view(cloud.data_ptr(), len(tgt)) == 0
    b.setInputTarget(tgt)
    da, db = a.dump_voxels(), b.dump_voxels()
    for k in ("keys", "counts", "mean", "cov", "icov"):
        assert np.array_equal(da[k], db[k])
    assert np.array_equal(a.point_keys(), b.point_keys())
    a.setInputSource(src); b.setInputSource(src)
    a.align(); b.align()
    assert np.array_equal(a.result()["final"], b.result()["final"])
    assert a.getFitnessScore() == b.getFitnessScore()
    c = a.clone()                       # a copy owns its own cloud
    del a
    c.align()
    assert np.array_equal(c.result()["final"], b.result()["final"]) and c.getFitnessScore() == b.getFitnessScore()


def test_run_pairs_matches_single_calls(nb):
    """ndtb200_run_pairs (batched scan-to-scan odometry: every lane = one handle + one C++ host thread inside the call) must
    return, for every pair, exactly what setInputTarget / setInputSource / align return on a handle of its own."""
    from util import as_xyzw_host
    pairs = []
    for seed in range(7):
        tgt, src = synthetic_scene(n_target=20000 + 3000 * seed, n_source=5000 + 500 * seed, seed=20 + seed)
        pairs.append((as_xyzw_host(tgt), as_xyzw_host(src)))
    guesses = [np.eye(4, dtype=np.float32) for _ in pairs]
    guesses[3][:3, 3] = (0.2, -0.1, 0.05)
    lanes = [nb.NormalDistributionsTransform() for _ in range(3)]
    for ln in lanes:
        ln.setTransformationEpsilon(0.01)
        ln.setMaximumIterations(40)
    pipe = nb.PairPipeline(lanes, [(t.ctypes.data, len(t)) for t, _ in pairs], [(s.ctypes.data, len(s)) for _, s in pairs], guesses)
    pipe.run()
    got = pipe.results()
    pipe.run()                                   # a second pass over the same pairs on the same lanes: identical bits
    again = pipe.results()
    for k, (t, s) in enumerate(pairs):
        one = nb.NormalDistributionsTransform()
        one.setTransformationEpsilon(0.01)
        one.setMaximumIterations(40)
        one.setInputTarget(t[:, :3])
        one.setInputSource(s[:, :3])
        one.align(guesses[k])
        r = one.result()
        for g in (got[k], again[k]):
            assert g["iterations"] == r["iterations"] and g["n_evaluations"] == r["n_evaluations"] and g["converged"] == r["converged"]
            assert np.array_equal(g["final"], r["final"])
            assert g["trans_probability"] == r["trans_probability"]
    # the oracle on one of the pairs (the pipeline is the same hot path, this is a plumbing check)
    ref = oracle.NormalDistributionsTransform()
    ref.setTransformationEpsilon(0.01)
    ref.setMaximumIterations(40)
    ref.setInputTarget(pairs[3][0][:, :3])
    ref.setInputSource(pairs[3][1][:, :3])
    ref.align(guesses[3])
    dt, dr = transform_delta(got[3]["final"], ref.result()["final"])
    assert dt < TRANS_TOL and dr < ROT_TOL
