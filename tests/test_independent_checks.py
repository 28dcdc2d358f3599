"""Independent checks (CPU, -m "not gpu") of the pieces no reference golden pins: the oracle AND the
library's host twin are compared with third-party implementations (scipy Rotation / polar, numpy pinv, an enumerated
neighbour set), not with one another — a mistake shared by the two hand-written restatements cannot hide here.

  Q5  Eigen 3.3 `rotation().eulerAngles(0,1,2)` on non-identity guesses (ndt_omp_impl.hpp:103-111)
  --  `JacobiSVD<6x6>(H).solve(-g)` on definite, indefinite and rank-deficient H (ndt_omp_impl.hpp:127-129)
  Q7  DIRECT26 = the 3x3x3 block minus the centre cell; DIRECT7 = centre + 6 faces (vgc_impl.hpp:407-433)
"""
import json
import os

import numpy as np
import pytest
from scipy.linalg import polar
from scipy.spatial.transform import Rotation

import oracle
from util import GOLDEN, load_pair, pose_matrix


def _wrap(a):
    return (np.asarray(a) + np.pi) % (2 * np.pi) - np.pi


def eigen33_euler_012_from_scipy(R):
    """What Eigen 3.3's eulerAngles(0,1,2) must return, derived from scipy's own XYZ decomposition: scipy gives the
    (r, p, y) with p in [-pi/2, pi/2]; Eigen 3.3 returns the decomposition whose FIRST angle lies in [0, pi], which is
    the same triple when r >= 0 and the 'other' solution (r + pi, pi - p, y + pi) otherwise."""
    r, p, y = Rotation.from_matrix(R).as_euler("XYZ")  # intrinsic: R = Rx(r) Ry(p) Rz(y)
    if r >= 0:
        return np.array([r, p, y])
    return np.array([r + np.pi, _wrap(np.pi - p), _wrap(y + np.pi)])


def _guess_cases():
    rng = np.random.default_rng(42)
    cases = [np.zeros(3), np.array([0.01, -0.02, 0.03]), np.array([-0.01, 0.02, -0.03]),   # small +/- roll (Q5)
             np.array([-1e-4, 1e-4, 0.5]), np.array([3.0, 0.2, -3.0]), np.array([-3.0, -0.2, 3.0]),
             np.array([1.2, 1.4, -2.2]), np.array([-2.9, -1.3, 0.1])]
    cases += [rng.uniform(-np.pi, np.pi, 3) * np.array([1.0, 0.49, 1.0]) for _ in range(200)]   # pitch away from the gimbal lock
    return cases


def _impls():
    import toyslam_b200 as nb      # host arithmetic only: runs without a GPU
    return [("oracle", oracle.matrix_to_pose), ("libndt_b200 host", nb.guess_to_pose)]


@pytest.mark.parametrize("name,fn", _impls())
def test_euler_angles_012_against_scipy(name, fn):
    for ang in _guess_cases():
        R = Rotation.from_euler("XYZ", ang).as_matrix()
        T = np.eye(4)
        T[:3, :3] = R
        T[:3, 3] = [0.3, -1.5, 2.25]
        T32 = T.astype(np.float32)
        p = fn(T32)
        assert np.array_equal(p[:3], T32[:3, 3].astype(np.float64))
        # (1) Eigen 3.3's range rule: first angle in [0, pi] (float32 pi), the others in [-pi, pi]
        assert -1e-6 <= p[3] <= np.pi + 1e-6 and abs(p[4]) <= np.pi + 1e-6 and abs(p[5]) <= np.pi + 1e-6, (name, ang, p)
        # (2) the angles re-compose to the rotation (scipy builds the matrix)
        R_back = Rotation.from_euler("XYZ", p[3:]).as_matrix()
        assert np.abs(R_back - T32[:3, :3].astype(np.float64)).max() < 5e-6, (name, ang, p)
        # (3) and they are the specific solution Eigen 3.3 returns, predicted from scipy's decomposition
        exp = eigen33_euler_012_from_scipy(T32[:3, :3].astype(np.float64))
        if abs(abs(exp[1]) - np.pi / 2) > 1e-2 and min(abs(exp[0]), abs(exp[0] - np.pi)) > 1e-5:   # away from the branch points
            assert np.abs(_wrap(p[3:] - exp)).max() < 2e-5, (name, ang, p[3:], exp)


@pytest.mark.parametrize("name,fn", _impls())
def test_negative_roll_guess_is_reparametrised(name, fn):
    """Q5 in words: a guess with a small NEGATIVE roll comes back as (pi + r, pi - p, pi + y)-like angles."""
    p = fn(pose_matrix([0.2, 0.05, -0.01, -0.01, 0.02, -0.03]))
    assert abs(p[3] - (np.pi - 0.01)) < 1e-5 and abs(abs(p[4]) - (np.pi - 0.02)) < 1e-5 and abs(abs(p[5]) - (np.pi - 0.03)) < 1e-5
    p = fn(pose_matrix([0.2, 0.05, -0.01, 0.01, -0.02, 0.03]))
    assert np.abs(p[3:] - np.array([0.01, -0.02, 0.03])).max() < 1e-6


@pytest.mark.parametrize("name,fn", _impls())
def test_rotation_is_the_polar_factor(name, fn):
    """Transform::rotation() is the orthogonal polar factor of the linear part (computeRotationScaling), so a guess
    whose 3x3 block is not exactly orthonormal (fp32 rounding, mild scale / shear) yields the angles of
    scipy.linalg.polar's U."""
    rng = np.random.default_rng(7)
    for _ in range(100):
        ang = rng.uniform(-np.pi, np.pi, 3) * np.array([1.0, 0.45, 1.0])
        R = Rotation.from_euler("XYZ", ang).as_matrix()
        M = R @ (np.eye(3) + 0.02 * rng.normal(size=(3, 3)))        # R times a near-identity stretch
        T = np.eye(4, dtype=np.float32)
        T[:3, :3] = M.astype(np.float32)
        U, _ = polar(T[:3, :3].astype(np.float64))
        p = fn(T)
        R_back = Rotation.from_euler("XYZ", p[3:]).as_matrix()
        assert np.abs(R_back - U).max() < 5e-6, (name, ang)


def _solve_cases():
    rng = np.random.default_rng(3)
    cases = []
    for _ in range(20):                                   # negative definite (the normal case near the optimum)
        A = rng.normal(size=(6, 6))
        cases.append((-(A @ A.T + 0.5 * np.eye(6)) * 10 ** rng.uniform(0, 6), rng.normal(size=6) * 1e3, "definite"))
    for _ in range(20):                                   # indefinite (DIRECT26's first Hessian)
        Q, _ = np.linalg.qr(rng.normal(size=(6, 6)))
        ev = rng.uniform(1, 100, 6) * rng.choice([-1, 1], 6)
        ev[0], ev[1] = abs(ev[0]), -abs(ev[1])
        cases.append((Q @ np.diag(ev) @ Q.T, rng.normal(size=6), "indefinite"))
    for rank in (5, 4, 2):                                # rank deficient: pseudo-inverse (minimum-norm) solution
        for _ in range(5):
            B = rng.normal(size=(6, rank))
            cases.append((-(B @ B.T), rng.normal(size=6), "rank%d" % rank))
    H = np.diag([-3.0, -2.0, 0.0, -5.0, 0.0, -1.0])       # unobservable axes: exact zero rows / columns
    cases.append((H, np.array([1.0, 2.0, 3.0, 4.0, 5.0, 6.0]), "zero rows"))
    cases.append((np.zeros((6, 6)), np.ones(6), "zero matrix"))
    return cases


def numpy_jacobisvd_solve(H, b):
    """JacobiSVD::solve with Eigen's default threshold: singular values <= eps * max(rows, cols)... Eigen uses
    threshold = NumTraits<double>::epsilon() * diagSize for JacobiSVD -> rank; pinv with that relative cut-off."""
    return np.linalg.pinv(H, rcond=6 * np.finfo(np.float64).eps, hermitian=False) @ b


def test_oracle_newton_solve_against_numpy_pinv():
    for H, g, kind in _solve_cases():
        x = oracle.svd_solve6(H, -g)
        ref = numpy_jacobisvd_solve(H, -g)
        scale = max(1e-300, np.abs(ref).max())
        tol = 1e-9 if kind in ("definite", "indefinite") else 1e-8
        assert np.abs(x - ref).max() <= tol * max(scale, 1e-12) + 1e-14, (kind, x, ref)


def enumerated_neighbours(keys_valid, grid, q, leaf, kind):
    """Keys of the valid voxels around query q by brute enumeration: kind 26 = 3x3x3 block minus the centre,
    7 = centre + the six face neighbours, 1 = centre."""
    min_b, max_b, div_b = grid
    c = np.floor(q.astype(np.float32) / np.float32(leaf)).astype(np.int64)
    out = set()
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                nz = abs(dx) + abs(dy) + abs(dz)
                if kind == 26 and nz == 0:
                    continue
                if kind == 7 and nz > 1:
                    continue
                if kind == 1 and nz != 0:
                    continue
                cc = c + np.array([dx, dy, dz])
                if np.any(cc < min_b) or np.any(cc > max_b):
                    continue
                key = int((cc[0] - min_b[0]) + (cc[1] - min_b[1]) * div_b[0] + (cc[2] - min_b[2]) * div_b[0] * div_b[1])
                if key in keys_valid:
                    out.add(key)
    return out


def lookup_queries():
    tgt, src = load_pair()
    rng = np.random.default_rng(11)
    return tgt, np.concatenate([src[:600], rng.uniform(-40, 40, size=(300, 3)).astype(np.float32),
                                np.round(src[:100])]).astype(np.float32)


@pytest.mark.parametrize("method,kind", [(oracle.DIRECT26, 26), (oracle.DIRECT7, 7), (oracle.DIRECT1, 1)])
def test_oracle_neighbourhoods_against_enumeration(method, kind):
    tgt, q = lookup_queries()
    n = oracle.NormalDistributionsTransform()
    n.setInputTarget(tgt)
    info = n.map_info()
    leaves = n.dump_leaves()
    valid = set(int(k) for k, c in zip(leaves["keys"], leaves["counts"]) if c >= 6)
    got = n.lookup(q, method)
    n_nonempty = 0
    for i in range(len(q)):
        exp = enumerated_neighbours(valid, (info["min_b"], info["max_b"], info["div_b"]), q[i], 1.0, kind)
        row = [int(k) for k in got[i] if k >= 0]
        assert len(row) == len(set(row))                  # no cell twice
        assert set(row) == exp, (i, q[i], sorted(row), sorted(exp))
        n_nonempty += bool(exp)
    assert n_nonempty > 500


def test_oracle_line_search_fixtures_match_survey_appendix_a():
    """Counts and final poses of SURVEY Appendix A (values of the survey's independent numpy restatement, committed in
    golden.json by make_fixtures.py): config 1 DIRECT26, fixture A (raw pair: 23 evaluations, 2 Hessian-only passes) and
    fixture B (0.3 m, node parameters)."""
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        ap = json.load(f)["appendix_a"]

    def run(tgt, src, method=oracle.DIRECT7, **kw):
        n = oracle.NormalDistributionsTransform()
        n.setNeighborhoodSearchMethod(method)
        if "eps" in kw:
            n.setTransformationEpsilon(kw["eps"]); n.setMaximumIterations(kw["max_iter"])
        n.setInputTarget(tgt); n.setInputSource(src); n.align()
        return n

    def check(n, exp):
        r = n.result()
        assert (r["iterations"], r["n_evaluations"], r["n_hessian_passes"]) == (exp["iterations"], exp["evaluations"], exp["hessian_passes"])
        assert np.abs(n.trace()["x"][-1] - np.array(exp["p"])).max() < 1e-6

    d = np.load(os.path.join(GOLDEN, "pair_raw.npz"))
    a = run(d["target"], d["source"])
    assert a.map_info()["n_valid"] == ap["fixture_A_raw_pair"]["valid_voxels"]
    check(a, ap["fixture_A_raw_pair"])
    tr = a.trace()
    assert list(tr["kind"]) == [0] + ([0] + [1] * 10 + [2]) * 2          # per iteration: full trial, 10 extra trials, Hessian pass
    assert np.all(tr["a_t"][1:13] == 0.05) and tr["a_t"][13] == 0.1 and np.all(tr["a_t"][14:] == 0.05)
    check(run(*load_pair("pair_ds0p3.npz"), eps=0.01, max_iter=64), ap["fixture_B_ds0p3_node_params"])
    c = run(*load_pair(), method=oracle.DIRECT26)
    check(c, ap["config1_DIRECT26"])
    assert "%.6f" % c.getFitnessScore() == "%.6f" % ap["config1_DIRECT26"]["fitness"]


def test_pose_to_matrix_against_scipy():
    """static convertTransform (ndt_omp.h:216-233) = Translation * Rx * Ry * Rz in fp32: the oracle's and the library's
    host function against scipy's intrinsic XYZ rotation (fp64) — agreement to fp32 rounding, and bit-equal to each other."""
    import toyslam_b200 as nb
    rng = np.random.default_rng(5)
    for _ in range(200):
        p = np.concatenate([rng.uniform(-50, 50, 3), rng.uniform(-np.pi, np.pi, 3)])
        R = Rotation.from_euler("XYZ", p[3:]).as_matrix()
        for fn in (oracle.pose_to_matrix, nb.pose_to_matrix):
            T = fn(p)
            assert T.dtype == np.float32 and np.array_equal(T[3], [0, 0, 0, 1])
            assert np.array_equal(T[:3, 3], p[:3].astype(np.float32))
            assert np.abs(T[:3, :3].astype(np.float64) - R).max() < 4e-7
        assert np.array_equal(oracle.pose_to_matrix(p), nb.pose_to_matrix(p))


def test_voxelgrid_non_cubic_leaf_oracle_against_numpy():
    """pcl::VoxelGrid::setLeafSize(lx, ly, lz) with a non-cubic leaf: the oracle against the independent numpy
    restatement of tests/golden/make_fixtures.py (bit for bit)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(GOLDEN, "make_fixtures.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    tgt, _ = load_pair()
    for leaf in ((0.3, 0.5, 0.2), (1.0, 0.25, 2.0), 0.3):
        assert np.array_equal(oracle.voxelgrid_downsample(tgt, leaf), mk.voxelgrid_downsample(tgt, leaf))
