"""ctypes front-end of the CPU oracle (oracle/ndt_oracle.hpp).

TEST INFRASTRUCTURE ONLY.  Allowed importers: tests/, __graft_entry__.smoke(), and bench.py's
`cpu_baseline` / `--impl reference` legs.  The product package (toyslam_b200) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libndt_oracle.so")
_lib = None

KDTREE, DIRECT26, DIRECT7, DIRECT1 = 0, 1, 2, 3
BUILD_OK, BUILD_NO_INPUT, BUILD_GRID_OVERFLOW = 0, 1, 2


def build(force=False):
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    try:
        L = C.CDLL(_LIB_PATH)
    except OSError:
        build(force=True)
        L = C.CDLL(_LIB_PATH)
    f32p, f64p, i32p, i64p = (C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_long))
    L.ndto_create.restype = C.c_void_p
    L.ndto_destroy.argtypes = [C.c_void_p]
    L.ndto_max_threads.restype = C.c_int
    L.ndto_set_params.argtypes = [C.c_void_p, C.c_float, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_double]
    L.ndto_set_target.argtypes = [C.c_void_p, f32p, C.c_size_t, C.c_int]
    L.ndto_set_target.restype = C.c_int
    L.ndto_set_source.argtypes = [C.c_void_p, f32p, C.c_size_t]
    L.ndto_align.argtypes = [C.c_void_p, f32p, f32p]
    L.ndto_get_result.argtypes = [C.c_void_p, f32p, i32p, i32p, f64p, i32p, i32p]
    L.ndto_fitness.argtypes = [C.c_void_p, C.c_double]
    L.ndto_fitness.restype = C.c_double
    L.ndto_calculate_score.argtypes = [C.c_void_p, f32p, C.c_size_t]
    L.ndto_calculate_score.restype = C.c_double
    L.ndto_gauss.argtypes = [C.c_void_p, f64p]
    L.ndto_map_info.argtypes = [C.c_void_p, i32p, i32p, i32p, i64p, i64p]
    L.ndto_point_keys.argtypes = [C.c_void_p, i32p]
    L.ndto_dump_leaves.argtypes = [C.c_void_p, i32p, i32p, f64p, f64p, f64p, i32p]
    L.ndto_dump_leaves.restype = C.c_long
    L.ndto_eval_derivatives.argtypes = [C.c_void_p, f64p, f32p, C.c_int, f64p]
    L.ndto_eval_derivatives.restype = C.c_long
    L.ndto_eval_hessian.argtypes = [C.c_void_p, f64p, f32p, f64p]
    L.ndto_lookup.argtypes = [C.c_void_p, f32p, C.c_size_t, C.c_int, i32p]
    L.ndto_lookup.restype = C.c_long
    L.ndto_trace.argtypes = [C.c_void_p, i32p, f64p, f64p, f64p, C.c_long]
    L.ndto_trace.restype = C.c_long
    L.ndto_voxelgrid.argtypes = [f32p, C.c_size_t, C.c_float, f32p, C.c_size_t]
    L.ndto_voxelgrid.restype = C.c_long
    L.ndto_voxelgrid3.argtypes = [f32p, C.c_size_t, f32p, f32p, C.c_size_t]
    L.ndto_voxelgrid3.restype = C.c_long
    L.ndto_pose_to_matrix.argtypes = [f64p, f32p]
    L.ndto_matrix_to_pose.argtypes = [f32p, f64p]
    L.ndto_transform.argtypes = [f32p, f32p, C.c_size_t, f32p]
    L.ndto_svd_solve6.argtypes = [f64p, f64p, f64p]
    _lib = L
    return L


def _f32(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f64(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i32(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def as_xyzw(points):
    """(n,3) or (n,4) float array -> contiguous (n,4) float32 with w = 1 (pcl::PointXYZ layout)."""
    p = np.asarray(points, dtype=np.float32)
    if p.ndim != 2 or p.shape[1] not in (3, 4):
        raise ValueError("points must be (n,3) or (n,4)")
    out = np.ones((p.shape[0], 4), dtype=np.float32)
    out[:, :3] = p[:, :3]
    return out


def max_threads():
    return int(lib().ndto_max_threads())


def voxelgrid_downsample(points, leaf):
    """pcl::VoxelGrid centroid downsample; `leaf` is a scalar or (lx, ly, lz)."""
    p = as_xyzw(points)
    out = np.empty_like(p)
    if np.ndim(leaf) == 0:
        n = lib().ndto_voxelgrid(_f32(p), p.shape[0], float(leaf), _f32(out), out.shape[0])
    else:
        l3 = np.ascontiguousarray(leaf, dtype=np.float32)
        n = lib().ndto_voxelgrid3(_f32(p), p.shape[0], _f32(l3), _f32(out), out.shape[0])
    if n < 0:
        raise OverflowError("leaf size too small (int32 voxel index overflow)")
    return out[:n, :3].copy()


def pose_to_matrix(p):
    p = np.ascontiguousarray(p, dtype=np.float64)
    T = np.empty(16, dtype=np.float32)
    lib().ndto_pose_to_matrix(_f64(p), _f32(T))
    return T.reshape(4, 4).T.copy()  # column-major buffer -> row-major numpy


def matrix_to_pose(T):
    Tc = np.ascontiguousarray(np.asarray(T, dtype=np.float32).T).reshape(-1)
    p = np.empty(6, dtype=np.float64)
    lib().ndto_matrix_to_pose(_f32(Tc), _f64(p))
    return p


def transform_points(T, points):
    Tc = np.ascontiguousarray(np.asarray(T, dtype=np.float32).T).reshape(-1)
    p = as_xyzw(points)
    out = np.empty_like(p)
    lib().ndto_transform(_f32(Tc), _f32(p), p.shape[0], _f32(out))
    return out


def svd_solve6(H, b):
    H = np.ascontiguousarray(H, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.empty(6, dtype=np.float64)
    lib().ndto_svd_solve6(_f64(H), _f64(b), _f64(x))
    return x


class NormalDistributionsTransform:
    """The oracle object, with the reference's method names (ndt_omp.h:115-238)."""

    def __init__(self):
        self._L = lib()
        self._h = C.c_void_p(self._L.ndto_create())
        self.resolution = 1.0
        self.step_size = 0.1
        self.outlier_ratio = 0.55
        self.trans_eps = 0.1
        self.max_iterations = 35
        self.search_method = DIRECT7
        self.num_threads = 0
        self.min_points_per_voxel = 6
        self.eig_ratio = 0.01
        self._n_target = 0
        self._n_source = 0
        self._push()

    def __del__(self):
        try:
            self._L.ndto_destroy(self._h)
        except Exception:
            pass

    def _push(self):
        self._L.ndto_set_params(self._h, self.resolution, self.step_size, self.outlier_ratio, self.trans_eps,
                                self.max_iterations, self.search_method, self.num_threads, self.min_points_per_voxel,
                                self.eig_ratio)

    # --- setters (names as in the reference) ---
    def setResolution(self, r):
        self.resolution = float(r); self._push()

    def setStepSize(self, s):
        self.step_size = float(s); self._push()

    def setOutlierRatio(self, o):
        self.outlier_ratio = float(o); self._push()

    def setTransformationEpsilon(self, e):
        self.trans_eps = float(e); self._push()

    def setMaximumIterations(self, n):
        self.max_iterations = int(n); self._push()

    def setNeighborhoodSearchMethod(self, m):
        self.search_method = int(m); self._push()

    def setNumThreads(self, n):
        self.num_threads = int(n); self._push()

    def setInputTarget(self, points, is_dense=True):
        p = as_xyzw(points)
        self._n_target = p.shape[0]
        self.build_status = self._L.ndto_set_target(self._h, _f32(p), p.shape[0], 1 if is_dense else 0)
        return self.build_status

    def setInputSource(self, points):
        p = as_xyzw(points)
        self._n_source = p.shape[0]
        self._L.ndto_set_source(self._h, _f32(p), p.shape[0])

    def align(self, guess=None):
        out = np.empty((self._n_source, 4), dtype=np.float32)
        g = None
        if guess is not None:
            g = np.ascontiguousarray(np.asarray(guess, dtype=np.float32).T).reshape(-1)
        self._L.ndto_align(self._h, _f32(g) if g is not None else None, _f32(out))
        return out

    def result(self):
        T = np.empty(16, dtype=np.float32)
        conv, it, ne, nh = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        tp = C.c_double()
        self._L.ndto_get_result(self._h, _f32(T), C.byref(conv), C.byref(it), C.byref(tp), C.byref(ne), C.byref(nh))
        return {"final": T.reshape(4, 4).T.copy(), "converged": bool(conv.value), "iterations": it.value,
                "trans_probability": tp.value, "n_evaluations": ne.value, "n_hessian_passes": nh.value}

    def getFinalTransformation(self):
        return self.result()["final"]

    def hasConverged(self):
        return self.result()["converged"]

    def getFitnessScore(self, max_range=np.finfo(np.float64).max):
        return float(self._L.ndto_fitness(self._h, float(max_range)))

    def calculateScore(self, points):
        p = as_xyzw(points)
        return float(self._L.ndto_calculate_score(self._h, _f32(p), p.shape[0]))

    # --- stage dumps ---
    def gauss(self):
        d = np.empty(3, dtype=np.float64)
        self._L.ndto_gauss(self._h, _f64(d))
        return d

    def map_info(self):
        mn, mx, dv = (np.empty(3, dtype=np.int32) for _ in range(3))
        nl, nv = C.c_long(), C.c_long()
        self._L.ndto_map_info(self._h, _i32(mn), _i32(mx), _i32(dv), C.byref(nl), C.byref(nv))
        return {"min_b": mn, "max_b": mx, "div_b": dv, "n_voxels": nl.value, "n_valid": nv.value}

    def point_keys(self):
        k = np.empty(self._n_target, dtype=np.int32)
        self._L.ndto_point_keys(self._h, _i32(k))
        return k

    def dump_leaves(self):
        n = self.map_info()["n_voxels"]
        keys = np.empty(n, dtype=np.int32)
        counts = np.empty(n, dtype=np.int32)
        mean = np.empty((n, 3), dtype=np.float64)
        cov = np.empty((n, 3, 3), dtype=np.float64)
        icov = np.empty((n, 3, 3), dtype=np.float64)
        infl = np.empty(n, dtype=np.int32)
        self._L.ndto_dump_leaves(self._h, _i32(keys), _i32(counts), _f64(mean), _f64(cov), _f64(icov), _i32(infl))
        return {"keys": keys, "counts": counts, "mean": mean, "cov": cov, "icov": icov, "inflated": infl}

    def eval_derivatives(self, p, T=None, compute_hessian=True):
        p = np.ascontiguousarray(p, dtype=np.float64)
        if T is not None and np.ndim(T) != 2:
            raise TypeError("T must be a 4x4 matrix or None (pass compute_hessian by keyword)")
        Tc = None if T is None else np.ascontiguousarray(np.asarray(T, dtype=np.float32).T).reshape(-1)
        out = np.empty(43, dtype=np.float64)
        hits = self._L.ndto_eval_derivatives(self._h, _f64(p), _f32(Tc) if Tc is not None else None,
                                             1 if compute_hessian else 0, _f64(out))
        return {"score": out[0], "gradient": out[1:7].copy(), "hessian": out[7:].reshape(6, 6).copy(), "hits": hits}

    def eval_hessian(self, p, T=None):
        p = np.ascontiguousarray(p, dtype=np.float64)
        if T is not None and np.ndim(T) != 2:
            raise TypeError("T must be a 4x4 matrix or None (pass compute_hessian by keyword)")
        Tc = None if T is None else np.ascontiguousarray(np.asarray(T, dtype=np.float32).T).reshape(-1)
        out = np.empty(36, dtype=np.float64)
        self._L.ndto_eval_hessian(self._h, _f64(p), _f32(Tc) if Tc is not None else None, _f64(out))
        return out.reshape(6, 6)

    def lookup(self, points, method=None):
        p = as_xyzw(points)
        keys = np.empty((p.shape[0], 26), dtype=np.int32)
        self._L.ndto_lookup(self._h, _f32(p), p.shape[0], self.search_method if method is None else method, _i32(keys))
        return keys

    def trace(self):
        cap = 4096
        kinds = np.empty(cap, dtype=np.int32)
        x = np.empty((cap, 6), dtype=np.float64)
        a = np.empty(cap, dtype=np.float64)
        s = np.empty(cap, dtype=np.float64)
        n = self._L.ndto_trace(self._h, _i32(kinds), _f64(x), _f64(a), _f64(s), cap)
        n = min(n, cap)
        return {"kind": kinds[:n].copy(), "x": x[:n].copy(), "a_t": a[:n].copy(), "score": s[:n].copy()}
