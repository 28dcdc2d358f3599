"""ctypes mirror of pclomp::NormalDistributionsTransform over libndt_b200.so (include/ndt_b200.h)."""
import ctypes as C
import os
import re

import numpy as np

from . import _build

KDTREE, DIRECT26, DIRECT7, DIRECT1 = 0, 1, 2, 3  # pclomp::NeighborSearchMethod (ndt_omp.h:52-57)

OK, ERR_NO_INPUT, ERR_GRID_OVERFLOW, ERR_CUDA, ERR_INVALID, ERR_NO_DEVICE = range(6)
_STATUS_NAMES = {0: "OK", 1: "NO_INPUT", 2: "GRID_OVERFLOW", 3: "CUDA", 4: "INVALID", 5: "NO_DEVICE"}


class NdtError(RuntimeError):
    def __init__(self, status, msg=""):
        super().__init__("ndtb200 status %s%s" % (_STATUS_NAMES.get(status, status), (": " + msg) if msg else ""))
        self.status = status


class Params(C.Structure):
    _fields_ = [("resolution", C.c_float), ("step_size", C.c_double), ("outlier_ratio", C.c_double),
                ("trans_eps", C.c_double), ("max_iterations", C.c_int32), ("search_method", C.c_int32),
                ("min_points_per_voxel", C.c_int32), ("eig_ratio", C.c_double)]


class Result(C.Structure):
    _fields_ = [("final_transformation", C.c_float * 16), ("last_increment", C.c_float * 16),
                ("converged", C.c_int32), ("iterations", C.c_int32), ("trans_probability", C.c_double),
                ("final_pose", C.c_double * 6), ("final_score", C.c_double), ("n_evaluations", C.c_int32),
                ("n_hessian_passes", C.c_int32), ("n_hits", C.c_int64)]


class MapInfo(C.Structure):
    _fields_ = [("min_b", C.c_int32 * 3), ("max_b", C.c_int32 * 3), ("div_b", C.c_int32 * 3),
                ("n_points", C.c_int64), ("n_voxels", C.c_int64), ("n_valid", C.c_int64),
                ("hash_capacity", C.c_int64)]


class MapperStep(C.Structure):
    _fields_ = [("transform", C.c_float * 16), ("pose", C.c_float * 16), ("fitness", C.c_double),
                ("converged", C.c_int32), ("iterations", C.c_int32), ("n_evaluations", C.c_int32), ("pad", C.c_int32),
                ("n_filtered", C.c_int64), ("n_map", C.c_int64)]


_lib = None


def library_path():
    # NDTB200_LIB: development only — lets a tuning run load an alternative build of the SAME library (other CTA shape)
    return os.environ.get("NDTB200_LIB") or _build.LIB_PATH


def exported_symbols():
    """Function names declared in include/ndt_b200.h (the drop-in boundary)."""
    hdr = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "ndt_b200.h")
    with open(hdr) as f:
        text = f.read()
    return sorted(set(re.findall(r"^\s*(?:int|void|void\*|const char\*|int64_t)\s*\*?\s*(ndtb200_[a-z0-9_]+)\s*\(", text, flags=re.M)))


def load_library():
    """Load libndt_b200.so.  Raises if it has not been built — never falls back to anything else."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise NdtError(ERR_NO_DEVICE, "libndt_b200.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(path)
    vp, f32p, f64p, i32p, i64p = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    L.ndtb200_create.argtypes = [C.POINTER(vp), C.c_int]
    L.ndtb200_destroy.argtypes = [vp]
    L.ndtb200_clone.argtypes = [vp, C.POINTER(vp)]
    L.ndtb200_last_error.argtypes = [vp]
    L.ndtb200_last_error.restype = C.c_char_p
    L.ndtb200_device_count.restype = C.c_int
    L.ndtb200_default_params.argtypes = [C.POINTER(Params)]
    L.ndtb200_set_params.argtypes = [vp, C.POINTER(Params)]
    L.ndtb200_get_params.argtypes = [vp, C.POINTER(Params)]
    L.ndtb200_set_target.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_int]
    L.ndtb200_set_source.argtypes = [vp, vp, C.c_size_t, C.c_size_t]
    L.ndtb200_set_target_device.argtypes = [vp, vp, C.c_size_t, C.c_int]
    L.ndtb200_set_source_device.argtypes = [vp, vp, C.c_size_t]
    L.ndtb200_set_target_device_view.argtypes = [vp, vp, C.c_size_t, C.c_int]
    L.ndtb200_align.argtypes = [vp, f32p, vp, C.c_size_t]
    L.ndtb200_align_async.argtypes = [vp, f32p]
    L.ndtb200_sync.argtypes = [vp]
    L.ndtb200_set_throughput_mode.argtypes = [vp, C.c_int]
    L.ndtb200_voxelgrid_filter.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_float, vp, C.c_size_t, C.c_size_t, i64p]
    L.ndtb200_voxelgrid_filter3.argtypes = [vp, vp, C.c_size_t, C.c_size_t, f32p, vp, C.c_size_t, C.c_size_t, i64p]
    L.ndtb200_pose_to_matrix.argtypes = [f64p, f32p]
    L.ndtb200_voxelgrid_filter_device.argtypes = [vp, vp, C.c_size_t, C.c_float, vp, C.c_size_t, i64p]
    L.ndtb200_mapper_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(Params), C.c_float, C.c_float, C.c_int]
    L.ndtb200_mapper_destroy.argtypes = [vp]
    L.ndtb200_mapper_push_scan.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.POINTER(MapperStep)]
    L.ndtb200_mapper_get_map.argtypes = [vp, vp, C.c_size_t, C.c_size_t, i64p]
    L.ndtb200_mapper_last_error.argtypes = [vp]
    L.ndtb200_mapper_last_error.restype = C.c_char_p
    L.ndtb200_mapper_launch_count.argtypes = [vp]
    L.ndtb200_mapper_launch_count.restype = C.c_int64
    L.ndtb200_cloud_bounds.argtypes = [vp, vp, C.c_size_t, C.c_int, f32p, f32p, i64p]
    L.ndtb200_build_partials.argtypes = [vp, f32p, f32p, i64p]
    L.ndtb200_copy_partials.argtypes = [vp, vp, vp, vp]
    L.ndtb200_build_from_partials.argtypes = [vp, f32p, f32p, C.c_int64, vp, vp, vp, C.c_size_t]
    L.ndtb200_partials_split.argtypes = [vp, i32p, C.c_int, i64p]
    L.ndtb200_merge_partials.argtypes = [vp, f32p, f32p, C.c_int64, vp, vp, vp, C.c_size_t, i64p]
    L.ndtb200_copy_records.argtypes = [vp, vp, vp]
    L.ndtb200_set_map_from_records.argtypes = [vp, f32p, f32p, C.c_int64, vp, vp, C.c_size_t]
    L.ndtb200_align_batch.argtypes = [C.POINTER(vp), C.c_int, f32p, C.POINTER(vp), C.c_size_t, C.POINTER(Result)]
    L.ndtb200_align_batch_async.argtypes = [C.POINTER(vp), C.c_int, f32p, C.POINTER(vp), C.c_size_t]
    L.ndtb200_run_pairs.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(vp), C.POINTER(C.c_size_t),
                                    C.c_size_t, f32p, C.c_int, C.POINTER(Result)]
    L.ndtb200_get_result.argtypes = [vp, C.POINTER(Result)]
    L.ndtb200_fitness_score.argtypes = [vp, C.c_double, f64p]
    L.ndtb200_fitness_sums.argtypes = [vp, C.c_double, f64p, i64p]
    L.ndtb200_calculate_score.argtypes = [vp, vp, C.c_size_t, C.c_size_t, f64p]
    L.ndtb200_calculate_score_batch.argtypes = [vp, vp, C.POINTER(C.c_size_t), C.c_int, C.c_size_t, f64p]
    L.ndtb200_score_poses.argtypes = [vp, f32p, C.c_int, f64p]
    L.ndtb200_get_map_info.argtypes = [vp, C.POINTER(MapInfo)]
    L.ndtb200_dump_point_keys.argtypes = [vp, i32p]
    L.ndtb200_dump_voxels.argtypes = [vp, i32p, i32p, f64p, f64p, f64p, i32p]
    L.ndtb200_eval_derivatives.argtypes = [vp, f64p, f32p, C.c_int, f64p, i64p]
    L.ndtb200_eval_hessian.argtypes = [vp, f64p, f32p, f64p]
    L.ndtb200_lookup.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_int, i32p]
    L.ndtb200_get_trace.argtypes = [vp, i32p, f64p, f64p, f64p, C.c_int, C.POINTER(C.c_int)]
    L.ndtb200_get_timeline.argtypes = [vp, f64p, C.c_int, C.POINTER(C.c_int)]
    L.ndtb200_comm_export.argtypes = [vp, vp]
    L.ndtb200_comm_attach.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int64]
    L.ndtb200_comm_detach.argtypes = [vp]
    L.ndtb200_align_emulated_ranks.argtypes = [vp, C.c_int, f32p, C.POINTER(Result)]
    L.ndtb200_debug_guess_to_pose.argtypes = [f32p, f64p]
    L.ndtb200_debug_newton_solve.argtypes = [vp, f64p, f64p, f64p, C.POINTER(C.c_int)]
    L.ndtb200_stream.argtypes = [vp]
    L.ndtb200_stream.restype = vp
    L.ndtb200_launch_count.argtypes = [vp]
    L.ndtb200_launch_count.restype = C.c_int64
    L.ndtb200_reset_launch_count.argtypes = [vp]
    L.ndtb200_reset_launch_count.restype = None
    L.ndtb200_last_align_ms.argtypes = [vp, f32p]
    _lib = L
    return L


class PairPipeline:
    """Batched scan-to-scan odometry through ndtb200_run_pairs: `lanes` NDT objects, each driven by its own C++ host
    thread inside the call (no Python threads, no GIL): pair k = setInputTarget(targets[k]) + setInputSource(sources[k])
    + align(guesses[k]).  Clouds are given as (host pointer, point count) with 16-byte records (pinned memory for speed);
    the ctypes arrays are built once per pair list."""

    def __init__(self, lanes, targets, sources, guesses=None):
        self._L = load_library()
        self.lanes = list(lanes)
        self.n_pairs = len(targets)
        assert len(sources) == self.n_pairs
        self._lanes = (C.c_void_p * len(self.lanes))(*[d._h for d in self.lanes])
        self._tp = (C.c_void_p * max(1, self.n_pairs))(*[int(p) for p, _ in targets])
        self._tn = (C.c_size_t * max(1, self.n_pairs))(*[int(n) for _, n in targets])
        self._sp = (C.c_void_p * max(1, self.n_pairs))(*[int(p) for p, _ in sources])
        self._sn = (C.c_size_t * max(1, self.n_pairs))(*[int(n) for _, n in sources])
        self._g = None if guesses is None else np.concatenate([_colmajor(T) for T in guesses]).astype(np.float32)
        self._res = (Result * max(1, self.n_pairs))()

    def run(self):
        st = self._L.ndtb200_run_pairs(self._lanes, len(self.lanes), self._tp, self._tn, self._sp, self._sn, 16,
                                       _ptr(self._g, C.c_float) if self._g is not None else None, self.n_pairs, self._res)
        if st != 0:
            raise NdtError(st, "ndtb200_run_pairs failed: " + "; ".join(self._L.ndtb200_last_error(d._h).decode() for d in self.lanes[:4]))

    def results(self):
        return [NormalDistributionsTransform._result_dict(r) for r in self._res[:self.n_pairs]]


class Batch:
    """A fixed set of NDT objects aligned together (ndtb200_align_batch): ctypes arrays are built once."""

    def __init__(self, ndts):
        self.ndts = list(ndts)
        self.n = len(self.ndts)
        self._L = load_library()
        self._arr = (C.c_void_p * max(1, self.n))(*[d._h for d in self.ndts])
        self._res = (Result * max(1, self.n))()

    def _guess(self, guesses):
        if guesses is None:
            return None
        return np.concatenate([_colmajor(T) for T in guesses]).astype(np.float32)

    def _outs(self, out_ptrs):
        if out_ptrs is None:
            return None
        return (C.c_void_p * self.n)(*[int(p) for p in out_ptrs])

    def align_async(self, guesses=None, out_ptrs=None, out_stride=16):
        g = self._guess(guesses)
        st = self._L.ndtb200_align_batch_async(self._arr, self.n, _ptr(g, C.c_float) if g is not None else None,
                                               self._outs(out_ptrs), out_stride)
        if st != 0:
            raise NdtError(st, "ndtb200_align_batch_async failed")

    def sync(self):
        for d in self.ndts:
            d.sync()

    def align(self, guesses=None, out_ptrs=None, out_stride=16):
        g = self._guess(guesses)
        st = self._L.ndtb200_align_batch(self._arr, self.n, _ptr(g, C.c_float) if g is not None else None,
                                         self._outs(out_ptrs), out_stride, self._res)
        if st != 0:
            raise NdtError(st, "ndtb200_align_batch failed: " + "; ".join(
                self._L.ndtb200_last_error(d._h).decode() for d in self.ndts[:4]))
        return [d.result() for d in self.ndts]


class Mapper:
    """The mapping-node loop (ndt_rosbag_mapping_node.cpp:42-161) as a device-resident pipeline: push raw scans, get
    per-step transforms / poses; the global map stays on the device until asked for."""

    def __init__(self, device=0, voxel_leaf=0.3, map_voxel=0.5, compute_fitness=True, **ndt_params):
        self._L = load_library()
        self._h = C.c_void_p()
        p = Params()
        self._L.ndtb200_default_params(C.byref(p))
        p.trans_eps, p.max_iterations = 0.01, 64     # the node's defaults
        for k, v in ndt_params.items():
            setattr(p, k, v)
        st = self._L.ndtb200_mapper_create(C.byref(self._h), int(device), C.byref(p), float(voxel_leaf), float(map_voxel),
                                           1 if compute_fitness else 0)
        if st != OK:
            self._h = C.c_void_p()
            raise NdtError(st, "ndtb200_mapper_create failed")

    def __del__(self):
        try:
            if self._h:
                self._L.ndtb200_mapper_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def push_scan_raw(self, ptr, n, stride=16):
        s = MapperStep()
        st = self._L.ndtb200_mapper_push_scan(self._h, ptr, n, stride, C.byref(s))
        if st != OK:
            raise NdtError(st, self._L.ndtb200_mapper_last_error(self._h).decode())
        return {"transform": np.array(s.transform, dtype=np.float32).reshape(4, 4).T.copy(),
                "pose": np.array(s.pose, dtype=np.float32).reshape(4, 4).T.copy(), "fitness": float(s.fitness),
                "converged": bool(s.converged), "iterations": int(s.iterations), "n_evaluations": int(s.n_evaluations),
                "n_filtered": int(s.n_filtered), "n_map": int(s.n_map)}

    def push_scan(self, points):
        p = as_xyzw(points)
        self._keep = p
        return self.push_scan_raw(p.ctypes.data, p.shape[0], 16)

    def global_map(self):
        n = C.c_int64(0)
        self._L.ndtb200_mapper_get_map(self._h, None, 0, 16, C.byref(n))
        out = np.empty((max(1, n.value), 4), dtype=np.float32)
        st = self._L.ndtb200_mapper_get_map(self._h, out.ctypes.data, out.shape[0], 16, C.byref(n))
        if st != OK:
            raise NdtError(st, self._L.ndtb200_mapper_last_error(self._h).decode())
        return np.ascontiguousarray(out[:n.value, :3])

    def launch_count(self):
        return int(self._L.ndtb200_mapper_launch_count(self._h))


def align_batch(ndts, guesses=None):
    """ndtb200_align_batch: align n independent (target, source) pairs — one object each — together.
    guesses: optional list of 4x4 matrices.  Returns the list of result dicts."""
    if len(ndts) == 0:
        return []
    return Batch(ndts).align(guesses)


def device_count():
    return int(load_library().ndtb200_device_count())


def pose_to_matrix(p):
    """ndtb200_pose_to_matrix = static convertTransform (ndt_omp.h:216-233).  Needs no device."""
    p = np.ascontiguousarray(p, dtype=np.float64)
    out = np.zeros(16, dtype=np.float32)
    st = load_library().ndtb200_pose_to_matrix(_ptr(p, C.c_double), _ptr(out, C.c_float))
    if st != OK:
        raise NdtError(st, "ndtb200_pose_to_matrix failed")
    return out.reshape(4, 4).T.copy()


def guess_to_pose(guess):
    """ndtb200_debug_guess_to_pose: [translation, rotation().eulerAngles(0,1,2)] of a 4x4 guess as the align entry point
    computes it on the host (ndt_omp_impl.hpp:103-111).  Needs no device."""
    g = _colmajor(guess)
    p = np.zeros(6, dtype=np.float64)
    st = load_library().ndtb200_debug_guess_to_pose(_ptr(g, C.c_float), _ptr(p, C.c_double))
    if st != OK:
        raise NdtError(st, "ndtb200_debug_guess_to_pose failed")
    return p


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def as_xyzw(points):
    """(n,3)/(n,4) array -> contiguous (n,4) float32, w = 1: the pcl::PointXYZ 16-byte layout."""
    p = np.asarray(points)
    if p.ndim != 2 or p.shape[1] not in (3, 4):
        raise ValueError("points must be (n,3) or (n,4)")
    if p.dtype == np.float32 and p.shape[1] == 4 and p.flags["C_CONTIGUOUS"]:
        return p
    out = np.ones((p.shape[0], 4), dtype=np.float32)
    out[:, :3] = p[:, :3]
    return out


def _colmajor(T):
    return np.ascontiguousarray(np.asarray(T, dtype=np.float32).T).reshape(-1)


class NormalDistributionsTransform:
    """Same method names and meaning as pclomp::NormalDistributionsTransform (ndt_omp.h:70-502)."""

    def __init__(self, device=0):
        self._L = load_library()
        self._h = C.c_void_p()
        st = self._L.ndtb200_create(C.byref(self._h), int(device))
        if st != OK:
            self._h = C.c_void_p()
            raise NdtError(st, "ndtb200_create failed (no CUDA device? this library has no CPU fallback)")
        self._p = Params()
        self._L.ndtb200_get_params(self._h, C.byref(self._p))
        self._n_source = 0
        self._n_target = 0
        self._keep = {}

    def __del__(self):
        try:
            if self._h:
                self._L.ndtb200_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # ---- helpers ----
    def _check(self, st, allow=()):
        if st != OK and st not in allow:
            raise NdtError(st, self._L.ndtb200_last_error(self._h).decode())
        return st

    def _push(self):
        return self._check(self._L.ndtb200_set_params(self._h, C.byref(self._p)), allow=(ERR_NO_INPUT, ERR_GRID_OVERFLOW))

    def clone(self):
        """Copy construction (the mapping node returns the object by value, ndt_omp_mapping_node.cpp:151-169)."""
        other = object.__new__(NormalDistributionsTransform)
        other._L = self._L
        other._h = C.c_void_p()
        self._check(self._L.ndtb200_clone(self._h, C.byref(other._h)))
        other._p = Params()
        self._L.ndtb200_get_params(other._h, C.byref(other._p))
        other._n_source, other._n_target, other._keep = self._n_source, self._n_target, {}
        return other

    # ---- setters / getters (reference names) ----
    def setResolution(self, r):
        self._p.resolution = float(r); self._push()

    def getResolution(self):
        return float(self._p.resolution)

    def setStepSize(self, s):
        self._p.step_size = float(s); self._push()

    def getStepSize(self):
        return float(self._p.step_size)

    def setOutlierRatio(self, o):
        self._p.outlier_ratio = float(o); self._push()

    def getOutlierRatio(self):
        return float(self._p.outlier_ratio)

    def setTransformationEpsilon(self, e):
        self._p.trans_eps = float(e); self._push()

    def setMaximumIterations(self, n):
        self._p.max_iterations = int(n); self._push()

    def setNeighborhoodSearchMethod(self, m):
        self._p.search_method = int(m); self._push()

    def setNumThreads(self, n):
        """Accepted and ignored: the device path has no host thread count (ndt_omp.h:115-117)."""

    def setMinPointPerVoxel(self, n):
        self._p.min_points_per_voxel = int(n); self._push()

    def setCovEigValueInflationRatio(self, r):
        self._p.eig_ratio = float(r); self._push()

    def setInputTarget(self, points, is_dense=True):
        p = as_xyzw(points)
        self._n_target = p.shape[0]
        st = self._L.ndtb200_set_target(self._h, p.ctypes.data, p.shape[0], 16, 1 if is_dense else 0)
        self.build_status = self._check(st, allow=(ERR_NO_INPUT, ERR_GRID_OVERFLOW))
        return self.build_status

    def setInputSource(self, points):
        p = as_xyzw(points)
        self._n_source = p.shape[0]
        self._check(self._L.ndtb200_set_source(self._h, p.ctypes.data, p.shape[0], 16))

    def set_target_raw(self, ptr, n, stride, is_dense=True):
        """Host pointer + stride, exactly as the C ABI takes it (pinned buffers, 32-byte point types...)."""
        self._n_target = int(n)
        self.build_status = self._check(self._L.ndtb200_set_target(self._h, ptr, n, stride, 1 if is_dense else 0),
                                        allow=(ERR_NO_INPUT, ERR_GRID_OVERFLOW))
        return self.build_status

    def set_source_raw(self, ptr, n, stride):
        self._n_source = int(n)
        self._check(self._L.ndtb200_set_source(self._h, ptr, n, stride))

    def set_target_device(self, dev_ptr, n, is_dense=True):
        self._n_target = int(n)
        self.build_status = self._check(self._L.ndtb200_set_target_device(self._h, dev_ptr, n, 1 if is_dense else 0),
                                        allow=(ERR_NO_INPUT, ERR_GRID_OVERFLOW))
        return self.build_status

    def set_target_device_view(self, dev_ptr, n, is_dense=True):
        """setInputTarget without a copy: the caller's device buffer must outlive the map (ndtb200_set_target_device_view)."""
        self._n_target = int(n)
        self.build_status = self._check(self._L.ndtb200_set_target_device_view(self._h, dev_ptr, n, 1 if is_dense else 0),
                                        allow=(ERR_NO_INPUT, ERR_GRID_OVERFLOW))
        return self.build_status

    def set_source_device(self, dev_ptr, n):
        self._n_source = int(n)
        self._check(self._L.ndtb200_set_source_device(self._h, dev_ptr, n))

    # ---- registration ----
    def align(self, guess=None, want_output=True, out=None):
        """align(output[, guess]).  Returns the (n,4) transformed source (or None)."""
        g = _colmajor(guess) if guess is not None else None
        if want_output and out is None:
            out = np.empty((self._n_source, 4), dtype=np.float32)
        self._check(self._L.ndtb200_align(self._h, _ptr(g, C.c_float) if g is not None else None,
                                          out.ctypes.data if want_output else None, 16))
        return out if want_output else None

    def align_raw(self, guess_ptr, out_ptr, out_stride):
        self._check(self._L.ndtb200_align(self._h, guess_ptr, out_ptr, out_stride))

    def align_async(self, guess=None):
        g = _colmajor(guess) if guess is not None else None
        self._check(self._L.ndtb200_align_async(self._h, _ptr(g, C.c_float) if g is not None else None))

    # ---- pcl::VoxelGrid centroid downsample on the device ----
    def voxelgrid_filter(self, points, leaf):
        """(n,3|4) cloud -> (m,3) float32 centroids, bit-identical to pcl::VoxelGrid (ascending cell index).
        `leaf`: scalar or (lx, ly, lz)."""
        p = as_xyzw(points)
        out = np.empty((max(1, p.shape[0]), 4), dtype=np.float32)
        m = C.c_int64(0)
        if np.ndim(leaf) == 0:
            st = self._L.ndtb200_voxelgrid_filter(self._h, p.ctypes.data, p.shape[0], 16, float(leaf), out.ctypes.data, out.shape[0], 16, C.byref(m))
        else:
            l3 = np.ascontiguousarray(leaf, dtype=np.float32)
            st = self._L.ndtb200_voxelgrid_filter3(self._h, p.ctypes.data, p.shape[0], 16, _ptr(l3, C.c_float), out.ctypes.data, out.shape[0], 16, C.byref(m))
        self._check(st)
        return np.ascontiguousarray(out[:m.value, :3])

    def voxelgrid_filter_device(self, dev_ptr, n, leaf, out_dev_ptr, out_capacity):
        m = C.c_int64(0)
        self._check(self._L.ndtb200_voxelgrid_filter_device(self._h, dev_ptr, n, float(leaf), out_dev_ptr, out_capacity, C.byref(m)))
        return int(m.value)

    # ---- sharded target-map build (ndtb200_cloud_bounds / build_partials / copy_partials / build_from_partials) ----
    def cloud_bounds(self, dev_ptr, n, is_dense=True):
        mn, mx = np.zeros(3, np.float32), np.zeros(3, np.float32)
        nf = C.c_int64(0)
        self._n_target = int(n)
        self._check(self._L.ndtb200_cloud_bounds(self._h, dev_ptr, n, 1 if is_dense else 0, _ptr(mn, C.c_float), _ptr(mx, C.c_float), C.byref(nf)))
        return mn, mx, int(nf.value)

    def build_partials(self, gmin, gmax):
        gmin, gmax = np.ascontiguousarray(gmin, np.float32), np.ascontiguousarray(gmax, np.float32)
        nv = C.c_int64(0)
        st = self._L.ndtb200_build_partials(self._h, _ptr(gmin, C.c_float), _ptr(gmax, C.c_float), C.byref(nv))
        self._check(st, allow=(ERR_GRID_OVERFLOW,))
        return st, int(nv.value)

    def copy_partials(self, keys_ptr, counts_ptr, moments_ptr):
        self._check(self._L.ndtb200_copy_partials(self._h, keys_ptr, counts_ptr, moments_ptr))

    def build_from_partials(self, gmin, gmax, n_finite_total, keys_ptr, counts_ptr, moments_ptr, n_total):
        gmin, gmax = np.ascontiguousarray(gmin, np.float32), np.ascontiguousarray(gmax, np.float32)
        return self._check(self._L.ndtb200_build_from_partials(self._h, _ptr(gmin, C.c_float), _ptr(gmax, C.c_float), int(n_finite_total),
                                                               keys_ptr, counts_ptr, moments_ptr, int(n_total)),
                           allow=(ERR_NO_INPUT, ERR_GRID_OVERFLOW))

    # ---- owner-partitioned sharded build (ndtb200_partials_split / merge_partials / copy_records / set_map_from_records) ----
    def partials_split(self, upper_keys, world):
        """Offsets (world + 1) of this handle's key-sorted partials at the owners' key boundaries."""
        up = np.ascontiguousarray(upper_keys, dtype=np.int32)
        offs = np.zeros(world + 1, dtype=np.int64)
        self._check(self._L.ndtb200_partials_split(self._h, _ptr(up, C.c_int32) if world > 1 else None, int(world), _ptr(offs, C.c_int64)))
        return offs

    def merge_partials(self, gmin, gmax, n_finite_total, keys_ptr, counts_ptr, moments_ptr, n_total):
        gmin, gmax = np.ascontiguousarray(gmin, np.float32), np.ascontiguousarray(gmax, np.float32)
        n = C.c_int64(0)
        st = self._check(self._L.ndtb200_merge_partials(self._h, _ptr(gmin, C.c_float), _ptr(gmax, C.c_float), int(n_finite_total),
                                                        keys_ptr, counts_ptr, moments_ptr, int(n_total), C.byref(n)),
                         allow=(ERR_NO_INPUT, ERR_GRID_OVERFLOW))
        return st, int(n.value)

    def copy_records(self, records_ptr, icov64_ptr):
        self._check(self._L.ndtb200_copy_records(self._h, records_ptr, icov64_ptr))

    def set_map_from_records(self, gmin, gmax, n_finite_total, records_ptr, icov64_ptr, n_total):
        gmin, gmax = np.ascontiguousarray(gmin, np.float32), np.ascontiguousarray(gmax, np.float32)
        return self._check(self._L.ndtb200_set_map_from_records(self._h, _ptr(gmin, C.c_float), _ptr(gmax, C.c_float), int(n_finite_total),
                                                                records_ptr, icov64_ptr, int(n_total)),
                           allow=(ERR_NO_INPUT, ERR_GRID_OVERFLOW))

    def set_throughput_mode(self, on=True):
        """Small-CTA solve kernel: several handles' solves share every SM (see ndtb200_set_throughput_mode)."""
        self._check(self._L.ndtb200_set_throughput_mode(self._h, 1 if on else 0))

    def sync(self):
        self._check(self._L.ndtb200_sync(self._h))

    @staticmethod
    def _result_dict(r):
        return {"final": np.array(r.final_transformation, dtype=np.float32).reshape(4, 4).T.copy(),
                "last_increment": np.array(r.last_increment, dtype=np.float32).reshape(4, 4).T.copy(),
                "converged": bool(r.converged), "iterations": int(r.iterations),
                "trans_probability": float(r.trans_probability), "final_pose": np.array(r.final_pose),
                "final_score": float(r.final_score), "n_evaluations": int(r.n_evaluations),
                "n_hessian_passes": int(r.n_hessian_passes), "n_hits": int(r.n_hits)}

    def result(self):
        r = Result()
        self._check(self._L.ndtb200_get_result(self._h, C.byref(r)))
        return self._result_dict(r)

    def align_emulated_ranks(self, world, guess=None):
        """ndtb200_align_emulated_ranks: the source-sharded solve with `world` ranks inside one cooperative launch on this
        GPU.  Returns one result dict per rank."""
        g = _colmajor(guess) if guess is not None else None
        res = (Result * int(world))()
        self._check(self._L.ndtb200_align_emulated_ranks(self._h, int(world), _ptr(g, C.c_float) if g is not None else None, res))
        return [self._result_dict(r) for r in res]

    def newton_solve(self, H, g):
        """ndtb200_debug_newton_solve: delta = solve(H, -g) by the kernel's warp solvers.  Returns (delta, path)."""
        H = np.ascontiguousarray(H, dtype=np.float64).reshape(36)
        g = np.ascontiguousarray(g, dtype=np.float64).reshape(6)
        d = np.zeros(6, dtype=np.float64)
        path = C.c_int(-1)
        self._check(self._L.ndtb200_debug_newton_solve(self._h, _ptr(H, C.c_double), _ptr(g, C.c_double), _ptr(d, C.c_double), C.byref(path)))
        return d, int(path.value)

    def getFinalTransformation(self):
        return self.result()["final"]

    def hasConverged(self):
        return self.result()["converged"]

    def getFinalNumIteration(self):
        return self.result()["iterations"]

    def getTransformationProbability(self):
        return self.result()["trans_probability"]

    def getFitnessScore(self, max_range=np.finfo(np.float64).max):
        v = C.c_double()
        self._check(self._L.ndtb200_fitness_score(self._h, float(max_range), C.byref(v)))
        return v.value

    def fitness_sums(self, max_range=np.finfo(np.float64).max):
        """(sum of accepted squared nearest-neighbour distances, count) of this handle's source: for sharded sources."""
        s, c = C.c_double(0), C.c_int64(0)
        self._check(self._L.ndtb200_fitness_sums(self._h, float(max_range), C.byref(s), C.byref(c)))
        return float(s.value), int(c.value)

    def calculateScore(self, points):
        p = as_xyzw(points)
        v = C.c_double()
        self._check(self._L.ndtb200_calculate_score(self._h, p.ctypes.data, p.shape[0], 16, C.byref(v)))
        return v.value

    def calculateScoreBatch(self, clouds):
        """ndtb200_calculate_score_batch: calculateScore of every (already transformed) cloud of the list, one call."""
        if len(clouds) == 0:
            return np.zeros(0)
        ps = [as_xyzw(c) for c in clouds]
        offs = np.zeros(len(ps) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([p.shape[0] for p in ps])
        allp = np.ascontiguousarray(np.concatenate(ps)) if offs[-1] else np.zeros((1, 4), np.float32)
        out = np.zeros(len(ps), dtype=np.float64)
        offs_c = (C.c_size_t * len(offs))(*[int(o) for o in offs])
        self._check(self._L.ndtb200_calculate_score_batch(self._h, allp.ctypes.data, offs_c, len(ps), 16, _ptr(out, C.c_double)))
        return out

    def scorePoses(self, poses):
        """ndtb200_score_poses: calculateScore of the current source under every candidate 4x4 pose."""
        if len(poses) == 0:
            return np.zeros(0)
        g = np.concatenate([_colmajor(T) for T in poses]).astype(np.float32)
        out = np.zeros(len(poses), dtype=np.float64)
        self._check(self._L.ndtb200_score_poses(self._h, _ptr(g, C.c_float), len(poses), _ptr(out, C.c_double)))
        return out

    # ---- stage dumps (parity API) ----
    def map_info(self):
        m = MapInfo()
        st = self._L.ndtb200_get_map_info(self._h, C.byref(m))
        return {"status": st, "min_b": np.array(m.min_b), "max_b": np.array(m.max_b), "div_b": np.array(m.div_b),
                "n_points": m.n_points, "n_voxels": m.n_voxels, "n_valid": m.n_valid, "hash_capacity": m.hash_capacity}

    def point_keys(self):
        k = np.empty(self._n_target, dtype=np.int32)
        self._check(self._L.ndtb200_dump_point_keys(self._h, _ptr(k, C.c_int32)))
        return k

    def dump_voxels(self):
        n = self.map_info()["n_voxels"]
        keys = np.empty(n, dtype=np.int32)
        counts = np.empty(n, dtype=np.int32)
        mean = np.empty((n, 3), dtype=np.float64)
        cov = np.empty((n, 3, 3), dtype=np.float64)
        icov = np.empty((n, 3, 3), dtype=np.float64)
        infl = np.empty(n, dtype=np.int32)
        self._check(self._L.ndtb200_dump_voxels(self._h, _ptr(keys, C.c_int32), _ptr(counts, C.c_int32),
                                                _ptr(mean, C.c_double), _ptr(cov, C.c_double), _ptr(icov, C.c_double),
                                                _ptr(infl, C.c_int32)))
        return {"keys": keys, "counts": counts, "mean": mean, "cov": cov, "icov": icov, "inflated": infl}

    def eval_derivatives(self, p, T=None, compute_hessian=True):
        p = np.ascontiguousarray(p, dtype=np.float64)
        if T is not None and np.ndim(T) != 2:
            raise TypeError("T must be a 4x4 matrix or None (pass compute_hessian by keyword)")
        Tc = None if T is None else _colmajor(T)
        out = np.empty(43, dtype=np.float64)
        hits = C.c_int64()
        self._check(self._L.ndtb200_eval_derivatives(self._h, _ptr(p, C.c_double),
                                                     _ptr(Tc, C.c_float) if Tc is not None else None,
                                                     1 if compute_hessian else 0, _ptr(out, C.c_double), C.byref(hits)))
        return {"score": out[0], "gradient": out[1:7].copy(), "hessian": out[7:].reshape(6, 6).copy(), "hits": hits.value}

    def eval_hessian(self, p, T=None):
        p = np.ascontiguousarray(p, dtype=np.float64)
        if T is not None and np.ndim(T) != 2:
            raise TypeError("T must be a 4x4 matrix or None (pass compute_hessian by keyword)")
        Tc = None if T is None else _colmajor(T)
        out = np.empty(36, dtype=np.float64)
        self._check(self._L.ndtb200_eval_hessian(self._h, _ptr(p, C.c_double),
                                                 _ptr(Tc, C.c_float) if Tc is not None else None, _ptr(out, C.c_double)))
        return out.reshape(6, 6)

    def lookup(self, points, method=None):
        p = as_xyzw(points)
        keys = np.empty((p.shape[0], 26), dtype=np.int32)
        self._check(self._L.ndtb200_lookup(self._h, p.ctypes.data, p.shape[0], 16,
                                           self._p.search_method if method is None else int(method),
                                           _ptr(keys, C.c_int32)))
        return keys

    def trace(self):
        cap = 1024
        kinds = np.empty(cap, dtype=np.int32)
        x = np.empty((cap, 6), dtype=np.float64)
        a = np.empty(cap, dtype=np.float64)
        s = np.empty(cap, dtype=np.float64)
        n = C.c_int()
        self._check(self._L.ndtb200_get_trace(self._h, _ptr(kinds, C.c_int32), _ptr(x, C.c_double), _ptr(a, C.c_double),
                                              _ptr(s, C.c_double), cap, C.byref(n)))
        m = min(n.value, cap)
        return {"kind": kinds[:m].copy(), "x": x[:m].copy(), "a_t": a[:m].copy(), "score": s[:m].copy()}

    def timeline(self):
        cap = 1024
        t = np.empty((cap, 4), dtype=np.float64)
        n = C.c_int()
        self._check(self._L.ndtb200_get_timeline(self._h, _ptr(t, C.c_double), cap, C.byref(n)))
        return t[:min(n.value, cap)].copy()

    # ---- multi-GPU source sharding ----
    def comm_export(self):
        buf = C.create_string_buffer(64)
        self._check(self._L.ndtb200_comm_export(self._h, buf))
        return buf.raw

    def comm_attach(self, rank, world, handles, n_source_total):
        blob = b"".join(handles)
        assert len(blob) == 64 * world
        self._check(self._L.ndtb200_comm_attach(self._h, int(rank), int(world), blob, int(n_source_total)))

    def comm_detach(self):
        self._check(self._L.ndtb200_comm_detach(self._h))

    # ---- plumbing ----
    def stream_ptr(self):
        return int(self._L.ndtb200_stream(self._h) or 0)

    def launch_count(self):
        return int(self._L.ndtb200_launch_count(self._h))

    def reset_launch_count(self):
        self._L.ndtb200_reset_launch_count(self._h)

    def last_align_ms(self):
        v = C.c_float()
        self._check(self._L.ndtb200_last_align_ms(self._h, C.byref(v)))
        return v.value
