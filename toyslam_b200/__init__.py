"""toyslam_b200 — B200-native NDT registration hot path (ToySLAM / ndt_omp), Python host mirror.

The product is libndt_b200.so (hand-written sm_100a CUDA behind the C ABI of include/ndt_b200.h)
and the C++ shim include/pclomp_b200/ndt_b200.hpp.  This package is the ctypes mirror of the
reference's class interface (same method names as pclomp::NormalDistributionsTransform,
ndt_omp/include/pclomp/ndt_omp.h:115-238) used by the parity tests and bench.py.

There is no CPU fallback: constructing an object without the CUDA library or without a GPU raises.
"""
from .ndt import (  # noqa: F401
    DIRECT1,
    DIRECT7,
    DIRECT26,
    KDTREE,
    NdtError,
    Batch,
    Mapper,
    NormalDistributionsTransform,
    align_batch,
    PairPipeline,
    device_count,
    exported_symbols,
    guess_to_pose,
    pose_to_matrix,
    library_path,
    load_library,
)

__all__ = ["NormalDistributionsTransform", "NdtError", "KDTREE", "DIRECT26", "DIRECT7", "DIRECT1",
           "Batch", "PairPipeline", "Mapper", "align_batch", "device_count", "load_library", "library_path", "exported_symbols", "guess_to_pose", "pose_to_matrix"]
