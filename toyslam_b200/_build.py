"""In-tree build of libndt_b200.so (sm_100a only) with nvcc.  No JIT cache, no torch extension:
the .so lands in toyslam_b200/lib/ and travels to the GPU box with the repo snapshot."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libndt_b200.so")
SOURCES = ["capi.cu"]
HEADERS = ["common.cuh", "map_build.cuh", "ndt_solve.cuh", "ndt_align.cuh", "ndt_aux.cuh",
           os.path.join("..", "..", "include", "ndt_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile the CUDA library for sm_100a.  Cross-compiles without a GPU."""
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    env = dict(os.environ)
    # the image exports CC/CXX=/opt/gcc/bin/* wrappers; nvcc wants the stock host compiler
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    subprocess.check_call(cmd, cwd=CSRC, env=env)
    return LIB_PATH


CHECKED_LIB_PATH = os.path.join(LIB_DIR, "libndt_b200_checked.so")


def build_checked(force=False):
    """The same library with -DNDTB200_CHECKED: device-side bounds / protocol assertions (NDT_CHECK) in the kernels.
    Run the parity suite on it with NDTB200_LIB=toyslam_b200/lib/libndt_b200_checked.so (compute-sanitizer is closed on
    this GPU pool)."""
    if not force and os.path.exists(CHECKED_LIB_PATH) and os.path.getmtime(CHECKED_LIB_PATH) >= max(
            os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC)):
        return CHECKED_LIB_PATH
    cmd = [_nvcc()] + (["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []) + NVCC_FLAGS + ["-DNDTB200_CHECKED", "-o", CHECKED_LIB_PATH] + SOURCES
    subprocess.check_call(cmd, cwd=CSRC)
    return CHECKED_LIB_PATH


APP_SRC = os.path.join(_HERE, "..", "apps", "align_b200.cpp")
APP_BIN = os.path.join(_HERE, "..", "apps", "align_b200")
REPLAY_SRC = os.path.join(_HERE, "..", "apps", "replay_b200.cpp")
REPLAY_BIN = os.path.join(_HERE, "..", "apps", "replay_b200")


def _build_app(src, out, force):
    inc = os.path.join(_HERE, "..", "include")
    hdrs = [os.path.join(inc, "ndt_b200.h")] + [os.path.join(inc, "pclomp_b200", h) for h in os.listdir(os.path.join(inc, "pclomp_b200"))]
    deps = [src, LIB_PATH] + hdrs
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-I", inc, src, "-o", out, "-L", LIB_DIR, "-lndt_b200", "-Wl,-rpath,$ORIGIN/../toyslam_b200/lib"]
    subprocess.check_call(cmd)
    return out


SHARDED_SRC = os.path.join(_HERE, "..", "apps", "sharded_b200.cpp")
SHARDED_BIN = os.path.join(_HERE, "..", "apps", "sharded_b200")


def build_sharded_app(force=False):
    """apps/sharded_b200: the multi-GPU paths through the C++ host (pclomp_b200::ShardedNdt over NCCL).  Needs nccl.h /
    libnccl and the CUDA runtime headers; returns None where they are missing."""
    cuda_inc = "/usr/local/cuda/include"
    if not (os.path.exists("/usr/include/nccl.h") or os.path.exists(os.path.join(cuda_inc, "nccl.h"))):
        return None
    inc = os.path.join(_HERE, "..", "include")
    deps = [SHARDED_SRC, LIB_PATH, os.path.join(inc, "pclomp_b200", "sharded_ndt.hpp"), os.path.join(inc, "ndt_b200.h")]
    if not force and os.path.exists(SHARDED_BIN) and all(os.path.getmtime(d) <= os.path.getmtime(SHARDED_BIN) for d in deps):
        return SHARDED_BIN
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-I", inc, "-I", cuda_inc, SHARDED_SRC, "-o", SHARDED_BIN, "-L", LIB_DIR, "-lndt_b200",
           "-L", "/usr/local/cuda/lib64", "-lcudart", "-lnccl", "-Wl,-rpath,$ORIGIN/../toyslam_b200/lib"]
    try:
        subprocess.check_call(cmd)
    except subprocess.CalledProcessError:
        return None
    return SHARDED_BIN


def build_apps(force=False):
    """The C++ programs over the header-only shim: apps/align_b200 (the reference's benchmark app: proves the
    reference-named API compiles and links) and apps/replay_b200 (PointCloud2 dump -> mapping loop).  Returns align_b200."""
    _build_app(REPLAY_SRC, REPLAY_BIN, force)
    build_sharded_app(force)
    return _build_app(APP_SRC, APP_BIN, force)


if __name__ == "__main__":
    print(build(force=True, verbose=True))
