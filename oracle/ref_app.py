"""Recipe: compile the NDT section of the REFERENCE's own benchmark app (ndt_omp/apps/align.cpp, read from
/root/reference where it lies — never copied into the repo) against pclomp_b200::NormalDistributionsTransform.

This is the drop-in proof the boundary owes (BASELINE north_star: "the ndt_omp apps ... can switch to it with no other
changes"): the translation unit is the reference file with

  * the include of the class swapped   `#include <pclomp/ndt_omp.h>`  ->  `#include <pclomp_b200/ndt_b200.hpp>`
  * the namespace swapped              `pclomp::`                     ->  `pclomp_b200::`
  * the parts of the app that are out of scope removed: the GICP / pcl::NDT benchmark blocks
    (:73-86, which need upstream PCL's own registration classes), their three includes (:6, :7, :12), the
    PCLVisualizer block (:107-115) and its include (:9); the declaration of `aligned`, which lived in the removed GICP
    block (:76), is re-inserted as a plain declaration.

Everything else — the `align(pcl::Registration<...>::Ptr, ...)` helper (:15-33), main's loading and 0.1 m
pcl::VoxelGrid downsampling (:36-71) and the pclomp::NDT benchmark loop (:88-105) — is compiled byte for byte.
PCL / ROS / Eigen are not installed here: tests/stubs/minipcl supplies PCL-shaped headers (see its README).

Outputs go to oracle/_ref/ (git-ignored; travels to the GPU box with the repo snapshot like every built binary).
"""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_APP = "/root/reference/ndt_omp/apps/align.cpp"
OUT_DIR = os.path.join(ROOT, "oracle", "_ref")
OUT_SRC = os.path.join(OUT_DIR, "align_ndt_section.cpp")
OUT_BIN = os.path.join(OUT_DIR, "align_ndt_section")

DROP_LINES = set([6, 7, 9, 12]) | set(range(73, 88)) | set(range(106, 117))   # 1-based, see the docstring


def available():
    return os.path.exists(REF_APP)


def generate():
    with open(REF_APP) as f:
        lines = f.read().split("\n")
    out = []
    for i, line in enumerate(lines, start=1):
        if i in DROP_LINES:
            if i == 76:  # `aligned` was declared by the removed GICP block
                out.append("  pcl::PointCloud<pcl::PointXYZ>::Ptr aligned;")
            continue
        if i == 11:
            assert line.strip() == "#include <pclomp/ndt_omp.h>", line
            line = "#include <pclomp_b200/ndt_b200.hpp>"
        out.append(line.replace("pclomp::", "pclomp_b200::"))
    os.makedirs(OUT_DIR, exist_ok=True)
    with open(OUT_SRC, "w") as f:
        f.write("\n".join(out))
    return OUT_SRC


def build(force=False):
    """Returns the path of the binary, or None when the reference is not present (GPU box: prebuilt file travels)."""
    if not available():
        return OUT_BIN if os.path.exists(OUT_BIN) else None
    lib_dir = os.path.join(ROOT, "toyslam_b200", "lib")
    deps = [REF_APP, os.path.join(lib_dir, "libndt_b200.so"), os.path.abspath(__file__)]
    for d, _, fs in os.walk(os.path.join(ROOT, "include")):
        deps += [os.path.join(d, f) for f in fs]
    for d, _, fs in os.walk(os.path.join(ROOT, "tests", "stubs", "minipcl")):
        deps += [os.path.join(d, f) for f in fs]
    if not force and os.path.exists(OUT_BIN) and all(os.path.getmtime(d) <= os.path.getmtime(OUT_BIN) for d in deps if os.path.exists(d)):
        return OUT_BIN
    src = generate()
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-fopenmp", "-Wall", "-I", os.path.join(ROOT, "tests", "stubs", "minipcl"),
           "-I", os.path.join(ROOT, "include"), src, "-o", OUT_BIN, "-L", lib_dir, "-lndt_b200",
           "-Wl,-rpath,$ORIGIN/../../toyslam_b200/lib"]
    subprocess.check_call(cmd)
    return OUT_BIN


if __name__ == "__main__":
    print(build(force=True))
