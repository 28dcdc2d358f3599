"""The C++ drop-in boundary: the shim class derives from pcl::Registration like the reference class
(ndt_omp.h:70-71), so a caller written against `pcl::Registration<...>::Ptr` compiles unchanged.

CPU tests compile; the GPU tests run the binaries (`-m gpu`)."""
import os
import re
import subprocess

import numpy as np
import pytest

from util import GOLDEN, ROOT, golden

INC = os.path.join(ROOT, "include")
MINIPCL = os.path.join(ROOT, "tests", "stubs", "minipcl")
LIBDIR = os.path.join(ROOT, "toyslam_b200", "lib")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def compile_cpp(src, out, extra=()):
    cmd = [CXX, "-O1", "-std=c++17", "-fopenmp", "-Wall", "-Werror=return-type", *extra, "-I", INC, src, "-o", out,
           "-L", LIBDIR, "-lndt_b200", "-Wl,-rpath," + LIBDIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


SURFACE_TU = r'''
// every public member of the reference class (ndt_omp.h:96-238, 499) used the way its callers use them
#include <pclomp_b200/ndt_b200.hpp>
typedef pcl::PointCloud<pcl::PointXYZ> Cloud;
typedef pclomp_b200::NormalDistributionsTransform<pcl::PointXYZ, pcl::PointXYZ> Ndt;
static_assert(std::is_base_of<pcl::Registration<pcl::PointXYZ, pcl::PointXYZ>, Ndt>::value, "derives from pcl::Registration (ndt_omp.h:70-71)");
Ndt make_by_value() {               // ndt_omp_mapping_node.cpp:151-169 returns the object by value
  Ndt ndt;
  ndt.setResolution(1.0f); ndt.setStepSize(0.1); ndt.setTransformationEpsilon(0.01); ndt.setMaximumIterations(64);
  ndt.setNumThreads(40); ndt.setNeighborhoodSearchMethod(pclomp_b200::DIRECT7); ndt.setOutlierRatio(0.55);
  return ndt;
}
int main() {
  Ndt::Ptr p(new Ndt());
  pcl::Registration<pcl::PointXYZ, pcl::PointXYZ>::Ptr base = p;           // apps/align.cpp:15, 103
  Ndt::ConstPtr cp = p;
  Cloud::Ptr a(new Cloud()), b(new Cloud());
  base->setInputTarget(a); base->setInputSource(b);
  Cloud out;
  base->align(out); base->align(out, Eigen::Matrix4f::Identity());
  (void)base->hasConverged(); (void)base->getFinalTransformation(); (void)base->getFitnessScore();
  base->setMaximumIterations(3); base->setTransformationEpsilon(0.5);
  Ndt v = make_by_value(); Ndt w(v); w = v;
  (void)v.getResolution(); (void)v.getStepSize(); (void)v.getOutlierRatio(); (void)v.getTransformationProbability();
  (void)v.getFinalNumIteration(); (void)v.calculateScore(out); (void)v.getFitnessScore(1.0); v.search_method = pclomp_b200::KDTREE;
  Eigen::Matrix<double, 6, 1> x; for (int i = 0; i < 6; ++i) x(i) = 0.1 * i;
  Eigen::Affine3f A; Eigen::Matrix4f M;
  Ndt::convertTransform(x, A); Ndt::convertTransform(x, M);                 // ndt_omp.h:216-233
  std::vector<const Cloud*> cands{&out, &out};
  std::vector<Eigen::Matrix4f> poses(2, M);
  (void)v.calculateScoreBatch(cands); (void)v.scorePoses(poses);            // batched calculateScore (loop-closure screening)
  pclomp_b200::VoxelGrid<pcl::PointXYZ> vg; vg.setLeafSize(0.1f, 0.2f, 0.3f); vg.setInputCloud(a); vg.filter(out);
  static_assert(pclomp_b200::KDTREE == 0 && pclomp_b200::DIRECT26 == 1 && pclomp_b200::DIRECT7 == 2 && pclomp_b200::DIRECT1 == 3, "ndt_omp.h:52-57");
  return M(3, 3) == 1.0f ? 0 : 1;
}
'''


@pytest.mark.parametrize("mode", ["compat", "minipcl"])
def test_public_surface_compiles_in_both_header_modes(tmp_path, mode):
    """Without PCL (pcl_compat.hpp stand-ins) and against a PCL-shaped header tree (the `__has_include(<pcl/...>)` branches
    a real PCL build takes)."""
    src = tmp_path / "surface.cpp"
    src.write_text("#include <type_traits>\n" + SURFACE_TU)
    extra = ["-I", MINIPCL] if mode == "minipcl" else []
    compile_cpp(str(src), str(tmp_path / "surface"), extra)
    compile_cpp(os.path.join(ROOT, "apps", "align_b200.cpp"), str(tmp_path / "app"), extra)


@pytest.mark.skipif(not os.path.exists("/root/reference/ndt_omp/apps/align.cpp"), reason="reference sources only exist in the build container")
def test_reference_app_ndt_section_compiles_unmodified():
    """ndt_omp/apps/align.cpp's NDT section (:15-33, 36-71, 88-105), read from /root/reference, with only the include
    line and the namespace changed (oracle/ref_app.py documents every dropped out-of-scope line)."""
    from oracle import ref_app
    out = ref_app.build(force=True)
    assert out and os.path.exists(out)
    with open(ref_app.OUT_SRC) as f, open(ref_app.REF_APP) as g:
        gen, ref = f.read().split("\n"), g.read().split("\n")
    kept = [l for i, l in enumerate(ref, start=1) if i not in ref_app.DROP_LINES]
    inserted = "  pcl::PointCloud<pcl::PointXYZ>::Ptr aligned;"           # declared by the removed GICP block (:76)
    assert gen.count(inserted) == 1
    gen.remove(inserted)
    assert len(gen) == len(kept)
    n_changed = 0
    for a, b in zip(gen, kept):                                        # the only edits: the include line + the namespace
        if a != b:
            n_changed += 1
            assert a == b.replace("pclomp::", "pclomp_b200::").replace("<pclomp/ndt_omp.h>", "<pclomp_b200/ndt_b200.hpp>"), (a, b)
    assert 1 <= n_changed <= 8


def write_pcd(path, xyz):   # the layout of ndt_omp/data/*.pcd: x y z intensity, float32, DATA binary
    rec = np.zeros((len(xyz), 4), dtype=np.float32)
    rec[:, :3] = xyz
    with open(path, "wb") as f:
        f.write(("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\n"
                 "COUNT 1 1 1 1\nWIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n" % (len(xyz), len(xyz))).encode())
        f.write(rec.tobytes())


@pytest.mark.gpu
def test_reference_app_ndt_section_reproduces_readme(tmp_path):
    """The reference's own app code (compiled by oracle/ref_app.py where /root/reference exists; the binary travels) on
    the bundled raw scans: `rosrun ndt_omp align 251370668.pcd 251371071.pcd` prints the README's fitness values
    (ndt_omp/README.md:21, 26, 31) for KDTREE / DIRECT7 / DIRECT1, for both thread settings — through a
    pcl::Registration pointer, i.e. with the base class's own (CPU) getFitnessScore on the device-computed transform."""
    from oracle import ref_app
    app = ref_app.build()
    if not app or not os.path.exists(app):
        pytest.skip("oracle/_ref/align_ndt_section was not built (needs /root/reference at build time)")
    d = np.load(os.path.join(GOLDEN, "pair_raw.npz"))
    tp, sp = str(tmp_path / "t.pcd"), str(tmp_path / "s.pcd")
    write_pcd(tp, d["target"]); write_pcd(sp, d["source"])
    out = subprocess.run([app, tp, sp], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr
    fit = [float(x) for x in re.findall(r"fitness: ([0-9.]+)", out.stdout)]
    g = golden()["fitness"]
    assert len(fit) == 6, out.stdout
    for got, name in zip(fit, ["KDTREE", "DIRECT7", "DIRECT1"] * 2):
        assert abs(got - g[name]) < 1e-6, (name, got, out.stdout)
    assert len(re.findall(r"single : [0-9.e+-]+\[msec\]", out.stdout)) == 6 and len(re.findall(r"10times: ", out.stdout)) == 6


@pytest.mark.gpu
def test_shim_app_against_minipcl_headers(tmp_path):
    """apps/align_b200.cpp built against the PCL-shaped header tree (real-PCL branches of the shim) gives the same
    answers as the stand-in build: README goldens, convertTransform round trip, copies, batch, mapper."""
    from util import load_pair
    app = compile_cpp(os.path.join(ROOT, "apps", "align_b200.cpp"), str(tmp_path / "app"), ["-I", MINIPCL])
    tgt, src = load_pair()
    tp, sp = str(tmp_path / "t.bin"), str(tmp_path / "s.bin")
    np.ascontiguousarray(tgt, dtype=np.float32).tofile(tp)
    np.ascontiguousarray(src, dtype=np.float32).tofile(sp)
    out = subprocess.run([app, tp, sp], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    fit = re.findall(r"fitness: ([0-9.]+)", out.stdout)
    assert abs(float(fit[0]) - golden()["fitness"]["DIRECT7"]) < 1e-6 and abs(float(fit[1]) - golden()["fitness"]["DIRECT1"]) < 1e-6
    assert "convertTransform(final pose) == getFinalTransformation(): yes" in out.stdout
    assert "copy converged: 1, iterations 5" in out.stdout
