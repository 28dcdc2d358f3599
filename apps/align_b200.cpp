// apps/align_b200.cpp — the NDT part of ndt_omp/apps/align.cpp (:15-33, 83-105) on the B200 library, through the
// header-only shim with the reference's method names.  Same measurement protocol: setInputTarget/Source once,
// time one align ("single"), then ten more ("10times"), then print getFitnessScore().
//
//   align_b200 target.{pcd|bin} source.{pcd|bin} [--leaf 0.1]
//
// .pcd: PCD v0.7 "DATA binary" files with float32 x y z first (the format of ndt_omp/data/*.pcd);
// .bin: raw float32 x,y,z triples.  The reference app first downsamples both clouds with a 0.1 m pcl::VoxelGrid
// (:57-69); pass clouds that are already downsampled (tests/golden/pair_ds0p1.npz exported as .bin), or see
// INTEGRATION.md for the PCL build where pcl::VoxelGrid is available.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include <pclomp_b200/ndt_b200.hpp>
#include <pclomp_b200/pcd_io.hpp>

typedef pcl::PointCloud<pcl::PointXYZ> Cloud;

static bool load_cloud(const std::string& path, Cloud& cloud) {
  if (path.size() > 4 && path.substr(path.size() - 4) == ".pcd") return pclomp_b200::io::loadPCDFile(path, cloud) == 0;
  std::ifstream f(path, std::ios::binary);  // .bin: raw float32 x,y,z triples
  if (!f) return false;
  std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  const size_t npts = raw.size() / 12;
  cloud.points.resize(npts);
  for (size_t i = 0; i < npts; ++i) {
    float xyz[3];
    std::memcpy(xyz, raw.data() + i * 12, 12);
    cloud.points[i] = pcl::PointXYZ(xyz[0], xyz[1], xyz[2]);
  }
  cloud.width = static_cast<uint32_t>(npts);
  cloud.height = 1;
  cloud.is_dense = true;
  return true;
}

// align point clouds and measure processing time — the helper of ndt_omp/apps/align.cpp:15-33 with its signature: it
// takes a pcl::Registration pointer, which the shim class is (it derives from pcl::Registration like the reference class)
static Cloud::Ptr align(pcl::Registration<pcl::PointXYZ, pcl::PointXYZ>::Ptr registration, const Cloud::Ptr& target_cloud,
                        const Cloud::Ptr& source_cloud) {
  registration->setInputTarget(target_cloud);
  registration->setInputSource(source_cloud);
  Cloud::Ptr aligned(new Cloud());
  auto t1 = std::chrono::steady_clock::now();
  registration->align(*aligned);
  auto t2 = std::chrono::steady_clock::now();
  std::cout << "single : " << std::chrono::duration<double, std::milli>(t2 - t1).count() << "[msec]" << std::endl;
  for (int i = 0; i < 10; i++) registration->align(*aligned);
  auto t3 = std::chrono::steady_clock::now();
  std::cout << "10times: " << std::chrono::duration<double, std::milli>(t3 - t2).count() << "[msec]" << std::endl;
  std::cout.precision(6);
  std::cout << "fitness: " << registration->getFitnessScore() << std::endl << std::endl;
  return aligned;
}

int main(int argc, char** argv) {
  if (argc < 3) {
    std::cout << "usage: align_b200 target.{pcd|bin} source.{pcd|bin}" << std::endl;
    return 0;
  }
  Cloud::Ptr target_cloud(new Cloud()), source_cloud(new Cloud());
  if (!load_cloud(argv[1], *target_cloud)) { std::cerr << "failed to load " << argv[1] << std::endl; return 1; }
  if (!load_cloud(argv[2], *source_cloud)) { std::cerr << "failed to load " << argv[2] << std::endl; return 1; }
  std::cout << "target " << target_cloud->size() << " pts, source " << source_cloud->size() << " pts" << std::endl;

  // downsampling (ndt_omp/apps/align.cpp:57-69): --leaf 0.1 reproduces the reference app on the raw PCD files
  float leaf = 0.f;
  for (int i = 3; i + 1 < argc; ++i)
    if (std::string(argv[i]) == "--leaf") leaf = static_cast<float>(std::atof(argv[i + 1]));
  if (leaf > 0.f) {
    pclomp_b200::VoxelGrid<pcl::PointXYZ> voxelgrid;
    voxelgrid.setLeafSize(leaf, leaf, leaf);
    Cloud::Ptr downsampled(new Cloud());
    voxelgrid.setInputCloud(target_cloud);
    voxelgrid.filter(*downsampled);
    *target_cloud = *downsampled;
    voxelgrid.setInputCloud(source_cloud);
    voxelgrid.filter(*downsampled);
    source_cloud = downsampled;
    std::cout << "downsampled (" << leaf << " m): target " << target_cloud->size() << " pts, source " << source_cloud->size() << " pts" << std::endl;
  }

  pclomp_b200::NormalDistributionsTransform<pcl::PointXYZ, pcl::PointXYZ>::Ptr ndt_omp(
      new pclomp_b200::NormalDistributionsTransform<pcl::PointXYZ, pcl::PointXYZ>());
  pclomp_b200::NormalDistributionsTransform<pcl::PointXYZ, pcl::PointXYZ>& ndt = *ndt_omp;
  if (!ndt.handle()) return 2;
  ndt_omp->setResolution(1.0);
  const std::pair<const char*, pclomp_b200::NeighborSearchMethod> methods[] = {
      {"DIRECT7", pclomp_b200::DIRECT7}, {"DIRECT1", pclomp_b200::DIRECT1}, {"DIRECT26", pclomp_b200::DIRECT26}};
  for (const auto& m : methods) {
    std::cout << "--- pclomp_b200::NDT (" << m.first << ") ---" << std::endl;
    ndt_omp->setNumThreads(1);
    ndt_omp->setNeighborhoodSearchMethod(m.second);
    align(ndt_omp, target_cloud, source_cloud);
  }
  // the mapping nodes copy the object (ndt_omp_mapping_node.cpp:151-169): a copy must give the same answer
  auto copy = ndt;
  copy.setNeighborhoodSearchMethod(pclomp_b200::DIRECT7);
  Cloud out;
  copy.align(out);
  {  // static convertTransform (ndt_omp.h:216-233): the final pose vector re-composes to the final transformation
    Eigen::Matrix<double, 6, 1> x;
    ndtb200_result r;
    ndtb200_get_result(copy.handle(), &r);
    for (int i = 0; i < 6; ++i) x(i) = r.final_pose[i];
    Eigen::Matrix4f M;
    pclomp_b200::NormalDistributionsTransform<pcl::PointXYZ, pcl::PointXYZ>::convertTransform(x, M);
    const Eigen::Matrix4f F = copy.getFinalTransformation();
    float d = 0.f;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) d = std::max(d, std::fabs(M(i, j) - F(i, j)));
    std::cout << "convertTransform(final pose) == getFinalTransformation(): " << (d == 0.f ? "yes" : "no") << std::endl;
  }
  std::cout << "copy converged: " << copy.hasConverged() << ", iterations " << copy.getFinalNumIteration() << std::endl;

  // independent pairs aligned together (not in the reference): four copies of the pair, one batch call
  typedef pclomp_b200::NormalDistributionsTransform<pcl::PointXYZ, pcl::PointXYZ> Ndt;
  std::vector<Ndt> objs(4, ndt);
  std::vector<Ndt*> ptrs;
  std::vector<Cloud> outs(objs.size());
  std::vector<Cloud*> out_ptrs;
  for (size_t i = 0; i < objs.size(); ++i) {
    objs[i].setNeighborhoodSearchMethod(pclomp_b200::DIRECT7);
    ptrs.push_back(&objs[i]);
    out_ptrs.push_back(&outs[i]);
  }
  auto b0 = std::chrono::steady_clock::now();
  Ndt::alignBatch(ptrs, out_ptrs);
  auto b1 = std::chrono::steady_clock::now();
  std::cout << "batch of " << objs.size() << ": " << std::chrono::duration<double, std::milli>(b1 - b0).count() << "[msec], iterations";
  for (auto& o : objs) std::cout << " " << o.getFinalNumIteration();
  std::cout << std::endl;

  // the mapping-node loop on the pipeline API: the two clouds as two consecutive scans
  {
    pclomp_b200::Mapper<pcl::PointXYZ> mapper(0.3f, 0.5f);
    if (mapper.ok()) {
      mapper.pushScan(*target_cloud);
      const auto step = mapper.pushScan(*source_cloud);
      Cloud gmap;
      mapper.globalMap(gmap);
      std::cout << "mapper: 2 scans, converged " << step.converged << ", iterations " << step.iterations << ", filtered " << step.n_filtered
                << " pts, map " << gmap.size() << " pts" << std::endl;
    }
  }

  // --save-aligned out.pcd: the aligned source of the DIRECT7 copy as a binary PCD (pcl::io::savePCDFileBinary)
  for (int i = 3; i + 1 < argc; ++i)
    if (std::string(argv[i]) == "--save-aligned") {
      if (pclomp_b200::io::savePCDFileBinary(argv[i + 1], out) != 0) { std::cerr << "failed to write " << argv[i + 1] << std::endl; return 1; }
      Cloud back;
      if (pclomp_b200::io::loadPCDFile(argv[i + 1], back) != 0 || back.size() != out.size()) { std::cerr << "PCD round trip failed" << std::endl; return 1; }
      std::cout << "saved " << back.size() << " aligned points to " << argv[i + 1] << std::endl;
    }
  return 0;
}
