// apps/sharded_b200.cpp — the multi-GPU paths through the C++ host (pclomp_b200::ShardedNdt over NCCL), one process per GPU:
//
//   sharded_b200 <rank> <world> <nccl-id-file> target.bin source.bin
//
// (.bin = raw float32 x,y,z triples).  Rank 0 creates the NCCL unique id and writes it to <nccl-id-file>; the other
// ranks wait for the file.  Part A: replicated target map, source sharded by contiguous ranges, align + getFitnessScore.
// Part B: the target map built by all ranks together (owner-partitioned sharded build), then the same align.
// Rank 0 prints one RESULT line per part; tests/test_gpu_multi.py compares them with the oracle.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include <pclomp_b200/sharded_ndt.hpp>

static bool load_bin(const std::string& path, std::vector<float>& xyzw) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  const size_t n = raw.size() / 12;
  xyzw.assign(n * 4, 1.0f);
  for (size_t i = 0; i < n; ++i) std::memcpy(&xyzw[i * 4], raw.data() + i * 12, 12);
  return true;
}

static void print_result(const char* tag, const ndtb200_result& r, double fitness, const ndtb200_map_info& mi) {
  std::printf("RESULT %s converged %d iterations %d evaluations %d hessian_passes %d fitness %.9g voxels %lld valid %lld final", tag, r.converged,
              r.iterations, r.n_evaluations, r.n_hessian_passes, fitness, (long long)mi.n_voxels, (long long)mi.n_valid);
  for (int i = 0; i < 16; ++i) std::printf(" %.9g", r.final_transformation[i]);
  std::printf(" tp %.12g\n", r.trans_probability);
}

int main(int argc, char** argv) {
  if (argc < 6) { std::cerr << "usage: sharded_b200 <rank> <world> <nccl-id-file> target.bin source.bin" << std::endl; return 64; }
  const int rank = std::atoi(argv[1]), world = std::atoi(argv[2]);
  const std::string idfile = argv[3];
  std::vector<float> tgt, src;
  if (!load_bin(argv[4], tgt) || !load_bin(argv[5], src)) { std::cerr << "failed to load the clouds" << std::endl; return 1; }
  int ndev = 0;
  cudaGetDeviceCount(&ndev);
  if (ndev < world) { std::cerr << "needs " << world << " GPUs" << std::endl; return 77; }
  cudaSetDevice(rank);
  ncclUniqueId id;
  if (rank == 0) {
    ncclGetUniqueId(&id);
    std::ofstream f(idfile + ".tmp", std::ios::binary);
    f.write(reinterpret_cast<const char*>(&id), sizeof(id));
    f.close();
    std::rename((idfile + ".tmp").c_str(), idfile.c_str());
  } else {
    for (int i = 0; i < 6000; ++i) {
      std::ifstream f(idfile, std::ios::binary);
      if (f && f.read(reinterpret_cast<char*>(&id), sizeof(id))) break;
      std::this_thread::sleep_for(std::chrono::milliseconds(10));
      if (i == 5999) { std::cerr << "no NCCL id" << std::endl; return 1; }
    }
  }
  ncclComm_t comm;
  if (ncclCommInitRank(&comm, world, id, rank) != ncclSuccess) { std::cerr << "ncclCommInitRank failed" << std::endl; return 1; }
  int rc = 0;
  {
    pclomp_b200::ShardedNdt ndt(comm, rank, world, rank);
    if (!ndt.ok()) return 2;
    const size_t n_t = tgt.size() / 4, n_s = src.size() / 4;
    ndtb200_result r;
    ndtb200_map_info mi;
    double fit = 0;
    // ---- A: replicated map, sharded source ----
    int st = ndt.setInputTarget(tgt.data(), n_t, 16);
    if (st == NDTB200_OK) st = ndt.setInputSource(src.data(), n_s, 16);
    if (st == NDTB200_OK) st = ndt.align(nullptr, &r);
    if (st == NDTB200_OK) st = ndt.getFitnessScore(std::numeric_limits<double>::max(), &fit);
    if (st != NDTB200_OK) { std::cerr << "rank " << rank << " part A failed: " << ndt.lastError() << std::endl; rc = 3; }
    ndtb200_get_map_info(ndt.handle(), &mi);
    if (rank == 0 && rc == 0) print_result("A", r, fit, mi);
    // every rank must hold the identical result: compare through the checksum of the final transform
    // ---- B: the map built by all ranks together ----
    if (rc == 0) {
      size_t lo, hi;
      pclomp_b200::ShardedNdt::pointRange(n_t, rank, world, lo, hi);
      void* d_local = nullptr;
      cudaMalloc(&d_local, std::max<size_t>(1, hi - lo) * 16);
      cudaMemcpy(d_local, tgt.data() + lo * 4, (hi - lo) * 16, cudaMemcpyHostToDevice);
      auto t0 = std::chrono::steady_clock::now();
      st = ndt.setInputTargetSharded(d_local, hi - lo);
      cudaDeviceSynchronize();
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      if (st == NDTB200_OK) st = ndt.setInputSource(src.data(), n_s, 16);
      if (st == NDTB200_OK) st = ndt.align(nullptr, &r);
      if (st != NDTB200_OK) { std::cerr << "rank " << rank << " part B failed (" << st << "): " << ndt.lastError() << std::endl; rc = 4; }
      ndtb200_get_map_info(ndt.handle(), &mi);
      if (rank == 0 && rc == 0) { print_result("B", r, 0.0, mi); std::printf("sharded build %.3f ms\n", ms); }
      cudaFree(d_local);
    }
  }
  ncclCommDestroy(comm);
  return rc;
}
