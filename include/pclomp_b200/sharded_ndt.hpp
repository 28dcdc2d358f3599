// sharded_ndt.hpp — C++ host of the multi-GPU paths (BASELINE north_star: "host code stays C++"): one process
// per GPU of an NVSwitch node, NCCL for everything that is an exchange step on the host side, the C ABI of
// libndt_b200.so for everything on the device.
//
//   source-sharded align   every rank holds the full target map and a contiguous 32-point-aligned range of the source;
//                          the 29 per-evaluation sums are exchanged INSIDE the persistent kernel (P2P mailbox stores
//                          over NVLink, ndtb200_comm_*); NCCL only carries the 64-byte IPC handles once per source and
//                          a 4-byte all-reduce that lines the ranks' launches up
//   sharded getFitnessScore  ndtb200_fitness_sums per rank + a 2-double all-reduce
//   sharded target-map build (owner-partitioned): min/max all-reduce of the bounding box -> per-rank sort + per-voxel
//                          partials with the common grid -> all-to-all of the 88-byte partials to the rank that owns
//                          their key range (grouped ncclSend / ncclRecv) -> owner merges in rank order + finalises ->
//                          all-gather of the finished records -> every rank builds its voxel index
//
// Header-only; needs <nccl.h> and <cuda_runtime.h>; link -lndt_b200 -lnccl -lcudart.  The Python twin is
// toyslam_b200/sharding.py (torch.distributed for the same plumbing).
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../ndt_b200.h"

namespace pclomp_b200 {

class ShardedNdt {
 public:
  // comm: an initialised NCCL communicator of `world` ranks (one per GPU); device: this rank's CUDA device
  ShardedNdt(ncclComm_t comm, int rank, int world, int device) : comm_(comm), rank_(rank), world_(world), device_(device) {
    cudaSetDevice(device_);
    cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking);
    if (ndtb200_create(&h_, device_) != NDTB200_OK) h_ = nullptr;
    ndtb200_default_params(&prm_);
  }
  ~ShardedNdt() {
    cudaSetDevice(device_);
    if (h_) { ndtb200_comm_detach(h_); ndtb200_destroy(h_); }
    for (void* p : scratch_) cudaFree(p);
    if (stream_) cudaStreamDestroy(stream_);
  }
  ShardedNdt(const ShardedNdt&) = delete;
  ShardedNdt& operator=(const ShardedNdt&) = delete;

  bool ok() const { return h_ != nullptr; }
  ndtb200_handle* handle() { return h_; }
  ndtb200_params& params() { return prm_; }
  void pushParams() { if (h_) ndtb200_set_params(h_, &prm_); }
  const std::string& lastError() const { return err_; }

  // [lo, hi) of `rank`'s contiguous slice of an n-point source: slices start on 32-point group boundaries
  static void sourceRange(size_t n, int rank, int world, size_t& lo, size_t& hi) {
    const size_t groups = (n + 31) / 32, base = groups / world, extra = groups % world;
    const size_t g_lo = rank * base + std::min<size_t>(rank, extra);
    const size_t g_hi = g_lo + base + (static_cast<size_t>(rank) < extra ? 1 : 0);
    lo = std::min(n, g_lo * 32);
    hi = std::min(n, g_hi * 32);
  }
  static void pointRange(size_t n, int rank, int world, size_t& lo, size_t& hi) {
    const size_t base = n / world, extra = n % world;
    lo = rank * base + std::min<size_t>(rank, extra);
    hi = lo + base + (static_cast<size_t>(rank) < extra ? 1 : 0);
  }

  // replicated map: every rank builds the full map from the full cloud (host points, stride bytes)
  int setInputTarget(const void* points, size_t n, size_t stride, bool dense = true) {
    pushParams();
    return ndtb200_set_target(h_, points, n, stride, dense ? 1 : 0);
  }

  // full source cloud on every rank (host points): this rank uploads its range, the ranks exchange their mailbox handles
  int setInputSource(const void* points, size_t n, size_t stride) {
    size_t lo, hi;
    sourceRange(n, rank_, world_, lo, hi);
    int st = ndtb200_set_source(h_, static_cast<const char*>(points) + lo * stride, hi - lo, stride);
    if (st != NDTB200_OK) return fail(st, "set_source");
    if (world_ > 1) {
      char mine[NDTB200_COMM_HANDLE_BYTES];
      st = ndtb200_comm_export(h_, mine);
      if (st != NDTB200_OK) return fail(st, "comm_export");
      char* d_all = static_cast<char*>(scratch(0, (size_t)world_ * NDTB200_COMM_HANDLE_BYTES));
      cudaMemcpyAsync(d_all + (size_t)rank_ * NDTB200_COMM_HANDLE_BYTES, mine, NDTB200_COMM_HANDLE_BYTES, cudaMemcpyHostToDevice, stream_);
      if (!nccl(ncclAllGather(d_all + (size_t)rank_ * NDTB200_COMM_HANDLE_BYTES, d_all, NDTB200_COMM_HANDLE_BYTES, ncclChar, comm_, stream_))) return NDTB200_ERR_CUDA;
      std::vector<char> all((size_t)world_ * NDTB200_COMM_HANDLE_BYTES);
      cudaMemcpyAsync(all.data(), d_all, all.size(), cudaMemcpyDeviceToHost, stream_);
      cudaStreamSynchronize(stream_);
      st = ndtb200_comm_attach(h_, rank_, world_, all.data(), static_cast<int64_t>(n));
      if (st != NDTB200_OK) return fail(st, "comm_attach");
      barrier();
    }
    n_source_total_ = n;
    return NDTB200_OK;
  }

  // all ranks call align together (like a collective); every rank ends with the identical result
  int align(const float* guess_colmajor16, ndtb200_result* out) {
    pushParams();
    barrier();  // the ranks' persistent kernels wait on one another: launch them together
    int st = ndtb200_align(h_, guess_colmajor16, nullptr, 0);
    if (st != NDTB200_OK) return fail(st, "align");
    return out ? ndtb200_get_result(h_, out) : NDTB200_OK;
  }

  // pcl::Registration::getFitnessScore of the whole sharded source (needs the raw target on every rank: replicated map)
  int getFitnessScore(double max_range, double* out) {
    double s = 0;
    int64_t c = 0;
    int st = ndtb200_fitness_sums(h_, max_range, &s, &c);
    if (st != NDTB200_OK) return fail(st, "fitness_sums");
    double v[2] = {s, static_cast<double>(c)};
    if (world_ > 1) {
      double* d = static_cast<double*>(scratch(1, 2 * sizeof(double)));
      cudaMemcpyAsync(d, v, sizeof(v), cudaMemcpyHostToDevice, stream_);
      if (!nccl(ncclAllReduce(d, d, 2, ncclDouble, ncclSum, comm_, stream_))) return NDTB200_ERR_CUDA;
      cudaMemcpyAsync(v, d, sizeof(v), cudaMemcpyDeviceToHost, stream_);
      cudaStreamSynchronize(stream_);
    }
    *out = v[1] > 0 ? v[0] / v[1] : std::numeric_limits<double>::max();
    return NDTB200_OK;
  }

  // Owner-partitioned sharded build: d_local = this rank's slice of the target cloud (float4 x,y,z,w on this device)
  int setInputTargetSharded(const void* d_local_xyzw, size_t n_local, bool dense = true) {
    pushParams();
    float mn[3], mx[3];
    int64_t nf = 0;
    int st = ndtb200_cloud_bounds(h_, d_local_xyzw, n_local, dense ? 1 : 0, mn, mx, &nf);
    if (st != NDTB200_OK) return fail(st, "cloud_bounds");
    float gmin[3] = {mn[0], mn[1], mn[2]}, gmax[3] = {mx[0], mx[1], mx[2]};
    long long nf_total = nf;
    if (world_ > 1) {
      float* d = static_cast<float*>(scratch(1, 64));
      float box[6] = {mn[0], mn[1], mn[2], -mx[0], -mx[1], -mx[2]};  // one min all-reduce: max = -min(-x)
      cudaMemcpyAsync(d, box, sizeof(box), cudaMemcpyHostToDevice, stream_);
      if (!nccl(ncclAllReduce(d, d, 6, ncclFloat, ncclMin, comm_, stream_))) return NDTB200_ERR_CUDA;
      cudaMemcpyAsync(box, d, sizeof(box), cudaMemcpyDeviceToHost, stream_);
      long long* dn = reinterpret_cast<long long*>(d + 8);
      cudaMemcpyAsync(dn, &nf_total, sizeof(long long), cudaMemcpyHostToDevice, stream_);
      if (!nccl(ncclAllReduce(dn, dn, 1, ncclInt64, ncclSum, comm_, stream_))) return NDTB200_ERR_CUDA;
      cudaMemcpyAsync(&nf_total, dn, sizeof(long long), cudaMemcpyDeviceToHost, stream_);
      cudaStreamSynchronize(stream_);
      for (int a = 0; a < 3; ++a) { gmin[a] = box[a]; gmax[a] = -box[3 + a]; }
    }
    int64_t nv = 0;
    st = ndtb200_build_partials(h_, gmin, gmax, &nv);
    if (st != NDTB200_OK && st != NDTB200_ERR_GRID_OVERFLOW) return fail(st, "build_partials");
    if (st != NDTB200_OK) nv = 0;
    const size_t cap_in = std::max<size_t>(1, static_cast<size_t>(nv));
    int32_t* keys = static_cast<int32_t*>(scratch(2, cap_in * 4));
    uint32_t* cnts = static_cast<uint32_t*>(scratch(3, cap_in * 4));
    double* moms = static_cast<double*>(scratch(4, cap_in * 72));
    if (nv > 0 && (st = ndtb200_copy_partials(h_, keys, cnts, moms)) != NDTB200_OK) return fail(st, "copy_partials");
    if (world_ == 1) return ndtb200_build_from_partials(h_, gmin, gmax, nf_total, keys, cnts, moms, static_cast<size_t>(nv));

    // owners = ascending key ranges balanced by voxel count: splitters from a sample of every rank's sorted keys
    const int S = 256;
    std::vector<int32_t> sample(S, std::numeric_limits<int32_t>::max());
    if (nv > 0) {
      std::vector<int32_t> hk(static_cast<size_t>(nv));  // bring-up simplicity: the sample is read through one host copy
      cudaMemcpy(hk.data(), keys, hk.size() * 4, cudaMemcpyDeviceToHost);
      for (int i = 0; i < S; ++i) sample[i] = hk[static_cast<size_t>((double)i * (nv - 1) / (S - 1))];
    }
    int32_t* d_samples = static_cast<int32_t*>(scratch(5, (size_t)world_ * S * 4));
    cudaMemcpyAsync(d_samples + (size_t)rank_ * S, sample.data(), S * 4, cudaMemcpyHostToDevice, stream_);
    if (!nccl(ncclAllGather(d_samples + (size_t)rank_ * S, d_samples, S, ncclInt32, comm_, stream_))) return NDTB200_ERR_CUDA;
    std::vector<int32_t> all((size_t)world_ * S);
    cudaMemcpyAsync(all.data(), d_samples, all.size() * 4, cudaMemcpyDeviceToHost, stream_);
    cudaStreamSynchronize(stream_);
    std::sort(all.begin(), all.end());
    std::vector<int32_t> upper(world_ - 1);
    for (int r = 0; r + 1 < world_; ++r) upper[r] = all[(size_t)(r + 1) * S];
    std::vector<int64_t> offs(world_ + 1);
    st = ndtb200_partials_split(h_, upper.data(), world_, offs.data());
    if (st != NDTB200_OK) return fail(st, "partials_split");
    // send counts -> receive counts
    std::vector<long long> send(world_), recv(world_);
    for (int r = 0; r < world_; ++r) send[r] = offs[r + 1] - offs[r];
    long long* d_cnt = static_cast<long long*>(scratch(6, (size_t)world_ * world_ * 8));
    cudaMemcpyAsync(d_cnt + (size_t)rank_ * world_, send.data(), world_ * 8, cudaMemcpyHostToDevice, stream_);
    if (!nccl(ncclAllGather(d_cnt + (size_t)rank_ * world_, d_cnt, world_, ncclInt64, comm_, stream_))) return NDTB200_ERR_CUDA;
    std::vector<long long> matrix((size_t)world_ * world_);
    cudaMemcpyAsync(matrix.data(), d_cnt, matrix.size() * 8, cudaMemcpyDeviceToHost, stream_);
    cudaStreamSynchronize(stream_);
    size_t n_in = 0;
    for (int r = 0; r < world_; ++r) { recv[r] = matrix[(size_t)r * world_ + rank_]; n_in += recv[r]; }
    const size_t cap_recv = std::max<size_t>(1, n_in);
    int32_t* r_keys = static_cast<int32_t*>(scratch(7, cap_recv * 4));
    uint32_t* r_cnts = static_cast<uint32_t*>(scratch(8, cap_recv * 4));
    double* r_moms = static_cast<double*>(scratch(9, cap_recv * 72));
    // the all-to-all: one grouped send / receive per peer and array (partials arrive in rank order)
    ncclGroupStart();
    size_t roff = 0;
    for (int r = 0; r < world_; ++r) {
      if (send[r]) {
        ncclSend(keys + offs[r], send[r], ncclInt32, r, comm_, stream_);
        ncclSend(cnts + offs[r], send[r], ncclUint32, r, comm_, stream_);
        ncclSend(moms + offs[r] * 9, send[r] * 9, ncclDouble, r, comm_, stream_);
      }
      if (recv[r]) {
        ncclRecv(r_keys + roff, recv[r], ncclInt32, r, comm_, stream_);
        ncclRecv(r_cnts + roff, recv[r], ncclUint32, r, comm_, stream_);
        ncclRecv(r_moms + roff * 9, recv[r] * 9, ncclDouble, r, comm_, stream_);
      }
      roff += recv[r];
    }
    if (!nccl(ncclGroupEnd())) return NDTB200_ERR_CUDA;
    cudaStreamSynchronize(stream_);
    int64_t n_own = 0;
    st = ndtb200_merge_partials(h_, gmin, gmax, nf_total, r_keys, r_cnts, r_moms, n_in, &n_own);
    if (st != NDTB200_OK && st != NDTB200_ERR_NO_INPUT) return fail(st, "merge_partials");
    if (st != NDTB200_OK) n_own = 0;
    // all-gather the finished records in rank (= key) order
    long long own = n_own;
    long long* d_own = static_cast<long long*>(scratch(6, (size_t)world_ * world_ * 8));
    cudaMemcpyAsync(d_own + rank_, &own, 8, cudaMemcpyHostToDevice, stream_);
    if (!nccl(ncclAllGather(d_own + rank_, d_own, 1, ncclInt64, comm_, stream_))) return NDTB200_ERR_CUDA;
    std::vector<long long> owns(world_);
    cudaMemcpyAsync(owns.data(), d_own, world_ * 8, cudaMemcpyDeviceToHost, stream_);
    cudaStreamSynchronize(stream_);
    size_t total = 0, my_off = 0;
    for (int r = 0; r < world_; ++r) { if (r < rank_) my_off += owns[r]; total += owns[r]; }
    const size_t cap_all = std::max<size_t>(1, total);
    char* g_rec = static_cast<char*>(scratch(10, cap_all * 64));
    double* g_ic = static_cast<double*>(scratch(11, cap_all * 48));
    if (n_own > 0 && (st = ndtb200_copy_records(h_, g_rec + my_off * 64, g_ic + my_off * 6)) != NDTB200_OK) return fail(st, "copy_records");
    // uneven all-gather = one broadcast per owner inside a group
    ncclGroupStart();
    size_t off = 0;
    for (int r = 0; r < world_; ++r) {
      if (owns[r]) {
        ncclBroadcast(g_rec + off * 64, g_rec + off * 64, owns[r] * 64, ncclChar, r, comm_, stream_);
        ncclBroadcast(g_ic + off * 6, g_ic + off * 6, owns[r] * 6, ncclDouble, r, comm_, stream_);
      }
      off += owns[r];
    }
    if (!nccl(ncclGroupEnd())) return NDTB200_ERR_CUDA;
    cudaStreamSynchronize(stream_);
    return ndtb200_set_map_from_records(h_, gmin, gmax, nf_total, g_rec, g_ic, total);
  }

  void barrier() {
    if (world_ <= 1) return;
    int* d = static_cast<int*>(scratch(1, 64));
    ncclAllReduce(d, d, 1, ncclInt32, ncclSum, comm_, stream_);
    cudaStreamSynchronize(stream_);
  }

 private:
  void* scratch(int slot, size_t bytes) {
    if (slot >= static_cast<int>(scratch_.size())) { scratch_.resize(slot + 1, nullptr); scratch_cap_.resize(slot + 1, 0); }
    if (scratch_cap_[slot] < bytes) {
      if (scratch_[slot]) cudaFree(scratch_[slot]);
      cudaMalloc(&scratch_[slot], bytes + bytes / 4 + 256);
      scratch_cap_[slot] = bytes + bytes / 4 + 256;
    }
    return scratch_[slot];
  }
  bool nccl(ncclResult_t r) {
    if (r == ncclSuccess) return true;
    err_ = std::string("NCCL: ") + ncclGetErrorString(r);
    return false;
  }
  int fail(int st, const char* what) {
    err_ = std::string(what) + ": " + (h_ ? ndtb200_last_error(h_) : "no handle");
    return st;
  }

  ncclComm_t comm_;
  int rank_, world_, device_;
  cudaStream_t stream_ = nullptr;
  ndtb200_handle* h_ = nullptr;
  ndtb200_params prm_;
  size_t n_source_total_ = 0;
  std::vector<void*> scratch_;
  std::vector<size_t> scratch_cap_;
  std::string err_;
};

}  // namespace pclomp_b200
