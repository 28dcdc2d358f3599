// capi.cu — host side of libndt_b200.so: handle, device memory, kernel orchestration, C ABI.
// See include/ndt_b200.h for the contract of every entry point and the reference method it replaces.
#include <sched.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ndt_b200.h"
#include "map_build.cuh"
#include "ndt_align.cuh"
#include "ndt_aux.cuh"
#include "small_build.cuh"

using namespace ndtb200;

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    // first allocation: 12.5 % headroom; a buffer that has to grow again (the mapper's global map, a longer scan)
    // grows by 50 % so that cudaFree + cudaMalloc (an implicit device synchronisation) stays rare
    const size_t want = (cap == 0) ? bytes + bytes / 8 + 256 : bytes + bytes / 2 + 256;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T> T* as() const { return static_cast<T*>(p); }
};

}  // namespace

struct ndtb200_handle {
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  ndtb200_params prm{};
  long long launches = 0;
  uint32_t launch_seq = 0;
  int last_blocks = 0;
  bool result_copy_enqueued = false;
  bool want_fused_out = false;  // set by the entry points that know the caller wants align()'s output cloud
  bool out_fused = false;       // the last solve wrote d_out itself

  // multi-GPU source sharding
  int comm_world = 1, comm_rank = 0;
  long long comm_n_total = 0;
  DevBuf d_mail;                               // own mailbox: [2][kMaxRanks][kNVP][2] uint64
  unsigned long long* mail_ptrs[kMaxRanks] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool mail_opened[kMaxRanks] = {false, false, false, false, false, false, false, false};

  // target cloud + map
  DevBuf d_target;
  const float4* target_view = nullptr;  // ndtb200_set_target_device_view: the caller's own device buffer, not copied
  size_t n_target = 0;
  bool target_dense = true, has_target = false;
  int map_status = NDTB200_ERR_NO_INPUT;
  GridDesc grid{};
  long long n_voxels = 0, n_valid = 0;
  uint32_t hash_cap = 0;
  int hash_shift = 0;
  DevBuf d_grid, d_mm_partial, d_mm_finite, d_keys_a, d_keys_b, d_vals_a, d_vals_b, d_hist, d_scan_tmp, d_scalar, d_sortmeta, d_status;
  DevBuf d_voxel_key, d_voxel_start, d_voxel_count, d_moments, d_records, d_icov64, d_hash, d_dense;
  size_t n_partials = 0;       // voxels of the last partial-only build (sharded build, before the exchange)
  bool map_is_merged = false;
  bool records_only_map = false;  // map installed from finished records (sharded build): no moments, no raw target
  bool moments_valid = false;  // d_moments holds the per-voxel moments of the current map (the staged build fuses them away)
  bool prefer_fused_build = false;  // set by the mapping pipeline: fused builds even with the small-CTA solve shape
  // payload sort (clouds above kPayloadSortMin points): the sort moves the points themselves (key in .w); the sorted cloud
  // ends in d_pay_a.  idx_lazy: the sorted point INDICES (d_vals_a: getFitnessScore, KDTREE centroids, parity dumps) were
  // not produced by the build and are computed on first use (ensure_sorted_indices)
  DevBuf d_pay_a, d_pay_b;
  bool idx_lazy = false;
  DevBuf d_centroid;              // KDTREE mode: fp32 centroid of every voxel
  bool centroids_valid = false;
  DevBuf d_cell_all, d_best;      // getFitnessScore: cell table over all occupied voxels, per-query results
  bool cell_all_valid = false;
  float vg_leaf[3] = {0.f, 0.f, 0.f};  // ndtb200_voxelgrid_filter*: per-axis leaf of the scratch handle (0 = use prm.resolution)
  ndtb200_handle* aux = nullptr;  // scratch state of ndtb200_voxelgrid_filter (keeps the map's build buffers untouched)  // map built from all ranks' partials: d_target holds only this rank's slice
  bool use_dense = false;

  // source cloud
  DevBuf d_source;
  size_t n_source = 0;
  bool has_source = false;
  // large sources: a copy ordered by voxel key (cells of the source's own grid at the map resolution), so that the 32
  // points of a warp fall into one or two voxels — coalesced probes, shared Gaussian records (SURVEY 7.1 last bullet)
  DevBuf d_source_sorted;
  bool source_sorted_valid = false, source_sort_failed = false;

  // align workspace
  DevBuf d_partials, d_totals, d_sync, d_result, d_trace, d_out, d_tmp, d_emu;
  int emu_world = 0;
  size_t emu_per_rank = 0, emu_result_off = 0;
  int coop_blocks[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};  // grid size per search method and CTA shape
  int shape = 0;  // 0 = latency shape (1024 threads, the whole GPU for one solve), 1 = throughput shape (256 threads, 1 CTA / SM)
  AlignResultDev* h_result = nullptr;  // pinned
  bool result_valid = false;
  float last_final_T[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  static constexpr int kTraceCap = 1024;
};

namespace {

inline const float4* target_pts(const ndtb200_handle* h) { return h->target_view ? h->target_view : h->d_target.as<float4>(); }

#define CK2(hh, call)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      (hh)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                             \
      return NDTB200_ERR_CUDA;                                                                     \
    }                                                                                              \
  } while (0)

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e__);                                \
      return NDTB200_ERR_CUDA;                                                                     \
    }                                                                                              \
  } while (0)

#define LAUNCHED(h) ((h)->launches++)

inline int grid_for(size_t n, int per_block, int cap) {
  size_t b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > static_cast<size_t>(cap)) b = cap;
  return static_cast<int>(b);
}

// ---- host copies of a strided host cloud into a packed float4 device buffer ------------------
int upload_points(ndtb200_handle* h, DevBuf& dst, const void* points, size_t n, size_t stride) {
  if (n == 0) return NDTB200_OK;
  if (stride < 12) { h->err = "stride_bytes must be >= 12"; return NDTB200_ERR_INVALID; }
  CK(dst.ensure(n * sizeof(float4)));
  if (stride == 16) {
    CK(cudaMemcpyAsync(dst.p, points, n * 16, cudaMemcpyHostToDevice, h->stream));
  } else {
    size_t width = stride < 16 ? stride : 16;
    if (width < 16) CK(cudaMemsetAsync(dst.p, 0, n * 16, h->stream));
    CK(cudaMemcpy2DAsync(dst.p, 16, points, stride, width, n, cudaMemcpyHostToDevice, h->stream));
  }
  return NDTB200_OK;
}

// ---- exclusive scan (recursive 3-kernel) ------------------------------------------------------
int exclusive_scan(ndtb200_handle* h, uint32_t* d_data, size_t n, uint32_t* d_tmp, uint32_t* d_total) {
  // in-place exclusive scan of d_data[0..n); d_tmp must hold sum over levels of ceil(n / tile^k)
  const int ntiles = static_cast<int>((n + kScanTile - 1) / kScanTile);
  scan_tiles_kernel<<<ntiles, kBuildThreads, 0, h->stream>>>(d_data, d_data, n, d_tmp);
  LAUNCHED(h);
  if (ntiles == 1) {
    if (d_total) CK(cudaMemcpyAsync(d_total, d_tmp, sizeof(uint32_t), cudaMemcpyDeviceToDevice, h->stream));
    return NDTB200_OK;
  }
  int st = exclusive_scan(h, d_tmp, ntiles, d_tmp + ((ntiles + 63) & ~63), d_total);
  if (st != NDTB200_OK) return st;
  scan_add_kernel<<<ntiles, kBuildThreads, 0, h->stream>>>(d_data, n, d_tmp);
  LAUNCHED(h);
  return NDTB200_OK;
}

size_t scan_tmp_elems(size_t n) {
  size_t total = 0;
  while (true) {
    size_t t = (n + kScanTile - 1) / kScanTile;
    total += (t + 63) & ~size_t(63);
    if (t <= 1) break;
    n = t;
  }
  return total + 64;
}

void clear_map(ndtb200_handle* h) {
  h->records_only_map = false;
  h->idx_lazy = false;
  h->cell_all_valid = false;
  h->centroids_valid = false;
  h->n_voxels = 0;
  h->n_valid = 0;
}

int ensure_empty_hash(ndtb200_handle* h) {
  h->hash_cap = 1024;
  h->hash_shift = 32 - 10;
  CK(h->d_hash.ensure(h->hash_cap * sizeof(HashSlot)));
  CK(cudaMemsetAsync(h->d_hash.p, 0xFF, h->hash_cap * sizeof(HashSlot), h->stream));
  CK(h->d_records.ensure(sizeof(VoxelRecord)));
  CK(h->d_icov64.ensure(6 * sizeof(double)));
  h->use_dense = false;
  return NDTB200_OK;
}

// ---- VoxelGridCovariance::applyFilter on the device -------------------------------------------
struct BuildOpts {
  // sharded build: the grid comes from the bounding box of the WHOLE cloud (all ranks), the handle
  // processes only its slice and stops after the per-voxel moments ("partials")
  const float* forced_min = nullptr;
  const float* forced_max = nullptr;
  bool partial_only = false;
};

// bounding box of the handle's target slice + grid description -> h->grid (host copy); one synchronisation
Leaf3 leaf_of(const ndtb200_handle* h) {
  Leaf3 l;
  for (int a = 0; a < 3; ++a) l.v[a] = h->vg_leaf[a] > 0.f ? h->vg_leaf[a] : h->prm.resolution;
  return l;
}

int compute_grid(ndtb200_handle* h, const float4* pts, size_t n, int dense, const BuildOpts& o) {
  const int mm_blocks = grid_for(n, kBuildThreads * 4, h->num_sms * 8);
  CK(h->d_mm_partial.ensure((size_t)mm_blocks * 6 * sizeof(float)));
  CK(h->d_mm_finite.ensure((size_t)mm_blocks * sizeof(unsigned int)));
  CK(h->d_grid.ensure(sizeof(GridDesc)));
  minmax3d_kernel<<<mm_blocks, kBuildThreads, 0, h->stream>>>(pts, n, dense, h->d_mm_partial.as<float>(),
                                                               h->d_mm_finite.as<unsigned int>());
  LAUNCHED(h);
  ForcedBox fb;
  fb.use = (o.forced_min && o.forced_max) ? 1 : 0;
  for (int a = 0; a < 3; ++a) { fb.mn[a] = fb.use ? o.forced_min[a] : 0.f; fb.mx[a] = fb.use ? o.forced_max[a] : 0.f; }
  grid_setup_kernel<<<1, kBuildThreads, 0, h->stream>>>(h->d_mm_partial.as<float>(), h->d_mm_finite.as<unsigned int>(),
                                                        mm_blocks, leaf_of(h), fb, h->d_grid.as<GridDesc>());
  LAUNCHED(h);
  CK(cudaMemcpyAsync(&h->grid, h->d_grid.p, sizeof(GridDesc), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return NDTB200_OK;
}

// One-sweep sort bookkeeping (map_build.cuh): [digit_hist 4 x 256 u64 | digit_base 4 x 256 u64 | tickets 4 x u32 (+ pad)]
constexpr size_t kSortMetaHistOff = 0, kSortMetaBaseOff = kMaxSortPasses * 256 * sizeof(unsigned long long),
                 kSortMetaTicketOff = 2 * kSortMetaBaseOff, kSortMetaBytes = kSortMetaTicketOff + 64;

// zero the histograms / tickets before the keys are produced (the key kernel accumulates the digit histograms)
int sort_meta_reset(ndtb200_handle* h) {
  CK(h->d_sortmeta.ensure(kSortMetaBytes));
  CK(cudaMemsetAsync(h->d_sortmeta.p, 0, kSortMetaBytes, h->stream));
  return NDTB200_OK;
}
unsigned long long* sort_meta_hist(ndtb200_handle* h) { return reinterpret_cast<unsigned long long*>(h->d_sortmeta.as<char>() + kSortMetaHistOff); }

// stable LSD radix sort (one-sweep passes) of the n (key, index) pairs whose keys sit in d_keys_a:
// sorted keys -> d_keys_a, sorted indices -> d_vals_a.  hist_ready: the digit histograms were accumulated by the key
// kernel (sort_meta_reset was called before it); otherwise they are computed from the keys here.  No synchronisation.
int sort_pairs(ndtb200_handle* h, size_t n, int passes, bool hist_ready) {
  const int ntiles = static_cast<int>((n + kSortTile - 1) / kSortTile);
  unsigned long long* hist = sort_meta_hist(h);
  unsigned long long* base = reinterpret_cast<unsigned long long*>(h->d_sortmeta.as<char>() + kSortMetaBaseOff);
  unsigned int* tickets = reinterpret_cast<unsigned int*>(h->d_sortmeta.as<char>() + kSortMetaTicketOff);
  uint32_t *ka = h->d_keys_a.as<uint32_t>(), *kb = h->d_keys_b.as<uint32_t>();
  uint32_t *va = h->d_vals_a.as<uint32_t>(), *vb = h->d_vals_b.as<uint32_t>();
  if (!hist_ready) {
    int st = sort_meta_reset(h);
    if (st != NDTB200_OK) return st;
    hist = sort_meta_hist(h);
    base = reinterpret_cast<unsigned long long*>(h->d_sortmeta.as<char>() + kSortMetaBaseOff);
    tickets = reinterpret_cast<unsigned int*>(h->d_sortmeta.as<char>() + kSortMetaTicketOff);
    digit_hist_kernel<<<grid_for(n, kBuildThreads * 8, h->num_sms * 8), kBuildThreads, 0, h->stream>>>(ka, n, hist, passes);
    LAUNCHED(h);
  }
  digit_base_kernel<<<1, 256, 0, h->stream>>>(hist, passes, base);
  LAUNCHED(h);
  const size_t status_bytes = (size_t)passes * ntiles * 256 * sizeof(unsigned long long);
  CK(h->d_status.ensure(status_bytes));
  CK(cudaMemsetAsync(h->d_status.p, 0, status_bytes, h->stream));
  for (int pass = 0; pass < passes; ++pass) {
    onesweep_kernel<<<ntiles, kBuildThreads, 0, h->stream>>>(ka, pass == 0 ? nullptr : va, n, pass * 8, base + pass * 256,
                                                              h->d_status.as<unsigned long long>() + (size_t)pass * ntiles * 256,
                                                              tickets + pass, kb, vb);
    LAUNCHED(h);
    std::swap(ka, kb);
    std::swap(va, vb);
  }
  // sorted keys in ka, sorted indices in va.  Keep them addressable through fixed members.
  if (ka != h->d_keys_a.as<uint32_t>()) { std::swap(h->d_keys_a, h->d_keys_b); std::swap(h->d_vals_a, h->d_vals_b); }
  return NDTB200_OK;
}

// sort + the segment heads: -> d_voxel_key / d_voxel_start (n_vox entries), sorted indices in d_vals_a.  One
// synchronisation (the voxel count, needed to size the voxel arrays).
int segment_heads(ndtb200_handle* h, size_t n, uint32_t sentinel, uint32_t* n_vox_out);

int sort_and_segment(ndtb200_handle* h, size_t n, uint32_t sentinel, int passes, uint32_t* n_vox_out, bool hist_ready = false) {
  {
    int st = sort_pairs(h, n, passes, hist_ready);
    if (st != NDTB200_OK) return st;
  }
  return segment_heads(h, n, sentinel, n_vox_out);
}

// occupied voxels = segment heads of the sorted keys in d_keys_a -> d_voxel_key / d_voxel_start; one synchronisation
int segment_heads(ndtb200_handle* h, size_t n, uint32_t sentinel, uint32_t* n_vox_out) {
  const int stiles = static_cast<int>((n + kScanTile - 1) / kScanTile);
  CK(h->d_hist.ensure((size_t)stiles * sizeof(uint32_t)));
  CK(h->d_scan_tmp.ensure(scan_tmp_elems(std::max<size_t>(stiles, 1)) * sizeof(uint32_t)));
  uint32_t* ka = h->d_keys_a.as<uint32_t>();
  uint32_t* tile_counts = h->d_hist.as<uint32_t>();
  head_count_kernel<<<stiles, kBuildThreads, 0, h->stream>>>(ka, n, sentinel, tile_counts);
  LAUNCHED(h);
  uint32_t* d_total = h->d_scalar.as<uint32_t>();
  {
    int st = exclusive_scan(h, tile_counts, stiles, h->d_scan_tmp.as<uint32_t>(), d_total);
    if (st != NDTB200_OK) return st;
  }
  uint32_t n_vox = 0;
  CK(cudaMemcpyAsync(&n_vox, d_total, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(h->d_voxel_key.ensure((size_t)std::max<uint32_t>(n_vox, 1u) * sizeof(int32_t)));
  CK(h->d_voxel_start.ensure((size_t)std::max<uint32_t>(n_vox, 1u) * sizeof(uint32_t)));
  head_write_kernel<<<stiles, kBuildThreads, 0, h->stream>>>(ka, n, sentinel, tile_counts, h->d_voxel_key.as<int32_t>(),
                                                              h->d_voxel_start.as<uint32_t>());
  LAUNCHED(h);
  *n_vox_out = n_vox;
  return NDTB200_OK;
}

// ---- payload sort (map_build.cuh: onesweep_payload_kernel) ------------------------------------------------------
// Clouds of kPayloadSortMin points and more: the (key, index) sort is bound by the gather that follows it (64 B of DRAM
// per point); here the passes carry the points.  Results are bit-identical to the (key, index) path (same stable order,
// same additions): tests/test_gpu_full_size.py compares the two.  NDTB200_PAYLOAD_SORT_MIN overrides the threshold.
constexpr size_t kPayloadSortMin = 4u << 20;
bool use_payload_sort(const ndtb200_handle* h, size_t n) {
  (void)h;
  size_t lo = kPayloadSortMin;
  if (const char* e = getenv("NDTB200_PAYLOAD_SORT_MIN")) lo = static_cast<size_t>(strtoull(e, nullptr, 10));
  return n >= lo && n < (1ull << 30);
}

// stable LSD radix sort of the target cloud by voxel key, points as payload: sorted cloud (key in .w) -> d_pay_a, sorted
// keys -> d_keys_a.  The digit histograms must already be in the sort meta block (voxel_key_kernel, keys == nullptr).
int sort_payload(ndtb200_handle* h, const float4* pts, size_t n, int dense, uint32_t sentinel, int passes) {
  const int ntiles = static_cast<int>((n + kPayTile - 1) / kPayTile);
  unsigned long long* hist = sort_meta_hist(h);
  unsigned long long* base = reinterpret_cast<unsigned long long*>(h->d_sortmeta.as<char>() + kSortMetaBaseOff);
  unsigned int* tickets = reinterpret_cast<unsigned int*>(h->d_sortmeta.as<char>() + kSortMetaTicketOff);
  CK(h->d_pay_a.ensure(n * sizeof(float4)));
  if (passes > 1) CK(h->d_pay_b.ensure(n * sizeof(float4)));
  CK(h->d_keys_a.ensure(n * sizeof(uint32_t)));
  digit_base_kernel<<<1, 256, 0, h->stream>>>(hist, passes, base);
  LAUNCHED(h);
  const size_t status_bytes = (size_t)passes * ntiles * 256 * sizeof(uint32_t);
  CK(h->d_status.ensure(status_bytes));
  CK(cudaMemsetAsync(h->d_status.p, 0, status_bytes, h->stream));
  // the last pass must land in d_pay_a: pass p writes a when (passes - 1 - p) is even
  const float4* in = pts;
  for (int pass = 0; pass < passes; ++pass) {
    float4* out = ((passes - 1 - pass) % 2 == 0) ? h->d_pay_a.as<float4>() : h->d_pay_b.as<float4>();
    uint32_t* st = h->d_status.as<uint32_t>() + (size_t)pass * ntiles * 256;
    uint32_t* keys_out = (pass == passes - 1) ? h->d_keys_a.as<uint32_t>() : nullptr;
    if (pass == 0)
      onesweep_payload_kernel<true><<<ntiles, kBuildThreads, 0, h->stream>>>(in, static_cast<uint32_t>(n), dense, h->d_grid.as<GridDesc>(), sentinel,
                                                                            0, base, st, tickets, out, keys_out);
    else
      onesweep_payload_kernel<false><<<ntiles, kBuildThreads, 0, h->stream>>>(in, static_cast<uint32_t>(n), dense, h->d_grid.as<GridDesc>(), sentinel,
                                                                             pass * 8, base + pass * 256, st, tickets + pass, out, keys_out);
    LAUNCHED(h);
    in = out;
  }
  return NDTB200_OK;
}

int passes_for(const GridDesc& g, bool with_sentinel, uint32_t* sentinel_out) {
  unsigned long long key_space = (unsigned long long)g.div_b[0] * (unsigned long long)g.div_b[1] * (unsigned long long)g.div_b[2];
  if (key_space > 0xFFFFFFFEull) key_space = 0xFFFFFFFEull;
  *sentinel_out = static_cast<uint32_t>(key_space);  // sorts after every real key
  const unsigned long long max_key = with_sentinel ? key_space : (key_space - 1);
  int bits = 1;
  while (bits < 32 && (max_key >> bits) != 0) ++bits;
  return (bits + 7) / 8;
}

// Which voxel index the map gets: the direct-mapped cell table (one 4-byte load per probe, no collisions) when dx*dy*dz
// int32 entries fit the budget (<= 4 GiB and <= 1/4 of the free device memory) — allocated and filled with -1 here —
// otherwise the open-addressing hash over the valid voxels (finish_hash_index).
int prepare_index(ndtb200_handle* h) {
  h->use_dense = false;
  h->n_valid = -1;  // fetched on demand (ndtb200_get_map_info) unless the hash needs it
  // cells actually addressed by keys: div_b product (the guard's dx*dy*dz, grid.ncell, is computed from the float
  // extents and can be smaller by one per axis)
  const unsigned long long ncell = static_cast<unsigned long long>(h->grid.div_b[0]) * static_cast<unsigned long long>(h->grid.div_b[1]) *
                                   static_cast<unsigned long long>(h->grid.div_b[2]);
  const unsigned long long bytes = ncell * sizeof(int32_t);
  const bool forced_hash = getenv("NDTB200_FORCE_HASH") != nullptr;
  bool fits = !forced_hash && ncell > 0 && bytes <= (4ull << 30);
  if (fits && bytes > h->d_dense.cap && bytes > (64ull << 20)) {  // cudaMemGetInfo is slow: only ask for large NEW tables
    size_t free_b = 0, total_b = 0;
    fits = cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && bytes <= (free_b + h->d_dense.cap) / 4;
  }
  if (fits) {
    CK(h->d_dense.ensure(bytes));
    CK(cudaMemsetAsync(h->d_dense.p, 0xFF, bytes, h->stream));
    h->use_dense = true;
  }
  return NDTB200_OK;
}

int finish_hash_index(ndtb200_handle* h, uint32_t n_vox);

// second pass of applyFilter + the voxel index.  counts: per-voxel point counts (merged partials) or nullptr
// (count = length of the voxel's sorted point range, n_finite closes the last one)
int finalize_and_index(ndtb200_handle* h, uint32_t n_vox, uint32_t n_finite, const uint32_t* counts, bool records_done = false) {
  unsigned int* d_nvalid = h->d_scalar.as<unsigned int>() + 4;
  const int vblocks = static_cast<int>(((size_t)n_vox + kBuildThreads - 1) / kBuildThreads);
  if (!records_done) {
    CK(h->d_records.ensure((size_t)std::max<uint32_t>(n_vox, 1u) * sizeof(VoxelRecord)));
    CK(h->d_icov64.ensure((size_t)std::max<uint32_t>(n_vox, 1u) * 6 * sizeof(double)));
    CK(cudaMemsetAsync(d_nvalid, 0, sizeof(unsigned int), h->stream));
  }
  if (n_vox > 0 && !records_done) {
    finalize_voxels_kernel<<<vblocks, kBuildThreads, 0, h->stream>>>(
        h->d_moments.as<double>(), h->d_voxel_key.as<int32_t>(), h->d_voxel_start.as<uint32_t>(), counts, n_vox, n_finite,
        h->prm.min_points_per_voxel, h->prm.eig_ratio, h->d_records.as<VoxelRecord>(), h->d_icov64.as<double>(),
        d_nvalid, nullptr, nullptr, nullptr, nullptr);
    LAUNCHED(h);
  }

  {
    int st = prepare_index(h);
    if (st != NDTB200_OK) return st;
  }
  if (h->use_dense && n_vox > 0) {
    dense_fill_kernel<<<vblocks, kBuildThreads, 0, h->stream>>>(h->d_records.as<VoxelRecord>(), n_vox,
                                                                h->prm.min_points_per_voxel, h->d_dense.as<int32_t>());
    LAUNCHED(h);
  }
  return finish_hash_index(h, n_vox);
}

// The hash form of the voxel index (grids whose cell table does not fit): needs n_valid on the host, the only
// synchronisation left after the voxel count.  No-op when the direct-mapped table is in use.
int finish_hash_index(ndtb200_handle* h, uint32_t n_vox) {
  unsigned int* d_nvalid = h->d_scalar.as<unsigned int>() + 4;
  const int vblocks = static_cast<int>(((size_t)n_vox + kBuildThreads - 1) / kBuildThreads);
  if (!h->use_dense) {
    unsigned int n_valid = 0;
    CK(cudaMemcpyAsync(&n_valid, d_nvalid, sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->n_valid = n_valid;
    uint32_t cap = 1024;
    int log2cap = 10;
    while (cap < 4ull * n_valid && log2cap < 31) { cap <<= 1; ++log2cap; }
    h->hash_cap = cap;
    h->hash_shift = 32 - log2cap;
    CK(h->d_hash.ensure((size_t)cap * sizeof(HashSlot)));
    CK(cudaMemsetAsync(h->d_hash.p, 0xFF, (size_t)cap * sizeof(HashSlot), h->stream));
    if (n_vox > 0) {
      hash_insert_kernel<<<vblocks, kBuildThreads, 0, h->stream>>>(h->d_records.as<VoxelRecord>(), n_vox,
                                                                   h->prm.min_points_per_voxel, h->d_hash.as<HashSlot>(),
                                                                   cap - 1, h->hash_shift);
      LAUNCHED(h);
    }
  }
  return NDTB200_OK;
}

// ---- scan-sized clouds: the whole build (mode 0) or the VoxelGrid downsample (mode 1) in ONE launch ----------------
// Rule: a scan-sized cloud on a handle in latency mode (one pipeline owns the GPU: a single caller, the mapping loop) is
// built by the fused kernel (small_build.cuh), a COOPERATIVE launch over up to all SMs with a counter barrier in global
// memory between the phases.  Handles in throughput mode (many pairs in flight on their own streams) keep the staged
// kernels by default.  For them the fused kernel exists in a second launch flavour — the grid is ONE THREAD-BLOCK
// CLUSTER of 8 CTAs, an ordinary launch, phases separated by the hardware cluster barrier (no co-residency admission
// of a cooperative launch, which made round 1's 16-CTA cooperative builds bimodal: 11.8 k pairs/s on one box, 2.9 k
// and 6.4 k on others) — selected by NDTB200_BUILD_PATH=fused.  Measured on c3, pairs/s at 1 / 2 / 4 GPUs of one
// host: staged 10.9 k / 21.6 k / 38.0 k (12 launches + 2 synchronisations per pair), 8-CTA clusters 11.1 k / 22.2 k /
// 33.0 k (3 launches + 1 synchronisation), 16-CTA clusters 9.6 k (one 16-SM cluster per GPC at a time), 4-CTA 8.6 k:
// fewer launches do not buy throughput — the lanes are bound by the solves sharing the SMs, not by the launch rate —
// so the staged path, which scales better, stays the default.
// NDTB200_BUILD_PATH=staged / fused forces one path, NDTB200_FUSED_LAUNCH=coop / cluster one flavour (tests: all agree
// bit for bit).
constexpr int kFusedCtasThroughput = 8;  // the portable cluster size
bool use_fused_build(const ndtb200_handle* h, size_t n) {
  const char* e = getenv("NDTB200_BUILD_PATH");  // tests: "staged" / "fused" force one path (bit-identical results)
  if (e && std::strcmp(e, "staged") == 0) return false;
  if (e && std::strcmp(e, "fused") == 0) return n > 0 && n <= (size_t)0x7fffffff / 64;
  return n > 0 && n <= kSmallMaxPoints && (h->shape == 0 || h->prefer_fused_build);
}

// cluster launch of the fused build: largest supported power-of-two cluster size <= want (16 needs the non-portable
// opt-in; 8 is always available on sm_100)
int fused_cluster_size(const void* fn, int want) {
  static std::mutex mu;  // lanes (host threads) build concurrently
  std::lock_guard<std::mutex> lock(mu);
  static bool nonportable_ok[2] = {false, false};
  static const void* fns[2] = {nullptr, nullptr};
  int slot = -1;
  for (int i = 0; i < 2; ++i) {
    if (fns[i] == fn) { slot = i; break; }
    if (fns[i] == nullptr) { fns[i] = fn; nonportable_ok[i] = cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess; slot = i; break; }
  }
  const int cap = (slot >= 0 && nonportable_ok[slot]) ? 16 : 8;
  int c = 1;
  while (c * 2 <= want && c * 2 <= cap) c *= 2;
  return c;
}

// Leaves: h->grid (host copy), *n_vox_out; mode 0: records / icov64 / moments / voxel lists / n_valid counter;
// mode 1: centroids in h->d_out.  The sorted point indices end up in h->d_vals_a.  One host synchronisation.
int run_fused_build(ndtb200_handle* h, const float4* pts, size_t n, int dense, int mode, uint32_t* n_vox_out) {
  int G = std::min(h->num_sms, grid_for(n, kBuildThreads, h->num_sms));
  // throughput mode: many builds are in flight on different streams — each one is a single cluster of <= 16 CTAs
  bool cluster = h->shape == 1 && !h->prefer_fused_build;
  if (const char* e = getenv("NDTB200_FUSED_LAUNCH")) cluster = std::strcmp(e, "cluster") == 0 ? true : (std::strcmp(e, "coop") == 0 ? false : cluster);
  if (const char* e = getenv("NDTB200_FUSED_CTAS")) { const int c = atoi(e); if (c >= 1) G = std::min(G, c); }  // tuning
  const void* fn = mode == 0 ? (cluster ? (const void*)small_build_kernel<0, true> : (const void*)small_build_kernel<0, false>)
                             : (cluster ? (const void*)small_build_kernel<1, true> : (const void*)small_build_kernel<1, false>);
  if (cluster) G = fused_cluster_size(fn, std::min(G, kFusedCtasThroughput));
  const int ntiles = static_cast<int>((n + kSmallTile - 1) / kSmallTile);
  const int stiles = static_cast<int>((n + kScanTile - 1) / kScanTile);
  CK(h->d_mm_partial.ensure((size_t)G * 6 * sizeof(float)));
  CK(h->d_mm_finite.ensure((size_t)G * sizeof(unsigned int)));
  CK(h->d_grid.ensure(sizeof(GridDesc)));
  CK(h->d_keys_a.ensure(n * sizeof(uint32_t)));
  CK(h->d_keys_b.ensure(n * sizeof(uint32_t)));
  CK(h->d_vals_a.ensure(n * sizeof(uint32_t)));
  CK(h->d_vals_b.ensure(n * sizeof(uint32_t)));
  CK(h->d_hist.ensure((size_t)256 * ntiles * sizeof(uint32_t)));
  CK(h->d_scan_tmp.ensure((size_t)(stiles + 64 + 1024) * sizeof(uint32_t)));
  CK(h->d_voxel_key.ensure(n * sizeof(int32_t)));
  CK(h->d_voxel_start.ensure(n * sizeof(uint32_t)));
  if (mode == 0) {
    CK(h->d_moments.ensure(n * 9 * sizeof(double)));
    CK(h->d_records.ensure(n * sizeof(VoxelRecord)));
    CK(h->d_icov64.ensure(n * 6 * sizeof(double)));
  } else {
    CK(h->d_out.ensure(n * sizeof(float4)));
  }
  char* sc = h->d_scalar.as<char>();
  CK(cudaMemsetAsync(sc, 0, 256, h->stream));  // n_vox (0), n_valid (16), sorted-index pointer (160), barrier (192)
  SmallBuildArgs a;
  a.pts = pts; a.n = static_cast<uint32_t>(n); a.is_dense = dense; a.leaf = leaf_of(h);
  a.min_points = h->prm.min_points_per_voxel; a.eig_ratio = h->prm.eig_ratio; a.mode = mode;
  a.mm_partial = h->d_mm_partial.as<float>(); a.mm_finite = h->d_mm_finite.as<unsigned int>();
  a.keys_a = h->d_keys_a.as<uint32_t>(); a.keys_b = h->d_keys_b.as<uint32_t>();
  a.vals_a = h->d_vals_a.as<uint32_t>(); a.vals_b = h->d_vals_b.as<uint32_t>();
  a.hist = h->d_hist.as<uint32_t>(); a.tile_heads = h->d_scan_tmp.as<uint32_t>();
  a.digit_totals = h->d_scan_tmp.as<uint32_t>() + ((stiles + 63) & ~63);
  CK(cudaMemsetAsync(a.digit_totals, 0, 1024 * sizeof(uint32_t), h->stream));
  a.barrier = reinterpret_cast<unsigned int*>(sc + 192);
  a.grid = h->d_grid.as<GridDesc>(); a.n_vox = reinterpret_cast<uint32_t*>(sc); a.n_valid = reinterpret_cast<unsigned int*>(sc + 16);
  a.voxel_key = h->d_voxel_key.as<int32_t>(); a.voxel_start = h->d_voxel_start.as<uint32_t>();
  a.moments = h->d_moments.as<double>(); a.records = h->d_records.as<VoxelRecord>(); a.icov64 = h->d_icov64.as<double>();
  a.centroids = h->d_out.as<float4>();
  a.sorted_idx_out = reinterpret_cast<uint32_t**>(sc + 160);
  void* args[] = {(void*)&a};
  if (cluster) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G);
    cfg.blockDim = dim3(kBuildThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = G;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CK(cudaLaunchKernelExC(&cfg, fn, args));
  } else {
    CK(cudaLaunchCooperativeKernel(fn, dim3(G), dim3(kBuildThreads), args, 0, h->stream));
  }
  LAUNCHED(h);
  struct { uint32_t n_vox; uint32_t pad[39]; uint32_t* sorted; } back;
  static_assert(offsetof(decltype(back), sorted) == 160, "scalar block layout");
  CK(cudaMemcpyAsync(&h->grid, h->d_grid.p, sizeof(GridDesc), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(&back, sc, sizeof(back), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (back.sorted == h->d_vals_b.as<uint32_t>()) { std::swap(h->d_vals_a, h->d_vals_b); std::swap(h->d_keys_a, h->d_keys_b); }
  *n_vox_out = back.n_vox;
  return NDTB200_OK;
}

// per-voxel fp64 moments of the sorted target into h->d_moments (the partial builds of the sharded path, and the parity
// dump of a map whose build fused the moments away)
int compute_moments(ndtb200_handle* h, uint32_t n_vox, uint32_t n_finite) {
  CK(h->d_moments.ensure((size_t)std::max<uint32_t>(n_vox, 1u) * 9 * sizeof(double)));
  h->moments_valid = true;
  if (n_vox == 0) return NDTB200_OK;
  // payload-sorted map: the sorted cloud is still in d_pay_a (sequential reads, no index array)
  const bool seq = h->idx_lazy;
  const float4* pts = seq ? h->d_pay_a.as<float4>() : target_pts(h);
  const uint32_t* va = seq ? nullptr : h->d_vals_a.as<uint32_t>();
  const double avg = static_cast<double>(n_finite) / std::max<uint32_t>(1u, n_vox);
  const int group = avg >= 48.0 ? 32 : (avg >= 10.0 ? 8 : 4);
  const int blocks = static_cast<int>(((size_t)n_vox * group + kBuildThreads - 1) / kBuildThreads);
#define NDTB200_MOMENTS(G, SEQ) \
  voxel_moments_kernel<G, SEQ><<<blocks, kBuildThreads, 0, h->stream>>>(pts, va, h->d_voxel_start.as<uint32_t>(), n_vox, n_finite, h->d_moments.as<double>())
  if (seq) {
    if (group == 32) NDTB200_MOMENTS(32, true);
    else if (group == 8) NDTB200_MOMENTS(8, true);
    else NDTB200_MOMENTS(4, true);
  } else {
    if (group == 32) NDTB200_MOMENTS(32, false);
    else if (group == 8) NDTB200_MOMENTS(8, false);
    else NDTB200_MOMENTS(4, false);
  }
#undef NDTB200_MOMENTS
  LAUNCHED(h);
  return NDTB200_OK;
}

int build_map_ex(ndtb200_handle* h, const BuildOpts& o) {
  clear_map(h);
  h->map_is_merged = false;
  std::memset(&h->grid, 0, sizeof(GridDesc));
  for (int a = 0; a < 3; ++a) h->grid.leaf[a] = h->prm.resolution;
  const size_t n = h->n_target;
  if (!h->has_target || (n == 0 && !o.partial_only)) {
    h->map_status = NDTB200_ERR_NO_INPUT;
    int st = ensure_empty_hash(h);
    return st != NDTB200_OK ? st : NDTB200_ERR_NO_INPUT;
  }
  if (n > 0xFFFFFFF0ull) { h->err = "target cloud too large (>= 2^32 points)"; return NDTB200_ERR_INVALID; }
  const float4* pts = target_pts(h);
  const int dense = h->target_dense ? 1 : 0;

  if (!o.partial_only && !o.forced_min && use_fused_build(h, n)) {  // scan-sized cloud: one cooperative launch
    uint32_t n_vox = 0;
    int st = run_fused_build(h, pts, n, dense, 0, &n_vox);
    if (st != NDTB200_OK) return st;
    if (h->grid.n_finite == 0) {
      h->map_status = NDTB200_ERR_NO_INPUT;
      st = ensure_empty_hash(h);
      return st != NDTB200_OK ? st : NDTB200_ERR_NO_INPUT;
    }
    if (h->grid.overflow) {
      h->map_status = NDTB200_ERR_GRID_OVERFLOW;
      st = ensure_empty_hash(h);
      return st != NDTB200_OK ? st : NDTB200_ERR_GRID_OVERFLOW;
    }
    h->n_voxels = n_vox;
    h->moments_valid = true;  // the fused kernel leaves the moment rows in d_moments
    st = finalize_and_index(h, n_vox, static_cast<uint32_t>(h->grid.n_finite), nullptr, /*records_done=*/true);
    if (st != NDTB200_OK) return st;
    h->map_status = NDTB200_OK;
    return NDTB200_OK;
  }

  // 1. bounding box + grid description
  {
    int st = compute_grid(h, pts, n, dense, o);
    if (st != NDTB200_OK) return st;
  }
  if (h->grid.n_finite == 0 && !o.partial_only) {
    h->map_status = NDTB200_ERR_NO_INPUT;
    int st = ensure_empty_hash(h);
    return st != NDTB200_OK ? st : NDTB200_ERR_NO_INPUT;
  }
  if (h->grid.overflow) {  // voxel_grid_covariance_omp_impl.hpp:79-84: warn, leave the map empty
    h->map_status = NDTB200_ERR_GRID_OVERFLOW;
    int st = ensure_empty_hash(h);
    return st != NDTB200_OK ? st : NDTB200_ERR_GRID_OVERFLOW;
  }
  h->n_partials = 0;
  if (h->grid.n_finite == 0) { h->map_status = NDTB200_OK; return NDTB200_OK; }  // empty slice of a sharded build

  // 2. keys
  uint32_t sentinel = 0;
  const int passes = passes_for(h->grid, !dense, &sentinel);
  const bool payload = use_payload_sort(h, n);
  {
    int st = sort_meta_reset(h);
    if (st != NDTB200_OK) return st;
  }
  const int key_blocks = grid_for(n, kBuildThreads * 4, h->num_sms * 8);
  uint32_t n_vox = 0;
  if (payload) {
    // large cloud: digit histograms only, then the passes carry the points (3. sort) + 4. segment heads
    voxel_key_kernel<<<key_blocks, kBuildThreads, 0, h->stream>>>(pts, n, dense, h->d_grid.as<GridDesc>(), sentinel, nullptr, nullptr,
                                                                   sort_meta_hist(h), passes);
    LAUNCHED(h);
    int st = sort_payload(h, pts, n, dense, sentinel, passes);
    if (st != NDTB200_OK) return st;
    st = segment_heads(h, n, sentinel, &n_vox);
    if (st != NDTB200_OK) return st;
    h->idx_lazy = true;
  } else {
    CK(h->d_keys_a.ensure(n * sizeof(uint32_t)));
    CK(h->d_keys_b.ensure(n * sizeof(uint32_t)));
    CK(h->d_vals_a.ensure(n * sizeof(uint32_t)));
    CK(h->d_vals_b.ensure(n * sizeof(uint32_t)));
    voxel_key_kernel<<<key_blocks, kBuildThreads, 0, h->stream>>>(pts, n, dense, h->d_grid.as<GridDesc>(), sentinel,
                                                                   h->d_keys_a.as<uint32_t>(), nullptr, sort_meta_hist(h), passes);
    LAUNCHED(h);
    // 3. stable sort of (key, point index), one-sweep passes + 4. occupied voxels = segment heads
    int st = sort_and_segment(h, n, sentinel, passes, &n_vox, /*hist_ready=*/true);
    if (st != NDTB200_OK) return st;
  }
  h->n_voxels = n_vox;
  const float4* mpts = payload ? h->d_pay_a.as<float4>() : pts;  // what the moments read: the sorted cloud, or a gather through the sorted indices
  const uint32_t* va = payload ? nullptr : h->d_vals_a.as<uint32_t>();
  const uint32_t n_finite = static_cast<uint32_t>(h->grid.n_finite);
  const double avg = static_cast<double>(n_finite) / std::max<uint32_t>(1u, n_vox);
  const int group = avg >= 48.0 ? 32 : (avg >= 10.0 ? 8 : 4);  // lanes per voxel
  const int blocks = static_cast<int>(((size_t)n_vox + kBuildThreads - 1) / kBuildThreads);  // voxel_build_kernel: kBuildThreads voxels per CTA

  if (o.partial_only) {  // 5'. moments only: leave {voxel_key, count, moments} for the exchange
    int st = compute_moments(h, n_vox, n_finite);
    if (st != NDTB200_OK) return st;
    CK(h->d_voxel_count.ensure((size_t)std::max<uint32_t>(n_vox, 1u) * sizeof(uint32_t)));
    const int vblocks = static_cast<int>(((size_t)n_vox + kBuildThreads - 1) / kBuildThreads);
    segment_counts_kernel<<<vblocks, kBuildThreads, 0, h->stream>>>(h->d_voxel_start.as<uint32_t>(), n_vox, n_finite,
                                                                     h->d_voxel_count.as<uint32_t>());
    LAUNCHED(h);
    h->n_partials = n_vox;
    h->map_status = NDTB200_OK;
    return NDTB200_OK;
  }

  // 5. + 6. moments, finalize and the cell-table entries in one kernel (voxel_build_kernel)
  {
    int st = prepare_index(h);
    if (st != NDTB200_OK) return st;
  }
  CK(h->d_records.ensure((size_t)std::max<uint32_t>(n_vox, 1u) * sizeof(VoxelRecord)));
  CK(h->d_icov64.ensure((size_t)std::max<uint32_t>(n_vox, 1u) * 6 * sizeof(double)));
  unsigned int* d_nvalid = h->d_scalar.as<unsigned int>() + 4;
  CK(cudaMemsetAsync(d_nvalid, 0, sizeof(unsigned int), h->stream));
  h->moments_valid = false;
  if (n_vox > 0) {
    int32_t* table = h->use_dense ? h->d_dense.as<int32_t>() : nullptr;
#define NDTB200_VOXEL_BUILD(G, SEQ)                                                                                         \
    voxel_build_kernel<G, SEQ><<<blocks, kBuildThreads, 0, h->stream>>>(mpts, va, h->d_voxel_key.as<int32_t>(), h->d_voxel_start.as<uint32_t>(), \
        n_vox, n_finite, h->prm.min_points_per_voxel, h->prm.eig_ratio, h->d_records.as<VoxelRecord>(), h->d_icov64.as<double>(), d_nvalid, table)
    if (payload) {
      if (group == 32) NDTB200_VOXEL_BUILD(32, true);
      else if (group == 8) NDTB200_VOXEL_BUILD(8, true);
      else NDTB200_VOXEL_BUILD(4, true);
    } else {
      if (group == 32) NDTB200_VOXEL_BUILD(32, false);
      else if (group == 8) NDTB200_VOXEL_BUILD(8, false);
      else NDTB200_VOXEL_BUILD(4, false);
    }
#undef NDTB200_VOXEL_BUILD
    LAUNCHED(h);
  }
  {
    int st = finish_hash_index(h, n_vox);
    if (st != NDTB200_OK) return st;
  }
  h->map_status = NDTB200_OK;
  return NDTB200_OK;
}

int build_map(ndtb200_handle* h) { return build_map_ex(h, BuildOpts()); }

// Merge the per-voxel partials of all ranks (concatenated in rank order) into this handle's map: stable sort by key,
// per-voxel sums in rank order (bit-identical on every rank), then the usual finalize + index.
// grid description of the COMMON bounding box (all ranks) -> h->grid; one synchronisation
int grid_from_global_box(ndtb200_handle* h, const float* gmin, const float* gmax, long long n_finite_total) {
  std::memset(&h->grid, 0, sizeof(GridDesc));
  float box[6] = {gmin[0], gmin[1], gmin[2], gmax[0], gmax[1], gmax[2]};
  CK(h->d_mm_partial.ensure(6 * sizeof(float)));
  CK(h->d_mm_finite.ensure(sizeof(unsigned int)));
  CK(h->d_grid.ensure(sizeof(GridDesc)));
  const unsigned int nf_flag = n_finite_total > 0 ? 1u : 0u;
  CK(cudaMemcpyAsync(h->d_mm_partial.p, box, sizeof(box), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_mm_finite.p, &nf_flag, sizeof(nf_flag), cudaMemcpyHostToDevice, h->stream));
  ForcedBox fb;
  fb.use = 0;
  for (int a = 0; a < 3; ++a) fb.mn[a] = fb.mx[a] = 0.f;
  grid_setup_kernel<<<1, kBuildThreads, 0, h->stream>>>(h->d_mm_partial.as<float>(), h->d_mm_finite.as<unsigned int>(), 1,
                                                        leaf_of(h), fb, h->d_grid.as<GridDesc>());
  LAUNCHED(h);
  CK(cudaMemcpyAsync(&h->grid, h->d_grid.p, sizeof(GridDesc), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->grid.n_finite = static_cast<int32_t>(std::min<long long>(n_finite_total, 0x7fffffffll));
  return NDTB200_OK;
}

// records_only: merge + finalize the given partials (this rank's share of the key space) and stop before the index —
// the records are then all-gathered and installed everywhere by set_map_from_records.
int build_from_partials(ndtb200_handle* h, const float* gmin, const float* gmax, long long n_finite_total,
                        const uint32_t* d_keys, const uint32_t* d_counts, const double* d_moments, size_t total,
                        bool records_only = false) {
  clear_map(h);
  {
    int st = grid_from_global_box(h, gmin, gmax, n_finite_total);
    if (st != NDTB200_OK) return st;
  }
  if (n_finite_total == 0 || (total == 0 && !records_only)) {
    h->map_status = NDTB200_ERR_NO_INPUT;
    int st = ensure_empty_hash(h);
    return st != NDTB200_OK ? st : NDTB200_ERR_NO_INPUT;
  }
  if (h->grid.overflow) {
    h->map_status = NDTB200_ERR_GRID_OVERFLOW;
    int st = ensure_empty_hash(h);
    return st != NDTB200_OK ? st : NDTB200_ERR_GRID_OVERFLOW;
  }
  if (total == 0) { h->n_voxels = 0; h->map_status = NDTB200_OK; return NDTB200_OK; }  // an owner without voxels
  uint32_t sentinel = 0;
  const int passes = passes_for(h->grid, false, &sentinel);
  CK(h->d_keys_a.ensure(total * sizeof(uint32_t)));
  CK(h->d_keys_b.ensure(total * sizeof(uint32_t)));
  CK(h->d_vals_a.ensure(total * sizeof(uint32_t)));
  CK(h->d_vals_b.ensure(total * sizeof(uint32_t)));
  CK(cudaMemcpyAsync(h->d_keys_a.p, d_keys, total * sizeof(uint32_t), cudaMemcpyDeviceToDevice, h->stream));
  uint32_t n_vox = 0;
  {
    int st = sort_and_segment(h, total, sentinel, passes, &n_vox);
    if (st != NDTB200_OK) return st;
  }
  h->n_voxels = n_vox;
  CK(h->d_moments.ensure((size_t)std::max<uint32_t>(n_vox, 1u) * 9 * sizeof(double)));
  CK(h->d_voxel_count.ensure((size_t)std::max<uint32_t>(n_vox, 1u) * sizeof(uint32_t)));
  const int vblocks = static_cast<int>(((size_t)n_vox + kBuildThreads - 1) / kBuildThreads);
  merge_partials_kernel<<<vblocks, kBuildThreads, 0, h->stream>>>(h->d_vals_a.as<uint32_t>(), h->d_voxel_start.as<uint32_t>(), n_vox,
                                                                   static_cast<uint32_t>(total), d_counts, d_moments,
                                                                   h->d_voxel_count.as<uint32_t>(), h->d_moments.as<double>());
  LAUNCHED(h);
  h->moments_valid = true;
  if (records_only) {  // finalize only: the index is built from the all-gathered records
    unsigned int* d_nvalid = h->d_scalar.as<unsigned int>() + 4;
    CK(h->d_records.ensure((size_t)std::max<uint32_t>(n_vox, 1u) * sizeof(VoxelRecord)));
    CK(h->d_icov64.ensure((size_t)std::max<uint32_t>(n_vox, 1u) * 6 * sizeof(double)));
    CK(cudaMemsetAsync(d_nvalid, 0, sizeof(unsigned int), h->stream));
    if (n_vox > 0) {
      finalize_voxels_kernel<<<vblocks, kBuildThreads, 0, h->stream>>>(
          h->d_moments.as<double>(), h->d_voxel_key.as<int32_t>(), h->d_voxel_start.as<uint32_t>(), h->d_voxel_count.as<uint32_t>(), n_vox, 0u,
          h->prm.min_points_per_voxel, h->prm.eig_ratio, h->d_records.as<VoxelRecord>(), h->d_icov64.as<double>(), d_nvalid, nullptr, nullptr,
          nullptr, nullptr);
      LAUNCHED(h);
    }
    h->map_is_merged = true;
    h->map_status = NDTB200_OK;
    return NDTB200_OK;
  }
  {
    int st = finalize_and_index(h, n_vox, 0u, h->d_voxel_count.as<uint32_t>());
    if (st != NDTB200_OK) return st;
  }
  h->map_is_merged = true;
  h->map_status = NDTB200_OK;
  return NDTB200_OK;
}

// Sorted point indices of a payload-sorted map, on first use (getFitnessScore, KDTREE centroids): the (key, index) sort
// of the same keys — stable, so the order equals the payload sort's and d_voxel_start stays valid.
int ensure_sorted_indices(ndtb200_handle* h) {
  if (!h->idx_lazy) return NDTB200_OK;
  const size_t n = h->n_target;
  uint32_t sentinel = 0;
  const int passes = passes_for(h->grid, !h->target_dense, &sentinel);
  CK(h->d_keys_a.ensure(n * sizeof(uint32_t)));
  CK(h->d_keys_b.ensure(n * sizeof(uint32_t)));
  CK(h->d_vals_a.ensure(n * sizeof(uint32_t)));
  CK(h->d_vals_b.ensure(n * sizeof(uint32_t)));
  int st = sort_meta_reset(h);
  if (st != NDTB200_OK) return st;
  voxel_key_kernel<<<grid_for(n, kBuildThreads * 4, h->num_sms * 8), kBuildThreads, 0, h->stream>>>(
      target_pts(h), n, h->target_dense ? 1 : 0, h->d_grid.as<GridDesc>(), sentinel, h->d_keys_a.as<uint32_t>(), nullptr, sort_meta_hist(h), passes);
  LAUNCHED(h);
  st = sort_pairs(h, n, passes, /*hist_ready=*/true);
  if (st != NDTB200_OK) return st;
  h->idx_lazy = false;  // d_vals_a is valid from here on (compute_moments goes back to the gather, same sums)
  return NDTB200_OK;
}

// cell table over ALL occupied voxels (shared by getFitnessScore and the KDTREE mode)
int ensure_cell_table(ndtb200_handle* h) {
  if (h->cell_all_valid) return NDTB200_OK;
  const unsigned long long ncell = static_cast<unsigned long long>(h->grid.div_b[0]) * static_cast<unsigned long long>(h->grid.div_b[1]) *
                                   static_cast<unsigned long long>(h->grid.div_b[2]);
  if (ncell == 0 || ncell * sizeof(int32_t) > (2ull << 30)) { h->err = "cell table does not fit (grid too large)"; return NDTB200_ERR_INVALID; }
  CK(h->d_cell_all.ensure(ncell * sizeof(int32_t)));
  CK(cudaMemsetAsync(h->d_cell_all.p, 0xFF, ncell * sizeof(int32_t), h->stream));
  const uint32_t nv = static_cast<uint32_t>(h->n_voxels);
  if (nv > 0) {
    cell_table_fill_kernel<<<(nv + 255) / 256, 256, 0, h->stream>>>(h->d_voxel_key.as<int32_t>(), nv, h->d_cell_all.as<int32_t>());
    LAUNCHED(h);
  }
  h->cell_all_valid = true;
  return NDTB200_OK;
}

// KDTREE mode (voxel_grid_covariance_omp.h:476-505): the fp32 centroid of every voxel (Leaf::centroid: points added in
// input order in fp32, then divided by the count, …_impl.hpp:228-262, 288) + the cell table over all occupied voxels
int ensure_kdtree_index(ndtb200_handle* h) {
  if (h->map_status != NDTB200_OK || h->n_voxels == 0) return NDTB200_OK;  // empty map: every probe misses
  if (h->map_is_merged) { h->err = "KDTREE mode is not available on a map merged from sharded partials"; return NDTB200_ERR_INVALID; }
  int st = ensure_cell_table(h);
  if (st != NDTB200_OK) return st;
  if (!h->centroids_valid) {
    st = ensure_sorted_indices(h);
    if (st != NDTB200_OK) return st;
    const uint32_t nv = static_cast<uint32_t>(h->n_voxels);
    CK(h->d_centroid.ensure((size_t)nv * sizeof(float4)));
    voxel_centroid_kernel<<<(nv + kBuildThreads - 1) / kBuildThreads, kBuildThreads, 0, h->stream>>>(
        target_pts(h), h->d_vals_a.as<uint32_t>(), h->d_voxel_start.as<uint32_t>(), nv,
        static_cast<uint32_t>(h->grid.n_finite), h->d_centroid.as<float4>());
    LAUNCHED(h);
    h->centroids_valid = true;
  }
  return NDTB200_OK;
}

MapView make_view(const ndtb200_handle* h) {
  MapView m;
  m.records = h->d_records.as<VoxelRecord>();
  m.icov64 = h->d_icov64.as<double>();
  m.hash = h->d_hash.as<HashSlot>();
  m.dense = (h->use_dense && h->n_voxels > 0) ? h->d_dense.as<int32_t>() : nullptr;
  m.cell_all = (h->cell_all_valid && h->n_voxels > 0) ? h->d_cell_all.as<int32_t>() : nullptr;
  m.centroids = h->centroids_valid ? h->d_centroid.as<float4>() : nullptr;
  m.kd_r2 = static_cast<float>(static_cast<double>(h->prm.resolution) * static_cast<double>(h->prm.resolution));
  m.hash_mask = h->hash_cap - 1;
  m.hash_shift = h->hash_shift;
  for (int a = 0; a < 3; ++a) {
    m.min_b[a] = h->grid.min_b[a];
    m.max_b[a] = h->grid.max_b[a];
    m.mul[a] = h->grid.mul[a];
    m.leaf[a] = h->grid.leaf[a];
    m.inv_leaf[a] = 1.0f / h->grid.leaf[a];
  }
  if (h->n_voxels == 0) {  // empty map: every bounds test fails
    for (int a = 0; a < 3; ++a) { m.min_b[a] = 1; m.max_b[a] = 0; m.mul[a] = 0; }
  }
  m.min_points = h->prm.min_points_per_voxel;
  m.n_cells = static_cast<unsigned long long>(h->grid.div_b[0]) * static_cast<unsigned long long>(h->grid.div_b[1]) *
              static_cast<unsigned long long>(h->grid.div_b[2]);
  m.n_records = static_cast<uint32_t>(h->n_voxels);
  m.pad = 0;
  return m;
}

// ---- guess -> pose vector (host; tiny) ---------------------------------------------------------
// Transform<float,3,Affine>::rotation() (polar factor via SVD) then eulerAngles(0,1,2) as Eigen 3.3
// (ndt_omp_impl.hpp:103-111).
void rotation_polar_host(const float* T /*row-major 3x4*/, float R[3][3]) {
  double W[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) W[i][j] = T[i * 4 + j];
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int k = 0; k < 3; ++k) { alpha += W[k][p] * W[k][p]; beta += W[k][q] * W[k][q]; gamma += W[k][p] * W[k][q]; }
        if (gamma == 0.0 || std::fabs(gamma) <= 1e-17 * std::sqrt(alpha * beta)) continue;
        rotated = true;
        double zeta = (beta - alpha) / (2.0 * gamma);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
        for (int k = 0; k < 3; ++k) {
          double wp = W[k][p], wq = W[k][q];
          W[k][p] = c * wp - s * wq; W[k][q] = s * wp + c * wq;
          double vp = V[k][p], vq = V[k][q];
          V[k][p] = c * vp - s * vq; V[k][q] = s * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  double U[3][3];
  for (int j = 0; j < 3; ++j) {
    double nrm = std::sqrt(W[0][j] * W[0][j] + W[1][j] * W[1][j] + W[2][j] * W[2][j]);
    for (int k = 0; k < 3; ++k) U[k][j] = (nrm > 0) ? W[k][j] / nrm : (k == j ? 1.0 : 0.0);
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += U[i][k] * V[j][k];
      R[i][j] = static_cast<float>(s);
    }
}

void euler_angles_012_host(const float R[3][3], float res[3]) {
  const float pi = static_cast<float>(M_PI);
  res[0] = std::atan2(R[1][2], R[2][2]);
  const float c2 = std::sqrt(R[0][0] * R[0][0] + R[0][1] * R[0][1]);
  if (res[0] > 0.0f) {
    res[0] -= pi;
    res[1] = std::atan2(-R[0][2], -c2);
  } else {
    res[1] = std::atan2(-R[0][2], c2);
  }
  const float s1 = std::sin(res[0]), c1 = std::cos(res[0]);
  res[2] = std::atan2(s1 * R[2][0] - c1 * R[1][0], c1 * R[1][1] - s1 * R[2][1]);
  res[0] = -res[0]; res[1] = -res[1]; res[2] = -res[2];
}

void gauss_constants(const ndtb200_params& p, double& d1, double& d2, double& d3) {
  // ndt_omp_impl.hpp:86-93
  const double c1 = 10.0 * (1 - p.outlier_ratio);
  const double c2 = p.outlier_ratio / std::pow(static_cast<double>(p.resolution), 3);
  d3 = -std::log(c2);
  d1 = -std::log(c1 + c2) - d3;
  d2 = -2 * std::log((-std::log(c1 * std::exp(-0.5) + c2) - d3) / d1);
}

template <int METHOD>
int query_coop_blocks(ndtb200_handle* h) {
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ndt_align_kernel<METHOD, kThreadsLatency>, kThreadsLatency, 0));
  if (per_sm < 1) { h->err = "align kernel does not fit on an SM"; return NDTB200_ERR_CUDA; }
  h->coop_blocks[METHOD][0] = h->num_sms;  // one 1024-thread CTA per SM
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ndt_align_kernel<METHOD, kThreadsThroughput>, kThreadsThroughput, 0));
  if (per_sm < 1) { h->err = "align kernel (throughput shape) does not fit on an SM"; return NDTB200_ERR_CUDA; }
  h->coop_blocks[METHOD][1] = h->num_sms;  // one 256-thread CTA per SM per solve: per_sm solves can be co-resident
  return NDTB200_OK;
}

int ensure_aux(ndtb200_handle* h) {  // scratch handle: VoxelGrid filters and the source sort keep the map's build buffers untouched
  if (h->aux) return NDTB200_OK;
  const int st = ndtb200_create(&h->aux, h->device);
  if (st != NDTB200_OK) h->err = "could not create the scratch state (ndtb200_voxelgrid_filter / source sort)";
  return st;
}

// Sources of at least this many points are aligned from a copy sorted by voxel key (NDTB200_SORT_SOURCE_MIN overrides;
// 0 = never).  A LiDAR scan in ring order is already coherent and small enough to live in L1 / L2; a merged or
// map-sized source in arbitrary order is not: its warps would touch 32 different voxels per probe.
size_t sort_source_min() {
  const char* e = getenv("NDTB200_SORT_SOURCE_MIN");  // read per call: tests switch it inside one process
  const long long v = e ? atoll(e) : 262144;
  return v <= 0 ? ~size_t(0) : static_cast<size_t>(v);
}

// The sort changes only the ORDER in which the per-point contributions are summed (fixed and reproducible for a given
// cloud): keys are the cells of the source's own bounding-box grid at the map resolution, untransformed — a rigid
// guess keeps neighbours neighbours.  Once per set_source; the output cloud of align() keeps the caller's order.
int ensure_sorted_source(ndtb200_handle* h) {
  if (h->source_sorted_valid || h->source_sort_failed) return NDTB200_OK;
  const size_t n = h->n_source;
  if (n < sort_source_min() || n > 0xFFFFFFF0ull) { h->source_sort_failed = true; return NDTB200_OK; }
  int st = ensure_aux(h);
  if (st != NDTB200_OK) return st;
  ndtb200_handle* a = h->aux;
  CK(cudaStreamSynchronize(h->stream));  // the source copy was enqueued on h's stream; the sort runs on the scratch handle's
  for (int k = 0; k < 3; ++k) a->vg_leaf[k] = h->prm.resolution;
  const float4* pts = h->d_source.as<float4>();
  st = compute_grid(a, pts, n, /*dense=*/0, BuildOpts());
  if (st != NDTB200_OK) { h->err = a->err; return st; }
  if (a->grid.overflow || a->grid.n_finite == 0) { h->source_sort_failed = true; return NDTB200_OK; }  // keep the caller's order
  uint32_t sentinel = 0;
  const int passes = passes_for(a->grid, true, &sentinel);
  CK2(a, a->d_keys_a.ensure(n * sizeof(uint32_t)));
  CK2(a, a->d_keys_b.ensure(n * sizeof(uint32_t)));
  CK2(a, a->d_vals_a.ensure(n * sizeof(uint32_t)));
  CK2(a, a->d_vals_b.ensure(n * sizeof(uint32_t)));
  st = sort_meta_reset(a);
  if (st != NDTB200_OK) return st;
  voxel_key_kernel<<<grid_for(n, kBuildThreads * 4, a->num_sms * 8), kBuildThreads, 0, a->stream>>>(
      pts, n, 0, a->d_grid.as<GridDesc>(), sentinel, a->d_keys_a.as<uint32_t>(), nullptr, sort_meta_hist(a), passes);
  LAUNCHED(a);
  st = sort_pairs(a, n, passes, /*hist_ready=*/true);
  if (st != NDTB200_OK) { h->err = a->err; return st; }
  CK(h->d_source_sorted.ensure(n * sizeof(float4)));
  gather_points_kernel<<<grid_for(n, kBuildThreads * 4, a->num_sms * 16), kBuildThreads, 0, a->stream>>>(
      pts, a->d_vals_a.as<uint32_t>(), n, h->d_source_sorted.as<float4>());
  LAUNCHED(a);
  CK(cudaStreamSynchronize(a->stream));
  h->source_sorted_valid = true;
  return NDTB200_OK;
}

// Enqueue one launch of the persistent kernel (no host synchronisation).
int launch_align(ndtb200_handle* h, int mode, const double p0[6], const float T0[12], int has_guess, int eval_hessian,
                 int emulate_world = 0) {
  if (!h->has_source || (h->n_source == 0 && h->comm_world == 1)) { h->err = "no input source"; return NDTB200_ERR_NO_INPUT; }
  const int method = h->prm.search_method;
  if (method < NDTB200_KDTREE || method > NDTB200_DIRECT1) {
    h->err = "invalid search method";
    return NDTB200_ERR_INVALID;
  }
  if (method == NDTB200_KDTREE) {
    const int st = ensure_kdtree_index(h);
    if (st != NDTB200_OK) return st;
  }
  {
    const int st = ensure_sorted_source(h);
    if (st != NDTB200_OK) return st;
  }
  const float4* d_src = h->source_sorted_valid ? h->d_source_sorted.as<float4>() : h->d_source.as<float4>();
  AlignParams prm;
  gauss_constants(h->prm, prm.d1, prm.d2, prm.d3);
  prm.step_size = h->prm.step_size;
  prm.trans_eps = h->prm.trans_eps;
  prm.max_iterations = h->prm.max_iterations;
  prm.mode = mode;
  prm.eval_hessian = eval_hessian;
  prm.has_guess = has_guess;
  for (int i = 0; i < 6; ++i) prm.p0[i] = p0[i];
  for (int i = 0; i < 12; ++i) prm.T0[i] = T0[i];
  prm.n_source = static_cast<int>(h->n_source);
  prm.trace_cap = ndtb200_handle::kTraceCap;
  static const int rot_env = getenv("NDTB200_ROTATE") ? atoi(getenv("NDTB200_ROTATE")) : 0;
  prm.rot = rot_env;
  prm.launch_tag = (++h->launch_seq) << 12;  // 4096 evaluations per launch before tags could repeat
  compute_angle_tables(p0, prm.tab0);

  const int shape = (h->comm_world > 1 || emulate_world > 1) ? 0 : h->shape;  // the sharded exchange needs >= 8 warps per CTA
  const int max_blocks = h->coop_blocks[method][shape];
  // every SM takes part as soon as there is one 32-point group per CTA (the kernel deals groups to warps round-robin)
  int blocks = grid_for(h->n_source, 32, max_blocks);
  if (const char* e = getenv("NDTB200_MAX_CTAS")) { const int c = atoi(e); if (c >= 1) blocks = std::min(blocks, c); }  // tests: small grids
  CK(h->d_partials.ensure((size_t)2 * max_blocks * kNVP * sizeof(double)));  // double-buffered by evaluation parity
  if (h->d_totals.p == nullptr) {
    CK(h->d_totals.ensure(3 * kNVP * sizeof(double)));
    CK(cudaMemsetAsync(h->d_totals.p, 0, 3 * kNVP * sizeof(double), h->stream));
  }
  if (h->d_sync.p == nullptr) {  // barrier words: zeroed once; every launch leaves them zeroed again
    CK(h->d_sync.ensure(64));
    CK(cudaMemsetAsync(h->d_sync.p, 0, 64, h->stream));
  }
  CK(h->d_result.ensure(sizeof(AlignResultDev)));
  CK(h->d_trace.ensure(ndtb200_handle::kTraceCap * sizeof(TraceRec)));

  AlignWorkspace ws;
  ws.partials = h->d_partials.as<double>();
  ws.totals = h->d_totals.as<double>();
  ws.sync = h->d_sync.as<unsigned int>();
  ws.result = h->d_result.as<AlignResultDev>();
  ws.trace = h->d_trace.as<TraceRec>();
  ws.world = h->comm_world;
  ws.rank = h->comm_rank;
  ws.n_source_total = h->comm_world > 1 ? h->comm_n_total : static_cast<long long>(h->n_source);
  for (int r = 0; r < kMaxRanks; ++r) ws.mail[r] = h->mail_ptrs[r];
  ws.vranks = 1;
  ws.pad = 0;
  ws.vr = nullptr;
  // the solve writes align()'s output cloud itself when the entry point knows the caller wants it (one launch less per
  // align); not for sharded / emulated solves (a rank holds a slice) and not for the parity-evaluation modes
  ws.out = nullptr;
  ws.out_src = nullptr;
  h->out_fused = false;
  if (h->want_fused_out && mode == MODE_ALIGN && h->comm_world == 1 && emulate_world <= 1 && h->n_source > 0) {
    CK(h->d_out.ensure(h->n_source * sizeof(float4)));
    ws.out = h->d_out.as<float4>();
    ws.out_src = h->d_source.as<float4>();
    h->out_fused = true;
  }
  h->want_fused_out = false;
  if (emulate_world > 1) {
    // `emulate_world` ranks of a source-sharded solve inside ONE cooperative launch on this GPU (see VirtualRank): the
    // grid is divided evenly, every rank gets its own rows / barrier words / result block / mailbox and the contiguous
    // source range a rank on its own GPU would get (slices start on a 32-point group boundary).
    const int W = emulate_world;
    const int cpr = std::max(1, std::min(blocks, max_blocks) / W);
    blocks = cpr * W;
    const size_t part_b = (size_t)cpr * kNVP * sizeof(double), tot_b = 3 * kNVP * sizeof(double), sync_b = 64;
    const size_t res_b = (sizeof(AlignResultDev) + 63) & ~size_t(63), mail_b = (size_t)2 * kMaxRanks * kNVP * 2 * sizeof(unsigned long long);
    const size_t per_rank = ((part_b + tot_b + sync_b + res_b + mail_b) + 255) & ~size_t(255);
    const size_t vr_b = ((size_t)W * sizeof(VirtualRank) + 255) & ~size_t(255);
    CK(h->d_emu.ensure(vr_b + per_rank * W));
    CK(cudaMemsetAsync(h->d_emu.p, 0, vr_b + per_rank * W, h->stream));
    std::vector<VirtualRank> vr(W);
    const long long groups = (static_cast<long long>(h->n_source) + 31) / 32;
    for (int r = 0; r < W; ++r) {
      char* base = h->d_emu.as<char>() + vr_b + per_rank * r;
      const long long g_lo = r * (groups / W) + std::min<long long>(r, groups % W);
      const long long g_hi = g_lo + groups / W + (r < groups % W ? 1 : 0);
      const long long lo = std::min<long long>(h->n_source, g_lo * 32), hi = std::min<long long>(h->n_source, g_hi * 32);
      vr[r].src = d_src + lo;
      vr[r].n_source = static_cast<int32_t>(hi - lo);
      vr[r].pad = 0;
      vr[r].partials = reinterpret_cast<double*>(base);
      vr[r].totals = reinterpret_cast<double*>(base + part_b);
      vr[r].sync = reinterpret_cast<unsigned int*>(base + part_b + tot_b);
      vr[r].result = reinterpret_cast<AlignResultDev*>(base + part_b + tot_b + sync_b);
      ws.mail[r] = reinterpret_cast<unsigned long long*>(base + part_b + tot_b + sync_b + res_b);
    }
    CK(cudaMemcpyAsync(h->d_emu.p, vr.data(), W * sizeof(VirtualRank), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));  // vr is a stack-lifetime staging buffer
    ws.world = W;
    ws.rank = 0;
    ws.vranks = W;
    ws.vr = h->d_emu.as<VirtualRank>();
    ws.n_source_total = static_cast<long long>(h->n_source);
    h->emu_world = W;
    h->emu_per_rank = per_rank;
    h->emu_result_off = vr_b + part_b + tot_b + sync_b;
  }
  MapView map = make_view(h);
  const float4* src = d_src;
  void* args[] = {(void*)&src, (void*)&map, (void*)&prm, (void*)&ws};
  const void* fn;
  if (shape == 0)
    fn = method == NDTB200_DIRECT1 ? (const void*)ndt_align_kernel<3, kThreadsLatency>
       : method == NDTB200_DIRECT7 ? (const void*)ndt_align_kernel<2, kThreadsLatency>
       : method == NDTB200_DIRECT26 ? (const void*)ndt_align_kernel<1, kThreadsLatency>
                                    : (const void*)ndt_align_kernel<0, kThreadsLatency>;
  else
    fn = method == NDTB200_DIRECT1 ? (const void*)ndt_align_kernel<3, kThreadsThroughput>
       : method == NDTB200_DIRECT7 ? (const void*)ndt_align_kernel<2, kThreadsThroughput>
       : method == NDTB200_DIRECT26 ? (const void*)ndt_align_kernel<1, kThreadsThroughput>
                                    : (const void*)ndt_align_kernel<0, kThreadsThroughput>;
  const int threads = shape == 0 ? kThreadsLatency : kThreadsThroughput;
  h->last_blocks = blocks;
  CK(cudaEventRecord(h->ev0, h->stream));
  const size_t smem = 0;
  static const bool plain_launch = getenv("NDTB200_PLAIN_LAUNCH") != nullptr;  // experiment only: no co-residency guarantee
  if (plain_launch) CK(cudaLaunchKernel(fn, dim3(blocks), dim3(threads), args, smem, h->stream));
  else CK(cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(threads), args, smem, h->stream));
  LAUNCHED(h);
  CK(cudaEventRecord(h->ev1, h->stream));
  h->result_valid = false;
  h->result_copy_enqueued = false;
  return NDTB200_OK;
}

int enqueue_result_copy(ndtb200_handle* h) {
  CK(cudaMemcpyAsync(h->h_result, h->d_result.p, sizeof(AlignResultDev), cudaMemcpyDeviceToHost, h->stream));
  h->result_copy_enqueued = true;
  return NDTB200_OK;
}

int fetch_result(ndtb200_handle* h) {
  if (!h->result_copy_enqueued) {
    int st = enqueue_result_copy(h);
    if (st != NDTB200_OK) return st;
  }
  if (h->shape == 1) {
    // throughput mode: many handles, each driven by its own host thread (the c3 pipeline): poll and YIELD instead of
    // spinning inside cudaStreamSynchronize, so that more pipeline threads than host cores still make progress
    while (true) {
      const cudaError_t q = cudaStreamQuery(h->stream);
      if (q == cudaSuccess) break;
      if (q != cudaErrorNotReady) { h->err = std::string("cudaStreamQuery: ") + cudaGetErrorString(q); return NDTB200_ERR_CUDA; }
      sched_yield();
    }
  }
  CK(cudaStreamSynchronize(h->stream));
  h->result_copy_enqueued = false;
  h->result_valid = true;
  return NDTB200_OK;
}

// parity dump of a map that was installed from finished records (the owner-partitioned sharded build): keys, counts,
// means and fp64 inverse covariances come from the records; covariances / inflation flags are not part of a record
int dump_records_only(ndtb200_handle* h, int32_t* keys, int32_t* counts, double* mean, double* cov, double* icov, int32_t* inflated) {
  const size_t V = static_cast<size_t>(h->n_voxels);
  std::vector<VoxelRecord> recs(V);
  std::vector<double> ic(V * 6);
  CK(cudaMemcpyAsync(recs.data(), h->d_records.p, V * sizeof(VoxelRecord), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(ic.data(), h->d_icov64.p, V * 6 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (size_t v = 0; v < V; ++v) {
    if (keys) keys[v] = recs[v].key;
    if (counts) counts[v] = recs[v].count;
    if (mean) for (int a = 0; a < 3; ++a) mean[v * 3 + a] = record_mean(recs[v], a);
    if (icov) {
      const double* c = &ic[v * 6];
      const double full[9] = {c[0], c[1], c[2], c[1], c[3], c[4], c[2], c[4], c[5]};
      for (int k = 0; k < 9; ++k) icov[v * 9 + k] = full[k];
    }
    if (cov) for (int k = 0; k < 9; ++k) cov[v * 9 + k] = std::nan("");
    if (inflated) inflated[v] = -1;
  }
  return NDTB200_OK;
}

void colmajor_to_T(const float* m, float T[12]) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) T[r * 4 + c] = m[c * 4 + r];
}
void T_to_colmajor(const float T[12], float* m) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) m[c * 4 + r] = T[r * 4 + c];
  m[3] = m[7] = m[11] = 0.0f;
  m[15] = 1.0f;
}
void fill_result(const AlignResultDev& r, ndtb200_result* out) {
  T_to_colmajor(r.final_T, out->final_transformation);
  float incr[12];
  pose_to_matrix(r.last_dp, incr);  // transformation_ (ndt_omp_impl.hpp:146-149); Identity if no step was taken
  T_to_colmajor(incr, out->last_increment);
  out->converged = r.converged;
  out->iterations = r.iterations;
  out->trans_probability = r.trans_probability;
  for (int i = 0; i < 6; ++i) out->final_pose[i] = r.final_pose[i];
  out->final_score = r.final_score;
  out->n_evaluations = r.n_evals;
  out->n_hessian_passes = r.n_hess;
  out->n_hits = r.n_hits;
}

// p = [translation, rotation().eulerAngles(0,1,2)] of the guess (ndt_omp_impl.hpp:103-111; Identity gives -0,0,-0)
void guess_to_pose(const float T0[12], double p0[6]) {
  float R[3][3], ang[3];
  rotation_polar_host(T0, R);
  euler_angles_012_host(R, ang);
  p0[0] = T0[3]; p0[1] = T0[7]; p0[2] = T0[11];
  p0[3] = ang[0]; p0[4] = ang[1]; p0[5] = ang[2];
}

bool is_identity_colmajor(const float* m) {
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c)
      if (m[c * 4 + r] != ((r == c) ? 1.0f : 0.0f)) return false;
  return true;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int ndtb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int ndtb200_default_params(ndtb200_params* p) {
  if (!p) return NDTB200_ERR_INVALID;
  p->resolution = 1.0f;
  p->step_size = 0.1;
  p->outlier_ratio = 0.55;
  p->trans_eps = 0.1;
  p->max_iterations = 35;
  p->search_method = NDTB200_DIRECT7;
  p->min_points_per_voxel = 6;
  p->eig_ratio = 0.01;
  return NDTB200_OK;
}

int ndtb200_create(ndtb200_handle** out, int device) {
  if (!out) return NDTB200_ERR_INVALID;
  *out = nullptr;
  // Independent handles run on their own streams; streams are multiplexed onto CUDA_DEVICE_MAX_CONNECTIONS hardware
  // queues, and kernels of streams that share a queue serialise.  Ask for the maximum (32) unless the caller chose a
  // value; this only takes effect if it happens before the process creates its CUDA context (measured on c2, 64 solves
  // in flight: 13.7 k -> 15.4 k aligns/s).
  setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev)
    return NDTB200_ERR_NO_DEVICE;  // no CPU fallback, by design
  ndtb200_handle* h = new ndtb200_handle();
  h->device = device;
  ndtb200_default_params(&h->prm);
  auto fail = [&](int code) { delete h; return code; };
  if (cudaSetDevice(device) != cudaSuccess) return fail(NDTB200_ERR_NO_DEVICE);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(NDTB200_ERR_NO_DEVICE);
  h->num_sms = prop.multiProcessorCount;
  if (!prop.cooperativeLaunch) return fail(NDTB200_ERR_NO_DEVICE);
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(NDTB200_ERR_CUDA);
  if (cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess) return fail(NDTB200_ERR_CUDA);
  if (cudaMallocHost(&h->h_result, sizeof(AlignResultDev)) != cudaSuccess) return fail(NDTB200_ERR_CUDA);
  std::memset(h->h_result, 0, sizeof(AlignResultDev));
  int st = query_coop_blocks<0>(h);
  if (st == NDTB200_OK) st = query_coop_blocks<1>(h);
  if (st == NDTB200_OK) st = query_coop_blocks<2>(h);
  if (st == NDTB200_OK) st = query_coop_blocks<3>(h);
  if (st == NDTB200_OK) st = ensure_empty_hash(h);
  if (st == NDTB200_OK && h->d_scalar.ensure(256) != cudaSuccess) st = NDTB200_ERR_CUDA;
  if (st != NDTB200_OK) {
    std::fprintf(stderr, "ndtb200_create: %s\n", h->err.c_str());
    return fail(st);
  }
  *out = h;
  return NDTB200_OK;
}

int ndtb200_destroy(ndtb200_handle* h) {
  if (!h) return NDTB200_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  ndtb200_comm_detach(h);
  if (h->aux) { ndtb200_destroy(h->aux); h->aux = nullptr; }
  DevBuf* bufs[] = {&h->d_target, &h->d_grid, &h->d_mm_partial, &h->d_mm_finite, &h->d_keys_a, &h->d_keys_b,
                    &h->d_vals_a, &h->d_vals_b, &h->d_pay_a, &h->d_pay_b, &h->d_hist, &h->d_scan_tmp, &h->d_scalar, &h->d_sortmeta, &h->d_status, &h->d_voxel_key,
                    &h->d_voxel_start, &h->d_voxel_count, &h->d_moments, &h->d_records, &h->d_icov64, &h->d_hash, &h->d_dense, &h->d_source,
                    &h->d_cell_all, &h->d_best, &h->d_centroid, &h->d_partials, &h->d_totals, &h->d_sync, &h->d_result, &h->d_trace, &h->d_out, &h->d_tmp, &h->d_mail, &h->d_emu, &h->d_source_sorted};
  for (DevBuf* b : bufs) b->release();
  if (h->h_result) cudaFreeHost(h->h_result);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return NDTB200_OK;
}

const char* ndtb200_last_error(const ndtb200_handle* h) { return h ? h->err.c_str() : "null handle"; }

int ndtb200_set_params(ndtb200_handle* h, const ndtb200_params* p) {
  if (!h || !p) return NDTB200_ERR_INVALID;
  if (!(p->resolution > 0) || p->search_method < 0 || p->search_method > 3) {
    h->err = "invalid parameters";
    return NDTB200_ERR_INVALID;
  }
  const bool res_changed = (p->resolution != h->prm.resolution);
  h->prm = *p;
  if (h->prm.min_points_per_voxel < 3) h->prm.min_points_per_voxel = 3;  // setMinPointPerVoxel, vgc.h:228-240
  // setResolution (ndt_omp.h:132-142): rebuild only if the value changed AND a source is set (sic)
  if (res_changed && h->has_source && h->has_target) {
    cudaSetDevice(h->device);
    int st = build_map(h);
    if (st == NDTB200_ERR_CUDA) return st;
  }
  return NDTB200_OK;
}

int ndtb200_get_params(const ndtb200_handle* h, ndtb200_params* p) {
  if (!h || !p) return NDTB200_ERR_INVALID;
  *p = h->prm;
  return NDTB200_OK;
}

int ndtb200_set_target(ndtb200_handle* h, const void* points, size_t n, size_t stride_bytes, int is_dense) {
  if (!h || (!points && n)) return NDTB200_ERR_INVALID;
  cudaSetDevice(h->device);
  h->target_view = nullptr;
  int st = upload_points(h, h->d_target, points, n, stride_bytes);
  if (st != NDTB200_OK) return st;
  h->n_target = n;
  h->target_dense = is_dense != 0;
  h->has_target = true;
  return build_map(h);
}

// setInputTarget for a cloud that already lives on this device and STAYS there: the reference keeps the caller's cloud by
// shared pointer (pcl::Registration::target_) and never copies it; here the handle keeps the caller's device pointer.
// The buffer must stay alive and unchanged until the next set_target call or the handle's destruction (getFitnessScore and
// the KDTREE centroids read the raw target later).
int ndtb200_set_target_device_view(ndtb200_handle* h, const void* d_points, size_t n, int is_dense) {
  if (!h || (!d_points && n)) return NDTB200_ERR_INVALID;
  cudaSetDevice(h->device);
  h->target_view = n ? static_cast<const float4*>(d_points) : nullptr;
  h->n_target = n;
  h->target_dense = is_dense != 0;
  h->has_target = true;
  return build_map(h);
}

int ndtb200_set_target_device(ndtb200_handle* h, const void* d_points, size_t n, int is_dense) {
  if (!h || (!d_points && n)) return NDTB200_ERR_INVALID;
  cudaSetDevice(h->device);
  h->target_view = nullptr;
  if (n) {
    CK(h->d_target.ensure(n * sizeof(float4)));
    CK(cudaMemcpyAsync(h->d_target.p, d_points, n * sizeof(float4), cudaMemcpyDeviceToDevice, h->stream));
  }
  h->n_target = n;
  h->target_dense = is_dense != 0;
  h->has_target = true;
  return build_map(h);
}

int ndtb200_set_source(ndtb200_handle* h, const void* points, size_t n, size_t stride_bytes) {
  if (!h || (!points && n)) return NDTB200_ERR_INVALID;
  if (n > 0x7FFFFFF0ull) { h->err = "source cloud too large"; return NDTB200_ERR_INVALID; }
  cudaSetDevice(h->device);
  int st = upload_points(h, h->d_source, points, n, stride_bytes);
  if (st != NDTB200_OK) return st;
  h->n_source = n;
  h->has_source = true;
  h->source_sorted_valid = h->source_sort_failed = false;
  return NDTB200_OK;
}

int ndtb200_set_source_device(ndtb200_handle* h, const void* d_points, size_t n) {
  if (!h || (!d_points && n)) return NDTB200_ERR_INVALID;
  if (n > 0x7FFFFFF0ull) { h->err = "source cloud too large"; return NDTB200_ERR_INVALID; }
  cudaSetDevice(h->device);
  if (n) {
    CK(h->d_source.ensure(n * sizeof(float4)));
    CK(cudaMemcpyAsync(h->d_source.p, d_points, n * sizeof(float4), cudaMemcpyDeviceToDevice, h->stream));
  }
  h->n_source = n;
  h->has_source = true;
  h->source_sorted_valid = h->source_sort_failed = false;
  return NDTB200_OK;
}

int ndtb200_align_async(ndtb200_handle* h, const float* guess) {
  if (!h) return NDTB200_ERR_INVALID;
  cudaSetDevice(h->device);
  float T0[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  double p0[6] = {0, 0, 0, 0, 0, 0};
  int has_guess = 0;
  if (guess && !is_identity_colmajor(guess)) {
    has_guess = 1;
    colmajor_to_T(guess, T0);
  }
  guess_to_pose(T0, p0);
  return launch_align(h, MODE_ALIGN, p0, T0, has_guess, 1);
}

int ndtb200_sync(ndtb200_handle* h) {
  if (!h) return NDTB200_ERR_INVALID;
  cudaSetDevice(h->device);
  int st = fetch_result(h);
  if (st != NDTB200_OK) return st;
  std::memcpy(h->last_final_T, h->h_result->final_T, sizeof(h->last_final_T));
  if (h->h_result->aborted) { h->err = "sharded solve aborted: a peer rank did not answer within 10 s"; return NDTB200_ERR_CUDA; }
  return NDTB200_OK;
}

// align()'s output cloud (pcl::Registration::align: source transformed by the final pose), enqueued behind the solve
static int enqueue_output(ndtb200_handle* h, void* out_points, size_t out_stride_bytes) {
  if (!out_points) return NDTB200_OK;
  if (out_stride_bytes < 16) { h->err = "out_stride_bytes must be >= 16"; return NDTB200_ERR_INVALID; }
  const int n = static_cast<int>(h->n_source);
  if (n == 0) return NDTB200_OK;
  if (!h->out_fused) {  // the solve was launched without knowing an output would be asked for (ndtb200_align_async)
    CK(h->d_out.ensure((size_t)n * sizeof(float4)));
    transform_output_kernel<<<grid_for(n, 256, h->num_sms * 8), 256, 0, h->stream>>>(
        h->d_source.as<float4>(), n, h->d_result.as<AlignResultDev>(), h->d_out.as<float4>());
    LAUNCHED(h);
  }
  if (out_stride_bytes == 16)
    CK(cudaMemcpyAsync(out_points, h->d_out.p, (size_t)n * 16, cudaMemcpyDeviceToHost, h->stream));
  else
    CK(cudaMemcpy2DAsync(out_points, out_stride_bytes, h->d_out.p, 16, 16, n, cudaMemcpyDeviceToHost, h->stream));
  return NDTB200_OK;
}

int ndtb200_align(ndtb200_handle* h, const float* guess, void* out_points, size_t out_stride_bytes) {
  if (!h) return NDTB200_ERR_INVALID;
  h->want_fused_out = out_points != nullptr && out_stride_bytes >= 16;
  int st = ndtb200_align_async(h, guess);
  h->want_fused_out = false;
  if (st != NDTB200_OK) return st;
  st = enqueue_output(h, out_points, out_stride_bytes);
  if (st != NDTB200_OK) return st;
  return ndtb200_sync(h);
}

int ndtb200_get_result(ndtb200_handle* h, ndtb200_result* out) {
  if (!h || !out) return NDTB200_ERR_INVALID;
  if (!h->result_valid && h->d_result.p == nullptr) {  // nothing aligned yet: pcl::Registration's initial state
    std::memset(out, 0, sizeof(*out));
    out->final_transformation[0] = out->final_transformation[5] = out->final_transformation[10] = out->final_transformation[15] = 1.0f;
    out->last_increment[0] = out->last_increment[5] = out->last_increment[10] = out->last_increment[15] = 1.0f;
    return NDTB200_OK;
  }
  if (!h->result_valid) {
    int st = ndtb200_sync(h);
    if (st != NDTB200_OK) return st;
  }
  fill_result(*h->h_result, out);
  return NDTB200_OK;
}

// A source-sharded solve with `world` ranks emulated inside ONE cooperative launch on this handle's GPU: every rank
// works on the contiguous source range it would own on its own GPU and the ranks exchange their 29 per-evaluation sums
// through the same tagged mailbox stores / polls as the multi-GPU path (ndtb200_comm_*).  results[r] is rank r's own
// result block: all ranks must report identical bits.  For boxes with fewer GPUs than ranks (tests, bring-up).
int ndtb200_align_emulated_ranks(ndtb200_handle* h, int world, const float* guess, ndtb200_result* results) {
  if (!h || !results || world < 2 || world > kMaxRanks) return NDTB200_ERR_INVALID;
  if (h->comm_world > 1) { h->err = "handle is attached to a real communicator"; return NDTB200_ERR_INVALID; }
  cudaSetDevice(h->device);
  float T0[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  double p0[6];
  int has_guess = 0;
  if (guess && !is_identity_colmajor(guess)) { has_guess = 1; colmajor_to_T(guess, T0); }
  guess_to_pose(T0, p0);
  int st = launch_align(h, MODE_ALIGN, p0, T0, has_guess, 1, world);
  if (st != NDTB200_OK) return st;
  std::vector<AlignResultDev> rr(world);
  for (int r = 0; r < world; ++r)
    CK(cudaMemcpyAsync(&rr[r], h->d_emu.as<char>() + h->emu_result_off + h->emu_per_rank * r, sizeof(AlignResultDev),
                       cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  int rc = NDTB200_OK;
  for (int r = 0; r < world; ++r) {
    fill_result(rr[r], results + r);
    if (rr[r].aborted) { h->err = "emulated sharded solve aborted"; rc = NDTB200_ERR_CUDA; }
  }
  std::memcpy(h->h_result, &rr[0], sizeof(AlignResultDev));
  std::memcpy(h->last_final_T, rr[0].final_T, sizeof(h->last_final_T));
  h->result_valid = true;
  return rc;
}

static int fitness_sums(ndtb200_handle* h, double max_range, double* sum_out, unsigned long long* count_out);

int ndtb200_fitness_score(ndtb200_handle* h, double max_range, double* out) {
  if (!h || !out) return NDTB200_ERR_INVALID;
  double s = 0;
  unsigned long long c = 0;
  const int st = fitness_sums(h, max_range, &s, &c);
  if (st != NDTB200_OK) return st;
  *out = c > 0 ? s / static_cast<double>(c) : 1.7976931348623157e308;
  return NDTB200_OK;
}

// The two sums behind getFitnessScore: with the source sharded over several GPUs (every rank holds the full raw target
// and its source range) the ranks all-reduce {sum of squared distances, accepted count} and divide.
int ndtb200_fitness_sums(ndtb200_handle* h, double max_range, double* sum_sq_dist, int64_t* n_accepted) {
  if (!h || !sum_sq_dist || !n_accepted) return NDTB200_ERR_INVALID;
  *sum_sq_dist = 0;
  *n_accepted = 0;
  if (h->has_source && h->has_target && h->n_source == 0 && h->n_target > 0) return NDTB200_OK;  // an empty source slice
  unsigned long long c = 0;
  const int st = fitness_sums(h, max_range, sum_sq_dist, &c);
  *n_accepted = static_cast<int64_t>(c);
  return st;
}

static int fitness_sums(ndtb200_handle* h, double max_range, double* sum_out, unsigned long long* count_out) {
  if (!h->has_source || !h->has_target || h->n_source == 0 || h->n_target == 0) {
    h->err = "fitness needs a source and a target";
    return NDTB200_ERR_NO_INPUT;
  }
  if (h->map_is_merged) {
    h->err = "getFitnessScore is not available on a map merged from sharded partials (this rank holds only its slice of the raw target)";
    return NDTB200_ERR_INVALID;
  }
  cudaSetDevice(h->device);
  const int n = static_cast<int>(h->n_source);
  const int blocks = (n + 255) / 256;
  CK(h->d_tmp.ensure((size_t)blocks * 16 + 64));
  double* d_sum = h->d_tmp.as<double>();
  unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(d_sum + blocks);
  float* d_T = reinterpret_cast<float*>(h->d_scalar.as<char>() + 64);
  CK(cudaMemcpyAsync(d_T, h->last_final_T, 12 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  // grid-accelerated exact search when the map's sorted point ranges are available and the cell table fits
  const unsigned long long ncell = static_cast<unsigned long long>(h->grid.div_b[0]) * static_cast<unsigned long long>(h->grid.div_b[1]) *
                                   static_cast<unsigned long long>(h->grid.div_b[2]);
  const bool force_brute = getenv("NDTB200_FITNESS_BRUTE") != nullptr;  // tests: compare with the O(N*M) scan
  const bool use_grid = !force_brute && h->map_status == NDTB200_OK && h->n_voxels > 0 && ncell > 0 && ncell * sizeof(int32_t) <= (2ull << 30) &&
                        h->n_target <= 0x7fffffffull;
  if (use_grid) {
    {
      const int st = ensure_sorted_indices(h);  // payload-sorted map: the index array is built on first use
      if (st != NDTB200_OK) return st;
    }
    {
      const int st = ensure_cell_table(h);
      if (st != NDTB200_OK) return st;
    }
    CK(h->d_best.ensure((size_t)n * sizeof(float) + (size_t)n * sizeof(int) + 64));
    float* d_best = h->d_best.as<float>();
    int* d_list = reinterpret_cast<int*>(d_best + n);
    int* d_count = reinterpret_cast<int*>(h->d_scalar.as<char>() + 128);
    CK(cudaMemsetAsync(d_count, 0, sizeof(int), h->stream));
    const int warp_blocks = static_cast<int>(((size_t)n * 32 + 255) / 256);  // one warp per query
    fitness_grid_kernel<<<warp_blocks, 256, 0, h->stream>>>(h->d_source.as<float4>(), n, target_pts(h), h->d_vals_a.as<uint32_t>(),
                                                        h->d_voxel_start.as<uint32_t>(), static_cast<uint32_t>(h->n_voxels),
                                                        static_cast<uint32_t>(h->grid.n_finite), h->d_cell_all.as<int32_t>(),
                                                        h->d_grid.as<GridDesc>(), d_T, d_best, d_list, d_count);
    LAUNCHED(h);
    fitness_fallback_kernel<<<std::min(n, h->num_sms * 8), 256, 0, h->stream>>>(h->d_source.as<float4>(), d_list, d_count, target_pts(h),
                                                            static_cast<int>(h->n_target), d_T, d_best);
    LAUNCHED(h);
    fitness_reduce_kernel<<<blocks, 256, 0, h->stream>>>(d_best, n, max_range, d_sum, d_cnt);
    LAUNCHED(h);
    if (getenv("NDTB200_DEBUG_STEP")) {
      int nf = 0;
      CK(cudaMemcpyAsync(&nf, d_count, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      std::fprintf(stderr, "  fitness: %d of %d queries went to the brute-force fallback\n", nf, n);
    }
  } else {
    fitness_bruteforce_kernel<<<blocks, 256, 0, h->stream>>>(h->d_source.as<float4>(), n, target_pts(h),
                                                             static_cast<int>(h->n_target), d_T, max_range, d_sum, d_cnt);
    LAUNCHED(h);
  }
  std::vector<double> hs(blocks);
  std::vector<unsigned long long> hc(blocks);
  CK(cudaMemcpyAsync(hs.data(), d_sum, blocks * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(hc.data(), d_cnt, blocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  double s = 0;
  unsigned long long c = 0;
  for (int b = 0; b < blocks; ++b) { s += hs[b]; c += hc[b]; }
  *sum_out = s;
  *count_out = c;
  return NDTB200_OK;
}

// shared by the three calculateScore entry points: cloud (float4, device) + segments or poses -> out[n_seg] (means)
static int calculate_scores(ndtb200_handle* h, const float4* d_cloud, const unsigned long long* h_seg_start, const float* h_poses12,
                            int n_seg, size_t n_points_per_pose, double* out) {
  if (h->prm.search_method == NDTB200_KDTREE) {
    const int st2 = ensure_kdtree_index(h);
    if (st2 != NDTB200_OK) return st2;
  }
  size_t longest = n_points_per_pose;
  if (h_seg_start)
    for (int s = 0; s < n_seg; ++s) longest = std::max<size_t>(longest, h_seg_start[s + 1] - h_seg_start[s]);
  const int chunks = static_cast<int>(std::max<size_t>(1, std::min<size_t>((longest + 255) / 256, std::max(1, 8 * h->num_sms / std::max(1, n_seg)))));
  CK(h->d_tmp.ensure((size_t)n_seg * chunks * sizeof(double) + (size_t)(n_seg + 1) * sizeof(unsigned long long) + (size_t)n_seg * 12 * sizeof(float) + 256));
  double* d_sum = h->d_tmp.as<double>();
  unsigned long long* d_seg = reinterpret_cast<unsigned long long*>(d_sum + (size_t)n_seg * chunks);
  float* d_pose = reinterpret_cast<float*>(d_seg + n_seg + 1);
  if (h_seg_start) CK(cudaMemcpyAsync(d_seg, h_seg_start, (size_t)(n_seg + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, h->stream));
  if (h_poses12) CK(cudaMemcpyAsync(d_pose, h_poses12, (size_t)n_seg * 12 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  double d1, d2, d3;
  gauss_constants(h->prm, d1, d2, d3);
  MapView map = make_view(h);
  const dim3 grid(chunks, n_seg);
  const float* poses = h_poses12 ? d_pose : nullptr;
  const int npp = static_cast<int>(n_points_per_pose);
  switch (h->prm.search_method) {
    case NDTB200_DIRECT1: calculate_score_kernel<3><<<grid, 256, 0, h->stream>>>(d_cloud, d_seg, poses, npp, map, d1, d2, d3, chunks, d_sum); break;
    case NDTB200_DIRECT7: calculate_score_kernel<2><<<grid, 256, 0, h->stream>>>(d_cloud, d_seg, poses, npp, map, d1, d2, d3, chunks, d_sum); break;
    case NDTB200_DIRECT26: calculate_score_kernel<1><<<grid, 256, 0, h->stream>>>(d_cloud, d_seg, poses, npp, map, d1, d2, d3, chunks, d_sum); break;
    default: calculate_score_kernel<0><<<grid, 256, 0, h->stream>>>(d_cloud, d_seg, poses, npp, map, d1, d2, d3, chunks, d_sum); break;
  }
  LAUNCHED(h);
  std::vector<double> hs((size_t)n_seg * chunks);
  CK(cudaMemcpyAsync(hs.data(), d_sum, hs.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int s = 0; s < n_seg; ++s) {
    double acc = 0;
    for (int c = 0; c < chunks; ++c) acc += hs[(size_t)s * chunks + c];
    const size_t cnt = h_seg_start ? static_cast<size_t>(h_seg_start[s + 1] - h_seg_start[s]) : n_points_per_pose;
    out[s] = cnt ? acc / static_cast<double>(cnt) : 0.0;
  }
  return NDTB200_OK;
}

int ndtb200_calculate_score(ndtb200_handle* h, const void* points, size_t n, size_t stride_bytes, double* out) {
  if (!h || !out || (!points && n)) return NDTB200_ERR_INVALID;
  if (n == 0) { *out = 0; return NDTB200_ERR_NO_INPUT; }
  cudaSetDevice(h->device);
  int st = upload_points(h, h->d_out, points, n, stride_bytes);
  if (st != NDTB200_OK) return st;
  const unsigned long long seg[2] = {0ull, static_cast<unsigned long long>(n)};
  return calculate_scores(h, h->d_out.as<float4>(), seg, nullptr, 1, 0, out);
}

int ndtb200_calculate_score_batch(ndtb200_handle* h, const void* points, const size_t* cloud_offsets, int n_clouds, size_t stride_bytes,
                                  double* out) {
  if (!h || !out || !cloud_offsets || n_clouds < 0 || (!points && n_clouds && cloud_offsets[n_clouds])) return NDTB200_ERR_INVALID;
  if (n_clouds == 0) return NDTB200_OK;
  const size_t n = cloud_offsets[n_clouds];
  std::vector<unsigned long long> seg(n_clouds + 1);
  for (int s = 0; s <= n_clouds; ++s) {
    if (s && cloud_offsets[s] < cloud_offsets[s - 1]) { h->err = "cloud_offsets must be non-decreasing"; return NDTB200_ERR_INVALID; }
    seg[s] = cloud_offsets[s];
  }
  cudaSetDevice(h->device);
  if (n == 0) { for (int s = 0; s < n_clouds; ++s) out[s] = 0.0; return NDTB200_OK; }
  int st = upload_points(h, h->d_out, points, n, stride_bytes);
  if (st != NDTB200_OK) return st;
  return calculate_scores(h, h->d_out.as<float4>(), seg.data(), nullptr, n_clouds, 0, out);
}

int ndtb200_score_poses(ndtb200_handle* h, const float* poses16, int n_poses, double* out) {
  if (!h || !out || (!poses16 && n_poses) || n_poses < 0) return NDTB200_ERR_INVALID;
  if (n_poses == 0) return NDTB200_OK;
  if (!h->has_source || h->n_source == 0) { h->err = "no input source"; return NDTB200_ERR_NO_INPUT; }
  cudaSetDevice(h->device);
  std::vector<float> T((size_t)n_poses * 12);
  for (int s = 0; s < n_poses; ++s) colmajor_to_T(poses16 + (size_t)s * 16, T.data() + (size_t)s * 12);
  return calculate_scores(h, h->d_source.as<float4>(), nullptr, T.data(), n_poses, h->n_source, out);
}

int ndtb200_get_map_info(const ndtb200_handle* h, ndtb200_map_info* out) {
  if (!h || !out) return NDTB200_ERR_INVALID;
  for (int a = 0; a < 3; ++a) {
    out->min_b[a] = h->grid.min_b[a];
    out->max_b[a] = h->grid.max_b[a];
    out->div_b[a] = h->grid.div_b[a];
  }
  out->n_points = static_cast<int64_t>(h->n_target);
  out->n_voxels = h->n_voxels;
  if (h->n_valid < 0 && h->n_voxels > 0) {  // counted on the device during the build, fetched on first request
    ndtb200_handle* hm = const_cast<ndtb200_handle*>(h);
    cudaSetDevice(h->device);
    unsigned int nv = 0;
    if (cudaMemcpyAsync(&nv, hm->d_scalar.as<unsigned int>() + 4, sizeof(unsigned int), cudaMemcpyDeviceToHost, hm->stream) != cudaSuccess ||
        cudaStreamSynchronize(hm->stream) != cudaSuccess)
      return NDTB200_ERR_CUDA;
    hm->n_valid = nv;
  }
  out->n_valid = h->n_valid < 0 ? 0 : h->n_valid;
  out->hash_capacity = h->hash_cap;
  return h->map_status;
}

int ndtb200_dump_point_keys(ndtb200_handle* h, int32_t* keys) {
  if (!h || !keys) return NDTB200_ERR_INVALID;
  if (h->map_status != NDTB200_OK) return h->map_status;
  cudaSetDevice(h->device);
  const size_t n = h->n_target;
  CK(h->d_tmp.ensure(n * sizeof(uint32_t)));
  voxel_key_kernel<<<grid_for(n, kBuildThreads * 4, h->num_sms * 16), kBuildThreads, 0, h->stream>>>(
      target_pts(h), n, h->target_dense ? 1 : 0, h->d_grid.as<GridDesc>(), 0xFFFFFFFFu,
      h->d_tmp.as<uint32_t>(), nullptr, nullptr, 0);
  LAUNCHED(h);
  CK(cudaMemcpyAsync(keys, h->d_tmp.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return NDTB200_OK;
}

int ndtb200_dump_voxels(ndtb200_handle* h, int32_t* keys, int32_t* counts, double* mean, double* cov, double* icov,
                        int32_t* inflated) {
  if (!h) return NDTB200_ERR_INVALID;
  if (h->map_status != NDTB200_OK) return h->map_status;
  cudaSetDevice(h->device);
  const uint32_t V = static_cast<uint32_t>(h->n_voxels);
  if (V == 0) return NDTB200_OK;
  if (h->records_only_map) return dump_records_only(h, keys, counts, mean, cov, icov, inflated);
  if (!h->moments_valid) {
    const int st = compute_moments(h, V, static_cast<uint32_t>(h->grid.n_finite));
    if (st != NDTB200_OK) return st;
  }
  DevBuf d_mean, d_cov, d_icov, d_infl, d_rec, d_ic64;
  CK(d_mean.ensure((size_t)V * 3 * sizeof(double)));
  CK(d_cov.ensure((size_t)V * 9 * sizeof(double)));
  CK(d_icov.ensure((size_t)V * 9 * sizeof(double)));
  CK(d_infl.ensure((size_t)V * sizeof(int)));
  CK(d_rec.ensure((size_t)V * sizeof(VoxelRecord)));
  CK(d_ic64.ensure((size_t)V * 6 * sizeof(double)));
  unsigned int* d_nvalid = h->d_scalar.as<unsigned int>() + 8;
  const int vblocks = static_cast<int>(((size_t)V + kBuildThreads - 1) / kBuildThreads);
  finalize_voxels_kernel<<<vblocks, kBuildThreads, 0, h->stream>>>(
      h->d_moments.as<double>(), h->d_voxel_key.as<int32_t>(), h->d_voxel_start.as<uint32_t>(),
      h->map_is_merged ? h->d_voxel_count.as<uint32_t>() : nullptr, V,
      static_cast<uint32_t>(h->grid.n_finite), h->prm.min_points_per_voxel, h->prm.eig_ratio, d_rec.as<VoxelRecord>(),
      d_ic64.as<double>(), d_nvalid, d_mean.as<double>(), d_cov.as<double>(), d_icov.as<double>(), d_infl.as<int>());
  LAUNCHED(h);
  std::vector<VoxelRecord> recs(V);
  CK(cudaMemcpyAsync(recs.data(), d_rec.p, (size_t)V * sizeof(VoxelRecord), cudaMemcpyDeviceToHost, h->stream));
  if (mean) CK(cudaMemcpyAsync(mean, d_mean.p, (size_t)V * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (cov) CK(cudaMemcpyAsync(cov, d_cov.p, (size_t)V * 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (icov) CK(cudaMemcpyAsync(icov, d_icov.p, (size_t)V * 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (inflated) CK(cudaMemcpyAsync(inflated, d_infl.p, (size_t)V * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (uint32_t v = 0; v < V; ++v) {
    if (keys) keys[v] = recs[v].key;
    if (counts) counts[v] = recs[v].count;
  }
  d_mean.release(); d_cov.release(); d_icov.release(); d_infl.release(); d_rec.release(); d_ic64.release();
  return NDTB200_OK;
}

int ndtb200_eval_derivatives(ndtb200_handle* h, const double p[6], const float* T, int compute_hessian,
                             double out43[43], int64_t* n_hits) {
  if (!h || !p || !out43) return NDTB200_ERR_INVALID;
  cudaSetDevice(h->device);
  float T0[12];
  if (T) colmajor_to_T(T, T0); else pose_to_matrix(p, T0);
  int st = launch_align(h, MODE_EVAL, p, T0, 0, compute_hessian);
  if (st != NDTB200_OK) return st;
  st = fetch_result(h);
  if (st != NDTB200_OK) return st;
  for (int i = 0; i < 43; ++i) out43[i] = h->h_result->totals[i];
  if (n_hits) *n_hits = h->h_result->n_hits;
  h->result_valid = false;
  return NDTB200_OK;
}

int ndtb200_eval_hessian(ndtb200_handle* h, const double p[6], const float* T, double out36[36]) {
  if (!h || !p || !out36) return NDTB200_ERR_INVALID;
  cudaSetDevice(h->device);
  float T0[12];
  if (T) colmajor_to_T(T, T0); else pose_to_matrix(p, T0);
  int st = launch_align(h, MODE_HESSIAN, p, T0, 0, 1);
  if (st != NDTB200_OK) return st;
  st = fetch_result(h);
  if (st != NDTB200_OK) return st;
  for (int i = 0; i < 36; ++i) out36[i] = h->h_result->totals[7 + i];
  h->result_valid = false;
  return NDTB200_OK;
}

int ndtb200_lookup(ndtb200_handle* h, const void* points, size_t n, size_t stride_bytes, int search_method,
                   int32_t* out_keys) {
  if (!h || !out_keys || (!points && n)) return NDTB200_ERR_INVALID;
  if (n == 0) return NDTB200_OK;
  cudaSetDevice(h->device);
  int st = upload_points(h, h->d_out, points, n, stride_bytes);
  if (st != NDTB200_OK) return st;
  CK(h->d_tmp.ensure(n * 26 * sizeof(int32_t)));
  MapView map = make_view(h);
  const int blocks = static_cast<int>((n + 255) / 256);
  const float4* q = h->d_out.as<float4>();
  int32_t* d_keys = h->d_tmp.as<int32_t>();
  switch (search_method) {
    case NDTB200_DIRECT1: lookup_kernel<3><<<blocks, 256, 0, h->stream>>>(q, (int)n, map, d_keys); break;
    case NDTB200_DIRECT7: lookup_kernel<2><<<blocks, 256, 0, h->stream>>>(q, (int)n, map, d_keys); break;
    case NDTB200_DIRECT26: lookup_kernel<1><<<blocks, 256, 0, h->stream>>>(q, (int)n, map, d_keys); break;
    default: h->err = "ndtb200_lookup dumps the DIRECT1/7/26 neighbourhoods (26 columns); KDTREE is checked through eval / align"; return NDTB200_ERR_INVALID;
  }
  LAUNCHED(h);
  CK(cudaMemcpyAsync(out_keys, d_keys, n * 26 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return NDTB200_OK;
}

// convertTransform (ndt_omp.h:216-233): [x, y, z, roll, pitch, yaw] -> Translation * Rx * Ry * Rz as a column-major fp32
// 4x4, the same un-fused fp32 arithmetic the solver builds its trial poses with.  Host only.
int ndtb200_pose_to_matrix(const double p[6], float out16[16]) {
  if (!p || !out16) return NDTB200_ERR_INVALID;
  float T[12];
  pose_to_matrix(p, T);
  T_to_colmajor(T, out16);
  return NDTB200_OK;
}

int ndtb200_debug_guess_to_pose(const float guess[16], double p_out[6]) {
  if (!guess || !p_out) return NDTB200_ERR_INVALID;
  float T0[12];
  colmajor_to_T(guess, T0);
  guess_to_pose(T0, p_out);
  return NDTB200_OK;
}

int ndtb200_debug_newton_solve(ndtb200_handle* h, const double H[36], const double g[6], double delta_out[6], int* path_out) {
  if (!h || !H || !g || !delta_out) return NDTB200_ERR_INVALID;
  cudaSetDevice(h->device);
  CK(h->d_tmp.ensure(64 * sizeof(double)));
  double* d = h->d_tmp.as<double>();
  CK(cudaMemcpyAsync(d, H, 36 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(d + 36, g, 6 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  newton_solve_debug_kernel<<<1, 32, 0, h->stream>>>(d, d + 36, d + 42, reinterpret_cast<int*>(d + 48));
  LAUNCHED(h);
  double back[7];
  CK(cudaMemcpyAsync(back, d + 42, 7 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < 6; ++i) delta_out[i] = back[i];
  if (path_out) std::memcpy(path_out, &back[6], sizeof(int));
  return NDTB200_OK;
}

// parity helper: the on-device line-search trace of the last align (kind, pose, a_t, score per evaluation)
// profiling helper: CTA-0 timeline (ns since the first evaluation started) of the last solve, 4 stamps per evaluation
int ndtb200_get_timeline(ndtb200_handle* h, double* t4, int cap, int* n_out) {
  // t4 rows: start, local, reduced, advanced (contract) — debug stamps are printed to stderr when NDTB200_DEBUG_STEP is set
  if (!h || !n_out || !t4) return NDTB200_ERR_INVALID;
  if (!h->result_valid) {
    int st = ndtb200_sync(h);
    if (st != NDTB200_OK) return st;
  }
  cudaSetDevice(h->device);
  int n = h->h_result->n_trace;
  *n_out = n;
  if (n > ndtb200_handle::kTraceCap) n = ndtb200_handle::kTraceCap;
  if (n > cap) n = cap;
  if (n <= 0) return NDTB200_OK;
  std::vector<TraceRec> tr(n);
  CK(cudaMemcpyAsync(tr.data(), h->d_trace.p, n * sizeof(TraceRec), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const unsigned long long t0 = tr[0].t_start;
  if (getenv("NDTB200_DEBUG_STEP")) {  // arrival-time spread of the CTAs at the LAST barrier
    const int G = static_cast<int>(h->last_blocks);
    std::vector<double> part((size_t)G * kNVP);
    const size_t row_off = (h->comm_world == 1) ? (size_t)((h->h_result->n_trace - 1) & 1) * G * kNVP : 0;  // rows are double-buffered by evaluation parity
    CK(cudaMemcpyAsync(part.data(), h->d_partials.as<double>() + row_off, part.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    std::vector<double> arr(G);
    const unsigned long long tl = tr[n - 1].t_start;
    for (int b = 0; b < G; ++b) {
      long long bits; std::memcpy(&bits, &part[(size_t)b * kNVP + 31], 8);
      arr[b] = (double)((long long)bits - (long long)tl) * 1e-3;
    }
    std::vector<double> s(arr); std::sort(s.begin(), s.end());
    std::fprintf(stderr, "  last eval: CTA arrival after eval start [us]: min %.2f p10 %.2f p50 %.2f p90 %.2f p99 %.2f max %.2f (CTA0 %.2f, G=%d)\n",
                 s[0], s[G / 10], s[G / 2], s[G * 9 / 10], s[G * 99 / 100], s[G - 1], arr[0], G);
    {  // the slowest CTAs and the SMs they ran on (hardware position vs data: do the same SMs repeat?)
      std::vector<int> order(G);
      for (int b = 0; b < G; ++b) order[b] = b;
      std::sort(order.begin(), order.end(), [&](int x, int y) { return arr[x] > arr[y]; });
      std::fprintf(stderr, "  slowest CTAs (cta:sm:us):");
      for (int i = 0; i < std::min(G, 8); ++i) std::fprintf(stderr, " %d:%d:%.2f(h%.0f)", order[i], (int)part[(size_t)order[i] * kNVP + 30], arr[order[i]], part[(size_t)order[i] * kNVP + 28]);
      std::vector<double> hh(G); for (int b = 0; b < G; ++b) hh[b] = part[(size_t)b * kNVP + 28];
      std::sort(hh.begin(), hh.end());
      std::fprintf(stderr, "\n  hits per CTA: min %.0f p50 %.0f max %.0f\n", hh[0], hh[G / 2], hh[G - 1]);
    }
    // by SM pairing: CTAs b and b+148 usually share an SM
    double mx_lo = 0, mx_hi = 0; for (int b = 0; b < G; ++b) { if (b < G / 2) mx_lo = std::max(mx_lo, arr[b]); else mx_hi = std::max(mx_hi, arr[b]); }
    std::fprintf(stderr, "  max arrival first half of CTAs %.2f, second half %.2f\n", mx_lo, mx_hi);
  }
  for (int i = 0; i < n; ++i) {
    if (tr[i].t_start == 0) {  // a Hessian pass fused into the preceding line-search trial: no device time of its own
      const double t_prev = i > 0 ? t4[(i - 1) * 4 + 3] : 0.0;
      t4[i * 4 + 0] = t4[i * 4 + 1] = t4[i * 4 + 2] = t4[i * 4 + 3] = t_prev;
      continue;
    }
    t4[i * 4 + 0] = static_cast<double>(tr[i].t_start - t0);
    t4[i * 4 + 1] = static_cast<double>(tr[i].t_local - t0);
    t4[i * 4 + 2] = static_cast<double>(tr[i].t_reduced - t0);
    t4[i * 4 + 3] = static_cast<double>(tr[i].t_advanced - t0);
    if (getenv("NDTB200_DEBUG_STEP"))
      std::fprintf(stderr, "  eval %d CTA0: thread 0 points + warp sum %.2f us, (unused %.2f,) CTA fold %.2f us\n", i, (double)((long long)tr[i].t_phase[0] - (long long)tr[i].t_start) * 1e-3,
                   (double)((long long)tr[i].t_phase[1] - (long long)tr[i].t_phase[0]) * 1e-3, (double)((long long)tr[i].t_local - (long long)tr[i].t_phase[1]) * 1e-3);
    if (getenv("NDTB200_DEBUG_STEP"))
      std::fprintf(stderr, "  step %d: grid complete %.2f us after CTA 0 finished, totals %.2f us later; advance %.2f us, solve %.2f us, post %.2f us, pose setup %.2f us\n", i,
                   (double)((long long)tr[i].t_dbg[3] - (long long)tr[i].t_local) * 1e-3, (double)((long long)tr[i].t_reduced - (long long)tr[i].t_dbg[3]) * 1e-3,
                   (tr[i].t_dbg[0] - tr[i].t_reduced) * 1e-3, (tr[i].t_dbg[1] - tr[i].t_dbg[0]) * 1e-3,
                   (tr[i].t_dbg[2] - tr[i].t_dbg[1]) * 1e-3, (tr[i].t_advanced - tr[i].t_dbg[2]) * 1e-3);
  }
  return NDTB200_OK;
}

int ndtb200_get_trace(ndtb200_handle* h, int32_t* kinds, double* x6, double* a_t, double* score, int cap, int* n_out) {
  if (!h || !n_out) return NDTB200_ERR_INVALID;
  if (!h->result_valid) {
    int st = ndtb200_sync(h);
    if (st != NDTB200_OK) return st;
  }
  cudaSetDevice(h->device);
  int n = h->h_result->n_trace;
  *n_out = n;
  if (n > ndtb200_handle::kTraceCap) n = ndtb200_handle::kTraceCap;
  if (n > cap) n = cap;
  if (n <= 0) return NDTB200_OK;
  std::vector<TraceRec> tr(n);
  CK(cudaMemcpyAsync(tr.data(), h->d_trace.p, n * sizeof(TraceRec), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < n; ++i) {
    if (kinds) kinds[i] = tr[i].kind;
    if (x6) for (int k = 0; k < 6; ++k) x6[i * 6 + k] = tr[i].x[k];
    if (a_t) a_t[i] = tr[i].a_t;
    if (score) score[i] = tr[i].score;
  }
  return NDTB200_OK;
}

// ---- multi-GPU source sharding -----------------------------------------------------------------------------
static constexpr size_t kMailBytes = (size_t)2 * kMaxRanks * kNVP * 2 * sizeof(unsigned long long);

int ndtb200_comm_export(ndtb200_handle* h, void* handle_out64) {
  if (!h || !handle_out64) return NDTB200_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == NDTB200_COMM_HANDLE_BYTES, "IPC handle size");
  cudaSetDevice(h->device);
  if (h->d_mail.p == nullptr) {
    CK(cudaMalloc(&h->d_mail.p, kMailBytes));   // a dedicated allocation: IPC handles cover whole allocations
    h->d_mail.cap = kMailBytes;
    CK(cudaMemset(h->d_mail.p, 0, kMailBytes));
  }
  cudaIpcMemHandle_t ipc;
  CK(cudaIpcGetMemHandle(&ipc, h->d_mail.p));
  std::memcpy(handle_out64, &ipc, sizeof(ipc));
  return NDTB200_OK;
}

int ndtb200_comm_detach(ndtb200_handle* h) {
  if (!h) return NDTB200_ERR_INVALID;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (int r = 0; r < kMaxRanks; ++r) {
    if (h->mail_opened[r] && h->mail_ptrs[r]) cudaIpcCloseMemHandle(h->mail_ptrs[r]);
    h->mail_opened[r] = false;
    h->mail_ptrs[r] = nullptr;
  }
  h->comm_world = 1;
  h->comm_rank = 0;
  h->comm_n_total = 0;
  return NDTB200_OK;
}

int ndtb200_comm_attach(ndtb200_handle* h, int rank, int world, const void* all_handles, int64_t n_source_total) {
  if (!h || !all_handles || world < 1 || world > kMaxRanks || rank < 0 || rank >= world) return NDTB200_ERR_INVALID;
  if (h->d_mail.p == nullptr) { h->err = "call ndtb200_comm_export first"; return NDTB200_ERR_INVALID; }
  ndtb200_comm_detach(h);
  cudaSetDevice(h->device);
  CK(cudaMemset(h->d_mail.p, 0, kMailBytes));
  const cudaIpcMemHandle_t* hs = static_cast<const cudaIpcMemHandle_t*>(all_handles);
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      h->mail_ptrs[r] = h->d_mail.as<unsigned long long>();
    } else {
      void* p = nullptr;
      CK(cudaIpcOpenMemHandle(&p, hs[r], cudaIpcMemLazyEnablePeerAccess));
      h->mail_ptrs[r] = static_cast<unsigned long long*>(p);
      h->mail_opened[r] = true;
    }
  }
  h->comm_world = world;
  h->comm_rank = rank;
  h->comm_n_total = n_source_total;
  h->launch_seq = 0;  // tags are compared across ranks: all ranks restart the launch sequence together
  return NDTB200_OK;
}

int ndtb200_clone(const ndtb200_handle* src, ndtb200_handle** out) {
  if (!src || !out) return NDTB200_ERR_INVALID;
  ndtb200_handle* h = nullptr;
  int st = ndtb200_create(&h, src->device);
  if (st != NDTB200_OK) return st;
  h->prm = src->prm;
  h->shape = src->shape;
  if (src->has_target) {
    cudaStreamSynchronize(src->stream);
    // the copy keeps the source object's map: build it with the resolution the source map was built with
    ndtb200_params p = src->prm;
    if (src->grid.leaf[0] > 0) h->prm.resolution = src->grid.leaf[0];
    st = ndtb200_set_target_device(h, target_pts(src), src->n_target, src->target_dense ? 1 : 0);
    h->prm = p;
    if (st == NDTB200_ERR_CUDA) { ndtb200_destroy(h); return st; }
  }
  if (src->has_source) {
    st = ndtb200_set_source_device(h, src->d_source.p, src->n_source);
    if (st != NDTB200_OK) { ndtb200_destroy(h); return st; }
  }
  std::memcpy(h->h_result, src->h_result, sizeof(AlignResultDev));
  h->result_valid = src->result_valid;
  std::memcpy(h->last_final_T, src->last_final_T, sizeof(h->last_final_T));
  cudaStreamSynchronize(h->stream);
  *out = h;
  return NDTB200_OK;
}

// ---- sharded target-map build: points by contiguous range, one exchange of per-voxel partials -------
int ndtb200_cloud_bounds(ndtb200_handle* h, const void* d_points, size_t n, int is_dense, float out_min[3], float out_max[3],
                         int64_t* n_finite) {
  if (!h || (!d_points && n) || !out_min || !out_max) return NDTB200_ERR_INVALID;
  cudaSetDevice(h->device);
  h->target_view = nullptr;
  if (n) {
    CK(h->d_target.ensure(n * sizeof(float4)));
    CK(cudaMemcpyAsync(h->d_target.p, d_points, n * sizeof(float4), cudaMemcpyDeviceToDevice, h->stream));
  }
  h->n_target = n;
  h->target_dense = is_dense != 0;
  h->has_target = true;
  std::memset(&h->grid, 0, sizeof(GridDesc));
  for (int a = 0; a < 3; ++a) { out_min[a] = 3.402823466e+38f; out_max[a] = -3.402823466e+38f; }
  if (n_finite) *n_finite = 0;
  if (n == 0) return NDTB200_OK;
  int st = compute_grid(h, h->d_target.as<float4>(), n, is_dense ? 1 : 0, BuildOpts());
  if (st != NDTB200_OK) return st;
  for (int a = 0; a < 3; ++a) { out_min[a] = h->grid.min_p[a]; out_max[a] = h->grid.max_p[a]; }
  if (n_finite) *n_finite = h->grid.n_finite;
  return NDTB200_OK;
}

int ndtb200_build_partials(ndtb200_handle* h, const float global_min[3], const float global_max[3], int64_t* n_partials) {
  if (!h || !global_min || !global_max || !n_partials) return NDTB200_ERR_INVALID;
  cudaSetDevice(h->device);
  BuildOpts o;
  o.forced_min = global_min;
  o.forced_max = global_max;
  o.partial_only = true;
  const int st = build_map_ex(h, o);
  *n_partials = (st == NDTB200_OK) ? static_cast<int64_t>(h->n_partials) : 0;
  return st;
}

int ndtb200_copy_partials(ndtb200_handle* h, void* d_keys, void* d_counts, void* d_moments) {
  if (!h) return NDTB200_ERR_INVALID;
  cudaSetDevice(h->device);
  const size_t v = h->n_partials;
  if (v == 0) return NDTB200_OK;
  if (!d_keys || !d_counts || !d_moments) return NDTB200_ERR_INVALID;
  CK(cudaMemcpyAsync(d_keys, h->d_voxel_key.p, v * sizeof(int32_t), cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemcpyAsync(d_counts, h->d_voxel_count.p, v * sizeof(uint32_t), cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemcpyAsync(d_moments, h->d_moments.p, v * 9 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return NDTB200_OK;
}

int ndtb200_build_from_partials(ndtb200_handle* h, const float global_min[3], const float global_max[3], int64_t n_finite_total,
                                const void* d_keys, const void* d_counts, const void* d_moments, size_t n_total) {
  if (!h || !global_min || !global_max || (n_total && (!d_keys || !d_counts || !d_moments))) return NDTB200_ERR_INVALID;
  if (n_total > 0xFFFFFFF0ull) { h->err = "too many partials"; return NDTB200_ERR_INVALID; }
  cudaSetDevice(h->device);
  h->has_target = true;
  return build_from_partials(h, global_min, global_max, n_finite_total, static_cast<const uint32_t*>(d_keys),
                             static_cast<const uint32_t*>(d_counts), static_cast<const double*>(d_moments), n_total);
}

// ---- sharded build, owner-partitioned: partials travel to the rank that OWNS their key range (all-to-all),
// the owner merges / finalises its voxels, the finished 64-byte records are all-gathered and installed everywhere ----
__global__ void lower_bound_kernel(const int32_t* __restrict__ keys, uint32_t n, const int32_t* __restrict__ bounds, int nb,
                                   long long* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  const int32_t b = bounds[i];
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = lo + (hi - lo) / 2;
    if (keys[mid] < b) lo = mid + 1; else hi = mid;
  }
  out[i] = lo;
}

int ndtb200_partials_split(ndtb200_handle* h, const int32_t* upper_keys, int world, int64_t* offsets_out) {
  if (!h || !offsets_out || world < 1 || (world > 1 && !upper_keys)) return NDTB200_ERR_INVALID;
  cudaSetDevice(h->device);
  const uint32_t n = static_cast<uint32_t>(h->n_partials);
  offsets_out[0] = 0;
  offsets_out[world] = n;
  if (world == 1) return NDTB200_OK;
  CK(h->d_tmp.ensure((size_t)world * (sizeof(int32_t) + sizeof(long long)) + 64));
  long long* d_out = h->d_tmp.as<long long>();
  int32_t* d_b = reinterpret_cast<int32_t*>(d_out + world);
  CK(cudaMemcpyAsync(d_b, upper_keys, (size_t)(world - 1) * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  lower_bound_kernel<<<1, 32 * ((world + 30) / 32), 0, h->stream>>>(h->d_voxel_key.as<int32_t>(), n, d_b, world - 1, d_out);
  LAUNCHED(h);
  std::vector<long long> back(world - 1);
  CK(cudaMemcpyAsync(back.data(), d_out, (size_t)(world - 1) * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int r = 1; r < world; ++r) offsets_out[r] = back[r - 1];
  return NDTB200_OK;
}

int ndtb200_merge_partials(ndtb200_handle* h, const float global_min[3], const float global_max[3], int64_t n_finite_total,
                           const void* d_keys, const void* d_counts, const void* d_moments, size_t n_total, int64_t* n_merged) {
  if (!h || !global_min || !global_max || !n_merged || (n_total && (!d_keys || !d_counts || !d_moments))) return NDTB200_ERR_INVALID;
  if (n_total > 0xFFFFFFF0ull) { h->err = "too many partials"; return NDTB200_ERR_INVALID; }
  cudaSetDevice(h->device);
  h->has_target = true;
  const int st = build_from_partials(h, global_min, global_max, n_finite_total, static_cast<const uint32_t*>(d_keys),
                                     static_cast<const uint32_t*>(d_counts), static_cast<const double*>(d_moments), n_total, /*records_only=*/true);
  *n_merged = (st == NDTB200_OK) ? h->n_voxels : 0;
  return st;
}

int ndtb200_copy_records(ndtb200_handle* h, void* d_records_out, void* d_icov64_out) {
  if (!h) return NDTB200_ERR_INVALID;
  cudaSetDevice(h->device);
  const size_t v = static_cast<size_t>(h->n_voxels);
  if (v == 0) return NDTB200_OK;
  if (!d_records_out || !d_icov64_out) return NDTB200_ERR_INVALID;
  CK(cudaMemcpyAsync(d_records_out, h->d_records.p, v * sizeof(VoxelRecord), cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemcpyAsync(d_icov64_out, h->d_icov64.p, v * 6 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return NDTB200_OK;
}

__global__ void record_keys_kernel(const VoxelRecord* __restrict__ records, uint32_t n, int32_t* __restrict__ keys, int min_points,
                                   unsigned int* __restrict__ n_valid) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  keys[v] = records[v].key;
  if (records[v].count >= min_points) atomicAdd(n_valid, 1u);
}

int ndtb200_set_map_from_records(ndtb200_handle* h, const float global_min[3], const float global_max[3], int64_t n_finite_total,
                                 const void* d_records, const void* d_icov64, size_t n_total) {
  if (!h || !global_min || !global_max || (n_total && (!d_records || !d_icov64))) return NDTB200_ERR_INVALID;
  if (n_total > 0x7FFFFFF0ull) { h->err = "too many voxels"; return NDTB200_ERR_INVALID; }
  cudaSetDevice(h->device);
  h->has_target = true;
  clear_map(h);
  int st = grid_from_global_box(h, global_min, global_max, n_finite_total);
  if (st != NDTB200_OK) return st;
  if (n_finite_total == 0 || n_total == 0) {
    h->map_status = NDTB200_ERR_NO_INPUT;
    st = ensure_empty_hash(h);
    return st != NDTB200_OK ? st : NDTB200_ERR_NO_INPUT;
  }
  if (h->grid.overflow) {
    h->map_status = NDTB200_ERR_GRID_OVERFLOW;
    st = ensure_empty_hash(h);
    return st != NDTB200_OK ? st : NDTB200_ERR_GRID_OVERFLOW;
  }
  const uint32_t n_vox = static_cast<uint32_t>(n_total);
  CK(h->d_records.ensure(n_total * sizeof(VoxelRecord)));
  CK(h->d_icov64.ensure(n_total * 6 * sizeof(double)));
  CK(h->d_voxel_key.ensure(n_total * sizeof(int32_t)));
  CK(cudaMemcpyAsync(h->d_records.p, d_records, n_total * sizeof(VoxelRecord), cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_icov64.p, d_icov64, n_total * 6 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  unsigned int* d_nvalid = h->d_scalar.as<unsigned int>() + 4;
  CK(cudaMemsetAsync(d_nvalid, 0, sizeof(unsigned int), h->stream));
  const int vblocks = static_cast<int>(((size_t)n_vox + kBuildThreads - 1) / kBuildThreads);
  record_keys_kernel<<<vblocks, kBuildThreads, 0, h->stream>>>(h->d_records.as<VoxelRecord>(), n_vox, h->d_voxel_key.as<int32_t>(),
                                                                h->prm.min_points_per_voxel, d_nvalid);
  LAUNCHED(h);
  h->n_voxels = n_vox;
  h->moments_valid = false;
  h->map_is_merged = true;
  h->records_only_map = true;
  st = finalize_and_index(h, n_vox, 0u, nullptr, /*records_done=*/true);
  if (st != NDTB200_OK) return st;
  h->map_status = NDTB200_OK;
  return NDTB200_OK;
}

// ---- pcl::VoxelGrid centroid downsample on the device (the step before the path in every caller) ----
static int voxelgrid_filter_impl(ndtb200_handle* h, const float leaf[3], int64_t* n_out) {
  // the cloud sits in a->d_target (n_target points); the result is left in a->d_out
  ndtb200_handle* a = h;
  *n_out = 0;
  if (!(leaf[0] > 0 && leaf[1] > 0 && leaf[2] > 0)) { a->err = "leaf size must be positive"; return NDTB200_ERR_INVALID; }
  const size_t n = a->n_target;
  if (n == 0) return NDTB200_OK;
  a->prm.resolution = leaf[0];
  for (int k = 0; k < 3; ++k) a->vg_leaf[k] = leaf[k];
  const float4* pts = a->d_target.as<float4>();
  std::memset(&a->grid, 0, sizeof(GridDesc));
  if (use_fused_build(a, n)) {  // scan-sized cloud: one cooperative launch, centroids left in a->d_out
    uint32_t n_vox = 0;
    const int stf = run_fused_build(a, pts, n, /*dense=*/0, 1, &n_vox);
    if (stf != NDTB200_OK) return stf;
    if (a->grid.n_finite == 0) return NDTB200_OK;
    if (a->grid.overflow) return NDTB200_ERR_GRID_OVERFLOW;
    *n_out = n_vox;
    return NDTB200_OK;
  }
  int st = compute_grid(a, pts, n, /*dense=*/0, BuildOpts());
  if (st != NDTB200_OK) return st;
  if (a->grid.n_finite == 0) return NDTB200_OK;
  if (a->grid.overflow) return NDTB200_ERR_GRID_OVERFLOW;  // pcl::VoxelGrid warns and passes the cloud through
  uint32_t sentinel = 0;
  const int passes = passes_for(a->grid, true, &sentinel);
  CK2(a, a->d_keys_a.ensure(n * sizeof(uint32_t)));
  CK2(a, a->d_keys_b.ensure(n * sizeof(uint32_t)));
  CK2(a, a->d_vals_a.ensure(n * sizeof(uint32_t)));
  CK2(a, a->d_vals_b.ensure(n * sizeof(uint32_t)));
  st = sort_meta_reset(a);
  if (st != NDTB200_OK) return st;
  const int key_blocks = grid_for(n, kBuildThreads * 4, a->num_sms * 8);
  voxel_key_kernel<<<key_blocks, kBuildThreads, 0, a->stream>>>(pts, n, 0, a->d_grid.as<GridDesc>(), sentinel,
                                                                 a->d_keys_a.as<uint32_t>(), nullptr, sort_meta_hist(a), passes);
  LAUNCHED(a);
  uint32_t n_vox = 0;
  st = sort_and_segment(a, n, sentinel, passes, &n_vox, /*hist_ready=*/true);
  if (st != NDTB200_OK) return st;
  CK2(a, a->d_out.ensure((size_t)std::max<uint32_t>(n_vox, 1u) * sizeof(float4)));
  if (n_vox > 0) {
    const int vblocks = static_cast<int>(((size_t)n_vox + kBuildThreads - 1) / kBuildThreads);
    voxel_centroid_kernel<<<vblocks, kBuildThreads, 0, a->stream>>>(pts, a->d_vals_a.as<uint32_t>(), a->d_voxel_start.as<uint32_t>(),
                                                                     n_vox, static_cast<uint32_t>(a->grid.n_finite), a->d_out.as<float4>());
    LAUNCHED(a);
  }
  *n_out = n_vox;
  return NDTB200_OK;
}


int ndtb200_voxelgrid_filter(ndtb200_handle* h, const void* points, size_t n, size_t stride_bytes, float leaf, void* out_points,
                             size_t out_capacity, size_t out_stride_bytes, int64_t* n_out) {
  const float l3[3] = {leaf, leaf, leaf};
  return ndtb200_voxelgrid_filter3(h, points, n, stride_bytes, l3, out_points, out_capacity, out_stride_bytes, n_out);
}

int ndtb200_voxelgrid_filter3(ndtb200_handle* h, const void* points, size_t n, size_t stride_bytes, const float leaf[3], void* out_points,
                              size_t out_capacity, size_t out_stride_bytes, int64_t* n_out) {
  if (!h || !n_out || !leaf || (!points && n)) return NDTB200_ERR_INVALID;
  if (n > 0xFFFFFFF0ull) { h->err = "cloud too large"; return NDTB200_ERR_INVALID; }
  cudaSetDevice(h->device);
  int st = ensure_aux(h);
  if (st != NDTB200_OK) return st;
  ndtb200_handle* a = h->aux;
  st = upload_points(a, a->d_target, points, n, stride_bytes);
  if (st != NDTB200_OK) { h->err = a->err; return st; }
  a->n_target = n;
  a->has_target = true;
  st = voxelgrid_filter_impl(a, leaf, n_out);
  if (st != NDTB200_OK) { h->err = a->err; return st; }
  const size_t m = static_cast<size_t>(*n_out);
  if (out_points && m > 0) {
    if (out_stride_bytes < 16) { h->err = "out_stride_bytes must be >= 16"; return NDTB200_ERR_INVALID; }
    if (m > out_capacity) { h->err = "output buffer too small"; return NDTB200_ERR_INVALID; }
    if (out_stride_bytes == 16) CK(cudaMemcpyAsync(out_points, a->d_out.p, m * 16, cudaMemcpyDeviceToHost, a->stream));
    else CK(cudaMemcpy2DAsync(out_points, out_stride_bytes, a->d_out.p, 16, 16, m, cudaMemcpyDeviceToHost, a->stream));
  }
  CK(cudaStreamSynchronize(a->stream));
  return NDTB200_OK;
}

int ndtb200_voxelgrid_filter_device(ndtb200_handle* h, const void* d_points_xyzw, size_t n, float leaf, void* d_out_xyzw,
                                    size_t out_capacity, int64_t* n_out) {
  if (!h || !n_out || (!d_points_xyzw && n)) return NDTB200_ERR_INVALID;
  if (n > 0xFFFFFFF0ull) { h->err = "cloud too large"; return NDTB200_ERR_INVALID; }
  cudaSetDevice(h->device);
  int st = ensure_aux(h);
  if (st != NDTB200_OK) return st;
  ndtb200_handle* a = h->aux;
  if (n) {
    CK(a->d_target.ensure(n * sizeof(float4)));
    CK(cudaMemcpyAsync(a->d_target.p, d_points_xyzw, n * sizeof(float4), cudaMemcpyDeviceToDevice, a->stream));
  }
  a->n_target = n;
  a->has_target = true;
  const float l3[3] = {leaf, leaf, leaf};
  st = voxelgrid_filter_impl(a, l3, n_out);
  if (st != NDTB200_OK) { h->err = a->err; return st; }
  const size_t m = static_cast<size_t>(*n_out);
  if (d_out_xyzw && m > 0) {
    if (m > out_capacity) { h->err = "output buffer too small"; return NDTB200_ERR_INVALID; }
    CK(cudaMemcpyAsync(d_out_xyzw, a->d_out.p, m * 16, cudaMemcpyDeviceToDevice, a->stream));
  }
  CK(cudaStreamSynchronize(a->stream));
  return NDTB200_OK;
}

// =================================================================================================
// The mapping-node loop as a device-resident pipeline (lidar_subscriber/src/ndt_rosbag_mapping_node.cpp
// :42-75 run loop, :108-118 downsample_cloud, :120-144 perform_registration, :146-161 update_global_map).
// Only the raw scan goes up and one small step record comes down; the filtered scan, the voxel maps, the aligned
// source and the global map never leave HBM.  Two NDT handles alternate: while handle A aligns scan k+1 against the
// map of scan k, handle B already builds the map of scan k+1 on its own stream.
// =================================================================================================
struct ndtb200_mapper {
  int device = 0;
  ndtb200_handle* ndt[2] = {nullptr, nullptr};  // ndt[cur] holds the map of the previous scan
  ndtb200_handle* vg = nullptr;                 // VoxelGrid filters (scan leaf, map leaf)
  int cur = 0;
  float voxel_leaf = 0.3f, map_voxel = 0.5f;
  int compute_fitness = 1;
  bool have_prev = false;
  float pose[16];            // column-major, accumulated pose (Eigen::Matrix4f pose = pose * transform)
  float pres_transform[16];  // column-major, guess of the next registration
  DevBuf d_raw, d_filtered, d_transformed, d_map[2];
  size_t n_map = 0;
  int map_cur = 0;
  long long scans = 0;
  std::string err;
};

static void mat4_identity(float* m) { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.0f : 0.0f; }
static void mat4_mul_colmajor(const float* a, const float* b, float* c) {  // c = a * b, fp32
  float r[16];
  for (int col = 0; col < 4; ++col)
    for (int row = 0; row < 4; ++row) {
      float s = 0.f;
      for (int k = 0; k < 4; ++k) s += a[k * 4 + row] * b[col * 4 + k];
      r[col * 4 + row] = s;
    }
  std::memcpy(c, r, sizeof(r));
}

#define MCK(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      m->err = std::string(#call) + ": " + cudaGetErrorString(e__);                                \
      return NDTB200_ERR_CUDA;                                                                     \
    }                                                                                              \
  } while (0)

int ndtb200_mapper_create(ndtb200_mapper** out, int device, const ndtb200_params* ndt_params, float voxel_leaf, float map_voxel,
                          int compute_fitness) {
  if (!out) return NDTB200_ERR_INVALID;
  *out = nullptr;
  ndtb200_mapper* m = new ndtb200_mapper();
  m->device = device;
  m->voxel_leaf = voxel_leaf;
  m->map_voxel = map_voxel;
  m->compute_fitness = compute_fitness;
  mat4_identity(m->pose);
  mat4_identity(m->pres_transform);
  int st = ndtb200_create(&m->ndt[0], device);
  if (st == NDTB200_OK) st = ndtb200_create(&m->ndt[1], device);
  if (st == NDTB200_OK) st = ndtb200_create(&m->vg, device);
  ndtb200_params p;
  ndtb200_default_params(&p);
  if (ndt_params) p = *ndt_params;
  else { p.trans_eps = 0.01; p.max_iterations = 64; }  // the node's defaults (ndt_rosbag_mapping_node.cpp:83-88)
  for (int i = 0; i < 2 && st == NDTB200_OK; ++i) st = ndtb200_set_params(m->ndt[i], &p);
  for (int i = 0; i < 2 && st == NDTB200_OK; ++i) {
    // small-CTA solve shape (16 k registers per SM) so that the other handle's fused build kernel (one 256-thread CTA per
    // SM) can be co-resident: the map of scan k+1 is built while scan k+1 is still being aligned
    m->ndt[i]->shape = 1;
    m->ndt[i]->prefer_fused_build = true;
  }
  if (st == NDTB200_OK) st = ensure_aux(m->vg);
  if (st == NDTB200_OK) {
    // reserve the buffers that grow with the drive up front (a 2 M-point global map, 256 k-point scans): cudaMalloc /
    // cudaFree in the middle of the loop are device-wide synchronisations worth tens of milliseconds
    const size_t map_pts = 2u << 20, scan_pts = 256u << 10;
    ndtb200_handle* a = m->vg->aux;
    bool ok = m->d_map[0].ensure(map_pts * 16) == cudaSuccess && m->d_map[1].ensure(map_pts * 16) == cudaSuccess &&
              m->d_raw.ensure(scan_pts * 16) == cudaSuccess && m->d_filtered.ensure(scan_pts * 16) == cudaSuccess &&
              a->d_target.ensure(map_pts * 16) == cudaSuccess && a->d_out.ensure(map_pts * 16) == cudaSuccess &&
              a->d_keys_a.ensure(map_pts * 4) == cudaSuccess && a->d_keys_b.ensure(map_pts * 4) == cudaSuccess &&
              a->d_vals_a.ensure(map_pts * 4) == cudaSuccess && a->d_vals_b.ensure(map_pts * 4) == cudaSuccess &&
              a->d_voxel_key.ensure(map_pts * 4) == cudaSuccess && a->d_voxel_start.ensure(map_pts * 4) == cudaSuccess &&
              a->d_status.ensure((size_t)kMaxSortPasses * (map_pts / kSortTile + 1) * 256 * sizeof(unsigned long long)) == cudaSuccess &&
              a->d_sortmeta.ensure(kSortMetaBytes) == cudaSuccess && a->d_hist.ensure((map_pts / kScanTile + 64) * 4) == cudaSuccess &&
              a->d_scan_tmp.ensure(scan_tmp_elems(map_pts / kScanTile + 64) * 4) == cudaSuccess;
    // the two registration handles: everything a 64 k-point scan touches (build scratch, records, cell tables over a
    // 32 MB = 8 M-cell grid, solver workspace), so that no scan of the drive is the first to allocate (a cudaFree +
    // cudaMalloc pair in the loop was measured at 4 - 40 ms)
    const size_t pts = kSmallMaxPoints, table_b = 32u << 20;
    for (int i = 0; i < 2 && ok; ++i) {
      ndtb200_handle* n = m->ndt[i];
      ok = n->d_target.ensure(pts * 16) == cudaSuccess && n->d_source.ensure(pts * 16) == cudaSuccess && n->d_out.ensure(pts * 16) == cudaSuccess &&
           n->d_keys_a.ensure(pts * 4) == cudaSuccess && n->d_keys_b.ensure(pts * 4) == cudaSuccess && n->d_vals_a.ensure(pts * 4) == cudaSuccess &&
           n->d_vals_b.ensure(pts * 4) == cudaSuccess && n->d_voxel_key.ensure(pts * 4) == cudaSuccess && n->d_voxel_start.ensure(pts * 4) == cudaSuccess &&
           n->d_moments.ensure(pts * 72) == cudaSuccess && n->d_records.ensure(pts * sizeof(VoxelRecord)) == cudaSuccess &&
           n->d_icov64.ensure(pts * 48) == cudaSuccess && n->d_hist.ensure((size_t)256 * (pts / kSmallTile + 1) * 4) == cudaSuccess &&
           n->d_scan_tmp.ensure((pts / kScanTile + 64 + 1024 + 64) * 4) == cudaSuccess && n->d_mm_partial.ensure((size_t)n->num_sms * 8 * 24) == cudaSuccess &&
           n->d_mm_finite.ensure((size_t)n->num_sms * 8 * 4) == cudaSuccess && n->d_grid.ensure(sizeof(GridDesc)) == cudaSuccess &&
           n->d_dense.ensure(table_b) == cudaSuccess && n->d_cell_all.ensure(table_b) == cudaSuccess &&
           n->d_best.ensure(pts * 8 + 64) == cudaSuccess && n->d_tmp.ensure(pts * 16) == cudaSuccess &&
           n->d_partials.ensure((size_t)2 * n->num_sms * kNVP * sizeof(double)) == cudaSuccess && n->d_result.ensure(sizeof(AlignResultDev)) == cudaSuccess &&
           n->d_trace.ensure(ndtb200_handle::kTraceCap * sizeof(TraceRec)) == cudaSuccess;
    }
    if (!ok) st = NDTB200_ERR_CUDA;
  }
  if (st != NDTB200_OK) { ndtb200_mapper_destroy(m); return st; }
  *out = m;
  return NDTB200_OK;
}

int ndtb200_mapper_destroy(ndtb200_mapper* m) {
  if (!m) return NDTB200_OK;
  cudaSetDevice(m->device);
  for (int i = 0; i < 2; ++i) if (m->ndt[i]) ndtb200_destroy(m->ndt[i]);
  if (m->vg) ndtb200_destroy(m->vg);
  m->d_raw.release(); m->d_filtered.release(); m->d_transformed.release(); m->d_map[0].release(); m->d_map[1].release();
  delete m;
  return NDTB200_OK;
}

const char* ndtb200_mapper_last_error(const ndtb200_mapper* m) { return m ? m->err.c_str() : "null mapper"; }

// update_global_map (:146-161): transform the filtered scan by the pose, append, re-voxelise the whole map
static int mapper_update_global_map(ndtb200_mapper* m, size_t n_f) {
  ndtb200_handle* v = m->vg;
  cudaStream_t s = v->stream;
  const size_t total = m->n_map + n_f;
  DevBuf& cat = m->d_map[1 - m->map_cur];
  MCK(cat.ensure(std::max<size_t>(total, 1) * sizeof(float4)));
  if (m->n_map) MCK(cudaMemcpyAsync(cat.p, m->d_map[m->map_cur].p, m->n_map * sizeof(float4), cudaMemcpyDeviceToDevice, s));
  if (n_f) {
    Mat34 T;
    colmajor_to_T(m->pose, T.m);
    transform_cloud_kernel<<<grid_for(n_f, 256, v->num_sms * 8), 256, 0, s>>>(m->d_filtered.as<float4>(), n_f, T,
                                                                              cat.as<float4>() + m->n_map);
    v->launches++;
  }
  MCK(cudaStreamSynchronize(s));
  DevBuf& dst = m->d_map[m->map_cur];
  MCK(dst.ensure(std::max<size_t>(total, 1) * sizeof(float4)));
  int64_t n_out = 0;
  const int st = ndtb200_voxelgrid_filter_device(v, cat.p, total, m->map_voxel, dst.p, total, &n_out);
  if (st == NDTB200_ERR_GRID_OVERFLOW) {  // pcl::VoxelGrid passes the cloud through
    MCK(cudaMemcpy(dst.p, cat.p, total * sizeof(float4), cudaMemcpyDeviceToDevice));
    n_out = static_cast<int64_t>(total);
  } else if (st != NDTB200_OK) {
    m->err = ndtb200_last_error(v);
    return st;
  }
  m->n_map = static_cast<size_t>(n_out);
  return NDTB200_OK;
}

int ndtb200_mapper_push_scan(ndtb200_mapper* m, const void* points, size_t n, size_t stride_bytes, ndtb200_mapper_step* out) {
  if (!m || !out || (!points && n)) return NDTB200_ERR_INVALID;
  cudaSetDevice(m->device);
  std::memset(out, 0, sizeof(*out));
  mat4_identity(out->transform);
  ndtb200_handle* v = m->vg;
  // downsample_cloud (:108-118): raw scan up, centroids stay on the device
  int st = upload_points(v, m->d_raw, points, n, stride_bytes);
  if (st != NDTB200_OK) { m->err = v->err; return st; }
  MCK(cudaStreamSynchronize(v->stream));
  MCK(m->d_filtered.ensure(std::max<size_t>(n, 1) * sizeof(float4)));
  int64_t n_f64 = 0;
  st = ndtb200_voxelgrid_filter_device(v, m->d_raw.p, n, m->voxel_leaf, m->d_filtered.p, n, &n_f64);
  if (st == NDTB200_ERR_GRID_OVERFLOW) {
    MCK(cudaMemcpy(m->d_filtered.p, m->d_raw.p, n * sizeof(float4), cudaMemcpyDeviceToDevice));
    n_f64 = static_cast<int64_t>(n);
  } else if (st != NDTB200_OK) {
    m->err = ndtb200_last_error(v);
    return st;
  }
  const size_t n_f = static_cast<size_t>(n_f64);
  out->n_filtered = n_f64;
  ndtb200_handle* a = m->ndt[m->cur];       // holds the map of the previous scan
  ndtb200_handle* b = m->ndt[1 - m->cur];   // builds the map of this scan meanwhile
  if (m->have_prev) {
    // perform_registration (:120-144): target = previous filtered scan (already built), source = this one
    st = ndtb200_set_source_device(a, m->d_filtered.p, n_f);
    if (st == NDTB200_OK) st = ndtb200_align_async(a, m->pres_transform);
    if (st != NDTB200_OK) { m->err = ndtb200_last_error(a); return st; }
  }
  // the map of THIS scan (next step's target) is built on the other handle's stream while the solve runs
  int st_b = ndtb200_set_target_device(b, m->d_filtered.p, n_f, 1);
  if (st_b == NDTB200_ERR_CUDA || st_b == NDTB200_ERR_INVALID) { m->err = ndtb200_last_error(b); return st_b; }
  if (m->have_prev) {
    ndtb200_result r;
    st = ndtb200_sync(a);
    if (st == NDTB200_OK) st = ndtb200_get_result(a, &r);
    if (st != NDTB200_OK) { m->err = ndtb200_last_error(a); return st; }
    out->converged = r.converged;
    out->iterations = r.iterations;
    out->n_evaluations = r.n_evaluations;
    if (m->compute_fitness) {
      double f = 0;
      if (ndtb200_fitness_score(a, 1.7976931348623157e308, &f) == NDTB200_OK) out->fitness = f;
    }
    if (r.converged) std::memcpy(out->transform, r.final_transformation, sizeof(out->transform));  // else Identity (:140-143)
    std::memcpy(m->pres_transform, out->transform, sizeof(m->pres_transform));
    mat4_mul_colmajor(m->pose, out->transform, m->pose);  // pose = pose * transform (:66)
  }
  std::memcpy(out->pose, m->pose, sizeof(out->pose));
  st = mapper_update_global_map(m, n_f);
  if (st != NDTB200_OK) return st;
  out->n_map = static_cast<int64_t>(m->n_map);
  m->cur = 1 - m->cur;
  m->have_prev = true;
  m->scans++;
  return NDTB200_OK;
}

int ndtb200_mapper_get_map(ndtb200_mapper* m, void* out_points, size_t capacity, size_t out_stride_bytes, int64_t* n_out) {
  if (!m || !n_out) return NDTB200_ERR_INVALID;
  cudaSetDevice(m->device);
  *n_out = static_cast<int64_t>(m->n_map);
  if (!out_points || m->n_map == 0) return NDTB200_OK;
  if (out_stride_bytes < 16 || capacity < m->n_map) { m->err = "output buffer too small / stride < 16"; return NDTB200_ERR_INVALID; }
  if (out_stride_bytes == 16) MCK(cudaMemcpy(out_points, m->d_map[m->map_cur].p, m->n_map * 16, cudaMemcpyDeviceToHost));
  else MCK(cudaMemcpy2D(out_points, out_stride_bytes, m->d_map[m->map_cur].p, 16, 16, m->n_map, cudaMemcpyDeviceToHost));
  return NDTB200_OK;
}

int64_t ndtb200_mapper_launch_count(const ndtb200_mapper* m) {
  if (!m) return 0;
  long long c = m->ndt[0]->launches + m->ndt[1]->launches + m->vg->launches;
  if (m->vg->aux) c += m->vg->aux->launches;
  return c;
}

int ndtb200_set_throughput_mode(ndtb200_handle* h, int on) {
  if (!h) return NDTB200_ERR_INVALID;
  h->shape = on ? 1 : 0;
  return NDTB200_OK;
}

// Independent scan pairs in flight together (SURVEY §8b "align_batch", §8e "batched scan pairs"): every handle owns
// its map, source and stream; all solves are enqueued with the throughput CTA shape before the first wait, so up to
// four persistent kernels share every SM and each one's barrier / Newton-step latency is covered by the others.
// Independent scan pairs in flight together (SURVEY §8b "align_batch", §8e "batched scan pairs"): every handle owns
// its map, source and stream; all solves are enqueued with the throughput CTA shape before the first wait, so up to
// four persistent kernels share every SM and each one's barrier / Newton-step latency is covered by the others.
// (Measured and rejected in round 2: dividing the SMs between the pairs instead — every solve a 1..8-CTA grid /
// thread-block cluster of 1024-thread CTAs, all pairs side by side — reaches 7.6 k aligns/s on c2 against 15.4 k for
// the shared full-width grids: with many points per thread the pass runs at about half the issue rate, and only as
// many kernels run concurrently as the process has hardware queues — see CUDA_DEVICE_MAX_CONNECTIONS in bench.py.)
int ndtb200_align_batch_async(ndtb200_handle* const* hs, int n, const float* guesses, void* const* out_points,
                              size_t out_stride_bytes) {
  if (!hs || n < 0) return NDTB200_ERR_INVALID;
  int rc = NDTB200_OK;
  for (int i = 0; i < n; ++i) {
    if (!hs[i]) return NDTB200_ERR_INVALID;
    const int keep = hs[i]->shape;
    hs[i]->shape = 1;
    hs[i]->want_fused_out = out_points != nullptr && out_points[i] != nullptr && out_stride_bytes >= 16;
    int st = ndtb200_align_async(hs[i], guesses ? guesses + 16 * (size_t)i : nullptr);
    hs[i]->want_fused_out = false;
    hs[i]->shape = keep;
    if (st == NDTB200_OK && out_points) st = enqueue_output(hs[i], out_points[i], out_stride_bytes);
    if (st == NDTB200_OK) st = enqueue_result_copy(hs[i]);
    if (st != NDTB200_OK && rc == NDTB200_OK) rc = st;
  }
  return rc;
}

int ndtb200_align_batch(ndtb200_handle* const* hs, int n, const float* guesses, void* const* out_points,
                        size_t out_stride_bytes, ndtb200_result* results) {
  int rc = ndtb200_align_batch_async(hs, n, guesses, out_points, out_stride_bytes);
  if (rc == NDTB200_ERR_INVALID) return rc;
  for (int i = 0; i < n; ++i) {
    int st = ndtb200_sync(hs[i]);
    if (st == NDTB200_OK && results) st = ndtb200_get_result(hs[i], results + i);
    if (st != NDTB200_OK && rc == NDTB200_OK) rc = st;
  }
  return rc;
}

int ndtb200_run_pairs(ndtb200_handle* const* lanes, int n_lanes, const void* const* targets, const size_t* n_targets,
                      const void* const* sources, const size_t* n_sources, size_t stride_bytes, const float* guesses16,
                      int n_pairs, ndtb200_result* results) {
  if (!lanes || n_lanes < 1 || n_pairs < 0 || (n_pairs && (!targets || !n_targets || !sources || !n_sources || !results))) return NDTB200_ERR_INVALID;
  for (int l = 0; l < n_lanes; ++l)
    if (!lanes[l]) return NDTB200_ERR_INVALID;
  std::vector<int> status(n_lanes, NDTB200_OK);
  auto lane_loop = [&](int l) {
    ndtb200_handle* h = lanes[l];
    cudaSetDevice(h->device);
    for (int k = l; k < n_pairs; k += n_lanes) {
      int st = ndtb200_set_target(h, targets[k], n_targets[k], stride_bytes, 1);
      if (st == NDTB200_OK) st = ndtb200_set_source(h, sources[k], n_sources[k], stride_bytes);
      if (st == NDTB200_OK) st = ndtb200_align_async(h, guesses16 ? guesses16 + 16 * (size_t)k : nullptr);
      if (st == NDTB200_OK) st = ndtb200_sync(h);  // the node reads every pose (yield-polling on throughput-mode handles)
      if (st == NDTB200_OK) st = ndtb200_get_result(h, results + k);
      if (st != NDTB200_OK) {
        std::memset(results + k, 0, sizeof(ndtb200_result));
        if (status[l] == NDTB200_OK) status[l] = st;
      }
    }
  };
  std::vector<std::thread> th;
  th.reserve(n_lanes);
  for (int l = 1; l < n_lanes; ++l) th.emplace_back(lane_loop, l);
  lane_loop(0);
  for (std::thread& t : th) t.join();
  for (int l = 0; l < n_lanes; ++l)
    if (status[l] != NDTB200_OK) return status[l];
  return NDTB200_OK;
}

void* ndtb200_stream(ndtb200_handle* h) { return h ? static_cast<void*>(h->stream) : nullptr; }
int64_t ndtb200_launch_count(const ndtb200_handle* h) { return h ? h->launches : 0; }
void ndtb200_reset_launch_count(ndtb200_handle* h) { if (h) h->launches = 0; }

int ndtb200_last_align_ms(ndtb200_handle* h, float* ms) {
  if (!h || !ms) return NDTB200_ERR_INVALID;
  cudaSetDevice(h->device);
  CK(cudaEventSynchronize(h->ev1));
  CK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  if (getenv("NDTB200_DEBUG_STEP") && h->result_valid)
    std::fprintf(stderr, "  kernel body (CTA 0, globaltimer): %.2f us\n",
                 (h->h_result->t_kernel_end - h->h_result->t_kernel_begin) * 1e-3);
  return NDTB200_OK;
}

}  // extern "C"
