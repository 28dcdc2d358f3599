// small_build.cuh — the whole target-map build (or the pcl::VoxelGrid downsample) of a SCAN-SIZED cloud as ONE
// persistent cooperative kernel.
//
// The staged pipeline of map_build.cuh costs ~27 launches + 2 host synchronisations per build; for a 30 k-point scan
// every one of its kernels is a few microseconds of work, so the batched-odometry pipeline (BASELINE configs[2]) and
// the mapping loop were bound by the host's launch rate, not by the GPU.  Here all phases run inside one launch,
// separated by a grid barrier (all CTAs co-resident: cooperative launch), and the host synchronises once, to read the
// grid description and the voxel count it needs for the index allocation:
//
//   bounding box -> grid description (every CTA derives the identical GridDesc from the CTA partials)
//   keys -> [ count | scatter ] x passes      (the same tile code as the staged path, 512-key tiles; the tile offsets
//           come from a tile-major histogram + digit totals, so no global scan phase is needed)
//   segment heads (count | scan (CTA 0) | write) -> voxel list
//   MODE 0: per-voxel fp64 moments + finalize (mean, covariance, eigen-regularisation, inverse) -> records
//   MODE 1: per-voxel fp32 centroid in input order (pcl::VoxelGrid)                              -> float4 cloud
//
// Same arithmetic and same order as the staged kernels: results are bit-identical (tested path against path).
//
// Two launch flavours (template parameter CLUSTER): a cooperative launch over up to all SMs with a counter barrier in
// global memory (latency-mode handles), or ONE thread-block cluster of <= 8 CTAs launched as an ordinary kernel with
// the hardware cluster barrier (throughput-mode handles, NDTB200_BUILD_PATH=fused) — see run_fused_build in capi.cu.
#pragma once
#include "map_build.cuh"

namespace ndtb200 {

constexpr int kSmallRounds = 2;                                 // 512-key sort tiles: a 30 k-point scan spreads over 59 CTAs
constexpr int kSmallTile = kBuildThreads * kSmallRounds;
constexpr size_t kSmallMaxPoints = 65536;                       // above this the staged streaming kernels win

struct SmallBuildArgs {
  const float4* pts;
  uint32_t n;
  int is_dense;
  Leaf3 leaf;
  int min_points;
  double eig_ratio;
  int mode;  // 0: NDT voxel map, 1: VoxelGrid centroids
  // scratch (sized by n)
  float* mm_partial;            // [grid][6]
  unsigned int* mm_finite;      // [grid]
  uint32_t *keys_a, *keys_b, *vals_a, *vals_b;
  uint32_t* hist;               // [ntiles][256] (tile-major)
  uint32_t* digit_totals;       // [4 passes][256], zeroed before the launch
  uint32_t* tile_heads;         // [ceil(n / kScanTile)]
  unsigned int* barrier;        // zeroed before the launch
  // outputs
  GridDesc* grid;
  uint32_t* n_vox;
  unsigned int* n_valid;        // zeroed before the launch
  int32_t* voxel_key;           // [n]
  uint32_t* voxel_start;        // [n]
  double* moments;              // [n][9]   (MODE 0)
  VoxelRecord* records;         // [n]      (MODE 0)
  double* icov64;               // [n][6]   (MODE 0)
  float4* centroids;            // [n]      (MODE 1)
  uint32_t** sorted_idx_out;    // receives which of vals_a / vals_b holds the sorted point indices
};

// CLUSTER: the grid is ONE thread-block cluster (8 portable / 16 CTAs, launched with cudaLaunchAttributeClusterDimension
// as an ordinary kernel): the phases are separated by the hardware cluster barrier (release / acquire at cluster scope)
// instead of a counter in global memory.  The hardware co-schedules a cluster, so many such builds run side by side
// (handles in throughput mode) without the co-residency requirement of a cooperative launch.
template <bool CLUSTER>
__device__ __forceinline__ void small_grid_sync(unsigned int* counter, unsigned int& phase) {
  if (CLUSTER) {
    __threadfence();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    ++phase;
    return;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    ++phase;
    atom_add_acq_rel_gpu(counter, 1u);
    const unsigned int target = phase * gridDim.x;
    while (ld_acquire_u32(counter) < target) {}
  } else {
    ++phase;
  }
  __syncthreads();
}

// in-place exclusive scan of data[0..n) by ONE CTA (256 threads, contiguous chunks); returns the total to all threads
__device__ __forceinline__ uint32_t cta_exclusive_scan_inplace(uint32_t* data, uint32_t n, uint32_t* s_scan) {
  const uint32_t per = (n + kBuildThreads - 1) / kBuildThreads;
  const uint32_t lo = min(n, threadIdx.x * per), hi = min(n, lo + per);
  uint32_t sum = 0;
  for (uint32_t i = lo; i < hi; ++i) sum += __ldcg(data + i);
  uint32_t total;
  uint32_t off = block_exclusive_scan(sum, s_scan, total);
  for (uint32_t i = lo; i < hi; ++i) {
    const uint32_t v = __ldcg(data + i);
    data[i] = off;
    off += v;
  }
  return total;
}

template <int MODE, bool CLUSTER = false>
__global__ void __launch_bounds__(kBuildThreads)
small_build_kernel(const SmallBuildArgs a) {
  __shared__ uint32_t warp_cnt[kSortWarps][256];
  __shared__ uint32_t s_dstart[256];
  __shared__ uint32_t s_gbase[256];
  __shared__ uint32_t s_scan[kBuildThreads / 32 + 1];
  __shared__ uint32_t s_key[kSmallTile];
  __shared__ uint32_t s_val[kSmallTile];
  __shared__ GridDesc g;
  __shared__ float s_mn[3][kBuildThreads / 32], s_mx[3][kBuildThreads / 32];
  __shared__ unsigned long long s_nf[kBuildThreads / 32];
  const uint32_t n = a.n;
  const int G = gridDim.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned int phase = 0;

  // ---- bounding box: per-CTA partials (pcl::getMinMax3D) ----
  {
    float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
    float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    unsigned int nf = 0;
    for (uint32_t i = blockIdx.x * kBuildThreads + threadIdx.x; i < n; i += G * kBuildThreads) {
      const float4 p = __ldg(a.pts + i);
      const bool ok = a.is_dense || (isfinite(p.x) && isfinite(p.y) && isfinite(p.z));
      if (ok) {
        mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
        mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
        mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
        ++nf;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
        mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
      }
      nf += __shfl_xor_sync(0xffffffffu, nf, o);
    }
    if (lane == 0) {
      for (int c = 0; c < 3; ++c) { s_mn[c][warp] = mn[c]; s_mx[c][warp] = mx[c]; }
      s_nf[warp] = nf;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long t = s_nf[0];
      for (int w = 1; w < kBuildThreads / 32; ++w) {
        for (int c = 0; c < 3; ++c) { mn[c] = fminf(mn[c], s_mn[c][w]); mx[c] = fmaxf(mx[c], s_mx[c][w]); }
        t += s_nf[w];
      }
      for (int c = 0; c < 3; ++c) { a.mm_partial[blockIdx.x * 6 + c] = mn[c]; a.mm_partial[blockIdx.x * 6 + 3 + c] = mx[c]; }
      a.mm_finite[blockIdx.x] = static_cast<unsigned int>(t);
    }
  }
  small_grid_sync<CLUSTER>(a.barrier, phase);

  // ---- grid description: every CTA reduces the partials itself (identical result everywhere) ----
  {
    float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
    float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    unsigned long long nf = 0;
    for (int b = threadIdx.x; b < G; b += kBuildThreads) {
      for (int c = 0; c < 3; ++c) { mn[c] = fminf(mn[c], __ldcg(a.mm_partial + b * 6 + c)); mx[c] = fmaxf(mx[c], __ldcg(a.mm_partial + b * 6 + 3 + c)); }
      nf += __ldcg(a.mm_finite + b);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
        mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
      }
      nf += __shfl_xor_sync(0xffffffffu, nf, o);
    }
    if (lane == 0) {
      for (int c = 0; c < 3; ++c) { s_mn[c][warp] = mn[c]; s_mx[c][warp] = mx[c]; }
      s_nf[warp] = nf;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < kBuildThreads / 32; ++w) {
        for (int c = 0; c < 3; ++c) { mn[c] = fminf(mn[c], s_mn[c][w]); mx[c] = fmaxf(mx[c], s_mx[c][w]); }
        nf += s_nf[w];
      }
      GridDesc gd;
      grid_from_box(mn, mx, nf, a.leaf, /*forced=*/false, gd);
      g = gd;
      if (blockIdx.x == 0) *a.grid = gd;
    }
    __syncthreads();
  }
  if (g.n_finite == 0 || g.overflow) {  // uniform: every CTA derived the same description
    if (blockIdx.x == 0 && threadIdx.x == 0) { *a.n_vox = 0; *a.sorted_idx_out = a.vals_a; }
    return;
  }

  // ---- keys ----
  unsigned long long key_space = (unsigned long long)g.div_b[0] * (unsigned long long)g.div_b[1] * (unsigned long long)g.div_b[2];
  if (key_space > 0xFFFFFFFEull) key_space = 0xFFFFFFFEull;
  const uint32_t sentinel = static_cast<uint32_t>(key_space);
  const bool with_sentinel = (MODE == 1) || !a.is_dense;
  const unsigned long long max_key = with_sentinel ? key_space : (key_space - 1);
  int bits = 1;
  while (bits < 32 && (max_key >> bits) != 0) ++bits;
  const int passes = (bits + 7) / 8;
  const bool check_finite = (MODE == 1) || !a.is_dense;
  for (uint32_t i = blockIdx.x * kBuildThreads + threadIdx.x; i < n; i += G * kBuildThreads) {
    const float4 p = __ldg(a.pts + i);
    const bool ok = !check_finite || (isfinite(p.x) && isfinite(p.y) && isfinite(p.z));
    a.keys_a[i] = ok ? static_cast<uint32_t>(voxel_key_of(p, g)) : sentinel;
  }
  small_grid_sync<CLUSTER>(a.barrier, phase);

  // ---- stable LSD radix sort of (key, point index) ----
  const int ntiles = static_cast<int>((n + kSmallTile - 1) / kSmallTile);
  uint32_t *ka = a.keys_a, *kb = a.keys_b, *va = a.vals_a, *vb = a.vals_b;
  for (int pass = 0; pass < passes; ++pass) {
    const int shift = pass * 8;
    uint32_t* totals = a.digit_totals + pass * 256;
    for (int t = blockIdx.x; t < ntiles; t += G) radix_count_tile<kSmallRounds>(ka, n, shift, a.hist, ntiles, t, warp_cnt, totals);
    small_grid_sync<CLUSTER>(a.barrier, phase);
    for (int t = blockIdx.x; t < ntiles; t += G)
      radix_scatter_tile<kSmallRounds>(ka, pass == 0 ? nullptr : va, n, shift, a.hist, ntiles, t, kb, vb, warp_cnt, s_dstart, s_gbase,
                                       s_scan, s_key, s_val, totals);
    small_grid_sync<CLUSTER>(a.barrier, phase);
    uint32_t* tk = ka; ka = kb; kb = tk;
    uint32_t* tv = va; va = vb; vb = tv;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.sorted_idx_out = va;

  // ---- segment heads -> voxel list ----
  const int stiles = static_cast<int>((n + kScanTile - 1) / kScanTile);
  for (int t = blockIdx.x; t < stiles; t += G) {
    const size_t base = (size_t)t * kScanTile + (size_t)threadIdx.x * kScanItems;
    uint32_t key[kScanItems];
    const uint32_t c = __popc(load_heads(ka, n, base, sentinel, key));
    uint32_t total;
    block_exclusive_scan(c, s_scan, total);
    if (threadIdx.x == 0) a.tile_heads[t] = total;
  }
  small_grid_sync<CLUSTER>(a.barrier, phase);
  if (blockIdx.x == 0) {
    const uint32_t total = cta_exclusive_scan_inplace(a.tile_heads, static_cast<uint32_t>(stiles), s_scan);
    if (threadIdx.x == 0) *a.n_vox = total;
  }
  small_grid_sync<CLUSTER>(a.barrier, phase);
  const uint32_t n_vox = __ldcg(a.n_vox);
  for (int t = blockIdx.x; t < stiles; t += G) {
    const size_t base = (size_t)t * kScanTile + (size_t)threadIdx.x * kScanItems;
    uint32_t key[kScanItems];
    const uint32_t flags = load_heads(ka, n, base, sentinel, key);
    uint32_t total;
    uint32_t off = block_exclusive_scan(__popc(flags), s_scan, total) + __ldcg(a.tile_heads + t);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      if (flags & (1u << k)) {
        a.voxel_key[off] = static_cast<int32_t>(key[k]);
        a.voxel_start[off] = static_cast<uint32_t>(base + k);
        ++off;
      }
    }
  }
  small_grid_sync<CLUSTER>(a.barrier, phase);

  // ---- per voxel ----
  const uint32_t n_finite = static_cast<uint32_t>(g.n_finite);
  if (MODE == 1) {
    for (uint32_t v = blockIdx.x * kBuildThreads + threadIdx.x; v < n_vox; v += G * kBuildThreads) {
      const uint32_t b = __ldcg(a.voxel_start + v);
      const uint32_t e = (v + 1 < n_vox) ? __ldcg(a.voxel_start + v + 1) : n_finite;
      float sx = 0.f, sy = 0.f, sz = 0.f;
      for (uint32_t i = b; i < e; ++i) {
        const float4 p = __ldg(a.pts + __ldcg(va + i));
        sx = __fadd_rn(sx, p.x);
        sy = __fadd_rn(sy, p.y);
        sz = __fadd_rn(sz, p.z);
      }
      const float cnt = static_cast<float>(e - b);
      a.centroids[v] = make_float4(__fdiv_rn(sx, cnt), __fdiv_rn(sy, cnt), __fdiv_rn(sz, cnt), 1.0f);
    }
  } else {
    // the staged path picks 4 / 8 / 32 lanes per voxel by mean occupancy (the fixed shuffle tree is part of the
    // arithmetic): same rule here
    const double avg = static_cast<double>(n_finite) / static_cast<double>(n_vox > 0 ? n_vox : 1u);
    const int group = avg >= 48.0 ? 32 : (avg >= 10.0 ? 8 : 4);
    const uint32_t groups_per_cta = kBuildThreads / group;
    const int gl = threadIdx.x % group;
    for (uint32_t v0 = blockIdx.x * groups_per_cta; v0 < n_vox; v0 += G * groups_per_cta) {  // CTA-uniform trip count
      const uint32_t v = v0 + threadIdx.x / group;
      const bool active = v < n_vox;
      double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      uint32_t b = 0, e = 0;
      if (active) {
        b = __ldcg(a.voxel_start + v);
        e = (v + 1 < n_vox) ? __ldcg(a.voxel_start + v + 1) : n_finite;
        for (uint32_t i0 = b + gl; i0 < e; i0 += 4 * group) {
          uint32_t idx[4];
          float4 p[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) idx[u] = (i0 + u * group < e) ? __ldcg(va + i0 + u * group) : 0xffffffffu;
#pragma unroll
          for (int u = 0; u < 4; ++u) p[u] = (idx[u] != 0xffffffffu) ? ldg_gather16(a.pts + idx[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (idx[u] == 0xffffffffu) continue;
            const double x = p[u].x, y = p[u].y, z = p[u].z;
            s[0] += x; s[1] += y; s[2] += z;
            s[3] += x * x; s[4] += x * y; s[5] += x * z;
            s[6] += y * y; s[7] += y * z; s[8] += z * z;
          }
        }
      }
      for (int o = group / 2; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 9; ++k) s[k] += __shfl_down_sync(0xffffffffu, s[k], o, group);
      }
      if (active && gl == 0) {
#pragma unroll
        for (int k = 0; k < 9; ++k) a.moments[(size_t)v * 9 + k] = s[k];  // kept for ndtb200_dump_voxels
        finalize_one(v, static_cast<int>(e - b), s, __ldcg(a.voxel_key + v), a.min_points, a.eig_ratio, a.records, a.icov64, a.n_valid,
                     nullptr, nullptr, nullptr, nullptr);
      }
    }
  }
}

}  // namespace ndtb200
