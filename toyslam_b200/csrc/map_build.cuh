// map_build.cuh — target-map build: pclomp::VoxelGridCovariance::applyFilter
// (ndt_omp/include/pclomp/voxel_grid_covariance_omp_impl.hpp:48-370) re-designed for B200:
//
//   minmax3d            -> bounding box (pcl::getMinMax3D, …_impl.hpp:72)
//   grid_setup          -> min_b/max_b/div_b/divb_mul + int32 overflow guard (…_impl.hpp:75-103)
//   voxel_key           -> per-point int32 key, bit-exact fp32 arithmetic (…_impl.hpp:218-223), and the 8-bit digit
//                          histograms of all radix passes at once
//   radix sort          -> hand-written stable LSD radix sort, 8-bit digits, ONE-SWEEP passes (decoupled look-back):
//                          (key, point index) pairs (onesweep_kernel), or — clouds of 4 M points and more — the points
//                          themselves with the key in .w (onesweep_payload_kernel), so that the moments read the sorted
//                          cloud sequentially instead of gathering 64 B of DRAM per point
//   segment heads       -> occupied-voxel list (= leaves_), per-voxel point ranges
//   voxel_build         -> per-voxel count, sum x, sum x x^T in fp64 (first pass, …_impl.hpp:233-262), then mean,
//                          covariance, eigen-regularisation, inverse (second pass, …_impl.hpp:282-367) and the cell-table
//                          entry, in one kernel (voxel_moments / finalize_voxels: the same two halves as separate
//                          kernels, used by the sharded build and the parity dumps)
//   hash_insert         -> open-addressing HBM hash over the valid voxels (grids whose cell table does not fit)
//
// All kernels are HBM / issue bound streaming kernels: no tensor cores on this path.  DESIGN.md §4.1 has the per-kernel
// times and the reasoning for what bounds them.
#pragma once
#include "common.cuh"

namespace ndtb200 {

constexpr int kBuildThreads = 256;

// ---------------------------------------------------------------------------------------------
// bounding box
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBuildThreads)
minmax3d_kernel(const float4* __restrict__ pts, size_t n, int is_dense, float* __restrict__ partial,
                unsigned int* __restrict__ finite_partial) {
  float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
  float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
  unsigned int nf = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float4 p = __ldg(pts + i);
    bool ok = is_dense || (isfinite(p.x) && isfinite(p.y) && isfinite(p.z));
    if (ok) {
      mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
      mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
      mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
      ++nf;
    }
  }
  __shared__ float s_mn[3][kBuildThreads / 32], s_mx[3][kBuildThreads / 32];
  __shared__ unsigned int s_nf[kBuildThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
    nf += __shfl_xor_sync(0xffffffffu, nf, o);
  }
  if (lane == 0) {
    for (int a = 0; a < 3; ++a) { s_mn[a][warp] = mn[a]; s_mx[a][warp] = mx[a]; }
    s_nf[warp] = nf;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kBuildThreads / 32; ++w) {
      for (int a = 0; a < 3; ++a) { mn[a] = fminf(mn[a], s_mn[a][w]); mx[a] = fmaxf(mx[a], s_mx[a][w]); }
      nf += s_nf[w];
    }
    for (int a = 0; a < 3; ++a) { partial[blockIdx.x * 6 + a] = mn[a]; partial[blockIdx.x * 6 + 3 + a] = mx[a]; }
    finite_partial[blockIdx.x] = nf;
  }
}

// VoxelGrid members from the bounding box (…_impl.hpp:75-103): inverse leaf, int64 overflow guard expression, min_b /
// max_b / div_b / divb_mul — fp32 arithmetic exactly as the reference
struct Leaf3 { float v[3]; };  // pcl::VoxelGrid::setLeafSize(lx, ly, lz); the NDT grid is cubic (resolution)

__device__ __forceinline__ void grid_from_box(const float (&mn)[3], const float (&mx)[3], unsigned long long nf, const Leaf3& leaf, bool forced,
                                              GridDesc& g) {
  long long d[3];
  for (int a = 0; a < 3; ++a) {
    const float inv = __fdiv_rn(1.0f, leaf.v[a]);  // pcl::VoxelGrid::setLeafSize: inverse_leaf_size_ = 1.0f / leaf (per axis)
    g.leaf[a] = leaf.v[a];
    g.inv_leaf[a] = inv;
    g.min_p[a] = mn[a];
    g.max_p[a] = mx[a];
    // …_impl.hpp:75-77  int64((max - min) * inv_leaf) + 1, fp32 arithmetic
    d[a] = static_cast<long long>(__fmul_rn(__fsub_rn(mx[a], mn[a]), inv)) + 1;
    // …_impl.hpp:87-92
    g.min_b[a] = static_cast<int>(floorf(__fmul_rn(mn[a], inv)));
    g.max_b[a] = static_cast<int>(floorf(__fmul_rn(mx[a], inv)));
    g.div_b[a] = g.max_b[a] - g.min_b[a] + 1;
  }
  g.ncell = d[0] * d[1] * d[2];
  g.overflow = ((nf > 0 || forced) && g.ncell > 2147483647ll) ? 1 : 0;
  g.mul[0] = 1;
  g.mul[1] = g.div_b[0];
  g.mul[2] = g.div_b[0] * g.div_b[1];
  g.n_finite = static_cast<int>(nf);
}

struct ForcedBox {  // sharded build: every rank keys its slice with the bounding box of the WHOLE cloud
  int use;
  float mn[3], mx[3];
};

__global__ void __launch_bounds__(kBuildThreads)
grid_setup_kernel(const float* __restrict__ partial, const unsigned int* __restrict__ finite_partial,
                  int nblocks, Leaf3 leaf, ForcedBox forced, GridDesc* __restrict__ out) {
  // one CTA: strided reduction of the per-CTA partials, then thread 0 derives the grid description
  __shared__ float s_mn[3][kBuildThreads / 32], s_mx[3][kBuildThreads / 32];
  __shared__ unsigned long long s_nf[kBuildThreads / 32];
  float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
  float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
  unsigned long long nf = 0;
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
    for (int a = 0; a < 3; ++a) { mn[a] = fminf(mn[a], partial[b * 6 + a]); mx[a] = fmaxf(mx[a], partial[b * 6 + 3 + a]); }
    nf += finite_partial[b];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
    nf += __shfl_xor_sync(0xffffffffu, nf, o);
  }
  if ((threadIdx.x & 31) == 0) {
    for (int a = 0; a < 3; ++a) { s_mn[a][threadIdx.x >> 5] = mn[a]; s_mx[a][threadIdx.x >> 5] = mx[a]; }
    s_nf[threadIdx.x >> 5] = nf;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  for (int w = 1; w < kBuildThreads / 32; ++w) {
    for (int a = 0; a < 3; ++a) { mn[a] = fminf(mn[a], s_mn[a][w]); mx[a] = fmaxf(mx[a], s_mx[a][w]); }
    nf += s_nf[w];
  }
  if (forced.use) {
    for (int a = 0; a < 3; ++a) { mn[a] = forced.mn[a]; mx[a] = forced.mx[a]; }
  }
  GridDesc g;
  grid_from_box(mn, mx, nf, leaf, forced.use != 0, g);
  *out = g;
}

// ---------------------------------------------------------------------------------------------
// keys  (…_impl.hpp:218-223) — bit-exact: fp32 multiply by inv_leaf, floor, fp32 subtract of min_b
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int32_t voxel_key_of(const float4& p, const GridDesc& g) {
  int ijk0 = static_cast<int>(__fsub_rn(floorf(__fmul_rn(p.x, g.inv_leaf[0])), static_cast<float>(g.min_b[0])));
  int ijk1 = static_cast<int>(__fsub_rn(floorf(__fmul_rn(p.y, g.inv_leaf[1])), static_cast<float>(g.min_b[1])));
  int ijk2 = static_cast<int>(__fsub_rn(floorf(__fmul_rn(p.z, g.inv_leaf[2])), static_cast<float>(g.min_b[2])));
  return ijk0 * g.mul[0] + ijk1 * g.mul[1] + ijk2 * g.mul[2];
}

// Digit histograms of ALL radix passes at once (one-sweep sort: the per-pass count kernels and their scans go away).
// s_hist: [kMaxSortPasses][256] shared counters; a warp whose 32 keys share the digit (clouds in scan order, and always
// the high digits of coherent clouds) adds once.
constexpr int kMaxSortPasses = 4;

__device__ __forceinline__ void digit_hist_add(uint32_t (*s_hist)[256], uint32_t key, bool valid, int passes) {
  for (int p = 0; p < passes; ++p) {
    const uint32_t d = (key >> (8 * p)) & 255u;
    const uint32_t d0 = __shfl_sync(0xffffffffu, d, 0);
    const unsigned int act = __ballot_sync(0xffffffffu, valid);
    if (__all_sync(0xffffffffu, !valid || d == d0)) {
      if ((threadIdx.x & 31) == 0 && act) atomicAdd(&s_hist[p][d0], static_cast<uint32_t>(__popc(act)));
    } else if (valid) {
      atomicAdd(&s_hist[p][d], 1u);
    }
  }
}

// keys[i] = voxel key of point i (sentinel for a skipped non-finite point); digit_hist != nullptr: also accumulate the
// 8-bit digit histograms of the `passes` radix passes into digit_hist[pass][256] (zeroed by the caller)
__global__ void __launch_bounds__(kBuildThreads)
voxel_key_kernel(const float4* __restrict__ pts, size_t n, int is_dense, const GridDesc* __restrict__ gd,
                 uint32_t sentinel, uint32_t* __restrict__ keys, uint32_t* __restrict__ idx,
                 unsigned long long* __restrict__ digit_hist, int passes) {
  __shared__ GridDesc g;
  __shared__ uint32_t s_hist[kMaxSortPasses][256];
  if (threadIdx.x == 0) g = *gd;
  if (digit_hist)
    for (int i = threadIdx.x; i < kMaxSortPasses * 256; i += kBuildThreads) (&s_hist[0][0])[i] = 0u;
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t n_round = ((n + 31) / 32) * 32;  // whole warps stay in the loop (the histogram uses warp votes)
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_round; i += stride) {
    const bool in = i < n;
    uint32_t key = sentinel;
    if (in) {
      float4 p = __ldg(pts + i);
      bool ok = is_dense || (isfinite(p.x) && isfinite(p.y) && isfinite(p.z));
      if (ok) key = static_cast<uint32_t>(voxel_key_of(p, g));
      if (keys) keys[i] = key;  // nullptr: histograms only (the payload sort computes the keys again in its first pass)
      if (idx) idx[i] = static_cast<uint32_t>(i);
    }
    if (digit_hist) digit_hist_add(s_hist, key, in, passes);
  }
  if (digit_hist) {
    __syncthreads();
    for (int i = threadIdx.x; i < passes * 256; i += kBuildThreads) {
      const uint32_t c = (&s_hist[0][0])[i];
      if (c) atomicAdd(digit_hist + i, static_cast<unsigned long long>(c));
    }
  }
}

// the same histograms for keys that already exist (merged partials, VoxelGrid filter)
__global__ void __launch_bounds__(kBuildThreads)
digit_hist_kernel(const uint32_t* __restrict__ keys, size_t n, unsigned long long* __restrict__ digit_hist, int passes) {
  __shared__ uint32_t s_hist[kMaxSortPasses][256];
  for (int i = threadIdx.x; i < kMaxSortPasses * 256; i += kBuildThreads) (&s_hist[0][0])[i] = 0u;
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t n_round = ((n + 31) / 32) * 32;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_round; i += stride) {
    const bool in = i < n;
    const uint32_t key = in ? __ldg(keys + i) : 0u;
    digit_hist_add(s_hist, key, in, passes);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < passes * 256; i += kBuildThreads) {
    const uint32_t c = (&s_hist[0][0])[i];
    if (c) atomicAdd(digit_hist + i, static_cast<unsigned long long>(c));
  }
}

// exclusive scan of every pass's 256 digit totals -> first output position of each digit; one CTA of 256 threads
__global__ void __launch_bounds__(256)
digit_base_kernel(const unsigned long long* __restrict__ digit_hist, int passes, unsigned long long* __restrict__ digit_base) {
  __shared__ unsigned long long s[256];
  for (int p = 0; p < passes; ++p) {
    const unsigned long long c = digit_hist[p * 256 + threadIdx.x];
    s[threadIdx.x] = c;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
      const unsigned long long t = (threadIdx.x >= o) ? s[threadIdx.x - o] : 0ull;
      __syncthreads();
      s[threadIdx.x] += t;
      __syncthreads();
    }
    digit_base[p * 256 + threadIdx.x] = s[threadIdx.x] - c;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// generic exclusive scan over uint32 (3-kernel, recursive)
// ---------------------------------------------------------------------------------------------
constexpr int kScanItems = 8;
constexpr int kScanTile = kBuildThreads * kScanItems;

__global__ void __launch_bounds__(kBuildThreads)
scan_tiles_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t n,
                  uint32_t* __restrict__ tile_sums) {
  __shared__ uint32_t smem[kBuildThreads / 32 + 1];
  const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0u;
    sum += v[k];
  }
  uint32_t total;
  uint32_t off = block_exclusive_scan(sum, smem, total);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n) out[base + k] = off;
    off += v[k];
  }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kBuildThreads)
scan_add_kernel(uint32_t* __restrict__ out, size_t n, const uint32_t* __restrict__ tile_offsets) {
  const uint32_t add = tile_offsets[blockIdx.x];
  const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (base + k < n) out[base + k] += add;
}

// ---------------------------------------------------------------------------------------------
// stable LSD radix sort pass, 8-bit digit.  Tile = 8 warps x 16 rounds x 32 keys = 4096 keys; warp w owns 512
// consecutive keys, lane = consecutive key, so global loads are fully coalesced and (warp, round, lane) order is the
// input order (stability).
//   * peers of a key (lanes of the round with the same digit) come from 8 ballots — __match_any_sync costs ~64 issue
//     cycles per warp on this part, the ballots ~25;
//   * count pass: per-warp private digit counters, the lowest peer adds popc(peers): no atomics, no ranks;
//   * scatter pass: the tile is first sorted by digit INSIDE shared memory (stable local ranks), then written out:
//     consecutive threads write consecutive addresses of a digit's run (>= 64 B runs on average) instead of 4-byte
//     scattered stores.
// ---------------------------------------------------------------------------------------------
#ifndef NDTB200_SORT_ROUNDS
#define NDTB200_SORT_ROUNDS 16
#endif
#ifndef NDTB200_SCATTER_MIN_BLOCKS
#define NDTB200_SCATTER_MIN_BLOCKS 3
#endif
constexpr int kSortRounds = NDTB200_SORT_ROUNDS;
constexpr int kSortWarps = kBuildThreads / 32;
constexpr int kSortTile = kBuildThreads * kSortRounds;

__device__ __forceinline__ unsigned int digit_peers(uint32_t d, bool valid) {
  unsigned int m = __ballot_sync(0xffffffffu, valid);
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const bool bit = (d >> b) & 1u;
    const unsigned int bal = __ballot_sync(0xffffffffu, bit);
    m &= bit ? bal : ~bal;
  }
  return m;
}

// Count pass: only the per-tile digit totals are needed (no ranks), so shared-memory atomics do: a round whose 32 keys
// share one digit (clouds in scan order) is added by one lane, otherwise every lane adds 1 to its warp's private bin.
// Tile `tile` of ROUNDS x 256 keys; called by all 256 threads of a CTA.  warp_cnt: [kSortWarps][256] shared words.
// digit_totals != nullptr (fused small build): the tile's counts go to hist[tile][digit] (tile-major) and are added to
// digit_totals[digit]; otherwise hist[digit][tile] for the staged path's global scan.
template <int ROUNDS>
__device__ __forceinline__ void radix_count_tile(const uint32_t* __restrict__ keys, size_t n, int shift, uint32_t* __restrict__ hist,
                                                 int ntiles, int tile, uint32_t (*warp_cnt)[256], uint32_t* digit_totals = nullptr) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = threadIdx.x; d < 256 * kSortWarps; d += kBuildThreads) (&warp_cnt[0][0])[d] = 0u;
  __syncthreads();
  const size_t wbase = (size_t)tile * (kBuildThreads * ROUNDS) + (size_t)warp * (32 * ROUNDS);
  uint32_t k[ROUNDS];
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const size_t i = wbase + r * 32 + lane;
    k[r] = (i < n) ? __ldcg(keys + i) : 0u;  // L2-coherent: inside the fused small build the keys were written by other CTAs
  }
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const size_t i = wbase + r * 32 + lane;
    const bool valid = i < n;
    const uint32_t d = (k[r] >> shift) & 255u;
    const uint32_t d0 = __shfl_sync(0xffffffffu, d, 0);
    if (__all_sync(0xffffffffu, valid && d == d0)) {
      if (lane == 0) atomicAdd(&warp_cnt[warp][d0], 32u);
    } else if (valid) {
      atomicAdd(&warp_cnt[warp][d], 1u);
    }
  }
  __syncthreads();
  const int d = threadIdx.x;  // blockDim == 256
  uint32_t sum = 0;
#pragma unroll
  for (int w = 0; w < kSortWarps; ++w) sum += warp_cnt[w][d];
  if (digit_totals) {
    hist[(size_t)tile * 256 + d] = sum;
    if (sum) atomicAdd(digit_totals + d, sum);
  } else {
    hist[(size_t)d * ntiles + tile] = sum;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kBuildThreads)
radix_count_kernel(const uint32_t* __restrict__ keys, size_t n, int shift, uint32_t* __restrict__ hist, int ntiles) {
  __shared__ uint32_t warp_cnt[kSortWarps][256];
  radix_count_tile<kSortRounds>(keys, n, shift, hist, ntiles, blockIdx.x, warp_cnt);
}

// Scatter pass of one tile (see the header comment).  Shared memory: warp_cnt [kSortWarps][256], s_dstart / s_gbase
// [256], s_scan [9], s_key / s_val [ROUNDS * 256].
template <int ROUNDS>
__device__ __forceinline__ void radix_scatter_tile(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, size_t n,
                                                   int shift, const uint32_t* __restrict__ hist_scanned, int ntiles, int tile,
                                                   uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                   uint32_t (*warp_cnt)[256], uint32_t* s_dstart, uint32_t* s_gbase, uint32_t* s_scan,
                                                   uint32_t* s_key, uint32_t* s_val, const uint32_t* digit_totals = nullptr) {
  constexpr int kTile = kBuildThreads * ROUNDS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = threadIdx.x; d < 256 * kSortWarps; d += kBuildThreads) (&warp_cnt[0][0])[d] = 0u;
  __syncthreads();
  const size_t tbase = (size_t)tile * kTile;
  const size_t wbase = tbase + (size_t)warp * (32 * ROUNDS);
  uint32_t k[ROUNDS], v[ROUNDS];
  uint16_t rank[ROUNDS];
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const size_t i = wbase + r * 32 + lane;
    k[r] = (i < n) ? __ldcg(keys_in + i) : 0u;
    v[r] = (i < n) ? (vals_in ? __ldcg(vals_in + i) : static_cast<uint32_t>(i)) : 0u;
  }
  // stable rank of every key among the keys of ITS WARP with the same digit
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const size_t i = wbase + r * 32 + lane;
    const bool valid = i < n;
    const uint32_t d = (k[r] >> shift) & 255u;
    const unsigned int m = digit_peers(d, valid);
    uint32_t pos = 0;
    if (valid) pos = warp_cnt[warp][d] + __popc(m & ((1u << lane) - 1u));
    __syncwarp();
    if (valid && lane == (__ffs(m) - 1)) warp_cnt[warp][d] += __popc(m);
    __syncwarp();
    rank[r] = static_cast<uint16_t>(pos);
  }
  __syncthreads();
  // per digit: exclusive prefix over the warps, then over the digits (tile-local sorted layout)
  {
    const int d = threadIdx.x;
    uint32_t off = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const uint32_t c = warp_cnt[w][d];
      warp_cnt[w][d] = off;
      off += c;
    }
    uint32_t total;
    const uint32_t dstart = block_exclusive_scan(off, s_scan, total);
    s_dstart[d] = dstart;
    uint32_t gpos;
    if (digit_totals) {
      // fused small build: no global scan — position of (digit d, this tile) = keys with a smaller digit (exclusive
      // scan of the digit totals) + keys with digit d in earlier tiles (tile-major histogram: coalesced, 8 loads in flight)
      uint32_t before = 0;
      int t = 0;
      for (; t + 8 <= tile; t += 8) {
        uint32_t c[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) c[u] = __ldcg(hist_scanned + (size_t)(t + u) * 256 + d);
#pragma unroll
        for (int u = 0; u < 8; ++u) before += c[u];
      }
      for (; t < tile; ++t) before += __ldcg(hist_scanned + (size_t)t * 256 + d);
      uint32_t tot_all;
      gpos = block_exclusive_scan(__ldcg(digit_totals + d), s_scan, tot_all) + before;
    } else {
      gpos = __ldcg(hist_scanned + (size_t)d * ntiles + tile);
    }
    s_gbase[d] = gpos - dstart;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const size_t i = wbase + r * 32 + lane;
    if (i < n) {
      const uint32_t d = (k[r] >> shift) & 255u;
      const uint32_t p = s_dstart[d] + warp_cnt[warp][d] + rank[r];
      s_key[p] = k[r];
      s_val[p] = v[r];
    }
  }
  __syncthreads();
  const uint32_t count = static_cast<uint32_t>(n - tbase < (size_t)kTile ? n - tbase : (size_t)kTile);
  for (uint32_t j = threadIdx.x; j < count; j += kBuildThreads) {
    const uint32_t key = s_key[j];
    const uint32_t pos = s_gbase[(key >> shift) & 255u] + j;
    NDT_CHECK(pos < n);
    keys_out[pos] = key;
    vals_out[pos] = s_val[j];
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kBuildThreads, NDTB200_SCATTER_MIN_BLOCKS)
radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, size_t n, int shift,
                     const uint32_t* __restrict__ hist_scanned, int ntiles, uint32_t* __restrict__ keys_out,
                     uint32_t* __restrict__ vals_out) {
  __shared__ uint32_t warp_cnt[kSortWarps][256];
  __shared__ uint32_t s_dstart[256];
  __shared__ uint32_t s_gbase[256];
  __shared__ uint32_t s_scan[kBuildThreads / 32 + 1];
  __shared__ uint32_t s_key[kSortTile];
  __shared__ uint32_t s_val[kSortTile];
  radix_scatter_tile<kSortRounds>(keys_in, vals_in, n, shift, hist_scanned, ntiles, blockIdx.x, keys_out, vals_out, warp_cnt, s_dstart,
                                  s_gbase, s_scan, s_key, s_val);
}

// ---------------------------------------------------------------------------------------------
// ONE-SWEEP pass: the scatter kernel finds its own global offsets.  Every tile publishes its 256 digit counts in a
// status row and obtains, per digit, the number of equal digits in all earlier tiles by decoupled look-back (walk the
// earlier tiles' rows backwards, adding "aggregate" words until an "inclusive prefix" word is found); the first output
// position of every digit comes from the up-front histograms (digit_base).  Per pass the keys are read ONCE and no
// count / scan kernels run.  Tiles are numbered in the order CTAs START (atomic ticket), so a tile only ever waits for
// tiles that are already running.  status word: bits 63..62 = 0 not ready / 1 aggregate / 2 inclusive prefix.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(kBuildThreads, NDTB200_SCATTER_MIN_BLOCKS)
onesweep_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, size_t n, int shift,
                const unsigned long long* __restrict__ digit_base, unsigned long long* __restrict__ status,
                unsigned int* __restrict__ ticket, uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
  constexpr int ROUNDS = kSortRounds;
  constexpr int kTile = kBuildThreads * ROUNDS;
  __shared__ uint32_t warp_cnt[kSortWarps][256];
  __shared__ uint32_t s_dstart[256];
  __shared__ unsigned long long s_gbase[256];
  __shared__ uint32_t s_scan[kBuildThreads / 32 + 1];
  __shared__ uint32_t s_key[kTile];
  __shared__ uint32_t s_val[kTile];
  __shared__ unsigned int s_tile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  for (int d = threadIdx.x; d < 256 * kSortWarps; d += kBuildThreads) (&warp_cnt[0][0])[d] = 0u;
  __syncthreads();
  const unsigned int tile = s_tile;
  NDT_CHECK((size_t)tile * kTile < n);
  const size_t tbase = (size_t)tile * kTile;
  const size_t wbase = tbase + (size_t)warp * (32 * ROUNDS);
  uint32_t k[ROUNDS], v[ROUNDS];
  uint16_t rank[ROUNDS];
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const size_t i = wbase + r * 32 + lane;
    k[r] = (i < n) ? __ldcs(keys_in + i) : 0u;   // streaming: every element is touched exactly once per pass
    v[r] = (i < n) ? (vals_in ? __ldcs(vals_in + i) : static_cast<uint32_t>(i)) : 0u;
  }
  // stable rank of every key among the keys of ITS WARP with the same digit
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const size_t i = wbase + r * 32 + lane;
    const bool valid = i < n;
    const uint32_t d = (k[r] >> shift) & 255u;
    const unsigned int m = digit_peers(d, valid);
    uint32_t pos = 0;
    if (valid) pos = warp_cnt[warp][d] + __popc(m & ((1u << lane) - 1u));
    __syncwarp();
    if (valid && lane == (__ffs(m) - 1)) warp_cnt[warp][d] += __popc(m);
    __syncwarp();
    rank[r] = static_cast<uint16_t>(pos);
  }
  __syncthreads();
  {
    const int d = threadIdx.x;  // one digit per thread (blockDim == 256)
    uint32_t off = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const uint32_t c = warp_cnt[w][d];
      warp_cnt[w][d] = off;
      off += c;
    }
    // publish this tile's count, then look back
    unsigned long long* row = status + (size_t)tile * 256;
    const unsigned long long kMask = (1ull << 62) - 1ull;
    unsigned long long excl = 0;
    if (tile == 0) {
      st_relaxed_u64(row + d, (2ull << 62) | off);
    } else {
      st_relaxed_u64(row + d, (1ull << 62) | off);
      long long t = static_cast<long long>(tile) - 1;
      while (true) {
        const unsigned long long w = ld_relaxed_u64(status + (size_t)t * 256 + d);
        const unsigned int flag = static_cast<unsigned int>(w >> 62);
        if (flag == 0u) continue;
        excl += w & kMask;
        if (flag == 2u) break;
        NDT_CHECK(t > 0);
        --t;
      }
      st_relaxed_u64(row + d, (2ull << 62) | (excl + off));
    }
    uint32_t total;
    const uint32_t dstart = block_exclusive_scan(off, s_scan, total);
    s_dstart[d] = dstart;
    s_gbase[d] = digit_base[d] + excl - dstart;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const size_t i = wbase + r * 32 + lane;
    if (i < n) {
      const uint32_t d = (k[r] >> shift) & 255u;
      const uint32_t p = s_dstart[d] + warp_cnt[warp][d] + rank[r];
      s_key[p] = k[r];
      s_val[p] = v[r];
    }
  }
  __syncthreads();
  const uint32_t count = static_cast<uint32_t>(n - tbase < (size_t)kTile ? n - tbase : (size_t)kTile);
  for (uint32_t j = threadIdx.x; j < count; j += kBuildThreads) {
    const uint32_t key = s_key[j];
    const unsigned long long pos = s_gbase[(key >> shift) & 255u] + j;
    NDT_CHECK(pos < n && s_val[j] < n);
    keys_out[pos] = key;
    vals_out[pos] = s_val[j];
  }
}

// ---------------------------------------------------------------------------------------------
// ONE-SWEEP pass that carries the POINT as the payload (clouds above kPayloadSortMin points).  Sorting (key, index)
// pairs leaves the per-voxel moments with a random 16-byte gather that drags 64 B out of DRAM per point (2.0 ms of a
// 5.8 ms build at 100 M points); here every pass moves the float4 itself with the key in .w (32 B of sequential traffic
// per point and pass), so the moments read the sorted cloud front to back.  The ranking work per key is the same as in
// onesweep_kernel (8 ballots, the issue-bound part), tiles are 2048 points (the payload lives in registers: 8 x float4).
//   FROM_CLOUD: first pass — reads the caller's cloud and computes the key on the fly (the up-front digit histograms
//   came from voxel_key_kernel with keys == nullptr); otherwise the key is the .w of the previous pass's output.
//   keys_out != nullptr (last pass): the sorted keys also go to a plain uint32 array for the segment-head kernels.
// status words are 32-bit (flag in bits 31..30, counts < 2^30): the host takes this path only for n < 2^30.
// ---------------------------------------------------------------------------------------------
#ifndef NDTB200_LOOKBACK_SLEEP_NS
#define NDTB200_LOOKBACK_SLEEP_NS 40
#endif
constexpr int kPayRounds = 8;
constexpr int kPayTile = kBuildThreads * kPayRounds;

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <bool FROM_CLOUD>
__global__ void __launch_bounds__(kBuildThreads, 4)
onesweep_payload_kernel(const float4* __restrict__ in, uint32_t n, int is_dense, const GridDesc* __restrict__ gd, uint32_t sentinel,
                        int shift, const unsigned long long* __restrict__ digit_base, uint32_t* __restrict__ status,
                        unsigned int* __restrict__ ticket, float4* __restrict__ out, uint32_t* __restrict__ keys_out) {
  constexpr int ROUNDS = kPayRounds;
  constexpr int kTile = kPayTile;
  __shared__ uint32_t warp_cnt[kSortWarps][256];
  __shared__ uint32_t s_dstart[256];
  __shared__ uint32_t s_gbase[256];
  __shared__ uint32_t s_scan[kBuildThreads / 32 + 1];
  __shared__ float4 s_pay[kTile];
  __shared__ unsigned int s_tile;
  __shared__ GridDesc g;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  if (FROM_CLOUD && threadIdx.x == 32) g = *gd;
  for (int d = threadIdx.x; d < 256 * kSortWarps; d += kBuildThreads) (&warp_cnt[0][0])[d] = 0u;
  __syncthreads();
  const unsigned int tile = s_tile;
  const uint32_t tbase = tile * (uint32_t)kTile;
  NDT_CHECK(tbase < n);
  const uint32_t wbase = tbase + (uint32_t)warp * (32 * ROUNDS);
  float4 p[ROUNDS];
  uint16_t rank[ROUNDS];
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const uint32_t i = wbase + r * 32 + lane;
    p[r] = (i < n) ? __ldcs(in + i) : make_float4(0.f, 0.f, 0.f, 0.f);  // streaming: every point is touched once per pass
  }
  if (FROM_CLOUD) {
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      const bool ok = is_dense || (isfinite(p[r].x) && isfinite(p[r].y) && isfinite(p[r].z));
      p[r].w = __uint_as_float(ok ? static_cast<uint32_t>(voxel_key_of(p[r], g)) : sentinel);
    }
  }
  // stable rank of every point among the points of ITS WARP with the same digit
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const uint32_t i = wbase + r * 32 + lane;
    const bool valid = i < n;
    const uint32_t d = (__float_as_uint(p[r].w) >> shift) & 255u;
    const unsigned int m = digit_peers(d, valid);
    uint32_t pos = 0;
    if (valid) pos = warp_cnt[warp][d] + __popc(m & ((1u << lane) - 1u));
    __syncwarp();
    if (valid && lane == (__ffs(m) - 1)) warp_cnt[warp][d] += __popc(m);
    __syncwarp();
    rank[r] = static_cast<uint16_t>(pos);
  }
  __syncthreads();
  {
    const int d = threadIdx.x;  // one digit per thread (blockDim == 256)
    uint32_t off = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const uint32_t c = warp_cnt[w][d];
      warp_cnt[w][d] = off;
      off += c;
    }
    // publish this tile's count right away; the look-back itself waits until the tile has been sorted into shared
    // memory (the predecessors get that much more time to publish: 40 % of all stall samples sat in the spin below
    // when the look-back came first, profiles/r02_build_kernels.md)
    uint32_t* row = status + (size_t)tile * 256;
    st_relaxed_u32(row + d, ((tile == 0 ? 2u : 1u) << 30) | off);
    uint32_t total;
    const uint32_t dstart = block_exclusive_scan(off, s_scan, total);
    s_dstart[d] = dstart;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      const uint32_t i = wbase + r * 32 + lane;
      if (i < n) {
        const uint32_t dd = (__float_as_uint(p[r].w) >> shift) & 255u;
        s_pay[s_dstart[dd] + warp_cnt[warp][dd] + rank[r]] = p[r];
      }
    }
    const uint32_t kMask = (1u << 30) - 1u;
    uint32_t excl = 0;
    if (tile != 0) {
      long long t = static_cast<long long>(tile) - 1;
      while (true) {
        const uint32_t w = ld_relaxed_u32(status + (size_t)t * 256 + d);
        const uint32_t flag = w >> 30;
        if (flag == 0u) { __nanosleep(NDTB200_LOOKBACK_SLEEP_NS); continue; }  // not published yet: leave the issue slots to the co-resident tiles
        excl += w & kMask;
        if (flag == 2u) break;
        NDT_CHECK(t > 0);
        --t;
      }
      st_relaxed_u32(row + d, (2u << 30) | (excl + off));
    }
    s_gbase[d] = static_cast<uint32_t>(digit_base[d]) + excl - dstart;
  }
  __syncthreads();
  const uint32_t count = (n - tbase < (uint32_t)kTile) ? n - tbase : (uint32_t)kTile;
  for (uint32_t j = threadIdx.x; j < count; j += kBuildThreads) {
    const float4 q = s_pay[j];
    const uint32_t key = __float_as_uint(q.w);
    const uint32_t pos = s_gbase[(key >> shift) & 255u] + j;
    NDT_CHECK(pos < n);
    out[pos] = q;
    if (keys_out) keys_out[pos] = key;
  }
}

// 8 consecutive sorted keys of this thread (two 16-byte loads) + the key before them; returns the head flags as bits
__device__ __forceinline__ uint32_t load_heads(const uint32_t* __restrict__ skeys, size_t n, size_t base, uint32_t sentinel,
                                               uint32_t (&key)[kScanItems]) {
  uint32_t prev = sentinel;  // key[-1]: the first element is a head iff it is not the sentinel
  if (base > 0 && base < n) prev = __ldcg(skeys + base - 1);
  if (base + kScanItems <= n) {
    const uint4 a = __ldcg(reinterpret_cast<const uint4*>(skeys + base));
    const uint4 b = __ldcg(reinterpret_cast<const uint4*>(skeys + base) + 1);
    key[0] = a.x; key[1] = a.y; key[2] = a.z; key[3] = a.w; key[4] = b.x; key[5] = b.y; key[6] = b.z; key[7] = b.w;
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) key[k] = (base + k < n) ? __ldcg(skeys + base + k) : sentinel;
  }
  uint32_t flags = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const bool head = (base + k < n) && (key[k] != sentinel) && (base + k == 0 || key[k] != prev);
    flags |= head ? (1u << k) : 0u;
    prev = key[k];
  }
  return flags;
}

__global__ void __launch_bounds__(kBuildThreads)
head_count_kernel(const uint32_t* __restrict__ skeys, size_t n, uint32_t sentinel, uint32_t* __restrict__ tile_counts) {
  static_assert(kScanItems == 8, "load_heads reads two uint4");
  __shared__ uint32_t smem[kBuildThreads / 32 + 1];
  const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
  uint32_t key[kScanItems];
  const uint32_t c = __popc(load_heads(skeys, n, base, sentinel, key));
  uint32_t total;
  block_exclusive_scan(c, smem, total);
  if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kBuildThreads)
head_write_kernel(const uint32_t* __restrict__ skeys, size_t n, uint32_t sentinel,
                  const uint32_t* __restrict__ tile_offsets, int32_t* __restrict__ voxel_key,
                  uint32_t* __restrict__ voxel_start) {
  __shared__ uint32_t smem[kBuildThreads / 32 + 1];
  const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
  uint32_t key[kScanItems];
  const uint32_t flags = load_heads(skeys, n, base, sentinel, key);
  uint32_t total;
  uint32_t off = block_exclusive_scan(__popc(flags), smem, total) + tile_offsets[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (flags & (1u << k)) {
      voxel_key[off] = static_cast<int32_t>(key[k]);
      voxel_start[off] = static_cast<uint32_t>(base + k);
      ++off;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// per-voxel moments: count, sum x, sum x x^T (fp64).  GROUP lanes per voxel stride through the
// voxel's sorted point range (gather through the sorted index), then a fixed-order shuffle tree.
// moments layout per voxel: [sx sy sz sxx sxy sxz syy syz szz] (9 doubles)
// ---------------------------------------------------------------------------------------------
// 16-byte gather load with the smallest L2 prefetch size: a random 16-byte read otherwise drags a 128-byte line out of
// DRAM (measured: 126 B of DRAM traffic per gathered point).
__device__ __forceinline__ float4 ldg_gather16(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L2::64B.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// the GROUP lanes of voxel `gid` sum its points (4 gathers in flight per lane, index order) and fold by a fixed shuffle
// tree: lane gl == 0 ends with the 9 sums.  Every lane of the warp must call this (shuffles).
// SEQ: `pts` is the cloud already in sorted order (payload sort): the loads are sequential, sorted_idx is not read;
// same lanes, same order of additions — bit-identical sums.
template <int GROUP, bool SEQ = false>
__device__ __forceinline__ void voxel_moments_group(const float4* __restrict__ pts, const uint32_t* __restrict__ sorted_idx,
                                                    const uint32_t* __restrict__ voxel_start, uint32_t n_voxels, uint32_t n_finite,
                                                    uint32_t gid, int gl, bool active, double (&s)[9], uint32_t& count) {
#pragma unroll
  for (int k = 0; k < 9; ++k) s[k] = 0.0;
  count = 0;
  if (active) {
    const uint32_t b = voxel_start[gid];
    const uint32_t e = (gid + 1 < n_voxels) ? voxel_start[gid + 1] : n_finite;
    count = e - b;
    // four gathers in flight per lane (index loads, then point loads), accumulated in index order
    for (uint32_t i0 = b + gl; i0 < e; i0 += 4 * GROUP) {
      uint32_t idx[4];
      float4 p[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) idx[u] = (i0 + u * GROUP < e) ? (SEQ ? i0 + u * GROUP : __ldg(sorted_idx + i0 + u * GROUP)) : 0xffffffffu;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        NDT_CHECK(idx[u] == 0xffffffffu || i0 + u * GROUP < n_finite);
        p[u] = (idx[u] != 0xffffffffu) ? (SEQ ? __ldcs(pts + idx[u]) : ldg_gather16(pts + idx[u])) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
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
#pragma unroll
  for (int o = GROUP / 2; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < 9; ++k) s[k] += __shfl_down_sync(0xffffffffu, s[k], o, GROUP);
  }
}

template <int GROUP, bool SEQ = false>
__global__ void __launch_bounds__(kBuildThreads)
voxel_moments_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ sorted_idx,
                     const uint32_t* __restrict__ voxel_start, uint32_t n_voxels, uint32_t n_finite,
                     double* __restrict__ moments) {
  const uint32_t gid = (blockIdx.x * (uint32_t)blockDim.x + threadIdx.x) / GROUP;
  const int gl = threadIdx.x % GROUP;
  const bool active = gid < n_voxels;
  double s[9];
  uint32_t count;
  voxel_moments_group<GROUP, SEQ>(pts, sorted_idx, voxel_start, n_voxels, n_finite, gid, gl, active, s, count);
  if (active && gl == 0) {
#pragma unroll
    for (int k = 0; k < 9; ++k) moments[(size_t)gid * 9 + k] = s[k];
  }
}

__device__ __forceinline__ void finalize_one(uint32_t v, int count, const double (&m)[9], int32_t key, int min_points, double eig_ratio,
                                             VoxelRecord* __restrict__ records, double* __restrict__ icov64,
                                             unsigned int* __restrict__ n_valid, double* __restrict__ dbg_mean,
                                             double* __restrict__ dbg_cov, double* __restrict__ dbg_icov, int* __restrict__ dbg_inflated);

// Both passes of applyFilter for the occupied voxels in ONE kernel (the map build of clouds above scan size): the
// GROUP lanes of a voxel sum its moments, lane 0 finalises the leaf (mean, covariance, eigen-regularisation, inverse:
// finalize_one) straight from registers — the 72-byte moment row never goes to memory and back — and enters a valid
// voxel into the direct-mapped cell table (nullptr: the hash index is filled afterwards).  Same arithmetic in the same
// order as voxel_moments_kernel + finalize_voxels_kernel + dense_fill_kernel.
#ifndef NDTB200_VB_MIN_BLOCKS
#define NDTB200_VB_MIN_BLOCKS 4
#endif
template <int GROUP, bool SEQ = false>
__global__ void __launch_bounds__(kBuildThreads, NDTB200_VB_MIN_BLOCKS)
voxel_build_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ sorted_idx, const int32_t* __restrict__ voxel_key,
                   const uint32_t* __restrict__ voxel_start, uint32_t n_voxels, uint32_t n_finite, int min_points, double eig_ratio,
                   VoxelRecord* __restrict__ records, double* __restrict__ icov64, unsigned int* __restrict__ n_valid,
                   int32_t* __restrict__ dense_table) {
  // A CTA owns kBuildThreads consecutive voxels: their moments are summed kBuildThreads / GROUP voxels at a time (GROUP
  // rounds), then EVERY thread finalises one leaf.  The finalize is a ~3 us serial fp64 chain per thread: with one
  // finalising warp per CTA (32 voxels per CTA, the first version) that latency was paid 8x as often per SM and bounded
  // the kernel (655 us for 3 M voxels); the arithmetic per voxel is unchanged.
  constexpr int kVoxPerRound = kBuildThreads / GROUP;
  constexpr int kVoxPerCta = kBuildThreads;
  __shared__ double s_m[kVoxPerCta][9];
  __shared__ uint32_t s_cnt[kVoxPerCta];
  const int gl = threadIdx.x % GROUP, gi = threadIdx.x / GROUP;
  const uint32_t v0 = blockIdx.x * (uint32_t)kVoxPerCta;
#pragma unroll 1
  for (int r = 0; r < GROUP; ++r) {
    const uint32_t slot = r * kVoxPerRound + gi;
    const uint32_t gid = v0 + slot;
    const bool active = gid < n_voxels;
    double s[9];
    uint32_t count;
    voxel_moments_group<GROUP, SEQ>(pts, sorted_idx, voxel_start, n_voxels, n_finite, gid, gl, active, s, count);
    if (gl == 0) {
#pragma unroll
      for (int k = 0; k < 9; ++k) s_m[slot][k] = s[k];
      s_cnt[slot] = count;
    }
  }
  __syncthreads();
  const uint32_t v = v0 + threadIdx.x;
  if (v < n_voxels) {
    double m[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) m[k] = s_m[threadIdx.x][k];
    const int32_t key = voxel_key[v];
    finalize_one(v, static_cast<int>(s_cnt[threadIdx.x]), m, key, min_points, eig_ratio, records, icov64, n_valid, nullptr, nullptr, nullptr, nullptr);
    if (dense_table != nullptr && records[v].count >= min_points) dense_table[key] = static_cast<int32_t>(v);
  }
}

// ---------------------------------------------------------------------------------------------
// finalize: second pass of applyFilter (…_impl.hpp:282-367), one thread per voxel, fp64
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sym_eig3_jacobi(const double A[3][3], double eval[3], double evec[3][3]) {
  // SelfAdjointEigenSolver<Matrix3d> equivalent: lower triangle in, ascending eigenvalues out.
  double a00 = A[0][0], a11 = A[1][1], a22 = A[2][2], a01 = A[1][0], a02 = A[2][0], a12 = A[2][1];
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
#define NDTB200_ROT(app, aqq, apq, apr, aqr, P, Q)                                                  \
  if (apq != 0.0) {                                                                                 \
    double theta = (aqq - app) / (2.0 * apq);                                                       \
    double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));               \
    double c = 1.0 / sqrt(t * t + 1.0), s = t * c;                                                  \
    app -= t * apq;                                                                                 \
    aqq += t * apq;                                                                                 \
    apq = 0.0;                                                                                      \
    double r_p = apr, r_q = aqr;                                                                    \
    apr = c * r_p - s * r_q;                                                                        \
    aqr = s * r_p + c * r_q;                                                                        \
    for (int k = 0; k < 3; ++k) {                                                                   \
      double vp = v[k][P], vq = v[k][Q];                                                            \
      v[k][P] = c * vp - s * vq;                                                                    \
      v[k][Q] = s * vp + c * vq;                                                                    \
    }                                                                                               \
  }
  for (int sweep = 0; sweep < 32; ++sweep) {
    if (a01 * a01 + a02 * a02 + a12 * a12 == 0.0) break;
    NDTB200_ROT(a00, a11, a01, a02, a12, 0, 1)
    NDTB200_ROT(a00, a22, a02, a01, a12, 0, 2)
    NDTB200_ROT(a11, a22, a12, a01, a02, 1, 2)
  }
#undef NDTB200_ROT
  double d[3] = {a00, a11, a22};
  int o0 = 0, o1 = 1, o2 = 2, t;
  if (d[o1] < d[o0]) { t = o0; o0 = o1; o1 = t; }
  if (d[o2] < d[o1]) { t = o1; o1 = o2; o2 = t; }
  if (d[o1] < d[o0]) { t = o0; o0 = o1; o1 = t; }
  const int ord[3] = {o0, o1, o2};
  for (int j = 0; j < 3; ++j) {
    eval[j] = d[ord[j]];
    for (int i = 0; i < 3; ++i) evec[i][j] = v[i][ord[j]];
  }
}

__device__ __forceinline__ void inv3_cofactor(const double a[3][3], double r[3][3]) {
  double c00 = a[1][1] * a[2][2] - a[1][2] * a[2][1];
  double c01 = a[1][2] * a[2][0] - a[1][0] * a[2][2];
  double c02 = a[1][0] * a[2][1] - a[1][1] * a[2][0];
  double det = a[0][0] * c00 + a[0][1] * c01 + a[0][2] * c02;
  double id = 1.0 / det;
  r[0][0] = c00 * id; r[1][0] = c01 * id; r[2][0] = c02 * id;
  r[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * id;
  r[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * id;
  r[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * id;
  r[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * id;
  r[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * id;
  r[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * id;
}

// dbg_cov / dbg_icov / dbg_inflated: optional full-precision dumps (parity API), NULL in production.
__global__ void __launch_bounds__(kBuildThreads)
finalize_voxels_kernel(const double* __restrict__ moments, const int32_t* __restrict__ voxel_key,
                       const uint32_t* __restrict__ voxel_start, const uint32_t* __restrict__ voxel_count,
                       uint32_t n_voxels, uint32_t n_finite,
                       int min_points, double eig_ratio, VoxelRecord* __restrict__ records,
                       double* __restrict__ icov64, unsigned int* __restrict__ n_valid,
                       double* __restrict__ dbg_mean, double* __restrict__ dbg_cov, double* __restrict__ dbg_icov,
                       int* __restrict__ dbg_inflated) {
  const uint32_t v = blockIdx.x * (uint32_t)blockDim.x + threadIdx.x;
  if (v >= n_voxels) return;
  int count;
  if (voxel_count) {  // merged partials: the count was summed with the moments
    count = static_cast<int>(voxel_count[v]);
  } else {
    const uint32_t b = voxel_start[v];
    const uint32_t e = (v + 1 < n_voxels) ? voxel_start[v + 1] : n_finite;
    count = static_cast<int>(e - b);
  }
  double m[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) m[k] = moments[(size_t)v * 9 + k];
  finalize_one(v, count, m, voxel_key[v], min_points, eig_ratio, records, icov64, n_valid, dbg_mean, dbg_cov, dbg_icov, dbg_inflated);
}

// second pass of applyFilter for ONE voxel (…_impl.hpp:282-367): count, moments {sum x (3), sum x x^T (6)} -> record
__device__ __forceinline__ void finalize_one(uint32_t v, int count, const double (&m)[9], int32_t key, int min_points, double eig_ratio,
                                             VoxelRecord* __restrict__ records, double* __restrict__ icov64,
                                             unsigned int* __restrict__ n_valid, double* __restrict__ dbg_mean,
                                             double* __restrict__ dbg_cov, double* __restrict__ dbg_icov, int* __restrict__ dbg_inflated) {
  const double n = static_cast<double>(count);
  const double pt_sum[3] = {m[0], m[1], m[2]};
  const double mean[3] = {m[0] / n, m[1] / n, m[2] / n};
  // Q1: the reference accumulates x x^T on top of an Identity-initialised cov_ (vgc.h:107)
  double S[3][3] = {{m[3] + 1.0, m[4], m[5]}, {m[4], m[6] + 1.0, m[7]}, {m[5], m[7], m[8] + 1.0}};
  double cov[3][3], icov[3][3];
  bool inflated = false;
  for (int a = 0; a < 3; ++a)
    for (int c = 0; c < 3; ++c) { cov[a][c] = (a == c) ? 1.0 : 0.0; icov[a][c] = 0.0; }
  if (count >= min_points) {
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c)  // …_impl.hpp:329
        cov[a][c] = (S[a][c] - 2 * (pt_sum[a] * mean[c])) / n + mean[a] * mean[c];
    const double scale = (n - 1.0) / n;  // Q3, …_impl.hpp:330
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c) cov[a][c] *= scale;
    // The eigen-decomposition only decides (a) rejection (a negative eigenvalue) and (b) inflation (ev0 < ratio * ev2,
    // …_impl.hpp:337-356).  Invariants of the symmetric matrix bound the spectrum: with trace t > 0, sum of principal
    // 2x2 minors c2 > 0 and det > 0 all eigenvalues are positive, ev2 < t and ev0 >= det / (ev1 ev2) >= det / c2.  So
    // det / c2 >= ratio * t proves "no rejection, no inflation" with a >= 2 % margin, cov stays as it is and the
    // Jacobi sweeps (the bulk of this kernel's fp64 work) are skipped for the well-conditioned voxels.
    const double tr = cov[0][0] + cov[1][1] + cov[2][2];
    const double c2 = (cov[0][0] * cov[1][1] - cov[0][1] * cov[1][0]) + (cov[0][0] * cov[2][2] - cov[0][2] * cov[2][0]) +
                      (cov[1][1] * cov[2][2] - cov[1][2] * cov[2][1]);
    const double det3 = cov[0][0] * (cov[1][1] * cov[2][2] - cov[1][2] * cov[2][1]) -
                        cov[0][1] * (cov[1][0] * cov[2][2] - cov[1][2] * cov[2][0]) +
                        cov[0][2] * (cov[1][0] * cov[2][1] - cov[1][1] * cov[2][0]);
    const bool well_conditioned = (tr > 0.0) && (c2 > 0.0) && (det3 > 0.0) && (eig_ratio <= 0.5) && (det3 >= eig_ratio * tr * c2) &&
                                  !isinf(tr) && !isinf(c2);
    double ev[3] = {1.0, 1.0, 1.0}, evec[3][3];
    if (!well_conditioned) sym_eig3_jacobi(cov, ev, evec);
    if (ev[0] < 0 || ev[1] < 0 || ev[2] <= 0) {
      count = -1;  // …_impl.hpp:337-341
    } else {
      const double min_ev = eig_ratio * ev[2];
      if (!well_conditioned && ev[0] < min_ev) {  // …_impl.hpp:346-356
        ev[0] = min_ev;
        if (ev[1] < min_ev) ev[1] = min_ev;
        double vinv[3][3], vd[3][3];
        inv3_cofactor(evec, vinv);
        for (int a = 0; a < 3; ++a)
          for (int c = 0; c < 3; ++c) vd[a][c] = evec[a][c] * ev[c];
        for (int a = 0; a < 3; ++a)
          for (int c = 0; c < 3; ++c)
            cov[a][c] = vd[a][0] * vinv[0][c] + vd[a][1] * vinv[1][c] + vd[a][2] * vinv[2][c];
        inflated = true;
      }
      inv3_cofactor(cov, icov);  // …_impl.hpp:359
      double mxc = icov[0][0], mnc = icov[0][0];
      for (int a = 0; a < 3; ++a)
        for (int c = 0; c < 3; ++c) { mxc = fmax(mxc, icov[a][c]); mnc = fmin(mnc, icov[a][c]); }
      if (isinf(mxc) && mxc > 0) count = -1;  // …_impl.hpp:360-364
      if (isinf(mnc) && mnc < 0) count = -1;
    }
  }
  VoxelRecord r;
  for (int a = 0; a < 3; ++a) {
    r.mean_hi[a] = static_cast<float>(mean[a]);
    r.mean_lo[a] = static_cast<float>(mean[a] - static_cast<double>(r.mean_hi[a]));
  }
  r.icov[0] = static_cast<float>(icov[0][0]); r.icov[1] = static_cast<float>(icov[0][1]);
  r.icov[2] = static_cast<float>(icov[0][2]); r.icov[3] = static_cast<float>(icov[1][1]);
  r.icov[4] = static_cast<float>(icov[1][2]); r.icov[5] = static_cast<float>(icov[2][2]);
  r.key = key;
  r.count = count;
  r.pad[0] = r.pad[1] = 0;
  records[v] = r;
  double* ic = icov64 + (size_t)v * 6;
  ic[0] = icov[0][0]; ic[1] = icov[0][1]; ic[2] = icov[0][2]; ic[3] = icov[1][1]; ic[4] = icov[1][2]; ic[5] = icov[2][2];
  if (count >= min_points) atomicAdd(n_valid, 1u);
  if (dbg_cov) {
    for (int a = 0; a < 3; ++a) dbg_mean[(size_t)v * 3 + a] = mean[a];
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c) {
        dbg_cov[(size_t)v * 9 + a * 3 + c] = cov[a][c];
        dbg_icov[(size_t)v * 9 + a * 3 + c] = icov[a][c];
      }
    dbg_inflated[v] = inflated ? 1 : 0;
  }
}

__global__ void __launch_bounds__(kBuildThreads)
hash_insert_kernel(const VoxelRecord* __restrict__ records, uint32_t n_voxels, int min_points,
                   HashSlot* __restrict__ table, uint32_t mask, int shift) {
  const uint32_t v = blockIdx.x * (uint32_t)blockDim.x + threadIdx.x;
  if (v >= n_voxels) return;
  const int count = records[v].count;
  if (count < min_points) return;  // sparse and rejected leaves are invisible to lookups (…_impl.hpp:395)
  const uint32_t key = static_cast<uint32_t>(records[v].key);
  const HashSlot slot = (static_cast<HashSlot>(v) << 32) | key;
  uint32_t h = hash_key(key, shift);
  while (true) {
    HashSlot prev = atomicCAS(table + h, NDTB200_HASH_EMPTY, slot);
    if (prev == NDTB200_HASH_EMPTY) return;
    h = (h + 1) & mask;
  }
}

// pcl::VoxelGrid centroid downsample (callers: ndt_omp/apps/align.cpp:57-69, ndt_rosbag_mapping_node.cpp:108-118): one
// thread per occupied cell adds its points in INPUT order in fp32 (the stable sort keeps it) and divides by the count
// — the arithmetic of pcl::VoxelGrid::applyFilter's centroid, so the output is bit-identical; cells ascend by index.
__global__ void __launch_bounds__(kBuildThreads)
voxel_centroid_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ sorted_idx,
                      const uint32_t* __restrict__ voxel_start, uint32_t n_voxels, uint32_t n_finite, float4* __restrict__ out) {
  const uint32_t v = blockIdx.x * (uint32_t)blockDim.x + threadIdx.x;
  if (v >= n_voxels) return;
  const uint32_t b = voxel_start[v];
  const uint32_t e = (v + 1 < n_voxels) ? voxel_start[v + 1] : n_finite;
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (uint32_t i = b; i < e; ++i) {
    const float4 p = __ldg(pts + __ldg(sorted_idx + i));
    sx = __fadd_rn(sx, p.x);
    sy = __fadd_rn(sy, p.y);
    sz = __fadd_rn(sz, p.z);
  }
  const float cnt = static_cast<float>(e - b);
  out[v] = make_float4(__fdiv_rn(sx, cnt), __fdiv_rn(sy, cnt), __fdiv_rn(sz, cnt), 1.0f);
}

// out[j] = pts[idx[j]]: the source cloud in voxel-key order (large sources: see ensure_sorted_source in capi.cu)
__global__ void __launch_bounds__(kBuildThreads)
gather_points_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ idx, size_t n, float4* __restrict__ out) {
  for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x)
    out[j] = ldg_gather16(pts + __ldg(idx + j));
}

// per-voxel point counts of a partial build (length of each sorted range)
__global__ void __launch_bounds__(kBuildThreads)
segment_counts_kernel(const uint32_t* __restrict__ voxel_start, uint32_t n_voxels, uint32_t n_finite, uint32_t* __restrict__ counts) {
  const uint32_t v = blockIdx.x * (uint32_t)blockDim.x + threadIdx.x;
  if (v >= n_voxels) return;
  counts[v] = ((v + 1 < n_voxels) ? voxel_start[v + 1] : n_finite) - voxel_start[v];
}

// sharded build: sum the partials {count, 9 moments} of one voxel in the order of the (stable) sort = rank order
__global__ void __launch_bounds__(kBuildThreads)
merge_partials_kernel(const uint32_t* __restrict__ sorted_idx, const uint32_t* __restrict__ voxel_start, uint32_t n_voxels,
                      uint32_t total, const uint32_t* __restrict__ part_counts, const double* __restrict__ part_moments,
                      uint32_t* __restrict__ counts, double* __restrict__ moments) {
  const uint32_t v = blockIdx.x * (uint32_t)blockDim.x + threadIdx.x;
  if (v >= n_voxels) return;
  const uint32_t b = voxel_start[v];
  const uint32_t e = (v + 1 < n_voxels) ? voxel_start[v + 1] : total;
  uint32_t c = 0;
  double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (uint32_t i = b; i < e; ++i) {
    const uint32_t j = sorted_idx[i];
    c += part_counts[j];
#pragma unroll
    for (int k = 0; k < 9; ++k) s[k] += part_moments[(size_t)j * 9 + k];
  }
  counts[v] = c;
#pragma unroll
  for (int k = 0; k < 9; ++k) moments[(size_t)v * 9 + k] = s[k];
}

// direct-mapped cell table: table[key] = record index of every VALID voxel (the table is pre-filled with -1)
__global__ void __launch_bounds__(kBuildThreads)
dense_fill_kernel(const VoxelRecord* __restrict__ records, uint32_t n_voxels, int min_points, int32_t* __restrict__ table) {
  const uint32_t v = blockIdx.x * (uint32_t)blockDim.x + threadIdx.x;
  if (v >= n_voxels) return;
  if (records[v].count < min_points) return;  // sparse and rejected leaves are invisible to lookups (…_impl.hpp:395)
  NDT_CHECK(records[v].key >= 0);
  table[records[v].key] = static_cast<int32_t>(v);
}

}  // namespace ndtb200
