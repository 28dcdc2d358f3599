// ndt_aux.cuh — kernels around the hot path: output-cloud transform (pcl::transformPointCloud),
// getFitnessScore (pcl::Registration), calculateScore (ndt_omp_impl.hpp:935-983), neighbour lookup dump.
#pragma once
#include "common.cuh"
#include "ndt_align.cuh"

namespace ndtb200 {

// align()'s output cloud: source transformed by final_transformation_, data[3] = 1.
__global__ void __launch_bounds__(256)
transform_output_kernel(const float4* __restrict__ src, int n, const AlignResultDev* __restrict__ res,
                        float4* __restrict__ out) {
  __shared__ float T[12];
  if (threadIdx.x < 12) T[threadIdx.x] = res->final_T[threadIdx.x];
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = __ldg(src + i);
    float4 o;
    transform_point(T, p.x, p.y, p.z, o.x, o.y, o.z);
    o.w = 1.0f;
    out[i] = o;
  }
}

// pcl::transformPointCloud with a matrix given by value (the mapping pipeline's update_global_map)
struct Mat34 { float m[12]; };
__global__ void __launch_bounds__(256)
transform_cloud_kernel(const float4* __restrict__ src, size_t n, Mat34 T, float4* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float4 p = __ldg(src + i);
    float4 o;
    transform_point(T.m, p.x, p.y, p.z, o.x, o.y, o.z);
    o.w = 1.0f;
    out[i] = o;
  }
}

// getFitnessScore: exact 1-NN of T*source in the RAW target, squared distance in fp32 with FLANN's
// L2_Simple order ((dx^2 + dy^2) + dz^2).  Brute force, target streamed through shared memory tiles.
// One partial (sum of accepted d2 in fp64, count) per CTA; the host adds the partials in CTA order.
constexpr int kNNTile = 2048;
__global__ void __launch_bounds__(256)
fitness_bruteforce_kernel(const float4* __restrict__ src, int n_src, const float4* __restrict__ tgt, int n_tgt,
                          const float* __restrict__ T_in, double max_range, double* __restrict__ part_sum,
                          unsigned long long* __restrict__ part_cnt) {
  __shared__ float4 tile[kNNTile];
  __shared__ float T[12];
  __shared__ double s_sum[8];
  __shared__ unsigned int s_cnt[8];
  if (threadIdx.x < 12) T[threadIdx.x] = T_in[threadIdx.x];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float qx = 0, qy = 0, qz = 0;
  if (i < n_src) {
    const float4 p = __ldg(src + i);
    transform_point(T, p.x, p.y, p.z, qx, qy, qz);
  }
  float best = __int_as_float(0x7f800000);
  for (int base = 0; base < n_tgt; base += kNNTile) {
    const int m = min(kNNTile, n_tgt - base);
    __syncthreads();
    for (int k = threadIdx.x; k < m; k += blockDim.x) tile[k] = __ldg(tgt + base + k);
    __syncthreads();
    if (i < n_src) {
#pragma unroll 8
      for (int k = 0; k < m; ++k) {
        const float4 t = tile[k];
        const float dx = qx - t.x, dy = qy - t.y, dz = qz - t.z;
        const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        best = fminf(best, d);
      }
    }
  }
  double v = 0;
  unsigned int c = 0;
  if (i < n_src && static_cast<double>(best) <= max_range) { v = best; c = 1; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v += __shfl_down_sync(0xffffffffu, v, o);
    c += __shfl_down_sync(0xffffffffu, c, o);
  }
  if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = v; s_cnt[threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    unsigned long long cc = 0;
    for (int w = 0; w < 8; ++w) { s += s_sum[w]; cc += s_cnt[w]; }
    part_sum[blockIdx.x] = s;
    part_cnt[blockIdx.x] = cc;
  }
}

// ---- getFitnessScore, grid-accelerated (exact) ----------------------------------------------------------------
// The build already sorted the raw target by NDT cell (sorted point indices + per-voxel ranges).  With a direct-mapped
// table over ALL occupied cells, the nearest raw target point of a query is found by scanning the (2r+1)^3 cells
// around it, r = 1, 2, ...: once the best squared distance is below the squared distance to the faces of the scanned
// cube (minus a rounding margin) nothing outside can be closer, so the result is the brute-force minimum — the same
// fp32 value whatever the scan order.  Queries outside the grid or unresolved after kMaxRing rings go to the
// brute-force kernel below.  Distances: FLANN L2_Simple order, un-fused fp32.
constexpr int kMaxRing = 6;

__device__ __forceinline__ float l2_simple(float qx, float qy, float qz, const float4 t) {
  const float dx = qx - t.x, dy = qy - t.y, dz = qz - t.z;
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// cell table over all occupied voxels (the lookup table of the derivative pass holds only the valid ones)
__global__ void __launch_bounds__(256)
cell_table_fill_kernel(const int32_t* __restrict__ voxel_key, uint32_t n_voxels, int32_t* __restrict__ table) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n_voxels) table[voxel_key[v]] = static_cast<int32_t>(v);
}

// One WARP per query: the lanes share the cells of the current ring (27 cells of ring 1 = one cell per lane), each
// lane scans the points of its cells, a shuffle tree takes the minimum.  A face of the scanned cube only limits the
// search if the grid continues behind it; a query outside the grid starts from the nearest grid cell.
__global__ void __launch_bounds__(256)
fitness_grid_kernel(const float4* __restrict__ src, int n_src, const float4* __restrict__ tgt, const uint32_t* __restrict__ sorted_idx,
                    const uint32_t* __restrict__ voxel_start, uint32_t n_voxels, uint32_t n_finite,
                    const int32_t* __restrict__ cell_table, const GridDesc* __restrict__ gd, const float* __restrict__ T_in,
                    float* __restrict__ best_out, int* __restrict__ fallback_list, int* __restrict__ fallback_count) {
  __shared__ float T[12];
  __shared__ GridDesc g;
  if (threadIdx.x < 12) T[threadIdx.x] = T_in[threadIdx.x];
  if (threadIdx.x == 32) g = *gd;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // warp-uniform query index
  if (i >= n_src) return;
  const float4 p = __ldg(src + i);
  float q[3];
  transform_point(T, p.x, p.y, p.z, q[0], q[1], q[2]);
  float cf[3];
  int c[3];
  bool finite = true;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    finite = finite && (q[a] == q[a]) && (fabsf(q[a]) < 3.0e38f);
    const float rel = __fsub_rn(floorf(__fmul_rn(q[a], g.inv_leaf[a])), static_cast<float>(g.min_b[a]));
    const float relc = fminf(fmaxf(rel, 0.0f), static_cast<float>(g.div_b[a] - 1));  // nearest grid cell along this axis
    c[a] = finite ? static_cast<int>(relc) : 0;
    cf[a] = relc + static_cast<float>(g.min_b[a]);
  }
  float best = __int_as_float(0x7f800000);
  bool resolved = false;
  if (finite) {
    for (int r = 1; r <= kMaxRing && !resolved; ++r) {
      const int side = 2 * r + 1, ncell = side * side * side;
      for (int k = lane; k < ncell; k += 32) {
        const int dx = k % side - r, dy = (k / side) % side - r, dz = k / (side * side) - r;
        if (r > 1 && abs(dx) < r && abs(dy) < r && abs(dz) < r) continue;  // inner cube: scanned by the previous ring
        const int cx = c[0] + dx, cy = c[1] + dy, cz = c[2] + dz;
        if (cx < 0 || cx >= g.div_b[0] || cy < 0 || cy >= g.div_b[1] || cz < 0 || cz >= g.div_b[2]) continue;
        const int v = __ldg(cell_table + (cx * g.mul[0] + cy * g.mul[1] + cz * g.mul[2]));
        if (v < 0) continue;
        const uint32_t b = __ldg(voxel_start + v);
        const uint32_t e = (static_cast<uint32_t>(v) + 1 < n_voxels) ? __ldg(voxel_start + v + 1) : n_finite;
        for (uint32_t j = b; j < e; ++j) best = fminf(best, l2_simple(q[0], q[1], q[2], __ldg(tgt + __ldg(sorted_idx + j))));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
      // distance from the query to those faces of the scanned cube that have grid cells behind them, minus a margin for
      // the fp32 rounding of the cell boundaries (coordinates up to tens of km)
      float margin = 3.402823466e+38f;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float eps = 1e-3f * g.leaf[a] + 4e-7f * (fabsf(q[a]) + g.leaf[a]);
        if (c[a] - r > 0) margin = fminf(margin, q[a] - (cf[a] - static_cast<float>(r)) * g.leaf[a] - eps);
        if (c[a] + r < g.div_b[a] - 1) margin = fminf(margin, (cf[a] + static_cast<float>(r) + 1.0f) * g.leaf[a] - q[a] - eps);
      }
      resolved = (margin > 0.0f) && (best <= margin * margin);
    }
  }
  if (lane == 0) {
    if (resolved) {
      best_out[i] = best;
    } else {
      best_out[i] = -1.0f;
      fallback_list[atomicAdd(fallback_count, 1)] = i;
    }
  }
}

// brute force over the whole raw target for the listed queries (no point within the scanned rings): one CTA per query,
// the 256 threads stride over the target and a block reduction takes the minimum — a handful of far-away queries must
// not cost a serial scan of the target each
__global__ void __launch_bounds__(256)
fitness_fallback_kernel(const float4* __restrict__ src, const int* __restrict__ list, const int* __restrict__ count,
                        const float4* __restrict__ tgt, int n_tgt, const float* __restrict__ T_in, float* __restrict__ best_out) {
  __shared__ float T[12];
  __shared__ float s_min[8];
  const int n = *count;
  if (static_cast<int>(blockIdx.x) >= n) return;
  if (threadIdx.x < 12) T[threadIdx.x] = T_in[threadIdx.x];
  __syncthreads();
  for (int j = blockIdx.x; j < n; j += gridDim.x) {
    const int i = list[j];
    const float4 p = __ldg(src + i);
    float qx, qy, qz;
    transform_point(T, p.x, p.y, p.z, qx, qy, qz);
    float best = __int_as_float(0x7f800000);
    for (int k = threadIdx.x; k < n_tgt; k += blockDim.x) best = fminf(best, l2_simple(qx, qy, qz, __ldg(tgt + k)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
      float m = s_min[0];
      for (int w = 1; w < 8; ++w) m = fminf(m, s_min[w]);
      best_out[i] = m;
    }
    __syncthreads();
  }
}

// mean of the accepted squared distances: one fp64 partial + count per CTA, the host adds them in CTA order
__global__ void __launch_bounds__(256)
fitness_reduce_kernel(const float* __restrict__ best, int n_src, double max_range, double* __restrict__ part_sum,
                      unsigned long long* __restrict__ part_cnt) {
  __shared__ double s_sum[8];
  __shared__ unsigned int s_cnt[8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0;
  unsigned int c = 0;
  if (i < n_src) {
    const float b = best[i];
    if (static_cast<double>(b) <= max_range) { v = b; c = 1; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v += __shfl_down_sync(0xffffffffu, v, o);
    c += __shfl_down_sync(0xffffffffu, c, o);
  }
  if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = v; s_cnt[threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    unsigned long long cc = 0;
    for (int w = 0; w < 8; ++w) { s += s_sum[w]; cc += s_cnt[w]; }
    part_sum[blockIdx.x] = s;
    part_cnt[blockIdx.x] = cc;
  }
}

// calculateScore (ndt_omp_impl.hpp:935-983) of ONE already transformed point: fp64, mean over its neighbours of
// (-d1 * e - d3).  A point without a neighbourhood (also a non-finite one, see point_is_finite) scores 0.
template <int METHOD>
__device__ __forceinline__ double score_of_point(float px, float py, float pz, const MapView& map, double d1, double d2, double d3) {
  if (!point_is_finite(px, py, pz)) return 0.0;
  const int ix = lookup_cell(px, map.leaf[0], map.inv_leaf[0]);
  const int iy = lookup_cell(py, map.leaf[1], map.inv_leaf[1]);
  const int iz = lookup_cell(pz, map.leaf[2], map.inv_leaf[2]);
  constexpr int K = num_offsets<METHOD>();
  int recs[K];
  int cnt = 0;
  for (int k = 0; k < K; ++k) {
    recs[k] = probe_neighbour<METHOD>(map, ix, iy, iz, k, px, py, pz);
    cnt += recs[k] >= 0;
  }
  double score = 0;
  for (int k = 0; k < K; ++k) {
    if (recs[k] < 0) continue;
    const VoxelRecord* R = map.records + recs[k];
    const double* ic = map.icov64 + (size_t)recs[k] * 6;
    const double r0 = static_cast<double>(px) - record_mean(*R, 0), r1 = static_cast<double>(py) - record_mean(*R, 1),
                 r2 = static_cast<double>(pz) - record_mean(*R, 2);
    const double u0 = ic[0] * r0 + ic[1] * r1 + ic[2] * r2, u1 = ic[1] * r0 + ic[3] * r1 + ic[4] * r2,
                 u2 = ic[2] * r0 + ic[4] * r1 + ic[5] * r2;
    const double e = exp(-d2 * (r0 * u0 + r1 * u1 + r2 * u2) / 2);
    score += (-d1 * e - d3) / cnt;
  }
  return score;
}

// Scores of many clouds / many candidate poses in ONE launch (loop-closure screening): segment s covers the
// points [seg_start[s], seg_start[s + 1]) of `cloud` (poses == nullptr: already transformed clouds, one segment each), or
// the WHOLE cloud transformed by poses[s] (row-major 3x4 fp32, pcl::transformPointCloud arithmetic).  One CTA-partial per
// (segment, chunk): part_sum[s * chunks + c]; the host adds a segment's partials in order.
template <int METHOD>
__global__ void __launch_bounds__(256)
calculate_score_kernel(const float4* __restrict__ cloud, const unsigned long long* __restrict__ seg_start, const float* __restrict__ poses,
                       int n_points_per_pose, const MapView map, double d1, double d2, double d3, int chunks,
                       double* __restrict__ part_sum) {
  __shared__ double s_sum[8];
  __shared__ float T[12];
  const int seg = blockIdx.y, chunk = blockIdx.x;
  unsigned long long lo, hi;
  if (poses) {
    if (threadIdx.x < 12) T[threadIdx.x] = poses[(size_t)seg * 12 + threadIdx.x];
    lo = 0; hi = static_cast<unsigned long long>(n_points_per_pose);
  } else {
    lo = seg_start[seg]; hi = seg_start[seg + 1];
  }
  __syncthreads();
  double score = 0;
  for (unsigned long long i = lo + (unsigned long long)chunk * 256 + threadIdx.x; i < hi; i += (unsigned long long)chunks * 256) {
    const float4 p = __ldg(cloud + i);
    float x = p.x, y = p.y, z = p.z;
    if (poses) transform_point(T, p.x, p.y, p.z, x, y, z);
    score += score_of_point<METHOD>(x, y, z, map, d1, d2, d3);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) score += __shfl_down_sync(0xffffffffu, score, o);
  if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = score;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int w = 0; w < 8; ++w) s += s_sum[w];
    part_sum[(size_t)seg * chunks + chunk] = s;
  }
}

// getNeighborhoodAtPoint{,7,1} dump: keys of the valid neighbour voxels, reference offset order, -1 padded.
template <int METHOD>
__global__ void __launch_bounds__(256)
lookup_kernel(const float4* __restrict__ q, int n, const MapView map, int32_t* __restrict__ out_keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = __ldg(q + i);
  const int ix = lookup_cell(p.x, map.leaf[0], map.inv_leaf[0]);
  const int iy = lookup_cell(p.y, map.leaf[1], map.inv_leaf[1]);
  const int iz = lookup_cell(p.z, map.leaf[2], map.inv_leaf[2]);
  constexpr int K = num_offsets<METHOD>();
  int w = 0;
  for (int k = 0; k < K && point_is_finite(p.x, p.y, p.z); ++k) {
    int dx, dy, dz;
    get_offset<METHOD>(k, dx, dy, dz);
    const int rec = probe_cell(map, ix + dx, iy + dy, iz + dz);
    if (rec >= 0) out_keys[(size_t)i * 26 + (w++)] = map.records[rec].key;
  }
  for (; w < 26; ++w) out_keys[(size_t)i * 26 + w] = -1;
}

}  // namespace ndtb200
