// exp_countsort.cu — micro-benchmark behind the map-build design decision: can a counting sort by voxel
// key that uses L2 atomics on a dense per-cell table (one REDG per point to count, one ATOMG per point to take a slot,
// a scattered 16-byte store) beat three ballot-ranked radix passes + the 64 B/point gather?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/exp_countsort tools/exp_countsort.cu
//   /tmp/exp_countsort [points] [reps]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include <cub/cub.cuh>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t pcg(uint32_t v) {
  uint32_t s = v * 747796405u + 2891336453u;
  uint32_t w = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
  return (w >> 22u) ^ w;
}
__device__ __forceinline__ float u01(uint32_t h) { return (h >> 8) * (1.0f / 16777216.0f); }

// the city-like surface samples of tools/build_bench.py (ground, walls of a building grid, roofs), random order
__global__ void gen_points(float4* pts, size_t n, uint32_t seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t h = pcg(static_cast<uint32_t>(i) ^ seed);
    const uint32_t kind = h & 3u;
    h = pcg(h); const float u = (u01(h) - 0.5f) * 600.f;
    h = pcg(h); const float v = (u01(h) - 0.5f) * 600.f;
    h = pcg(h); const float hh = u01(h) * 30.f;
    float x = kind == 1 ? rintf(u / 40.f) * 40.f : u;
    float y = kind == 2 ? rintf(v / 40.f) * 40.f : v;
    float z = kind == 0 ? 0.f : (kind == 3 ? floorf(hh / 10.f) * 10.f : hh);
    h = pcg(h); x += 0.04f * (u01(h) - 0.5f);
    h = pcg(h); y += 0.04f * (u01(h) - 0.5f);
    h = pcg(h); z += 0.04f * (u01(h) - 0.5f);
    pts[i] = make_float4(x, y, z, 1.f);
  }
}

struct Grid { float inv; int minb[3]; int mul[3]; uint32_t ncell; };

__device__ __forceinline__ uint32_t key_of(const float4& p, const Grid& g) {
  const int i = static_cast<int>(floorf(p.x * g.inv)) - g.minb[0];
  const int j = static_cast<int>(floorf(p.y * g.inv)) - g.minb[1];
  const int k = static_cast<int>(floorf(p.z * g.inv)) - g.minb[2];
  return static_cast<uint32_t>(i * g.mul[0] + j * g.mul[1] + k * g.mul[2]);
}

// pass A: per-cell counts, one REDG per point
__global__ void __launch_bounds__(256) count_kernel(const float4* __restrict__ pts, size_t n, Grid g, uint32_t* __restrict__ cnt, uint32_t* __restrict__ keys) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float4 p = __ldcs(pts + i);
    const uint32_t k = key_of(p, g);
    if (keys) keys[i] = k;
    atomicAdd(cnt + k, 1u);
  }
}

// pass C: slot = start[key] + (cursor[key]++), scattered 16-byte store (w carries the key)
__global__ void __launch_bounds__(256) scatter_kernel(const float4* __restrict__ pts, size_t n, Grid g, const uint32_t* __restrict__ start,
                                                      uint32_t* __restrict__ cursor, float4* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float4 p = __ldcs(pts + i);
    const uint32_t k = key_of(p, g);
    const uint32_t pos = __ldg(start + k) + atomicAdd(cursor + k, 1u);
    p.w = __uint_as_float(k);
    out[pos] = p;
  }
}
// same, the cursor table pre-loaded with the start offsets (one table access less per point)
__global__ void __launch_bounds__(256) scatter2_kernel(const float4* __restrict__ pts, size_t n, Grid g, uint32_t* __restrict__ cursor, float4* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float4 p = __ldcs(pts + i);
    const uint32_t k = key_of(p, g);
    const uint32_t pos = atomicAdd(cursor + k, 1u);
    p.w = __uint_as_float(k);
    __stcs(out + pos, p);
  }
}

// pass D: sequential segmented moments over the cell-sorted points: 8 lanes per occupied cell
__global__ void __launch_bounds__(256) moments_kernel(const float4* __restrict__ sorted, const uint32_t* __restrict__ start, const uint32_t* __restrict__ cnt,
                                                      uint32_t ncell, double* __restrict__ out) {
  const uint32_t c = (blockIdx.x * 256u + threadIdx.x) >> 3;
  const int gl = threadIdx.x & 7;
  double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t n = 0;
  if (c < ncell) {
    n = cnt[c];
    const uint32_t b = start[c];
    for (uint32_t i = b + gl; i < b + n; i += 8) {
      const float4 p = __ldcs(sorted + i);
      const double x = p.x, y = p.y, z = p.z;
      s[0] += x; s[1] += y; s[2] += z; s[3] += x * x; s[4] += x * y; s[5] += x * z; s[6] += y * y; s[7] += y * z; s[8] += z * z;
    }
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < 9; ++k) s[k] += __shfl_down_sync(0xffffffffu, s[k], o, 8);
  if (c < ncell && n > 0 && gl == 0) {
    double t = 0;
    for (int k = 0; k < 9; ++k) t += s[k];
    out[c] = t;
  }
}


// ---- V5: two-level counting sort.  Level 1: bucket = key >> kBucketShift, slot from an ATOMG on the bucket cursor, the
// 16-byte stores of a bucket form a moving frontier (write-combined in L2).  Level 2: one CTA per bucket places every
// point at its cell's exact range (cell_start from the full-resolution histogram): scattered stores inside a small window.
constexpr int kBucketShift = 12;
__global__ void __launch_bounds__(256) part1_kernel(const float4* __restrict__ pts, size_t n, Grid g, uint32_t* __restrict__ bucket_cursor, float4* __restrict__ tmp) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float4 p = __ldcs(pts + i);
    const uint32_t k = key_of(p, g);
    const uint32_t pos = atomicAdd(bucket_cursor + (k >> kBucketShift), 1u);
    p.w = __uint_as_float(k);
    tmp[pos] = p;
  }
}
__global__ void bucket_init_kernel(const uint32_t* __restrict__ start, uint32_t ncell, uint32_t nb, uint32_t n, uint32_t* __restrict__ bucket_start, uint32_t* __restrict__ bucket_cursor) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nb) return;
  const uint32_t c = b << kBucketShift;
  const uint32_t s = (b == nb || c >= ncell) ? n : start[c];
  bucket_start[b] = s;
  if (b < nb) bucket_cursor[b] = s;
}
template <bool SMEM_CURSOR>
__global__ void __launch_bounds__(512) part2_kernel(const float4* __restrict__ tmp, const uint32_t* __restrict__ bucket_start, uint32_t nb, uint32_t ncell,
                                                    uint32_t* __restrict__ cell_cursor, unsigned int* __restrict__ ticket, float4* __restrict__ out) {
  __shared__ unsigned int s_b;
  __shared__ uint32_t s_cur[SMEM_CURSOR ? (1 << kBucketShift) : 1];
  while (true) {
    __syncthreads();
    if (threadIdx.x == 0) s_b = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t b = s_b;
    if (b >= nb) return;
    const uint32_t lo = bucket_start[b], hi = bucket_start[b + 1];
    if (SMEM_CURSOR) {
      for (uint32_t c = threadIdx.x; c < (1u << kBucketShift); c += blockDim.x) {
        const uint32_t cell = (b << kBucketShift) + c;
        s_cur[c] = cell < ncell ? cell_cursor[cell] : 0u;
      }
      __syncthreads();
    }
    for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      const float4 p = __ldcs(tmp + i);
      const uint32_t k = __float_as_uint(p.w);
      const uint32_t pos = SMEM_CURSOR ? atomicAdd(&s_cur[k & ((1u << kBucketShift) - 1u)], 1u) : atomicAdd(cell_cursor + k, 1u);
      out[pos] = p;
    }
  }
}

// sequential segmented moments, v2: a warp owns 32 consecutive cells and walks its occupied ones, 8 lanes per cell
__global__ void __launch_bounds__(256) moments2_kernel(const float4* __restrict__ sorted, const uint32_t* __restrict__ start, const uint32_t* __restrict__ cnt,
                                                       uint32_t ncell, double* __restrict__ out) {
  const uint32_t warp = (blockIdx.x * 256u + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31, g = lane >> 3, gl = lane & 7;
  const uint32_t c = warp * 32u + lane;
  uint32_t n = 0, b = 0;
  if (c < ncell) { n = __ldg(cnt + c); b = __ldg(start + c); }
  unsigned int occ = __ballot_sync(0xffffffffu, n > 0);
  while (occ) {
    const unsigned int src = __fns(occ, 0, g + 1);  // the (g+1)-th occupied cell of this round, 0xffffffff if none
    const bool act = src != 0xffffffffu;
    const uint32_t nn = __shfl_sync(0xffffffffu, n, act ? src : 0), bb = __shfl_sync(0xffffffffu, b, act ? src : 0);
    double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (act)
      for (uint32_t i = bb + gl; i < bb + nn; i += 8) {
        const float4 p = __ldcs(sorted + i);
        const double x = p.x, y = p.y, z = p.z;
        s[0] += x; s[1] += y; s[2] += z; s[3] += x * x; s[4] += x * y; s[5] += x * z; s[6] += y * y; s[7] += y * z; s[8] += z * z;
      }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1)
#pragma unroll
      for (int k = 0; k < 9; ++k) s[k] += __shfl_down_sync(0xffffffffu, s[k], o, 8);
    if (act && gl == 0) {
      double t = 0;
      for (int k = 0; k < 9; ++k) t += s[k];
      out[warp * 32u + src] = t;
    }
    // drop the (up to) four cells just processed
    for (int k = 0; k < 4 && occ; ++k) occ &= occ - 1;
  }
}

int main(int argc, char** argv) {
  const size_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 100000000ull;
  const int reps = argc > 2 ? atoi(argv[2]) : 3;
  Grid g;
  g.inv = 1.0f;
  g.minb[0] = -301; g.minb[1] = -301; g.minb[2] = -1;
  const int dx = 603, dy = 603, dz = 33;
  g.mul[0] = 1; g.mul[1] = dx; g.mul[2] = dx * dy;
  g.ncell = static_cast<uint32_t>(dx) * dy * dz;
  printf("points %zu, cells %u (%.1f MB table)\n", n, g.ncell, g.ncell * 4 / 1e6);
  float4 *pts, *out;
  uint32_t *cnt, *start, *cursor, *keys;
  double* mom;
  CK(cudaMalloc(&pts, n * 16)); CK(cudaMalloc(&out, n * 16));
  CK(cudaMalloc(&cnt, (size_t)g.ncell * 4)); CK(cudaMalloc(&start, (size_t)g.ncell * 4)); CK(cudaMalloc(&cursor, (size_t)g.ncell * 4));
  CK(cudaMalloc(&keys, n * 4)); CK(cudaMalloc(&mom, (size_t)g.ncell * 8));
  void* tmp = nullptr; size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, start, g.ncell);
  CK(cudaMalloc(&tmp, tmp_bytes));
  gen_points<<<148 * 8, 256>>>(pts, n, 12345u);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e[8];
  for (auto& x : e) CK(cudaEventCreate(&x));
  const int grid = 148 * 8;
  for (int r = 0; r < reps; ++r) {
    CK(cudaMemsetAsync(cnt, 0, (size_t)g.ncell * 4));
    CK(cudaMemsetAsync(cursor, 0, (size_t)g.ncell * 4));
    CK(cudaEventRecord(e[0]));
    count_kernel<<<grid, 256>>>(pts, n, g, cnt, nullptr);
    CK(cudaEventRecord(e[1]));
    cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, start, g.ncell);
    CK(cudaEventRecord(e[2]));
    scatter_kernel<<<grid, 256>>>(pts, n, g, start, cursor, out);
    CK(cudaEventRecord(e[3]));
    moments_kernel<<<(g.ncell * 8 + 255) / 256, 256>>>(out, start, cnt, g.ncell, mom);
    CK(cudaEventRecord(e[4]));
    CK(cudaMemcpyAsync(cursor, start, (size_t)g.ncell * 4, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e[5]));
    scatter2_kernel<<<grid, 256>>>(pts, n, g, cursor, out);
    CK(cudaEventRecord(e[6]));
    CK(cudaDeviceSynchronize());
    float t[6];
    for (int k = 0; k < 6; ++k) CK(cudaEventElapsedTime(&t[k], e[k], e[k + 1]));
    printf("rep %d: count %.3f ms | scan %.3f | scatter(start+cursor) %.3f | moments(seq) %.3f | copy %.3f | scatter(cursor only, st.cs) %.3f  => total(A+scan+C+D) %.3f ms\n",
           r, t[0], t[1], t[2], t[3], t[4], t[5], t[0] + t[1] + t[2] + t[3]);
  }

  {
    const uint32_t nb = (g.ncell + (1u << kBucketShift) - 1) >> kBucketShift;
    uint32_t *bstart, *bcur; unsigned int* ticket;
    CK(cudaMalloc(&bstart, (nb + 1) * 4)); CK(cudaMalloc(&bcur, (nb + 1) * 4)); CK(cudaMalloc(&ticket, 4));
    float4* tmpbuf; CK(cudaMalloc(&tmpbuf, n * 16));
    cudaEvent_t f[10];
    for (auto& x : f) CK(cudaEventCreate(&x));
    for (int variant = 0; variant < 2; ++variant)
    for (int r = 0; r < reps; ++r) {
      CK(cudaMemsetAsync(cnt, 0, (size_t)g.ncell * 4));
      CK(cudaMemsetAsync(ticket, 0, 4));
      CK(cudaEventRecord(f[0]));
      count_kernel<<<grid, 256>>>(pts, n, g, cnt, nullptr);
      CK(cudaEventRecord(f[1]));
      cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, start, g.ncell);
      CK(cudaMemcpyAsync(cursor, start, (size_t)g.ncell * 4, cudaMemcpyDeviceToDevice));
      bucket_init_kernel<<<(nb + 256) / 256, 256>>>(start, g.ncell, nb, (uint32_t)n, bstart, bcur);
      CK(cudaEventRecord(f[2]));
      part1_kernel<<<grid, 256>>>(pts, n, g, bcur, tmpbuf);
      CK(cudaEventRecord(f[3]));
      if (variant == 0) part2_kernel<false><<<148 * 2, 512>>>(tmpbuf, bstart, nb, g.ncell, cursor, ticket, out);
      else part2_kernel<true><<<148 * 2, 512>>>(tmpbuf, bstart, nb, g.ncell, cursor, ticket, out);
      CK(cudaEventRecord(f[4]));
      moments2_kernel<<<((g.ncell + 31) / 32 * 32 + 255) / 256, 256>>>(out, start, cnt, g.ncell, mom);
      CK(cudaEventRecord(f[5]));
      CK(cudaDeviceSynchronize());
      float t[5];
      for (int k = 0; k < 5; ++k) CK(cudaEventElapsedTime(&t[k], f[k], f[k + 1]));
      printf("V5%s rep %d: count %.3f | scan+init %.3f | part1 %.3f | part2 %.3f | moments2 %.3f => total %.3f ms (%u buckets)\n", variant ? "(smem cursors)" : "(ATOMG cursors)", r,
             t[0], t[1], t[2], t[3], t[4], t[0] + t[1] + t[2] + t[3] + t[4], nb);
    }
    // check: every point of out sits in its cell's range
    {
      std::vector<float4> ho(n < 4000000 ? n : 4000000);
      CK(cudaMemcpy(ho.data(), out, ho.size() * 16, cudaMemcpyDeviceToHost));
      std::vector<uint32_t> hs(g.ncell), hcnt(g.ncell);
      CK(cudaMemcpy(hs.data(), start, (size_t)g.ncell * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hcnt.data(), cnt, (size_t)g.ncell * 4, cudaMemcpyDeviceToHost));
      size_t bad = 0;
      for (size_t i = 0; i < ho.size(); ++i) {
        uint32_t k; memcpy(&k, &ho[i].w, 4);
        if (k >= g.ncell || i < hs[k] || i >= (size_t)hs[k] + hcnt[k]) ++bad;
      }
      printf("V5 placement check over the first %zu points: %zu misplaced\n", ho.size(), bad);
    }
  }
  // sanity: every point landed in its cell's range
  uint32_t* hc = (uint32_t*)malloc(16 * 4);
  CK(cudaMemcpy(hc, cnt, 64, cudaMemcpyDeviceToHost));
  unsigned long long occupied = 0;
  {
    uint32_t* h = (uint32_t*)malloc((size_t)g.ncell * 4);
    CK(cudaMemcpy(h, cnt, (size_t)g.ncell * 4, cudaMemcpyDeviceToHost));
    unsigned long long tot = 0;
    for (uint32_t c = 0; c < g.ncell; ++c) { tot += h[c]; occupied += h[c] != 0; }
    printf("sum of counts %llu (n %zu), occupied cells %llu\n", tot, n, occupied);
    free(h);
  }
  return 0;
}
