// ndt_align.cuh — the registration hot path as ONE persistent kernel:
//
//   derivative pass   computeDerivatives / updateDerivatives / computePointDerivatives
//                     (ndt_omp_impl.hpp:179-285, 398-440, 484-537), fp32 per-hit math
//   Hessian-only pass computeHessian / updateHessian (ndt_omp_impl.hpp:540-645), fp64 math, fp64 tables
//   Newton + line search  computeTransformation / computeStepLengthMT (ndt_omp_impl.hpp:80-171, 772-932)
//
// Design (B200-first, not the reference's per-point-slot + serial-sum structure):
//   * one thread per source point, float4 loads, in-register fp32 transform (never materialises
//     trans_cloud), DIRECT1/7/26 probes into the HBM voxel hash (all probes of a point issued
//     together), 64-byte Gaussian records prefetched, then consumed hit by hit;
//   * warps take 32-point groups interleaved over the whole grid, so every CTA sees the same mix of
//     dense and empty regions of the scan (static, deterministic load balance);
//   * accumulation: a thread adds the (fp32) contributions of at most kFlushPoints points in fp32,
//     then the warp folds them into fp64 (shuffle tree) — fp64 everywhere a long sum is formed,
//     without the fp32->fp64 conversion per hit that saturated the XU pipe in the first version;
//     warp sums -> one fp64 partial per CTA -> the LAST-ARRIVING CTA sums the partials in CTA order
//     (bit-reproducible) and publishes the 28 totals;
//   * every CTA then runs the identical Newton / More-Thuente step (scalar part in one thread, the
//     trigonometry and the 69 table entries spread over a warp), so ONE grid barrier per evaluation
//     is the only synchronisation and the whole align() never returns to the host;
//   * launched cooperatively (all CTAs co-resident), gridDim = #SMs x occupancy.
#pragma once
#include "common.cuh"
#include "ndt_solve.cuh"

namespace ndtb200 {

constexpr int kAlignThreads = 256;
constexpr int kAlignWarps = kAlignThreads / 32;
constexpr int kNV = 29;          // score, g[6], H upper triangle[21], hit count
constexpr int kNVP = 32;         // padded row length of the partial / total buffers
constexpr int kFlushPoints = 8;  // fp32 run length (points per thread) before folding into fp64

enum { ACT_DONE = 0, ACT_EVAL_FULL = 1, ACT_EVAL_NOHESS = 2, ACT_HESS_ONLY = 3, ACT_SOLVE = 4 };
enum { ST_INITIAL = 0, ST_MT_FIRST = 1, ST_MT_LOOP = 2, ST_MT_HESS = 3, ST_SINGLE = 4 };
enum { MODE_ALIGN = 0, MODE_EVAL = 1, MODE_HESSIAN = 2 };

struct TraceRec {
  int32_t kind;  // 0 derivatives+hessian, 1 derivatives only, 2 hessian only
  int32_t pad;
  double x[6];
  double a_t;
  double score;
  // CTA-0 timeline of this evaluation (globaltimer ns): start, local work done, totals available, step decided
  unsigned long long t_start, t_local, t_reduced, t_advanced;
  unsigned long long t_phase[2];  // CTA 0: first chunk's phase A done, phase B done
  unsigned long long t_dbg[4];  // step breakdown: after advance(), after the warp solve, after newton_post(), (spare)
};

struct AlignResultDev {
  float final_T[12];
  double last_dp[6];
  int32_t converged, iterations, n_evals, n_hess;
  double trans_probability;
  double final_pose[6];
  double final_score;
  double totals[43];  // score, g[6], H[36] of the last evaluation
  long long n_hits;
  int32_t n_trace, aborted;
  unsigned long long t_kernel_begin, t_kernel_end;  // globaltimer ns, CTA 0
};

struct AlignParams {
  double d1, d2, d3;
  double step_size, trans_eps;
  int32_t max_iterations;
  int32_t mode;
  int32_t eval_hessian;
  int32_t has_guess;
  double p0[6];
  float T0[12];
  int32_t n_source;
  int32_t trace_cap;
  uint32_t launch_tag;  // launch sequence number << 10: makes the tags of the published totals unique per launch
  AngleTables tab0;     // angle tables of p0 (host-computed: the kernel prologue has no trigonometry)
};

constexpr int kMaxRanks = 8;  // GPUs of one NVSwitch node

struct AlignWorkspace {
  double* partials;      // [gridDim][kNVP]
  double* totals;        // [2][kNVP]
  unsigned int* sync;    // [0] arrive counter, [2] wrapping exit counter, [3] abort flag (zeroed once, self-resetting)
  AlignResultDev* result;
  TraceRec* trace;
  // source-sharded multi-GPU solve: every rank holds a slice of the source and the full map; the
  // 29 per-evaluation sums are exchanged by direct stores into every peer's mailbox over NVLink.
  int32_t world, rank;
  unsigned long long* mail[kMaxRanks];  // mail[r]: rank r's mailbox [2 parities][kMaxRanks][kNVP][2 words] (own or IPC-mapped peer memory)
  long long n_source_total;              // points of the whole (unsharded) source cloud
};

struct EvalCtx {
  float T[12];
  AngleTables tab;
};

struct SolverState {
  double p[6], score, g[6], H[36];
  double dir[6];
  double x_t[6];
  double phi_0, d_phi_0, a_l, f_l, g_l, a_u, f_u, g_u, a_t;
  double phi_t, d_phi_t, psi_t, d_psi_t;
  double step_max, step_min;
  long long n_hits;
  int32_t step_iterations, interval_converged, open_interval;
  int32_t nr_iterations, converged, state, n_evals, n_hess, n_trace;
  int32_t need_pose;  // 1 => ctx (matrix + tables) must be rebuilt from x_t before the next evaluation
  double last_dp[6];  // last Newton increment (transformation_ is built from it once, at the end)
  double delta[6];    // Newton step H^-1 (-g), written by the warp solver
  float final_T[12];
};

// ---------------------------------------------------------------------------------------------
// one (point, voxel) contribution: updateDerivatives (T = float) / updateHessian (T = double)
// pj = j_ang * x (8 values), ph = h_ang * x (15 values); r = x' - mean; c = inverse covariance.
// acc layout: [0] score, [1..6] gradient, [7..27] Hessian upper triangle row-major.
// d1 is passed in T: for the fp32 path the reference forms -d1*e and d1*(d2*e) in fp64 and rounds
// to fp32 (ndt_omp_impl.hpp:501, 510); multiplying by fl32(d1) in fp32 differs from that by at most
// 1.2e-7 relative, and keeps the per-hit path free of fp64 and conversion instructions.
// ---------------------------------------------------------------------------------------------
template <typename T, bool SCORE_GRAD, bool HESS>
__device__ __forceinline__ void hit_contribution(const T r0, const T r1, const T r2, const T c00, const T c01,
                                                 const T c02, const T c11, const T c12, const T c22, const T* pj,
                                                 const T* ph, const T d2, const T d1, T* acc) {
  const T u0 = c00 * r0 + c01 * r1 + c02 * r2;
  const T u1 = c01 * r0 + c11 * r1 + c12 * r2;
  const T u2 = c02 * r0 + c12 * r1 + c22 * r2;
  const T q = r0 * u0 + r1 * u1 + r2 * u2;
  T w;
  if constexpr (sizeof(T) == 4) {
    // ndt_omp_impl.hpp:499-510
    const float e = expf(-d2 * q * 0.5f);
    const float e2 = d2 * e;
    if (!(e2 <= 1.0f && e2 >= 0.0f)) return;  // e2 > 1 || e2 < 0 || NaN  -> contributes nothing
    if (SCORE_GRAD) acc[0] -= d1 * e;
    w = e2 * d1;
  } else {
    // ndt_omp_impl.hpp:622-629
    const double e2 = d2 * exp(-d2 * q / 2);
    if (!(e2 <= 1.0 && e2 >= 0.0)) return;
    w = e2 * d1;
  }
  const T s0 = u0, s1 = u1, s2 = u2;
  const T s3 = u1 * pj[0] + u2 * pj[1];
  const T s4 = u0 * pj[2] + u1 * pj[3] + u2 * pj[4];
  const T s5 = u0 * pj[5] + u1 * pj[6] + u2 * pj[7];
  if (SCORE_GRAD) {
    acc[1] += w * s0;
    acc[2] += w * s1;
    acc[3] += w * s2;
    acc[4] += w * s3;
    acc[5] += w * s4;
    acc[6] += w * s5;
  }
  if (HESS) {
    // C * J_i for the three rotational columns (J_3 = (0,j0,j1), J_4 = (j2,j3,j4), J_5 = (j5,j6,j7))
    const T v3x = c01 * pj[0] + c02 * pj[1], v3y = c11 * pj[0] + c12 * pj[1], v3z = c12 * pj[0] + c22 * pj[1];
    const T v4x = c00 * pj[2] + c01 * pj[3] + c02 * pj[4], v4y = c01 * pj[2] + c11 * pj[3] + c12 * pj[4],
            v4z = c02 * pj[2] + c12 * pj[3] + c22 * pj[4];
    const T v5x = c00 * pj[5] + c01 * pj[6] + c02 * pj[7], v5y = c01 * pj[5] + c11 * pj[6] + c12 * pj[7],
            v5z = c02 * pj[5] + c12 * pj[6] + c22 * pj[7];
    const T md2 = -d2;
#define NDTB200_H(idx, si, sj, extra) acc[7 + idx] += w * (md2 * si * sj + (extra));
    NDTB200_H(0, s0, s0, c00)
    NDTB200_H(1, s0, s1, c01)
    NDTB200_H(2, s0, s2, c02)
    NDTB200_H(3, s0, s3, v3x)
    NDTB200_H(4, s0, s4, v4x)
    NDTB200_H(5, s0, s5, v5x)
    NDTB200_H(6, s1, s1, c11)
    NDTB200_H(7, s1, s2, c12)
    NDTB200_H(8, s1, s3, v3y)
    NDTB200_H(9, s1, s4, v4y)
    NDTB200_H(10, s1, s5, v5y)
    NDTB200_H(11, s2, s2, c22)
    NDTB200_H(12, s2, s3, v3z)
    NDTB200_H(13, s2, s4, v4z)
    NDTB200_H(14, s2, s5, v5z)
    // rotational block: u . H_E[i][j]  +  J_j . (C J_i)
    NDTB200_H(15, s3, s3, (u1 * ph[0] + u2 * ph[1]) + (pj[0] * v3y + pj[1] * v3z))
    NDTB200_H(16, s3, s4, (u1 * ph[2] + u2 * ph[3]) + (pj[2] * v3x + pj[3] * v3y + pj[4] * v3z))
    NDTB200_H(17, s3, s5, (u1 * ph[4] + u2 * ph[5]) + (pj[5] * v3x + pj[6] * v3y + pj[7] * v3z))
    NDTB200_H(18, s4, s4, (u0 * ph[6] + u1 * ph[7] + u2 * ph[8]) + (pj[2] * v4x + pj[3] * v4y + pj[4] * v4z))
    NDTB200_H(19, s4, s5, (u0 * ph[9] + u1 * ph[10] + u2 * ph[11]) + (pj[5] * v4x + pj[6] * v4y + pj[7] * v4z))
    NDTB200_H(20, s5, s5, (u0 * ph[12] + u1 * ph[13] + u2 * ph[14]) + (pj[5] * v5x + pj[6] * v5y + pj[7] * v5z))
#undef NDTB200_H
  }
}

__constant__ int8_t c_off26[26][3] = {
    // pcl::getAllNeighborCellIndices(): 13 "half" offsets then their negations (centre excluded, Q7)
    {-1, -1, -1}, {-1, 0, -1}, {-1, 1, -1}, {0, -1, -1}, {0, 0, -1}, {0, 1, -1}, {1, -1, -1}, {1, 0, -1}, {1, 1, -1},
    {-1, -1, 0}, {0, -1, 0}, {1, -1, 0}, {-1, 0, 0},
    {1, 1, 1}, {1, 0, 1}, {1, -1, 1}, {0, 1, 1}, {0, 0, 1}, {0, -1, 1}, {-1, 1, 1}, {-1, 0, 1}, {-1, -1, 1},
    {1, 1, 0}, {0, 1, 0}, {-1, 1, 0}, {1, 0, 0}};

template <int METHOD>
__device__ __forceinline__ constexpr int num_offsets() {
  return METHOD == 3 ? 1 : (METHOD == 2 ? 7 : 26);
}

// DIRECT7 order (voxel_grid_covariance_omp_impl.hpp:423-430): centre, +x, -x, +y, -y, +z, -z
template <int METHOD>
__device__ __forceinline__ void get_offset(int k, int& dx, int& dy, int& dz) {
  if (METHOD == 3) {
    dx = dy = dz = 0;
  } else if (METHOD == 2) {
    dx = (k == 1) - (k == 2);
    dy = (k == 3) - (k == 4);
    dz = (k == 5) - (k == 6);
  } else {
    dx = c_off26[k][0]; dy = c_off26[k][1]; dz = c_off26[k][2];
  }
}

// Probe one neighbour cell (voxel_grid_covariance_omp_impl.hpp:388-400): bounds test, key, hash find.
__device__ __forceinline__ int probe_cell(const MapView& m, int cx, int cy, int cz) {
  if (cx < m.min_b[0] || cx > m.max_b[0] || cy < m.min_b[1] || cy > m.max_b[1] || cz < m.min_b[2] || cz > m.max_b[2])
    return -1;
  const int key = (cx - m.min_b[0]) * m.mul[0] + (cy - m.min_b[1]) * m.mul[1] + (cz - m.min_b[2]) * m.mul[2];
  return map_find(m, key);
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// All K probes of a point with their first hash slots in flight together (memory-level parallelism),
// then resolved; collisions (rare at load <= 0.25) continue with the plain linear probe.
template <int K>
__device__ __forceinline__ void probe_cells(const MapView& m, int ix, int iy, int iz, int (&rec)[K]) {
  static_assert(K == 1 || K == 7, "unrolled probe is for DIRECT1 / DIRECT7");
  HashSlot slot[K];
  uint32_t key[K], h[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    int dx, dy, dz;
    get_offset<(K == 1 ? 3 : 2)>(k, dx, dy, dz);
    const int cx = ix + dx, cy = iy + dy, cz = iz + dz;
    const bool inside = !(cx < m.min_b[0] || cx > m.max_b[0] || cy < m.min_b[1] || cy > m.max_b[1] ||
                          cz < m.min_b[2] || cz > m.max_b[2]);
    key[k] = static_cast<uint32_t>((cx - m.min_b[0]) * m.mul[0] + (cy - m.min_b[1]) * m.mul[1] + (cz - m.min_b[2]) * m.mul[2]);
    h[k] = hash_key(key[k], m.hash_shift);
    slot[k] = inside ? __ldg(m.hash + h[k]) : NDTB200_HASH_EMPTY;
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    int r = -1;
    HashSlot s = slot[k];
    uint32_t hh = h[k];
    while (s != NDTB200_HASH_EMPTY) {
      if (static_cast<uint32_t>(s) == key[k]) { r = static_cast<int>(s >> 32); break; }
      hh = (hh + 1) & m.hash_mask;
      s = __ldg(m.hash + hh);
    }
    rec[k] = r;
    if (r >= 0) prefetch_l2(m.records + r);
  }
}

// ---------------------------------------------------------------------------------------------
// all hits of one source point, fp32 path (computeDerivatives inner loop, ndt_omp_impl.hpp:207-275)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void hit_from_record_f32(const VoxelRecord* R, float tx, float ty, float tz, float& r0,
                                                    float& r1, float& r2, float (&c)[6]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(R));
  const float4 b = __ldg(reinterpret_cast<const float4*>(R) + 1);
  const float4 cc = __ldg(reinterpret_cast<const float4*>(R) + 2);
  // x_trans = fl32(double(x') - mean) (ndt_omp_impl.hpp:259-262, 492) via the exact hi/lo split of the mean
  r0 = __fsub_rn(__fsub_rn(tx, a.x), a.w);
  r1 = __fsub_rn(__fsub_rn(ty, a.y), b.x);
  r2 = __fsub_rn(__fsub_rn(tz, a.z), b.y);
  c[0] = b.z; c[1] = b.w; c[2] = cc.x; c[3] = cc.y; c[4] = cc.z; c[5] = cc.w;
}

template <int METHOD, bool HESS>
__device__ __forceinline__ void eval_point_f32(const float4 pt, const EvalCtx& c, const MapView& m, const float d2f,
                                               const float d1f, float* acc) {
  float tx, ty, tz;
  transform_point(c.T, pt.x, pt.y, pt.z, tx, ty, tz);
  // getNeighborhoodAtPoint (…_impl.hpp:379-381): cell = floor(x' / leaf), fp32 DIVISION (Q8)
  const int ix = static_cast<int>(floorf(__fdiv_rn(tx, m.leaf[0])));
  const int iy = static_cast<int>(floorf(__fdiv_rn(ty, m.leaf[1])));
  const int iz = static_cast<int>(floorf(__fdiv_rn(tz, m.leaf[2])));
  constexpr int K = num_offsets<METHOD>();
  if constexpr (METHOD != 1) {
    int rec[K];
    probe_cells<K>(m, ix, iy, iz, rec);
    // computePointDerivatives (fp32 overload, ndt_omp_impl.hpp:398-440): depends on the ORIGINAL point only
    float pj[8], ph[15];
#pragma unroll
    for (int r = 0; r < 8; ++r) pj[r] = c.tab.jf[r][0] * pt.x + c.tab.jf[r][1] * pt.y + c.tab.jf[r][2] * pt.z;
    if (HESS) {
#pragma unroll
      for (int r = 0; r < 15; ++r) ph[r] = c.tab.hf[r][0] * pt.x + c.tab.hf[r][1] * pt.y + c.tab.hf[r][2] * pt.z;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (rec[k] < 0) continue;
      float r0, r1, r2, cv[6];
      hit_from_record_f32(m.records + rec[k], tx, ty, tz, r0, r1, r2, cv);
      acc[28] += 1.0f;
      hit_contribution<float, true, HESS>(r0, r1, r2, cv[0], cv[1], cv[2], cv[3], cv[4], cv[5], pj, ph, d2f, d1f, acc);
    }
  } else {
    float pj[8], ph[15];
#pragma unroll
    for (int r = 0; r < 8; ++r) pj[r] = c.tab.jf[r][0] * pt.x + c.tab.jf[r][1] * pt.y + c.tab.jf[r][2] * pt.z;
    if (HESS) {
#pragma unroll
      for (int r = 0; r < 15; ++r) ph[r] = c.tab.hf[r][0] * pt.x + c.tab.hf[r][1] * pt.y + c.tab.hf[r][2] * pt.z;
    }
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
      int dx, dy, dz;
      get_offset<METHOD>(k, dx, dy, dz);
      const int rec = probe_cell(m, ix + dx, iy + dy, iz + dz);
      if (rec < 0) continue;
      float r0, r1, r2, cv[6];
      hit_from_record_f32(m.records + rec, tx, ty, tz, r0, r1, r2, cv);
      acc[28] += 1.0f;
      hit_contribution<float, true, HESS>(r0, r1, r2, cv[0], cv[1], cv[2], cv[3], cv[4], cv[5], pj, ph, d2f, d1f, acc);
    }
  }
}

// fp64 Hessian-only path (computeHessian inner loop, ndt_omp_impl.hpp:565-609)
template <int METHOD>
__device__ __forceinline__ void eval_point_f64(const float4 pt, const EvalCtx& c, const MapView& m, const double d2,
                                               const double d1, double* acc) {
  float tx, ty, tz;
  transform_point(c.T, pt.x, pt.y, pt.z, tx, ty, tz);
  const int ix = static_cast<int>(floorf(__fdiv_rn(tx, m.leaf[0])));
  const int iy = static_cast<int>(floorf(__fdiv_rn(ty, m.leaf[1])));
  const int iz = static_cast<int>(floorf(__fdiv_rn(tz, m.leaf[2])));
  const double x = pt.x, y = pt.y, z = pt.z;
  double pj[8], ph[15];
#pragma unroll
  for (int r = 0; r < 8; ++r) pj[r] = x * c.tab.jd[r][0] + y * c.tab.jd[r][1] + z * c.tab.jd[r][2];
#pragma unroll
  for (int r = 0; r < 15; ++r) ph[r] = x * c.tab.hd[r][0] + y * c.tab.hd[r][1] + z * c.tab.hd[r][2];
  constexpr int K = num_offsets<METHOD>();
#pragma unroll 1
  for (int k = 0; k < K; ++k) {
    int dx, dy, dz;
    get_offset<METHOD>(k, dx, dy, dz);
    const int rec = probe_cell(m, ix + dx, iy + dy, iz + dz);
    if (rec < 0) continue;
    const VoxelRecord* R = m.records + rec;
    const double* ic = m.icov64 + (size_t)rec * 6;
    const double r0 = static_cast<double>(tx) - (static_cast<double>(__ldg(&R->mean_hi[0])) + static_cast<double>(__ldg(&R->mean_lo[0])));
    const double r1 = static_cast<double>(ty) - (static_cast<double>(__ldg(&R->mean_hi[1])) + static_cast<double>(__ldg(&R->mean_lo[1])));
    const double r2 = static_cast<double>(tz) - (static_cast<double>(__ldg(&R->mean_hi[2])) + static_cast<double>(__ldg(&R->mean_lo[2])));
    acc[28] += 1.0;
    hit_contribution<double, false, true>(r0, r1, r2, __ldg(ic), __ldg(ic + 1), __ldg(ic + 2), __ldg(ic + 3),
                                          __ldg(ic + 4), __ldg(ic + 5), pj, ph, d2, d1, acc);
  }
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
template <typename A>
__device__ __forceinline__ void warp_flush(A* acc, double* s_warp_row);

// The whole fp64 Hessian-only pass of one warp.  Rare (only after a multi-trial line search) and register hungry
// (46 fp64 coefficients + 22 fp64 accumulators): kept out of line so its register pressure and spills never touch the
// fp32 hot path.
template <int METHOD>
__device__ __noinline__ void eval_hessian_f64(const float4* __restrict__ src, int n, int n_groups, const EvalCtx* ctx,
                                              const MapView* map, double d2, double d1, double* s_warp_row) {
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * kAlignWarps + (threadIdx.x >> 5);
  const int warps_total = gridDim.x * kAlignWarps;
  double acc[kNV];
#pragma unroll
  for (int k = 0; k < kNV; ++k) acc[k] = 0.0;
  for (int g = warp_global; g < n_groups; g += warps_total) {
    const int i = (g << 5) + lane;
    if (i < n) eval_point_f64<METHOD>(__ldg(src + i), *ctx, *map, d2, d1, acc);
  }
  warp_flush(acc, s_warp_row);
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Fold the per-thread accumulators of a warp into the warp's fp64 sums (fixed shuffle tree) and clear them.
template <typename A>
__device__ __forceinline__ void warp_flush(A* acc, double* s_warp_row) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < kNV; ++k) {
    double v = static_cast<double>(acc[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) s_warp_row[k] += v;
    acc[k] = A(0);
  }
}

// ---- grid-wide sum of the CTA partials: ONE barrier per evaluation ---------------------------------
// 1. every CTA publishes its 29 fp64 partials, then one acq_rel atomic "arrive";
// 2. the LAST CTA to arrive (whichever it is) sums the partials in a fixed (slice, row) order — totals are
//    bit-reproducible — and publishes each total as two 64-bit words that carry a 32-bit tag next to each
//    32-bit half of the double (flag-in-data, as NCCL's LL protocol): readers need a single round trip;
// 3. 29 threads of every CTA poll their own total until both tags match this evaluation.
__device__ __forceinline__ unsigned int atom_add_acq_rel_gpu(unsigned int* p, unsigned int v) {
  unsigned int old;
  asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

constexpr int kRedRows = 20;  // partial rows per thread in flight while the last CTA reduces

__device__ __forceinline__ unsigned long long pack_lo(unsigned int tag, unsigned long long bits) {
  return (static_cast<unsigned long long>(tag) << 32) | (bits & 0xffffffffull);
}
__device__ __forceinline__ unsigned long long pack_hi(unsigned int tag, unsigned long long bits) {
  return (static_cast<unsigned long long>(tag) << 32) | (bits >> 32);
}

// Poll one tagged (lo, hi) word pair until both carry `tag`.  With peers involved the spin is bounded (10 s): a rank
// that never launches must not hang the GPU; on timeout the abort flag is raised and the solve ends with an error.
__device__ __forceinline__ double poll_tagged(const unsigned long long* slot, unsigned int tag, bool bounded,
                                              unsigned int* abort_flag) {
  unsigned long long w0, w1;
  unsigned long long t0 = 0;
  unsigned int spins = 0;
  while (true) {
    w0 = ld_volatile_u64(slot);
    w1 = ld_volatile_u64(slot + 1);
    if (static_cast<unsigned int>(w0 >> 32) == tag && static_cast<unsigned int>(w1 >> 32) == tag) break;
    if (bounded && ((++spins) & 0x3ffu) == 0u) {
      const unsigned long long now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > 10000000000ull || *reinterpret_cast<volatile unsigned int*>(abort_flag) != 0u) {
        atomicExch(abort_flag, 1u);
        return 0.0;
      }
    }
  }
  return __longlong_as_double(static_cast<long long>(((w1 & 0xffffffffull) << 32) | (w0 & 0xffffffffull)));
}

__device__ __forceinline__ void grid_allreduce(const double* s_block, double* s_tot, const AlignWorkspace& ws,
                                               unsigned int& epoch, int* s_flag, double (*s_red8)[kNVP],
                                               unsigned int launch_tag) {
  const unsigned int G = gridDim.x;
  const int W = ws.world;
  if (G == 1 && W == 1) {
    if (threadIdx.x < kNV) s_tot[threadIdx.x] = s_block[threadIdx.x];
    __syncthreads();
    ++epoch;
    return;
  }
  const unsigned int tag = launch_tag + epoch + 1u;
  const int par = epoch & 1u;
  bool last = true;
  if (G > 1) {
    if (threadIdx.x < kNV) ws.partials[(size_t)blockIdx.x * kNVP + threadIdx.x] = s_block[threadIdx.x];
    if (threadIdx.x == 31) ws.partials[(size_t)blockIdx.x * kNVP + 31] = __longlong_as_double(static_cast<long long>(globaltimer_ns()));  // profiling: arrival time of this CTA
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int ticket = atom_add_acq_rel_gpu(&ws.sync[0], 1u);  // release: the CTA's partials; acquire: everyone's
      *s_flag = (ticket == (epoch + 1u) * G - 1u) ? 1 : 0;
    }
    __syncthreads();
    last = (*s_flag != 0);
  }
  unsigned long long* slots = reinterpret_cast<unsigned long long*>(ws.totals);
  if (last) {  // last CTA of this GPU to arrive: every partial is visible
    double t = 0.0;
    if (G > 1) {
      if (threadIdx.x == 0) ws.totals[2 * kNVP + (epoch & 1u)] = __longlong_as_double(static_cast<long long>(globaltimer_ns()));  // profiling stamp
      const int k = threadIdx.x & 31, slice = threadIdx.x >> 5;  // lane = value, warp = slice of the CTA rows
      double s = 0.0;
      for (unsigned int b0 = 0; b0 < G; b0 += kAlignWarps * kRedRows) {
        double v[kRedRows];
#pragma unroll
        for (int i = 0; i < kRedRows; ++i) {
          const unsigned int b = b0 + slice + kAlignWarps * i;
          v[i] = (b < G && k < kNV) ? __ldcg(ws.partials + (size_t)b * kNVP + k) : 0.0;
        }
#pragma unroll
        for (int i = 0; i < kRedRows; ++i) s += v[i];
      }
      s_red8[slice][k] = s;
      __syncthreads();
      if (threadIdx.x < kNV) {
#pragma unroll
        for (int w = 0; w < kAlignWarps; ++w) t += s_red8[w][threadIdx.x];
      }
      __syncthreads();
    } else if (threadIdx.x < kNV) {
      t = s_block[threadIdx.x];
    }
    if (threadIdx.x < kNV) {
      const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(t));
      if (W == 1) {
        st_volatile_u64(slots + 2 * threadIdx.x, pack_lo(tag, bits));
        st_volatile_u64(slots + 2 * threadIdx.x + 1, pack_hi(tag, bits));
      } else {
        // this GPU's sums go straight into every rank's mailbox (P2P stores over NVLink; each 64-bit word is
        // self-validating, so no ordering between them is needed)
        const size_t off = (((size_t)par * kMaxRanks + ws.rank) * kNVP + threadIdx.x) * 2;
        for (int r = 0; r < W; ++r) {
          st_volatile_u64(ws.mail[r] + off, pack_lo(tag, bits));
          st_volatile_u64(ws.mail[r] + off + 1, pack_hi(tag, bits));
        }
      }
    }
  }
  if (W == 1) {
    if (threadIdx.x < kNV) s_tot[threadIdx.x] = poll_tagged(slots + 2 * threadIdx.x, tag, false, &ws.sync[3]);
    __syncthreads();
  } else {
    const int r = threadIdx.x >> 5, k = threadIdx.x & 31;
    if (r < W && k < kNV)
      s_red8[r][k] = poll_tagged(ws.mail[ws.rank] + (((size_t)par * kMaxRanks + r) * kNVP + k) * 2, tag, true, &ws.sync[3]);
    __syncthreads();
    if (threadIdx.x < kNV) {  // rank order: every GPU forms identical bits and takes the identical Newton step
      double t = 0.0;
      for (int rr = 0; rr < W; ++rr) t += s_red8[rr][threadIdx.x];
      s_tot[threadIdx.x] = t;
    }
    __syncthreads();
  }
  ++epoch;
}

// ---------------------------------------------------------------------------------------------
// pose -> (fp32 matrix, fp32 + fp64 angle tables), spread over one warp.
// Table entries are sums of at most two signed products of {1,cx,sx,cy,sy,cz,sz}; each lane evaluates
// its entries from a compact term list with un-fused fp64 multiplies/adds — the same values the
// straight-line formulas of computeAngleDerivatives give (ndt_omp_impl.hpp:329-393).
// ---------------------------------------------------------------------------------------------
// factor codes: 0:1  1:cx 2:sx 3:cy 4:sy 5:cz 6:sz ; term = sign * (f[a]*f[b])*f[c]; sign 0 => absent
struct TableTerm { int8_t s1, a1, b1, c1, s2, a2, b2, c2; };
#define TT(s1, a1, b1, c1, s2, a2, b2, c2) {s1, a1, b1, c1, s2, a2, b2, c2}
__device__ const TableTerm g_table_terms[69] = {
    // j_ang rows a..h (8 x 3)
    TT(-1, 2, 6, 0, +1, 1, 4, 5), TT(-1, 2, 5, 0, -1, 1, 4, 6), TT(-1, 1, 3, 0, 0, 0, 0, 0),   // a
    TT(+1, 1, 6, 0, +1, 2, 4, 5), TT(+1, 1, 5, 0, -1, 2, 4, 6), TT(-1, 2, 3, 0, 0, 0, 0, 0),   // b
    TT(-1, 4, 5, 0, 0, 0, 0, 0), TT(+1, 4, 6, 0, 0, 0, 0, 0), TT(+1, 3, 0, 0, 0, 0, 0, 0),     // c
    TT(+1, 2, 3, 5, 0, 0, 0, 0), TT(-1, 2, 3, 6, 0, 0, 0, 0), TT(+1, 2, 4, 0, 0, 0, 0, 0),     // d
    TT(-1, 1, 3, 5, 0, 0, 0, 0), TT(+1, 1, 3, 6, 0, 0, 0, 0), TT(-1, 1, 4, 0, 0, 0, 0, 0),     // e
    TT(-1, 3, 6, 0, 0, 0, 0, 0), TT(-1, 3, 5, 0, 0, 0, 0, 0), TT(0, 0, 0, 0, 0, 0, 0, 0),      // f
    TT(+1, 1, 5, 0, -1, 2, 4, 6), TT(-1, 1, 6, 0, -1, 2, 4, 5), TT(0, 0, 0, 0, 0, 0, 0, 0),    // g
    TT(+1, 2, 5, 0, +1, 1, 4, 6), TT(+1, 1, 4, 5, -1, 2, 6, 0), TT(0, 0, 0, 0, 0, 0, 0, 0),    // h
    // h_ang rows a2 a3 b2 b3 c2 c3 d1 d2 d3 e1 e2 e3 f1 f2 f3 (15 x 3), fp64 table (d1.z = -sy)
    TT(-1, 1, 6, 0, -1, 2, 4, 5), TT(-1, 1, 5, 0, +1, 2, 4, 6), TT(+1, 2, 3, 0, 0, 0, 0, 0),   // a2
    TT(-1, 2, 6, 0, +1, 1, 4, 5), TT(-1, 1, 4, 6, -1, 2, 5, 0), TT(-1, 1, 3, 0, 0, 0, 0, 0),   // a3
    TT(+1, 1, 3, 5, 0, 0, 0, 0), TT(-1, 1, 3, 6, 0, 0, 0, 0), TT(+1, 1, 4, 0, 0, 0, 0, 0),     // b2
    TT(+1, 2, 3, 5, 0, 0, 0, 0), TT(-1, 2, 3, 6, 0, 0, 0, 0), TT(+1, 2, 4, 0, 0, 0, 0, 0),     // b3
    TT(-1, 2, 5, 0, -1, 1, 4, 6), TT(+1, 2, 6, 0, -1, 1, 4, 5), TT(0, 0, 0, 0, 0, 0, 0, 0),    // c2
    TT(+1, 1, 5, 0, -1, 2, 4, 6), TT(-1, 2, 4, 5, -1, 1, 6, 0), TT(0, 0, 0, 0, 0, 0, 0, 0),    // c3
    TT(-1, 3, 5, 0, 0, 0, 0, 0), TT(+1, 3, 6, 0, 0, 0, 0, 0), TT(-1, 4, 0, 0, 0, 0, 0, 0),     // d1
    TT(-1, 2, 4, 5, 0, 0, 0, 0), TT(+1, 2, 4, 6, 0, 0, 0, 0), TT(+1, 2, 3, 0, 0, 0, 0, 0),     // d2
    TT(+1, 1, 4, 5, 0, 0, 0, 0), TT(-1, 1, 4, 6, 0, 0, 0, 0), TT(-1, 1, 3, 0, 0, 0, 0, 0),     // d3
    TT(+1, 4, 6, 0, 0, 0, 0, 0), TT(+1, 4, 5, 0, 0, 0, 0, 0), TT(0, 0, 0, 0, 0, 0, 0, 0),      // e1
    TT(-1, 2, 3, 6, 0, 0, 0, 0), TT(-1, 2, 3, 5, 0, 0, 0, 0), TT(0, 0, 0, 0, 0, 0, 0, 0),      // e2
    TT(+1, 1, 3, 6, 0, 0, 0, 0), TT(+1, 1, 3, 5, 0, 0, 0, 0), TT(0, 0, 0, 0, 0, 0, 0, 0),      // e3
    TT(-1, 3, 5, 0, 0, 0, 0, 0), TT(+1, 3, 6, 0, 0, 0, 0, 0), TT(0, 0, 0, 0, 0, 0, 0, 0),      // f1
    TT(-1, 1, 6, 0, -1, 2, 4, 5), TT(-1, 1, 5, 0, +1, 2, 4, 6), TT(0, 0, 0, 0, 0, 0, 0, 0),    // f2
    TT(-1, 2, 6, 0, +1, 1, 4, 5), TT(-1, 1, 4, 6, -1, 2, 5, 0), TT(0, 0, 0, 0, 0, 0, 0, 0)};   // f3
#undef TT

// Called by all threads of warp 0 (others return immediately).  s_trig: 16 doubles of shared scratch.
__device__ __forceinline__ void setup_pose_warp(const double* x_t, EvalCtx& ctx, float* final_T, double* s_trig,
                                                const TableTerm* s_terms, bool build_matrix) {
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  // lanes 0..5 : fp64 cos/sin of the three angles with the small-angle snap (ndt_omp_impl.hpp:292-326)
  // lanes 6..11: fp32 sin/cos of the fp32-cast angles for the matrix (Eigen::AngleAxis<float>)
  if (lane < 6) {
    const double a = x_t[3 + (lane >> 1)];
    double v;
    if (fabs(a) < 10e-5) v = (lane & 1) ? 0.0 : 1.0;
    else v = (lane & 1) ? sin(a) : cos(a);
    s_trig[1 + lane] = v;  // order cx sx cy sy cz sz = factor codes 1..6
  } else if (lane < 12) {
    const int l = lane - 6;
    const double a = static_cast<double>(static_cast<float>(x_t[3 + (l >> 1)]));
    s_trig[8 + l] = static_cast<double>(static_cast<float>((l & 1) ? sin(a) : cos(a)));  // cx sx cy sy cz sz (fp32 values)
  } else if (lane == 12) {
    s_trig[0] = 1.0;
  }
  __syncwarp();
  for (int e = lane; e < 69; e += 32) {
    const TableTerm t = s_terms[e];
    double v = 0.0;
    if (t.s1) v = __dmul_rn(__dmul_rn(s_trig[t.a1], s_trig[t.b1]), s_trig[t.c1]) * static_cast<double>(t.s1);
    if (t.s2) v = __dadd_rn(v, __dmul_rn(__dmul_rn(s_trig[t.a2], s_trig[t.b2]), s_trig[t.c2]) * static_cast<double>(t.s2));
    if (e < 24) {
      ctx.tab.jd[e / 3][e % 3] = v;
      ctx.tab.jf[e / 3][e % 3] = static_cast<float>(v);
    } else {
      const int r = (e - 24) / 3, c = (e - 24) % 3;
      ctx.tab.hd[r][c] = v;
      // Q2: the fp32 table carries +sy in row d1 (ndt_omp_impl.hpp:383), the fp64 table -sy (:361)
      ctx.tab.hf[r][c] = (r == 6 && c == 2) ? static_cast<float>(s_trig[4]) : static_cast<float>(v);
    }
  }
  if (build_matrix && lane == 0) {
    // Translation * Rx * Ry * Rz in fp32 from the fp32 trig values (same arithmetic as pose_to_matrix)
    const float cx = static_cast<float>(s_trig[8]), sx = static_cast<float>(s_trig[9]);
    const float cy = static_cast<float>(s_trig[10]), sy = static_cast<float>(s_trig[11]);
    const float cz = static_cast<float>(s_trig[12]), sz = static_cast<float>(s_trig[13]);
    float Rx[3][3], Ry[3][3], Rz[3][3], Rxy[3][3], R[3][3];
    angle_axis_from_sc(sx, cx, 0, Rx);
    angle_axis_from_sc(sy, cy, 1, Ry);
    angle_axis_from_sc(sz, cz, 2, Rz);
    mul33f(Rx, Ry, Rxy);
    mul33f(Rxy, Rz, R);
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) ctx.T[i * 4 + j] = R[i][j];
      ctx.T[i * 4 + 3] = static_cast<float>(x_t[i]);
    }
    for (int i = 0; i < 12; ++i) final_T[i] = ctx.T[i];  // final_transformation_ = T(x_t) (ndt_omp_impl.hpp:827-830)
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// optimiser state machine — scalar part (thread 0 of every CTA, identical inputs -> identical decisions)
// ---------------------------------------------------------------------------------------------
// Newton step H * delta = -g (ndt_omp_impl.hpp:127-129), rows spread over the lanes of warp 0:
// Gaussian elimination with partial pivoting in registers + shuffles (~1k cycles instead of the ~10k a
// single thread spends walking a 6x7 array in local memory).  A (numerically) rank-deficient H falls
// back to the one-sided Jacobi SVD pseudo-inverse (Eigen's JacobiSVD::solve semantics) in lane 0.
__device__ __forceinline__ void warp_newton_solve(SolverState& st) {
  const int lane = threadIdx.x & 31;
  const int r = lane < 6 ? lane : 5;
  double a[7];
#pragma unroll
  for (int j = 0; j < 6; ++j) a[j] = st.H[r * 6 + j];
  a[6] = -st.g[r];
  double pmin = 1e300, pmax = 0.0, amax = 0.0;
  bool ok = true;
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    double v = (lane >= c && lane < 6) ? fabs(a[c]) : -1.0;
    int idx = lane;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    const int piv = __shfl_sync(0xffffffffu, idx, 0);
    const double best = __shfl_sync(0xffffffffu, v, 0);
    if (!(best > 0.0) || isinf(best)) ok = false;
    pmin = fmin(pmin, best);
    pmax = fmax(pmax, best);
    amax = fmax(amax, best);
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      if (j < c) continue;
      const double from_piv = __shfl_sync(0xffffffffu, a[j], piv);
      const double from_c = __shfl_sync(0xffffffffu, a[j], c);
      if (lane == c) a[j] = from_piv;
      else if (lane == piv) a[j] = from_c;
    }
    const double pc = __shfl_sync(0xffffffffu, a[c], c);
    const double f = a[c] * (1.0 / pc);
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      if (j < c) continue;
      const double pj = __shfl_sync(0xffffffffu, a[j], c);
      if (lane > c && lane < 6) a[j] = (j == c) ? 0.0 : a[j] - f * pj;
    }
  }
  ok = ok && (pmin > 1e-9 * pmax);
  double x[6];
  double dinv = 1.0;  // 1 / diagonal of this lane's row, all lanes at once
#pragma unroll
  for (int j = 0; j < 6; ++j)
    if (j == r) dinv = 1.0 / a[j];
#pragma unroll
  for (int rr = 5; rr >= 0; --rr) {
    const double xr = __shfl_sync(0xffffffffu, a[6] * dinv, rr);
    x[rr] = xr;
    if (lane < rr) a[6] -= a[rr] * xr;
  }
  if (lane == 0) {
    if (ok) {
#pragma unroll
      for (int i = 0; i < 6; ++i) st.delta[i] = x[i];
    } else {
      double neg_g[6];
      for (int i = 0; i < 6; ++i) neg_g[i] = -st.g[i];
      svd_solve6(st.H, neg_g, st.delta);
    }
  }
  __syncwarp();
}

// After the Newton solve: step-length bookkeeping and start of the line search
// (ndt_omp_impl.hpp:131-142, 772-837).  Returns the next action.
__device__ __noinline__ int newton_post(SolverState& st, const AlignParams& prm) {
  while (true) {
    const double* delta = st.delta;  // a zero-length step leaves H and g unchanged: same solution next round
    double nrm = 0;
    for (int i = 0; i < 6; ++i) nrm += delta[i] * delta[i];
    nrm = sqrt(nrm);
    if (nrm == 0 || nrm != nrm) {  // ndt_omp_impl.hpp:134-139
      st.converged = (nrm == nrm) ? 1 : 0;
      return ACT_DONE;
    }
    for (int i = 0; i < 6; ++i) st.dir[i] = delta[i] / nrm;
    // computeStepLengthMT(p, dir, nrm, step_size, eps/2, ...)
    st.step_max = prm.step_size;
    st.step_min = prm.trans_eps / 2;
    st.phi_0 = -st.score;
    double dphi = 0;
    for (int i = 0; i < 6; ++i) dphi += st.g[i] * st.dir[i];
    st.d_phi_0 = -dphi;
    double a_ret = 0;
    bool evaluate = true;
    if (st.d_phi_0 >= 0) {
      if (st.d_phi_0 == 0) {
        a_ret = 0;  // "return 0": no trial is evaluated (ndt_omp_impl.hpp:787-788)
        evaluate = false;
      } else {
        st.d_phi_0 *= -1;
        for (int i = 0; i < 6; ++i) st.dir[i] *= -1;
      }
    }
    if (evaluate) {
      const double mu = 1.e-4;
      st.step_iterations = 0;
      st.a_l = 0; st.a_u = 0;
      st.f_l = st.phi_0 - st.phi_0 - mu * st.d_phi_0 * st.a_l;
      st.g_l = st.d_phi_0 - mu * st.d_phi_0;
      st.f_u = st.phi_0 - st.phi_0 - mu * st.d_phi_0 * st.a_u;
      st.g_u = st.d_phi_0 - mu * st.d_phi_0;
      st.interval_converged = ((st.step_max - st.step_min) < 0) ? 1 : 0;
      st.open_interval = 1;
      double a_t = nrm;
      a_t = std_min(a_t, st.step_max);
      a_t = std_max(a_t, st.step_min);
      st.a_t = a_t;
      for (int i = 0; i < 6; ++i) st.x_t[i] = st.p[i] + st.dir[i] * a_t;
      st.need_pose = 1;
      st.state = ST_MT_FIRST;
      return ACT_EVAL_FULL;
    }
    // zero-length step: finish this Newton iteration without an evaluation
    for (int i = 0; i < 6; ++i) st.last_dp[i] = st.dir[i] * a_ret;
    for (int i = 0; i < 6; ++i) st.p[i] += st.last_dp[i];
    if (st.nr_iterations > prm.max_iterations || (st.nr_iterations && (fabs(a_ret) < prm.trans_eps))) st.converged = 1;
    st.nr_iterations++;
    if (st.converged) return ACT_DONE;
  }
}

// Called after every evaluation with the reduced totals; returns the next action.
__device__ __noinline__ int advance(SolverState& st, const AlignParams& prm, const double* tot, int kind, TraceRec* trace) {
  st.need_pose = 0;
  if (kind != ACT_HESS_ONLY) {
    st.score = tot[0];
    for (int i = 0; i < 6; ++i) st.g[i] = tot[1 + i];
  }
  if (kind == ACT_EVAL_NOHESS) {
    for (int i = 0; i < 36; ++i) st.H[i] = 0.0;  // computeDerivatives(..., false) leaves H zeroed
  } else {
    int k = 7;
    for (int i = 0; i < 6; ++i)
      for (int j = i; j < 6; ++j) { st.H[i * 6 + j] = tot[k]; st.H[j * 6 + i] = tot[k]; ++k; }
  }
  st.n_hits += static_cast<long long>(tot[28]);
  if (kind == ACT_HESS_ONLY) st.n_hess++; else st.n_evals++;
  if (trace && st.n_trace < prm.trace_cap) {
    TraceRec& r = trace[st.n_trace];
    r.kind = kind - 1;
    r.pad = 0;
    for (int i = 0; i < 6; ++i) r.x[i] = (st.state == ST_INITIAL || st.state == ST_SINGLE) ? st.p[i] : st.x_t[i];
    r.a_t = (st.state == ST_INITIAL || st.state == ST_SINGLE) ? 0.0 : st.a_t;
    r.score = st.score;
  }
  st.n_trace++;

  const double mu = 1.e-4, nu = 0.9;
  switch (st.state) {
    case ST_SINGLE:
      return ACT_DONE;
    case ST_INITIAL:
      return ACT_SOLVE;
    case ST_MT_FIRST: {
      st.phi_t = -st.score;
      double d = 0;
      for (int i = 0; i < 6; ++i) d += st.g[i] * st.dir[i];
      st.d_phi_t = -d;
      st.psi_t = st.phi_t - st.phi_0 - mu * st.d_phi_0 * st.a_t;
      st.d_psi_t = st.d_phi_t - mu * st.d_phi_0;
      break;
    }
    case ST_MT_LOOP: {
      st.phi_t = -st.score;
      double d = 0;
      for (int i = 0; i < 6; ++i) d += st.g[i] * st.dir[i];
      st.d_phi_t = -d;
      st.psi_t = st.phi_t - st.phi_0 - mu * st.d_phi_0 * st.a_t;
      st.d_psi_t = st.d_phi_t - mu * st.d_phi_0;
      if (st.open_interval && (st.psi_t <= 0 && st.d_psi_t >= 0)) {  // ndt_omp_impl.hpp:894-905
        st.open_interval = 0;
        st.f_l = st.f_l + st.phi_0 - mu * st.d_phi_0 * st.a_l;
        st.g_l = st.g_l + mu * st.d_phi_0;
        st.f_u = st.f_u + st.phi_0 - mu * st.d_phi_0 * st.a_u;
        st.g_u = st.g_u + mu * st.d_phi_0;
      }
      if (st.open_interval)
        st.interval_converged = mt_update_interval(st.a_l, st.f_l, st.g_l, st.a_u, st.f_u, st.g_u, st.a_t, st.psi_t, st.d_psi_t) ? 1 : 0;
      else
        st.interval_converged = mt_update_interval(st.a_l, st.f_l, st.g_l, st.a_u, st.f_u, st.g_u, st.a_t, st.phi_t, st.d_phi_t) ? 1 : 0;
      st.step_iterations++;
      break;
    }
    case ST_MT_HESS:
      goto mt_finish;
  }
  // line-search loop condition (ndt_omp_impl.hpp:850)
  if (!st.interval_converged && st.step_iterations < 10 && !(st.psi_t <= 0 && st.d_phi_t <= -nu * st.d_phi_0)) {
    double a_t;
    if (st.open_interval)
      a_t = mt_trial_value(st.a_l, st.f_l, st.g_l, st.a_u, st.f_u, st.g_u, st.a_t, st.psi_t, st.d_psi_t);
    else
      a_t = mt_trial_value(st.a_l, st.f_l, st.g_l, st.a_u, st.f_u, st.g_u, st.a_t, st.phi_t, st.d_phi_t);
    a_t = std_min(a_t, st.step_max);
    a_t = std_max(a_t, st.step_min);
    st.a_t = a_t;
    for (int i = 0; i < 6; ++i) st.x_t[i] = st.p[i] + st.dir[i] * a_t;
    st.need_pose = 1;
    st.state = ST_MT_LOOP;
    return ACT_EVAL_NOHESS;
  }
  if (st.step_iterations) {  // ndt_omp_impl.hpp:928-929: same pose, same tables
    st.state = ST_MT_HESS;
    return ACT_HESS_ONLY;
  }
mt_finish: {
    // back in computeTransformation (ndt_omp_impl.hpp:143-164)
    const double a = st.a_t;
    for (int i = 0; i < 6; ++i) st.last_dp[i] = st.dir[i] * a;
    for (int i = 0; i < 6; ++i) st.p[i] = st.p[i] + st.last_dp[i];
    if (st.nr_iterations > prm.max_iterations || (st.nr_iterations && (fabs(a) < prm.trans_eps))) st.converged = 1;
    st.nr_iterations++;
    if (st.converged) return ACT_DONE;
    return ACT_SOLVE;
  }
}

// ---------------------------------------------------------------------------------------------
// fp32 derivative pass of one CTA — two phases per chunk of kChunk points ("hit queue"):
//   A  one thread per point: float4 load, transform, cell, all K hash probes in flight together, record
//      prefetch; the point's transformed position and its J_E / H_E coefficients are staged in shared
//      memory; every (point, voxel) hit is appended to the warp's queue segment (ballot compaction,
//      deterministic order);
//   B  the CTA's 256 threads consume the pooled queue densely, one hit per thread per round, so lanes
//      stay ~100 % busy whatever the per-point hit count (0..K) and the work is balanced over the CTA.
// ---------------------------------------------------------------------------------------------
constexpr int kPtVec = 7;              // float4 per staged point: x'(3) + pj(8) + ph(15) = 26 floats
constexpr int kRedStride = kAlignThreads + 8;  // padded row of the final fold (conflict-free banks)
constexpr int kFlushChunks = 2;        // fp32 run length bound: chunks before folding into fp64 (large clouds only)

// 32-point groups per warp per chunk: 2 for DIRECT1/7 (512 points staged per CTA: ~86 KB of shared memory, two
// CTAs per SM), 1 for DIRECT26 whose queue is 3.7x larger.
template <int METHOD>
__host__ __device__ constexpr int groups_per_warp() { return METHOD == 1 ? 1 : 2; }
template <int METHOD>
__host__ __device__ constexpr int queue_cap_per_warp() {
  return groups_per_warp<METHOD>() * 32 * (METHOD == 3 ? 1 : (METHOD == 2 ? 7 : 26));
}
template <int METHOD>
__host__ __device__ constexpr size_t eval_smem_bytes() {
  return (size_t)groups_per_warp<METHOD>() * kAlignThreads * kPtVec * 16 + (size_t)kAlignWarps * queue_cap_per_warp<METHOD>() * 8;
}

// Fold the per-thread fp32 partials into the fp64 CTA sums through shared memory (the staging buffers are free when
// this is called): thread (k, sub) adds 32 of the 256 partials of value k — groups of 4 in fp32 (short runs), then
// fp64 — and 8 lanes combine in a fixed tree.  ~8x fewer instructions than a 29-value warp shuffle tree per warp.
// Clears acc.  Must be called by all threads of the CTA.
__device__ __forceinline__ void fold_partials(float (&acc)[kNV], float* s_red, double* s_extra) {
#pragma unroll
  for (int k = 0; k < kNV; ++k) { s_red[k * kRedStride + threadIdx.x] = acc[k]; acc[k] = 0.0f; }
  __syncthreads();
  const int k = threadIdx.x >> 3, sub = threadIdx.x & 7;
  double s = 0.0;
  if (k < kNV) {
    const float* row = s_red + k * kRedStride + sub;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      s0 += static_cast<double>((row[8 * i] + row[8 * (i + 1)]) + (row[8 * (i + 2)] + row[8 * (i + 3)]));
      s1 += static_cast<double>((row[8 * (i + 4)] + row[8 * (i + 5)]) + (row[8 * (i + 6)] + row[8 * (i + 7)]));
    }
    s = s0 + s1;
  }
  s += __shfl_down_sync(0xffffffffu, s, 4, 8);
  s += __shfl_down_sync(0xffffffffu, s, 2, 8);
  s += __shfl_down_sync(0xffffffffu, s, 1, 8);
  if (k < kNV && sub == 0) s_extra[k] += s;
}

struct HitOperands {  // what phase B needs for one hit before the arithmetic starts
  uint2 e;
  float4 a, b, c;  // the 48 hot bytes of the voxel record
};

// q-th entry of the pooled queue: the per-warp segments are walked forward only (q grows by 256 per round)
__device__ __forceinline__ HitOperands fetch_hit(const uint2* s_q, const int* s_wcount, int qcap, int q, int& seg, int& base,
                                                 const VoxelRecord* records) {
  while (q >= base + s_wcount[seg]) { base += s_wcount[seg]; ++seg; }
  HitOperands h;
  h.e = s_q[seg * qcap + (q - base)];
  const float4* R = reinterpret_cast<const float4*>(records + h.e.x);
  h.a = __ldg(R);
  h.b = __ldg(R + 1);
  h.c = __ldg(R + 2);
  return h;
}

template <int METHOD, bool HESS>
__device__ __forceinline__ void eval_chunked_f32(const float4* __restrict__ src, int n, int n_groups, const EvalCtx& ctx,
                                                 const MapView& m, float d2f, float d1f, float4* s_pts, uint2* s_q,
                                                 int* s_wcount, double (*s_warp)[kNVP], double* s_extra,
                                                 unsigned long long* s_dbg) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int K = num_offsets<METHOD>();
  constexpr int GPW = groups_per_warp<METHOD>();
  constexpr int QCAP = queue_cap_per_warp<METHOD>();
  const unsigned int lt_mask = (1u << lane) - 1u;
  float acc[kNV];
#pragma unroll
  for (int k = 0; k < kNV; ++k) acc[k] = 0.0f;
  int chunks_since = 0;
  // this CTA's 32-point groups: g = blockIdx.x + j * gridDim.x (interleaved over the grid)
  const int groups_mine = (n_groups > (int)blockIdx.x) ? (n_groups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  for (int j0 = 0; j0 < groups_mine; j0 += kAlignWarps * GPW) {
    // ---------------- phase A ----------------
    int wcount = 0;
    uint2* qw = s_q + warp * QCAP;
#pragma unroll
    for (int s = 0; s < GPW; ++s) {
      const int j = j0 + s * kAlignWarps + warp;
      if (j < groups_mine) {  // warp-uniform
        const int i = ((blockIdx.x + j * gridDim.x) << 5) + lane;
        const int slot = s * kAlignThreads + threadIdx.x;
        const bool valid = i < n;
        const float4 pt = valid ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        float tx, ty, tz;
        transform_point(ctx.T, pt.x, pt.y, pt.z, tx, ty, tz);
        // getNeighborhoodAtPoint (…_impl.hpp:379-381): cell = floor(x' / leaf), fp32 DIVISION (Q8)
        const int ix = static_cast<int>(floorf(__fdiv_rn(tx, m.leaf[0])));
        const int iy = static_cast<int>(floorf(__fdiv_rn(ty, m.leaf[1])));
        const int iz = static_cast<int>(floorf(__fdiv_rn(tz, m.leaf[2])));
        if constexpr (METHOD != 1) {
          int rec[K];
          probe_cells<K>(m, ix, iy, iz, rec);
#pragma unroll
          for (int k = 0; k < K; ++k) {
            const bool hit = valid && rec[k] >= 0;
            const unsigned int mask = __ballot_sync(0xffffffffu, hit);
            if (hit) qw[wcount + __popc(mask & lt_mask)] = make_uint2(static_cast<unsigned int>(rec[k]), slot);
            wcount += __popc(mask);
          }
        } else {
#pragma unroll 1
          for (int k = 0; k < K; ++k) {
            int dx, dy, dz;
            get_offset<METHOD>(k, dx, dy, dz);
            const int rec = valid ? probe_cell(m, ix + dx, iy + dy, iz + dz) : -1;
            const bool hit = rec >= 0;
            const unsigned int mask = __ballot_sync(0xffffffffu, hit);
            if (hit) {
              qw[wcount + __popc(mask & lt_mask)] = make_uint2(static_cast<unsigned int>(rec), slot);
              prefetch_l2(m.records + rec);
            }
            wcount += __popc(mask);
          }
        }
        // computePointDerivatives (fp32 overload, ndt_omp_impl.hpp:398-440): depends on the ORIGINAL point only
        float pj[8], ph[15];
#pragma unroll
        for (int r = 0; r < 8; ++r) pj[r] = ctx.tab.jf[r][0] * pt.x + ctx.tab.jf[r][1] * pt.y + ctx.tab.jf[r][2] * pt.z;
        float4* P = s_pts + slot * kPtVec;
        P[0] = make_float4(tx, ty, tz, pj[0]);
        P[1] = make_float4(pj[1], pj[2], pj[3], pj[4]);
        if (HESS) {
#pragma unroll
          for (int r = 0; r < 15; ++r) ph[r] = ctx.tab.hf[r][0] * pt.x + ctx.tab.hf[r][1] * pt.y + ctx.tab.hf[r][2] * pt.z;
          P[2] = make_float4(pj[5], pj[6], pj[7], ph[0]);
          P[3] = make_float4(ph[1], ph[2], ph[3], ph[4]);
          P[4] = make_float4(ph[5], ph[6], ph[7], ph[8]);
          P[5] = make_float4(ph[9], ph[10], ph[11], ph[12]);
          P[6] = make_float4(ph[13], ph[14], 0.f, 0.f);
        } else {
          P[2] = make_float4(pj[5], pj[6], pj[7], 0.f);
        }
      }
    }
    if (lane == 0) s_wcount[warp] = wcount;
    __syncthreads();
    if (threadIdx.x == 0 && j0 == 0) s_dbg[0] = globaltimer_ns();
    // ---------------- phase B ----------------
    int Q = 0;
#pragma unroll
    for (int w = 0; w < kAlignWarps; ++w) Q += s_wcount[w];
    int seg = 0, base = 0;
    for (int q = threadIdx.x; q < Q; q += kAlignThreads) {
      const HitOperands cur = fetch_hit(s_q, s_wcount, QCAP, q, seg, base, m.records);
      const float4* P = s_pts + cur.e.y * kPtVec;
      const float4 v0 = P[0], v1 = P[1], v2 = P[2];
      float pj[8] = {v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z};
      float ph[15];
      if (HESS) {
        const float4 v3 = P[3], v4 = P[4], v5 = P[5], v6 = P[6];
        ph[0] = v2.w; ph[1] = v3.x; ph[2] = v3.y; ph[3] = v3.z; ph[4] = v3.w; ph[5] = v4.x; ph[6] = v4.y; ph[7] = v4.z;
        ph[8] = v4.w; ph[9] = v5.x; ph[10] = v5.y; ph[11] = v5.z; ph[12] = v5.w; ph[13] = v6.x; ph[14] = v6.y;
      }
      // x_trans = fl32(double(x') - mean) (ndt_omp_impl.hpp:259-262, 492) via the exact hi/lo split of the mean
      const float r0 = __fsub_rn(__fsub_rn(v0.x, cur.a.x), cur.a.w);
      const float r1 = __fsub_rn(__fsub_rn(v0.y, cur.a.y), cur.b.x);
      const float r2 = __fsub_rn(__fsub_rn(v0.z, cur.a.z), cur.b.y);
      acc[28] += 1.0f;
      hit_contribution<float, true, HESS>(r0, r1, r2, cur.b.z, cur.b.w, cur.c.x, cur.c.y, cur.c.z, cur.c.w, pj, ph, d2f, d1f, acc);
    }
    __syncthreads();  // the next chunk overwrites the staging buffers
    if (threadIdx.x == 0 && j0 == 0) s_dbg[1] = globaltimer_ns();
    if (++chunks_since == kFlushChunks) {  // only reached by large clouds: bound the fp32 run length
      fold_partials(acc, reinterpret_cast<float*>(s_pts), s_extra);
      __syncthreads();
      chunks_since = 0;
    }
  }
  fold_partials(acc, reinterpret_cast<float*>(s_pts), s_extra);
}

// ---------------------------------------------------------------------------------------------
// the persistent kernel
// ---------------------------------------------------------------------------------------------
template <int METHOD>
__global__ void __launch_bounds__(kAlignThreads, 2)
ndt_align_kernel(const float4* __restrict__ src, const MapView map, const AlignParams prm, const AlignWorkspace ws) {
  __shared__ SolverState st;
  __shared__ EvalCtx ctx;
  __shared__ double s_warp[kAlignWarps][kNVP];
  __shared__ double s_block[kNVP];
  __shared__ double s_tot[kNVP];
  __shared__ double s_extra[kNVP];
  __shared__ unsigned long long s_dbg[4];
  __shared__ double s_trig[16];
  __shared__ TableTerm s_terms[69];
  __shared__ int s_action;
  __shared__ int s_flag;
  __shared__ int s_wcount[kAlignWarps];
  extern __shared__ float4 dyn_smem[];
  float4* s_pts = dyn_smem;                                                                      // [GPW * 256][kPtVec]
  uint2* s_q = reinterpret_cast<uint2*>(dyn_smem + groups_per_warp<METHOD>() * kAlignThreads * kPtVec);  // [kAlignWarps][QCAP]

  const unsigned long long t_kernel_begin = globaltimer_ns();
  const int n = prm.n_source;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // 32-point groups interleaved over all warps of the grid (deterministic static balance)
  const int n_groups = (n + 31) >> 5;
  const int warp_global = blockIdx.x * kAlignWarps + warp;
  const int warps_total = gridDim.x * kAlignWarps;
  unsigned int epoch = 0;

  if (threadIdx.x == 0) {
    // align() prologue + computeTransformation up to the first computeDerivatives (ndt_omp_impl.hpp:83-119)
    for (int i = 0; i < 6; ++i) st.p[i] = prm.p0[i];
    st.score = 0;
    for (int i = 0; i < 6; ++i) st.g[i] = 0;
    for (int i = 0; i < 36; ++i) st.H[i] = 0;
    st.n_hits = 0;
    st.nr_iterations = 0;
    st.converged = 0;
    st.n_evals = st.n_hess = st.n_trace = 0;
    st.a_t = 0;
    st.need_pose = 0;
    for (int i = 0; i < 6; ++i) st.x_t[i] = prm.p0[i];
    for (int i = 0; i < 12; ++i) {
      ctx.T[i] = prm.T0[i];       // first evaluation: source transformed by the guess matrix itself (:100)
      st.final_T[i] = prm.T0[i];  // final_transformation_ = guess (:98) or Identity (align())
    }
    for (int i = 0; i < 6; ++i) st.last_dp[i] = 0.0;
    if (prm.mode == MODE_ALIGN) {
      st.state = ST_INITIAL;
      s_action = ACT_EVAL_FULL;
    } else if (prm.mode == MODE_EVAL) {
      st.state = ST_SINGLE;
      s_action = prm.eval_hessian ? ACT_EVAL_FULL : ACT_EVAL_NOHESS;
    } else {
      st.state = ST_SINGLE;
      s_action = ACT_HESS_ONLY;
    }
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 69) s_terms[threadIdx.x - 64] = g_table_terms[threadIdx.x - 64];
  if (threadIdx.x == 32) ctx.tab = prm.tab0;  // tables for p0 (computed on the host); matrix = T0
  __syncthreads();

  const float d2f = static_cast<float>(prm.d2);
  const float d1f = static_cast<float>(prm.d1);
  while (true) {
    const int action = s_action;
    if (action == ACT_DONE) break;
    unsigned long long t_start = 0, t_local = 0, t_reduced = 0;
    const bool timing = (blockIdx.x == 0 && threadIdx.x == 0 && ws.trace != nullptr);
    if (timing) t_start = globaltimer_ns();
    if (threadIdx.x < kAlignWarps * kNVP) (&s_warp[0][0])[threadIdx.x] = 0.0;
    if (threadIdx.x < kNVP) s_extra[threadIdx.x] = 0.0;
    __syncthreads();

    if (action == ACT_HESS_ONLY) {
      eval_hessian_f64<METHOD>(src, n, n_groups, &ctx, &map, prm.d2, prm.d1, s_warp[warp]);
    } else if (action == ACT_EVAL_FULL) {
      eval_chunked_f32<METHOD, true>(src, n, n_groups, ctx, map, d2f, d1f, s_pts, s_q, s_wcount, s_warp, s_extra, s_dbg);
    } else {
      eval_chunked_f32<METHOD, false>(src, n, n_groups, ctx, map, d2f, d1f, s_pts, s_q, s_wcount, s_warp, s_extra, s_dbg);
    }
    __syncthreads();
    if (threadIdx.x < kNV) {
      double s = 0;
#pragma unroll
      for (int w = 0; w < kAlignWarps; ++w) s += s_warp[w][threadIdx.x];
      s_block[threadIdx.x] = s + s_extra[threadIdx.x];
    }
    __syncthreads();
    if (timing) t_local = globaltimer_ns();
    grid_allreduce(s_block, s_tot, ws, epoch, &s_flag, s_warp, prm.launch_tag);
    if (ws.world > 1 && *reinterpret_cast<volatile unsigned int*>(&ws.sync[3]) != 0u) break;  // a peer never answered (uniform: checked after a barrier)
    if (timing) t_reduced = globaltimer_ns();
    unsigned long long t_last_arrive = 0;
    if (timing && gridDim.x > 1) t_last_arrive = static_cast<unsigned long long>(__double_as_longlong(__ldcg(ws.totals + 2 * kNVP + ((epoch - 1u) & 1u))));
    int slot = 0;
    unsigned long long t_d0 = 0, t_d1 = 0, t_d2 = 0;
    if (threadIdx.x == 0) {
      slot = st.n_trace;
      s_action = advance(st, prm, s_tot, action, blockIdx.x == 0 ? ws.trace : nullptr);
    }
    __syncthreads();
    if (timing) t_d0 = t_d1 = t_d2 = globaltimer_ns();
    if (s_action == ACT_SOLVE) {  // block-uniform
      if (warp == 0) warp_newton_solve(st);
      __syncthreads();
      if (timing) t_d1 = globaltimer_ns();
      if (threadIdx.x == 0) s_action = newton_post(st, prm);
      __syncthreads();
      if (timing) t_d2 = globaltimer_ns();
    }
    if (st.need_pose) setup_pose_warp(st.x_t, ctx, st.final_T, s_trig, s_terms, /*build_matrix=*/true);
    if (timing && slot < prm.trace_cap) {
      TraceRec& r = ws.trace[slot];
      r.t_start = t_start; r.t_local = t_local; r.t_reduced = t_reduced; r.t_advanced = globaltimer_ns();
      r.t_dbg[0] = t_d0; r.t_dbg[1] = t_d1; r.t_dbg[2] = t_d2; r.t_dbg[3] = t_last_arrive;
      r.t_phase[0] = s_dbg[0]; r.t_phase[1] = s_dbg[1];
    }
    __syncthreads();
  }

  if (blockIdx.x == 0 && threadIdx.x == 0) {
    AlignResultDev& r = *ws.result;
    for (int i = 0; i < 12; ++i) r.final_T[i] = st.final_T[i];
    for (int i = 0; i < 6; ++i) r.last_dp[i] = st.last_dp[i];  // transformation_ = T(last_dp) is formed on the host
    r.converged = st.converged;
    r.iterations = st.nr_iterations;
    r.n_evals = st.n_evals;
    r.n_hess = st.n_hess;
    r.trans_probability = st.score / static_cast<double>(ws.world > 1 ? ws.n_source_total : (long long)n);  // ndt_omp_impl.hpp:136, 170
    r.aborted = (ws.world > 1) ? static_cast<int32_t>(*reinterpret_cast<volatile unsigned int*>(&ws.sync[3])) : 0;
    const bool trial = (st.n_evals > 1) || prm.mode != MODE_ALIGN;
    for (int i = 0; i < 6; ++i) r.final_pose[i] = trial ? st.x_t[i] : prm.p0[i];
    r.final_score = st.score;
    r.totals[0] = st.score;
    for (int i = 0; i < 6; ++i) r.totals[1 + i] = st.g[i];
    for (int i = 0; i < 36; ++i) r.totals[7 + i] = st.H[i];
    r.n_hits = st.n_hits;
    r.n_trace = st.n_trace;
    r.t_kernel_begin = t_kernel_begin;
    r.t_kernel_end = globaltimer_ns();
  }
  // Leave the barrier words zeroed for the next launch: the LAST CTA to get here resets them (every other
  // CTA has finished its last barrier by then); sync[2] is a wrapping exit counter that resets itself.
  if (threadIdx.x == 0 && gridDim.x > 1) {
    const unsigned int old = atomicInc(&ws.sync[2], gridDim.x - 1);
    if (old == gridDim.x - 1) {
      ws.sync[0] = 0u;
      ws.sync[3] = 0u;
      __threadfence();
    }
  }
}

}  // namespace ndtb200
