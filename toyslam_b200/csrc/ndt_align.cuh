// ndt_align.cuh — the registration hot path as ONE persistent kernel:
//
//   derivative pass   computeDerivatives / updateDerivatives / computePointDerivatives
//                     (ndt_omp_impl.hpp:179-285, 398-440, 484-537), fp32 per-hit math, fp64 accumulation
//   Hessian-only pass computeHessian / updateHessian (ndt_omp_impl.hpp:540-645), fp64 math, fp64 tables
//   Newton + line search  computeTransformation / computeStepLengthMT (ndt_omp_impl.hpp:80-171, 772-932)
//
// Design (B200-first, not the reference's per-point-slot + serial-sum structure):
//   * one thread per source point, float4 loads, in-register fp32 transform (never materialises
//     trans_cloud), DIRECT1/7/26 probes into the HBM voxel hash, 64-byte Gaussian records;
//   * 28 (+1 hit counter) fp64 accumulators per thread -> warp shuffle tree -> shared memory ->
//     one partial per CTA -> the LAST-ARRIVING CTA sums the partials in CTA order (bit-reproducible)
//     and publishes the 28 totals; every CTA then runs the identical Newton / More-Thuente step
//     redundantly, so a single grid barrier per evaluation is the only synchronisation and the
//     whole align() never returns to the host;
//   * the kernel is launched cooperatively (all CTAs co-resident) with gridDim = #SMs x occupancy.
#pragma once
#include "common.cuh"
#include "ndt_solve.cuh"

namespace ndtb200 {

constexpr int kAlignThreads = 256;
constexpr int kNV = 29;   // score, g[6], H upper triangle[21], hit count
constexpr int kNVP = 32;  // padded row length of the partial / total buffers

enum { ACT_DONE = 0, ACT_EVAL_FULL = 1, ACT_EVAL_NOHESS = 2, ACT_HESS_ONLY = 3 };
enum { ST_INITIAL = 0, ST_MT_FIRST = 1, ST_MT_LOOP = 2, ST_MT_HESS = 3, ST_SINGLE = 4 };
enum { MODE_ALIGN = 0, MODE_EVAL = 1, MODE_HESSIAN = 2 };

struct TraceRec {
  int32_t kind;  // 0 derivatives+hessian, 1 derivatives only, 2 hessian only
  int32_t pad;
  double x[6];
  double a_t;
  double score;
};

struct AlignResultDev {
  float final_T[12];
  float incr_T[12];
  int32_t converged, iterations, n_evals, n_hess;
  double trans_probability;
  double final_pose[6];
  double final_score;
  double totals[43];  // score, g[6], H[36] of the last evaluation
  long long n_hits;
  int32_t n_trace, pad;
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
};

struct AlignWorkspace {
  double* partials;      // [gridDim][kNVP]
  double* totals;        // [2][kNVP]
  unsigned int* sync;    // [0] arrive counter, [1] epoch flag   (zeroed before each launch)
  AlignResultDev* result;
  TraceRec* trace;
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
  float final_T[12], incr_T[12];
};

// ---------------------------------------------------------------------------------------------
// one (point, voxel) contribution: updateDerivatives (fp32, T=float) / updateHessian (fp64, T=double)
// pj = j_ang * x (8 values), ph = h_ang * x (15 values); r = x' - mean; c = inverse covariance.
// acc layout: [0] score, [1..6] gradient, [7..27] Hessian upper triangle row-major.
// ---------------------------------------------------------------------------------------------
template <typename T, bool SCORE_GRAD, bool HESS>
__device__ __forceinline__ void hit_contribution(const T r0, const T r1, const T r2, const T c00, const T c01,
                                                 const T c02, const T c11, const T c12, const T c22, const T* pj,
                                                 const T* ph, const T d2, const double d1, double* acc) {
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
    if (SCORE_GRAD) acc[0] += static_cast<double>(e) * (-d1);
    w = static_cast<float>(static_cast<double>(e2) * d1);
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
    acc[1] += static_cast<double>(w * s0);
    acc[2] += static_cast<double>(w * s1);
    acc[3] += static_cast<double>(w * s2);
    acc[4] += static_cast<double>(w * s3);
    acc[5] += static_cast<double>(w * s4);
    acc[6] += static_cast<double>(w * s5);
  }
  if (HESS) {
    // C * J_i for the three rotational columns (J_3 = (0,j0,j1), J_4 = (j2,j3,j4), J_5 = (j5,j6,j7))
    const T v3x = c01 * pj[0] + c02 * pj[1], v3y = c11 * pj[0] + c12 * pj[1], v3z = c12 * pj[0] + c22 * pj[1];
    const T v4x = c00 * pj[2] + c01 * pj[3] + c02 * pj[4], v4y = c01 * pj[2] + c11 * pj[3] + c12 * pj[4],
            v4z = c02 * pj[2] + c12 * pj[3] + c22 * pj[4];
    const T v5x = c00 * pj[5] + c01 * pj[6] + c02 * pj[7], v5y = c01 * pj[5] + c11 * pj[6] + c12 * pj[7],
            v5z = c02 * pj[5] + c12 * pj[6] + c22 * pj[7];
    const T md2 = -d2;
#define NDTB200_H(idx, si, sj, extra) acc[7 + idx] += static_cast<double>(w * (md2 * si * sj + (extra)));
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
__constant__ int8_t c_off7[7][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};

template <int METHOD>
__device__ __forceinline__ constexpr int num_offsets() {
  return METHOD == 3 ? 1 : (METHOD == 2 ? 7 : 26);
}

template <int METHOD>
__device__ __forceinline__ void get_offset(int k, int& dx, int& dy, int& dz) {
  if (METHOD == 3) { dx = dy = dz = 0; }
  else if (METHOD == 2) { dx = c_off7[k][0]; dy = c_off7[k][1]; dz = c_off7[k][2]; }
  else { dx = c_off26[k][0]; dy = c_off26[k][1]; dz = c_off26[k][2]; }
}

// Probe one neighbour cell (voxel_grid_covariance_omp_impl.hpp:388-400): bounds test, key, hash find.
__device__ __forceinline__ int probe_cell(const MapView& m, int cx, int cy, int cz) {
  if (cx < m.min_b[0] || cx > m.max_b[0] || cy < m.min_b[1] || cy > m.max_b[1] || cz < m.min_b[2] || cz > m.max_b[2])
    return -1;
  const int key = (cx - m.min_b[0]) * m.mul[0] + (cy - m.min_b[1]) * m.mul[1] + (cz - m.min_b[2]) * m.mul[2];
  return map_find(m, key);
}

// ---------------------------------------------------------------------------------------------
// all hits of one source point, fp32 path (computeDerivatives inner loop, ndt_omp_impl.hpp:207-275)
// ---------------------------------------------------------------------------------------------
template <int METHOD, bool HESS>
__device__ __forceinline__ void eval_point_f32(const float4 pt, const EvalCtx& c, const MapView& m, const float d2f,
                                               const double d1, double* acc) {
  float tx, ty, tz;
  transform_point(c.T, pt.x, pt.y, pt.z, tx, ty, tz);
  // getNeighborhoodAtPoint (…_impl.hpp:379-381): cell = floor(x' / leaf), fp32 DIVISION (Q8)
  const int ix = static_cast<int>(floorf(__fdiv_rn(tx, m.leaf[0])));
  const int iy = static_cast<int>(floorf(__fdiv_rn(ty, m.leaf[1])));
  const int iz = static_cast<int>(floorf(__fdiv_rn(tz, m.leaf[2])));
  // computePointDerivatives (fp32 overload, ndt_omp_impl.hpp:398-440): depends on the ORIGINAL point only
  float pj[8], ph[15];
#pragma unroll
  for (int r = 0; r < 8; ++r) pj[r] = c.tab.jf[r][0] * pt.x + c.tab.jf[r][1] * pt.y + c.tab.jf[r][2] * pt.z;
  if (HESS) {
#pragma unroll
    for (int r = 0; r < 15; ++r) ph[r] = c.tab.hf[r][0] * pt.x + c.tab.hf[r][1] * pt.y + c.tab.hf[r][2] * pt.z;
  }
  const double dtx = tx, dty = ty, dtz = tz;
  constexpr int K = num_offsets<METHOD>();
#pragma unroll(METHOD == 1 ? 1 : K)
  for (int k = 0; k < K; ++k) {
    int dx, dy, dz;
    get_offset<METHOD>(k, dx, dy, dz);
    const int rec = probe_cell(m, ix + dx, iy + dy, iz + dz);
    if (rec < 0) continue;
    const VoxelRecord* R = m.records + rec;
    const double2 m01 = __ldg(reinterpret_cast<const double2*>(R));
    const float4 b = __ldg(reinterpret_cast<const float4*>(R) + 1);
    const float4 cc = __ldg(reinterpret_cast<const float4*>(R) + 2);
    const double m2 = __hiloint2double(__float_as_int(b.y), __float_as_int(b.x));
    // x_trans = double(x') - mean, THEN cast to fp32 (ndt_omp_impl.hpp:259-262, 492)
    const float r0 = static_cast<float>(dtx - m01.x);
    const float r1 = static_cast<float>(dty - m01.y);
    const float r2 = static_cast<float>(dtz - m2);
    hit_contribution<float, true, HESS>(r0, r1, r2, b.z, b.w, cc.x, cc.y, cc.z, cc.w, pj, ph, d2f, d1, acc);
    acc[28] += 1.0;
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
    const double r0 = static_cast<double>(tx) - __ldg(&R->mean[0]);
    const double r1 = static_cast<double>(ty) - __ldg(&R->mean[1]);
    const double r2 = static_cast<double>(tz) - __ldg(&R->mean[2]);
    hit_contribution<double, false, true>(r0, r1, r2, __ldg(ic), __ldg(ic + 1), __ldg(ic + 2), __ldg(ic + 3),
                                          __ldg(ic + 4), __ldg(ic + 5), pj, ph, d2, d1, acc);
    acc[28] += 1.0;
  }
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// acc[kNV] per thread -> s_block[kNV] (block sum, fixed order)
__device__ __forceinline__ void block_reduce(double* acc, double (*s_warp)[kNVP], double* s_block) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < kNV; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) s_warp[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < kNV) {
    double s = 0;
#pragma unroll
    for (int w = 0; w < kAlignThreads / 32; ++w) s += s_warp[w][threadIdx.x];
    s_block[threadIdx.x] = s;
  }
  __syncthreads();
}

// s_block (this CTA's sums) -> s_tot (sum over all CTAs, identical bits in every CTA).
__device__ __forceinline__ void grid_allreduce(const double* s_block, double* s_tot, const AlignWorkspace& ws,
                                               unsigned int& epoch, int* s_flag) {
  const unsigned int G = gridDim.x;
  if (G == 1) {
    if (threadIdx.x < kNV) s_tot[threadIdx.x] = s_block[threadIdx.x];
    __syncthreads();
    return;
  }
  if (threadIdx.x < kNV) {
    ws.partials[(size_t)blockIdx.x * kNVP + threadIdx.x] = s_block[threadIdx.x];
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(&ws.sync[0], 1u);
    *s_flag = (ticket == (epoch + 1u) * G - 1u) ? 1 : 0;
  }
  __syncthreads();
  const int par = epoch & 1u;
  if (*s_flag) {  // last CTA to arrive: every partial is visible
    __threadfence();
    if (threadIdx.x < kNV) {
      double s = 0;
      const double* p = ws.partials + threadIdx.x;
#pragma unroll 8
      for (unsigned int b = 0; b < G; ++b) s += __ldcg(p + (size_t)b * kNVP);
      ws.totals[par * kNVP + threadIdx.x] = s;
      __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) atomicExch(&ws.sync[1], epoch + 1u);
  }
  if (threadIdx.x == 0) {
    while (ld_acquire_u32(&ws.sync[1]) < epoch + 1u) {
    }
  }
  __syncthreads();
  if (threadIdx.x < kNV) s_tot[threadIdx.x] = __ldcg(ws.totals + par * kNVP + threadIdx.x);
  __syncthreads();
  ++epoch;
}

// ---------------------------------------------------------------------------------------------
// optimiser state machine (thread 0 of every CTA, identical inputs -> identical decisions)
// ---------------------------------------------------------------------------------------------
__device__ inline void set_eval_pose(SolverState& st, EvalCtx& ctx, const double x_t[6]) {
  // final_transformation_ = T(x_t) in fp32 (ndt_omp_impl.hpp:827-830, 871-874)
  pose_to_matrix(x_t, ctx.T);
  for (int i = 0; i < 12; ++i) st.final_T[i] = ctx.T[i];
  compute_angle_tables(x_t, ctx.tab);
}

// Newton step + start of the line search (ndt_omp_impl.hpp:121-142, 772-837).  Returns the next action.
__device__ inline int newton_top(SolverState& st, EvalCtx& ctx, const AlignParams& prm) {
  while (true) {
    double neg_g[6], delta[6];
    for (int i = 0; i < 6; ++i) neg_g[i] = -st.g[i];
    newton_solve6(st.H, neg_g, delta);
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
    double a_ret;
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
      set_eval_pose(st, ctx, st.x_t);
      st.state = ST_MT_FIRST;
      return ACT_EVAL_FULL;
    }
    // zero-length step: finish this Newton iteration without an evaluation
    double dp[6];
    for (int i = 0; i < 6; ++i) dp[i] = st.dir[i] * a_ret;
    pose_to_matrix(dp, st.incr_T);
    for (int i = 0; i < 6; ++i) st.p[i] += dp[i];
    if (st.nr_iterations > prm.max_iterations || (st.nr_iterations && (fabs(a_ret) < prm.trans_eps))) st.converged = 1;
    st.nr_iterations++;
    if (st.converged) return ACT_DONE;
  }
}

// Called after every evaluation with the reduced totals; returns the next action.
__device__ inline int advance(SolverState& st, EvalCtx& ctx, const AlignParams& prm, const double* tot, int kind,
                              TraceRec* trace) {
  // unpack totals
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
      return newton_top(st, ctx, prm);
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
    set_eval_pose(st, ctx, st.x_t);
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
    double dp[6];
    for (int i = 0; i < 6; ++i) dp[i] = st.dir[i] * a;
    pose_to_matrix(dp, st.incr_T);
    for (int i = 0; i < 6; ++i) st.p[i] = st.p[i] + dp[i];
    if (st.nr_iterations > prm.max_iterations || (st.nr_iterations && (fabs(a) < prm.trans_eps))) st.converged = 1;
    st.nr_iterations++;
    if (st.converged) return ACT_DONE;
    return newton_top(st, ctx, prm);
  }
}

// ---------------------------------------------------------------------------------------------
// the persistent kernel
// ---------------------------------------------------------------------------------------------
template <int METHOD>
__global__ void __launch_bounds__(kAlignThreads)
ndt_align_kernel(const float4* __restrict__ src, const MapView map, const AlignParams prm, const AlignWorkspace ws) {
  __shared__ SolverState st;
  __shared__ EvalCtx ctx;
  __shared__ double s_warp[kAlignThreads / 32][kNVP];
  __shared__ double s_block[kNVP];
  __shared__ double s_tot[kNVP];
  __shared__ int s_action;
  __shared__ int s_flag;

  const int n = prm.n_source;
  // contiguous chunk of source points per CTA (multiple of 32 so warps stay coalesced)
  const int chunk = ((n + gridDim.x - 1) / gridDim.x + 31) & ~31;
  const int begin = min(n, (int)blockIdx.x * chunk);
  const int end = min(n, begin + chunk);
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
    for (int i = 0; i < 6; ++i) st.x_t[i] = prm.p0[i];
    for (int i = 0; i < 12; ++i) {
      ctx.T[i] = prm.T0[i];     // first evaluation: source transformed by the guess matrix itself (:100)
      st.final_T[i] = prm.T0[i];  // final_transformation_ = guess (:98) or Identity (align())
      st.incr_T[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    }
    compute_angle_tables(prm.p0, ctx.tab);
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
  __syncthreads();

  const float d2f = static_cast<float>(prm.d2);
  while (true) {
    const int action = s_action;
    if (action == ACT_DONE) break;
    double acc[kNV];
#pragma unroll
    for (int k = 0; k < kNV; ++k) acc[k] = 0.0;
    if (action == ACT_EVAL_FULL) {
      for (int i = begin + threadIdx.x; i < end; i += kAlignThreads)
        eval_point_f32<METHOD, true>(__ldg(src + i), ctx, map, d2f, prm.d1, acc);
    } else if (action == ACT_EVAL_NOHESS) {
      for (int i = begin + threadIdx.x; i < end; i += kAlignThreads)
        eval_point_f32<METHOD, false>(__ldg(src + i), ctx, map, d2f, prm.d1, acc);
    } else {
      for (int i = begin + threadIdx.x; i < end; i += kAlignThreads)
        eval_point_f64<METHOD>(__ldg(src + i), ctx, map, prm.d2, prm.d1, acc);
    }
    block_reduce(acc, s_warp, s_block);
    grid_allreduce(s_block, s_tot, ws, epoch, &s_flag);
    if (threadIdx.x == 0) s_action = advance(st, ctx, prm, s_tot, action, blockIdx.x == 0 ? ws.trace : nullptr);
    __syncthreads();
  }

  if (blockIdx.x == 0 && threadIdx.x == 0) {
    AlignResultDev& r = *ws.result;
    for (int i = 0; i < 12; ++i) { r.final_T[i] = st.final_T[i]; r.incr_T[i] = st.incr_T[i]; }
    r.converged = st.converged;
    r.iterations = st.nr_iterations;
    r.n_evals = st.n_evals;
    r.n_hess = st.n_hess;
    r.trans_probability = st.score / static_cast<double>(n);  // ndt_omp_impl.hpp:136, 170
    const bool trial = (st.n_evals > 1) || prm.mode != MODE_ALIGN;
    for (int i = 0; i < 6; ++i) r.final_pose[i] = trial ? st.x_t[i] : prm.p0[i];
    r.final_score = st.score;
    r.totals[0] = st.score;
    for (int i = 0; i < 6; ++i) r.totals[1 + i] = st.g[i];
    for (int i = 0; i < 36; ++i) r.totals[7 + i] = st.H[i];
    r.n_hits = st.n_hits;
    r.n_trace = st.n_trace;
    r.pad = 0;
  }
}

}  // namespace ndtb200
