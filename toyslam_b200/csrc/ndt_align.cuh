// ndt_align.cuh — the registration hot path as ONE persistent kernel:
//
//   derivative pass   computeDerivatives / updateDerivatives / computePointDerivatives
//                     (ndt_omp_impl.hpp:179-285, 398-440, 484-537), fp32 per-hit math
//   Hessian-only pass computeHessian / updateHessian (ndt_omp_impl.hpp:540-645), fp64 math, fp64 tables
//   Newton + line search  computeTransformation / computeStepLengthMT (ndt_omp_impl.hpp:80-171, 772-932)
//
// Design (B200-first, not the reference's per-point-slot + serial-sum structure):
//   * one thread per source point, float4 loads, in-register fp32 transform (never materialises
//     trans_cloud), DIRECT1/7/26 probes into the voxel index (direct-mapped cell table, or the open-addressing
//     hash for grids whose table does not fit; all probes of a point issued together), 64-byte Gaussian
//     records prefetched, then consumed hit by hit;
//   * warps take 32-point groups interleaved over the whole grid, so every CTA sees the same mix of
//     dense and empty regions of the scan (static, deterministic load balance);
//   * accumulation: a thread adds the (fp32) contributions of at most kFlushPoints points in fp32,
//     then the warp folds them into fp64 (shuffle tree) — fp64 everywhere a long sum is formed;
//     warp sums -> one fp64 partial row per CTA (rows double-buffered by evaluation parity);
//   * grid reduction, single GPU: one acq_rel "arrive" per CTA on a counter whose target for evaluation e is
//     (e + 1) * G; as soon as it is reached EVERY CTA sums the G rows itself in the same fixed (slice, row) order:
//     identical bits everywhere, no broadcast.  The counter and the exit counter reset themselves at kernel exit;
//   * grid reduction, source-sharded multi-GPU (world > 1): the last CTA of a rank to arrive sums the rank's rows
//     and stores the 29 totals as tagged (flag-in-data) 64-bit words into every rank's mailbox (P2P stores over
//     NVLink); all CTAs poll their own mailbox and add the W rows in rank order.  tag = launch_seq << 12 | (epoch+1);
//   * every CTA then runs the identical Newton / More-Thuente step (scalar part in one thread, the
//     trigonometry and the 69 table entries spread over a warp), so ONE grid barrier per evaluation
//     is the only synchronisation and the whole align() never returns to the host;
//   * launched cooperatively (all CTAs co-resident), gridDim = #SMs (x emulated ranks' share, see VirtualRank).
#pragma once
#include "common.cuh"
#include "ndt_solve.cuh"

namespace ndtb200 {

// CTA shapes of the persistent kernel (template parameter THREADS), always 64 registers per thread:
//   latency shape    1024 threads, one CTA per SM: a single align() owns the GPU (148 CTAs: few rows in the grid
//                    reduction, few arrivals on the barrier word);
//   throughput shape  256 threads, one CTA per SM PER KERNEL: up to four independent aligns (different handles /
//                    streams) are co-resident on every SM, so one solve's barrier + Newton-step latency is covered by
//                    the others' derivative passes (ndtb200_align_batch).
constexpr int kThreadsLatency = 1024;
#ifndef NDTB200_THROUGHPUT_THREADS
#define NDTB200_THROUGHPUT_THREADS 256
#endif
constexpr int kThreadsThroughput = NDTB200_THROUGHPUT_THREADS;
__host__ __device__ constexpr int min_blocks_for(int threads) { return 1024 / threads; }
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
  int32_t rot;          // profiling: rotate the CTA -> point-group assignment by this many CTAs
  uint32_t launch_tag;  // launch sequence number << 12 (kMaxEpochsSharded evaluations per launch): makes the mailbox tags unique per launch
  AngleTables tab0;     // angle tables of p0 (host-computed: the kernel prologue has no trigonometry)
};

constexpr int kMaxRanks = 8;  // GPUs of one NVSwitch node
constexpr unsigned int kMaxEpochsSharded = 4095;  // mailbox tag = launch_seq << 12 | (epoch + 1): evaluations per sharded launch

// One emulated rank of a source-sharded solve run as ONE cooperative launch on ONE GPU (the CTAs of the grid are
// divided evenly between the ranks): every rank has its own source slice, partial rows, barrier words, result block
// and mailbox, exactly as a rank on its own GPU, and the ranks exchange their 29 sums through the same tagged
// mailbox stores and polls.  This is how a one-GPU box exercises the multi-GPU exchange code (separate launches that
// wait on one another are not guaranteed to be co-resident on one device).
struct VirtualRank {
  const float4* src;
  int32_t n_source, pad;
  double* partials;
  double* totals;
  unsigned int* sync;
  struct AlignResultDev* result;
};

struct AlignWorkspace {
  double* partials;      // [2][gridDim][kNVP] (single GPU: rows double-buffered by evaluation parity)
  double* totals;        // [3][kNVP] (profiling stamps of the sharded path)
  unsigned int* sync;    // [0] arrive counter, [2] wrapping exit counter, [3] abort flag (zeroed once, self-resetting)
  AlignResultDev* result;
  TraceRec* trace;
  // source-sharded multi-GPU solve: every rank holds a slice of the source and the full map; the
  // 29 per-evaluation sums are exchanged by direct stores into every peer's mailbox over NVLink.
  int32_t world, rank;
  unsigned long long* mail[kMaxRanks];  // mail[r]: rank r's mailbox [2 parities][kMaxRanks][kNVP][2 words] (own or IPC-mapped peer memory)
  long long n_source_total;              // points of the whole (unsigned) source cloud
  int32_t vranks, pad;                   // > 1: `world` emulated ranks inside this one launch
  const VirtualRank* vr;                 // [vranks] (device memory)
  // align()'s output cloud written by the solve itself (single-rank MODE_ALIGN launches whose caller asked for it):
  // out[i] = final_transformation * out_src[i] in the CALLER's point order (out_src is the unsorted source)
  float4* out;
  const float4* out_src;
};

// What a CTA needs to know about the rank it works for (shared memory; filled once at kernel start).
struct RankCtx {
  double* partials;
  double* totals;
  unsigned int* sync;
  AlignResultDev* result;
  unsigned int G, bid;  // CTAs of this rank, index of this CTA inside the rank
  int32_t rank, pad;
};

struct EvalCtx {
  float T[12];
  float4 tf[23];  // the fp32 tables as (c0, c1, c2, 0) rows: j_ang rows 0..7, h_ang rows 8..22 (one LDS.128 per row)
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
  double H_ls[36];    // Hessian of the LAST line-search trial, computeHessian's table variant (see unpack_hessian_warp)
  float final_T[12];
  // copies of the launch parameters the step functions need (the parameter struct itself is never address-taken, so it
  // stays in the constant bank instead of being copied to every thread's local memory)
  double step_size, trans_eps;
  int32_t max_iterations, trace_cap;
};

__constant__ int8_t c_off26[26][3] = {
    // pcl::getAllNeighborCellIndices(): 13 "half" offsets then their negations (centre excluded, Q7)
    {-1, -1, -1}, {-1, 0, -1}, {-1, 1, -1}, {0, -1, -1}, {0, 0, -1}, {0, 1, -1}, {1, -1, -1}, {1, 0, -1}, {1, 1, -1},
    {-1, -1, 0}, {0, -1, 0}, {1, -1, 0}, {-1, 0, 0},
    {1, 1, 1}, {1, 0, 1}, {1, -1, 1}, {0, 1, 1}, {0, 0, 1}, {0, -1, 1}, {-1, 1, 1}, {-1, 0, 1}, {-1, -1, 1},
    {1, 1, 0}, {0, 1, 0}, {-1, 1, 0}, {1, 0, 0}};

// METHOD: the reference's enum values (ndt_omp.h:52-57): 0 KDTREE, 1 DIRECT26, 2 DIRECT7, 3 DIRECT1
template <int METHOD>
__device__ __forceinline__ constexpr int num_offsets() {
  return METHOD == 3 ? 1 : (METHOD == 2 ? 7 : (METHOD == 1 ? 26 : 27));
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
  } else if (METHOD == 1) {
    dx = c_off26[k][0]; dy = c_off26[k][1]; dz = c_off26[k][2];
  } else {  // KDTREE: the 27 cells a centroid within one resolution of the point can sit in
    dx = k % 3 - 1; dy = (k / 3) % 3 - 1; dz = k / 9 - 1;
  }
}

// Probe one neighbour cell (voxel_grid_covariance_omp_impl.hpp:388-400): bounds test, key, hash find.
__device__ __forceinline__ int probe_cell(const MapView& m, int cx, int cy, int cz) {
  if (cx < m.min_b[0] || cx > m.max_b[0] || cy < m.min_b[1] || cy > m.max_b[1] || cz < m.min_b[2] || cz > m.max_b[2])
    return -1;
  const int key = (cx - m.min_b[0]) * m.mul[0] + (cy - m.min_b[1]) * m.mul[1] + (cz - m.min_b[2]) * m.mul[2];
  return map_find(m, key);
}

// KDTREE mode: radiusSearch(x', resolution) over the voxel-centroid cloud (voxel_grid_covariance_omp.h:476-505,
// ndt_omp_impl.hpp:234-236).  FLANN L2_Simple: fp32 squared distance, strict "< r^2".  The centroid cloud holds every
// leaf that had >= min_points when it was pushed (…_impl.hpp:311-317) — also the ones rejected afterwards
// (count == -1, zero / infinite icov): the reference does not re-check nr_points on this path (quirk Q10).
__device__ __forceinline__ int probe_cell_kdtree(const MapView& m, int cx, int cy, int cz, float tx, float ty, float tz) {
  if (cx < m.min_b[0] || cx > m.max_b[0] || cy < m.min_b[1] || cy > m.max_b[1] || cz < m.min_b[2] || cz > m.max_b[2])
    return -1;
  const int key = (cx - m.min_b[0]) * m.mul[0] + (cy - m.min_b[1]) * m.mul[1] + (cz - m.min_b[2]) * m.mul[2];
  const int v = __ldg(m.cell_all + key);
  if (v < 0) return -1;
  const int count = __ldg(&m.records[v].count);
  if (!(count >= m.min_points || count == -1)) return -1;
  const float4 c = __ldg(m.centroids + v);
  const float dx = tx - c.x, dy = ty - c.y, dz = tz - c.z;
  const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  return (d < m.kd_r2) ? v : -1;
}

template <int METHOD>
__device__ __forceinline__ int probe_neighbour(const MapView& m, int ix, int iy, int iz, int k, float tx, float ty, float tz) {
  int dx, dy, dz;
  get_offset<METHOD>(k, dx, dy, dz);
  if constexpr (METHOD == 0) return probe_cell_kdtree(m, ix + dx, iy + dy, iz + dz, tx, ty, tz);
  else return probe_cell(m, ix + dx, iy + dy, iz + dz);
}

// A non-finite transformed point (NaN / inf in the source, or 0 * inf in the transform) has no neighbourhood in the
// reference: int(floor(NaN)) is INT_MIN on the reference's platform (x86 cvttss2si), far outside every grid, so
// getNeighborhoodAtPoint returns nothing and the point adds exactly 0 (voxel_grid_covariance_omp_impl.hpp:379-391,
// ndt_omp_impl.hpp:506-507 never runs).  CUDA's float->int conversion maps NaN to 0, i.e. to a real cell: skip the
// point instead.  (A finite sum of three finite coordinates that overflows belongs to a point outside every grid too.)
__device__ __forceinline__ bool point_is_finite(float tx, float ty, float tz) {
  return fabsf(__fadd_rn(__fadd_rn(tx, ty), tz)) <= 3.402823466e+38f;
}

// All K probes of a point (DIRECT1 / DIRECT7) with their loads in flight together.  In-bounds tests are unsigned
// compares of the min-relative cell coordinates; neighbour keys are key0 +- mul[axis].
template <int K>
__device__ __forceinline__ void probe_cells(const MapView& m, int ix, int iy, int iz, int (&rec)[K]) {
  static_assert(K == 1 || K == 7, "unrolled probe is for DIRECT1 / DIRECT7");
  const int rx = ix - m.min_b[0], ry = iy - m.min_b[1], rz = iz - m.min_b[2];
  const unsigned int ex = static_cast<unsigned int>(m.max_b[0] - m.min_b[0]);
  const unsigned int ey = static_cast<unsigned int>(m.max_b[1] - m.min_b[1]);
  const unsigned int ez = static_cast<unsigned int>(m.max_b[2] - m.min_b[2]);
  // (unsigned)(r + d) <= e  <=>  min_b <= i + d <= max_b ; an empty map has max_b < min_b: e wraps to 0xffffffff only
  // if max_b - min_b == -1, so the empty view sets mul = 0 and its table holds -1 everywhere (see make_view)
  const bool inx = static_cast<unsigned int>(rx) <= ex, iny = static_cast<unsigned int>(ry) <= ey,
             inz = static_cast<unsigned int>(rz) <= ez;
  const int key0 = rx * m.mul[0] + ry * m.mul[1] + rz * m.mul[2];
  bool ok[K];
  int key[K];
  ok[0] = inx && iny && inz;
  key[0] = key0;
  if constexpr (K == 7) {
    // DIRECT7 order (voxel_grid_covariance_omp_impl.hpp:423-430): centre, +x, -x, +y, -y, +z, -z
    ok[1] = (static_cast<unsigned int>(rx + 1) <= ex) && iny && inz; key[1] = key0 + m.mul[0];
    ok[2] = (static_cast<unsigned int>(rx - 1) <= ex) && iny && inz; key[2] = key0 - m.mul[0];
    ok[3] = inx && (static_cast<unsigned int>(ry + 1) <= ey) && inz; key[3] = key0 + m.mul[1];
    ok[4] = inx && (static_cast<unsigned int>(ry - 1) <= ey) && inz; key[4] = key0 - m.mul[1];
    ok[5] = inx && iny && (static_cast<unsigned int>(rz + 1) <= ez); key[5] = key0 + m.mul[2];
    ok[6] = inx && iny && (static_cast<unsigned int>(rz - 1) <= ez); key[6] = key0 - m.mul[2];
  }
  if (m.dense != nullptr) {  // uniform
#pragma unroll
    for (int k = 0; k < K; ++k) {
      NDT_CHECK(!ok[k] || (key[k] >= 0 && static_cast<unsigned long long>(key[k]) < m.n_cells));
      rec[k] = ok[k] ? __ldg(m.dense + key[k]) : -1;
      NDT_CHECK(rec[k] >= -1 && (rec[k] < 0 || static_cast<uint32_t>(rec[k]) < m.n_records));
    }
  } else {
    HashSlot slot[K];
    uint32_t h[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      h[k] = hash_key(static_cast<uint32_t>(key[k]), m.hash_shift);
      slot[k] = ok[k] ? __ldg(m.hash + h[k]) : NDTB200_HASH_EMPTY;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      int r = -1;
      HashSlot sl = slot[k];
      uint32_t hh = h[k];
      while (sl != NDTB200_HASH_EMPTY) {
        if (static_cast<uint32_t>(sl) == static_cast<uint32_t>(key[k])) { r = static_cast<int>(sl >> 32); break; }
        hh = (hh + 1) & m.hash_mask;
        sl = __ldg(m.hash + hh);
      }
      rec[k] = r;
    }
  }
  // (an L2 prefetch of the found records here was measured at -2 % on the city-map workload and +-0 on c2: the record
  // loads follow a few instructions later anyway)
}

// The optimiser state and the evaluation context of the CTA live in file-scope shared memory, so the out-of-line step
// functions reach them with shared-space loads/stores at fixed addresses (not generic pointers).
__shared__ SolverState g_st;
__shared__ EvalCtx g_ctx;

// ---------------------------------------------------------------------------------------------
// fp64 Hessian-only pass (computeHessian / updateHessian, ndt_omp_impl.hpp:540-645): fp64 math, fp64 tables (Q2: -sy),
// same per-point factorisation as the fp32 pass (see point_f32 below): A = sum w u, M = sum w (C - d2 u u^T),
// H_ij = J_i^T M J_j + A . H_E[i][j].
// ---------------------------------------------------------------------------------------------
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

template <int METHOD>
__device__ __forceinline__ void point_hessian_f64(const float4 pt, const EvalCtx& c, const MapView& m, const double d2,
                                                  const double d1, double* acc /*[22]: H upper triangle, hits*/) {
  float tx, ty, tz;
  transform_point(c.T, pt.x, pt.y, pt.z, tx, ty, tz);
  if (!point_is_finite(tx, ty, tz)) return;  // no neighbourhood, no contribution (see point_is_finite)
  const int ix = lookup_cell(tx, m.leaf[0], m.inv_leaf[0]);
  const int iy = lookup_cell(ty, m.leaf[1], m.inv_leaf[1]);
  const int iz = lookup_cell(tz, m.leaf[2], m.inv_leaf[2]);
  constexpr int K = num_offsets<METHOD>();
  double A[3] = {0, 0, 0}, M[6] = {0, 0, 0, 0, 0, 0};
  int nh = 0;
  auto hit = [&](int rec) {
    ++nh;
    const float4* R = reinterpret_cast<const float4*>(m.records + rec);
    const float4 ra = __ldg(R), rb = __ldg(R + 1);  // mean_hi[3], mean_lo[3] (+ 2 icov words)
    const double2* ic = reinterpret_cast<const double2*>(m.icov64 + (size_t)rec * 6);
    const double2 i0 = __ldg(ic), i1 = __ldg(ic + 1), i2 = __ldg(ic + 2);
    const double r0 = static_cast<double>(tx) - (static_cast<double>(ra.x) + static_cast<double>(ra.w));
    const double r1 = static_cast<double>(ty) - (static_cast<double>(ra.y) + static_cast<double>(rb.x));
    const double r2 = static_cast<double>(tz) - (static_cast<double>(ra.z) + static_cast<double>(rb.y));
    const double c00 = i0.x, c01 = i0.y, c02 = i1.x, c11 = i1.y, c12 = i2.x, c22 = i2.y;
    const double u0 = c00 * r0 + c01 * r1 + c02 * r2;
    const double u1 = c01 * r0 + c11 * r1 + c12 * r2;
    const double u2 = c02 * r0 + c12 * r1 + c22 * r2;
    const double q = r0 * u0 + r1 * u1 + r2 * u2;
    const double e2 = d2 * exp(-d2 * q / 2);  // ndt_omp_impl.hpp:622-626
    if (!(e2 <= 1.0 && e2 >= 0.0)) return;
    const double w = e2 * d1;
    A[0] += w * u0; A[1] += w * u1; A[2] += w * u2;
    const double kk = -d2 * w;
    const double t0 = kk * u0, t1 = kk * u1, t2 = kk * u2;
    M[0] += w * c00 + t0 * u0;
    M[1] += w * c01 + t0 * u1;
    M[2] += w * c02 + t0 * u2;
    M[3] += w * c11 + t1 * u1;
    M[4] += w * c12 + t1 * u2;
    M[5] += w * c22 + t2 * u2;
  };
  if constexpr (METHOD == 2 || METHOD == 3) {
    int rec[K];
    probe_cells<K>(m, ix, iy, iz, rec);  // all lookups of the point in flight together
#pragma unroll
    for (int k = 0; k < K; ++k)
      if (rec[k] >= 0) hit(rec[k]);
  } else {
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
      const int rec = probe_neighbour<METHOD>(m, ix, iy, iz, k, tx, ty, tz);
      if (rec >= 0) hit(rec);
    }
  }
  if (nh == 0) return;
  acc[21] += static_cast<double>(nh);
  const double x = pt.x, y = pt.y, z = pt.z;
  double pj[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) pj[r] = x * c.tab.jd[r][0] + y * c.tab.jd[r][1] + z * c.tab.jd[r][2];
  const double m3x = M[1] * pj[0] + M[2] * pj[1], m3y = M[3] * pj[0] + M[4] * pj[1], m3z = M[4] * pj[0] + M[5] * pj[1];
  const double m4x = M[0] * pj[2] + M[1] * pj[3] + M[2] * pj[4], m4y = M[1] * pj[2] + M[3] * pj[3] + M[4] * pj[4],
               m4z = M[2] * pj[2] + M[4] * pj[3] + M[5] * pj[4];
  const double m5x = M[0] * pj[5] + M[1] * pj[6] + M[2] * pj[7], m5y = M[1] * pj[5] + M[3] * pj[6] + M[4] * pj[7],
               m5z = M[2] * pj[5] + M[4] * pj[6] + M[5] * pj[7];
  acc[0] += M[0]; acc[1] += M[1]; acc[2] += M[2]; acc[3] += m3x; acc[4] += m4x; acc[5] += m5x;
  acc[6] += M[3]; acc[7] += M[4]; acc[8] += m3y; acc[9] += m4y; acc[10] += m5y;
  acc[11] += M[5]; acc[12] += m3z; acc[13] += m4z; acc[14] += m5z;
#define NDTB200_PH(r) (x * c.tab.hd[r][0] + y * c.tab.hd[r][1] + z * c.tab.hd[r][2])
  acc[15] += (pj[0] * m3y + pj[1] * m3z) + (A[1] * NDTB200_PH(0) + A[2] * NDTB200_PH(1));
  acc[16] += (pj[0] * m4y + pj[1] * m4z) + (A[1] * NDTB200_PH(2) + A[2] * NDTB200_PH(3));
  acc[17] += (pj[0] * m5y + pj[1] * m5z) + (A[1] * NDTB200_PH(4) + A[2] * NDTB200_PH(5));
  acc[18] += (pj[2] * m4x + pj[3] * m4y + pj[4] * m4z) + (A[0] * NDTB200_PH(6) + A[1] * NDTB200_PH(7) + A[2] * NDTB200_PH(8));
  acc[19] += (pj[2] * m5x + pj[3] * m5y + pj[4] * m5z) + (A[0] * NDTB200_PH(9) + A[1] * NDTB200_PH(10) + A[2] * NDTB200_PH(11));
  acc[20] += (pj[5] * m5x + pj[6] * m5y + pj[7] * m5z) + (A[0] * NDTB200_PH(12) + A[1] * NDTB200_PH(13) + A[2] * NDTB200_PH(14));
#undef NDTB200_PH
}

// The whole fp64 Hessian-only pass of one warp.  Rare (only after a multi-trial line search) and register hungry:
// kept out of line so its register pressure never touches the fp32 hot path.  Writes the warp's fp64 sums
// (fixed shuffle tree) to s_warp_row[0..31] (entries 0..6: score / gradient are not produced by this pass).
template <int METHOD>
__device__ __noinline__ void eval_hessian_f64(const float4* __restrict__ src, int n, int n_groups, int warp_global,
                                              int warps_total, const EvalCtx* ctx, const MapView* map, double d2, double d1,
                                              double* s_warp_row) {
  const int lane = threadIdx.x & 31;
  double acc[22];
#pragma unroll
  for (int k = 0; k < 22; ++k) acc[k] = 0.0;
  for (int g = warp_global; g < n_groups; g += warps_total) {
    const int i = (g << 5) + lane;
    if (i < n) point_hessian_f64<METHOD>(__ldg(src + i), *ctx, *map, d2, d1, acc);
  }
  if (lane < 7) s_warp_row[lane] = 0.0;
  if (lane >= 29) s_warp_row[lane] = 0.0;
#pragma unroll
  for (int k = 0; k < 22; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) s_warp_row[7 + k] = v;
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

// partial rows per thread in flight while the last CTA reduces
__host__ __device__ constexpr int red_rows_for(int nwarps) { return nwarps >= 32 ? 8 : 20; }

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

template <int NW>
__device__ __forceinline__ void grid_allreduce(const double* s_block, double* s_tot, const AlignWorkspace& ws, const RankCtx& rc,
                                               unsigned int& epoch, int* s_flag, double (*s_red8)[kNVP],
                                               unsigned int launch_tag) {
  const unsigned int G = rc.G, bid = rc.bid;
  const int W = ws.world;
  NDT_CHECK(bid < G && W >= 1 && W <= kMaxRanks && rc.rank >= 0 && rc.rank < W);
  if (G == 1 && W == 1) {
    if (threadIdx.x < kNV) s_tot[threadIdx.x] = s_block[threadIdx.x];
    __syncthreads();
    ++epoch;
    return;
  }
  if (W == 1) {
    // Single GPU: EVERY CTA sums all partial rows itself (same fixed (slice, row) order -> identical bits everywhere)
    // as soon as the arrival counter shows the grid complete: one round trip less than "the last CTA reduces and
    // publishes".  Rows are double-buffered by evaluation parity: a CTA can only reach evaluation e+1's arrive after
    // it finished reading evaluation e's rows, so a row is never overwritten while somebody still reads it.
    double* rows = rc.partials + (size_t)(epoch & 1u) * G * kNVP;
    if (threadIdx.x < kNV) rows[(size_t)bid * kNVP + threadIdx.x] = s_block[threadIdx.x];
    if (threadIdx.x == 31) rows[(size_t)bid * kNVP + 31] = __longlong_as_double(static_cast<long long>(globaltimer_ns()));  // profiling: arrival time
    if (threadIdx.x == 30) { unsigned int smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); rows[(size_t)bid * kNVP + 30] = static_cast<double>(smid); }
    __syncthreads();
    if (threadIdx.x == 0) {
      atom_add_acq_rel_gpu(&rc.sync[0], 1u);  // release: this CTA's row
      const unsigned int target = (epoch + 1u) * G;
      while (ld_acquire_u32(&rc.sync[0]) < target) {}  // acquire: everyone's rows
      s_tot[31] = __longlong_as_double(static_cast<long long>(globaltimer_ns()));  // profiling: grid complete
    }
    __syncthreads();
    constexpr int kRedRows = red_rows_for(NW);
    const int k = threadIdx.x & 31, slice = threadIdx.x >> 5;  // lane = value, warp = slice of the CTA rows
    double sum = 0.0;
    for (unsigned int b0 = 0; b0 < G; b0 += NW * kRedRows) {
      double v[kRedRows];
#pragma unroll
      for (int i = 0; i < kRedRows; ++i) {
        const unsigned int b = b0 + slice + NW * i;
        v[i] = (b < G && k < kNV) ? __ldcg(rows + (size_t)b * kNVP + k) : 0.0;
      }
#pragma unroll
      for (int i = 0; i < kRedRows; ++i) sum += v[i];
    }
    s_red8[slice][k] = sum;
    __syncthreads();
    if (threadIdx.x < kNV) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < NW; ++w) t += s_red8[w][threadIdx.x];
      s_tot[threadIdx.x] = t;
    }
    __syncthreads();
    ++epoch;
    return;
  }
  // ---- source-sharded solve (W ranks, one per GPU or emulated inside this launch) ----
  const unsigned int tag = launch_tag + epoch + 1u;
  const int par = epoch & 1u;
  bool last = true;
  if (G > 1) {
    if (threadIdx.x < kNV) rc.partials[(size_t)bid * kNVP + threadIdx.x] = s_block[threadIdx.x];
    if (threadIdx.x == 31) rc.partials[(size_t)bid * kNVP + 31] = __longlong_as_double(static_cast<long long>(globaltimer_ns()));  // profiling: arrival time of this CTA
    if (threadIdx.x == 30) { unsigned int smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); rc.partials[(size_t)bid * kNVP + 30] = static_cast<double>(smid); }
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int ticket = atom_add_acq_rel_gpu(&rc.sync[0], 1u);  // release: the CTA's partials; acquire: everyone's
      *s_flag = (ticket == (epoch + 1u) * G - 1u) ? 1 : 0;
    }
    __syncthreads();
    last = (*s_flag != 0);
  }
  if (last) {  // last CTA of this rank to arrive: every partial of the rank is visible
    double t = 0.0;
    if (G > 1) {
      if (threadIdx.x == 0) rc.totals[2 * kNVP + (epoch & 1u)] = __longlong_as_double(static_cast<long long>(globaltimer_ns()));  // profiling stamp
      const int k = threadIdx.x & 31, slice = threadIdx.x >> 5;  // lane = value, warp = slice of the CTA rows
      double s = 0.0;
      constexpr int kRedRows = red_rows_for(NW);
      for (unsigned int b0 = 0; b0 < G; b0 += NW * kRedRows) {
        double v[kRedRows];
#pragma unroll
        for (int i = 0; i < kRedRows; ++i) {
          const unsigned int b = b0 + slice + NW * i;
          v[i] = (b < G && k < kNV) ? __ldcg(rc.partials + (size_t)b * kNVP + k) : 0.0;
        }
#pragma unroll
        for (int i = 0; i < kRedRows; ++i) s += v[i];
      }
      s_red8[slice][k] = s;
      __syncthreads();
      if (threadIdx.x < kNV) {
#pragma unroll
        for (int w = 0; w < NW; ++w) t += s_red8[w][threadIdx.x];
      }
    } else if (threadIdx.x < kNV) {
      t = s_block[threadIdx.x];
    }
    if (threadIdx.x < kNV) {
      // this rank's sums go straight into every rank's mailbox (P2P stores over NVLink; each 64-bit word is
      // self-validating, so no ordering between them is needed)
      const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(t));
      const size_t off = (((size_t)par * kMaxRanks + rc.rank) * kNVP + threadIdx.x) * 2;
      for (int r = 0; r < W; ++r) {
        st_volatile_u64(ws.mail[r] + off, pack_lo(tag, bits));
        st_volatile_u64(ws.mail[r] + off + 1, pack_hi(tag, bits));
      }
    }
  }
  __syncthreads();  // s_red8 (the rank-row buffer below) aliases the CTA-sum scratch read above
  {
    const int r = threadIdx.x >> 5, k = threadIdx.x & 31;
    if (r < W && k < kNV)
      s_red8[r][k] = poll_tagged(ws.mail[rc.rank] + (((size_t)par * kMaxRanks + r) * kNVP + k) * 2, tag, true, &rc.sync[3]);
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

// pose -> evaluation context in three pieces so that two warps share the work:
//   pose_trig   (warp 0, lanes 0..5)  one sincos per lane, a single code path: lanes 0..2 the fp64 angles with the
//               small-angle snap (ndt_omp_impl.hpp:292-326), lanes 3..5 the fp32-cast angles (Eigen::AngleAxis<float>)
//   pose_tables (warp 0)              the 69 table entries, two or three per lane
//   pose_matrix (warp 1, lane 0)      Translation * Rx * Ry * Rz in fp32 (same arithmetic as pose_to_matrix)
// s_trig: [0] = 1, [1..6] = cx sx cy sy cz sz (fp64, factor codes), [8..13] = the fp32 values of cx sx cy sy cz sz.
__device__ __forceinline__ void pose_trig(const double* x_t, double* s_trig) {
  const int lane = threadIdx.x & 31;
  if (lane < 6) {
    const int ax = lane < 3 ? lane : lane - 3;
    const double a64 = x_t[3 + ax];
    const double a = lane < 3 ? a64 : static_cast<double>(static_cast<float>(a64));
    double sv, cv;
    sincos(a, &sv, &cv);
    if (lane < 3) {
      if (fabs(a64) < 10e-5) { cv = 1.0; sv = 0.0; }
      s_trig[1 + 2 * ax] = cv;
      s_trig[2 + 2 * ax] = sv;
    } else {
      s_trig[8 + 2 * ax] = static_cast<double>(static_cast<float>(cv));
      s_trig[9 + 2 * ax] = static_cast<double>(static_cast<float>(sv));
    }
  } else if (lane == 12) {
    s_trig[0] = 1.0;
  }
}

__device__ __forceinline__ void pose_tables(EvalCtx& ctx, const double* s_trig, const TableTerm* s_terms, bool line_search_trial) {
  const int lane = threadIdx.x & 31;
  for (int e = lane; e < 69; e += 32) {
    const TableTerm t = s_terms[e];
    double v = 0.0;
    if (t.s1) v = __dmul_rn(__dmul_rn(s_trig[t.a1], s_trig[t.b1]), s_trig[t.c1]) * static_cast<double>(t.s1);
    if (t.s2) v = __dadd_rn(v, __dmul_rn(__dmul_rn(s_trig[t.a2], s_trig[t.b2]), s_trig[t.c2]) * static_cast<double>(t.s2));
    if (e < 24) {
      ctx.tab.jd[e / 3][e % 3] = v;
      reinterpret_cast<float*>(&ctx.tf[e / 3])[e % 3] = static_cast<float>(v);
    } else {
      const int r = (e - 24) / 3, c = (e - 24) % 3;
      ctx.tab.hd[r][c] = v;
      // Q2: the fp32 table carries +sy in row d1 (ndt_omp_impl.hpp:383), the fp64 table -sy (:361).  computeDerivatives
      // with compute_hessian = true (the first trial of a line search) reads the fp32 table: +sy.  A later trial runs
      // without a Hessian in the reference and is followed — if it was the last — by computeHessian, which reads the
      // fp64 table: its Hessian sums are formed during the trial itself, so those evaluations take -sy (= v).
      reinterpret_cast<float*>(&ctx.tf[8 + r])[c] =
          (r == 6 && c == 2 && !line_search_trial) ? static_cast<float>(s_trig[4]) : static_cast<float>(v);
    }
  }
}

__device__ __forceinline__ void pose_matrix(const double* x_t, EvalCtx& ctx, float* final_T, const double* s_trig) {
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

__device__ __forceinline__ void bar_sync_pair() {  // named barrier 1 shared by warps 0 and 1 (64 threads)
  asm volatile("bar.sync 1, 64;" ::: "memory");
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

// H (6x6, both triangles) from the 21 reduced upper-triangle totals, one or two entries per lane.
// computeDerivatives(..., false) leaves H zeroed (ndt_omp_impl.hpp:186, 218): the solver sees H = 0 after a line-search
// trial.  The trial's own Hessian sums are kept aside in H_ls: they were formed with computeHessian's angle table (the
// fp64 table's -sy in row d1, quirk Q2 — see pose_tables) at the pose computeHessian would be called with, so when the
// line search ends after extra trials (ndt_omp_impl.hpp:928-929) the "Hessian-only pass" is already done.
__device__ __forceinline__ void unpack_hessian_warp(const double* tot, int kind) {
  const int lane = threadIdx.x & 31;
  for (int idx = lane; idx < 36; idx += 32) {
    const int i = idx / 6, j = idx - i * 6;
    const int a = i < j ? i : j, b = i < j ? j : i;
    const int k = 7 + a * 6 - (a * (a - 1)) / 2 + (b - a);
    g_st.H[idx] = (kind == ACT_EVAL_NOHESS) ? 0.0 : tot[k];
    if (kind == ACT_EVAL_NOHESS) g_st.H_ls[idx] = tot[k];
  }
  __syncwarp();
}

// Fast path of the Newton solve: elimination WITHOUT pivot search / row swaps, rows spread over lanes 0..5.  Accepted
// only when every pivot has the same sign (H definite — the normal case near the optimum, where no-pivot elimination
// is backward stable) and the pivots span less than 1e9; otherwise the caller runs the pivoted solver.  Either way
// the solution equals JacobiSVD::solve (ndt_omp_impl.hpp:127-129) up to cond(H) * eps.
__device__ __forceinline__ bool warp_solve_definite() {
  SolverState& st = g_st;
  const int lane = threadIdx.x & 31;
  const int r = lane < 6 ? lane : 5;
  double a[7];
#pragma unroll
  for (int j = 0; j < 6; ++j) a[j] = st.H[r * 6 + j];
  a[6] = -st.g[r];
  double pmin = 1e300, pmax = 0.0;
  int npos = 0, nneg = 0;
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    const double pc = __shfl_sync(0xffffffffu, a[c], c);
    npos += (pc > 0.0);
    nneg += (pc < 0.0);
    pmin = fmin(pmin, fabs(pc));
    pmax = fmax(pmax, fabs(pc));
    const double f = a[c] * (1.0 / pc);
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      if (j <= c) continue;
      const double pj = __shfl_sync(0xffffffffu, a[j], c);
      if (lane > c && lane < 6) a[j] -= f * pj;
    }
  }
  const bool ok = (npos == 6 || nneg == 6) && (pmin > 1e-9 * pmax) && !isinf(pmax);
  if (!ok) return false;  // warp-uniform
  double dinv = 1.0;
#pragma unroll
  for (int j = 0; j < 6; ++j)
    if (j == r) dinv = 1.0 / a[j];
  double x[6];
#pragma unroll
  for (int rr = 5; rr >= 0; --rr) {
    const double xr = __shfl_sync(0xffffffffu, a[6] * dinv, rr);
    x[rr] = xr;
    if (lane < rr) a[6] -= a[rr] * xr;
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) st.delta[i] = x[i];
  }
  __syncwarp();
  return true;
}

// parity hook (ndtb200_debug_newton_solve): the two warp solvers on a caller-supplied system, one warp
__global__ void __launch_bounds__(32) newton_solve_debug_kernel(const double* __restrict__ H, const double* __restrict__ g,
                                                                double* __restrict__ delta, int* __restrict__ path) {
  for (int i = threadIdx.x; i < 36; i += 32) g_st.H[i] = H[i];
  if (threadIdx.x < 6) { g_st.g[threadIdx.x] = g[threadIdx.x]; g_st.delta[threadIdx.x] = 0.0; }
  __syncwarp();
  const bool fast = warp_solve_definite();
  if (!fast) warp_newton_solve(g_st);
  __syncwarp();
  if (threadIdx.x < 6) delta[threadIdx.x] = g_st.delta[threadIdx.x];
  if (threadIdx.x == 0) *path = fast ? 0 : 1;
}

// After the Newton solve: step-length bookkeeping and start of the line search
// (ndt_omp_impl.hpp:131-142, 772-837).  Returns the next action.
__device__ __noinline__ int newton_post() {
  SolverState& st = g_st;
  const SolverState& prm = g_st;
  while (true) {
    const double* delta = st.delta;  // a zero-length step leaves H and g unchanged: same solution next round
    double nrm = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) nrm += delta[i] * delta[i];
    nrm = sqrt(nrm);
    if (nrm == 0 || nrm != nrm) {  // ndt_omp_impl.hpp:134-139
      st.converged = (nrm == nrm) ? 1 : 0;
      return ACT_DONE;
    }
    const double inv_nrm = 1.0 / nrm;  // one division instead of six (differs from x / nrm by at most one fp64 ulp)
#pragma unroll
    for (int i = 0; i < 6; ++i) st.dir[i] = delta[i] * inv_nrm;
    // computeStepLengthMT(p, dir, nrm, step_size, eps/2, ...)
    st.step_max = prm.step_size;
    st.step_min = prm.trans_eps / 2;
    st.phi_0 = -st.score;
    double dphi = 0;
#pragma unroll
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
#pragma unroll
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
#pragma unroll
      for (int i = 0; i < 6; ++i) st.x_t[i] = st.p[i] + st.dir[i] * a_t;
      st.need_pose = 1;
      st.state = ST_MT_FIRST;
      return ACT_EVAL_FULL;
    }
    // zero-length step: finish this Newton iteration without an evaluation
#pragma unroll
    for (int i = 0; i < 6; ++i) st.last_dp[i] = st.dir[i] * a_ret;
#pragma unroll
    for (int i = 0; i < 6; ++i) st.p[i] += st.last_dp[i];
    if (st.nr_iterations > prm.max_iterations || (st.nr_iterations && (fabs(a_ret) < prm.trans_eps))) st.converged = 1;
    st.nr_iterations++;
    if (st.converged) return ACT_DONE;
  }
}

// Called after every evaluation with the reduced totals; returns the next action.
__device__ __noinline__ int advance(const double* tot, int kind, TraceRec* trace) {
  SolverState& st = g_st;
  const SolverState& prm = g_st;
  st.need_pose = 0;
  if (kind != ACT_HESS_ONLY) {
    st.score = tot[0];
#pragma unroll
    for (int i = 0; i < 6; ++i) st.g[i] = tot[1 + i];
  }
  // st.H was filled by the lanes of the warp (unpack_hessian_warp) before this call
  st.n_hits += static_cast<long long>(tot[28]);
  if (kind == ACT_HESS_ONLY) st.n_hess++; else st.n_evals++;
  if (trace && st.n_trace < prm.trace_cap) {
    TraceRec& r = trace[st.n_trace];
    r.kind = kind - 1;
    r.pad = 0;
#pragma unroll
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
#pragma unroll
      for (int i = 0; i < 6; ++i) d += st.g[i] * st.dir[i];
      st.d_phi_t = -d;
      st.psi_t = st.phi_t - st.phi_0 - mu * st.d_phi_0 * st.a_t;
      st.d_psi_t = st.d_phi_t - mu * st.d_phi_0;
      break;
    }
    case ST_MT_LOOP: {
      st.phi_t = -st.score;
      double d = 0;
#pragma unroll
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
#pragma unroll
    for (int i = 0; i < 6; ++i) st.x_t[i] = st.p[i] + st.dir[i] * a_t;
    st.need_pose = 1;
    st.state = ST_MT_LOOP;
    return ACT_EVAL_NOHESS;
  }
  if (st.step_iterations) {
    // ndt_omp_impl.hpp:928-929: computeHessian(hessian, trans_cloud, x_t) — same pose as the last trial, whose
    // evaluation already summed this Hessian (H_ls).  Recorded in the trace like the pass it replaces.
#pragma unroll
    for (int i = 0; i < 36; ++i) st.H[i] = st.H_ls[i];
    st.n_hess++;
    if (trace && st.n_trace < prm.trace_cap) {
      TraceRec& r = trace[st.n_trace];
      r.kind = 2;
      r.pad = 0;
#pragma unroll
      for (int i = 0; i < 6; ++i) r.x[i] = st.x_t[i];
      r.a_t = st.a_t;
      r.score = st.score;
      r.t_start = r.t_local = r.t_reduced = r.t_advanced = 0;
    }
    st.n_trace++;
  }
mt_finish: {
    // back in computeTransformation (ndt_omp_impl.hpp:143-164)
    const double a = st.a_t;
#pragma unroll
    for (int i = 0; i < 6; ++i) st.last_dp[i] = st.dir[i] * a;
#pragma unroll
    for (int i = 0; i < 6; ++i) st.p[i] = st.p[i] + st.last_dp[i];
    if (st.nr_iterations > prm.max_iterations || (st.nr_iterations && (fabs(a) < prm.trans_eps))) st.converged = 1;
    st.nr_iterations++;
    if (st.converged) return ACT_DONE;
    return ACT_SOLVE;
  }
}

// ---------------------------------------------------------------------------------------------
// fp32 derivative pass: ONE THREAD PER SOURCE POINT, contributions factorised per point.
//
// The reference adds, for every (point, voxel) hit, 6 gradient and 36 Hessian terms (ndt_omp_impl.hpp:512-534).  All of
// them are linear in three per-hit quantities once the point's J_E / H_E are factored out (they depend on the point
// only, :267):
//     A = sum_hits w u                (3)      u = C (x' - mean),  w = d1 d2 exp(-d2/2 (x'-mean)^T u)
//     M = sum_hits w (C - d2 u u^T)   (6, symmetric)
//     S = sum_hits -d1 e              (1)
//     g_i  = J_i . A                      H_ij = J_i^T M J_j + A . H_E[i][j]
// so a hit costs ~45 fp32 instructions (u, q, exp, 3 + 12 FMAs) instead of ~220, and the 21 Hessian / 6 gradient terms
// are formed once per point (~90 instructions).  J_E / H_E are only needed in that epilogue, which keeps the hit loop's
// register footprint small.  Same validity test per hit (e2 > 1 || e2 < 0 || NaN contributes nothing, :506-507).
// ---------------------------------------------------------------------------------------------
struct RecordRegs { float4 a, b, c; };  // the 48 hot bytes of a voxel record

__device__ __forceinline__ RecordRegs load_record(const VoxelRecord* R) {
  RecordRegs r;
  NDT_CHECK((reinterpret_cast<unsigned long long>(R) & 15ull) == 0ull);
  r.a = __ldg(reinterpret_cast<const float4*>(R));
  r.b = __ldg(reinterpret_cast<const float4*>(R) + 1);
  r.c = __ldg(reinterpret_cast<const float4*>(R) + 2);
  return r;
}

template <bool HESS>
__device__ __forceinline__ void hit_f32(const RecordRegs& rr, float tx, float ty, float tz, float d2f, float d1f,
                                        float& S, float (&A)[3], float (&M)[6]) {
  const float4 a = rr.a, b = rr.b, c = rr.c;
  // x_trans = fl32(double(x') - mean) (ndt_omp_impl.hpp:259-262, 492) via the exact hi/lo split of the mean
  const float r0 = __fsub_rn(__fsub_rn(tx, a.x), a.w);
  const float r1 = __fsub_rn(__fsub_rn(ty, a.y), b.x);
  const float r2 = __fsub_rn(__fsub_rn(tz, a.z), b.y);
  const float c00 = b.z, c01 = b.w, c02 = c.x, c11 = c.y, c12 = c.z, c22 = c.w;
  const float u0 = c00 * r0 + c01 * r1 + c02 * r2;
  const float u1 = c01 * r0 + c11 * r1 + c12 * r2;
  const float u2 = c02 * r0 + c12 * r1 + c22 * r2;
  const float q = r0 * u0 + r1 * u1 + r2 * u2;
  const float e = expf(-d2f * q * 0.5f);  // ndt_omp_impl.hpp:499
  const float e2 = d2f * e;
  if (!(e2 <= 1.0f && e2 >= 0.0f)) return;  // :506-507
  S -= d1f * e;
  const float w = e2 * d1f;
  A[0] += w * u0;
  A[1] += w * u1;
  A[2] += w * u2;
  if (HESS) {
    const float k = -d2f * w;
    const float t0 = k * u0, t1 = k * u1, t2 = k * u2;
    M[0] += w * c00; M[0] += t0 * u0;
    M[1] += w * c01; M[1] += t0 * u1;
    M[2] += w * c02; M[2] += t0 * u2;
    M[3] += w * c11; M[3] += t1 * u1;
    M[4] += w * c12; M[4] += t1 * u2;
    M[5] += w * c22; M[5] += t2 * u2;
  }
}

// acc: [0] score, [1..6] gradient, [7..27] Hessian upper triangle row-major, [28] hits found, [29..31] unused (zero)
template <int METHOD, bool HESS>
__device__ __forceinline__ void point_f32(float px, float py, float pz, const EvalCtx& ctx, const MapView& m, float d2f,
                                          float d1f, float (&acc)[32]) {
  float tx, ty, tz;
  transform_point(ctx.T, px, py, pz, tx, ty, tz);
  if (!point_is_finite(tx, ty, tz)) return;  // no neighbourhood, no contribution (see point_is_finite)
  // getNeighborhoodAtPoint (…_impl.hpp:379-381): cell = floor(x' / leaf), fp32 DIVISION (Q8)
  const int ix = lookup_cell(tx, m.leaf[0], m.inv_leaf[0]);
  const int iy = lookup_cell(ty, m.leaf[1], m.inv_leaf[1]);
  const int iz = lookup_cell(tz, m.leaf[2], m.inv_leaf[2]);
  constexpr int K = num_offsets<METHOD>();
  float S = 0.f, A[3] = {0.f, 0.f, 0.f}, M[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int nh = 0;
  if constexpr (METHOD == 2 || METHOD == 3) {
    int rec[K];
    probe_cells<K>(m, ix, iy, iz, rec);
    // (a depth-1 software pipeline of the record loads was measured: the extra live registers spill at the 64-register
    // cap and cost 12 % — the 32 resident warps already cover the L2 latency)
#ifdef NDTB200_ROLLED_HITS
    // one copy of the hit body (instruction-cache footprint): the neighbour's record index is selected from the registers
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
      int r = rec[0];
#pragma unroll
      for (int j = 1; j < K; ++j) r = (k == j) ? rec[j] : r;
      if (r < 0) continue;
      ++nh;
      hit_f32<HESS>(load_record(m.records + r), tx, ty, tz, d2f, d1f, S, A, M);
    }
#else
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (rec[k] < 0) continue;
      ++nh;
      NDT_CHECK(static_cast<uint32_t>(rec[k]) < m.n_records);
      hit_f32<HESS>(load_record(m.records + rec[k]), tx, ty, tz, d2f, d1f, S, A, M);
    }
#endif
  } else {
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
      const int rec = probe_neighbour<METHOD>(m, ix, iy, iz, k, tx, ty, tz);
      if (rec < 0) continue;
      NDT_CHECK(static_cast<uint32_t>(rec) < m.n_records);
      ++nh;
      hit_f32<HESS>(load_record(m.records + rec), tx, ty, tz, d2f, d1f, S, A, M);
    }
  }
  if (nh == 0) return;
  acc[28] += static_cast<float>(nh);
  acc[0] += S;
  // computePointDerivatives (fp32 overload, ndt_omp_impl.hpp:398-440): J_E entries pj[0..7], H_E entries ph[0..14]
  float pj[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const float4 t = ctx.tf[r];
    pj[r] = t.x * px + t.y * py + t.z * pz;
  }
  acc[1] += A[0];
  acc[2] += A[1];
  acc[3] += A[2];
  acc[4] += A[1] * pj[0] + A[2] * pj[1];
  acc[5] += A[0] * pj[2] + A[1] * pj[3] + A[2] * pj[4];
  acc[6] += A[0] * pj[5] + A[1] * pj[6] + A[2] * pj[7];
  if (HESS) {
    float ph[15];
#pragma unroll
    for (int r = 0; r < 15; ++r) {
      const float4 t = ctx.tf[8 + r];
      ph[r] = t.x * px + t.y * py + t.z * pz;
    }
    // M J_i for the rotational columns J_3 = (0,pj0,pj1), J_4 = (pj2,pj3,pj4), J_5 = (pj5,pj6,pj7)
    const float m3x = M[1] * pj[0] + M[2] * pj[1], m3y = M[3] * pj[0] + M[4] * pj[1], m3z = M[4] * pj[0] + M[5] * pj[1];
    const float m4x = M[0] * pj[2] + M[1] * pj[3] + M[2] * pj[4], m4y = M[1] * pj[2] + M[3] * pj[3] + M[4] * pj[4],
                m4z = M[2] * pj[2] + M[4] * pj[3] + M[5] * pj[4];
    const float m5x = M[0] * pj[5] + M[1] * pj[6] + M[2] * pj[7], m5y = M[1] * pj[5] + M[3] * pj[6] + M[4] * pj[7],
                m5z = M[2] * pj[5] + M[4] * pj[6] + M[5] * pj[7];
    acc[7] += M[0];  acc[8] += M[1];  acc[9] += M[2];  acc[10] += m3x; acc[11] += m4x; acc[12] += m5x;
    acc[13] += M[3]; acc[14] += M[4]; acc[15] += m3y;  acc[16] += m4y; acc[17] += m5y;
    acc[18] += M[5]; acc[19] += m3z;  acc[20] += m4z;  acc[21] += m5z;
    acc[22] += (pj[0] * m3y + pj[1] * m3z) + (A[1] * ph[0] + A[2] * ph[1]);
    acc[23] += (pj[0] * m4y + pj[1] * m4z) + (A[1] * ph[2] + A[2] * ph[3]);
    acc[24] += (pj[0] * m5y + pj[1] * m5z) + (A[1] * ph[4] + A[2] * ph[5]);
    acc[25] += (pj[2] * m4x + pj[3] * m4y + pj[4] * m4z) + (A[0] * ph[6] + A[1] * ph[7] + A[2] * ph[8]);
    acc[26] += (pj[2] * m5x + pj[3] * m5y + pj[4] * m5z) + (A[0] * ph[9] + A[1] * ph[10] + A[2] * ph[11]);
    acc[27] += (pj[5] * m5x + pj[6] * m5y + pj[7] * m5z) + (A[0] * ph[12] + A[1] * ph[13] + A[2] * ph[14]);
  }
}

// Sum 32 per-lane quantities over the 32 lanes of a warp with 31 shuffles (recursive halving: at every level a lane
// keeps one half of its values and sends the other half to its partner).  Lane l returns the warp sum of v[l].
// Fixed pairwise tree: deterministic, error <= 5 ulp of sum|v| — not a long fp32 run.
template <typename T>
__device__ __forceinline__ T warp_transpose_sum(T (&v)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const T send = upper ? v[i] : v[i + half];
      const T keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return v[0];
}

// The fp32 derivative pass of one warp over its 32-point groups g = warp_global + j * warps_total.  The first two
// points of every thread live in registers for the whole solve (the source never changes between evaluations; clouds
// of up to 2 x (threads in the grid) points are never re-read from memory after the first evaluation).
// Returns (lane k) the warp's fp64 sum of quantity k.
template <int METHOD, bool HESS>
__device__ __forceinline__ double eval_warp_f32(const float4* __restrict__ src, int n, int n_groups, int warp_global,
                                                int warps_total, const float (&cached)[2][3], const EvalCtx& ctx,
                                                const MapView& m, float d2f, float d1f) {
  const int lane = threadIdx.x & 31;
  float acc[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) acc[k] = 0.0f;
  double wsum = 0.0;
  int since = 0;
  int j = 0;
  for (int g = warp_global; g < n_groups; g += warps_total, ++j) {
    const int i = (g << 5) + lane;
    if (i < n) {
      NDT_CHECK(i >= 0);
      float px, py, pz;
#ifdef NDTB200_NO_CACHED_POINTS
      { const float4 pt = __ldg(src + i); px = pt.x; py = pt.y; pz = pt.z; }
#else
      if (j == 0) { px = cached[0][0]; py = cached[0][1]; pz = cached[0][2]; }
      else if (j == 1) { px = cached[1][0]; py = cached[1][1]; pz = cached[1][2]; }
      else { const float4 pt = __ldg(src + i); px = pt.x; py = pt.y; pz = pt.z; }
#endif
      point_f32<METHOD, HESS>(px, py, pz, ctx, m, d2f, d1f, acc);
    }
    if (++since == kFlushPoints) {  // bound the fp32 run length of a thread (large clouds only)
      wsum += static_cast<double>(warp_transpose_sum(acc));
#pragma unroll
      for (int k = 0; k < 32; ++k) acc[k] = 0.0f;
      since = 0;
    }
  }
  if (since) wsum += static_cast<double>(warp_transpose_sum(acc));
  return wsum;
}

// ---------------------------------------------------------------------------------------------
// the persistent kernel
// ---------------------------------------------------------------------------------------------
template <int METHOD, int THREADS>
__global__ void __launch_bounds__(THREADS, min_blocks_for(THREADS))
ndt_align_kernel(const float4* __restrict__ src_in, const MapView map, const AlignParams prm, const AlignWorkspace ws) {
  constexpr int kAlignWarps = THREADS / 32;
  SolverState& st = g_st;
  EvalCtx& ctx = g_ctx;
  __shared__ double s_warp[kAlignWarps][kNVP];
  __shared__ double s_block[kNVP];
  __shared__ double s_tot[kNVP];
  __shared__ unsigned long long s_dbg[4];
  __shared__ double s_trig[16];
  __shared__ TableTerm s_terms[69];
  __shared__ int s_action;
  __shared__ int s_flag;
  __shared__ int s_abort;
  __shared__ RankCtx s_rc;
  __shared__ MapView s_map;  // for the out-of-line fp64 pass (taking the parameter's address would copy it to local memory)

  const unsigned long long t_kernel_begin = globaltimer_ns();
  // the rank this CTA works for: the whole grid (one rank per GPU), or one of `vranks` emulated ranks sharing the launch
  const float4* __restrict__ src = src_in;
  int n = prm.n_source;
  unsigned int G = gridDim.x, bid = blockIdx.x;
  int my_rank = ws.rank;
  if (ws.vranks > 1) {  // uniform
    G = gridDim.x / static_cast<unsigned int>(ws.vranks);
    my_rank = static_cast<int>(blockIdx.x / G);
    bid = blockIdx.x - static_cast<unsigned int>(my_rank) * G;
    src = ws.vr[my_rank].src;
    n = ws.vr[my_rank].n_source;
  }
  if (threadIdx.x == 64 % THREADS) {
    RankCtx rc;
    if (ws.vranks > 1) {
      const VirtualRank v = ws.vr[my_rank];
      rc.partials = v.partials; rc.totals = v.totals; rc.sync = v.sync; rc.result = v.result;
    } else {
      rc.partials = ws.partials; rc.totals = ws.totals; rc.sync = ws.sync; rc.result = ws.result;
    }
    rc.G = G; rc.bid = bid; rc.rank = my_rank; rc.pad = 0;
    s_rc = rc;
    s_abort = 0;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // 32-point groups interleaved over all warps of the rank's CTAs (deterministic static balance)
  const int n_groups = (n + 31) >> 5;
  const int warp_global = warp * G + (bid + prm.rot) % G;  // consecutive groups go to different SMs
  const int warps_total = G * kAlignWarps;
  unsigned int epoch = 0;

  // this thread's first two source points stay in registers for the whole solve
  float cached[2][3];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const long long i = (static_cast<long long>(warp_global + j * warps_total) << 5) + lane;
    float4 pt = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) pt = __ldg(src + i);
    cached[j][0] = pt.x; cached[j][1] = pt.y; cached[j][2] = pt.z;
  }

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
    st.step_size = prm.step_size;
    st.trans_eps = prm.trans_eps;
    st.max_iterations = prm.max_iterations;
    st.trace_cap = prm.trace_cap;
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
  for (int i = threadIdx.x; i < 69; i += THREADS) s_terms[i] = g_table_terms[i];
  if (threadIdx.x == 33) s_map = map;
  if (threadIdx.x == 32) {  // tables for p0 (computed on the host); matrix = T0
    ctx.tab = prm.tab0;
#pragma unroll
    for (int r = 0; r < 8; ++r) ctx.tf[r] = make_float4(prm.tab0.jf[r][0], prm.tab0.jf[r][1], prm.tab0.jf[r][2], 0.f);
#pragma unroll
    for (int r = 0; r < 15; ++r) ctx.tf[8 + r] = make_float4(prm.tab0.hf[r][0], prm.tab0.hf[r][1], prm.tab0.hf[r][2], 0.f);
  }
  __syncthreads();
  const bool trace_cta = (blockIdx.x == 0);  // the CTA that records the optimiser trace / timeline

  const float d2f = static_cast<float>(prm.d2);
  const float d1f = static_cast<float>(prm.d1);
  while (true) {
    const int action = s_action;
    if (action == ACT_DONE) break;
    unsigned long long t_start = 0, t_local = 0, t_reduced = 0;
    const bool timing = (trace_cta && threadIdx.x == 0 && ws.trace != nullptr);
    if (timing) t_start = globaltimer_ns();

    if (action == ACT_HESS_ONLY) {
      eval_hessian_f64<METHOD>(src, n, n_groups, warp_global, warps_total, &ctx, &s_map, prm.d2, prm.d1, s_warp[warp]);
    } else {
      double wsum;
      // a line-search trial (ACT_EVAL_NOHESS) sums its Hessian as well — with computeHessian's table variant, see
      // pose_tables — so that no separate Hessian-only pass is needed when the line search ends on it.  A
      // derivatives-only evaluation requested through the parity API (MODE_EVAL, no pose update) skips the Hessian.
      if (action == ACT_EVAL_FULL || prm.mode == MODE_ALIGN)
        wsum = eval_warp_f32<METHOD, true>(src, n, n_groups, warp_global, warps_total, cached, ctx, map, d2f, d1f);
      else
        wsum = eval_warp_f32<METHOD, false>(src, n, n_groups, warp_global, warps_total, cached, ctx, map, d2f, d1f);
      s_warp[warp][lane] = wsum;
    }
    if (timing) s_dbg[0] = globaltimer_ns();
    __syncthreads();
    if (threadIdx.x < kNV) {  // fp64 sum over the CTA's warps, fixed order
      double s = 0;
#pragma unroll 8
      for (int w = 0; w < kAlignWarps; ++w) s += s_warp[w][threadIdx.x];
      s_block[threadIdx.x] = s;
    }
    if (timing) t_local = globaltimer_ns();
    grid_allreduce<kAlignWarps>(s_block, s_tot, ws, s_rc, epoch, &s_flag, s_warp, prm.launch_tag);
    if (ws.world > 1) {
      // a peer that never answered (bounded poll) or a launch that ran out of mailbox tags: every thread of the CTA must
      // take the same decision, so ONE thread reads the flag and the CTA breaks on the shared copy
      if (threadIdx.x == 0) {
        unsigned int ab = *reinterpret_cast<volatile unsigned int*>(&s_rc.sync[3]);
        if (epoch >= kMaxEpochsSharded) ab = 1u;
        s_abort = static_cast<int>(ab);
      }
      __syncthreads();
      if (s_abort) break;
    }
    if (timing) t_reduced = globaltimer_ns();
    unsigned long long t_last_arrive = 0;
    if (timing && G > 1)
      t_last_arrive = (ws.world == 1) ? static_cast<unsigned long long>(__double_as_longlong(s_tot[31]))
                                      : static_cast<unsigned long long>(__double_as_longlong(__ldcg(s_rc.totals + 2 * kNVP + ((epoch - 1u) & 1u))));
    int slot = 0;
    unsigned long long t_d0 = 0, t_d1 = 0, t_d2 = 0;
    // the step: warp 0 of every CTA runs the identical Newton / More-Thuente state machine on the identical totals
    if (warp == 0) {
      unpack_hessian_warp(s_tot, action);
      if (lane == 0) {
        slot = st.n_trace;
        s_action = advance(s_tot, action, trace_cta ? ws.trace : nullptr);
      }
      __syncwarp();
      if (timing) t_d0 = t_d1 = t_d2 = globaltimer_ns();
      if (s_action == ACT_SOLVE) {  // warp-uniform
        if (!warp_solve_definite()) warp_newton_solve(st);
        if (timing) t_d1 = globaltimer_ns();
        if (lane == 0) s_action = newton_post();
        __syncwarp();
        if (timing) t_d2 = globaltimer_ns();
      }
      if (st.need_pose) pose_trig(st.x_t, s_trig);
      bar_sync_pair();  // warp 1 builds the matrix while this warp fills the tables
      if (st.need_pose) pose_tables(ctx, s_trig, s_terms, s_action == ACT_EVAL_NOHESS);
      if (timing && slot < prm.trace_cap) {
        TraceRec& r = ws.trace[slot];
        r.t_start = t_start; r.t_local = t_local; r.t_reduced = t_reduced; r.t_advanced = globaltimer_ns();
        r.t_dbg[0] = t_d0; r.t_dbg[1] = t_d1; r.t_dbg[2] = t_d2; r.t_dbg[3] = t_last_arrive;
        r.t_phase[0] = s_dbg[0]; r.t_phase[1] = t_local;
      }
    } else if (warp == 1) {
      bar_sync_pair();
      if (st.need_pose && lane == 0) pose_matrix(st.x_t, ctx, st.final_T, s_trig);
    }
    __syncthreads();
  }

  // pcl::Registration::align's output (transformPointCloud(source, final_transformation_)): every CTA holds the identical
  // final_T in shared memory, so the cloud is written here instead of by a second kernel (same arithmetic as
  // transform_output_kernel: bit-identical)
  if (ws.out != nullptr) {  // uniform
    const float4* __restrict__ osrc = ws.out_src;
    for (int i = static_cast<int>(bid) * THREADS + threadIdx.x; i < n; i += static_cast<int>(G) * THREADS) {
      const float4 p = __ldg(osrc + i);
      float4 o;
      transform_point(st.final_T, p.x, p.y, p.z, o.x, o.y, o.z);
      o.w = 1.0f;
      ws.out[i] = o;
    }
  }
  if (bid == 0 && threadIdx.x == 0) {
    AlignResultDev& r = *s_rc.result;
    for (int i = 0; i < 12; ++i) r.final_T[i] = st.final_T[i];
    for (int i = 0; i < 6; ++i) r.last_dp[i] = st.last_dp[i];  // transformation_ = T(last_dp) is formed on the host
    r.converged = st.converged;
    r.iterations = st.nr_iterations;
    r.n_evals = st.n_evals;
    r.n_hess = st.n_hess;
    r.trans_probability = st.score / static_cast<double>(ws.world > 1 ? ws.n_source_total : (long long)n);  // ndt_omp_impl.hpp:136, 170
    r.aborted = (ws.world > 1) ? s_abort : 0;
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
  // Leave the barrier words zeroed for the next launch: the LAST CTA of the rank to get here resets them (every other
  // CTA has finished its last barrier by then); sync[2] is a wrapping exit counter that resets itself.
  if (threadIdx.x == 0) {
    if (G > 1) {
      const unsigned int old = atomicInc(&s_rc.sync[2], G - 1);
      if (old == G - 1) {
        s_rc.sync[0] = 0u;
        s_rc.sync[3] = 0u;
        __threadfence();
      }
    } else {
      s_rc.sync[3] = 0u;  // a one-CTA rank of a sharded solve: nobody else reads the flag of this launch any more
    }
  }
}

}  // namespace ndtb200
