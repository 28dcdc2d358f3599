// ndt_solve.cuh — on-device optimiser pieces: pose -> fp32 matrix, angle tables, 6x6 Newton solve,
// More-Thuente line search.  These run in ONE thread per CTA inside the persistent align kernel, so
// the Newton loop never returns to the host (ndt_omp_impl.hpp:80-171, 648-932).
#pragma once
#include "common.cuh"

namespace ndtb200 {

// ---------------------------------------------------------------------------------------------
// Translation * AngleAxis(rx,X) * AngleAxis(ry,Y) * AngleAxis(rz,Z) in fp32, evaluated as Eigen
// does (ndt_omp_impl.hpp:146-149, 827-830).  sin/cos of the fp32 angle are taken in fp64 and
// rounded to fp32 (correctly-rounded fp32 values on host libm and device alike).
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline void angle_axis_unit(float angle, int axis, float R[3][3]) {
  float ax[3] = {0.f, 0.f, 0.f};
  ax[axis] = 1.0f;
  const float s = static_cast<float>(sin(static_cast<double>(angle)));
  const float c = static_cast<float>(cos(static_cast<double>(angle)));
  const float sin_axis[3] = {s * ax[0], s * ax[1], s * ax[2]};
  const float c1 = 1.0f - c;
  const float cos1_axis[3] = {c1 * ax[0], c1 * ax[1], c1 * ax[2]};
  float tmp;
  tmp = cos1_axis[0] * ax[1];
  R[0][1] = tmp - sin_axis[2];
  R[1][0] = tmp + sin_axis[2];
  tmp = cos1_axis[0] * ax[2];
  R[0][2] = tmp + sin_axis[1];
  R[2][0] = tmp - sin_axis[1];
  tmp = cos1_axis[1] * ax[2];
  R[1][2] = tmp - sin_axis[0];
  R[2][1] = tmp + sin_axis[0];
#ifdef __CUDA_ARCH__
  for (int i = 0; i < 3; ++i) R[i][i] = __fadd_rn(__fmul_rn(cos1_axis[i], ax[i]), c);
#else
  for (int i = 0; i < 3; ++i) { volatile float t = cos1_axis[i] * ax[i]; R[i][i] = t + c; }
#endif
}

// Same rotation matrix from precomputed fp32 sine / cosine (the persistent kernel computes the
// trigonometry in parallel lanes).
__host__ __device__ inline void angle_axis_from_sc(float s, float c, int axis, float R[3][3]) {
  float ax[3] = {0.f, 0.f, 0.f};
  ax[axis] = 1.0f;
  const float sin_axis[3] = {s * ax[0], s * ax[1], s * ax[2]};
  const float c1 = 1.0f - c;
  const float cos1_axis[3] = {c1 * ax[0], c1 * ax[1], c1 * ax[2]};
  float tmp;
  tmp = cos1_axis[0] * ax[1];
  R[0][1] = tmp - sin_axis[2];
  R[1][0] = tmp + sin_axis[2];
  tmp = cos1_axis[0] * ax[2];
  R[0][2] = tmp + sin_axis[1];
  R[2][0] = tmp - sin_axis[1];
  tmp = cos1_axis[1] * ax[2];
  R[1][2] = tmp - sin_axis[0];
  R[2][1] = tmp + sin_axis[0];
#ifdef __CUDA_ARCH__
  for (int i = 0; i < 3; ++i) R[i][i] = __fadd_rn(__fmul_rn(cos1_axis[i], ax[i]), c);
#else
  for (int i = 0; i < 3; ++i) { volatile float t = cos1_axis[i] * ax[i]; R[i][i] = t + c; }
#endif
}

__host__ __device__ inline void mul33f(const float A[3][3], const float B[3][3], float C[3][3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
#ifdef __CUDA_ARCH__
      float s = __fmul_rn(A[i][0], B[0][j]);
      s = __fadd_rn(s, __fmul_rn(A[i][1], B[1][j]));
      s = __fadd_rn(s, __fmul_rn(A[i][2], B[2][j]));
#else
      volatile float t0 = A[i][0] * B[0][j], t1 = A[i][1] * B[1][j], t2 = A[i][2] * B[2][j];
      volatile float s01 = t0 + t1;
      float s = s01 + t2;
#endif
      C[i][j] = s;
    }
}

// T: row-major 3x4
__host__ __device__ inline void pose_to_matrix(const double p[6], float T[12]) {
  float Rx[3][3], Ry[3][3], Rz[3][3], Rxy[3][3], R[3][3];
  angle_axis_unit(static_cast<float>(p[3]), 0, Rx);
  angle_axis_unit(static_cast<float>(p[4]), 1, Ry);
  angle_axis_unit(static_cast<float>(p[5]), 2, Rz);
  mul33f(Rx, Ry, Rxy);
  mul33f(Rxy, Rz, R);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T[i * 4 + j] = R[i][j];
    T[i * 4 + 3] = static_cast<float>(p[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// computeAngleDerivatives (ndt_omp_impl.hpp:288-395).  jd/hd: fp64 tables (j_ang_a_.., h_ang_a2_..),
// jf/hf: the fp32 tables j_ang / h_ang.  Q2: fp32 row d1 carries +sy, fp64 row d1 carries -sy.
// ---------------------------------------------------------------------------------------------
struct AngleTables {
  double jd[8][3];
  double hd[15][3];
  float jf[8][3];
  float hf[15][3];
};

__host__ __device__ inline void compute_angle_tables(const double p[6], AngleTables& t) {
  double cx, cy, cz, sx, sy, sz;
  if (fabs(p[3]) < 10e-5) { cx = 1.0; sx = 0.0; } else { cx = cos(p[3]); sx = sin(p[3]); }
  if (fabs(p[4]) < 10e-5) { cy = 1.0; sy = 0.0; } else { cy = cos(p[4]); sy = sin(p[4]); }
  if (fabs(p[5]) < 10e-5) { cz = 1.0; sz = 0.0; } else { cz = cos(p[5]); sz = sin(p[5]); }
  const double J[8][3] = {
      {(-sx * sz + cx * sy * cz), (-sx * cz - cx * sy * sz), (-cx * cy)},
      {(cx * sz + sx * sy * cz), (cx * cz - sx * sy * sz), (-sx * cy)},
      {(-sy * cz), sy * sz, cy},
      {sx * cy * cz, (-sx * cy * sz), sx * sy},
      {(-cx * cy * cz), cx * cy * sz, (-cx * sy)},
      {(-cy * sz), (-cy * cz), 0},
      {(cx * cz - sx * sy * sz), (-cx * sz - sx * sy * cz), 0},
      {(sx * cz + cx * sy * sz), (cx * sy * cz - sx * sz), 0}};
  const double H[15][3] = {
      {(-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), sx * cy},
      {(-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), (-cx * cy)},
      {(cx * cy * cz), (-cx * cy * sz), (cx * sy)},
      {(sx * cy * cz), (-sx * cy * sz), (sx * sy)},
      {(-sx * cz - cx * sy * sz), (sx * sz - cx * sy * cz), 0},
      {(cx * cz - sx * sy * sz), (-sx * sy * cz - cx * sz), 0},
      {(-cy * cz), (cy * sz), (-sy)},
      {(-sx * sy * cz), (sx * sy * sz), (sx * cy)},
      {(cx * sy * cz), (-cx * sy * sz), (-cx * cy)},
      {(sy * sz), (sy * cz), 0},
      {(-sx * cy * sz), (-sx * cy * cz), 0},
      {(cx * cy * sz), (cx * cy * cz), 0},
      {(-cy * cz), (cy * sz), 0},
      {(-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), 0},
      {(-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), 0}};
  for (int r = 0; r < 8; ++r)
    for (int c = 0; c < 3; ++c) { t.jd[r][c] = J[r][c]; t.jf[r][c] = static_cast<float>(J[r][c]); }
  for (int r = 0; r < 15; ++r)
    for (int c = 0; c < 3; ++c) { t.hd[r][c] = H[r][c]; t.hf[r][c] = static_cast<float>(H[r][c]); }
  t.hf[6][2] = static_cast<float>(sy);  // ndt_omp_impl.hpp:383 (fp32 table) vs :361 (fp64 table)
}

// ---------------------------------------------------------------------------------------------
// Newton step: the reference solves H * delta = -g with JacobiSVD (ndt_omp_impl.hpp:127-129).
// Fast path: LU with partial pivoting (identical to the SVD solution up to cond(H)*eps when H has
// full rank).  If the pivots say H is (numerically) rank deficient, fall back to the one-sided
// Jacobi SVD pseudo-inverse with Eigen's default rank threshold, as the oracle does.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline void svd_solve6(const double* H /*row-major 36*/, const double* b, double* x) {
  double W[6][6], V[6][6];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) { W[i][j] = H[i * 6 + j]; V[i][j] = (i == j) ? 1.0 : 0.0; }
  const double eps = 2.220446049250313e-16;
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 5; ++p)
      for (int q = p + 1; q < 6; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int k = 0; k < 6; ++k) { alpha += W[k][p] * W[k][p]; beta += W[k][q] * W[k][q]; gamma += W[k][p] * W[k][q]; }
        if (gamma == 0.0 || fabs(gamma) <= eps * sqrt(alpha * beta)) continue;
        rotated = true;
        double zeta = (beta - alpha) / (2.0 * gamma);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int k = 0; k < 6; ++k) {
          double wp = W[k][p], wq = W[k][q];
          W[k][p] = c * wp - s * wq; W[k][q] = s * wp + c * wq;
          double vp = V[k][p], vq = V[k][q];
          V[k][p] = c * vp - s * vq; V[k][q] = s * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  double sig2[6], smax2 = 0;
  for (int j = 0; j < 6; ++j) {
    double s = 0;
    for (int k = 0; k < 6; ++k) s += W[k][j] * W[k][j];
    sig2[j] = s;
    smax2 = fmax(smax2, s);
  }
  const double thr = fmax(sqrt(smax2) * 6.0 * eps, 2.2250738585072014e-308);
  for (int i = 0; i < 6; ++i) x[i] = 0;
  for (int j = 0; j < 6; ++j) {
    if (!(sqrt(sig2[j]) > thr)) continue;
    double wb = 0;
    for (int k = 0; k < 6; ++k) wb += W[k][j] * b[k];
    const double coef = wb / sig2[j];
    for (int i = 0; i < 6; ++i) x[i] += V[i][j] * coef;
  }
}

__host__ __device__ inline void newton_solve6(const double* H, const double* b, double* x) {
  double A[6][7];
  double amax = 0;
  for (int i = 0; i < 6; ++i) {
    for (int j = 0; j < 6; ++j) { A[i][j] = H[i * 6 + j]; amax = fmax(amax, fabs(A[i][j])); }
    A[i][6] = b[i];
  }
  bool ok = (amax > 0) && (amax == amax) && !isinf(amax);
  double pmin = 1e300, pmax = 0;
  for (int c = 0; c < 6 && ok; ++c) {
    int piv = c;
    double best = fabs(A[c][c]);
    for (int r = c + 1; r < 6; ++r)
      if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); piv = r; }
    if (!(best > 0)) { ok = false; break; }
    pmin = fmin(pmin, best);
    pmax = fmax(pmax, best);
    if (piv != c)
      for (int k = c; k < 7; ++k) { double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
    const double inv = 1.0 / A[c][c];
    for (int r = c + 1; r < 6; ++r) {
      const double f = A[r][c] * inv;
      if (f != 0.0)
        for (int k = c; k < 7; ++k) A[r][k] -= f * A[c][k];
    }
  }
  if (ok && pmin > 1e-9 * pmax) {
    for (int r = 5; r >= 0; --r) {
      double s = A[r][6];
      for (int k = r + 1; k < 6; ++k) s -= A[r][k] * x[k];
      x[r] = s / A[r][r];
    }
    return;
  }
  svd_solve6(H, b, x);
}

// ---------------------------------------------------------------------------------------------
// More-Thuente pieces (ndt_omp_impl.hpp:648-769, ndt_omp.h:430-447)
// ---------------------------------------------------------------------------------------------
// std::min / std::max semantics (NaN handling differs from fmin / fmax)
__host__ __device__ inline double std_min(double a, double b) { return (b < a) ? b : a; }
__host__ __device__ inline double std_max(double a, double b) { return (a < b) ? b : a; }

__host__ __device__ inline bool mt_update_interval(double& a_l, double& f_l, double& g_l, double& a_u, double& f_u,
                                                   double& g_u, double a_t, double f_t, double g_t) {
  if (f_t > f_l) {
    a_u = a_t; f_u = f_t; g_u = g_t;
    return false;
  } else if (g_t * (a_l - a_t) > 0) {
    a_l = a_t; f_l = f_t; g_l = g_t;
    return false;
  } else if (g_t * (a_l - a_t) < 0) {
    a_u = a_l; f_u = f_l; g_u = g_l;
    a_l = a_t; f_l = f_t; g_l = g_t;
    return false;
  }
  return true;
}

__host__ __device__ inline double mt_trial_value(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u,
                                                 double a_t, double f_t, double g_t) {
  if (f_t > f_l) {
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
    if (fabs(a_c - a_l) < fabs(a_q - a_l)) return a_c;
    return 0.5 * (a_q + a_c);
  } else if (g_t * g_l < 0) {
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    if (fabs(a_c - a_t) >= fabs(a_s - a_t)) return a_c;
    return a_s;
  } else if (fabs(g_t) <= fabs(g_l)) {
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    double a_t_next = (fabs(a_c - a_t) < fabs(a_s - a_t)) ? a_c : a_s;
    if (a_t > a_l) return std_min(a_t + 0.66 * (a_u - a_t), a_t_next);
    return std_max(a_t + 0.66 * (a_u - a_t), a_t_next);
  }
  double z = 3 * (f_t - f_u) / (a_t - a_u) - g_t - g_u;
  double w = sqrt(z * z - g_t * g_u);
  return (a_u + (a_t - a_u) * (w - g_u - z) / (g_t - g_u + 2 * w));
}

}  // namespace ndtb200
