// oracle/ndt_oracle_capi.cpp — C API over the CPU oracle for ctypes (TEST INFRASTRUCTURE ONLY).
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this library.
#include "ndt_oracle.hpp"

using namespace ndt_oracle;

namespace {
inline Mat4f from_colmajor(const float* m) {
  Mat4f T;
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) T.m[r][c] = m[c * 4 + r];
  return T;
}
inline void to_colmajor(const Mat4f& T, float* m) {
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) m[c * 4 + r] = T.m[r][c];
}
inline const P4* as_p4(const float* xyzw) { return reinterpret_cast<const P4*>(xyzw); }
}  // namespace

extern "C" {

void* ndto_create() { return new NormalDistributionsTransform(); }
void ndto_destroy(void* h) { delete static_cast<NormalDistributionsTransform*>(h); }

int ndto_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void ndto_set_params(void* h, float resolution, double step_size, double outlier_ratio, double trans_eps,
                     int max_iterations, int search_method, int num_threads, int min_points_per_voxel,
                     double eig_ratio) {
  auto* o = static_cast<NormalDistributionsTransform*>(h);
  o->step_size_ = step_size;
  o->outlier_ratio_ = outlier_ratio;
  o->transformation_epsilon_ = trans_eps;
  o->max_iterations_ = max_iterations;
  o->search_method = search_method;
  if (num_threads > 0) o->num_threads_ = num_threads;
  o->target_cells_.min_points_per_voxel = min_points_per_voxel;
  o->target_cells_.min_covar_eigvalue_mult = eig_ratio;
  o->resolution_ = resolution;  // plain store: callers set params before the target (no rebuild rule here)
  o->computeGaussConstants();
}

int ndto_set_target(void* h, const float* xyzw, size_t n, int is_dense) {
  auto* o = static_cast<NormalDistributionsTransform*>(h);
  o->setInputTarget(as_p4(xyzw), n, is_dense != 0);
  return static_cast<int>(o->build_status_);
}

void ndto_set_source(void* h, const float* xyzw, size_t n) {
  static_cast<NormalDistributionsTransform*>(h)->setInputSource(as_p4(xyzw), n);
}

void ndto_align(void* h, const float* guess_colmajor, float* out_xyzw) {
  auto* o = static_cast<NormalDistributionsTransform*>(h);
  Mat4f guess = guess_colmajor ? from_colmajor(guess_colmajor) : identity4f();
  std::vector<P4> out;
  o->align(out, guess);
  if (out_xyzw) std::memcpy(out_xyzw, out.data(), out.size() * sizeof(P4));
}

void ndto_get_result(void* h, float* final_colmajor, int* converged, int* iterations, double* trans_probability,
                     int* n_evaluations, int* n_hessian_passes) {
  auto* o = static_cast<NormalDistributionsTransform*>(h);
  to_colmajor(o->final_transformation_, final_colmajor);
  *converged = o->converged_ ? 1 : 0;
  *iterations = o->nr_iterations_;
  *trans_probability = o->trans_probability_;
  int ne = 0, nh = 0;
  for (const auto& r : o->trace) (r.kind == 2 ? nh : ne)++;
  *n_evaluations = ne;
  *n_hessian_passes = nh;
}

double ndto_fitness(void* h, double max_range) {
  return static_cast<NormalDistributionsTransform*>(h)->getFitnessScore(max_range);
}

double ndto_calculate_score(void* h, const float* xyzw, size_t n) {
  auto* o = static_cast<NormalDistributionsTransform*>(h);
  std::vector<P4> cloud(as_p4(xyzw), as_p4(xyzw) + n);
  return o->calculateScore(cloud);
}

void ndto_gauss(void* h, double* d3) {
  auto* o = static_cast<NormalDistributionsTransform*>(h);
  o->computeGaussConstants();
  d3[0] = o->gauss_d1_; d3[1] = o->gauss_d2_; d3[2] = o->gauss_d3_;
}

void ndto_map_info(void* h, int* min_b, int* max_b, int* div_b, long* n_leaves, long* n_valid) {
  auto* o = static_cast<NormalDistributionsTransform*>(h);
  const auto& g = o->target_cells_;
  for (int a = 0; a < 3; ++a) { min_b[a] = g.min_b[a]; max_b[a] = g.max_b[a]; div_b[a] = g.div_b[a]; }
  *n_leaves = static_cast<long>(g.leaves.size());
  long v = 0;
  for (const auto& kv : g.leaves) v += (kv.second.nr_points >= g.min_points_per_voxel);
  *n_valid = v;
}

void ndto_point_keys(void* h, int* keys) {
  auto* o = static_cast<NormalDistributionsTransform*>(h);
  std::memcpy(keys, o->target_cells_.point_keys.data(), o->target_cells_.point_keys.size() * sizeof(int));
}

// Leaves in ascending key order.  counts: nr_points (-1 for rejected leaves).
long ndto_dump_leaves(void* h, int* keys, int* counts, double* mean, double* cov, double* icov, int* inflated) {
  auto* o = static_cast<NormalDistributionsTransform*>(h);
  long i = 0;
  for (const auto& kv : o->target_cells_.leaves) {
    const Leaf& l = kv.second;
    keys[i] = static_cast<int>(kv.first);
    counts[i] = l.nr_points;
    for (int a = 0; a < 3; ++a) mean[i * 3 + a] = l.mean[a];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        cov[i * 9 + a * 3 + b] = l.cov.m[a][b];
        icov[i * 9 + a * 3 + b] = l.icov.m[a][b];
      }
    inflated[i] = l.inflated ? 1 : 0;
    ++i;
  }
  return i;
}

// out43 = score, g[6], H[36] (row-major).  T_colmajor may be NULL (=> matrix built from p).
long ndto_eval_derivatives(void* h, const double* p, const float* T_colmajor, int compute_hessian, double* out43) {
  auto* o = static_cast<NormalDistributionsTransform*>(h);
  o->computeGaussConstants();
  Mat4f T = T_colmajor ? from_colmajor(T_colmajor) : pose_to_matrix(p);
  std::vector<P4> trans;
  o->transformCloud(o->input_, trans, T);
  double g[6], H[6][6];
  out43[0] = o->computeDerivatives(g, H, trans, p, compute_hessian != 0);
  for (int i = 0; i < 6; ++i) out43[1 + i] = g[i];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) out43[7 + i * 6 + j] = H[i][j];
  return o->n_hits_last;
}

void ndto_eval_hessian(void* h, const double* p, const float* T_colmajor, double* out36) {
  auto* o = static_cast<NormalDistributionsTransform*>(h);
  o->computeGaussConstants();
  Mat4f T = T_colmajor ? from_colmajor(T_colmajor) : pose_to_matrix(p);
  std::vector<P4> trans;
  o->transformCloud(o->input_, trans, T);
  o->computeAngleDerivatives(p);
  double H[6][6];
  o->computeHessian(H, trans);
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) out36[i * 6 + j] = H[i][j];
}

// Neighbour keys of each query point: out_keys[n][26], -1 padded.  Returns total hits.
long ndto_lookup(void* h, const float* xyzw, size_t n, int search_method, int* out_keys) {
  auto* o = static_cast<NormalDistributionsTransform*>(h);
  int rel[26][3];
  int K = neighbor_offsets(search_method, rel);
  long total = 0;
  std::vector<const Leaf*> nb;
  std::vector<int> keys;
  for (size_t i = 0; i < n; ++i) {
    o->target_cells_.getNeighborhoodAtPoint(rel, K, as_p4(xyzw)[i], nb, &keys);
    for (int k = 0; k < 26; ++k) out_keys[i * 26 + k] = (k < static_cast<int>(keys.size())) ? keys[k] : -1;
    total += static_cast<long>(keys.size());
  }
  return total;
}

long ndto_trace(void* h, int* kinds, double* x6, double* a_t, double* score, long cap) {
  auto* o = static_cast<NormalDistributionsTransform*>(h);
  long n = std::min<long>(cap, static_cast<long>(o->trace.size()));
  for (long i = 0; i < n; ++i) {
    kinds[i] = o->trace[i].kind;
    for (int k = 0; k < 6; ++k) x6[i * 6 + k] = o->trace[i].x[k];
    a_t[i] = o->trace[i].a_t;
    score[i] = o->trace[i].score;
  }
  return static_cast<long>(o->trace.size());
}

long ndto_voxelgrid(const float* xyzw, size_t n, float leaf, float* out_xyzw, size_t cap) {
  std::vector<P4> out;
  long r = voxelgrid_downsample(as_p4(xyzw), n, leaf, out);
  if (r < 0) return r;
  size_t m = std::min(cap, out.size());
  if (out_xyzw) std::memcpy(out_xyzw, out.data(), m * sizeof(P4));
  return r;
}

long ndto_voxelgrid3(const float* xyzw, size_t n, const float* leaf3, float* out_xyzw, size_t cap) {
  std::vector<P4> out;
  long r = voxelgrid_downsample(as_p4(xyzw), n, leaf3, out);
  if (r < 0) return r;
  size_t m = std::min(cap, out.size());
  if (out_xyzw) std::memcpy(out_xyzw, out.data(), m * sizeof(P4));
  return r;
}

void ndto_pose_to_matrix(const double* p, float* T_colmajor) { to_colmajor(pose_to_matrix(p), T_colmajor); }

void ndto_matrix_to_pose(const float* T_colmajor, double* p) {
  Mat4f T = from_colmajor(T_colmajor);
  float R[3][3], ang[3];
  rotation_polar(T, R);
  euler_angles_012(R, ang);
  p[0] = T.m[0][3]; p[1] = T.m[1][3]; p[2] = T.m[2][3];
  p[3] = ang[0]; p[4] = ang[1]; p[5] = ang[2];
}

void ndto_transform(const float* T_colmajor, const float* xyzw, size_t n, float* out_xyzw) {
  Mat4f T = from_colmajor(T_colmajor);
  for (size_t i = 0; i < n; ++i) reinterpret_cast<P4*>(out_xyzw)[i] = transform_point(T, as_p4(xyzw)[i]);
}

void ndto_svd_solve6(const double* H36, const double* b6, double* x6) {
  double H[6][6];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) H[i][j] = H36[i * 6 + j];
  svd_solve6(H, b6, x6);
}

}  // extern "C"
