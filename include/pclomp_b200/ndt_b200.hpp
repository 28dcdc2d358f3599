// ndt_b200.hpp — header-only C++ shim: the public method set of pclomp::NormalDistributionsTransform
// (ndt_omp/include/pclomp/ndt_omp.h:96-238, 499) forwarded to the C ABI of libndt_b200.so (include/ndt_b200.h).
//
// A caller of the reference switches by changing
//     #include <pclomp/ndt_omp.h>                      ->  #include <pclomp_b200/ndt_b200.hpp>
//     pclomp::NormalDistributionsTransform<P, P>        ->  pclomp_b200::NormalDistributionsTransform<P, P>
// and linking -lndt_b200 instead of -lndt_omp.  Methods, argument meaning and observable behaviour
// (setInputTarget builds the map immediately, setResolution rebuilds only if the value changed and a source is set,
// align() may be repeated, final transform = pose of the last line-search trial, hasConverged() is also true at the
// iteration cap, copies are independent objects) follow the reference; errors are reported like PCL does
// (a message on stderr + empty map / converged == false), never by exception.
#pragma once
#include <cstdio>
#include <limits>
#include <memory>
#include <utility>
#include <vector>

#include "../ndt_b200.h"
#include "pcl_compat.hpp"

namespace pclomp_b200 {

enum NeighborSearchMethod { KDTREE = NDTB200_KDTREE, DIRECT26 = NDTB200_DIRECT26, DIRECT7 = NDTB200_DIRECT7, DIRECT1 = NDTB200_DIRECT1 };

template <typename PointSource, typename PointTarget = PointSource>
class NormalDistributionsTransform {
 public:
  typedef pcl::PointCloud<PointSource> PointCloudSource;
  typedef typename PointCloudSource::Ptr PointCloudSourcePtr;
  typedef typename PointCloudSource::ConstPtr PointCloudSourceConstPtr;
  typedef pcl::PointCloud<PointTarget> PointCloudTarget;
  typedef typename PointCloudTarget::ConstPtr PointCloudTargetConstPtr;
  typedef std::shared_ptr<NormalDistributionsTransform<PointSource, PointTarget>> Ptr;
  typedef std::shared_ptr<const NormalDistributionsTransform<PointSource, PointTarget>> ConstPtr;

  explicit NormalDistributionsTransform(int device = 0) : search_method(DIRECT7), h_(nullptr), device_(device) {
    const int st = ndtb200_create(&h_, device);
    if (st != NDTB200_OK) {
      std::fprintf(stderr, "[pclomp_b200::NormalDistributionsTransform] no usable CUDA device (status %d); "
                           "this library has no CPU fallback\n", st);
      h_ = nullptr;
    }
    ndtb200_default_params(&prm_);
  }
  ~NormalDistributionsTransform() { if (h_) ndtb200_destroy(h_); }

  // copy construction / assignment: the mapping node returns the object by value (ndt_omp_mapping_node.cpp:151-169)
  NormalDistributionsTransform(const NormalDistributionsTransform& o)
      : search_method(o.search_method), h_(nullptr), device_(o.device_), prm_(o.prm_), target_(o.target_), input_(o.input_) {
    if (o.h_) ndtb200_clone(o.h_, &h_);
  }
  NormalDistributionsTransform& operator=(const NormalDistributionsTransform& o) {
    if (this != &o) {
      if (h_) ndtb200_destroy(h_);
      h_ = nullptr;
      search_method = o.search_method; device_ = o.device_; prm_ = o.prm_; target_ = o.target_; input_ = o.input_;
      if (o.h_) ndtb200_clone(o.h_, &h_);
    }
    return *this;
  }
  NormalDistributionsTransform(NormalDistributionsTransform&& o) noexcept
      : search_method(o.search_method), h_(o.h_), device_(o.device_), prm_(o.prm_), target_(std::move(o.target_)), input_(std::move(o.input_)) {
    o.h_ = nullptr;
  }

  // ---- setters / getters (ndt_omp.h:115-209 + pcl::Registration) ----
  void setNumThreads(int) {}  // accepted, meaningless on the device path
  void setResolution(float resolution) { prm_.resolution = resolution; push(); }
  float getResolution() const { return prm_.resolution; }
  double getStepSize() const { return prm_.step_size; }
  void setStepSize(double step_size) { prm_.step_size = step_size; push(); }
  double getOutlierRatio() const { return prm_.outlier_ratio; }
  void setOutlierRatio(double outlier_ratio) { prm_.outlier_ratio = outlier_ratio; push(); }
  void setNeighborhoodSearchMethod(NeighborSearchMethod method) { search_method = method; }
  void setTransformationEpsilon(double epsilon) { prm_.trans_eps = epsilon; push(); }
  double getTransformationEpsilon() const { return prm_.trans_eps; }
  void setMaximumIterations(int nr_iterations) { prm_.max_iterations = nr_iterations; push(); }
  int getMaximumIterations() const { return prm_.max_iterations; }

  void setInputTarget(const PointCloudTargetConstPtr& cloud) {
    target_ = cloud;
    if (!h_) return;
    const void* p = (cloud && !cloud->points.empty()) ? static_cast<const void*>(cloud->points.data()) : nullptr;
    const size_t n = cloud ? cloud->points.size() : 0;
    const int st = ndtb200_set_target(h_, p, n, sizeof(PointTarget), cloud ? (cloud->is_dense ? 1 : 0) : 1);
    if (st == NDTB200_ERR_NO_INPUT)
      std::fprintf(stderr, "[pclomp_b200::VoxelGridCovariance::applyFilter] No input dataset given!\n");
    else if (st == NDTB200_ERR_GRID_OVERFLOW)
      std::fprintf(stderr, "[pclomp_b200::VoxelGridCovariance::applyFilter] Leaf size is too small for the input dataset. "
                           "Integer indices would overflow.\n");
    else if (st != NDTB200_OK)
      std::fprintf(stderr, "[pclomp_b200] setInputTarget failed: %s\n", ndtb200_last_error(h_));
  }
  void setInputSource(const PointCloudSourceConstPtr& cloud) {
    input_ = cloud;
    if (!h_) return;
    const void* p = (cloud && !cloud->points.empty()) ? static_cast<const void*>(cloud->points.data()) : nullptr;
    const int st = ndtb200_set_source(h_, p, cloud ? cloud->points.size() : 0, sizeof(PointSource));
    if (st != NDTB200_OK) std::fprintf(stderr, "[pclomp_b200] setInputSource failed: %s\n", ndtb200_last_error(h_));
  }
  PointCloudTargetConstPtr getInputTarget() const { return target_; }
  PointCloudSourceConstPtr getInputSource() const { return input_; }

  // ---- registration (pcl::Registration::align -> computeTransformation, ndt_omp_impl.hpp:80-171) ----
  void align(PointCloudSource& output) { align(output, Eigen::Matrix4f::Identity()); }
  void align(PointCloudSource& output, const Eigen::Matrix4f& guess) {
    if (!h_ || !input_) return;
    prm_.search_method = static_cast<int32_t>(search_method);  // public field, read at align time like the reference
    ndtb200_set_params(h_, &prm_);
    output.points.resize(input_->points.size());
    output.width = static_cast<uint32_t>(output.points.size());
    output.height = 1;
    output.is_dense = input_->is_dense;
    for (size_t i = 0; i < output.points.size(); ++i) output.points[i] = input_->points[i];  // copies the non-xyz fields
    const int st = ndtb200_align(h_, guess.data(), output.points.empty() ? nullptr : output.points.data(), sizeof(PointSource));
    if (st != NDTB200_OK) std::fprintf(stderr, "[pclomp_b200] align failed: %s\n", ndtb200_last_error(h_));
  }

  Eigen::Matrix4f getFinalTransformation() { return matrix_of(true); }
  Eigen::Matrix4f getLastIncrementalTransformation() { return matrix_of(false); }
  bool hasConverged() { ndtb200_result r; return h_ && ndtb200_get_result(h_, &r) == NDTB200_OK && r.converged != 0; }
  int getFinalNumIteration() { ndtb200_result r; return (h_ && ndtb200_get_result(h_, &r) == NDTB200_OK) ? r.iterations : 0; }
  double getTransformationProbability() {
    ndtb200_result r;
    return (h_ && ndtb200_get_result(h_, &r) == NDTB200_OK) ? r.trans_probability : 0.0;
  }
  double getFitnessScore(double max_range = std::numeric_limits<double>::max()) {
    double v = std::numeric_limits<double>::max();
    if (h_) ndtb200_fitness_score(h_, max_range, &v);
    return v;
  }
  // negative log likelihood of an already transformed cloud (ndt_omp_impl.hpp:935-983)
  double calculateScore(const PointCloudSource& cloud) {
    double v = 0.0;
    if (h_ && !cloud.points.empty()) {
      prm_.search_method = static_cast<int32_t>(search_method);
      ndtb200_set_params(h_, &prm_);
      ndtb200_calculate_score(h_, cloud.points.data(), cloud.points.size(), sizeof(PointSource), &v);
    }
    return v;
  }

  ndtb200_handle* handle() { return h_; }

  // ---- not in the reference: independent scan pairs aligned together (ndtb200_align_batch) ----
  // Every object already holds its own target and source; outputs[i] receives the aligned cloud of objects[i] like
  // align(output, guess) does.  guesses may be empty (identity) or hold one matrix per object.
  static void alignBatch(const std::vector<NormalDistributionsTransform*>& objects, const std::vector<PointCloudSource*>& outputs,
                         const std::vector<Eigen::Matrix4f>& guesses = std::vector<Eigen::Matrix4f>()) {
    const int n = static_cast<int>(objects.size());
    if (n == 0) return;
    std::vector<ndtb200_handle*> hs(n, nullptr);
    std::vector<void*> outs(n, nullptr);
    std::vector<float> g;
    if (!guesses.empty()) g.resize(16 * static_cast<size_t>(n));
    for (int i = 0; i < n; ++i) {
      NormalDistributionsTransform* o = objects[i];
      if (!o || !o->h_ || !o->input_) { std::fprintf(stderr, "[pclomp_b200] alignBatch: object %d has no device handle / source\n", i); return; }
      o->prm_.search_method = static_cast<int32_t>(o->search_method);
      ndtb200_set_params(o->h_, &o->prm_);
      hs[i] = o->h_;
      if (i < static_cast<int>(outputs.size()) && outputs[i]) {
        PointCloudSource& out = *outputs[i];
        out.points.resize(o->input_->points.size());
        out.width = static_cast<uint32_t>(out.points.size());
        out.height = 1;
        out.is_dense = o->input_->is_dense;
        for (size_t k = 0; k < out.points.size(); ++k) out.points[k] = o->input_->points[k];
        outs[i] = out.points.empty() ? nullptr : static_cast<void*>(out.points.data());
      }
      if (!guesses.empty()) {
        const Eigen::Matrix4f& G = guesses[static_cast<size_t>(i) < guesses.size() ? i : guesses.size() - 1];
        for (int k = 0; k < 16; ++k) g[16 * static_cast<size_t>(i) + k] = G.data()[k];
      }
    }
    const int st = ndtb200_align_batch(hs.data(), n, guesses.empty() ? nullptr : g.data(), outs.data(), sizeof(PointSource), nullptr);
    if (st != NDTB200_OK) std::fprintf(stderr, "[pclomp_b200] alignBatch failed (status %d): %s\n", st, ndtb200_last_error(hs[0]));
  }

  NeighborSearchMethod search_method;  // public in the reference too (ndt_omp.h:499)

 private:
  void push() {
    if (!h_) return;
    prm_.search_method = static_cast<int32_t>(search_method);
    const int st = ndtb200_set_params(h_, &prm_);
    if (st == NDTB200_ERR_CUDA) std::fprintf(stderr, "[pclomp_b200] set_params failed: %s\n", ndtb200_last_error(h_));
  }
  Eigen::Matrix4f matrix_of(bool final_not_increment) {
    Eigen::Matrix4f M = Eigen::Matrix4f::Identity();
    ndtb200_result r;
    if (h_ && ndtb200_get_result(h_, &r) == NDTB200_OK) {
      const float* src = final_not_increment ? r.final_transformation : r.last_increment;
      for (int c = 0; c < 4; ++c)
        for (int rr = 0; rr < 4; ++rr) M(rr, c) = src[c * 4 + rr];
    }
    return M;
  }

  ndtb200_handle* h_;
  int device_;
  ndtb200_params prm_;
  PointCloudTargetConstPtr target_;
  PointCloudSourceConstPtr input_;
};

// pcl::VoxelGrid<PointT> as the callers use it before the registration (ndt_omp/apps/align.cpp:57-69,
// ndt_rosbag_mapping_node.cpp:108-118): setLeafSize / setInputCloud / filter, centroids computed on the device
// (ndtb200_voxelgrid_filter) with pcl::VoxelGrid's own arithmetic.  Only x, y, z are produced (the reference's callers
// run it on PointXYZ / use xyz only); the other fields of the output points are value-initialised.
template <typename PointT>
class VoxelGrid {
 public:
  typedef pcl::PointCloud<PointT> PointCloud;
  typedef typename PointCloud::ConstPtr PointCloudConstPtr;
  explicit VoxelGrid(int device = 0) : h_(nullptr), leaf_(0.f) { if (ndtb200_create(&h_, device) != NDTB200_OK) h_ = nullptr; }
  ~VoxelGrid() { if (h_) ndtb200_destroy(h_); }
  VoxelGrid(const VoxelGrid&) = delete;
  VoxelGrid& operator=(const VoxelGrid&) = delete;
  void setLeafSize(float lx, float, float) { leaf_ = lx; }  // the callers use cubic leaves
  void setInputCloud(const PointCloudConstPtr& cloud) { input_ = cloud; }
  void filter(PointCloud& output) {
    output.points.clear();
    if (!h_ || !input_ || input_->points.empty()) { output.width = 0; output.height = 1; return; }
    struct Rec { float x, y, z, w; };
    std::vector<Rec> out(input_->points.size());
    int64_t m = 0;
    const int st = ndtb200_voxelgrid_filter(h_, input_->points.data(), input_->points.size(), sizeof(PointT), leaf_, out.data(),
                                            out.size(), sizeof(Rec), &m);
    if (st == NDTB200_ERR_GRID_OVERFLOW) {  // pcl::VoxelGrid: warn and pass the cloud through
      std::fprintf(stderr, "[pclomp_b200::VoxelGrid::applyFilter] Leaf size is too small for the input dataset. Integer indices would overflow.\n");
      output = *input_;
      return;
    }
    if (st != NDTB200_OK) { std::fprintf(stderr, "[pclomp_b200::VoxelGrid] filter failed: %s\n", ndtb200_last_error(h_)); return; }
    output.points.resize(static_cast<size_t>(m));
    for (int64_t i = 0; i < m; ++i) {
      PointT p = PointT();
      p.x = out[i].x; p.y = out[i].y; p.z = out[i].z;
      output.points[static_cast<size_t>(i)] = p;
    }
    output.width = static_cast<uint32_t>(m);
    output.height = 1;
    output.is_dense = true;
  }

 private:
  ndtb200_handle* h_;
  float leaf_;
  PointCloudConstPtr input_;
};

// The mapping-node loop (lidar_subscriber/src/ndt_rosbag_mapping_node.cpp:42-161) on the device-resident pipeline
// (ndtb200_mapper_*): one pushScan() per raw scan replaces downsample_cloud + perform_registration + pose chaining +
// update_global_map; globalMap() is what publish_global_map() sends.
template <typename PointT>
class Mapper {
 public:
  struct Step {
    Eigen::Matrix4f transform, pose;
    double fitness;
    bool converged;
    int iterations;
    size_t n_filtered, n_map;
  };
  explicit Mapper(float voxel_leaf = 0.3f, float map_voxel = 0.5f, int device = 0, bool compute_fitness = true) : m_(nullptr) {
    if (ndtb200_mapper_create(&m_, device, nullptr, voxel_leaf, map_voxel, compute_fitness ? 1 : 0) != NDTB200_OK) m_ = nullptr;
  }
  ~Mapper() { if (m_) ndtb200_mapper_destroy(m_); }
  Mapper(const Mapper&) = delete;
  Mapper& operator=(const Mapper&) = delete;
  bool ok() const { return m_ != nullptr; }
  Step pushScan(const pcl::PointCloud<PointT>& cloud) {
    Step out;
    out.transform = Eigen::Matrix4f::Identity();
    out.pose = Eigen::Matrix4f::Identity();
    out.fitness = 0; out.converged = false; out.iterations = 0; out.n_filtered = 0; out.n_map = 0;
    if (!m_) return out;
    ndtb200_mapper_step s;
    const int st = ndtb200_mapper_push_scan(m_, cloud.points.empty() ? nullptr : cloud.points.data(), cloud.points.size(), sizeof(PointT), &s);
    if (st != NDTB200_OK) { std::fprintf(stderr, "[pclomp_b200::Mapper] push_scan failed: %s\n", ndtb200_mapper_last_error(m_)); return out; }
    for (int c = 0; c < 4; ++c)
      for (int r = 0; r < 4; ++r) { out.transform(r, c) = s.transform[c * 4 + r]; out.pose(r, c) = s.pose[c * 4 + r]; }
    out.fitness = s.fitness; out.converged = s.converged != 0; out.iterations = s.iterations;
    out.n_filtered = static_cast<size_t>(s.n_filtered); out.n_map = static_cast<size_t>(s.n_map);
    return out;
  }
  void globalMap(pcl::PointCloud<PointT>& out) {
    out.points.clear();
    int64_t n = 0;
    if (!m_ || ndtb200_mapper_get_map(m_, nullptr, 0, sizeof(PointT), &n) != NDTB200_OK || n == 0) { out.width = 0; out.height = 1; return; }
    struct Rec { float x, y, z, w; };
    std::vector<Rec> tmp(static_cast<size_t>(n));
    ndtb200_mapper_get_map(m_, tmp.data(), tmp.size(), sizeof(Rec), &n);
    out.points.resize(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) { PointT p = PointT(); p.x = tmp[i].x; p.y = tmp[i].y; p.z = tmp[i].z; out.points[static_cast<size_t>(i)] = p; }
    out.width = static_cast<uint32_t>(n); out.height = 1; out.is_dense = true;
  }

 private:
  ndtb200_mapper* m_;
};

}  // namespace pclomp_b200
