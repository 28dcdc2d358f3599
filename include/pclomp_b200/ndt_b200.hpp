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
#if __has_include(<omp.h>)
#include <omp.h>  // callers of the reference get omp_get_max_threads() through its headers (ndt_omp/apps/align.cpp:88)
#endif

namespace pclomp_b200 {

namespace detail {
template <typename P, typename T> struct rebind_ptr;
template <template <typename...> class SP, typename U, typename T> struct rebind_ptr<SP<U>, T> { typedef SP<T> type; };
}  // namespace detail

enum NeighborSearchMethod { KDTREE = NDTB200_KDTREE, DIRECT26 = NDTB200_DIRECT26, DIRECT7 = NDTB200_DIRECT7, DIRECT1 = NDTB200_DIRECT1 };

template <typename PointSource, typename PointTarget = PointSource>
class NormalDistributionsTransform : public pcl::Registration<PointSource, PointTarget> {
  // Derived from pcl::Registration exactly like the reference class (ndt_omp.h:70-71): the real one when PCL is
  // installed, the stand-in of pcl_compat.hpp otherwise.  A `pcl::Registration<P, P>::Ptr` can therefore hold this
  // object (ndt_omp/apps/align.cpp:15, 95-103) and the base class's align() wrapper calls computeTransformation() below.
  typedef pcl::Registration<PointSource, PointTarget> Base;

 protected:
  using Base::input_;
  using Base::target_;
  using Base::final_transformation_;
  using Base::transformation_;
  using Base::previous_transformation_;
  using Base::converged_;
  using Base::nr_iterations_;
  using Base::max_iterations_;
  using Base::transformation_epsilon_;
  using Base::reg_name_;

 public:
  typedef pcl::PointCloud<PointSource> PointCloudSource;
  typedef typename PointCloudSource::Ptr PointCloudSourcePtr;
  typedef typename PointCloudSource::ConstPtr PointCloudSourceConstPtr;
  typedef pcl::PointCloud<PointTarget> PointCloudTarget;
  typedef typename PointCloudTarget::ConstPtr PointCloudTargetConstPtr;
  // the smart pointer template of the base class's Ptr (boost::shared_ptr up to PCL 1.10, std::shared_ptr from 1.11 and in
  // the stand-in), so that a Ptr of this class converts to pcl::Registration<...>::Ptr as the reference's does
  typedef typename detail::rebind_ptr<typename Base::Ptr, NormalDistributionsTransform<PointSource, PointTarget>>::type Ptr;
  typedef typename detail::rebind_ptr<typename Base::Ptr, const NormalDistributionsTransform<PointSource, PointTarget>>::type ConstPtr;
  static_assert(sizeof(PointSource) >= 16 && sizeof(PointTarget) >= 16,
                "PCL point types start with float data[4] = {x, y, z, 1}: align() writes those 16 bytes of every output point");

  // ndt_omp_impl.hpp:46-76: resolution 1.0, step 0.1, outlier ratio 0.55, epsilon 0.1, 35 iterations, DIRECT7
  explicit NormalDistributionsTransform(int device = 0) : search_method(DIRECT7), h_(nullptr), device_(device), trans_probability_(0.0) {
    reg_name_ = "NormalDistributionsTransform";
    const int st = ndtb200_create(&h_, device);
    if (st != NDTB200_OK) {
      std::fprintf(stderr, "[pclomp_b200::NormalDistributionsTransform] no usable CUDA device (status %d); "
                           "this library has no CPU fallback\n", st);
      h_ = nullptr;
    }
    ndtb200_default_params(&prm_);
    transformation_epsilon_ = prm_.trans_eps;
    max_iterations_ = prm_.max_iterations;
  }
  ~NormalDistributionsTransform() { if (h_) ndtb200_destroy(h_); }

  // copy construction / assignment: the mapping node returns the object by value (ndt_omp_mapping_node.cpp:151-169)
  NormalDistributionsTransform(const NormalDistributionsTransform& o)
      : Base(o), search_method(o.search_method), h_(nullptr), device_(o.device_), prm_(o.prm_), trans_probability_(o.trans_probability_) {
    if (o.h_) ndtb200_clone(o.h_, &h_);
  }
  NormalDistributionsTransform& operator=(const NormalDistributionsTransform& o) {
    if (this != &o) {
      Base::operator=(o);
      if (h_) ndtb200_destroy(h_);
      h_ = nullptr;
      search_method = o.search_method; device_ = o.device_; prm_ = o.prm_; trans_probability_ = o.trans_probability_;
      if (o.h_) ndtb200_clone(o.h_, &h_);
    }
    return *this;
  }
  NormalDistributionsTransform(NormalDistributionsTransform&& o) noexcept
      : Base(o), search_method(o.search_method), h_(o.h_), device_(o.device_), prm_(o.prm_), trans_probability_(o.trans_probability_) {
    o.h_ = nullptr;
  }

  // ---- setters / getters (ndt_omp.h:115-209); setMaximumIterations / setTransformationEpsilon / hasConverged /
  //      getFinalTransformation / getLastIncrementalTransformation are pcl::Registration's own ----
  void setNumThreads(int) {}  // accepted, meaningless on the device path
  void setResolution(float resolution) { prm_.resolution = resolution; push(); }
  float getResolution() const { return prm_.resolution; }
  double getStepSize() const { return prm_.step_size; }
  void setStepSize(double step_size) { prm_.step_size = step_size; }
  double getOutlierRatio() const { return prm_.outlier_ratio; }
  void setOutlierRatio(double outlier_ratio) { prm_.outlier_ratio = outlier_ratio; }
  void setNeighborhoodSearchMethod(NeighborSearchMethod method) { search_method = method; }

  void setInputTarget(const PointCloudTargetConstPtr& cloud) override {
    Base::setInputTarget(cloud);  // ndt_omp.h:122-127: pcl::Registration::setInputTarget(cloud); init();
    if (!h_) return;
    const void* p = (cloud && !cloud->points.empty()) ? static_cast<const void*>(cloud->points.data()) : nullptr;
    const size_t n = cloud ? cloud->points.size() : 0;
    push();
    const int st = ndtb200_set_target(h_, p, n, sizeof(PointTarget), cloud ? (cloud->is_dense ? 1 : 0) : 1);
    if (st == NDTB200_ERR_NO_INPUT)
      std::fprintf(stderr, "[pclomp_b200::VoxelGridCovariance::applyFilter] No input dataset given!\n");
    else if (st == NDTB200_ERR_GRID_OVERFLOW)
      std::fprintf(stderr, "[pclomp_b200::VoxelGridCovariance::applyFilter] Leaf size is too small for the input dataset. "
                           "Integer indices would overflow.\n");
    else if (st != NDTB200_OK)
      std::fprintf(stderr, "[pclomp_b200] setInputTarget failed: %s\n", ndtb200_last_error(h_));
  }
  void setInputSource(const PointCloudSourceConstPtr& cloud) override {
    Base::setInputSource(cloud);
    if (!h_) return;
    const void* p = (cloud && !cloud->points.empty()) ? static_cast<const void*>(cloud->points.data()) : nullptr;
    const int st = ndtb200_set_source(h_, p, cloud ? cloud->points.size() : 0, sizeof(PointSource));
    if (st != NDTB200_OK) std::fprintf(stderr, "[pclomp_b200] setInputSource failed: %s\n", ndtb200_last_error(h_));
  }

  int getFinalNumIteration() const { return nr_iterations_; }
  double getTransformationProbability() const { return trans_probability_; }
  // pcl::Registration::getFitnessScore — mean squared distance of every transformed source point to its nearest raw
  // target point — on the device (exact nearest neighbours).  With the real PCL base class a call through a
  // pcl::Registration pointer runs upstream's own kd-tree implementation on the same final_transformation_.
  double getFitnessScore(double max_range = std::numeric_limits<double>::max())
#if !PCLOMP_B200_HAVE_PCL_REGISTRATION
      override
#endif
  {
    double v = std::numeric_limits<double>::max();
    if (h_) ndtb200_fitness_score(h_, max_range, &v);
    return v;
  }
  // negative log likelihood of an already transformed cloud (ndt_omp_impl.hpp:935-983)
  double calculateScore(const PointCloudSource& cloud) {
    double v = 0.0;
    if (h_ && !cloud.points.empty()) {
      push();
      ndtb200_calculate_score(h_, cloud.points.data(), cloud.points.size(), sizeof(PointSource), &v);
    }
    return v;
  }

  // calculateScore of MANY already transformed clouds in one upload + one launch (loop-closure screening; not in the
  // reference, which would loop over calculateScore)
  std::vector<double> calculateScoreBatch(const std::vector<const PointCloudSource*>& clouds) {
    std::vector<double> out(clouds.size(), 0.0);
    if (!h_ || clouds.empty()) return out;
    push();
    std::vector<size_t> offs(clouds.size() + 1, 0);
    for (size_t c = 0; c < clouds.size(); ++c) offs[c + 1] = offs[c] + clouds[c]->points.size();
    std::vector<PointSource> all;
    all.reserve(offs.back());
    for (const PointCloudSource* c : clouds) all.insert(all.end(), c->points.begin(), c->points.end());
    if (!all.empty())
      ndtb200_calculate_score_batch(h_, all.data(), offs.data(), static_cast<int>(clouds.size()), sizeof(PointSource), out.data());
    return out;
  }
  // calculateScore of the CURRENT source under many candidate poses (only the poses are uploaded)
  template <class PoseContainer>  // any sequence of Eigen::Matrix4f (std::vector with or without Eigen's aligned allocator)
  std::vector<double> scorePoses(const PoseContainer& poses) {
    std::vector<double> out(poses.size(), 0.0);
    if (!h_ || poses.empty()) return out;
    push();
    std::vector<float> flat(poses.size() * 16);
    for (size_t k = 0; k < poses.size(); ++k) std::memcpy(&flat[k * 16], poses[k].data(), 16 * sizeof(float));
    ndtb200_score_poses(h_, flat.data(), static_cast<int>(poses.size()), out.data());
    return out;
  }

  // [x, y, z, roll, pitch, yaw] -> Translation * AngleAxis(roll, X) * AngleAxis(pitch, Y) * AngleAxis(yaw, Z), fp32
  // (ndt_omp.h:216-233)
  static void convertTransform(const Eigen::Matrix<double, 6, 1>& x, Eigen::Affine3f& trans) {
#if PCLOMP_B200_HAVE_EIGEN
    trans = Eigen::Translation<float, 3>(float(x(0)), float(x(1)), float(x(2))) *
            Eigen::AngleAxis<float>(float(x(3)), Eigen::Vector3f::UnitX()) *
            Eigen::AngleAxis<float>(float(x(4)), Eigen::Vector3f::UnitY()) *
            Eigen::AngleAxis<float>(float(x(5)), Eigen::Vector3f::UnitZ());
#else
    const double p[6] = {x(0), x(1), x(2), x(3), x(4), x(5)};
    ndtb200_pose_to_matrix(p, trans.matrix().data());  // the same fp32 arithmetic the solver uses for its trial poses
#endif
  }
  static void convertTransform(const Eigen::Matrix<double, 6, 1>& x, Eigen::Matrix4f& trans) {
    Eigen::Affine3f _affine;
    convertTransform(x, _affine);
    trans = _affine.matrix();
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
      o->push();
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
    for (int i = 0; i < n; ++i) objects[i]->fetch_result();
  }

  NeighborSearchMethod search_method;  // public in the reference too (ndt_omp.h:499)

 protected:
  // pcl::Registration::align() has already copied the source into `output` (data[3] = 1) and reset the transforms
  // (ndt_omp_impl.hpp:80-171 runs from here in the reference)
  void computeTransformation(PointCloudSource& output, const Eigen::Matrix4f& guess) override {
    nr_iterations_ = 0;
    converged_ = false;
    if (!h_ || !input_) return;
    push();  // the public search_method field and the base class's epsilon / iteration cap are read at align time
    const int st = ndtb200_align(h_, guess.data(), output.points.empty() ? nullptr : output.points.data(), sizeof(PointSource));
    if (st != NDTB200_OK) { std::fprintf(stderr, "[pclomp_b200] align failed: %s\n", ndtb200_last_error(h_)); return; }
    fetch_result();
  }

 private:
  // The C ABI keeps the reference's rule itself (ndtb200_set_params: a CHANGED resolution rebuilds the map only if a
  // source is set, ndt_omp.h:132-142), so the whole parameter block can be pushed at any time.
  void push() {
    if (!h_) return;
    prm_.search_method = static_cast<int32_t>(search_method);
    prm_.trans_eps = transformation_epsilon_;
    prm_.max_iterations = max_iterations_;
    const int st = ndtb200_set_params(h_, &prm_);
    if (st == NDTB200_ERR_CUDA) std::fprintf(stderr, "[pclomp_b200] set_params failed: %s\n", ndtb200_last_error(h_));
  }
  void fetch_result() {
    ndtb200_result r;
    if (!h_ || ndtb200_get_result(h_, &r) != NDTB200_OK) return;
    for (int c = 0; c < 4; ++c)
      for (int rr = 0; rr < 4; ++rr) {
        final_transformation_(rr, c) = r.final_transformation[c * 4 + rr];
        transformation_(rr, c) = r.last_increment[c * 4 + rr];
      }
    converged_ = r.converged != 0;
    nr_iterations_ = r.iterations;
    trans_probability_ = r.trans_probability;
  }

  ndtb200_handle* h_;
  int device_;
  ndtb200_params prm_;
  double trans_probability_;
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
  explicit VoxelGrid(int device = 0) : h_(nullptr) { leaf_[0] = leaf_[1] = leaf_[2] = 0.f; if (ndtb200_create(&h_, device) != NDTB200_OK) h_ = nullptr; }
  ~VoxelGrid() { if (h_) ndtb200_destroy(h_); }
  VoxelGrid(const VoxelGrid&) = delete;
  VoxelGrid& operator=(const VoxelGrid&) = delete;
  void setLeafSize(float lx, float ly, float lz) { leaf_[0] = lx; leaf_[1] = ly; leaf_[2] = lz; }
  void setInputCloud(const PointCloudConstPtr& cloud) { input_ = cloud; }
  void filter(PointCloud& output) {
    output.points.clear();
    if (!h_ || !input_ || input_->points.empty()) { output.width = 0; output.height = 1; return; }
    struct Rec { float x, y, z, w; };
    std::vector<Rec> out(input_->points.size());
    int64_t m = 0;
    const int st = ndtb200_voxelgrid_filter3(h_, input_->points.data(), input_->points.size(), sizeof(PointT), leaf_, out.data(),
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
  float leaf_[3];
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
