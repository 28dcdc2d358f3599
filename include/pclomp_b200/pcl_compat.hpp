// pcl_compat.hpp — the few PCL / Eigen types the NDT callers touch, for builds WITHOUT PCL.
//
// With PCL installed, include <pcl/point_types.h> / <pcl/point_cloud.h> before ndt_b200.hpp and this header is
// skipped; the shim then works on the real pcl::PointCloud.  Layouts match PCL's: PointXYZ is a 16-byte record
// (x, y, z, padding), PointXYZI / PointXYZRGB are 32-byte records whose first 12 bytes are x, y, z
// (ndt_omp/src/pclomp/ndt_omp.cpp:4-6 instantiates exactly these three).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#if __has_include(<pcl/point_cloud.h>) && __has_include(<pcl/point_types.h>) && !defined(PCLOMP_B200_FORCE_COMPAT)
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#define PCLOMP_B200_HAVE_PCL 1
#else
#define PCLOMP_B200_HAVE_PCL 0

namespace pcl {

struct alignas(16) PointXYZ {
  union { float data[4]; struct { float x, y, z; }; };
  PointXYZ() : data{0.f, 0.f, 0.f, 1.f} {}
  PointXYZ(float x_, float y_, float z_) : data{x_, y_, z_, 1.f} {}
};
struct alignas(16) PointXYZI {
  union { float data[4]; struct { float x, y, z; }; };
  union { struct { float intensity; }; float data_c[4]; };
  PointXYZI() : data{0.f, 0.f, 0.f, 1.f}, data_c{0.f, 0.f, 0.f, 0.f} {}
};
struct alignas(16) PointXYZRGB {
  union { float data[4]; struct { float x, y, z; }; };
  union { struct { float rgb; }; float data_c[4]; };
  PointXYZRGB() : data{0.f, 0.f, 0.f, 1.f}, data_c{0.f, 0.f, 0.f, 0.f} {}
};
static_assert(sizeof(PointXYZ) == 16 && sizeof(PointXYZI) == 32 && sizeof(PointXYZRGB) == 32, "PCL point layouts");

template <typename PointT>
class PointCloud {
 public:
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  std::vector<PointT> points;
  uint32_t width = 0, height = 1;
  bool is_dense = true;
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); width = 0; height = 1; }
  void resize(size_t n) { points.resize(n); width = static_cast<uint32_t>(n); height = 1; }
  void push_back(const PointT& p) { points.push_back(p); width = static_cast<uint32_t>(points.size()); height = 1; }
  PointT& operator[](size_t i) { return points[i]; }
  const PointT& operator[](size_t i) const { return points[i]; }
  PointCloud& operator+=(const PointCloud& o) {  // the mapping nodes accumulate clouds with += (ndt_rosbag_mapping_node.cpp:155)
    points.insert(points.end(), o.points.begin(), o.points.end());
    width = static_cast<uint32_t>(points.size());
    height = 1;
    is_dense = is_dense && o.is_dense;
    return *this;
  }
};

}  // namespace pcl
#endif  // PCL available?

#include "eigen_compat.hpp"

// ---------------------------------------------------------------------------------------------------------------
// pcl::Registration<PointSource, PointTarget> stand-in for builds WITHOUT PCL: the members and the align() wrapper
// (upstream pcl/registration/registration.h + impl/registration.hpp) that pclomp::NormalDistributionsTransform relies on
// (ndt_omp.h:70-71, 242-257), so that a caller written against `pcl::Registration<...>::Ptr` — the helper of
// ndt_omp/apps/align.cpp:15 — compiles unchanged.  With PCL installed the real class is used instead.
// ---------------------------------------------------------------------------------------------------------------
#if __has_include(<pcl/registration/registration.h>) && !defined(PCLOMP_B200_FORCE_COMPAT)
#include <pcl/registration/registration.h>
#define PCLOMP_B200_HAVE_PCL_REGISTRATION 1
#else
#define PCLOMP_B200_HAVE_PCL_REGISTRATION 0
#include <limits>
#include <string>
namespace pcl {
template <typename PointSource, typename PointTarget, typename Scalar = float>
class Registration {
 public:
  typedef Eigen::Matrix4f Matrix4;
  typedef std::shared_ptr<Registration<PointSource, PointTarget, Scalar>> Ptr;
  typedef std::shared_ptr<const Registration<PointSource, PointTarget, Scalar>> ConstPtr;
  typedef pcl::PointCloud<PointSource> PointCloudSource;
  typedef typename PointCloudSource::Ptr PointCloudSourcePtr;
  typedef typename PointCloudSource::ConstPtr PointCloudSourceConstPtr;
  typedef pcl::PointCloud<PointTarget> PointCloudTarget;
  typedef typename PointCloudTarget::Ptr PointCloudTargetPtr;
  typedef typename PointCloudTarget::ConstPtr PointCloudTargetConstPtr;

  Registration()
      : reg_name_(), nr_iterations_(0), max_iterations_(10), final_transformation_(Matrix4::Identity()),
        transformation_(Matrix4::Identity()), previous_transformation_(Matrix4::Identity()), transformation_epsilon_(0.0),
        converged_(false) {}
  virtual ~Registration() {}

  virtual void setInputSource(const PointCloudSourceConstPtr& cloud) { input_ = cloud; }
  PointCloudSourceConstPtr const getInputSource() { return input_; }
  virtual void setInputTarget(const PointCloudTargetConstPtr& cloud) { target_ = cloud; }
  PointCloudTargetConstPtr const getInputTarget() { return target_; }
  Matrix4 getFinalTransformation() { return final_transformation_; }
  Matrix4 getLastIncrementalTransformation() { return transformation_; }
  void setMaximumIterations(int nr_iterations) { max_iterations_ = nr_iterations; }
  int getMaximumIterations() { return max_iterations_; }
  void setTransformationEpsilon(double epsilon) { transformation_epsilon_ = epsilon; }
  double getTransformationEpsilon() { return transformation_epsilon_; }
  bool hasConverged() const { return converged_; }
  // upstream: mean squared distance of the transformed source to its nearest target point (kd-tree); the stand-in
  // dispatches to the registration object, which computes it on the device
  virtual double getFitnessScore(double max_range = std::numeric_limits<double>::max()) = 0;

  void align(PointCloudSource& output) { align(output, Matrix4::Identity()); }
  void align(PointCloudSource& output, const Matrix4& guess) {
    if (!input_ || !target_) return;  // initCompute() fails without both clouds
    output.points.resize(input_->points.size());
    output.width = input_->width;
    output.height = input_->height;
    output.is_dense = input_->is_dense;
    for (size_t i = 0; i < input_->points.size(); ++i) output.points[i] = input_->points[i];
    converged_ = false;
    final_transformation_ = transformation_ = previous_transformation_ = Matrix4::Identity();
    for (size_t i = 0; i < output.points.size(); ++i) output.points[i].data[3] = 1.0f;
    computeTransformation(output, guess);
  }

 protected:
  virtual void computeTransformation(PointCloudSource& output, const Matrix4& guess) = 0;

  std::string reg_name_;
  int nr_iterations_, max_iterations_;
  PointCloudSourceConstPtr input_;
  PointCloudTargetConstPtr target_;
  Matrix4 final_transformation_, transformation_, previous_transformation_;
  double transformation_epsilon_;
  bool converged_;
};
}  // namespace pcl
#endif
