// minipcl: pcl/registration/registration.h — the shape of upstream pcl::Registration (PCL 1.10): protected state,
// virtual setInputSource / setInputTarget, NON-virtual align() and getFitnessScore(), pure virtual
// computeTransformation().  Eigen::Matrix4f comes from the repo's stand-in (Eigen is not installed here).
#pragma once
#include <limits>
#include <string>
#include <vector>

#include <pcl/point_cloud.h>
#include <pclomp_b200/eigen_compat.hpp>

namespace pcl {
template <typename PointSource, typename PointTarget, typename Scalar = float>
class Registration {
 public:
  typedef Eigen::Matrix4f Matrix4;
  typedef shared_ptr<Registration<PointSource, PointTarget, Scalar>> Ptr;
  typedef shared_ptr<const Registration<PointSource, PointTarget, Scalar>> ConstPtr;
  typedef pcl::PointCloud<PointSource> PointCloudSource;
  typedef typename PointCloudSource::Ptr PointCloudSourcePtr;
  typedef typename PointCloudSource::ConstPtr PointCloudSourceConstPtr;
  typedef pcl::PointCloud<PointTarget> PointCloudTarget;
  typedef typename PointCloudTarget::Ptr PointCloudTargetPtr;
  typedef typename PointCloudTarget::ConstPtr PointCloudTargetConstPtr;

  Registration()
      : reg_name_(), nr_iterations_(0), max_iterations_(10), final_transformation_(Matrix4::Identity()),
        transformation_(Matrix4::Identity()), previous_transformation_(Matrix4::Identity()), transformation_epsilon_(0.0),
        converged_(false), target_cloud_updated_(true), source_cloud_updated_(true) {}
  virtual ~Registration() {}

  virtual void setInputSource(const PointCloudSourceConstPtr& cloud) { source_cloud_updated_ = true; input_ = cloud; }
  PointCloudSourceConstPtr const getInputSource() { return input_; }
  virtual inline void setInputTarget(const PointCloudTargetConstPtr& cloud) {
    if (!cloud || cloud->points.empty()) return;  // upstream: PCL_ERROR "Invalid or empty point cloud dataset given!"
    target_ = cloud;
    target_cloud_updated_ = true;
  }
  PointCloudTargetConstPtr const getInputTarget() { return target_; }
  inline Matrix4 getFinalTransformation() { return final_transformation_; }
  inline Matrix4 getLastIncrementalTransformation() { return transformation_; }
  inline void setMaximumIterations(int nr_iterations) { max_iterations_ = nr_iterations; }
  inline int getMaximumIterations() { return max_iterations_; }
  inline void setTransformationEpsilon(double epsilon) { transformation_epsilon_ = epsilon; }
  inline double getTransformationEpsilon() { return transformation_epsilon_; }
  inline bool hasConverged() const { return converged_; }

  // upstream: transform the source by final_transformation_, nearest neighbour of every point in the raw target through
  // the kd-tree, mean of the squared distances <= max_range.  Brute force here; fp32 squared distances (FLANN L2_Simple).
  inline double getFitnessScore(double max_range = std::numeric_limits<double>::max()) {
    if (!input_ || !target_ || target_->points.empty()) return std::numeric_limits<double>::max();
    const Matrix4& T = final_transformation_;
    double sum = 0.0;
    int nr = 0;
    for (const PointSource& p : input_->points) {
      const float x = T(0, 0) * p.x + (T(0, 1) * p.y + (T(0, 2) * p.z + T(0, 3)));
      const float y = T(1, 0) * p.x + (T(1, 1) * p.y + (T(1, 2) * p.z + T(1, 3)));
      const float z = T(2, 0) * p.x + (T(2, 1) * p.y + (T(2, 2) * p.z + T(2, 3)));
      float best = std::numeric_limits<float>::max();
      for (const PointTarget& q : target_->points) {
        const float dx = x - q.x, dy = y - q.y, dz = z - q.z;
        const float d = (dx * dx + dy * dy) + dz * dz;
        if (d < best) best = d;
      }
      if (static_cast<double>(best) <= max_range) { sum += best; ++nr; }
    }
    return nr > 0 ? sum / nr : std::numeric_limits<double>::max();
  }

  inline void align(PointCloudSource& output) { align(output, Matrix4::Identity()); }
  inline void align(PointCloudSource& output, const Matrix4& guess) {
    if (!initCompute()) return;
    output.points.resize(input_->points.size());
    output.header = input_->header;
    output.width = input_->width;
    output.height = input_->height;
    output.is_dense = input_->is_dense;
    for (size_t i = 0; i < input_->points.size(); ++i) output.points[i] = input_->points[i];
    converged_ = false;
    final_transformation_ = transformation_ = previous_transformation_ = Matrix4::Identity();
    for (size_t i = 0; i < output.points.size(); ++i) output.points[i].data[3] = 1.0;
    computeTransformation(output, guess);
  }

 protected:
  bool initCompute() {
    if (!target_ || !input_) return false;   // upstream: PCL_ERROR "No input target dataset was given!"
    target_cloud_updated_ = false;           // upstream builds tree_ over the target here (the `single` - `10times/10` gap)
    return true;
  }
  virtual void computeTransformation(PointCloudSource& output, const Matrix4& guess) = 0;

  std::string reg_name_;
  int nr_iterations_, max_iterations_;
  PointCloudSourceConstPtr input_;
  PointCloudTargetConstPtr target_;
  Matrix4 final_transformation_, transformation_, previous_transformation_;
  double transformation_epsilon_;
  bool converged_, target_cloud_updated_, source_cloud_updated_;
};
}  // namespace pcl
