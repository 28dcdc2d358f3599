// minipcl: pcl/point_cloud.h
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>
namespace pcl {
template <typename T> using shared_ptr = std::shared_ptr<T>;  // PCL >= 1.11
struct PCLHeader { uint32_t seq = 0; uint64_t stamp = 0; };
template <typename PointT>
class PointCloud {
 public:
  typedef shared_ptr<PointCloud<PointT>> Ptr;
  typedef shared_ptr<const PointCloud<PointT>> ConstPtr;
  PCLHeader header;
  std::vector<PointT> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); width = 0; height = 0; }
  void push_back(const PointT& p) { points.push_back(p); width = static_cast<uint32_t>(points.size()); height = 1; }
  PointT& operator[](size_t i) { return points[i]; }
  const PointT& operator[](size_t i) const { return points[i]; }
  PointCloud& operator+=(const PointCloud& o) {
    points.insert(points.end(), o.points.begin(), o.points.end());
    width = static_cast<uint32_t>(points.size());
    height = 1;
    is_dense = is_dense && o.is_dense;
    return *this;
  }
};
}  // namespace pcl
