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
};

}  // namespace pcl
#endif  // PCL available?

#if __has_include(<Eigen/Core>) && !defined(PCLOMP_B200_FORCE_COMPAT)
#include <Eigen/Core>
#define PCLOMP_B200_HAVE_EIGEN 1
#else
#define PCLOMP_B200_HAVE_EIGEN 0
namespace Eigen {
// Column-major 4x4 float with the handful of members the NDT callers use.
struct Matrix4f {
  float m[16];
  Matrix4f() { std::memset(m, 0, sizeof(m)); }
  static Matrix4f Identity() {
    Matrix4f r;
    r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f;
    return r;
  }
  float& operator()(int r, int c) { return m[c * 4 + r]; }
  float operator()(int r, int c) const { return m[c * 4 + r]; }
  const float* data() const { return m; }
  float* data() { return m; }
  Matrix4f operator*(const Matrix4f& o) const {
    Matrix4f r;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        float s = 0.f;
        for (int k = 0; k < 4; ++k) s += (*this)(i, k) * o(k, j);
        r(i, j) = s;
      }
    return r;
  }
  bool operator!=(const Matrix4f& o) const { return std::memcmp(m, o.m, sizeof(m)) != 0; }
};
}  // namespace Eigen
#endif
