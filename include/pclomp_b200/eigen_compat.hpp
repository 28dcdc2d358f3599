// eigen_compat.hpp — the few Eigen types on the NDT class's public surface, for builds WITHOUT Eigen
// (Eigen::Matrix4f of align / getFinalTransformation; Eigen::Affine3f and Eigen::Matrix<double, 6, 1> of the static
// convertTransform, ndt_omp.h:216-233).  With Eigen installed the real headers are used.
#pragma once
#include <cstring>

#if __has_include(<Eigen/Core>) && __has_include(<Eigen/Geometry>) && !defined(PCLOMP_B200_FORCE_COMPAT)
#include <Eigen/Core>
#include <Eigen/Geometry>
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
  bool operator==(const Matrix4f& o) const { return std::memcmp(m, o.m, sizeof(m)) == 0; }
};
// the two other Eigen types on the class's public surface (convertTransform, ndt_omp.h:216-233)
template <typename Scalar, int Rows, int Cols> struct Matrix;
template <> struct Matrix<double, 6, 1> {
  double v[6];
  Matrix() { for (double& x : v) x = 0.0; }
  static Matrix Zero() { return Matrix(); }
  double& operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
};
struct Affine3f {
  Matrix4f m;
  Affine3f() : m(Matrix4f::Identity()) {}
  Matrix4f& matrix() { return m; }
  const Matrix4f& matrix() const { return m; }
};
}  // namespace Eigen
#endif

