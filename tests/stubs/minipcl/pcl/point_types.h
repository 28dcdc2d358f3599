// minipcl: pcl/point_types.h — the three point layouts ndt_omp instantiates (ndt_omp/src/pclomp/ndt_omp.cpp:4-6)
#pragma once
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
}  // namespace pcl
