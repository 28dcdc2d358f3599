// pcd_io.hpp — PCD v0.7 reader / writer for the formats the reference's drivers use:
// `pcl::io::loadPCDFile` on ndt_omp/data/*.pcd (ndt_omp/apps/align.cpp:48-55: "DATA binary", float32 x y z intensity)
// and `pcl::io::savePCDFileBinary` of the accumulated map (lidar_subscriber/src/lidar_subscriber_node.cpp:38-46).
// Header-only, no PCL.  Reads "DATA binary" and "DATA ascii" files whose fields include float32 x, y, z; writes
// "DATA binary" with FIELDS x y z.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <cstdlib>
#include <sstream>
#include <string>
#include <vector>

#include "pcl_compat.hpp"

namespace pclomp_b200 {
namespace io {

// returns 0 on success, -1 on failure (like pcl::io::loadPCDFile)
template <typename PointT>
int loadPCDFile(const std::string& path, pcl::PointCloud<PointT>& cloud) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return -1;
  std::vector<std::string> fields;
  std::vector<int> sizes, counts;
  std::vector<char> types;
  size_t npts = 0, width = 0, height = 1;
  std::string data_kind, line;
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty() || line[0] == '#') continue;
    std::istringstream ls(line);
    std::string key;
    ls >> key;
    if (key == "FIELDS") { std::string t; while (ls >> t) fields.push_back(t); }
    else if (key == "SIZE") { int v; while (ls >> v) sizes.push_back(v); }
    else if (key == "TYPE") { char c; while (ls >> c) types.push_back(c); }
    else if (key == "COUNT") { int v; while (ls >> v) counts.push_back(v); }
    else if (key == "WIDTH") ls >> width;
    else if (key == "HEIGHT") ls >> height;
    else if (key == "POINTS") ls >> npts;
    else if (key == "DATA") { ls >> data_kind; break; }
  }
  if (fields.empty() || sizes.size() != fields.size() || data_kind.empty()) return -1;
  if (counts.empty()) counts.assign(fields.size(), 1);
  if (types.size() != fields.size()) types.assign(fields.size(), 'F');
  if (npts == 0) npts = width * height;
  int off[3] = {-1, -1, -1}, col[3] = {-1, -1, -1};
  size_t stride = 0, ncols = 0;
  for (size_t i = 0; i < fields.size(); ++i) {
    for (int a = 0; a < 3; ++a)
      if (fields[i] == std::string(1, static_cast<char>('x' + a)) && sizes[i] == 4 && types[i] == 'F') { off[a] = static_cast<int>(stride); col[a] = static_cast<int>(ncols); }
    stride += static_cast<size_t>(sizes[i]) * counts[i];
    ncols += counts[i];
  }
  if (off[0] < 0 || off[1] < 0 || off[2] < 0) return -1;
  cloud.points.clear();
  cloud.points.reserve(npts);
  bool dense = true;
  if (data_kind == "binary") {
    std::vector<char> raw(npts * stride);
    f.read(raw.data(), static_cast<std::streamsize>(raw.size()));
    if (static_cast<size_t>(f.gcount()) != raw.size()) return -1;
    for (size_t i = 0; i < npts; ++i) {
      float xyz[3];
      for (int a = 0; a < 3; ++a) std::memcpy(&xyz[a], raw.data() + i * stride + off[a], 4);
      PointT p = PointT();
      p.x = xyz[0]; p.y = xyz[1]; p.z = xyz[2];
      dense = dense && (xyz[0] == xyz[0]) && (xyz[1] == xyz[1]) && (xyz[2] == xyz[2]);
      cloud.points.push_back(p);
    }
  } else if (data_kind == "ascii") {
    for (size_t i = 0; i < npts && std::getline(f, line); ++i) {
      std::istringstream ls(line);
      std::vector<double> v;
      std::string tok;
      while (ls >> tok) v.push_back(tok == "nan" ? std::numeric_limits<double>::quiet_NaN() : std::atof(tok.c_str()));
      if (v.size() < ncols) return -1;
      PointT p = PointT();
      p.x = static_cast<float>(v[col[0]]); p.y = static_cast<float>(v[col[1]]); p.z = static_cast<float>(v[col[2]]);
      dense = dense && (p.x == p.x) && (p.y == p.y) && (p.z == p.z);
      cloud.points.push_back(p);
    }
  } else {
    return -1;  // binary_compressed is not used by the reference's data or drivers
  }
  cloud.width = static_cast<uint32_t>(cloud.points.size());
  cloud.height = 1;
  cloud.is_dense = dense;
  return 0;
}

// "DATA binary", FIELDS x y z (what savePCDFileBinary writes for PointXYZ, minus the padding word)
template <typename PointT>
int savePCDFileBinary(const std::string& path, const pcl::PointCloud<PointT>& cloud) {
  std::ofstream f(path, std::ios::binary);
  if (!f) return -1;
  const size_t n = cloud.points.size();
  f << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n"
    << "WIDTH " << n << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA binary\n";
  std::vector<float> raw(n * 3);
  for (size_t i = 0; i < n; ++i) { raw[3 * i] = cloud.points[i].x; raw[3 * i + 1] = cloud.points[i].y; raw[3 * i + 2] = cloud.points[i].z; }
  f.write(reinterpret_cast<const char*>(raw.data()), static_cast<std::streamsize>(raw.size() * sizeof(float)));
  return f ? 0 : -1;
}

}  // namespace io
}  // namespace pclomp_b200
