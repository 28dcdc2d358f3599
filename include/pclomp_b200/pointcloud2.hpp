// pointcloud2.hpp — ROS-free sensor_msgs/PointCloud2 for the replay of recorded scans: the message the
// reference's nodes receive (lidar_subscriber/src/ndt_rosbag_mapping_node.cpp:42-60 reads it from a bag,
// lidar_subscriber_node.cpp:35-39 from a topic) in its ROS 1 wire format, and the pcl::fromROSMsg conversion to a point
// cloud.  Header-only, no ROS, no PCL.
//
// Wire format (ROS 1 message serialization, little endian, no padding):
//   std_msgs/Header  : uint32 seq | uint32 stamp.sec | uint32 stamp.nsec | string frame_id
//   uint32 height | uint32 width
//   PointField[]     : uint32 count, then per field: string name | uint32 offset | uint8 datatype | uint32 count
//   uint8 is_bigendian | uint32 point_step | uint32 row_step
//   uint8[] data     : uint32 length + bytes
//   uint8 is_dense
//   (string = uint32 length + bytes; datatypes: 1 INT8 2 UINT8 3 INT16 4 UINT16 5 INT32 6 UINT32 7 FLOAT32 8 FLOAT64)
// A "dump" is a plain concatenation of records [uint32 message_length][message bytes] — what `rosbag` stores as the
// data of each message record, without the bag container.
#pragma once
#include <cstdint>
#include <cstring>
#include <fstream>
#include <limits>
#include <string>
#include <vector>

#include "pcl_compat.hpp"

namespace pclomp_b200 {

struct PointField {
  enum { INT8 = 1, UINT8 = 2, INT16 = 3, UINT16 = 4, INT32 = 5, UINT32 = 6, FLOAT32 = 7, FLOAT64 = 8 };
  std::string name;
  uint32_t offset = 0;
  uint8_t datatype = 0;
  uint32_t count = 1;
};

struct PointCloud2 {
  uint32_t seq = 0, stamp_sec = 0, stamp_nsec = 0;
  std::string frame_id;
  uint32_t height = 1, width = 0;
  std::vector<PointField> fields;
  bool is_bigendian = false;
  uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  bool is_dense = true;
};

namespace detail {
struct Reader {
  const uint8_t* p;
  size_t left;
  bool ok = true;
  template <typename T> T get() {
    T v = T();
    if (left < sizeof(T)) { ok = false; left = 0; return v; }
    std::memcpy(&v, p, sizeof(T));
    p += sizeof(T); left -= sizeof(T);
    return v;
  }
  std::string str() {
    const uint32_t n = get<uint32_t>();
    if (!ok || left < n) { ok = false; left = 0; return std::string(); }
    std::string s(reinterpret_cast<const char*>(p), n);
    p += n; left -= n;
    return s;
  }
};
template <typename T> inline void put(std::vector<uint8_t>& out, T v) {
  const uint8_t* b = reinterpret_cast<const uint8_t*>(&v);
  out.insert(out.end(), b, b + sizeof(T));
}
inline void put_str(std::vector<uint8_t>& out, const std::string& s) {
  put<uint32_t>(out, static_cast<uint32_t>(s.size()));
  out.insert(out.end(), s.begin(), s.end());
}
}  // namespace detail

// returns false on a truncated / malformed message
inline bool deserialize(const uint8_t* buf, size_t len, PointCloud2& m) {
  detail::Reader r{buf, len};
  m.seq = r.get<uint32_t>(); m.stamp_sec = r.get<uint32_t>(); m.stamp_nsec = r.get<uint32_t>();
  m.frame_id = r.str();
  m.height = r.get<uint32_t>(); m.width = r.get<uint32_t>();
  const uint32_t nf = r.get<uint32_t>();
  if (!r.ok || nf > 4096) return false;
  m.fields.resize(nf);
  for (uint32_t i = 0; i < nf; ++i) {
    m.fields[i].name = r.str();
    m.fields[i].offset = r.get<uint32_t>();
    m.fields[i].datatype = r.get<uint8_t>();
    m.fields[i].count = r.get<uint32_t>();
  }
  m.is_bigendian = r.get<uint8_t>() != 0;
  m.point_step = r.get<uint32_t>(); m.row_step = r.get<uint32_t>();
  const uint32_t nd = r.get<uint32_t>();
  if (!r.ok || r.left < nd) return false;
  m.data.assign(r.p, r.p + nd);
  r.p += nd; r.left -= nd;
  m.is_dense = r.get<uint8_t>() != 0;
  return r.ok;
}

inline void serialize(const PointCloud2& m, std::vector<uint8_t>& out) {
  out.clear();
  detail::put<uint32_t>(out, m.seq); detail::put<uint32_t>(out, m.stamp_sec); detail::put<uint32_t>(out, m.stamp_nsec);
  detail::put_str(out, m.frame_id);
  detail::put<uint32_t>(out, m.height); detail::put<uint32_t>(out, m.width);
  detail::put<uint32_t>(out, static_cast<uint32_t>(m.fields.size()));
  for (const PointField& f : m.fields) {
    detail::put_str(out, f.name);
    detail::put<uint32_t>(out, f.offset); detail::put<uint8_t>(out, f.datatype); detail::put<uint32_t>(out, f.count);
  }
  detail::put<uint8_t>(out, m.is_bigendian ? 1 : 0);
  detail::put<uint32_t>(out, m.point_step); detail::put<uint32_t>(out, m.row_step);
  detail::put<uint32_t>(out, static_cast<uint32_t>(m.data.size()));
  out.insert(out.end(), m.data.begin(), m.data.end());
  detail::put<uint8_t>(out, m.is_dense ? 1 : 0);
}

namespace detail {
inline float read_scalar_as_float(const uint8_t* p, uint8_t datatype, bool swap) {
  auto load = [&](void* dst, size_t n) {
    uint8_t tmp[8];
    for (size_t i = 0; i < n; ++i) tmp[i] = swap ? p[n - 1 - i] : p[i];
    std::memcpy(dst, tmp, n);
  };
  switch (datatype) {
    case PointField::INT8: { int8_t v; load(&v, 1); return static_cast<float>(v); }
    case PointField::UINT8: { uint8_t v; load(&v, 1); return static_cast<float>(v); }
    case PointField::INT16: { int16_t v; load(&v, 2); return static_cast<float>(v); }
    case PointField::UINT16: { uint16_t v; load(&v, 2); return static_cast<float>(v); }
    case PointField::INT32: { int32_t v; load(&v, 4); return static_cast<float>(v); }
    case PointField::UINT32: { uint32_t v; load(&v, 4); return static_cast<float>(v); }
    case PointField::FLOAT32: { float v; load(&v, 4); return v; }
    case PointField::FLOAT64: { double v; load(&v, 8); return static_cast<float>(v); }
    default: return std::numeric_limits<float>::quiet_NaN();
  }
}
}  // namespace detail

// pcl::fromROSMsg for the x / y / z (and, when the point type has it and the message carries it, intensity) fields.
// Honours field offsets / datatypes, point_step / row_step, is_bigendian and is_dense.  Returns false when the message
// has no x, y, z fields or its sizes are inconsistent.
template <typename PointT>
bool fromROSMsg(const PointCloud2& m, pcl::PointCloud<PointT>& cloud) {
  const PointField *fx = nullptr, *fy = nullptr, *fz = nullptr;
  for (const PointField& f : m.fields) {
    if (f.name == "x") fx = &f; else if (f.name == "y") fy = &f; else if (f.name == "z") fz = &f;
  }
  if (!fx || !fy || !fz || m.point_step == 0) return false;
  const size_t npts = static_cast<size_t>(m.width) * m.height;
  const size_t row_step = m.row_step ? m.row_step : static_cast<size_t>(m.width) * m.point_step;
  if (npts && (static_cast<size_t>(m.height - 1) * row_step + static_cast<size_t>(m.width) * m.point_step > m.data.size())) return false;
  const uint16_t one = 1;
  const bool host_big = *reinterpret_cast<const uint8_t*>(&one) == 0;
  const bool swap = m.is_bigendian != host_big;
  cloud.points.resize(npts);
  cloud.width = m.width;
  cloud.height = m.height;
  cloud.is_dense = m.is_dense;
  for (uint32_t r = 0; r < m.height; ++r)
    for (uint32_t c = 0; c < m.width; ++c) {
      const uint8_t* p = m.data.data() + static_cast<size_t>(r) * row_step + static_cast<size_t>(c) * m.point_step;
      PointT q = PointT();
      q.x = detail::read_scalar_as_float(p + fx->offset, fx->datatype, swap);
      q.y = detail::read_scalar_as_float(p + fy->offset, fy->datatype, swap);
      q.z = detail::read_scalar_as_float(p + fz->offset, fz->datatype, swap);
      cloud.points[static_cast<size_t>(r) * m.width + c] = q;
    }
  return true;
}

// pcl::toROSMsg for x, y, z (FLOAT32, point_step 16 like PointXYZ)
template <typename PointT>
void toROSMsg(const pcl::PointCloud<PointT>& cloud, PointCloud2& m) {
  m.height = 1;
  m.width = static_cast<uint32_t>(cloud.points.size());
  m.fields.resize(3);
  const char* names[3] = {"x", "y", "z"};
  for (int i = 0; i < 3; ++i) { m.fields[i].name = names[i]; m.fields[i].offset = 4 * i; m.fields[i].datatype = PointField::FLOAT32; m.fields[i].count = 1; }
  m.is_bigendian = false;
  m.point_step = 16;
  m.row_step = m.point_step * m.width;
  m.data.assign(static_cast<size_t>(m.row_step), 0);
  for (size_t i = 0; i < cloud.points.size(); ++i) {
    const float xyz[3] = {cloud.points[i].x, cloud.points[i].y, cloud.points[i].z};
    std::memcpy(m.data.data() + i * 16, xyz, 12);
  }
  m.is_dense = cloud.is_dense;
}

// Dump reader: [uint32 length][message] records.  next() returns false at the end of the file or on a bad record.
class PointCloud2DumpReader {
 public:
  explicit PointCloud2DumpReader(const std::string& path) : f_(path, std::ios::binary) {}
  bool ok() const { return static_cast<bool>(f_); }
  bool next(PointCloud2& m) {
    uint32_t len = 0;
    if (!f_.read(reinterpret_cast<char*>(&len), 4)) return false;
    buf_.resize(len);
    if (len && !f_.read(reinterpret_cast<char*>(buf_.data()), len)) return false;
    return deserialize(buf_.data(), buf_.size(), m);
  }

 private:
  std::ifstream f_;
  std::vector<uint8_t> buf_;
};

inline bool appendToDump(std::ofstream& f, const PointCloud2& m) {
  std::vector<uint8_t> buf;
  serialize(m, buf);
  const uint32_t len = static_cast<uint32_t>(buf.size());
  f.write(reinterpret_cast<const char*>(&len), 4);
  f.write(reinterpret_cast<const char*>(buf.data()), buf.size());
  return static_cast<bool>(f);
}

}  // namespace pclomp_b200
