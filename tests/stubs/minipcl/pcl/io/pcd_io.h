// minipcl: pcl/io/pcd_io.h — pcl::io::loadPCDFile / savePCDFileBinary on the repo's PCD v0.7 reader / writer
#pragma once
#include <string>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pclomp_b200/pcd_io.hpp>
namespace pcl {
namespace io {
template <typename PointT> int loadPCDFile(const std::string& path, pcl::PointCloud<PointT>& cloud) { return pclomp_b200::io::loadPCDFile(path, cloud); }
template <typename PointT> int savePCDFileBinary(const std::string& path, const pcl::PointCloud<PointT>& cloud) { return pclomp_b200::io::savePCDFileBinary(path, cloud); }
}  // namespace io
}  // namespace pcl
