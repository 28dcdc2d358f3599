// minipcl: pcl/filters/voxel_grid.h — pcl::VoxelGrid<PointT> (setLeafSize / setInputCloud / filter) on the device filter
#pragma once
#include <pclomp_b200/ndt_b200.hpp>
namespace pcl {
template <typename PointT> using VoxelGrid = pclomp_b200::VoxelGrid<PointT>;
}  // namespace pcl
