// apps/replay_b200.cpp — ROS-free replay of recorded sensor_msgs/PointCloud2 messages through the mapping-node loop
// (lidar_subscriber/src/ndt_rosbag_mapping_node.cpp:42-75: for every message of the bag: fromROSMsg -> downsample ->
// NDT against the previous scan with the previous transform as guess -> pose chaining -> global map), on the
// device-resident pipeline (pclomp_b200::Mapper).  Input: a dump of [uint32 length][serialized PointCloud2] records
// (pclomp_b200/pointcloud2.hpp); output: one line per scan and, optionally, the global map as a binary PCD.
//
//   replay_b200 scans.pc2dump [--voxel-leaf 0.3] [--map-voxel 0.5] [--save-map map.pcd] [--rewrite out.pc2dump]
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>

#include <pclomp_b200/ndt_b200.hpp>
#include <pclomp_b200/pcd_io.hpp>
#include <pclomp_b200/pointcloud2.hpp>

int main(int argc, char** argv) {
  if (argc < 2) {
    std::cout << "usage: replay_b200 scans.pc2dump [--voxel-leaf 0.3] [--map-voxel 0.5] [--save-map map.pcd] [--rewrite out.pc2dump]" << std::endl;
    return 0;
  }
  float voxel_leaf = 0.3f, map_voxel = 0.5f;  // the node's defaults (ndt_rosbag_mapping_node.cpp:88-89)
  std::string save_map, rewrite;
  for (int i = 2; i + 1 < argc; i += 2) {
    const std::string k = argv[i];
    if (k == "--voxel-leaf") voxel_leaf = static_cast<float>(std::atof(argv[i + 1]));
    else if (k == "--map-voxel") map_voxel = static_cast<float>(std::atof(argv[i + 1]));
    else if (k == "--save-map") save_map = argv[i + 1];
    else if (k == "--rewrite") rewrite = argv[i + 1];
  }
  pclomp_b200::PointCloud2DumpReader reader(argv[1]);
  if (!reader.ok()) { std::cerr << "failed to open " << argv[1] << std::endl; return 1; }
  pclomp_b200::Mapper<pcl::PointXYZ> mapper(voxel_leaf, map_voxel);
  if (!mapper.ok()) { std::cerr << "no usable CUDA device" << std::endl; return 2; }
  std::ofstream re;
  if (!rewrite.empty()) re.open(rewrite, std::ios::binary);
  pclomp_b200::PointCloud2 msg;
  pcl::PointCloud<pcl::PointXYZ> cloud;
  int k = 0;
  std::printf("# scan seq stamp n_raw n_filtered converged iterations fitness  pose(tx ty tz)  n_map\n");
  while (reader.next(msg)) {
    if (!pclomp_b200::fromROSMsg(msg, cloud)) { std::cerr << "message " << k << ": no x/y/z fields" << std::endl; return 1; }
    const auto s = mapper.pushScan(cloud);
    std::printf("scan %d %u %u.%09u %zu %zu %d %d %.6f  %.6f %.6f %.6f  %zu\n", k, msg.seq, msg.stamp_sec, msg.stamp_nsec, cloud.size(),
                s.n_filtered, s.converged ? 1 : 0, s.iterations, s.fitness, s.pose(0, 3), s.pose(1, 3), s.pose(2, 3), s.n_map);
    if (re.is_open()) {  // PointCloud2 round trip: toROSMsg(fromROSMsg(msg)) keeps the header and the points
      pclomp_b200::PointCloud2 out;
      pclomp_b200::toROSMsg(cloud, out);
      out.seq = msg.seq; out.stamp_sec = msg.stamp_sec; out.stamp_nsec = msg.stamp_nsec; out.frame_id = msg.frame_id;
      pclomp_b200::appendToDump(re, out);
    }
    ++k;
  }
  std::printf("replayed %d scans\n", k);
  if (!save_map.empty()) {
    pcl::PointCloud<pcl::PointXYZ> gmap;
    mapper.globalMap(gmap);
    if (pclomp_b200::io::savePCDFileBinary(save_map, gmap) != 0) { std::cerr << "failed to write " << save_map << std::endl; return 1; }
    std::printf("saved %zu map points to %s\n", gmap.size(), save_map.c_str());
  }
  return 0;
}
