// minipcl: ros/ros.h — the two ROS calls of ndt_omp/apps/align.cpp (ros::WallTime::now(), ros::Time::init())
#pragma once
#include <chrono>
namespace ros {
struct WallDuration { double s; double toSec() const { return s; } };
struct WallTime {
  std::chrono::steady_clock::time_point t;
  static WallTime now() { WallTime w; w.t = std::chrono::steady_clock::now(); return w; }
  WallDuration operator-(const WallTime& o) const { return WallDuration{std::chrono::duration<double>(t - o.t).count()}; }
};
struct Time { static void init() {} };
}  // namespace ros
