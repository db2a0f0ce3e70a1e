// Test stand-in for sensor_msgs/PointCloud2 (ROS 1 message definition): same member names and types. ROS is not installed
// in the build image; the shim's adapters are compiled against this to prove they use the message the way roscpp lays it out.
#pragma once
#include <cstdint>
#include <vector>
#include "../std_msgs/Header.h"
#include "PointField.h"
namespace sensor_msgs {
struct PointCloud2 {
  std_msgs::Header header;
  uint32_t height = 0, width = 0;
  std::vector<PointField> fields;
  uint8_t is_bigendian = 0;
  uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  uint8_t is_dense = 0;
};
}  // namespace sensor_msgs
