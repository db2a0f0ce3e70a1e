// Test stand-in for std_msgs/Header (ROS 1): the members the shim touches, same names and types.
#pragma once
#include <cstdint>
#include <string>
namespace ros { struct Time { uint32_t sec = 0, nsec = 0; }; }
namespace std_msgs { struct Header { uint32_t seq = 0; ros::Time stamp; std::string frame_id; }; }
