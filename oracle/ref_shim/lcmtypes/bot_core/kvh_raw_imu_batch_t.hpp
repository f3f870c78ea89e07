// stand-in for the lcm-gen C++ bindings of bot_core's kvh_raw_imu_t / kvh_raw_imu_batch_t (libbot2 lcmtypes, absent): the
// fields estimate_tools/src/estimate_tools/imu_stream.cpp reads and writes.  TEST INFRASTRUCTURE ONLY (oracle/_ref).
#pragma once
#include <cstdint>
#include <vector>
namespace bot_core {
struct kvh_raw_imu_t {
  int64_t utime;
  int64_t packet_count;
  double delta_rotation[3];
  double linear_acceleration[3];
};
struct kvh_raw_imu_batch_t {
  int64_t utime;
  int32_t num_packets;
  std::vector<kvh_raw_imu_t> raw_imu;
};
}  // namespace bot_core
