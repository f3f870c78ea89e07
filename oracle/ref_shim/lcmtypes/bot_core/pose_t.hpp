// struct-only stand-in for the lcm-gen output of bot_core_pose_t.lcm (field order and sizes as in the .lcm file)
#pragma once
#include <stdint.h>
namespace bot_core {
struct pose_t {
  int64_t utime;
  double pos[3], vel[3], orientation[4], rotation_rate[3], accel[3];
};
}  // namespace bot_core
