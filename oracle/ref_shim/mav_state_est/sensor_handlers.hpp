// stand-in for MSE/sensor_handlers.hpp as motion_estimate/src/mav_est_legodo/rbis_legodo_common.hpp includes it: that header
// only needs the names lcm::LCM, BotParam, BotTrans and libbot2's small math helpers from it (the handlers themselves are out
// of scope).  TEST INFRASTRUCTURE ONLY.  libbot2 semantics are [RECALLED] (bot_core/rotations.c, small_linalg.h).
#pragma once
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <iostream>
#include <string>

#include <bot_param/param_client.h>

namespace lcm {
class LCM;
}
typedef struct _BotTrans {
  double rot_quat[4];   // w, x, y, z
  double trans_vec[3];
} BotTrans;
static inline double bot_sq(double v) { return v * v; }
// libbot2 bot_quat_to_roll_pitch_yaw (rotations.c): q = (w, x, y, z)
static inline void bot_quat_to_roll_pitch_yaw(const double q[4], double rpy[3]) {
  const double roll_a = 2 * (q[0] * q[1] + q[2] * q[3]);
  const double roll_b = 1 - 2 * (q[1] * q[1] + q[2] * q[2]);
  rpy[0] = atan2(roll_a, roll_b);
  const double pitch_sin = 2 * (q[0] * q[2] - q[3] * q[1]);
  rpy[1] = asin(pitch_sin);
  const double yaw_a = 2 * (q[0] * q[3] + q[1] * q[2]);
  const double yaw_b = 1 - 2 * (q[2] * q[2] + q[3] * q[3]);
  rpy[2] = atan2(yaw_a, yaw_b);
}
