// stand-in for the lcm-gen C binding of pronto-lcmtypes/lcmtypes/pronto_optical_flow_t.lcm (only named by the
// out-of-scope optical-flow update, never called here)
#pragma once
#include <stdint.h>
#include <lcm/lcm.h>
typedef struct _pronto_optical_flow_t {
  int64_t utime;
  double dt, ux, uy, scale, theta;
  double alpha1, alpha2, gamma;
  double conf_rs, conf_xy;
} pronto_optical_flow_t;
static inline int pronto_optical_flow_t_publish(lcm_t*, const char*, const pronto_optical_flow_t*) { return 0; }
