// stand-in for <lcm/lcm.h>: only what RBISOpticalFlowMeasurement::publish (out of scope, never called here) names
#pragma once
typedef struct _lcm_t lcm_t;
static inline lcm_t* lcm_create(const char*) { return 0; }
static inline void lcm_destroy(lcm_t*) {}
