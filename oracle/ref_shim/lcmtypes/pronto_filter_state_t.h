// stand-in for the lcm-gen C binding of pronto-lcmtypes/lcmtypes/pronto_filter_state_t.lcm (field order of the .lcm file)
#pragma once
#include <stdint.h>
typedef struct _pronto_filter_state_t {
  int64_t utime;
  double quat[4];
  int32_t num_states;
  double* state;
  int32_t num_cov_elements;
  double* cov;
} pronto_filter_state_t;
