// stand-in for the lcm-gen C++ binding of pronto-lcmtypes/lcmtypes/pronto_filter_state_t.lcm
#pragma once
#include <stdint.h>
#include <vector>
namespace pronto {
class filter_state_t {
 public:
  int64_t utime;
  double quat[4];
  int32_t num_states;
  std::vector<double> state;
  int32_t num_cov_elements;
  std::vector<double> cov;
};
}  // namespace pronto
