// stand-in for the lcm_utils pod (absent): noise_id.cpp's loadFilterHistory reads an LCM log with it; no log reader exists
// here, so the loader returns nothing.  TEST INFRASTRUCTURE ONLY (oracle/_ref).
#pragma once
#include <string>
#include <vector>
namespace lcm_utils {
template <class T>
std::vector<T> loadMsgsFromLog(const std::string&, const std::string&) { return std::vector<T>(); }
}  // namespace lcm_utils
