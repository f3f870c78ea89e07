// stand-in for MIT stl_utils (absent): the one helper MSE/update_history.cpp:44 uses.
// [RECALLED] stlmultimap_get_lower: iterator to the LAST entry whose key is <= the query; false if there is none.
#pragma once
#include <map>
namespace stl_utils {
template <class K, class V>
bool stlmultimap_get_lower(std::multimap<K, V>& m, const K& key, typename std::multimap<K, V>::iterator& lower) {
  typename std::multimap<K, V>::iterator it = m.upper_bound(key);
  if (it == m.begin()) return false;
  --it;
  lower = it;
  return true;
}
}  // namespace stl_utils
