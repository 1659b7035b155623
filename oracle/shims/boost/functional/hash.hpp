// TEST INFRASTRUCTURE ONLY -- stand-in for boost::hash_range used by the reference's
// LRU.h:13 (PATTERN analysis tool, outside the MPC hot path).
#pragma once
#include <cstddef>
#include <functional>
namespace boost {
template <class It>
inline std::size_t hash_range(It first, It last) {
  std::size_t seed = 0;
  for (; first != last; ++first)
    seed ^= std::hash<typename std::iterator_traits<It>::value_type>()(*first) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
  return seed;
}
}  // namespace boost
