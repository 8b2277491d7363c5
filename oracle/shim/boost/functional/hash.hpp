// Stand-in for <boost/functional/hash.hpp>: the reference only needs boost::hash_range
// (mcts/stage_node.h:16,30) to key its child map. TEST INFRASTRUCTURE ONLY.
// The hash only affects unordered_map iteration order (printing / visitors), not search results.
#ifndef MAMCTS_ORACLE_SHIM_BOOST_HASH_H_
#define MAMCTS_ORACLE_SHIM_BOOST_HASH_H_
#include <cstddef>
#include <functional>
namespace boost {
template <class It>
inline std::size_t hash_range(It first, It last) {
  std::size_t seed = 0;
  for (; first != last; ++first) {
    seed ^= std::hash<typename std::iterator_traits<It>::value_type>()(*first) +
            0x9e3779b97f4a7c15ULL + (seed << 6) + (seed >> 2);
  }
  return seed;
}
}  // namespace boost
#endif
