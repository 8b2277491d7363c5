// search_more.cu - the device-resident UCT search (search.cu, search_impl.cuh) for CrossingState<int> with 4, 5 or
// 6 other agents (the reference takes the number of agents from CrossingStateParameters::NUM_OTHER_AGENTS,
// environments/crossing_state_parameters.h:15-33; its own tests use 1, 2 and 3). A translation unit of its own so
// that these instantiations compile next to search.cu's instead of after them.
#include "search_impl.cuh"

namespace mamcts {

mamcts_search* make_search_more(const mamcts_search_config& cfg, uint32_t n_ego, uint32_t n_oth, uint32_t max_nodes) {
  const bool compact = cfg.layout != MAMCTS_LAYOUT_CLASSIC;
  const uint32_t it = cfg.mcts.max_number_of_iterations;
  // the counts-only compact layout wherever it applies (search.cu: make_search), else the classic layout
  const bool multi = compact && cfg.env == MAMCTS_ENV_CROSSING_I32 && cfg.mcts.discount_factor >= 0.0 && max_nodes < 32767;
  if (n_ego == 4 && n_oth == 7) {
    switch (cfg.crossing.num_other_agents) {
      case 4: return multi ? make_search_multi<CrossingEnv<4>, 4, 7>(it) : make_search_cidx<CrossingEnv<4>, 4, 7, 0>(max_nodes, false, it);
      case 5: return multi ? make_search_multi<CrossingEnv<5>, 4, 7>(it) : make_search_cidx<CrossingEnv<5>, 4, 7, 0>(max_nodes, false, it);
      case 6: return multi ? make_search_multi<CrossingEnv<6>, 4, 7>(it) : make_search_cidx<CrossingEnv<6>, 4, 7, 0>(max_nodes, false, it);
      default: return nullptr;
    }
  }
  return nullptr;
}

}  // namespace mamcts
