// device_hypothesis_mcts.h - DeviceHypothesisMcts: the surface of
//   Mcts<CrossingState<int>, UctStatistic, HypothesisStatistic, RandomHeuristic>
// for hypothesis-MCTS (reference mcts/mcts.h:43-76: search(state, belief_tracker), returnBestAction,
// numIterations, numNodes, searchTime) over the device-resident hypothesis search of the C ABI
// (mamcts_hyp_search_*): num_trees trees live in HBM, whole iterations - the ego's UctStatistic, the other
// agents' HypothesisStatistic, the hypothesis policies, the rollouts - run on the GPU.
//
//   CudaState<CrossingState<int>>::set_parameters(params);           // as for every CUDA host class
//   CudaState<CrossingState<int>>::set_hypotheses({{4,5},{5,6}});    // the ranges given to add_hypothesis()
//   DeviceHypothesisMcts mcts(mcts_parameters, /*num_trees=*/1);
//   mcts.search(*state, belief_tracker);
//   ActionIdx a = mcts.returnBestAction();
//
// The belief tracker is used exactly as Mcts::single_search uses it (mcts.h:158): sample_current_hypothesis()
// is called once per iteration ON THE CALLER'S TRACKER, so it is left in the state the reference's search
// leaves it in; the sampled ids travel to the device as the hypothesis sequence. With num_trees = 1 (or
// decorrelate = false) a tree is the reference's tree bit for bit (tests/dropin/dropin_hypothesis.cc); with
// decorrelate = true the trees are independent root-parallel trees merged like Mcts::parallel_search merges
// them (UctStatistic::merge_node_statistics of the ego roots, mcts.h:270-292).
#ifndef MAMCTS_B200_HOST_DEVICE_HYPOTHESIS_MCTS_H_
#define MAMCTS_B200_HOST_DEVICE_HYPOTHESIS_MCTS_H_

#include <chrono>
#include <cstring>
#include <map>
#include <stdexcept>
#include <utility>
#include <vector>

#include "cuda_random_heuristic.h"  // CudaEngine, to_mamcts_params
#include "cuda_state.h"
#include "mamcts_b200.h"
#include "mcts/hypothesis/hypothesis_belief_tracker.h"
#include "mcts/mcts_parameters.h"

namespace mcts {

inline mamcts_hypothesis_params to_mamcts_hypothesis_params(const MctsParameters& r, unsigned policy_random_seed) {
  mamcts_hypothesis_params h;
  mamcts_default_hypothesis_params(&h);
  h.upper_cost_bound = r.hypothesis_statistic.UPPER_COST_BOUND;
  h.lower_cost_bound = r.hypothesis_statistic.LOWER_COST_BOUND;
  h.pw_k = r.hypothesis_statistic.PROGRESSIVE_WIDENING_K;
  h.pw_alpha = r.hypothesis_statistic.PROGRESSIVE_WIDENING_ALPHA;
  h.exploration_constant = r.hypothesis_statistic.EXPLORATION_CONSTANT;
  h.cost_based_action_selection = r.hypothesis_statistic.COST_BASED_ACTION_SELECTION ? 1 : 0;
  h.progressive_widening_hypothesis_based = r.hypothesis_statistic.PROGRESSIVE_WIDENING_HYPOTHESIS_BASED ? 1 : 0;
  const auto& cc = r.cost_constrained_statistic.USE_CHANCE_CONSTRAINED_UPDATES;
  h.chance_constrained_update = (!cc.empty() && cc.at(0)) ? 1 : 0;   // hypothesis_statistic.h:116
  h.policy_random_seed = policy_random_seed;
  return h;
}

class DeviceHypothesisMcts {
 public:
  using S = CrossingState<int>;
  struct RootStatistic {   // the ego root in UctStatistic's vocabulary (uct_statistic.h:272-289)
    double value = 0.0;
    unsigned total_node_visits = 0;
    std::map<ActionIdx, std::pair<unsigned, double>> ucb_statistics;
  };

  DeviceHypothesisMcts(const MctsParameters& mcts_parameters, unsigned num_trees = 1, bool decorrelate = false,
                       unsigned first_tree_id = 0)
      : mcts_parameters_(mcts_parameters), num_trees_(num_trees), decorrelate_(decorrelate), first_tree_id_(first_tree_id) {
    CudaEngine::get();
  }
  ~DeviceHypothesisMcts() { if (search_) mamcts_search_destroy(search_); }
  DeviceHypothesisMcts(const DeviceHypothesisMcts&) = delete;
  DeviceHypothesisMcts& operator=(const DeviceHypothesisMcts&) = delete;

  // Mcts::search(state, belief_tracker) (mcts.h:143-164), bounded by MAX_NUMBER_OF_ITERATIONS
  void search(const S& current_state, HypothesisBeliefTracker& belief_tracker) {
    using CS = CudaState<S>;
    const auto start = std::chrono::high_resolution_clock::now();
    const mamcts_crossing_params env = CS::params();
    CS::check_state(current_state, env);
    int32_t x[MAMCTS_MAX_AGENTS] = {0}, last[MAMCTS_MAX_AGENTS] = {0};
    uint8_t flags = 0;
    CS::pack(current_state, x, last, &flags);
    if (flags & 1u) throw std::logic_error("DeviceHypothesisMcts::search: the state is terminal");
    const auto& gaps = CS::hypotheses();
    if (gaps.empty() || gaps.size() > MAMCTS_MAX_HYPOTHESES)
      throw std::logic_error("DeviceHypothesisMcts::search: register 1..4 hypotheses with CudaState::set_hypotheses()");
    const std::vector<AgentIdx> others = current_state.get_other_agent_idx();
    const unsigned J = (unsigned)others.size();
    const unsigned iterations = mcts_parameters_.MAX_NUMBER_OF_ITERATIONS;
    // one sample per iteration from the caller's tracker, as single_search draws them (mcts.h:158)
    std::vector<uint8_t> sequence((size_t)iterations * J);
    for (unsigned it = 0; it < iterations; ++it) {
      const auto& cur = belief_tracker.sample_current_hypothesis();
      for (unsigned j = 0; j < J; ++j) {
        const HypothesisId h = cur.at(others[j]);
        if (h >= gaps.size()) throw std::logic_error("DeviceHypothesisMcts::search: sampled hypothesis id not registered");
        sequence[(size_t)it * J + j] = (uint8_t)h;
      }
    }
    mamcts_hyp_search_config cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.num_trees = num_trees_;
    cfg.first_tree_id = first_tree_id_;
    cfg.num_hypotheses = (uint32_t)gaps.size();
    for (size_t h = 0; h < gaps.size(); ++h) { cfg.gap_lo[h] = gaps[h].first; cfg.gap_hi[h] = gaps[h].second; }
    cfg.decorrelate_trees = decorrelate_ ? 1 : 0;
    cfg.mcts = to_mamcts_params(mcts_parameters_);
    cfg.hypothesis = to_mamcts_hypothesis_params(mcts_parameters_, CS::policy_seed());
    cfg.crossing = env;
    if (search_ && std::memcmp(&created_, &cfg, sizeof(cfg)) != 0) {   // parameters / hypotheses changed
      mamcts_search_destroy(search_);
      search_ = nullptr;
    }
    if (!search_) {
      CudaEngine::check(mamcts_hyp_search_create(CudaEngine::get(), &cfg, x, last, sequence.data(), &search_),
                        "mamcts_hyp_search_create");
      created_ = cfg;
    } else {
      CudaEngine::check(mamcts_hyp_search_reset(search_, x, last, sequence.data()), "mamcts_hyp_search_reset");
    }
    CudaEngine::check(mamcts_search_run(search_, iterations), "mamcts_search_run");
    num_ego_actions_ = (unsigned)current_state.get_num_actions(current_state.get_ego_agent_idx());
    CudaEngine::check(mamcts_search_merge_roots(search_, merged_q_, merged_n_, merged_occ_, &merged_visits_, &merged_value_,
                                                &merged_best_), "mamcts_search_merge_roots");
    merged_valid_ = true;
    num_iterations_ = iterations;
    search_time_ = (unsigned)std::chrono::duration_cast<std::chrono::milliseconds>(
                       std::chrono::high_resolution_clock::now() - start).count();
  }

  ActionIdx returnBestAction() const { require(); return merged_best_; }
  unsigned numIterations() const { return num_iterations_; }
  unsigned searchTime() const { return search_time_; }
  unsigned numTrees() const { return num_trees_; }
  uint64_t numNodes() const {
    require();
    uint64_t nodes = 0;
    CudaEngine::check(mamcts_search_counters(search_, nullptr, nullptr, nullptr, &nodes, nullptr), "mamcts_search_counters");
    return nodes;
  }
  uint64_t numEnvSteps() const {
    require();
    uint64_t r = 0, x = 0;
    CudaEngine::check(mamcts_search_counters(search_, nullptr, &r, &x, nullptr, nullptr), "mamcts_search_counters");
    return r + x;
  }
  RootStatistic get_root_statistic(unsigned tree) const {   // the ego root of one tree
    require();
    double value = 0.0, q[MAMCTS_MAX_ACTIONS];
    uint32_t visits = 0, n[MAMCTS_MAX_ACTIONS], mask = 0, nodes = 0;
    CudaEngine::check(mamcts_search_root_stats(search_, tree, 0, &value, &visits, q, n, &mask, &nodes), "mamcts_search_root_stats");
    RootStatistic s;
    s.value = value;
    s.total_node_visits = visits;
    for (unsigned a = 0; a < MAMCTS_MAX_ACTIONS; ++a)
      if ((mask >> a) & 1u) s.ucb_statistics[a] = {n[a], q[a]};
    return s;
  }
  RootStatistic get_merged_root_statistic() const {
    require();
    RootStatistic s;
    s.value = merged_value_;
    s.total_node_visits = (unsigned)merged_visits_;
    for (unsigned a = 0; a < num_ego_actions_; ++a)
      if (merged_occ_[a]) s.ucb_statistics[a] = {(unsigned)merged_n_[a], merged_q_[a]};
    return s;
  }

 private:
  void require() const {
    if (!search_ || !merged_valid_) throw std::logic_error("DeviceHypothesisMcts: search() has not been called");
  }
  const MctsParameters mcts_parameters_;
  unsigned num_trees_;
  bool decorrelate_;
  unsigned first_tree_id_;
  mamcts_search* search_ = nullptr;
  mamcts_hyp_search_config created_{};
  unsigned num_iterations_ = 0, search_time_ = 0, num_ego_actions_ = 0;
  bool merged_valid_ = false;
  double merged_q_[MAMCTS_MAX_ACTIONS] = {0.0}, merged_value_ = 0.0;
  uint64_t merged_n_[MAMCTS_MAX_ACTIONS] = {0}, merged_visits_ = 0;
  uint32_t merged_occ_[MAMCTS_MAX_ACTIONS] = {0}, merged_best_ = 0;
};

}  // namespace mcts
#endif
