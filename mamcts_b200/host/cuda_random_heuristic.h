// cuda_random_heuristic.h - CudaRandomHeuristic: drop-in for mcts::RandomHeuristic
// (reference mcts/heuristics/random_heuristic.h) as the template parameter H of
// Mcts<S, SE, SO, H> (mcts/mcts.h:35-36). Same surface (Heuristic concept, mcts/heuristic.h:24-41):
//   ctor(const MctsParameters&), calculate_heuristic_values(node), static rollout<S,SE,SO,H>(...),
//   sprintf(). It holds no device logic: every rollout is mamcts_rollout_* through the C ABI.
//
//   Mcts<CrossingState<int>, UctStatistic, UctStatistic, CudaRandomHeuristic> mcts(params);
//
// With the default REFERENCE_CONST action source the result is bit-identical to the stock
// RandomHeuristic; PHILOX gives real uniform rollouts, K per leaf.
// One leaf per call is launch-latency bound by construction (Mcts::iterate is synchronous,
// mcts.h:310); throughput comes from batched_mcts.h or the device-resident search.
#ifndef MAMCTS_B200_HOST_CUDA_RANDOM_HEURISTIC_H_
#define MAMCTS_B200_HOST_CUDA_RANDOM_HEURISTIC_H_

#include <cstdlib>
#include <cstring>
#include <iostream>
#include <mutex>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "cuda_state.h"
#include "mamcts_b200.h"
#include "mcts/heuristic.h"
#include "mcts/mcts_parameters.h"
#include "mcts/stage_node.h"

namespace mcts {

inline mamcts_params to_mamcts_params(const MctsParameters& r) {
  mamcts_params p;
  mamcts_default_params(&p);
  p.discount_factor = r.DISCOUNT_FACTOR;
  p.random_seed = r.RANDOM_SEED;
  p.max_number_of_iterations = r.MAX_NUMBER_OF_ITERATIONS;
  p.max_search_depth = r.MAX_SEARCH_DEPTH;
  p.max_number_of_nodes = r.MAX_NUMBER_OF_NODES;
  p.use_bound_estimation = r.USE_BOUND_ESTIMATION ? 1 : 0;
  p.rh_max_number_of_iterations = r.random_heuristic.MAX_NUMBER_OF_ITERATIONS;
  p.num_parallel_mcts = r.NUM_PARALLEL_MCTS;
  p.uct_lower_bound = r.uct_statistic.LOWER_BOUND;
  p.uct_upper_bound = r.uct_statistic.UPPER_BOUND;
  p.uct_exploration_constant = r.uct_statistic.EXPLORATION_CONSTANT;
  p.uct_pw_k = r.uct_statistic.PROGRESSIVE_WIDENING_K;
  p.uct_pw_alpha = r.uct_statistic.PROGRESSIVE_WIDENING_ALPHA;
  return p;
}

// Process-wide engine (one CUDA device + stream). mamcts_engine serialises calls internally, so
// the Mcts objects of parallel_search's threads (mcts.h:198-210) may share it.
class CudaEngine {
 public:
  static mamcts_engine* get() {
    static CudaEngine instance;
    return instance.engine_;
  }
  // Errors follow the reference's convention (MCTS_EXPECT_TRUE -> std::terminate, common.h:57-64).
  static void check(int rc, const char* what) {
    if (rc != MAMCTS_OK) {
      std::cerr << what << " failed (" << rc << "): " << mamcts_last_error(nullptr) << std::endl;
      std::terminate();
    }
  }

 private:
  CudaEngine() {
    mamcts_config cfg;
    const char* dev = std::getenv("MAMCTS_DEVICE");
    cfg.device = dev ? std::atoi(dev) : 0;
    cfg.flags = 0;
    check(mamcts_create(&cfg, &engine_), "mamcts_create");
  }
  ~CudaEngine() { mamcts_destroy(engine_); }
  mamcts_engine* engine_ = nullptr;
};

class CudaRandomHeuristic : public mcts::Heuristic<CudaRandomHeuristic> {
 public:
  struct Options {
    int action_source = MAMCTS_SRC_REFERENCE_CONST;
    uint32_t rollouts_per_leaf = 1;  // K (PHILOX only)
    uint64_t philox_seed = 1000;
    uint32_t tree_id = 0;            // Philox stream_hi
  };
  static Options& options() {
    static Options o;
    return o;
  }

  explicit CudaRandomHeuristic(const MctsParameters& mcts_parameters)
      : mcts::Heuristic<CudaRandomHeuristic>(mcts_parameters), calls_(0) {
    CudaEngine::get();
  }

  // Same signature and tuple as RandomHeuristic::rollout (random_heuristic.h:27-33): ego return,
  // other returns keyed by agent id, accumulated ego costs, executed step length.
  template <class S, class SE, class SO, class H>
  static std::tuple<Reward, std::unordered_map<AgentIdx, Reward>, EgoCosts, double> rollout(
      const std::shared_ptr<S>& starting_state, const MctsParameters& mcts_parameters,
      unsigned current_depth) {
    static thread_local uint32_t static_calls = 0;
    return rollout_impl<S>(*starting_state, mcts_parameters, current_depth, ++static_calls);
  }

  // random_heuristic.h:106-134
  template <class S, class SE, class SO, class H>
  std::pair<SE, std::unordered_map<AgentIdx, SO>> calculate_heuristic_values(
      const std::shared_ptr<StageNode<S, SE, SO, H>>& node) {
    Reward ego_accum_reward;
    std::unordered_map<AgentIdx, Reward> other_accum_rewards;
    EgoCosts accum_cost;
    double executed_step_length;
    std::tie(ego_accum_reward, other_accum_rewards, accum_cost, executed_step_length) =
        rollout_impl<S>(*node->get_state(), mcts_parameters_, node->get_depth(), ++calls_);
    SE ego_heuristic(0, node->get_state()->get_ego_agent_idx(), mcts_parameters_);
    ego_heuristic.set_heuristic_estimate(ego_accum_reward, accum_cost);
    std::unordered_map<AgentIdx, SO> other_heuristic_estimates;
    for (auto agent_idx : node->get_state()->get_other_agent_idx()) {
      SO statistic(0, agent_idx, mcts_parameters_);
      statistic.set_heuristic_estimate(other_accum_rewards[agent_idx], accum_cost);
      other_heuristic_estimates.insert(std::pair<AgentIdx, SO>(agent_idx, statistic));
    }
    return std::pair<SE, std::unordered_map<AgentIdx, SO>>(ego_heuristic, other_heuristic_estimates);
  }

  std::string sprintf() const { return "CudaRandomHeuristic"; }

 private:
  template <class S>
  static std::tuple<Reward, std::unordered_map<AgentIdx, Reward>, EgoCosts, double> rollout_impl(
      const S& state, const MctsParameters& mcts_parameters, unsigned current_depth,
      uint32_t call_index) {
    using CS = CudaState<S>;
    const Options& opt = options();
    const mamcts_params mp = to_mamcts_params(mcts_parameters);
    const uint32_t A = CS::num_agents(state);
    int32_t x[MAMCTS_MAX_AGENTS] = {0}, last[MAMCTS_MAX_AGENTS] = {0};
    uint8_t flags = 0;
    CS::check_state(state, CS::params());   // throws if set_parameters() was never called or does not fit
    CS::pack(state, x, last, &flags);
    const uint32_t depth = current_depth;

    mamcts_action_source src;
    std::memset(&src, 0, sizeof(src));
    src.kind = opt.action_source;
    src.mem = MAMCTS_MEM_HOST;
    const std::vector<AgentIdx> others = state.get_other_agent_idx();
    if (opt.action_source == MAMCTS_SRC_REFERENCE_CONST) {
      // the constant first draw of a fresh statistic with that many actions
      for (uint32_t a = 0; a < A; ++a) {
        const AgentIdx id = a == 0 ? state.get_ego_agent_idx() : others[a - 1];
        uint32_t order[MAMCTS_MAX_ACTIONS];
        CudaEngine::check(mamcts_reference_expand_order(mcts_parameters.RANDOM_SEED,
                                                        (uint32_t)state.get_num_actions(id), order),
                          "mamcts_reference_expand_order");
        src.const_actions[a] = order[0];
      }
    } else {
      src.philox_seed = opt.philox_seed;
      src.stream_hi_base = opt.tree_id;
      src.stream_lo_base = call_index * opt.rollouts_per_leaf;
    }
    // HYPOTHESIS (hypothesis-MCTS, S derives from HypothesisStateInterface): the other agents follow the
    // hypotheses the belief tracker sampled for this iteration (mcts.h:143-164)
    int32_t gap_lo[MAMCTS_MAX_AGENTS] = {0}, gap_hi[MAMCTS_MAX_AGENTS] = {0};
    if (opt.action_source == MAMCTS_SRC_HYPOTHESIS) {
      pack_hypotheses_if_any<S>(state, gap_lo, gap_hi);
      src.hyp_gap_lo = gap_lo;
      src.hyp_gap_hi = gap_hi;
    }
    const uint32_t K = opt.action_source == MAMCTS_SRC_REFERENCE_CONST ? 1 : opt.rollouts_per_leaf;

    double ego_ret = 0.0, cost0 = 0.0, step_len = 0.0;
    double other_ret[MAMCTS_MAX_AGENTS] = {0.0};
    mamcts_rollout_out out;
    std::memset(&out, 0, sizeof(out));
    out.mem = MAMCTS_MEM_HOST;
    out.ego_return = &ego_ret;
    out.other_return = other_ret;
    out.ego_cost = &cost0;
    out.step_length = &step_len;
    if (CS::kEnv == MAMCTS_ENV_CROSSING_I32) {
      CudaEngine::check(mamcts_rollout_crossing_i32(CudaEngine::get(), &mp, crossing_params<S>(), 1, K,
                                                    MAMCTS_MEM_HOST, x, last, &flags, &depth, &src, &out),
                        "mamcts_rollout_crossing_i32");
    } else {
      CudaEngine::check(mamcts_rollout_simple(CudaEngine::get(), &mp, simple_params<S>(), 1, K,
                                              MAMCTS_MEM_HOST, x, &depth, &src, &out),
                        "mamcts_rollout_simple");
    }
    std::unordered_map<AgentIdx, Reward> other_accum_rewards;
    for (uint32_t j = 0; j < others.size(); ++j) other_accum_rewards[others[j]] = other_ret[j];
    EgoCosts accum_cost(state.get_num_costs(), 0.0);
    if (!accum_cost.empty()) accum_cost[0] = cost0;
    return std::make_tuple(ego_ret, other_accum_rewards, accum_cost, step_len);
  }

  template <class S>
  static auto pack_hypotheses_if_any(const S& state, int32_t* lo, int32_t* hi)
      -> decltype(CudaState<S>::pack_hypotheses(state, lo, hi), void()) {
    CudaState<S>::pack_hypotheses(state, lo, hi);
  }
  template <class S>
  static void pack_hypotheses_if_any(...) {
    std::cerr << "MAMCTS_SRC_HYPOTHESIS needs a hypothesis state (CrossingState)" << std::endl;
    std::terminate();
  }

  template <class S>
  static const mamcts_crossing_params* crossing_params() {
    if constexpr (CudaState<S>::kEnv == MAMCTS_ENV_CROSSING_I32) return &CudaState<S>::params();
    else return nullptr;
  }
  template <class S>
  static const mamcts_simple_params* simple_params() {
    if constexpr (CudaState<S>::kEnv == MAMCTS_ENV_SIMPLE) return &CudaState<S>::params();
    else return nullptr;
  }

  uint32_t calls_;
};

}  // namespace mcts
#endif
