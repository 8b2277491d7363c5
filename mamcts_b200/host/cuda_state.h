// cuda_state.h - CudaState<S>: how a reference State is packed into the engine's positional POD
// layout (include/mamcts_b200.h). Compiled against the reference's headers (juloberno/mamcts):
// add this repository's mamcts_b200/host and include/ to the include path of the host program.
//
//   State concept      : mcts/state.h:99-136
//   CrossingState<int> : environments/crossing_state.h (accessors :261-271, :173-185)
//   SimpleState        : test/uct/simple_state.h
//
// Agents are positional on the device (ego = 0, others in get_other_agent_idx() order); agent ids
// (4 / 5 in SimpleState) stay a host concept (SURVEY.md §7 "agent identity").
#ifndef MAMCTS_B200_HOST_CUDA_STATE_H_
#define MAMCTS_B200_HOST_CUDA_STATE_H_

#include <cstdint>
#include <stdexcept>
#include <utility>
#include <vector>

#include "mamcts_b200.h"
#include "mcts/state.h"

namespace mcts {

template <class S, class Enable = void>
struct CudaState;  // specialise per environment

// ---- CrossingState<int> -----------------------------------------------------------------------
#ifdef CROSSING_STATE_H
template <>
struct CudaState<CrossingState<int>> {
  static constexpr int kEnv = MAMCTS_ENV_CROSSING_I32;
  using EnvParams = mamcts_crossing_params;
  // CrossingState keeps its parameters private (crossing_state.h:320), and the reference's Mcts<> constructs
  // its heuristic from MctsParameters alone (mcts.h:43-48), so the host registers the same
  // CrossingStateParameters<int> object it built its states from. There is no silent default: params()
  // throws until set_parameters() was called, and pack() checks everything a state lets it check (number
  // of other agents, action counts) against the registered parameters. The drivers that own their
  // configuration (DeviceRootParallelMcts, BatchedRootParallelMcts) also take the parameters in their
  // constructor and re-create the device search when they change.
  static bool& registered() { static bool r = false; return r; }
  static mamcts_crossing_params& storage() {
    static mamcts_crossing_params p = [] { mamcts_crossing_params d; mamcts_default_crossing_params(&d); return d; }();
    return p;
  }
  static mamcts_crossing_params& params() {
    if (!registered())
      throw std::logic_error("CudaState<CrossingState<int>>: call set_parameters(CrossingStateParameters<int>) with the "
                             "parameters the states were built from before using the CUDA engine");
    return storage();
  }
  static mamcts_crossing_params to_params(const CrossingStateParameters<int>& c) {
    mamcts_crossing_params p;
    mamcts_default_crossing_params(&p);
    p.num_other_agents = c.NUM_OTHER_AGENTS;
    p.num_other_actions = c.NUM_OTHER_ACTIONS;
    p.cost_only_collision = c.COST_ONLY_COLLISION ? 1 : 0;
    p.min_velocity_ego = c.MIN_VELOCITY_EGO;
    p.max_velocity_ego = c.MAX_VELOCITY_EGO;
    p.min_velocity_other = c.MIN_VELOCITY_OTHER;
    p.max_velocity_other = c.MAX_VELOCITY_OTHER;
    p.chain_length = c.CHAIN_LENGTH;
    p.ego_goal_pos = c.EGO_GOAL_POS;
    p.reward_collision = c.REWARD_COLLISION;
    p.reward_goal_reached = c.REWARD_GOAL_REACHED;
    p.reward_step = c.REWARD_STEP;
    return p;
  }
  static void set_parameters(const CrossingStateParameters<int>& c) {
    storage() = to_params(c);
    policy_seed() = c.OTHER_AGENTS_POLICY_RANDOM_SEED;
    registered() = true;
  }
  // the seed every AgentPolicyCrossingState generator starts from (crossing_state_agent_policy.h:25)
  static unsigned& policy_seed() { static unsigned s = 1000; return s; }
  // what a state reveals about its parameters must agree with `p`
  static void check_state(const CrossingState<int>& s, const mamcts_crossing_params& p) {
    if (s.get_agent_states().size() != p.num_other_agents)
      throw std::logic_error("CudaState<CrossingState<int>>: the registered parameters do not match the state (other agents)");
    if ((int64_t)s.get_num_actions(s.get_ego_agent_idx()) != (int64_t)p.max_velocity_ego - p.min_velocity_ego + 1)
      throw std::logic_error("CudaState<CrossingState<int>>: the registered parameters do not match the state (ego actions)");
    const std::vector<AgentIdx> others = s.get_other_agent_idx();
    if (!others.empty() && s.get_num_actions(others[0]) != p.num_other_actions)
      throw std::logic_error("CudaState<CrossingState<int>>: the registered parameters do not match the state (other actions)");
  }
  // Behaviour hypotheses (MAMCTS_SRC_HYPOTHESIS): CrossingState keeps its AgentPolicyCrossingState
  // objects private (crossing_state.h:313) and a policy hides its desired-gap range
  // (crossing_state_agent_policy.h:61), so the host registers the same ranges it passed to
  // add_hypothesis(), in the same order (index = HypothesisId).
  static std::vector<std::pair<int32_t, int32_t>>& hypotheses() {
    static std::vector<std::pair<int32_t, int32_t>> h;
    return h;
  }
  static void set_hypotheses(const std::vector<std::pair<int32_t, int32_t>>& gap_ranges) { hypotheses() = gap_ranges; }
  // desired-gap range of the hypothesis currently assigned to every other agent (positional)
  static void pack_hypotheses(const CrossingState<int>& s, int32_t* gap_lo, int32_t* gap_hi) {
    const std::vector<AgentIdx> others = s.get_other_agent_idx();
    for (size_t j = 0; j < others.size(); ++j) {
      const HypothesisId h = s.get_current_hypothesis(others[j]);
      if (h >= hypotheses().size())
        throw std::logic_error("CudaState<CrossingState<int>>: set_hypotheses() does not cover the state's hypotheses");
      gap_lo[j] = hypotheses()[h].first;
      gap_hi[j] = hypotheses()[h].second;
    }
  }
  static uint32_t num_agents(const CrossingState<int>& s) { return s.get_num_agents(); }
  // x / last: [A] positional; flags bit0 = terminal
  static void pack(const CrossingState<int>& s, int32_t* x, int32_t* last, uint8_t* flags) {
    x[0] = s.get_ego_state().x_pos;
    last[0] = s.get_ego_state().last_action;
    const auto& others = s.get_agent_states();
    for (size_t i = 0; i < others.size(); ++i) {
      x[i + 1] = others[i].x_pos;
      last[i + 1] = others[i].last_action;
    }
    *flags = s.is_terminal() ? 1 : 0;
  }
};
#endif

// ---- SimpleState --------------------------------------------------------------------------------
#ifdef SIMPLESTATE_H
template <>
struct CudaState<SimpleState> {
  static constexpr int kEnv = MAMCTS_ENV_SIMPLE;
  using EnvParams = mamcts_simple_params;
  static void check_state(const SimpleState&, const mamcts_simple_params&) {}
  static mamcts_simple_params& params() {
    static mamcts_simple_params p = [] { mamcts_simple_params d; mamcts_default_simple_params(&d); return d; }();
    return p;
  }
  static uint32_t num_agents(const SimpleState&) { return 2; }
  static void pack(const SimpleState& s, int32_t* x, int32_t* last, uint8_t* flags) {
    x[0] = s.get_state_length();
    last[0] = 0;
    *flags = s.is_terminal() ? 1 : 0;
  }
};
#endif

}  // namespace mcts
#endif
