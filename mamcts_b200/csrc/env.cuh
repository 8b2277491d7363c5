// env.cuh - device restatement of the two environments' execute(), one lane per episode, state in
// registers. All floating point is IEEE binary64 in the reference's operation order; the library
// is compiled with -fmad=false so no multiply-add is contracted.
#pragma once
#include <stdint.h>

#include "mamcts_b200.h"

namespace mamcts {

// Everything execute() reads that is constant over a launch.
struct EnvParams {
  // CrossingState<int> (reference environments/crossing_state_parameters.h:15-33)
  double r_goal, r_coll, r_step;
  int32_t min_v_ego;
  int32_t crossing_point;   // CROSSING_POINT() = (CHAIN_LENGTH-1)/2+1, :28
  int32_t goal_pos;
  int32_t cost_only_collision;
  int32_t min_v_other, max_v_other;   // MIN/MAX_VELOCITY_OTHER (behaviour-hypothesis policy)
  uint32_t n_ego_actions;   // NUM_EGO_ACTIONS(), :25
  uint32_t n_other_actions;
  // SimpleState (reference test/uct/simple_state.h:16)
  int32_t win_len, lose_len;
  // CrossingState<float> (Domain-typed parameters, crossing_state_parameters.h:15-28)
  float f_min_v_ego, f_crossing_point, f_goal_pos;
  float f_min_v_other, f_max_v_other;   // MIN/MAX_VELOCITY_OTHER of CrossingState<float> (behaviour-hypothesis policy)
  // CrossingState: the ego reward and cost of a step only depend on (goal, collision, out_of_map);
  // the 8 possible values are evaluated once on the host in the reference's operation order
  // (crossing_state.h:146-153) and looked up with index goal | collision << 1 | out_of_map << 2.
  double r_tab[8], c_tab[8];
};

// CrossingStateParameters<float>: the Domain-typed fields are floats and CROSSING_POINT() is evaluated
// in float ((CHAIN_LENGTH - 1) / 2 + 1, crossing_state_parameters.h:28); rewards stay doubles.
inline void make_crossing_env_f32(const mamcts_crossing_params_f32& cp, EnvParams& ep);

// Fills every CrossingState field of EnvParams from the public parameter struct.
inline void make_crossing_env(const mamcts_crossing_params& cp, EnvParams& ep) {
  ep.r_goal = cp.reward_goal_reached;
  ep.r_coll = cp.reward_collision;
  ep.r_step = cp.reward_step;
  ep.min_v_ego = cp.min_velocity_ego;
  ep.crossing_point = (cp.chain_length - 1) / 2 + 1;   // CROSSING_POINT(), crossing_state_parameters.h:28
  ep.goal_pos = cp.ego_goal_pos;
  ep.cost_only_collision = cp.cost_only_collision;
  ep.min_v_other = cp.min_velocity_other;
  ep.max_v_other = cp.max_velocity_other;
  ep.n_ego_actions = (uint32_t)(cp.max_velocity_ego - cp.min_velocity_ego + 1);
  ep.n_other_actions = cp.num_other_actions;
  for (int f = 0; f < 8; ++f) {
    const bool goal = f & 1, coll = f & 2, oob = f & 4;
    // :146-148 evaluated left to right in double; volatile keeps every product and sum individually
    // rounded whatever the host compiler's contraction setting
    volatile double t0 = (double)(int)goal * ep.r_goal;
    volatile double t1 = (double)(int)coll * ep.r_coll;
    volatile double r = t0 + t1;
    volatile double t2 = ep.r_coll * (double)(int)oob;
    r = r + t2;
    r = r + ep.r_step;
    ep.r_tab[f] = r;
    ep.c_tab[f] = cp.cost_only_collision ? (double)(int)coll : -r;   // :149-153
  }
}

inline void make_crossing_env_f32(const mamcts_crossing_params_f32& cp, EnvParams& ep) {
  mamcts_crossing_params ip;   // the Domain-independent part (rewards, cost mode, counts) through the int path
  ip.reward_collision = cp.reward_collision;
  ip.reward_goal_reached = cp.reward_goal_reached;
  ip.reward_step = cp.reward_step;
  ip.num_other_agents = cp.num_other_agents;
  ip.num_other_actions = cp.num_other_actions;
  ip.cost_only_collision = cp.cost_only_collision;
  ip.min_velocity_ego = 0; ip.max_velocity_ego = 0; ip.min_velocity_other = 0; ip.max_velocity_other = 0;
  ip.chain_length = 3; ip.ego_goal_pos = 0; ip.reserved = 0;
  make_crossing_env(ip, ep);
  // the reference's cost for Domain = float is built from float literals: {collision * 1.0f, 0.0f} or
  // {-1.0f * rewards[0], 0.0f} (crossing_state.h:149-153) - the products are doubles either way
  volatile float chain_m1 = cp.chain_length - 1.0f;
  volatile float half = chain_m1 / 2.0f;
  ep.f_crossing_point = half + 1.0f;
  ep.f_goal_pos = cp.ego_goal_pos;
  ep.f_min_v_ego = cp.min_velocity_ego;
  ep.f_min_v_other = cp.min_velocity_other;
  ep.f_max_v_other = cp.max_velocity_other;
  // NUM_EGO_ACTIONS() is a float for this Domain and get_num_actions converts it to ActionIdx (:165-171)
  volatile float n_ego = cp.max_velocity_ego - cp.min_velocity_ego + 1.0f;
  ep.n_ego_actions = (uint32_t)n_ego;
  ep.n_other_actions = cp.num_other_actions;
}

// CrossingState<int> with NO other agents. Positions / last actions live in registers.
template <int NO>
struct CrossingEnv {
  using Domain = int32_t;
  // a joint-action entry given as an action INDEX, as the statistics of the reference produce it: the
  // ego's stays an index, an other agent's is bit-cast to the Domain (crossing_state_common.h:19-27, F5)
  __device__ __forceinline__ static int32_t action_from_index(int, uint32_t idx) { return (int32_t)idx; }
  static constexpr int A = NO + 1;
  static constexpr int kStateInts = A;     // x_pos per agent
  static constexpr bool kHasCost = true;   // get_num_costs() == 1, crossing_state.h:101-103
  // rewards[1..] are always 0 (:145): with a non-negative discount every return, V and Q of the other
  // agents is exactly +0.0 for the whole search
  static constexpr bool kOtherRewardZero = true;
  struct State {
    int32_t x[A];
    int32_t last[A];
    uint32_t flags;  // bit0 terminal, bit1 goal, bit2 collided
  };

  __device__ __forceinline__ static uint32_t num_actions(const EnvParams& p, int agent) {
    return agent == 0 ? p.n_ego_actions : p.n_other_actions;
  }
  __device__ __forceinline__ static bool terminal(const EnvParams&, const State& s) { return s.flags & 1u; }

  // act[0] = ego action index; act[i>0] = the Domain value of other i (for index-valued sources the
  // reference's bit-cast of the raw index, i.e. +index).
  // Restates CrossingState<int>::execute, reference environments/crossing_state.h:115-163.
  __device__ __forceinline__ static void step(const EnvParams& p, State& s, const int32_t (&act)[A],
                                              double& r_ego, double& r_other, double& cost) {
    const uint32_t f = step_outcome(p, s, act);
    // :146-153: one of 8 host-evaluated values (see EnvParams::r_tab)
    r_ego = p.r_tab[f];
    r_other = 0.0;                                         // rewards[1..] stay 0 (:145 resize)
    cost = p.c_tab[f];
  }

  // The reward and the cost of a step are functions of its outcome code f = goal | collision << 1 |
  // out of bounds << 2 alone (r_tab / c_tab), and f != 0 exactly when the step ends the episode: a
  // caller that only needs the rewards of an episode can look them up afterwards (lane_rollout).
  static constexpr bool kOutcomeRewards = true;
  __device__ __forceinline__ static double outcome_reward(const EnvParams& p, uint32_t f) { return p.r_tab[f]; }
  __device__ __forceinline__ static double outcome_cost(const EnvParams& p, uint32_t f) { return p.c_tab[f]; }
  __device__ __forceinline__ static uint32_t step_outcome(const EnvParams& p, State& s, const int32_t (&act)[A]) {
    const int32_t old_x = s.x[0];
    const int32_t v_ego = act[0] + p.min_v_ego;            // :306-309
    const int32_t new_x = old_x + v_ego;                   // :119 (ego not clamped)
    const bool oob = new_x < 0;                            // :121-123
    const bool ego_encloses = (new_x >= p.crossing_point) & (old_x <= p.crossing_point);
    bool coll = false;
#pragma unroll
    for (int i = 1; i < A; ++i) {
      const int32_t ox = s.x[i];
      const int32_t nx = max(ox + act[i], 0);              // :130-131
      coll |= ego_encloses & (nx >= p.crossing_point) & (ox <= p.crossing_point);  // :135-139
      s.x[i] = nx;
      s.last[i] = act[i];
    }
    s.x[0] = new_x;
    s.last[0] = v_ego;
    const bool goal = (new_x >= p.goal_pos) & !coll;       // :142
    s.flags = (uint32_t)(goal | coll | oob) | ((uint32_t)goal << 1) | ((uint32_t)coll << 2);  // :144
    return (uint32_t)goal | ((uint32_t)coll << 1) | ((uint32_t)oob << 2);
  }
};

// CrossingState<float> with NO other agents (reference environments/crossing_state.h with Domain = float;
// 30 other actions by default, crossing_state_parameters.h:50-51). Positions and velocities are FP32 and
// every operation is the reference's single IEEE operation (no contraction: -fmad=false), so REPLAY is
// bit-exact. Index-valued other actions are bit-cast like the reference does: index i is the denormal
// float with bit pattern i (F5), which the adds preserve (no flush-to-zero in this build).
template <int NO>
struct CrossingEnvF {
  using Domain = float;
  static constexpr int A = NO + 1;
  static constexpr int kStateInts = A;
  static constexpr bool kHasCost = true;
  static constexpr bool kOutcomeRewards = false;
  struct State {
    float x[A];
    float last[A];
    uint32_t flags;  // bit0 terminal, bit1 goal, bit2 collided
  };
  __device__ __forceinline__ static float action_from_index(int agent, uint32_t idx) {
    return agent == 0 ? (float)idx : __uint_as_float(idx);   // ego: ActionIdx -> Domain conversion (:306-309)
  }
  __device__ __forceinline__ static uint32_t num_actions(const EnvParams& p, int agent) {
    return agent == 0 ? p.n_ego_actions : p.n_other_actions;
  }
  __device__ __forceinline__ static bool terminal(const EnvParams&, const State& s) { return s.flags & 1u; }
  // act[0] = the ego action index as a float; act[i>0] = the velocity of other i.
  __device__ __forceinline__ static void step(const EnvParams& p, State& s, const float (&act)[A],
                                              double& r_ego, double& r_other, double& cost) {
    const float old_x = s.x[0];
    const float v_ego = __fadd_rn(act[0], p.f_min_v_ego);       // :306-309
    const float new_x = __fadd_rn(old_x, v_ego);                // :119
    const bool oob = new_x < 0.0f;                              // :121-123
    const bool ego_encloses = (new_x >= p.f_crossing_point) & (old_x <= p.f_crossing_point);
    bool coll = false;
#pragma unroll
    for (int i = 1; i < A; ++i) {
      const float ox = s.x[i];
      const float moved = __fadd_rn(ox, act[i]);                // :130
      const float nx = (moved >= 0.0f) ? moved : 0.0f;          // :131
      coll |= ego_encloses & (nx >= p.f_crossing_point) & (ox <= p.f_crossing_point);  // :135-139
      s.x[i] = nx;
      s.last[i] = act[i];
    }
    s.x[0] = new_x;
    s.last[0] = v_ego;
    const bool goal = (new_x >= p.f_goal_pos) & !coll;          // :142
    s.flags = (uint32_t)(goal | coll | oob) | ((uint32_t)goal << 1) | ((uint32_t)coll << 2);
    const uint32_t f = (uint32_t)goal | ((uint32_t)coll << 1) | ((uint32_t)oob << 2);
    r_ego = p.r_tab[f];
    r_other = 0.0;
    cost = p.c_tab[f];
  }
};

// AgentPolicyCrossingState<int>::calculate_action (reference
// environments/crossing_state_agent_policy.h:35-55): the velocity an other agent chooses for a desired
// gap to the ego agent, whose position is predicted one step ahead from its last action.
__device__ __forceinline__ int32_t crossing_policy_action(const EnvParams& p, int32_t agent_x, int32_t agent_last,
                                                          int32_t ego_x, int32_t ego_last, int32_t desired_gap) {
  if (agent_x >= p.crossing_point) return agent_last;                  // :51-53
  const int32_t gap_error = ego_x + ego_last - agent_x - desired_gap;  // :39
  if (desired_gap > 0)                                                 // :41-46
    return gap_error < 0 ? max(gap_error, p.min_v_other) : min(gap_error, p.max_v_other);
  return max(min(gap_error, p.max_v_other), agent_last);               // :47-50
}

// AgentPolicyCrossingState<float>::calculate_action (the same lines with Domain = float): every operation one
// FP32 operation in the reference's order; std::max(a, b) = a < b ? b : a and std::min(a, b) = b < a ? b : a,
// which also fixes the signs of zeros.
__device__ __forceinline__ float crossing_policy_action_f32(const EnvParams& p, float agent_x, float agent_last,
                                                            float ego_x, float ego_last, float desired_gap) {
  if (!(agent_x < p.f_crossing_point)) return agent_last;                                   // :51-53
  const float gap_error = __fsub_rn(__fsub_rn(__fadd_rn(ego_x, ego_last), agent_x), desired_gap);   // :39
  if (desired_gap > 0.0f)                                                                  // :41-46
    return gap_error < 0.0f ? (gap_error < p.f_min_v_other ? p.f_min_v_other : gap_error)
                            : (p.f_max_v_other < gap_error ? p.f_max_v_other : gap_error);
  const float lim = p.f_max_v_other < gap_error ? p.f_max_v_other : gap_error;             // :47-50
  return lim < agent_last ? agent_last : lim;
}

// SimpleState: one int32 of state, 2 agents x 2 actions (reference test/uct/simple_state.h).
struct SimpleEnv {
  using Domain = int32_t;
  __device__ __forceinline__ static int32_t action_from_index(int, uint32_t idx) { return (int32_t)idx; }
  static constexpr int A = 2;
  static constexpr int kStateInts = 1;     // state_length_
  static constexpr bool kHasCost = false;  // get_num_costs() == 0, :60-62
  static constexpr bool kOtherRewardZero = false;
  static constexpr bool kOutcomeRewards = false;
  struct State {
    int32_t x[A];     // x[0] = state_length_, x[1] unused (kept for a uniform interface)
    int32_t last[A];  // unused
    uint32_t flags;   // bit0 terminal
  };
  __device__ __forceinline__ static uint32_t num_actions(const EnvParams&, int) { return 2; }
  __device__ __forceinline__ static bool terminal(const EnvParams& p, const State& s) {
    return (s.x[0] >= p.win_len) | (s.x[0] <= p.lose_len);  // is_terminal :72-74
  }
  // Restates SimpleState::execute, reference test/uct/simple_state.h:26-54.
  __device__ __forceinline__ static void step(const EnvParams& p, State& s, const int32_t (&act)[A],
                                              double& r_ego, double& r_other, double& cost) {
    r_ego = 0.0;
    r_other = 0.0;
    cost = 0.0;
    if (act[0] != act[1]) {                                // :35
      int32_t nl = s.x[0] + 1;
      r_ego = 1.0;
      r_other = 1.0;
      if (nl >= p.win_len) {                               // :43-46
        r_ego = 5.0;
        r_other = 10.0;
        nl = p.win_len;
      }
      s.x[0] = nl;
    }
    s.last[0] = act[0];
    s.last[1] = act[1];
    s.flags = (uint32_t)terminal(p, s);
  }
};

}  // namespace mamcts
