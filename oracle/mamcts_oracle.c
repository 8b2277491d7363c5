/*
 * oracle/mamcts_oracle.c - CPU restatement (plain C) of the MA-MCTS rollout-and-backup hot path.
 * TEST INFRASTRUCTURE ONLY - see mamcts_oracle.h. Compiled with -ffp-contract=off so that no
 * fused multiply-add is formed (the reference is built for baseline x86-64, which has none).
 */
#include "mamcts_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10                                                                              */
/* ------------------------------------------------------------------------------------------ */

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Stream layout. The draw of agent `agent` at rollout step `step` with A agents:
 *   A <= 8: S = 8 / A steps share one block; block = step / S, lane = (step % S) * A + agent;
 *   A  > 8: B = ceil(A / 8) blocks per step; block = step * B + agent / 8, lane = agent % 8.
 * The 16-bit word is lane `lane` of philox(ctr = {block, lo, hi, 0}, key = seed). Lemire's
 * multiply-shift on 16 bits with rejection makes the draw exactly uniform; the r-th retry
 * (r = 0, 1, ...) reads lane (r & 7) of philox(ctr = {step * A + agent, lo, hi, 0x10000 + (r >> 3)}). */
void orc_draw_position(uint32_t A, uint32_t step, uint32_t agent, uint32_t* block, uint32_t* lane) {
  if (A <= 8) {
    const uint32_t S = 8 / A;
    *block = step / S;
    *lane = (step % S) * A + agent;
  } else {
    const uint32_t B = (A + 7) / 8;
    *block = step * B + agent / 8;
    *lane = agent % 8;
  }
}

uint32_t orc_philox_draw(uint64_t seed, uint32_t stream_hi, uint32_t stream_lo, uint32_t A,
                         uint32_t step, uint32_t agent, uint32_t n) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t block, h;
  orc_draw_position(A, step, agent, &block, &h);
  uint32_t ctr[4] = {block, stream_lo, stream_hi, 0u};
  uint32_t w[4];
  orc_philox4x32_10(ctr, key, w);
  uint32_t x = (w[h >> 1] >> (16u * (h & 1u))) & 0xFFFFu;
  uint32_t m = x * n;
  uint32_t l = m & 0xFFFFu;
  if (l < n) {
    uint32_t t = 65536u % n;
    uint32_t r = 0;
    while (l < t) {
      uint32_t rc[4] = {step * A + agent, stream_lo, stream_hi, 0x10000u + (r >> 3)};
      orc_philox4x32_10(rc, key, w);
      uint32_t hh = r & 7u;
      x = (w[hh >> 1] >> (16u * (hh & 1u))) & 0xFFFFu;
      m = x * n;
      l = m & 0xFFFFu;
      ++r;
    }
  }
  return m >> 16;
}

/* ------------------------------------------------------------------------------------------ */
/* Environments                                                                               */
/* ------------------------------------------------------------------------------------------ */

/* environments/crossing_state.h:115-163; CROSSING_POINT at crossing_state_parameters.h:28. */
void orc_crossing_execute(const mamcts_crossing_params* cp, const orc_state* s,
                          const int32_t* action, orc_state* next, double* ego_reward,
                          double* cost0) {
  const int32_t crossing_point = (cp->chain_length - 1) / 2 + 1;
  const int32_t old_x_ego = s->x[0];
  const int32_t v_ego = action[0] + cp->min_velocity_ego; /* idx_to_ego_crossing_action :306-309 */
  const int32_t new_x_ego = old_x_ego + v_ego;            /* ego is NOT clamped :119 */
  const int ego_out_of_map = new_x_ego < 0;               /* :121-123 */
  orc_state out;
  memset(&out, 0, sizeof(out));
  out.x[0] = new_x_ego;
  out.last[0] = v_ego;
  int collision = 0;
  for (uint32_t i = 0; i < cp->num_other_agents; ++i) {
    const int32_t v = action[i + 1];                      /* aconv<int> bit-cast :130-131 */
    const int32_t nx_raw = s->x[i + 1] + v;
    const int32_t nx = nx_raw >= 0 ? nx_raw : 0;          /* clamp others at 0 :131 */
    out.x[i + 1] = nx;
    out.last[i + 1] = v;
    if (new_x_ego >= crossing_point && old_x_ego <= crossing_point && nx >= crossing_point &&
        s->x[i + 1] <= crossing_point) {                  /* :135-139 */
      collision = 1;
    }
  }
  const int goal_reached = (new_x_ego >= cp->ego_goal_pos) && !collision; /* :142 */
  out.terminal = (uint8_t)(goal_reached || collision || ego_out_of_map);  /* :144 */
  out.goal = (uint8_t)goal_reached;
  out.collided = (uint8_t)collision;
  /* :146-148, evaluated left to right in double */
  double r = goal_reached * cp->reward_goal_reached + collision * cp->reward_collision;
  r = r + cp->reward_collision * ego_out_of_map;
  r = r + cp->reward_step;
  *ego_reward = r;
  if (cp->cost_only_collision) {
    *cost0 = (double)(collision * 1.0f);                  /* :149-150 */
  } else {
    *cost0 = -1.0f * r;                                   /* :152 (float * double = double) */
  }
  *next = out;
}

/* test/uct/simple_state.h:26-54 */
void orc_simple_execute(const mamcts_simple_params* sp, const orc_state* s, const int32_t* action,
                        orc_state* next, double* rewards) {
  orc_state out = *s;
  rewards[0] = 0.0;
  rewards[1] = 0.0;
  if (action[0] != action[1]) {                           /* {0,1} or {1,0} :35 */
    int32_t nl = s->x[0] + 1;
    rewards[0] = 1.0;
    rewards[1] = 1.0;
    if (nl >= sp->winning_state_length) {                 /* :43-46 */
      rewards[0] = 5.0;
      rewards[1] = 10.0;
      nl = sp->winning_state_length;
    }
    out.x[0] = nl;
  }
  out.terminal = (uint8_t)(out.x[0] >= sp->winning_state_length ||
                           out.x[0] <= sp->loosing_state_length); /* is_terminal :72-74 */
  *next = out;
}

void orc_env_crossing(orc_env* env, const mamcts_crossing_params* cp) {
  memset(env, 0, sizeof(*env));
  env->env = MAMCTS_ENV_CROSSING_I32;
  env->num_agents = cp->num_other_agents + 1;
  env->num_actions[0] = (uint32_t)(cp->max_velocity_ego - cp->min_velocity_ego + 1);
  for (uint32_t i = 1; i < env->num_agents; ++i) env->num_actions[i] = cp->num_other_actions;
  env->crossing = *cp;
}

void orc_env_simple(orc_env* env, const mamcts_simple_params* sp) {
  memset(env, 0, sizeof(*env));
  env->env = MAMCTS_ENV_SIMPLE;
  env->num_agents = 2;
  env->num_actions[0] = 2;
  env->num_actions[1] = 2;
  env->simple = *sp;
}

static int orc_is_terminal(const orc_env* env, const orc_state* s) {
  if (env->env == MAMCTS_ENV_SIMPLE) {
    return s->x[0] >= env->simple.winning_state_length ||
           s->x[0] <= env->simple.loosing_state_length;
  }
  return s->terminal;
}

/* Executes one joint action given as action INDICES (tree policy / const / philox source):
 * the reference's UctStatistic hands CrossingState<int> the raw index as the other agents'
 * ActionIdx, whose bit-cast is +index. rewards[A], cost0. */
static void orc_execute_idx(const orc_env* env, const orc_state* s, const uint32_t* ja,
                            orc_state* next, double* rewards, double* cost0) {
  int32_t act[ORC_MAX_AGENTS];
  for (uint32_t i = 0; i < env->num_agents; ++i) act[i] = (int32_t)ja[i];
  if (env->env == MAMCTS_ENV_SIMPLE) {
    orc_simple_execute(&env->simple, s, act, next, rewards);
    *cost0 = 0.0;
  } else {
    for (uint32_t i = 0; i < env->num_agents; ++i) rewards[i] = 0.0;
    orc_crossing_execute(&env->crossing, s, act, next, &rewards[0], cost0);
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Rollout                                                                                    */
/* ------------------------------------------------------------------------------------------ */

int32_t orc_crossing_policy_action(const mamcts_crossing_params* cp, int32_t agent_x, int32_t agent_last,
                                   int32_t ego_x, int32_t ego_last, int32_t desired_gap) {
  const int32_t crossing_point = (cp->chain_length - 1) / 2 + 1;  /* crossing_state_parameters.h:28 */
  if (agent_x < crossing_point) {                                 /* agent_policy.h:37 */
    /* forward-predicted ego position from its last action, :39 */
    const int32_t gap_error = ego_x + ego_last - agent_x - desired_gap;
    if (desired_gap > 0) {                                        /* :41-46 */
      if (gap_error < 0) return gap_error > cp->min_velocity_other ? gap_error : cp->min_velocity_other;
      return gap_error < cp->max_velocity_other ? gap_error : cp->max_velocity_other;
    }
    /* already ahead of the ego agent: never brake below the last velocity, :47-50 */
    const int32_t v = gap_error < cp->max_velocity_other ? gap_error : cp->max_velocity_other;
    return v > agent_last ? v : agent_last;
  }
  return agent_last;                                              /* past the crossing point, :51-53 */
}

void orc_rollout(const orc_env* env, const mamcts_params* mp, const orc_state* start,
                 uint32_t start_depth, int32_t kind, const int8_t* replay, uint32_t replay_steps,
                 const uint32_t* const_actions, uint64_t seed, uint32_t stream_hi,
                 uint32_t stream_lo, const int32_t* gap_lo, const int32_t* gap_hi,
                 orc_rollout_result* out, double* reward_trace, uint32_t trace_cap) {
  const uint32_t A = env->num_agents;
  orc_state state = *start;
  double ego = 0.0;                         /* random_heuristic.h:38 */
  double other[ORC_MAX_AGENTS];
  for (uint32_t j = 0; j < ORC_MAX_AGENTS; ++j) other[j] = 0.0; /* :39-42 */
  double cost0 = 0.0;                       /* :44 */
  const double gamma = mp->discount_factor; /* :45 */
  double m = gamma;                         /* :46 */
  uint32_t it = 0;                          /* :47 */
  double len = 0.0;                         /* :48 */
  double prob = 1.0;                        /* :49 */
  uint32_t depth = start_depth;
  const int has_cost = env->env == MAMCTS_ENV_CROSSING_I32; /* get_num_costs :101-103 / simple :60 */
  uint32_t cap = mp->rh_max_number_of_iterations;
  if (kind == MAMCTS_SRC_REPLAY && replay_steps < cap) cap = replay_steps;
  while (!orc_is_terminal(env, &state) && it < cap && depth < mp->max_search_depth) { /* :51-54 */
    int32_t act[ORC_MAX_AGENTS];
    for (uint32_t a = 0; a < A; ++a) {      /* :56-66 */
      if (kind == MAMCTS_SRC_REPLAY) {
        act[a] = replay[(size_t)it * A + a];
      } else if (kind == MAMCTS_SRC_REFERENCE_CONST) {
        act[a] = (int32_t)const_actions[a];
      } else if (kind == MAMCTS_SRC_HYPOTHESIS && a > 0) {
        /* HypothesisStatistic::choose_next_action of a fresh statistic samples from the hypothesis
         * (hypothesis_statistic.h:70-81) -> CrossingState::plan_action_current_hypothesis
         * (crossing_state.h:78-82) -> AgentPolicyCrossingState<int>::act (agent_policy.h:68-75) */
        const uint32_t span = (uint32_t)(gap_hi[a - 1] - gap_lo[a - 1] + 1);
        const int32_t gap = gap_lo[a - 1] + (int32_t)orc_philox_draw(seed, stream_hi, stream_lo, A, it, a, span);
        act[a] = orc_crossing_policy_action(&env->crossing, state.x[a], state.last[a], state.x[0],
                                            state.last[0], gap);
      } else {
        act[a] = (int32_t)orc_philox_draw(seed, stream_hi, stream_lo, A, it, a,
                                          env->num_actions[a]);
      }
    }
    orc_state next;
    double rewards[ORC_MAX_AGENTS];
    double c0 = 0.0;
    if (env->env == MAMCTS_ENV_SIMPLE) {
      orc_simple_execute(&env->simple, &state, act, &next, rewards);
    } else {
      for (uint32_t a = 0; a < A; ++a) rewards[a] = 0.0;
      orc_crossing_execute(&env->crossing, &state, act, &next, &rewards[0], &c0);
    }
    if (reward_trace && it < trace_cap) reward_trace[it] = rewards[0];
    m = gamma * m;                          /* :71 */
    ego += m * rewards[0];                  /* :72 */
    for (uint32_t j = 1; j < A; ++j) other[j - 1] = m * rewards[j]; /* ASSIGN :73-77 */
    if (has_cost) cost0 += c0;              /* undiscounted :79-84 */
    m = m * gamma;                          /* :85 */
    len += 1.0;                             /* get_execution_step_length() == 1.0 :87 */
    prob *= 1.0 / (double)env->num_actions[0]; /* :88 */
    it += 1;                                /* :89 */
    depth += 1;                             /* :90 */
    state = next;                           /* :91 */
  }
  ego *= prob;                              /* :95 */
  if (has_cost) cost0 *= prob;              /* :96-101 */
  out->ego_return = ego;
  for (uint32_t j = 0; j < ORC_MAX_AGENTS; ++j) out->other_return[j] = other[j];
  out->cost0 = cost0;
  out->step_length = len;
  out->steps = it;
  out->final_state = state;
  out->final_state.terminal = (uint8_t)orc_is_terminal(env, &state);
}

/* mcts/heuristics/action_value_random_heuristic.h:27-112 (see the header for the arithmetic). */
void orc_action_values(const orc_env* env, const mamcts_params* mp, const orc_state* start, uint32_t start_depth,
                       int32_t kind, const uint32_t* const_actions, uint64_t seed, uint32_t stream_hi,
                       uint32_t stream_lo, orc_action_values_result* out) {
  const uint32_t A = env->num_agents;
  const uint32_t nA = env->num_actions[0];
  const double gamma = mp->discount_factor;
  memset(out, 0, sizeof(*out));
  if (orc_is_terminal(env, start)) return;                       /* :38 - no action is tried */
  for (uint32_t a = 0; a < nA; ++a) {
    int32_t act[ORC_MAX_AGENTS];
    const uint32_t lo = stream_lo * nA + a;
    act[0] = (int32_t)a;                                         /* :41 */
    for (uint32_t j = 1; j < A; ++j)                             /* :43-47: a fresh statistic's choice */
      act[j] = kind == MAMCTS_SRC_REFERENCE_CONST
                   ? (int32_t)const_actions[j]
                   : (int32_t)orc_philox_draw(seed, stream_hi ^ 0x80000000u, lo, A, 0, j, env->num_actions[j]);
    orc_state next;
    double r0 = 0.0, c0 = 0.0;
    orc_crossing_execute(&env->crossing, start, act, &next, &r0, &c0);   /* :52 */
    orc_rollout_result rr;
    orc_rollout(env, mp, &next, start_depth + 1, kind, NULL, 0, const_actions, seed, stream_hi, lo, NULL, NULL, &rr,
                NULL, 0);                                        /* :58-61 */
    out->action_return[a] = 0.0 + (gamma * r0 + gamma * rr.ego_return);  /* :64-65 */
    for (uint32_t j = 0; j + 1 < A; ++j)                         /* :66-71 on the tied map: rewards[1..] are 0 */
      out->other_return[j] = rr.other_return[j] + (gamma * 0.0 + gamma * 0.0);
    out->action_cost[a] = 0.0 + (c0 + rr.cost0);                 /* :73-81 */
    out->action_len[a] = 0.0 + (1.0 + rr.step_length);           /* :83-84 */
    out->steps += 1 + rr.steps;
  }
  double mean = 0.0;                                             /* :97-103, unordered_map order: last first */
  for (uint32_t a = nA; a-- > 0;) mean += out->action_cost[a];
  out->mean_cost = mean / (double)nA;
}

/* ------------------------------------------------------------------------------------------ */
/* CrossingState<float>                                                                       */
/* ------------------------------------------------------------------------------------------ */

void orc_crossing_f32_execute(const mamcts_crossing_params_f32* cp, const orc_state_f32* s, const float* action,
                              orc_state_f32* next, double* ego_reward, double* cost0) {
  const uint32_t A = cp->num_other_agents + 1;
  const float crossing_point = (cp->chain_length - 1.0f) / 2.0f + 1.0f; /* CROSSING_POINT(), parameters.h:28 */
  const float old_x_ego = s->x[0];
  const float v_ego = action[0] + cp->min_velocity_ego;                 /* :306-309 (ActionIdx -> float) */
  const float new_x_ego = s->x[0] + v_ego;                              /* :119 */
  const int out_of_map = new_x_ego < 0.0f;                              /* :121-123 */
  int collision = 0;
  memset(next, 0, sizeof(*next));
  next->x[0] = new_x_ego;
  next->last[0] = v_ego;
  for (uint32_t i = 1; i < A; ++i) {
    const float new_x = s->x[i] + action[i];                            /* :130 */
    next->x[i] = (new_x >= 0.0f) ? new_x : 0.0f;                        /* :131 */
    next->last[i] = action[i];
    if (new_x_ego >= crossing_point && old_x_ego <= crossing_point && next->x[i] >= crossing_point &&
        s->x[i] <= crossing_point)
      collision = 1;                                                    /* :135-139 */
  }
  const int goal = (new_x_ego >= cp->ego_goal_pos) && !collision;       /* :142 */
  next->terminal = (uint8_t)(goal || collision || out_of_map);          /* :144 */
  next->goal = (uint8_t)goal;
  next->collided = (uint8_t)collision;
  double r = (double)goal * cp->reward_goal_reached + (double)collision * cp->reward_collision; /* :146-148 */
  r = r + cp->reward_collision * (double)out_of_map;
  r = r + cp->reward_step;
  *ego_reward = r;
  *cost0 = cp->cost_only_collision ? (double)((float)collision * 1.0f) : (double)-1.0f * r; /* :149-153 */
}

void orc_rollout_f32(const mamcts_crossing_params_f32* cp, const mamcts_params* mp, const orc_state_f32* start,
                     uint32_t start_depth, int32_t kind, const float* replay, uint32_t replay_steps,
                     const uint32_t* const_actions, uint64_t seed, uint32_t stream_hi, uint32_t stream_lo,
                     orc_rollout_result_f32* out) {
  const uint32_t A = cp->num_other_agents + 1;
  const uint32_t n_ego = (uint32_t)(cp->max_velocity_ego - cp->min_velocity_ego + 1.0f); /* NUM_EGO_ACTIONS() */
  orc_state_f32 state = *start;
  double ego = 0.0, cost0 = 0.0, len = 0.0, prob = 1.0;
  double other[ORC_MAX_AGENTS];
  for (uint32_t j = 0; j < ORC_MAX_AGENTS; ++j) other[j] = 0.0;
  const double gamma = mp->discount_factor;
  double m = gamma;
  uint32_t it = 0, depth = start_depth;
  uint32_t cap = mp->rh_max_number_of_iterations;
  if (kind == MAMCTS_SRC_REPLAY && replay_steps < cap) cap = replay_steps;
  while (!state.terminal && it < cap && depth < mp->max_search_depth) {
    float act[ORC_MAX_AGENTS];
    for (uint32_t a = 0; a < A; ++a) {
      if (kind == MAMCTS_SRC_REPLAY) {
        act[a] = replay[(size_t)it * A + a];
      } else {
        const uint32_t idx = kind == MAMCTS_SRC_REFERENCE_CONST
                                 ? const_actions[a]
                                 : orc_philox_draw(seed, stream_hi, stream_lo, A, it, a, a == 0 ? n_ego : cp->num_other_actions);
        if (a == 0) {
          act[a] = (float)idx;
        } else {
          memcpy(&act[a], &idx, sizeof(float)); /* aconv<float>(ActionIdx): the low bytes reinterpreted */
        }
      }
    }
    orc_state_f32 next;
    double r, c0;
    orc_crossing_f32_execute(cp, &state, act, &next, &r, &c0);
    m = gamma * m;
    ego += m * r;
    for (uint32_t j = 1; j < A; ++j) other[j - 1] = m * 0.0;
    cost0 += c0;
    m = m * gamma;
    len += 1.0;
    prob *= 1.0 / (double)n_ego;
    it += 1;
    depth += 1;
    state = next;
  }
  out->ego_return = ego * prob;
  for (uint32_t j = 0; j < ORC_MAX_AGENTS; ++j) out->other_return[j] = other[j];
  out->cost0 = cost0 * prob;
  out->step_length = len;
  out->steps = it;
  out->final_state = state;
}

/* environments/crossing_state_agent_policy.h:35-55, Domain = float */
float orc_crossing_policy_action_f32(const mamcts_crossing_params_f32* cp, float agent_x, float agent_last, float ego_x,
                                     float ego_last, float desired_gap) {
  volatile float chain_m1 = cp->chain_length - 1.0f;      /* CROSSING_POINT() in float, crossing_state_parameters.h:28 */
  volatile float half = chain_m1 / 2.0f;
  const float crossing_point = half + 1.0f;
  if (agent_x < crossing_point) {                          /* :37 */
    volatile float t0 = ego_x + ego_last;                  /* :39, left to right in float */
    volatile float t1 = t0 - agent_x;
    const float gap_error = t1 - desired_gap;
    if (desired_gap > 0) {                                 /* :41-46 */
      /* std::max(a, b) = a < b ? b : a; std::min(a, b) = b < a ? b : a (the signs of zeros follow from this) */
      if (gap_error < 0) return gap_error < cp->min_velocity_other ? cp->min_velocity_other : gap_error;
      return cp->max_velocity_other < gap_error ? cp->max_velocity_other : gap_error;
    }
    {                                                      /* :47-50 */
      const float lim = cp->max_velocity_other < gap_error ? cp->max_velocity_other : gap_error;
      return lim < agent_last ? agent_last : lim;
    }
  }
  return agent_last;                                       /* :51-53 */
}

void orc_rollout_f32_hypothesis(const mamcts_crossing_params_f32* cp, const mamcts_params* mp, const orc_state_f32* start,
                                uint32_t start_depth, uint64_t seed, uint32_t stream_hi, uint32_t stream_lo,
                                const float* gap_lo, const float* gap_hi, orc_rollout_result_f32* out) {
  const uint32_t A = cp->num_other_agents + 1;
  const uint32_t n_ego = (uint32_t)(cp->max_velocity_ego - cp->min_velocity_ego + 1.0f);
  orc_state_f32 state = *start;
  double ego = 0.0, cost0 = 0.0, len = 0.0, prob = 1.0;
  double other[ORC_MAX_AGENTS];
  for (uint32_t j = 0; j < ORC_MAX_AGENTS; ++j) other[j] = 0.0;
  const double gamma = mp->discount_factor;
  double m = gamma;
  uint32_t it = 0, depth = start_depth;
  const uint32_t cap = mp->rh_max_number_of_iterations;
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  while (!state.terminal && it < cap && depth < mp->max_search_depth) {
    float act[ORC_MAX_AGENTS];
    act[0] = (float)orc_philox_draw(seed, stream_hi, stream_lo, A, it, 0, n_ego);
    for (uint32_t a = 1; a < A; ++a) {
      uint32_t block, h, w[4];
      orc_draw_position(A, it, a, &block, &h);
      const uint32_t ctr[4] = {block, stream_lo, stream_hi, 0u};
      orc_philox4x32_10(ctr, key, w);
      const uint32_t word = (w[h >> 1] >> (16u * (h & 1u))) & 0xFFFFu;
      volatile float span = gap_hi[a - 1] - gap_lo[a - 1];
      volatile float frac = (float)word * (1.0f / 65536.0f);
      volatile float scaled = span * frac;
      const float gap = gap_lo[a - 1] + scaled;
      act[a] = orc_crossing_policy_action_f32(cp, state.x[a], state.last[a], state.x[0], state.last[0], gap);
    }
    orc_state_f32 next;
    double r, c0;
    orc_crossing_f32_execute(cp, &state, act, &next, &r, &c0);
    m = gamma * m;
    ego += m * r;
    for (uint32_t j = 1; j < A; ++j) other[j - 1] = m * 0.0;
    cost0 += c0;
    m = m * gamma;
    len += 1.0;
    prob *= 1.0 / (double)n_ego;
    it += 1;
    depth += 1;
    state = next;
  }
  out->ego_return = ego * prob;
  for (uint32_t j = 0; j < ORC_MAX_AGENTS; ++j) out->other_return[j] = other[j];
  out->cost0 = cost0 * prob;
  out->step_length = len;
  out->steps = it;
  out->final_state = state;
}

static double orc_cost_mean(const double* c, uint32_t C) {
  double acc = 0.0; /* std::accumulate(begin, end, 0.0) / size(), hypothesis_statistic.h:120,159 */
  for (uint32_t i = 0; i < C; ++i) acc += c[i];
  return acc / (double)C;
}

void orc_backup_hypothesis(double gamma, uint32_t num_others, uint32_t n_paths, const uint32_t* leaf_node,
                           const uint8_t* leaf_has_heuristic, const double* leaf_cost, uint32_t num_costs,
                           const uint32_t* path_offset, const uint32_t* path_len, const uint32_t* edge_node,
                           const uint8_t* edge_hypothesis, const uint8_t* edge_action, const double* edge_cost,
                           int32_t chance_constrained_update, mamcts_hyp_stats* st) {
  const size_t J = num_others, H = st->num_hypotheses, MA = st->max_actions, C = num_costs;
  for (uint32_t p = 0; p < n_paths; ++p) {
    for (size_t j = 0; j < J; ++j) {
      const size_t leaf = (size_t)leaf_node[p] * J + j;
      double latest_child;
      if (leaf_has_heuristic[p]) {
        /* set_heuristic_estimate :158-161 then update_from_heuristic :100-110 of a statistic whose
         * hypothesis_id_current_iteration_ is still HYPOTHESIS_ID_NOT_SET */
        const double h = orc_cost_mean(leaf_cost + (size_t)p * C, (uint32_t)C);
        st->ego_cost_value[leaf] = h;
        st->latest_ego_cost[leaf] = h;
        st->visits_hypothesis[leaf * (H + 1) + H] += 1;
        st->total_visits[leaf] += 1;
        latest_child = h;
      } else {
        latest_child = st->latest_ego_cost[leaf];
      }
      for (uint32_t e = 0; e < path_len[p]; ++e) {
        const size_t idx = (size_t)path_offset[p] + e;
        const size_t slot = (size_t)edge_node[idx] * J + j;
        const size_t hyp = edge_hypothesis[idx * J + j], act = edge_action[idx * J + j];
        /* update_statistic :112-134 */
        const double cw = orc_cost_mean(edge_cost + idx * C, (uint32_t)C);
        double latest;
        if (chance_constrained_update) latest = cw < latest_child ? latest_child : cw; /* std::max */
        else latest = cw + gamma * latest_child;
        st->latest_ego_cost[slot] = latest;
        const size_t ia = (slot * H + hyp) * MA + act;
        st->action_count[ia] += 1;
        st->action_cost[ia] = st->action_cost[ia] + (latest - st->action_cost[ia]) / (double)st->action_count[ia];
        const size_t ih = slot * (H + 1) + hyp;
        st->visits_hypothesis[ih] += 1;
        st->ego_cost_value[slot] =
            st->ego_cost_value[slot] + (latest - st->ego_cost_value[slot]) / (double)st->visits_hypothesis[ih];
        st->total_visits[slot] += 1;
        st->num_expanded[slot] += 1;
        latest_child = latest;
      }
    }
  }
}

double orc_reduce_mean(const double* vals, uint32_t K) {
  if (K == 1) return vals[0];
  double v[32], nv[32];
  if (K < 32) {
    for (uint32_t l = 0; l < K; ++l) v[l] = vals[l];
    for (uint32_t off = K / 2; off >= 1; off /= 2) {
      for (uint32_t l = 0; l < K; ++l) nv[l] = v[l] + v[l ^ off];
      for (uint32_t l = 0; l < K; ++l) v[l] = nv[l];
    }
    return v[0] / (double)K;
  }
  uint32_t R = K / 32 < 8 ? K / 32 : 8;
  uint32_t chunk = 32 * R;
  uint32_t nchunks = K / chunk;
  double total = 0.0;
  for (uint32_t c = 0; c < nchunks; ++c) {
    for (uint32_t l = 0; l < 32; ++l) {
      double acc = 0.0;
      for (uint32_t j = 0; j < R; ++j) acc += vals[(size_t)c * chunk + j * 32 + l];
      v[l] = acc;
    }
    for (uint32_t off = 16; off >= 1; off /= 2) {
      for (uint32_t l = 0; l < 32; ++l) nv[l] = v[l] + v[l ^ off];
      for (uint32_t l = 0; l < 32; ++l) v[l] = nv[l];
    }
    total += v[0];
  }
  return total / (double)K;
}

/* ------------------------------------------------------------------------------------------ */
/* UctStatistic                                                                               */
/* ------------------------------------------------------------------------------------------ */

void orc_stat_init(orc_stat* s, uint32_t num_actions) {
  memset(s, 0, sizeof(*s));
  s->num_actions = num_actions;
}

void orc_stat_update_from_heuristic(orc_stat* s, double h) { /* uct_statistic.h:106-110 */
  s->value = h;
  s->latest_return = s->value;
  s->total_visits += 1;
}

void orc_stat_backprop(orc_stat* s, double gamma, double child_latest_return) {
  const uint32_t a = s->collected_action;                 /* uct_statistic.h:136 */
  if (!s->expanded[a]) {                                  /* operator[] default-inserts */
    s->expanded[a] = 1;
    s->num_expanded += 1;
    s->n[a] = 0;
    s->q[a] = 0.0;
  }
  s->latest_return = s->collected_reward + gamma * child_latest_return; /* :138 */
  s->n[a] += 1;                                                         /* :139 */
  s->q[a] = s->q[a] + (s->latest_return - s->q[a]) / s->n[a];           /* :140 */
  s->total_visits += 1;                                                 /* :142 */
  s->value = s->value + (s->latest_return - s->value) / s->total_visits;/* :143 */
}

/* uct_statistic.h:122-132 */
static void orc_bounds_from_stat(const orc_stat* s, double* lo, double* hi) {
  int first = 1;
  double mn = 0.0, mx = 0.0;
  for (uint32_t a = 0; a < s->num_actions; ++a) {
    if (!s->expanded[a]) continue;
    if (first) { mn = mx = s->q[a]; first = 0; continue; }
    if (s->q[a] < mn) mn = s->q[a];
    if (!(s->q[a] < mx)) mx = s->q[a];
  }
  if (mn == mx) {
    if (mx != 0.0) { *lo = 0.0; *hi = mx; }
    else { *lo = 0.0; *hi = 1.0; }
    return;
  }
  *lo = mn; *hi = mx;
}

uint32_t orc_stat_choose_next_action(orc_stat* s, const mamcts_params* mp,
                                     const uint32_t* expand_order) {
  /* require_progressive_widening_total, uct_statistic.h:256-262 */
  const double widening_term = mp->uct_pw_k * pow((double)s->total_visits, mp->uct_pw_alpha);
  if ((double)s->num_expanded <= widening_term && s->num_expanded < s->num_actions) {
    const uint32_t a = expand_order[s->num_expanded];     /* :64-68 (deterministic, F2) */
    s->expanded[a] = 1;
    s->num_expanded += 1;
    s->n[a] = 0;
    s->q[a] = 0.0;
    return a;
  }
  /* calculate_ucb_and_max_action, :223-244 */
  uint32_t best = 0;
  double max_value = DBL_MIN;
  double lo = mp->uct_lower_bound, hi = mp->uct_upper_bound;
  if (mp->use_bound_estimation) orc_bounds_from_stat(s, &lo, &hi);
  for (uint32_t a = 0; a < s->num_actions; ++a) {
    if (!s->expanded[a]) continue;
    const double norm = (s->q[a] - lo) / (hi - lo);
    const double v =
        norm + 2 * mp->uct_exploration_constant *
                   sqrt((2 * log((double)s->total_visits)) / (double)s->n[a]);
    if (v > max_value) { max_value = v; best = a; }
  }
  return best;
}

uint32_t orc_stat_best_action(const orc_stat* s) {         /* uct_statistic.h:78-89 */
  int first = 1;
  double temp = 0.0;
  uint32_t best = 0;
  for (uint32_t a = 0; a < s->num_actions; ++a) {
    if (!s->expanded[a]) continue;
    if (first) { temp = s->q[a]; best = a; first = 0; continue; }
    if (s->q[a] > temp) { temp = s->q[a]; best = a; }
  }
  return best;
}

void orc_stat_merge(const orc_stat* const* stats, uint32_t n, orc_stat* merged) {
  /* uct_statistic.h:165-184; `merged` keeps its own total_visits like a fresh root's (0). */
  uint32_t occ[ORC_MAX_ACTIONS];
  memset(occ, 0, sizeof(occ));
  for (uint32_t a = 0; a < ORC_MAX_ACTIONS; ++a) {
    merged->expanded[a] = 0; merged->q[a] = 0.0; merged->n[a] = 0;
  }
  merged->num_expanded = 0;
  for (uint32_t t = 0; t < n; ++t) {
    const orc_stat* st = stats[t];
    for (uint32_t a = 0; a < st->num_actions; ++a) {
      if (!st->expanded[a]) continue;
      if (!merged->expanded[a]) { merged->expanded[a] = 1; merged->num_expanded += 1; }
      merged->q[a] += st->q[a];
      merged->n[a] += st->n[a];
      occ[a] += 1;
    }
    merged->total_visits += st->total_visits;
  }
  merged->value = 0.0;
  for (uint32_t a = 0; a < ORC_MAX_ACTIONS; ++a) {
    if (!merged->expanded[a]) continue;
    merged->q[a] /= occ[a];
    merged->value += merged->q[a];
  }
  merged->value /= merged->num_expanded;
}

/* ------------------------------------------------------------------------------------------ */
/* Arena tree search                                                                          */
/* ------------------------------------------------------------------------------------------ */

typedef struct orc_child {
  uint32_t ja[ORC_MAX_AGENTS];
  uint32_t node;
  double rewards[ORC_MAX_AGENTS]; /* joint_rewards_[ja], stage_node.h:62 */
} orc_child;

typedef struct orc_node {
  int32_t parent;
  uint32_t depth;
  uint32_t visit_count;
  uint32_t ja[ORC_MAX_AGENTS]; /* joint action leading here */
  orc_state state;
  orc_stat* stats;             /* [A] */
  orc_child* children;
  uint32_t num_children, cap_children;
} orc_node;

struct orc_tree {
  orc_search_config cfg;
  orc_node* nodes;
  uint32_t num_nodes, cap_nodes;
  uint64_t rollout_steps;
  uint64_t expand_steps;
  uint32_t iterations;
};

static uint32_t orc_tree_new_node(orc_tree* t, int32_t parent, uint32_t depth, const uint32_t* ja,
                                  const orc_state* st) {
  if (t->num_nodes == t->cap_nodes) {
    t->cap_nodes = t->cap_nodes ? t->cap_nodes * 2 : 1024;
    t->nodes = (orc_node*)realloc(t->nodes, (size_t)t->cap_nodes * sizeof(orc_node));
  }
  orc_node* nd = &t->nodes[t->num_nodes];
  memset(nd, 0, sizeof(*nd));
  nd->parent = parent;
  nd->depth = depth;
  if (ja) memcpy(nd->ja, ja, sizeof(nd->ja));
  nd->state = *st;
  nd->stats = (orc_stat*)calloc(t->cfg.env.num_agents, sizeof(orc_stat));
  for (uint32_t a = 0; a < t->cfg.env.num_agents; ++a)
    orc_stat_init(&nd->stats[a], t->cfg.env.num_actions[a]);
  return t->num_nodes++;
}

orc_tree* orc_tree_create(const orc_search_config* cfg, const orc_state* root) {
  orc_tree* t = (orc_tree*)calloc(1, sizeof(orc_tree));
  t->cfg = *cfg;
  orc_state r = *root;
  r.terminal = (uint8_t)orc_is_terminal(&cfg->env, root);
  orc_tree_new_node(t, -1, 0, NULL, &r);
  return t;
}

void orc_tree_destroy(orc_tree* t) {
  if (!t) return;
  for (uint32_t i = 0; i < t->num_nodes; ++i) {
    free(t->nodes[i].stats);
    free(t->nodes[i].children);
  }
  free(t->nodes);
  free(t);
}

static void orc_tree_iterate(orc_tree* t) {
  const orc_env* env = &t->cfg.env;
  const mamcts_params* mp = &t->cfg.mcts;
  const uint32_t A = env->num_agents;
  uint32_t node = 0;
  int run_heuristic = 0;
  /* select & expand, mcts.h:300-305 + stage_node.h:167-232 */
  for (;;) {
    orc_node* nd = &t->nodes[node];
    nd->visit_count++;                                               /* :180 */
    if (orc_is_terminal(env, &nd->state) || nd->depth == mp->max_search_depth ||
        t->num_nodes > mp->max_number_of_nodes) {                    /* :183-187 */
      run_heuristic = 0;
      break;
    }
    uint32_t ja[ORC_MAX_AGENTS];
    memset(ja, 0, sizeof(ja));
    for (uint32_t a = 0; a < A; ++a)                                 /* :190-195 */
      ja[a] = orc_stat_choose_next_action(&nd->stats[a], mp, t->cfg.expand_order[a]);
    int found = -1;
    for (uint32_t c = 0; c < nd->num_children; ++c) {                /* :198 */
      if (memcmp(nd->children[c].ja, ja, sizeof(uint32_t) * A) == 0) { found = (int)c; break; }
    }
    if (found >= 0) {                                                /* :199-206 */
      for (uint32_t a = 0; a < A; ++a) {
        nd->stats[a].collected_action = ja[a];
        nd->stats[a].collected_reward = nd->children[found].rewards[a];
      }
      node = nd->children[found].node;
      continue;
    }
    /* expand, :207-231 */
    orc_state next;
    double rewards[ORC_MAX_AGENTS];
    double cost0;
    orc_execute_idx(env, &nd->state, ja, &next, rewards, &cost0);
    t->expand_steps++;
    next.terminal = (uint8_t)orc_is_terminal(env, &next);
    const uint32_t parent_depth = nd->depth;
    const uint32_t child = orc_tree_new_node(t, (int32_t)node, parent_depth + 1, ja, &next);
    nd = &t->nodes[node]; /* realloc may have moved the arena */
    if (nd->num_children == nd->cap_children) {
      nd->cap_children = nd->cap_children ? nd->cap_children * 2 : 4;
      nd->children = (orc_child*)realloc(nd->children, nd->cap_children * sizeof(orc_child));
    }
    orc_child* ch = &nd->children[nd->num_children++];
    memset(ch, 0, sizeof(*ch));
    memcpy(ch->ja, ja, sizeof(ch->ja));
    ch->node = child;
    for (uint32_t a = 0; a < A; ++a) {
      ch->rewards[a] = rewards[a];
      nd->stats[a].collected_action = ja[a];
      nd->stats[a].collected_reward = rewards[a];
    }
    node = child;
    run_heuristic = 1;
    break;
  }
  /* heuristic, mcts.h:309-312 + random_heuristic.h:106-134 + stage_node.h:249-256 */
  if (run_heuristic) {
    orc_node* nd = &t->nodes[node];
    const uint32_t K = t->cfg.rollouts_per_leaf;
    double* ego = (double*)malloc(sizeof(double) * K);
    double* oth = (double*)malloc(sizeof(double) * K * ORC_MAX_AGENTS);
    for (uint32_t k = 0; k < K; ++k) {
      orc_rollout_result rr;
      uint32_t const_actions[ORC_MAX_AGENTS];
      for (uint32_t a = 0; a < A; ++a) const_actions[a] = t->cfg.expand_order[a][0];
      orc_rollout(env, mp, &nd->state, nd->depth, t->cfg.action_source, NULL, 0, const_actions,
                  t->cfg.philox_seed, t->cfg.tree_id, node * K + k, NULL, NULL, &rr, NULL, 0);
      t->rollout_steps += rr.steps;
      ego[k] = rr.ego_return;
      for (uint32_t j = 0; j + 1 < A; ++j) oth[(size_t)j * K + k] = rr.other_return[j];
    }
    orc_stat_update_from_heuristic(&nd->stats[0], orc_reduce_mean(ego, K));
    for (uint32_t j = 0; j + 1 < A; ++j)
      orc_stat_update_from_heuristic(&nd->stats[j + 1], orc_reduce_mean(oth + (size_t)j * K, K));
    free(ego);
    free(oth);
  }
  /* backpropagation, mcts.h:314-329 + stage_node.h:259-271 */
  int32_t child = (int32_t)node;
  int32_t p = t->nodes[node].parent;
  while (p >= 0) {
    for (uint32_t a = 0; a < A; ++a)
      orc_stat_backprop(&t->nodes[p].stats[a], mp->discount_factor,
                        t->nodes[child].stats[a].latest_return);
    child = p;
    p = t->nodes[p].parent;
  }
  t->iterations++;
}

void orc_tree_run(orc_tree* t, uint32_t iterations) {
  for (uint32_t i = 0; i < iterations; ++i) orc_tree_iterate(t);
}

uint32_t orc_tree_num_nodes(const orc_tree* t) { return t->num_nodes; }
uint64_t orc_tree_rollout_steps(const orc_tree* t) { return t->rollout_steps; }
const orc_stat* orc_tree_root_stat(const orc_tree* t, uint32_t agent) {
  return &t->nodes[0].stats[agent];
}

static double orc_ja_code(const orc_tree* t, const uint32_t* ja, uint32_t max_actions) {
  double code = 0.0, mul = 1.0;
  for (uint32_t a = 0; a < t->cfg.env.num_agents; ++a) {
    code += mul * (double)ja[a];
    mul *= (double)max_actions;
  }
  return code;
}

static uint32_t orc_dump_rec(const orc_tree* t, uint32_t node, double* rec, uint32_t cap,
                             uint32_t count, uint32_t max_actions) {
  const uint32_t A = t->cfg.env.num_agents;
  const uint32_t stride = 4 + A * (3 + 2 * max_actions);
  const orc_node* nd = &t->nodes[node];
  if (count < cap) {
    double* r = rec + (size_t)count * stride;
    r[0] = nd->depth;
    r[1] = nd->parent < 0 ? -1.0 : orc_ja_code(t, nd->ja, max_actions);
    r[2] = nd->visit_count;
    r[3] = nd->num_children;
    for (uint32_t a = 0; a < A; ++a) {
      double* s = r + 4 + a * (3 + 2 * max_actions);
      const orc_stat* st = &nd->stats[a];
      s[0] = st->total_visits;
      s[1] = st->value;
      s[2] = st->latest_return;
      for (uint32_t k = 0; k < max_actions; ++k) {
        const int ex = k < st->num_actions && st->expanded[k];
        s[3 + k] = ex ? (double)st->n[k] : -1.0;
        s[3 + max_actions + k] = ex ? st->q[k] : 0.0;
      }
    }
  }
  count++;
  /* children in ascending joint-action code */
  uint32_t nc = nd->num_children;
  uint32_t* order = (uint32_t*)malloc(sizeof(uint32_t) * (nc ? nc : 1));
  for (uint32_t i = 0; i < nc; ++i) order[i] = i;
  for (uint32_t i = 1; i < nc; ++i) {
    uint32_t k = order[i];
    double ck = orc_ja_code(t, nd->children[k].ja, max_actions);
    uint32_t j = i;
    while (j > 0 && orc_ja_code(t, nd->children[order[j - 1]].ja, max_actions) > ck) {
      order[j] = order[j - 1];
      --j;
    }
    order[j] = k;
  }
  for (uint32_t i = 0; i < nc; ++i)
    count = orc_dump_rec(t, nd->children[order[i]].node, rec, cap, count, max_actions);
  free(order);
  return count;
}

uint32_t orc_tree_dump(const orc_tree* t, double* records, uint32_t capacity_records,
                       uint32_t max_actions) {
  return orc_dump_rec(t, 0, records, capacity_records, 0, max_actions);
}
