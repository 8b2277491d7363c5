/*
 * oracle/mamcts_oracle.h - CPU restatement (plain C) of the MA-MCTS rollout-and-backup hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under mamcts_b200/ may include, link or call this; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and only as the checker. Each function cites the reference (juloberno/mamcts) file:line it
 * restates. Parity status: PINNED - validated against the unmodified reference compiled into
 * oracle/_ref/libmamcts_ref.so (tests/test_oracle_cpu.py::test_oracle_*_equal_reference*) and against the golden
 * vectors in tests/golden/ generated from it (tests/golden/make_golden.py).
 */
#ifndef MAMCTS_ORACLE_H_
#define MAMCTS_ORACLE_H_

#include <stdint.h>
#include "mamcts_b200.h" /* parameter structs only (the public C ABI header) */

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_AGENTS MAMCTS_MAX_AGENTS
#define ORC_MAX_ACTIONS MAMCTS_MAX_ACTIONS

/* ---- Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3",
 * SC'11; the Random123 reference constants). The PHILOX action source is NOT in the reference
 * (its rollouts are deterministic); the stream layout is specified in DESIGN.md §5. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* Where the draw of (step, agent) with A agents sits in its stream: block index and 16-bit lane. */
void orc_draw_position(uint32_t A, uint32_t step, uint32_t agent, uint32_t* block, uint32_t* lane);
/* Exactly-uniform draw in [0, n) for (step, agent) of stream (stream_hi, stream_lo). */
uint32_t orc_philox_draw(uint64_t seed, uint32_t stream_hi, uint32_t stream_lo, uint32_t A,
                         uint32_t step, uint32_t agent, uint32_t n);

/* ---- environments ------------------------------------------------------------------------ */

typedef struct orc_state {
  int32_t x[ORC_MAX_AGENTS];    /* Crossing: x_pos per agent (ego first); Simple: x[0] = length */
  int32_t last[ORC_MAX_AGENTS]; /* Crossing: last_action */
  uint8_t terminal, goal, collided, pad;
} orc_state;

/* CrossingState<int>::execute (environments/crossing_state.h:115-163).
 * action[0] = ego action index; action[1..] = the int the reference bit-casts out of the
 * other agent's ActionIdx (crossing_state_common.h:19-27). Writes the ego reward and cost[0]. */
void orc_crossing_execute(const mamcts_crossing_params* cp, const orc_state* s,
                          const int32_t* action, orc_state* next, double* ego_reward,
                          double* cost0);
/* SimpleState::execute (test/uct/simple_state.h:26-54). rewards[2]. */
void orc_simple_execute(const mamcts_simple_params* sp, const orc_state* s, const int32_t* action,
                        orc_state* next, double* rewards);

/* ---- rollout: RandomHeuristic::rollout (mcts/heuristics/random_heuristic.h:27-104) -------- */

typedef struct orc_env {
  int32_t env;                       /* MAMCTS_ENV_* */
  uint32_t num_agents;               /* A */
  uint32_t num_actions[ORC_MAX_AGENTS];
  mamcts_crossing_params crossing;
  mamcts_simple_params simple;
} orc_env;

void orc_env_crossing(orc_env* env, const mamcts_crossing_params* cp);
void orc_env_simple(orc_env* env, const mamcts_simple_params* sp);

typedef struct orc_rollout_result {
  double ego_return;
  double other_return[ORC_MAX_AGENTS];   /* index j = other j (position j+1) */
  double cost0;
  double step_length;
  uint32_t steps;
  orc_state final_state;
} orc_rollout_result;

/* AgentPolicyCrossingState<int>::calculate_action (environments/crossing_state_agent_policy.h:35-55):
 * the velocity an other agent chooses for a desired gap to the (forward-predicted) ego agent. */
int32_t orc_crossing_policy_action(const mamcts_crossing_params* cp, int32_t agent_x, int32_t agent_last,
                                   int32_t ego_x, int32_t ego_last, int32_t desired_gap);

/* One rollout. Action source as in mamcts_action_source but for ONE rollout:
 *   kind REPLAY: replay[step*A + agent] (int8), replay_steps steps available;
 *   kind REFERENCE_CONST: const_actions[agent]; kind PHILOX: (seed, stream_hi, stream_lo);
 *   kind HYPOTHESIS: other agent j (position j+1) draws a gap uniformly in [gap_lo[j], gap_hi[j]] from
 *   its Philox position and acts by orc_crossing_policy_action; the ego draws like PHILOX.
 * reward_trace (optional) receives the ego step rewards, up to trace_cap entries. */
void orc_rollout(const orc_env* env, const mamcts_params* mp, const orc_state* start,
                 uint32_t start_depth, int32_t kind, const int8_t* replay, uint32_t replay_steps,
                 const uint32_t* const_actions, uint64_t seed, uint32_t stream_hi,
                 uint32_t stream_lo, const int32_t* gap_lo, const int32_t* gap_hi,
                 orc_rollout_result* out, double* reward_trace, uint32_t trace_cap);

/* ---- per-ego-action heuristic: ActionValueRandomHeuristic::calculate_heuristic_values
 * (mcts/heuristics/action_value_random_heuristic.h:27-112), the part before the statistics are filled:
 * for every ego action a one first step (the other agents act as a fresh statistic makes them act) and one
 * RandomHeuristic::rollout from the resulting state at depth + 1 (:37-61), then
 *   action_return[a] = gamma * r_ego + gamma * rollout return      (:64-65)
 *   action_cost[a]   = first-step cost[0] + rollout cost[0]         (:73-81, not discounted)
 *   action_len[a]    = step length + rollout length                (:83-84)
 *   other_return[j]  = what the LAST action's rollout returned for other j, plus gamma * its first-step reward
 *                      (:58-60 ties the rollout's map to the accumulator itself, so earlier actions are overwritten)
 *   mean_cost        = the action costs added up in the iteration order of the reference's
 *                      std::unordered_map<ActionIdx, EgoCosts> (libstdc++: last inserted first) / nA_ego (:97-103)
 * A terminal start state runs no action at all (:38): everything is 0.
 * Action sources: REFERENCE_CONST (const_actions[agent], the first-draw constants) and PHILOX: the first
 * step's other-agent draws are (step 0, agent j) of stream (stream_hi ^ 0x80000000, stream_lo * nA_ego + a), the
 * rollout of action a is stream (stream_hi, stream_lo * nA_ego + a). CrossingState<int> only. */
typedef struct orc_action_values_result {
  double action_return[ORC_MAX_ACTIONS];
  double action_cost[ORC_MAX_ACTIONS];
  double action_len[ORC_MAX_ACTIONS];
  double other_return[ORC_MAX_AGENTS];
  double mean_cost;
  uint32_t steps;   /* execute() calls: first steps + rollout steps */
} orc_action_values_result;

void orc_action_values(const orc_env* env, const mamcts_params* mp, const orc_state* start, uint32_t start_depth,
                       int32_t kind, const uint32_t* const_actions, uint64_t seed, uint32_t stream_hi,
                       uint32_t stream_lo, orc_action_values_result* out);

/* The K-rollout mean of one leaf in the engine's fixed reduction order.
 * vals[k], k < K. K in {1,2,4,8,16} or a multiple of 32 (<=256) or a multiple of 256. */
double orc_reduce_mean(const double* vals, uint32_t K);

/* ---- CrossingState<float> (environments/crossing_state.h with Domain = float) ------------------ */

typedef struct orc_state_f32 {
  float x[ORC_MAX_AGENTS];
  float last[ORC_MAX_AGENTS];
  uint8_t terminal, goal, collided, pad;
} orc_state_f32;

typedef struct orc_rollout_result_f32 {
  double ego_return;
  double other_return[ORC_MAX_AGENTS];
  double cost0;
  double step_length;
  uint32_t steps;
  orc_state_f32 final_state;
} orc_rollout_result_f32;

/* execute (:115-163). action[0] = ego action index (as float), action[1..] = the others' velocities. */
void orc_crossing_f32_execute(const mamcts_crossing_params_f32* cp, const orc_state_f32* s, const float* action,
                              orc_state_f32* next, double* ego_reward, double* cost0);
/* RandomHeuristic::rollout over it. kind REPLAY: replay[step*A + agent] floats; REFERENCE_CONST / PHILOX:
 * index-valued, an other agent's index i acts as the float with bit pattern i (crossing_state_common.h:
 * 19-27). */
void orc_rollout_f32(const mamcts_crossing_params_f32* cp, const mamcts_params* mp, const orc_state_f32* start,
                     uint32_t start_depth, int32_t kind, const float* replay, uint32_t replay_steps,
                     const uint32_t* const_actions, uint64_t seed, uint32_t stream_hi, uint32_t stream_lo,
                     orc_rollout_result_f32* out);
/* AgentPolicyCrossingState<float>::calculate_action (environments/crossing_state_agent_policy.h:35-55 with
 * Domain = float): every operation a single FP32 operation in the reference's order. */
float orc_crossing_policy_action_f32(const mamcts_crossing_params_f32* cp, float agent_x, float agent_last, float ego_x,
                                     float ego_last, float desired_gap);
/* orc_rollout_f32 with kind HYPOTHESIS: the ego draws uniformly (like PHILOX), other agent j draws its desired
 * gap as gap_lo[j] + (gap_hi[j] - gap_lo[j]) * (w / 65536) from the 16-bit word w of its Philox position (the
 * engine's stand-in for AgentPolicyCrossingState<float>::act's uniform_real_distribution, agent_policy.h:78-84;
 * exactness for the reference's own generator is REPLAY) and acts by orc_crossing_policy_action_f32. */
void orc_rollout_f32_hypothesis(const mamcts_crossing_params_f32* cp, const mamcts_params* mp, const orc_state_f32* start,
                                uint32_t start_depth, uint64_t seed, uint32_t stream_hi, uint32_t stream_lo,
                                const float* gap_lo, const float* gap_hi, orc_rollout_result_f32* out);

/* ---- backup: HypothesisStatistic (mcts/hypothesis/hypothesis_statistic.h:100-134,158-161) ---- */
/* Same arrays as mamcts_backup_hypothesis / mamcts_hyp_stats (HOST), plain sequential restatement. */
void orc_backup_hypothesis(double gamma, uint32_t num_others, uint32_t n_paths, const uint32_t* leaf_node,
                           const uint8_t* leaf_has_heuristic, const double* leaf_cost, uint32_t num_costs,
                           const uint32_t* path_offset, const uint32_t* path_len, const uint32_t* edge_node,
                           const uint8_t* edge_hypothesis, const uint8_t* edge_action, const double* edge_cost,
                           int32_t chance_constrained_update, mamcts_hyp_stats* stats);

/* ---- backup: UctStatistic (mcts/statistics/uct_statistic.h) -------------------------------- */

typedef struct orc_stat {
  double value;                       /* value_ */
  double latest_return;               /* latest_return_ */
  uint32_t total_visits;              /* total_node_visits_ */
  uint32_t num_actions;
  uint32_t num_expanded;
  uint8_t expanded[ORC_MAX_ACTIONS];  /* action present in ucb_statistics_ */
  uint32_t n[ORC_MAX_ACTIONS];        /* action_count_ */
  double q[ORC_MAX_ACTIONS];          /* action_value_ */
  uint32_t collected_action;          /* collected_reward_.first (node_statistic.h:115-121) */
  double collected_reward;            /* collected_reward_.second */
} orc_stat;

void orc_stat_init(orc_stat* s, uint32_t num_actions);
/* UctStatistic::update_from_heuristic (uct_statistic.h:100-110). */
void orc_stat_update_from_heuristic(orc_stat* s, double h);
/* UctStatistic::update_statistics_from_backpropagated (uct_statistic.h:134-144). */
void orc_stat_backprop(orc_stat* s, double gamma, double child_latest_return);
/* UctStatistic::choose_next_action (uct_statistic.h:61-76, 223-244, 256-262, 122-132);
 * expand_order = the statistic's deterministic widening order (mamcts_reference_expand_order). */
uint32_t orc_stat_choose_next_action(orc_stat* s, const mamcts_params* mp,
                                     const uint32_t* expand_order);
/* UctStatistic::get_best_action (uct_statistic.h:78-89). */
uint32_t orc_stat_best_action(const orc_stat* s);
/* UctStatistic::merge_node_statistics (uct_statistic.h:165-184) over n ego root statistics. */
void orc_stat_merge(const orc_stat* const* stats, uint32_t n, orc_stat* merged);

/* ---- whole search: Mcts::search / iterate (mcts/mcts.h:166-186,294-333) + StageNode
 * (mcts/stage_node.h:167-271) on an arena tree --------------------------------------------- */

typedef struct orc_tree orc_tree;

typedef struct orc_search_config {
  orc_env env;
  mamcts_params mcts;
  int32_t action_source;       /* REFERENCE_CONST or PHILOX */
  uint32_t rollouts_per_leaf;  /* K */
  uint64_t philox_seed;
  uint32_t tree_id;            /* global tree id = Philox stream_hi */
  /* per-agent widening order, [A][ORC_MAX_ACTIONS] (mamcts_reference_expand_order) */
  uint32_t expand_order[ORC_MAX_AGENTS][ORC_MAX_ACTIONS];
} orc_search_config;

orc_tree* orc_tree_create(const orc_search_config* cfg, const orc_state* root);
void orc_tree_destroy(orc_tree* t);
/* Runs `iterations` MCTS iterations. */
void orc_tree_run(orc_tree* t, uint32_t iterations);
uint32_t orc_tree_num_nodes(const orc_tree* t);
uint64_t orc_tree_rollout_steps(const orc_tree* t);
const orc_stat* orc_tree_root_stat(const orc_tree* t, uint32_t agent);
/* Canonical dump, same record format as mamcts_search_dump_tree. Returns number of records. */
uint32_t orc_tree_dump(const orc_tree* t, double* records, uint32_t capacity_records,
                       uint32_t max_actions);

#ifdef __cplusplus
}
#endif
#endif
