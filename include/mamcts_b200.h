/*
 * mamcts_b200 - C ABI of the B200-native rollout-and-backup engine for multi-agent MCTS.
 *
 * This header is the drop-in boundary: plain C types, caller-owned buffers, integer status
 * codes (0 = ok, <0 = error; message via mamcts_last_error). Nothing here throws and no
 * torch / STL type crosses the boundary.
 *
 * Every entry point names the reference interface (juloberno/mamcts, file:line) it replaces.
 * Layout conventions:
 *   - agents are POSITIONAL: index 0 is the ego agent, 1..A-1 the other agents in the order
 *     of State::get_other_agent_idx() (mcts/stage_node.h:172-176,190-195); the host adapter
 *     maps agent *ids* (4/5 in SimpleState) to positions.
 *   - batched state is SoA over leaves: x[a * n + leaf].
 *   - all floating point is IEEE binary64 evaluated in the reference's operation order with no
 *     fused multiply-add, so REPLAY / REFERENCE_CONST results are bit-identical to the CPU.
 */
#ifndef MAMCTS_B200_H_
#define MAMCTS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAMCTS_ABI_VERSION 1
#define MAMCTS_MAX_AGENTS 16        /* ego + up to 15 others */
#define MAMCTS_MAX_ACTIONS 32       /* per agent (CrossingState<int>: 4 ego / 7 other) */

/* status codes */
#define MAMCTS_OK 0
#define MAMCTS_ERR_INVALID (-1)     /* bad argument */
#define MAMCTS_ERR_CUDA (-2)        /* CUDA runtime error (message has cudaGetErrorString) */
#define MAMCTS_ERR_NCCL (-3)        /* NCCL error or NCCL not loadable */
#define MAMCTS_ERR_NOMEM (-4)       /* device arena exhausted */
#define MAMCTS_ERR_UNSUPPORTED (-5) /* e.g. too many agents for the compiled kernels */

/* where caller buffers live */
#define MAMCTS_MEM_HOST 0
#define MAMCTS_MEM_DEVICE 1

/* rollout action sources */
#define MAMCTS_SRC_REPLAY 0          /* consume recorded reference joint actions           */
#define MAMCTS_SRC_REFERENCE_CONST 1 /* the reference's constant first-draw action (F1)    */
#define MAMCTS_SRC_PHILOX 2          /* counter-based Philox4x32-10 uniform sampling       */
#define MAMCTS_SRC_HYPOTHESIS 3      /* CrossingState: other agents follow their behaviour
                                        hypothesis (gap-keeping policy), ego uniform (C4)  */

/* environments */
#define MAMCTS_ENV_SIMPLE 0          /* test/uct/simple_state.h                            */
#define MAMCTS_ENV_CROSSING_I32 1    /* environments/crossing_state.h, Domain = int        */

typedef struct mamcts_engine mamcts_engine;

typedef struct mamcts_config {
  int32_t device;        /* CUDA device ordinal */
  uint32_t flags;        /* reserved, 0 */
} mamcts_config;

/* The MctsParameters fields consumed on the hot path
 * (reference: mcts/mcts_parameters.h:18-41, defaults :97-116). Wall-clock limits
 * (MAX_SEARCH_TIME, random_heuristic.MAX_SEARCH_TIME) have no device counterpart: the engine
 * is bounded by iteration / depth caps only. */
typedef struct mamcts_params {
  double discount_factor;              /* DISCOUNT_FACTOR */
  double uct_lower_bound;              /* uct_statistic.LOWER_BOUND */
  double uct_upper_bound;              /* uct_statistic.UPPER_BOUND */
  double uct_exploration_constant;     /* uct_statistic.EXPLORATION_CONSTANT */
  double uct_pw_k;                     /* uct_statistic.PROGRESSIVE_WIDENING_K */
  double uct_pw_alpha;                 /* uct_statistic.PROGRESSIVE_WIDENING_ALPHA */
  uint32_t random_seed;                /* RANDOM_SEED */
  uint32_t max_number_of_iterations;   /* MAX_NUMBER_OF_ITERATIONS */
  uint32_t max_search_depth;           /* MAX_SEARCH_DEPTH */
  uint32_t max_number_of_nodes;        /* MAX_NUMBER_OF_NODES */
  uint32_t use_bound_estimation;       /* USE_BOUND_ESTIMATION */
  uint32_t rh_max_number_of_iterations;/* random_heuristic.MAX_NUMBER_OF_ITERATIONS */
  uint32_t num_parallel_mcts;          /* NUM_PARALLEL_MCTS (root-parallel trees) */
  uint32_t reserved;
} mamcts_params;

/* CrossingStateParameters<int> (reference: environments/crossing_state_parameters.h:15-33). */
typedef struct mamcts_crossing_params {
  double reward_collision;
  double reward_goal_reached;
  double reward_step;
  uint32_t num_other_agents;
  uint32_t num_other_actions;
  int32_t cost_only_collision;
  int32_t min_velocity_ego;
  int32_t max_velocity_ego;
  int32_t min_velocity_other;
  int32_t max_velocity_other;
  int32_t chain_length;
  int32_t ego_goal_pos;
  int32_t reserved;
} mamcts_crossing_params;

/* CrossingStateParameters<float> (same struct with Domain = float; defaults
 * default_crossing_state_parameters<float>(): 30 other actions, crossing_state_parameters.h:50-51). */
typedef struct mamcts_crossing_params_f32 {
  double reward_collision;
  double reward_goal_reached;
  double reward_step;
  uint32_t num_other_agents;
  uint32_t num_other_actions;
  int32_t cost_only_collision;
  float min_velocity_ego;
  float max_velocity_ego;
  float min_velocity_other;
  float max_velocity_other;
  float chain_length;
  float ego_goal_pos;
  int32_t reserved;
} mamcts_crossing_params_f32;

/* SimpleState constants (reference: test/uct/simple_state.h:16,98-101). */
typedef struct mamcts_simple_params {
  int32_t winning_state_length;  /* 10 */
  int32_t loosing_state_length;  /* -1 */
} mamcts_simple_params;

/* Where rollout joint actions come from. Exactly one block is read, chosen by `kind`. */
typedef struct mamcts_action_source {
  int32_t kind;                     /* MAMCTS_SRC_* */
  int32_t mem;                      /* MAMCTS_MEM_* of the pointers below */
  /* REPLAY: actions[(leaf * replay_stride + step) * A + agent], int8. Ego entry = action index;
   * CrossingState other entry = the Domain value the reference bit-casts out of the ActionIdx
   * (environments/crossing_state_common.h:19-27), SimpleState entry = action index.
   * replay_steps[leaf] = number of recorded steps for that leaf (rollout_per_leaf must be 1). */
  const int8_t* replay_actions;
  const float* replay_actions_f32;  /* the same for CrossingState<float> (mamcts_rollout_crossing_f32):
                                       ego entry = action index, other entry = velocity */
  const uint32_t* replay_steps;
  uint32_t replay_stride;
  uint32_t reserved;
  /* REFERENCE_CONST: the one joint action every rollout step applies (F1), as action indices. */
  uint32_t const_actions[MAMCTS_MAX_AGENTS];
  /* PHILOX: key = philox_seed; the stream of rollout k of leaf i is
   * (stream_hi[i] if given else stream_hi_base,
   *  stream_lo[i] + k if given else stream_lo_base + i*K + k), modulo 2^32. */
  uint64_t philox_seed;
  const uint32_t* stream_hi;        /* optional per-leaf, may be NULL */
  const uint32_t* stream_lo;        /* optional per-leaf, may be NULL */
  uint32_t stream_hi_base;
  uint32_t stream_lo_base;
  /* HYPOTHESIS (CrossingState only; Philox streams as above): the desired-gap range [lo, hi] of the
   * behaviour hypothesis currently assigned to other agent j of leaf i, at [j * n + i] - what
   * HypothesisBeliefTracker::sample_current_hypothesis (mcts/hypothesis/hypothesis_belief_tracker.h:
   * 136-149) selected for this iteration. Every step the agent draws a gap uniformly from the range
   * (AgentPolicyCrossingState<int>::act, environments/crossing_state_agent_policy.h:68-75) and moves
   * by calculate_action (:35-55); the ego agent draws uniformly among its actions. */
  const int32_t* hyp_gap_lo;
  const int32_t* hyp_gap_hi;
  /* HYPOTHESIS for mamcts_rollout_crossing_f32 (AgentPolicyCrossingState<float>::act, agent_policy.h:78-84): the
   * same ranges as floats; every step the agent's desired gap is lo + (hi - lo) * (w / 65536) with w the 16-bit
   * word of its Philox position (the reference draws a std::uniform_real_distribution from the policy's own
   * mt19937 - exactness for that generator is REPLAY), then calculate_action in FP32. */
  const float* hyp_gap_lo_f32;
  const float* hyp_gap_hi_f32;
} mamcts_action_source;

/* Outputs of a batched rollout. Any pointer may be NULL (not produced). `mem` says where they
 * live. Per-leaf arrays hold the MEAN over the K rollouts of the leaf (K = 1: the rollout's own
 * values, bit-identical to RandomHeuristic::rollout). */
typedef struct mamcts_rollout_out {
  int32_t mem;
  int32_t reserved;
  double* ego_return;          /* [n]        std::get<0> of rollout()                        */
  double* other_return;        /* [(A-1)*n]  SoA: other j of leaf i at [j*n + i]             */
  double* ego_cost;            /* [n]        accum_cost[0] (CrossingState; SimpleState: none) */
  double* step_length;         /* [n]        executed_step_length                            */
  uint64_t* total_steps;       /* [1]        sum of execute() calls over all n*K rollouts    */
  /* per-rollout parity outputs, index r = leaf*K + k */
  uint32_t* steps;             /* [n*K] */
  int32_t* final_x;            /* [A*n*K]  SoA by agent: x positions of the last state       */
  int32_t* final_last_action;  /* [A*n*K]  SoA by agent: last_action (CrossingState)         */
  uint8_t* final_flags;        /* [n*K]    bit0 terminal, bit1 goal, bit2 collided (Crossing) */
  float* final_x_f32;          /* [A*n*K]  the two above for mamcts_rollout_crossing_f32        */
  float* final_last_action_f32;/* [A*n*K]                                                     */
  double* reward_trace;        /* [n*K*trace_stride] ego step rewards, replay/const parity   */
  uint32_t trace_stride;
  uint32_t reserved2;
  /* one decision, many rollouts (BASELINE configs[2]): {sum, sum of squares} per agent position over ALL n*K
   * per-rollout returns, then the rollout count - 2A+1 doubles in the layout of mamcts_rollout_moments,
   * reduced on the device right after the rollouts (and all-reduced over the engine's communicator when one
   * is attached), so a 2^20-rollout decision returns 2A+1 doubles instead of 2^20 values. With K = 1 the
   * per-leaf arrays ego_return (and other_return when A > 1) must be requested too. */
  double* moments;
} mamcts_rollout_out;

/* ---- lifetime ------------------------------------------------------------------------- */

/* Engine = one CUDA device + one stream + scratch + (optional) NCCL communicator.
 * Replaces nothing in the reference (it has no device); created once per host `Mcts`/thread.
 * Handles may be used from several host threads; calls on one engine are serialised. */
int mamcts_create(const mamcts_config* cfg, mamcts_engine** out);
void mamcts_destroy(mamcts_engine* e);
const char* mamcts_last_error(const mamcts_engine* e); /* NULL engine: last global error */
int mamcts_abi_version(void);
/* Use a caller-owned cudaStream_t (e.g. torch's current stream) instead of the engine's own. */
int mamcts_set_stream(mamcts_engine* e, void* cuda_stream);
void* mamcts_get_stream(mamcts_engine* e);
/* Blocks until everything queued on the engine stream has finished. */
int mamcts_sync(mamcts_engine* e);
/* Number of kernels this engine has launched since creation (bench `gpu_launches`). */
uint64_t mamcts_launch_count(const mamcts_engine* e);

/* torch-free device buffers for C/C++ callers. */
int mamcts_device_alloc(mamcts_engine* e, size_t bytes, void** dptr);
int mamcts_device_free(mamcts_engine* e, void* dptr);
int mamcts_memcpy_h2d(mamcts_engine* e, void* dst, const void* src, size_t bytes);
int mamcts_memcpy_d2h(mamcts_engine* e, void* dst, const void* src, size_t bytes);

/* Defaults identical to mcts_default_parameters() (mcts/mcts_parameters.h:97-135) and
 * default_crossing_state_parameters<int>() (environments/crossing_state_parameters.h:36-59). */
void mamcts_default_params(mamcts_params* p);
void mamcts_default_crossing_params(mamcts_crossing_params* p);
void mamcts_default_simple_params(mamcts_simple_params* p);

/* The reference's deterministic per-statistic draws, computed on the host
 * with libstdc++'s std::mt19937 + std::uniform_int_distribution exactly as
 * UctStatistic::choose_next_action does (mcts/statistics/uct_statistic.h:61-69):
 * order[k] = the k-th action a statistic with `num_actions` actions expands. order[0] is the
 * constant rollout action of REFERENCE_CONST. */
int mamcts_reference_expand_order(uint32_t random_seed, uint32_t num_actions, uint32_t* order);

/* ---- rollouts: RandomHeuristic::rollout (mcts/heuristics/random_heuristic.h:27-104) ------ */

/* n_leaves start states, rollouts_per_leaf (K) rollouts each, through
 * CrossingState<int>::execute (environments/crossing_state.h:115-163).
 *   state_x, state_last_action: [A*n] SoA (agent-major), int32; state_last_action may be NULL
 *   (defaults to 2, environments/crossing_state_common.h:31).
 *   state_flags: [n] bit0 = terminal (rollout loop skipped), may be NULL.
 *   start_depth: [n] current tree depth of the leaf (MAX_SEARCH_DEPTH cut-off), may be NULL = 0.
 * K must be 1, 2, 4, 8, 16 or a multiple of 32. mem = MAMCTS_MEM_* of the state arrays. */
int mamcts_rollout_crossing_i32(mamcts_engine* e, const mamcts_params* mp,
                                const mamcts_crossing_params* cp, uint32_t n_leaves,
                                uint32_t rollouts_per_leaf, int32_t mem,
                                const int32_t* state_x, const int32_t* state_last_action,
                                const uint8_t* state_flags, const uint32_t* start_depth,
                                const mamcts_action_source* src, mamcts_rollout_out* out);

/* Same through CrossingState<float>::execute (Domain = float: FP32 positions and velocities, every
 * operation the reference's own, so REPLAY is bit-exact). state_x / state_last_action: [A*n] float SoA
 * (last action default 2.0f). Action sources: REPLAY (src->replay_actions_f32), REFERENCE_CONST, PHILOX
 * (index-valued: an other agent's index i is the float with bit pattern i, as the reference's bit-cast
 * yields). Per-rollout parity outputs go to out->final_x_f32 / final_last_action_f32. */
int mamcts_rollout_crossing_f32(mamcts_engine* e, const mamcts_params* mp,
                                const mamcts_crossing_params_f32* cp, uint32_t n_leaves,
                                uint32_t rollouts_per_leaf, int32_t mem, const float* state_x,
                                const float* state_last_action, const uint8_t* state_flags,
                                const uint32_t* start_depth, const mamcts_action_source* src,
                                mamcts_rollout_out* out);
void mamcts_default_crossing_params_f32(mamcts_crossing_params_f32* p);

/* Same through SimpleState::execute (test/uct/simple_state.h:26-54); A = 2.
 *   state_len: [n] int32 state_length_. */
int mamcts_rollout_simple(mamcts_engine* e, const mamcts_params* mp,
                          const mamcts_simple_params* sp, uint32_t n_leaves,
                          uint32_t rollouts_per_leaf, int32_t mem, const int32_t* state_len,
                          const uint32_t* start_depth, const mamcts_action_source* src,
                          mamcts_rollout_out* out);

/* ---- per-ego-action heuristic: ActionValueRandomHeuristic::calculate_heuristic_values ---------
 * (mcts/heuristics/action_value_random_heuristic.h:27-112), its compute part (:37-85): for every leaf and every
 * ego action a, ONE first step of CrossingState<int>::execute with the joint action (a, what a fresh statistic of
 * every other agent chooses) and ONE RandomHeuristic::rollout from the resulting state at depth + 1; n_leaves *
 * nA_ego rollouts in one launch. Outputs per leaf (HOST or DEVICE, any may be NULL):
 *   action_return[leaf * nA + a]      = gamma * first-step ego reward + gamma * rollout return          (:64-65)
 *   action_cost[leaf * nA + a]        = first-step cost[0] + rollout cost[0], not discounted           (:73-81)
 *   action_step_length[leaf * nA + a] = 1 + rollout length                                             (:83-84)
 *   other_return[j * n + leaf]        = the LAST action's rollout return of other j (+ gamma * 0): the reference
 *                                       ties the rollout's map to the accumulator itself (:58-60)
 *   mean_cost[leaf]                   = the mean of the action costs, added in the order the reference's
 *                                       std::unordered_map yields them (:97-103); what every other agent's
 *                                       statistic receives through set_heuristic_estimate
 * A terminal leaf tries no action (:38): all of its outputs are 0. These arrays are the arguments of the
 * map-valued SE::set_heuristic_estimate(action_returns, action_costs) (:93; CostConstrainedStatistic, or
 * UctStatistic::set_heuristic_estimate_from_backpropagated(map), uct_statistic.h:156-163).
 * Action sources: MAMCTS_SRC_REFERENCE_CONST (src->const_actions[1..]: the others' first-draw constants; the
 * rollouts use const_actions[0..]) and MAMCTS_SRC_PHILOX: with (hi, lo) = the leaf's stream (stream_hi[leaf] or
 * stream_hi_base, stream_lo[leaf] or stream_lo_base + leaf), action a's rollout draws from stream
 * (hi, lo * nA + a) and its first step's other agents from (step 0, agent j) of stream (hi ^ 0x80000000, lo * nA + a). */
typedef struct mamcts_action_values_out {
  int32_t mem;
  int32_t reserved;
  double* action_return;        /* [n * nA_ego] */
  double* action_cost;          /* [n * nA_ego] */
  double* action_step_length;   /* [n * nA_ego] */
  double* other_return;         /* [(A-1) * n] */
  double* mean_cost;            /* [n] */
  uint64_t* total_steps;        /* [1] execute() calls: first steps + rollout steps */
} mamcts_action_values_out;

int mamcts_action_values_crossing_i32(mamcts_engine* e, const mamcts_params* mp,
                                      const mamcts_crossing_params* cp, uint32_t n_leaves, int32_t mem,
                                      const int32_t* state_x, const int32_t* state_last_action,
                                      const uint8_t* state_flags, const uint32_t* start_depth,
                                      const mamcts_action_source* src, mamcts_action_values_out* out);

/* ---- backup: UctStatistic (mcts/statistics/uct_statistic.h) ------------------------------ */

/* Statistic storage, SoA over (node, agent) slots s = node * A + agent; action stride
 * MAMCTS_STAT_STRIDE... see mamcts_uct_stats. Mirrors UctStatistic's members
 * (uct_statistic.h:272-289): value_, latest_return_, total_node_visits_,
 * ucb_statistics_[a].{action_count_, action_value_}. */
typedef struct mamcts_uct_stats {
  int32_t mem;
  uint32_t num_slots;       /* nodes * A */
  uint32_t max_actions;     /* row stride of q / n */
  uint32_t reserved;
  double* value;            /* [num_slots]              value_               */
  double* latest_return;    /* [num_slots]              latest_return_       */
  uint32_t* total_visits;   /* [num_slots]              total_node_visits_   */
  double* q;                /* [num_slots*max_actions]  action_value_        */
  uint32_t* n;              /* [num_slots*max_actions]  action_count_        */
} mamcts_uct_stats;

/* One batch of independent backup paths (one per tree / leaf), replacing the loop
 * mcts/mcts.h:309-329:
 *   leaf update  : UctStatistic::update_from_heuristic (uct_statistic.h:100-110) when
 *                  leaf_has_heuristic[p] != 0, with the rollout values ego_return / other_return.
 *   ancestors    : UctStatistic::update_statistics_from_backpropagated (uct_statistic.h:134-144)
 *                  from the leaf's parent up to the root.
 * Path p has path_len[p] edges; edge e (0 = nearest the leaf) at flat index
 * path_offset[p] + e: node index edge_node[.], and per agent a the collected (action, reward)
 * of NodeStatistic::collect (mcts/node_statistic.h:115-121): edge_action[idx*A + a],
 * edge_reward[idx*A + a]. Paths must not share nodes within one call (different trees). */
int mamcts_backup_uct(mamcts_engine* e, const mamcts_params* mp, uint32_t num_agents,
                      uint32_t n_paths, int32_t mem, const uint32_t* leaf_node,
                      const uint8_t* leaf_has_heuristic, const double* ego_return,
                      const double* other_return /* [(A-1)*n_paths] SoA */,
                      const uint32_t* path_offset, const uint32_t* path_len,
                      const uint32_t* edge_node, const uint8_t* edge_action,
                      const double* edge_reward, mamcts_uct_stats* stats);

/* ---- backup: HypothesisStatistic (mcts/hypothesis/hypothesis_statistic.h), C4 ----------------- */

/* Statistic storage of the OTHER agents in hypothesis-MCTS, SoA over slots s = node * num_others + j
 * (j = other agent position - 1). Mirrors HypothesisStatistic's members (hypothesis_statistic.h:
 * 310-318): ego_cost_value_, latest_ego_cost_, total_node_visits_, num_expanded_actions_,
 * total_node_visits_hypothesis_[h] and ucb_statistics_[h][a].{action_count_, action_ego_cost_}.
 * visits_hypothesis has num_hypotheses + 1 entries per slot: the last one is the key
 * HYPOTHESIS_ID_NOT_SET that a statistic which never chose an action counts its heuristic update
 * under (update_from_heuristic :100-110 reads hypothesis_id_current_iteration_ of a fresh object).
 * Actions are small indices chosen by the caller (for CrossingState: velocity - MIN_VELOCITY_OTHER;
 * the reference keys its maps by the bit-cast Domain value). */
typedef struct mamcts_hyp_stats {
  int32_t mem;
  uint32_t num_slots;         /* nodes * num_others */
  uint32_t num_hypotheses;    /* H */
  uint32_t max_actions;       /* MA */
  double* ego_cost_value;     /* [num_slots] */
  double* latest_ego_cost;    /* [num_slots] */
  uint32_t* total_visits;     /* [num_slots] */
  uint32_t* num_expanded;     /* [num_slots] */
  uint32_t* visits_hypothesis;/* [num_slots * (H + 1)] */
  double* action_cost;        /* [num_slots * H * MA] */
  uint32_t* action_count;     /* [num_slots * H * MA] */
} mamcts_hyp_stats;

/* One batch of independent backup paths for the other agents' HypothesisStatistic, the C4 counterpart
 * of mamcts_backup_uct (same path / edge layout; the ego agent's UctStatistic goes through
 * mamcts_backup_uct):
 *   leaf update : update_from_heuristic (hypothesis_statistic.h:100-110) with the heuristic's ego cost
 *                 estimate = mean of leaf_cost[p*C .. p*C+C) (set_heuristic_estimate :158-161) when
 *                 leaf_has_heuristic[p] != 0.
 *   ancestors   : update_statistic (:112-134): latest = mean(edge_cost) + gamma * latest(child), or
 *                 max(mean(edge_cost), latest(child)) when chance_constrained_update != 0; running means
 *                 of ucb_statistics_[h][a] and ego_cost_value_ (the latter over the visits UNDER h).
 * edge_hypothesis / edge_action: [edge * num_others + j]; edge_cost: [edge * C + c], the ego cost vector
 * collected on the edge (CrossingState: C = 2, crossing_state.h:149-153). */
int mamcts_backup_hypothesis(mamcts_engine* e, const mamcts_params* mp, uint32_t num_others,
                             uint32_t n_paths, int32_t mem, const uint32_t* leaf_node,
                             const uint8_t* leaf_has_heuristic, const double* leaf_cost,
                             uint32_t num_costs /* C */, const uint32_t* path_offset,
                             const uint32_t* path_len, const uint32_t* edge_node,
                             const uint8_t* edge_hypothesis, const uint8_t* edge_action,
                             const double* edge_cost, int32_t chance_constrained_update,
                             mamcts_hyp_stats* stats);

/* ---- root-parallel merge: UctStatistic::merge_node_statistics (uct_statistic.h:165-184) --- */

/* Reduces the ego root statistics of n_trees local trees (slot = root_slot[t]) to
 * {sum_q[a], sum_n[a], occurrences[a]} + total visits, all-reduces them over the engine's
 * communicator when one is attached (one ncclAllReduce(sum, f64) of 3*nA+1 values per decision),
 * then finishes as the reference does: q[a] = sum_q[a]/occ[a], value = mean_a q[a] over seen
 * actions, best = first argmax (uct_statistic.h:78-89). Outputs are HOST arrays of num_actions. */
int mamcts_root_merge(mamcts_engine* e, const mamcts_uct_stats* stats, uint32_t n_trees,
                      int32_t mem, const uint32_t* root_slot, uint32_t num_actions,
                      double* merged_q, uint64_t* merged_n, uint32_t* merged_occurrences,
                      uint64_t* merged_total_visits, double* merged_value,
                      uint32_t* best_action);

/* ---- one decision, many rollouts (BASELINE configs[2]): the exchange step ------------------------ */

/* {sum, sum of squares} of the per-rollout returns of every agent, then the rollout count:
 * moments[2*a], moments[2*a+1] for agent position a (0 = ego_return[n], a > 0 = other_return[(a-1)*n ..]),
 * moments[2*(num_others+1)] = n. Reduced on the device in a fixed order, then all-reduced over the
 * engine's communicator when one is attached (one ncclAllReduce of 2A+1 doubles per decision,
 * SURVEY.md 8e; the reference has no counterpart - it runs one rollout per leaf, mcts/mcts.h:310).
 * mem = where ego_return / other_return AND moments live. */
int mamcts_rollout_moments(mamcts_engine* e, int32_t mem, const double* ego_return,
                           const double* other_return, uint32_t num_others, uint32_t n,
                           double* moments);

/* ---- NCCL communicator over the GPUs of one host (NVLink 5 / NVSwitch) -------------------- */

#define MAMCTS_NCCL_UNIQUE_ID_BYTES 128
/* rank 0 creates the id; the caller ships the bytes to the other ranks (any side channel). */
int mamcts_comm_get_unique_id(mamcts_engine* e, uint8_t id[MAMCTS_NCCL_UNIQUE_ID_BYTES]);
int mamcts_comm_init_rank(mamcts_engine* e, const uint8_t id[MAMCTS_NCCL_UNIQUE_ID_BYTES],
                          int32_t world_size, int32_t rank);
int mamcts_comm_destroy(mamcts_engine* e);
/* In-place sum all-reduce of `count` doubles that live on the device. */
int mamcts_allreduce_f64(mamcts_engine* e, double* dptr, size_t count);

/* ---- device-resident root-parallel search (next row §8f-1) -------------------------------- */
/* Replaces Mcts<S, UctStatistic, UctStatistic, RandomHeuristic>::search for T independent trees
 * (mcts/mcts.h:133-141,166-221,294-333; mcts/stage_node.h:167-271). See mamcts_search_* below. */

/* Tree storage of the device-resident search.
 * CLASSIC: one fixed-size record per node (any number of agents, any depth).
 * COMPACT: nodes that were never selected again keep only their heuristic values (16 B for 2
 *          agents) and are promoted to a full record on their second selection; needs a direct child
 *          table (2 agents) and trees no deeper than 64 levels. Same results, a third of the memory.
 * AUTO   : COMPACT where possible, else CLASSIC. */
#define MAMCTS_LAYOUT_AUTO 0
#define MAMCTS_LAYOUT_CLASSIC 1
#define MAMCTS_LAYOUT_COMPACT 2

typedef struct mamcts_search_config {
  int32_t env;                 /* MAMCTS_ENV_* */
  int32_t action_source;       /* MAMCTS_SRC_REFERENCE_CONST or MAMCTS_SRC_PHILOX */
  uint32_t num_trees;          /* local trees on this engine */
  uint32_t first_tree_id;      /* global id of local tree 0 (root-parallel sharding, §8e) */
  uint32_t rollouts_per_leaf;  /* K = 1, 2, 4, 8 or 16 (K > 1 only with PHILOX): leaf value = mean of K rollouts */
  uint32_t max_nodes_per_tree; /* arena size; iterations stop expanding past it like
                                  MAX_NUMBER_OF_NODES (mcts/stage_node.h:184) */
  uint32_t layout;             /* MAMCTS_LAYOUT_*: how a tree is stored in HBM */
  uint32_t max_inner_nodes_per_tree; /* compact layout: records for nodes visited more than once;
                                  0 = max_nodes_per_tree (always enough). A 10 000-iteration C2 tree
                                  needs about 2 500 (max 2 734 over 120 trees). A tree that runs out
                                  makes every later read-back fail with MAMCTS_ERR_NOMEM. */
  uint64_t philox_seed;
  mamcts_params mcts;
  mamcts_crossing_params crossing;
  mamcts_simple_params simple;
} mamcts_search_config;

typedef struct mamcts_search mamcts_search;

/* ---- hypothesis-MCTS (BASELINE configs[3]): Mcts<CrossingState<int>, UctStatistic, HypothesisStatistic,
 * RandomHeuristic>::search(state, belief_tracker) on the device (mcts/mcts.h:143-164) -------------------------- */

/* MctsParameters::hypothesis_statistic (mcts/mcts_parameters.h:71-79, defaults :118-124) and the one
 * cost_constrained_statistic field HypothesisStatistic reads (USE_CHANCE_CONSTRAINED_UPDATES[0],
 * hypothesis_statistic.h:116). */
typedef struct mamcts_hypothesis_params {
  double upper_cost_bound;                    /* UPPER_COST_BOUND */
  double lower_cost_bound;                    /* LOWER_COST_BOUND */
  double pw_k;                                /* PROGRESSIVE_WIDENING_K */
  double pw_alpha;                            /* PROGRESSIVE_WIDENING_ALPHA */
  double exploration_constant;                /* EXPLORATION_CONSTANT */
  uint32_t cost_based_action_selection;       /* COST_BASED_ACTION_SELECTION */
  uint32_t progressive_widening_hypothesis_based; /* PROGRESSIVE_WIDENING_HYPOTHESIS_BASED: must be 1 (the default) */
  uint32_t chance_constrained_update;         /* USE_CHANCE_CONSTRAINED_UPDATES[0] (empty vector = 0) */
  uint32_t policy_random_seed;                /* CrossingStateParameters::OTHER_AGENTS_POLICY_RANDOM_SEED */
} mamcts_hypothesis_params;
void mamcts_default_hypothesis_params(mamcts_hypothesis_params* p);

#define MAMCTS_MAX_HYPOTHESES 4
#define MAMCTS_HYP_MAX_OTHER_ACTIONS 8     /* MAX_VELOCITY_OTHER - MIN_VELOCITY_OTHER + 1 (default 7) */

typedef struct mamcts_hyp_search_config {
  uint32_t num_trees;
  uint32_t first_tree_id;
  uint32_t max_nodes_per_tree;         /* 0 = min(MAX_NUMBER_OF_NODES, iterations) + 1 */
  uint32_t num_hypotheses;             /* H <= MAMCTS_MAX_HYPOTHESES: CrossingState::add_hypothesis calls, in order */
  int32_t gap_lo[MAMCTS_MAX_HYPOTHESES];   /* desired-gap range of hypothesis h (AgentPolicyCrossingState ctor) */
  int32_t gap_hi[MAMCTS_MAX_HYPOTHESES];
  /* 0: every tree is the reference's tree (all random streams start where the reference's generators start).
   * 1: tree t starts its policy / statistic generators and its hypothesis sequence at offsets derived from its
   *    global id, i.e. differently seeded generators per tree: independent root-parallel trees. */
  uint32_t decorrelate_trees;
  uint32_t reserved;
  mamcts_params mcts;
  mamcts_hypothesis_params hypothesis;
  mamcts_crossing_params crossing;
} mamcts_hyp_search_config;

/* The hypothesis every other agent is assigned in every iteration, as HypothesisBeliefTracker::
 * sample_current_hypothesis (mcts/hypothesis/hypothesis_belief_tracker.h:136-149) draws it: a
 * std::discrete_distribution over the agent's beliefs from the tracker's mt19937, agents in the order the
 * tracker's std::unordered_map<AgentIdx, ...> yields them (libstdc++, like mamcts_reference_expand_order).
 * beliefs: [num_others][num_hypotheses] (agent ids 1..num_others); the generator is seeded with `seed`
 * (RANDOM_SEED_HYPOTHESIS_SAMPLING) and `skip_iterations` earlier sample_current_hypothesis() calls with the
 * same beliefs are discarded first. out: [n_iterations][num_others] (positional). A C++ host that owns a real
 * HypothesisBeliefTracker copies it and samples from the copy instead (mamcts_b200/host/device_hypothesis_mcts.h). */
int mamcts_reference_hypothesis_sequence(uint32_t seed, const float* beliefs, uint32_t num_others,
                                         uint32_t num_hypotheses, uint32_t skip_iterations,
                                         uint32_t n_iterations, uint8_t* out);

/* T device-resident hypothesis-MCTS trees: the ego agent's UctStatistic, every other agent's HypothesisStatistic
 * (choose_next_action with hypothesis-based progressive widening, random or worst-case selection among the
 * expanded actions, hypothesis_statistic.h:62-98,178-202; update_statistic :112-134), the hypothesis policies
 * acting inside the tree and the rollouts (AgentPolicyCrossingState<int>::act, crossing_state_agent_policy.h:
 * 35-55,68-75), all on the device. hypothesis_sequence: [MAX_NUMBER_OF_ITERATIONS][num_others] HOST, see above.
 * The reference's generators (one mt19937 per policy object copied with every state, one per statistic) are
 * reproduced by stream POSITIONS into host-tabulated draws of the same libstdc++ distributions, so with
 * decorrelate_trees = 0 every tree equals the reference's tree bit for bit. The handle is a mamcts_search:
 * mamcts_search_run / _reset / _counters / _root_stats (agent 0) / _merge_roots / _dump_tree / _destroy apply.
 * Dump record of a node (doubles): depth, joint-action code (ego + M * sum_j (v_j - MIN_VELOCITY_OTHER) M^j,
 * M = max(nA_ego, MA)), visit_count, children; ego {N, V, G, n[NE], q[NE]}; per other agent {N, ego cost value,
 * latest ego cost, num_expanded_actions_, visits under HYPOTHESIS_ID_NOT_SET, 0; per hypothesis {visits, number of
 * expanded actions, the actions in the iteration order of the reference's unordered_map (-1 padded)}; per
 * (hypothesis, action) {count (-1 = absent), cost}} with MA = MAX_VELOCITY_OTHER - MIN_VELOCITY_OTHER + 1. */
int mamcts_hyp_search_create(mamcts_engine* e, const mamcts_hyp_search_config* cfg, const int32_t* root_x,
                             const int32_t* root_last_action, const uint8_t* hypothesis_sequence,
                             mamcts_search** out);
/* New decision: new root state and a new hypothesis sequence (NULL keeps the resident one). */
int mamcts_hyp_search_reset(mamcts_search* s, const int32_t* root_x, const int32_t* root_last_action,
                            const uint8_t* hypothesis_sequence);

/* Allocates T tree arenas in HBM and resets every tree to the given root state
 * (root_x / root_last_action: [A] for CrossingState, root_x[0] = state_length for SimpleState). */
int mamcts_search_create(mamcts_engine* e, const mamcts_search_config* cfg, const int32_t* root_x,
                         const int32_t* root_last_action, mamcts_search** out);
int mamcts_search_destroy(mamcts_search* s);
/* Re-roots every tree (new decision) without reallocating. root_x == NULL keeps the root state
 * already resident on the device (no host->device copy). */
int mamcts_search_reset(mamcts_search* s, const int32_t* root_x, const int32_t* root_last_action);
/* Runs `iterations` more MCTS iterations on every tree (select/expand, rollout, backup),
 * asynchronously on the engine stream. */
int mamcts_search_run(mamcts_search* s, uint32_t iterations);
/* Totals since the last reset: iterations done, execute() calls in rollouts, in expansions, nodes,
 * and statistic levels walked by the backups (sum of leaf depths). Any pointer may be NULL. */
int mamcts_search_counters(mamcts_search* s, uint64_t* iterations, uint64_t* rollout_steps,
                           uint64_t* expand_steps, uint64_t* nodes, uint64_t* backup_levels);
/* Diagnostic: how many UCB arg-max selections since the last reset needed the exact FP64 evaluation
 * after the FP32 screen could not prove the winner (selection results are exact either way). */
int mamcts_search_exact_ucb_evaluations(mamcts_search* s, uint64_t* count);
/* Root statistics of local tree t, agent position a: HOST outputs, arrays of max_actions. */
int mamcts_search_root_stats(mamcts_search* s, uint32_t tree, uint32_t agent, double* value,
                             uint32_t* total_visits, double* q, uint32_t* n, uint32_t* expanded_mask,
                             uint32_t* num_nodes);
/* Root-parallel merge of all local trees' ego root statistics (+ all-reduce when a communicator is
 * attached); same outputs as mamcts_root_merge. */
int mamcts_search_merge_roots(mamcts_search* s, double* merged_q, uint64_t* merged_n,
                              uint32_t* merged_occurrences, uint64_t* merged_total_visits,
                              double* merged_value, uint32_t* best_action);
/* Canonical dump of local tree t for whole-tree parity: nodes in depth-first order with children
 * sorted by joint action; per node a record of (4 + A*(3 + 2*max_actions)) doubles:
 * depth, joint_action_code, visit_count, num_children, then per agent
 * {total_visits, value, latest_return, n[0..max_actions), q[0..max_actions)} (n = -1 if unexpanded).
 * Returns the number of records written (<= capacity_records) in *num_records. */
int mamcts_search_dump_tree(mamcts_search* s, uint32_t tree, double* records,
                            uint32_t capacity_records, uint32_t* num_records);

#ifdef __cplusplus
}
#endif
#endif /* MAMCTS_B200_H_ */
