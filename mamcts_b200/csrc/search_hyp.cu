// search_hyp.cu - device-resident hypothesis-MCTS (BASELINE configs[3]): T independent trees of
//   Mcts<CrossingState<int>, UctStatistic, HypothesisStatistic, RandomHeuristic>::search(state, belief_tracker)
// (reference mcts/mcts.h:143-164,294-333), one lane per tree, whole iterations on the device:
//   ego agent      UctStatistic: the selection and back-propagation code of the UCT search (search_kernel.cuh)
//   other agents   HypothesisStatistic (mcts/hypothesis/hypothesis_statistic.h): the hypothesis of the iteration
//                  (:62-66), hypothesis-based progressive widening (:70-81, :270-276), an action sampled from the
//                  hypothesis policy inside the tree (environments/crossing_state.h:78-82 ->
//                  crossing_state_agent_policy.h:35-55,68-75), otherwise a random (:86-93, :204-212) or the
//                  worst-case (:178-202) expanded action, ego-cost back-propagation (:100-134)
//   rollouts       RandomHeuristic::rollout with a fresh UctStatistic for the ego (its first draw)
//                  and a fresh HypothesisStatistic per other agent (always the hypothesis policy)
//
// Randomness. The reference has three kinds of generators on this path, all std::mt19937: the belief
// tracker's (one hypothesis per agent and iteration, hypothesis_belief_tracker.h:136-149) - its draws arrive
// as a host-computed sequence; one per POLICY OBJECT (the desired gap of every act(), agent_policy.h:68-75),
// copied with every state copy, so a node's state carries the generator position its parent's state had when
// the node was created, advanced by the draws made at the node itself; and one per STATISTIC (the random pick
// among the expanded actions). All policy objects are seeded alike and all statistics are seeded alike, so a
// generator is fully described by its POSITION in one common stream. The kernels keep positions (per node and
// hypothesis, per statistic) and look the draws up in tables the host fills with the real libstdc++
// distributions over the real mt19937 stream (draw value and words consumed, per position): nothing about a
// std:: distribution is re-derived on the device, and with decorrelate_trees = 0 every tree is the
// reference's tree bit for bit. The iteration order of the reference's std::unordered_map<ActionIdx, UcbPair>
// (it decides which expanded action a random index or a tie selects) is the reverse insertion order whenever
// every action key lands in a bucket of its own (libstdc++: 13 buckets, identity hash, new nodes linked at the
// front) - checked on the host for the configured velocity range; other ranges are refused.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

#include "engine.h"
#include "root_merge.h"
#include "search_host.h"
#include "search_kernel.cuh"

namespace mamcts {

#ifndef MAMCTS_HYP_RECONVERGE
#define MAMCTS_HYP_RECONVERGE 1   // 4 096 trees: 336 -> 319 ms; 12 288 trees: 908 -> 763 ms (profiles/r02_hyp_search.txt)
#endif
constexpr int kHMax = MAMCTS_MAX_HYPOTHESES;
constexpr int kHypMA = MAMCTS_HYP_MAX_OTHER_ACTIONS;
constexpr int kHypPathMax = 96;
constexpr uint32_t kNotSetHyp = 0xFFu;

// A node's record, laid out by 128-byte line in the way a visit uses it (round 2; the first version kept one array
// per member over all hypotheses, ~9 lines touched and ~6 dirtied per visit, 1 104 bytes zero-filled per new node):
//   line 0      header, state, reward / cost of the incoming edge, the ego agent's UctStatistic
//   line 1      the generator positions of the state's policy objects and, per other agent, the members of its
//               HypothesisStatistic that do not depend on the hypothesis
//   lines 2..   one line per (other agent, hypothesis): visits, the expanded actions in insertion order, their
//               counts and costs
// A visit under hypotheses (h_1 .. h_J) touches lines 0, 1 and the J lines (j, h_j). A hypothesis line is
// initialised when the agent first chooses under that hypothesis at the node (HypothesisStatistic::
// init_hypothesis_variables, hypothesis_statistic.h:52-61 - the reference does the same), so a new node costs two
// lines of stores and the three quarters of the nodes that are never selected again never touch the rest.
struct alignas(128) HypHyp {
  uint32_t visits;                   // total_node_visits_hypothesis_[h]
  uint8_t nexp;                      // ucb_statistics_[h].size()
  uint8_t present;                   // bit a: action a is a key of ucb_statistics_[h]
  uint8_t pad[2];
  uint8_t order[kHypMA];             // insertion order of the actions of ucb_statistics_[h]
  uint32_t count[kHypMA];            // action_count_
  double cost[kHypMA];               // action_ego_cost_
};
static_assert(sizeof(HypHyp) == 128, "one line per (agent, hypothesis)");

struct HypCommon {
  double cost_value;                 // ego_cost_value_
  double latest;                     // latest_ego_cost_
  uint32_t N;                        // total_node_visits_
  uint32_t nexp_total;               // num_expanded_actions_ (one per back-propagation, :132)
  uint32_t stat_pos;                 // words this statistic's generator has produced
  uint32_t visits_notset;            // total_node_visits_hypothesis_[HYPOTHESIS_ID_NOT_SET] (:106)
  uint8_t seen[kHMax];               // init_hypothesis_variables(h) ran (:52-61): the line (agent, h) is valid
  uint32_t pad;
};

template <int J, int NE>
struct alignas(128) HypNode {
  uint32_t parent, visit_count, depth;
  uint8_t flags;                     // bit0 terminal, bit1 goal, bit2 collided
  uint8_t ja[J + 1];                 // ego action index, others: velocity - MIN_VELOCITY_OTHER
  int32_t x[J + 1], last[J + 1];
  double r_in;                       // ego reward collected on the incoming edge
  double c_in;                       // mean of the ego cost vector collected on the incoming edge (:120)
  StatRec<NE> ego;
  alignas(128) uint32_t pol_pos[kHMax]; // position of the state's policy generator of hypothesis h
  HypCommon com[J];
  HypHyp hyp[J][kHMax];
};

struct HypTables {
  uint32_t pw_min_visits[kHypMA + 1];  // smallest visits_h with nexp <= pw_k * pow(visits_h, pw_alpha)
};

struct HypParams {
  void* nodes;
  uint32_t* hash;
  const uint8_t* seq;        // [seq_len][J] hypothesis of agent j in iteration i
  const uint16_t* gap_tab;   // [H][pol_len]: low byte = gap - gap_lo[h], high byte = words consumed
  const uint16_t* sel_tab;   // [kHypMA][sel_len]: low byte = index, high byte = words consumed (size s at row s-1)
  uint32_t seq_len, pol_len, sel_len;
  uint32_t H, MA, n_oth_actions;
  int32_t gap_lo[kHMax];
  int32_t min_v_other;
  uint32_t cost_based, chance, decorrelate;
  double hyp_two_c, hyp_lb, hyp_ub;
  HypTables tab;
};

// The lines a visit will touch, requested at once when the node index becomes known: one round trip to HBM for the
// level instead of one per field (profiles/r02_hyp_search.txt).
template <class Node, int J>
__device__ __forceinline__ void hyp_prefetch_node(const Node* nd, const uint8_t (&hyp)[J]) {
#pragma unroll
  for (size_t off = 0; off < offsetof(Node, hyp); off += 128) prefetch_l1(reinterpret_cast<const char*>(nd) + off);
#pragma unroll
  for (int j = 0; j < J; ++j) prefetch_l1(&nd->hyp[j][hyp[j]]);
}

__device__ __forceinline__ uint32_t hyp_hash(uint32_t parent, const uint8_t* ja, int n) {
  uint32_t h = parent * 0x9E3779B1u;
  for (int a = 0; a < n; ++a) h = (h ^ ja[a]) * 0x01000193u;
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 12;
  return h;
}

// get_worst_case_action (hypothesis_statistic.h:178-202) over the actions of hypothesis h in the container's
// iteration order (last inserted first): normalised ego cost + 2c sqrt(2 ln(visits_h) / count), strict >.
__device__ __noinline__ uint32_t hyp_worst_case(const HypHyp& o, const SearchParams& p, const HypParams& hp) {
  const uint32_t nexp = o.nexp;
  double lo = hp.hyp_lb, hi = hp.hyp_ub;
  if (p.use_bound_est) {   // calculate_bounds_based_on_stat :214-222
    double mn = 0.0, mx = 0.0;
    for (uint32_t k = 0; k < nexp; ++k) {
      const double c = o.cost[o.order[nexp - 1u - k]];
      if (k == 0) { mn = c; mx = c; }
      else { mn = c < mn ? c : mn; mx = c < mx ? mx : c; }   // std::minmax_element semantics on values
    }
    if (mn == mx) { lo = 0.0; hi = mx; } else { lo = mn; hi = mx; }
  }
  const double range = __dsub_rn(hi, lo);
  const double log2N = __dmul_rn(2.0, __ldg(p.log_table + o.visits));
  double largest = -1.7976931348623157e308;   // std::numeric_limits<double>::lowest()
  uint32_t worst = o.order[nexp - 1u];        // ucb_statistics.begin()->first
  for (uint32_t k = 0; k < nexp; ++k) {
    const uint32_t a = o.order[nexp - 1u - k];
    const double norm = __ddiv_rn(__dsub_rn(o.cost[a], lo), range);
    const double ucb = __dadd_rn(norm, __dmul_rn(hp.hyp_two_c, __dsqrt_rn(__ddiv_rn(log2N, (double)o.count[a]))));
    if (ucb > largest) { largest = ucb; worst = a; }
  }
  return worst;
}

#ifndef MAMCTS_HYP_MIN_BLOCKS
#define MAMCTS_HYP_MIN_BLOCKS 8
#endif
template <int J, int NE>
__global__ void __launch_bounds__(kSearchThreads, MAMCTS_HYP_MIN_BLOCKS) hyp_search_kernel(const __grid_constant__ SearchParams p,
                                                                   const __grid_constant__ HypParams hp) {
  using Env = CrossingEnv<J>;
  using Node = HypNode<J, NE>;
  constexpr int A = J + 1;
  __shared__ SearchTables tab;
  stage_tables(tab, p.tab);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t tree = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * p.lanes_per_warp + lane;
  // the lanes that own a tree; with MAMCTS_HYP_RECONVERGE they meet again once per iteration
  const unsigned live = __ballot_sync(0xFFFFFFFFu, lane < p.lanes_per_warp && tree < p.T);
  (void)live;
  if (lane >= p.lanes_per_warp || tree >= p.T) return;
  Node* nodes = static_cast<Node*>(hp.nodes) + (size_t)tree * p.max_nodes;
  uint32_t* hash = hp.hash + (size_t)tree * (p.hash_mask + 1);
  uint32_t num_nodes = p.tree_nodes[tree];
  const uint32_t iter0 = p.tree_iters[tree];
  unsigned long long rollout_steps = 0, expand_steps = 0, backup_levels = 0, exact_evals = 0;
  const uint32_t nodes_before = num_nodes;
  const uint32_t gid = p.first_tree_id + tree;
  // decorrelated trees: generators that start elsewhere in the stream, a rotated hypothesis sequence
  const uint32_t pol_off = hp.decorrelate ? ((gid * 2654435761u) >> 20) & 4095u : 0u;
  const uint32_t sel_off = hp.decorrelate ? ((gid * 40503u + 977u) * 2654435761u >> 20) & 4095u : 0u;
  const uint32_t seq_off = hp.decorrelate ? (gid * 7919u) % hp.seq_len : 0u;
  const uint8_t ego_const = tab.order[0][0];   // a fresh UctStatistic's first draw
  bool failed = false;
  uint32_t done = 0;

  uint32_t path_node[kHypPathMax];
  uint8_t path_ja[kHypPathMax][A];
  BackRegs path_back[kHypPathMax];

#pragma unroll 1
  for (uint32_t iter = 0; iter < p.iterations; ++iter) {   // a failed lane idles through the barriers
#if MAMCTS_HYP_RECONVERGE
    __syncwarp(live);
#endif
    if (failed) continue;
    // the hypotheses of this iteration: HypothesisBeliefTracker::sample_current_hypothesis (mcts.h:158)
    uint8_t hyp[J];
    {
      const uint32_t si = (iter0 + iter + seq_off) % hp.seq_len;
#pragma unroll
      for (int j = 0; j < J; ++j) hyp[j] = __ldg(hp.seq + (size_t)si * J + j);
    }
    uint32_t node = 0, depth = 0;
    bool do_rollout = false;
    typename Env::State st;
    uint32_t leaf_pos[kHMax];
    // ---- select & expand: mcts.h:300-305, stage_node.h:167-232 ------------------------------------------
    hyp_prefetch_node<Node, J>(nodes, hyp);   // the root
#pragma unroll 1
    while (true) {
      Node* nd = nodes + node;
      nd->visit_count += 1;                                                     // :180
      st.flags = nd->flags;
      if ((st.flags & 1u) || depth == p.max_depth || num_nodes > p.node_budget) break;   // :183-187
      if (depth >= (uint32_t)kHypPathMax) { failed = true; atomicAdd(p.counters + 7, 1ull); break; }
#pragma unroll
      for (int i = 0; i < A; ++i) { st.x[i] = nd->x[i]; st.last[i] = nd->last[i]; }
      uint8_t ja[A];
      {
        const StatRegs<NE> ego = load_stat<NE>(&nd->ego);
        BackRegs back;
        ja[0] = (uint8_t)choose_action<NE>(&nd->ego, ego, p, tab, 0, p.n_ego, exact_evals, back);   // :190
        path_back[depth] = back;
      }
      // what the other agents' choices read, requested together (one round trip for the level, see the backup):
      // the generator positions of the node's policy objects (one per hypothesis; two agents under the same
      // hypothesis draw from the same object one after the other, so they are carried in registers through the
      // agent loop and written back once) and, per agent, the counters of its hypothesis
      uint32_t pp[kHMax];
      {
        static_assert(kHMax == 4 && offsetof(Node, pol_pos) % 16 == 0, "the positions are one 16-byte load");
        const uint4 w = *reinterpret_cast<const uint4*>(&nd->pol_pos[0]);
        pp[0] = w.x; pp[1] = w.y; pp[2] = w.z; pp[3] = w.w;
      }
      uint32_t nexp0[J], vis0[J], pres0[J], spos0[J];
      bool seen0[J];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const HypHyp& o = nd->hyp[j][hyp[j]];
        nexp0[j] = o.nexp; vis0[j] = o.visits; pres0[j] = o.present;
        spos0[j] = nd->com[j].stat_pos; seen0[j] = nd->com[j].seen[hyp[j]] != 0;
      }
#pragma unroll
      for (int j = 0; j < J; ++j) {
        // the line of a hypothesis the agent has not chosen under at this node holds nothing yet (see HypNode)
        if (!seen0[j]) { nexp0[j] = 0; vis0[j] = 0; pres0[j] = 0; }
      }
      bool pp_moved = false;
#pragma unroll
      for (int j = 0; j < J; ++j) {                                             // :191-194, agent order
        HypHyp& o = nd->hyp[j][hyp[j]];
        const uint32_t h = hyp[j];                                              // hypothesis_statistic.h:64-65
        if (!seen0[j]) {                                                        // init_hypothesis_variables
          uint4* w = reinterpret_cast<uint4*>(&o);
#pragma unroll
          for (size_t i = 0; i < sizeof(HypHyp) / 16; ++i) w[i] = make_uint4(0u, 0u, 0u, 0u);
          nd->com[j].seen[h] = 1;
        }
        const uint32_t nexp = nexp0[j];
        uint32_t a = 0;
        if (failed) {
        } else if (nexp < hp.n_oth_actions && vis0[j] >= hp.tab.pw_min_visits[nexp]) {   // :270-276
          // sample from the hypothesis: the policy object of the NODE's state draws a gap (:73-78)
          const uint32_t pph = h == 0 ? pp[0] : h == 1 ? pp[1] : h == 2 ? pp[2] : pp[3];
          const uint32_t pos = pph + pol_off;
          if (pos >= hp.pol_len) { failed = true; atomicAdd(p.counters + 6, 1ull); }
          else {
            const uint32_t e = __ldg(hp.gap_tab + (size_t)h * hp.pol_len + pos);
            const uint32_t adv = e >> 8;
            pp[0] += h == 0 ? adv : 0u; pp[1] += h == 1 ? adv : 0u; pp[2] += h == 2 ? adv : 0u; pp[3] += h == 3 ? adv : 0u;
            pp_moved = true;
            const int32_t gap = hp.gap_lo[h] + (int32_t)(e & 0xFFu);
            const int32_t v = crossing_policy_action(p.env, st.x[j + 1], st.last[j + 1], st.x[0], st.last[0], gap);
            a = (uint32_t)(v - hp.min_v_other);
            if (!((pres0[j] >> a) & 1u)) {                                      // ucb_statistics_[h][a] created
              o.present = (uint8_t)(pres0[j] | (1u << a));
              o.order[nexp] = (uint8_t)a;
              o.nexp = (uint8_t)(nexp + 1u);
              o.cost[a] = 0.0;
              o.count[a] = 0;
            }
          }
        } else if (hp.cost_based) {
          a = hyp_worst_case(o, p, hp);                                         // :85-87
        } else {
          // sample_from_hypothesis_expanded_actions (:204-212): the statistic's own generator picks an index
          // into the container's iteration order (last inserted first)
          const uint32_t pos = spos0[j] + sel_off;
          if (pos >= hp.sel_len) { failed = true; atomicAdd(p.counters + 6, 1ull); }
          else {
            const uint32_t e = __ldg(hp.sel_tab + (size_t)(nexp - 1u) * hp.sel_len + pos);
            nd->com[j].stat_pos = spos0[j] + (e >> 8);
            a = o.order[nexp - 1u - (e & 0xFFu)];
          }
        }
        ja[j + 1] = (uint8_t)a;
      }
      if (pp_moved) *reinterpret_cast<uint4*>(&nd->pol_pos[0]) = make_uint4(pp[0], pp[1], pp[2], pp[3]);
      if (failed) break;
      path_node[depth] = node;
#pragma unroll
      for (int i = 0; i < A; ++i) path_ja[depth][i] = ja[i];
      // child lookup, :198
      uint32_t slot = hyp_hash(node, ja, A) & p.hash_mask;
      uint32_t child = 0;
      while (true) {
        const uint32_t c = hash[slot];
        if (c == 0) break;
        const Node* cn = nodes + c;
        hyp_prefetch_node<Node, J>(cn, hyp);   // nearly always the child the descent goes on with
        bool same = cn->parent == node;
#pragma unroll
        for (int i = 0; i < A; ++i) same &= cn->ja[i] == ja[i];
        if (same) { child = c; break; }
        slot = (slot + 1) & p.hash_mask;
      }
      if (child) { node = child; depth += 1; continue; }                        // :199-206
      // expand, :207-231: the child's state copies the policy objects as they are NOW
      int32_t act[A];
      act[0] = ja[0];
#pragma unroll
      for (int i = 1; i < A; ++i) act[i] = (int32_t)ja[i] + hp.min_v_other;
      double r_ego, r_other, c0;
      st.flags = 0;
      Env::step(p.env, st, act, r_ego, r_other, c0);
      expand_steps += 1;
      child = num_nodes++;
      Node* cn = nodes + child;
      {
        // zero the header, the ego statistic and the agents' common members (a fresh node: every statistic
        // empty); the per-hypothesis lines are initialised when first used (HypNode)
        uint4* w = reinterpret_cast<uint4*>(cn);
        for (size_t i = 0; i < offsetof(Node, hyp) / 16; ++i) w[i] = make_uint4(0u, 0u, 0u, 0u);
      }
      cn->parent = node;
      cn->depth = depth + 1;
      cn->flags = (uint8_t)st.flags;
#pragma unroll
      for (int i = 0; i < A; ++i) { cn->ja[i] = ja[i]; cn->x[i] = st.x[i]; cn->last[i] = st.last[i]; }
      cn->r_in = r_ego;
      // collected_cost_worst_case: std::accumulate over the ego cost vector {c0, 0} / its size (:120)
      cn->c_in = __ddiv_rn(__dadd_rn(__dadd_rn(0.0, c0), 0.0), 2.0);
#pragma unroll
      for (int h = 0; h < kHMax; ++h) { cn->pol_pos[h] = pp[h]; leaf_pos[h] = pp[h]; }
      hash[slot] = child;
      node = child;
      depth += 1;
      do_rollout = true;
      break;
    }
#if MAMCTS_HYP_RECONVERGE
    __syncwarp(live);   // the descent leaves its loop at different depths
#endif
    if (failed) continue;
    // ---- heuristic: mcts.h:309-312 -> RandomHeuristic::rollout<S, UctStatistic, HypothesisStatistic, H> --------
    Node* leaf = nodes + node;
    double G_ego, L[J];
    if (do_rollout) {
      double ego = 0.0, cost = 0.0;
      uint32_t it = 0;
      const uint32_t depth_room = p.max_depth > depth ? p.max_depth - depth : 0u;
      const uint32_t limit = p.rh_cap < depth_room ? p.rh_cap : depth_room;
      bool fin = Env::terminal(p.env, st) | (it >= limit);
#pragma unroll 1
      while (!fin) {
        int32_t act[A];
        act[0] = ego_const;
#pragma unroll 1
        for (int j = 0; j < J; ++j) {   // a fresh HypothesisStatistic always samples from its hypothesis (:70-78)
          const uint32_t h = hyp[j];
          const uint32_t pos = leaf_pos[h] + pol_off;
          if (pos >= hp.pol_len) { failed = true; break; }
          const uint32_t e = __ldg(hp.gap_tab + (size_t)h * hp.pol_len + pos);
          leaf_pos[h] += e >> 8;        // the rollout's state copies carry the generator on, never the node's
          act[j + 1] = (int32_t)(e & 0xFFu);
        }
        if (failed) break;
        // every policy reads the PRE-step state (the joint action is complete before execute, random_heuristic.h:57-70)
        {
          int32_t v[A];
#pragma unroll
          for (int j = 1; j < A; ++j)
            v[j] = crossing_policy_action(p.env, st.x[j], st.last[j], st.x[0], st.last[0], hp.gap_lo[hyp[j - 1]] + act[j]);
#pragma unroll
          for (int j = 1; j < A; ++j) act[j] = v[j];
        }
        double r_ego, r_other, c;
        Env::step(p.env, st, act, r_ego, r_other, c);
        if (r_ego != 0.0) ego = __dadd_rn(ego, __dmul_rn(__ldg(p.disc + it), r_ego));   // random_heuristic.h:71-72
        if (c != 0.0) cost = __dadd_rn(cost, c);                                         // :79-84
        it += 1;
        fin = Env::terminal(p.env, st) | (it >= limit);
      }
      if (failed) { atomicAdd(p.counters + 6, 1ull); continue; }
      rollout_steps += it;
      const double prob = __ldg(p.disc + p.disc_len + it);
      const double ego_h = __dmul_rn(ego, prob);                                         // :95
      const double cost_h = __dmul_rn(cost, prob);                                       // :96-101
      // UctStatistic::update_from_heuristic (uct_statistic.h:100-110)
      leaf->ego.V = ego_h; leaf->ego.G = ego_h; leaf->ego.N = 1;
      G_ego = ego_h;
      // HypothesisStatistic::set_heuristic_estimate (:158-161: mean of the one-entry cost vector) and
      // update_from_heuristic (:100-110): the visit is counted under HYPOTHESIS_ID_NOT_SET (a new node never chose)
      const double hc = __ddiv_rn(__dadd_rn(0.0, cost_h), 1.0);
#pragma unroll
      for (int j = 0; j < J; ++j) {
        leaf->com[j].cost_value = hc; leaf->com[j].latest = hc;
        leaf->com[j].visits_notset += 1; leaf->com[j].N += 1;
        L[j] = hc;
      }
    } else {
      G_ego = leaf->ego.G;
#pragma unroll
      for (int j = 0; j < J; ++j) L[j] = leaf->com[j].latest;
    }
    // ---- back-propagation: mcts.h:314-329, stage_node.h:259-271 -------------------------------------------
    backup_levels += depth;
    uint32_t level = depth, child = node;
    // the path's records were read a few microseconds ago; ask for them again in one batch in case the L1 let go
    for (uint32_t l = 0; l < depth; ++l) hyp_prefetch_node<Node, J>(nodes + path_node[l], hyp);
#pragma unroll 1
    while (level > 0) {
      level -= 1;
      Node* pn = nodes + path_node[level];
      const Node* cn = nodes + child;
      // every value the level reads, requested before anything is computed or stored: the loads are independent
      // of each other, so the level costs one memory round trip instead of one per field (stores to the same
      // record with run-time indices in between would otherwise keep the compiler from hoisting them)
      const double r0 = cn->r_in, cw = cn->c_in;
      uint32_t cnt0[J], nh0[J], N0[J], nt0[J];
      double ac0[J], cv0[J];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const HypHyp& o = pn->hyp[j][hyp[j]];
        const HypCommon& cm = pn->com[j];
        const uint32_t a = path_ja[level][j + 1];
        cnt0[j] = o.count[a]; ac0[j] = o.cost[a]; nh0[j] = o.visits; cv0[j] = cm.cost_value;
        N0[j] = cm.N; nt0[j] = cm.nexp_total;
      }
      G_ego = back_apply<NE>(&pn->ego, path_ja[level][0], path_back[level], r0, p.gamma, G_ego);
#pragma unroll
      for (int j = 0; j < J; ++j) {   // HypothesisStatistic::update_statistic, :112-134
        HypHyp& o = pn->hyp[j][hyp[j]];
        HypCommon& cm = pn->com[j];
        const uint32_t a = path_ja[level][j + 1];
        const double Lc = L[j];
        const double Ln = hp.chance ? (cw < Lc ? Lc : cw) : __dadd_rn(cw, __dmul_rn(p.gamma, Lc));
        cm.latest = Ln;
        const uint32_t cnt = cnt0[j] + 1;
        o.count[a] = cnt;
        o.cost[a] = __dadd_rn(ac0[j], __ddiv_rn(__dsub_rn(Ln, ac0[j]), (double)cnt));
        const uint32_t nh = nh0[j] + 1;
        o.visits = nh;
        cm.cost_value = __dadd_rn(cv0[j], __ddiv_rn(__dsub_rn(Ln, cv0[j]), (double)nh));
        cm.N = N0[j] + 1;
        cm.nexp_total = nt0[j] + 1;
        L[j] = Ln;
      }
      child = path_node[level];
    }
    done += 1;
  }
  p.tree_iters[tree] = iter0 + done;
  p.tree_nodes[tree] = num_nodes;
  atomicAdd(p.counters + 0, (unsigned long long)done);
  atomicAdd(p.counters + 1, rollout_steps);
  atomicAdd(p.counters + 2, expand_steps);
  atomicAdd(p.counters + 3, (unsigned long long)(num_nodes - nodes_before));
  atomicAdd(p.counters + 4, backup_levels);
  atomicAdd(p.counters + 5, exact_evals);
}

template <int J, int NE>
__global__ void hyp_reset_kernel(void* nodes_v, uint32_t* tree_nodes, uint32_t* tree_iters, uint32_t T,
                                 uint32_t max_nodes, const int32_t* root_x, const int32_t* root_last) {
  using Node = HypNode<J, NE>;
  const uint32_t tree = blockIdx.x * blockDim.x + threadIdx.x;
  if (tree >= T) return;
  Node* root = static_cast<Node*>(nodes_v) + (size_t)tree * max_nodes;
  uint4* w = reinterpret_cast<uint4*>(root);
  for (size_t i = 0; i < offsetof(Node, hyp) / 16; ++i) w[i] = make_uint4(0u, 0u, 0u, 0u);   // see HypNode
  root->parent = kNoParent;
  for (int i = 0; i <= J; ++i) { root->x[i] = root_x[i]; root->last[i] = root_last[i]; }
  tree_nodes[tree] = 1;
  tree_iters[tree] = 0;
}

// ---- host side -----------------------------------------------------------------------------------------------

// A uniform random bit generator that replays words of a recorded mt19937 stream from a given position and
// counts how many it handed out: the std:: distributions run unchanged on top of it.
struct ReplayUrng {
  using result_type = std::mt19937::result_type;
  static constexpr result_type min() { return std::mt19937::min(); }
  static constexpr result_type max() { return std::mt19937::max(); }
  const std::vector<result_type>* words;
  size_t pos;
  result_type operator()() { return (*words)[pos++]; }
};

static std::vector<std::mt19937::result_type> mt_stream(uint32_t seed, size_t n) {
  std::mt19937 gen(seed);
  std::vector<std::mt19937::result_type> w(n);
  for (auto& x : w) x = gen();
  return w;
}

template <int J, int NE>
struct SearchH final : public mamcts_search {
  using Node = HypNode<J, NE>;
  mamcts_hyp_search_config hcfg;
  HypParams hp;
  uint8_t* d_seq = nullptr;
  uint16_t* d_gap = nullptr;
  uint16_t* d_sel = nullptr;
  int32_t* d_root_last = nullptr;

  size_t node_bytes() const override { return sizeof(Node); }
  bool uses_hash() const override { return true; }
  void free_extra() override { cudaFree(d_seq); cudaFree(d_gap); cudaFree(d_sel); cudaFree(d_root_last); }

  int upload_root(const int32_t* root_x, const int32_t* root_last) override {
    int32_t rx[MAMCTS_MAX_AGENTS] = {0}, rl[MAMCTS_MAX_AGENTS] = {0};
    for (int i = 0; i <= J; ++i) { rx[i] = root_x[i]; rl[i] = root_last ? root_last[i] : 2; }
    MAMCTS_CUDA(e, cudaMemcpyAsync(d_root_x, rx, sizeof(rx), cudaMemcpyHostToDevice, e->stream));
    MAMCTS_CUDA(e, cudaMemcpyAsync(d_root_last, rl, sizeof(rl), cudaMemcpyHostToDevice, e->stream));
    MAMCTS_CUDA(e, cudaStreamSynchronize(e->stream));
    return MAMCTS_OK;
  }
  int upload_sequence(const uint8_t* seq) {
    MAMCTS_CUDA(e, cudaMemcpyAsync(d_seq, seq, (size_t)hp.seq_len * J, cudaMemcpyHostToDevice, e->stream));
    MAMCTS_CUDA(e, cudaStreamSynchronize(e->stream));
    return MAMCTS_OK;
  }
  int health() override {
    unsigned long long h[2] = {0, 0};
    MAMCTS_CUDA(e, cudaMemcpyAsync(h, d_counters + 6, sizeof(h), cudaMemcpyDeviceToHost, e->stream));
    MAMCTS_CUDA(e, cudaStreamSynchronize(e->stream));
    if (h[0]) return fail(e, MAMCTS_ERR_NOMEM, "mamcts_hyp_search: " + std::to_string(h[0]) + " tree(s) ran past the tabulated generator positions");
    if (h[1]) return fail(e, MAMCTS_ERR_UNSUPPORTED, "mamcts_hyp_search: " + std::to_string(h[1]) + " tree(s) grew deeper than 96 levels");
    return MAMCTS_OK;
  }
  int launch_reset() override {
    const uint32_t T = cfg.num_trees;
    MAMCTS_CUDA(e, cudaMemsetAsync(d_hash, 0, sizeof(uint32_t) * (size_t)T * (p.hash_mask + 1), e->stream));
    MAMCTS_CUDA(e, cudaMemsetAsync(d_counters, 0, sizeof(unsigned long long) * 8, e->stream));
    hyp_reset_kernel<J, NE><<<(T + 255) / 256, 256, 0, e->stream>>>(d_nodes, d_tree_nodes, d_tree_iters, T, p.max_nodes,
                                                                    d_root_x, d_root_last);
    e->launches++;
    MAMCTS_CUDA(e, cudaGetLastError());
    iters_done = 0;
    return MAMCTS_OK;
  }
  int launch_run(uint32_t iterations) override {
    SearchParams q = p;
    q.iterations = iterations;
    const uint32_t T = cfg.num_trees;
    const uint32_t trees_per_block = (uint32_t)(block_threads / 32) * q.lanes_per_warp;
    const int grid = (int)((T + trees_per_block - 1) / trees_per_block);
    hyp_search_kernel<J, NE><<<grid, block_threads, 0, e->stream>>>(q, hp);
    e->launches++;
    MAMCTS_CUDA(e, cudaGetLastError());
    return MAMCTS_OK;
  }
  void read_header(const void* node, uint32_t* parent, uint32_t* depth, uint32_t* visits, uint8_t* ja) const override {
    const Node* nd = static_cast<const Node*>(node);
    *parent = nd->parent; *depth = nd->depth; *visits = nd->visit_count;
    for (int a = 0; a <= J; ++a) ja[a] = nd->ja[a];
  }
  void read_stat(const void* node, uint32_t agent, double* value, double* latest, uint32_t* visits, uint32_t* nexp,
                 double* q, uint32_t* n) const override {
    const Node* nd = static_cast<const Node*>(node);
    (void)agent;   // the ego agent's UctStatistic; the other agents' statistics are in the dump
    *value = nd->ego.V; *latest = nd->ego.G; *visits = nd->ego.N; *nexp = nd->ego.nexp;
    for (int a = 0; a < NE; ++a) { q[a] = nd->ego.Q[a]; n[a] = nd->ego.n[a]; }
  }
  void root_gather(RootGather* g) const override {
    const char* base = static_cast<const char*>(d_nodes);
    const size_t tree_bytes = sizeof(Node) * (size_t)p.max_nodes;
    g->q = reinterpret_cast<const double*>(base + offsetof(Node, ego) + offsetof(StatRec<NE>, Q));
    g->n = base + offsetof(Node, ego) + offsetof(StatRec<NE>, n);
    g->visits = base + offsetof(Node, ego) + offsetof(StatRec<NE>, N);
    g->count_bytes = 4;
    g->root_slot = nullptr;
    g->stride_q = tree_bytes / sizeof(double);
    g->stride_n = tree_bytes / sizeof(uint32_t);
    g->stride_v = tree_bytes / sizeof(uint32_t);
    g->max_actions = NE;
  }
  // canonical dump: include/mamcts_b200.h (mamcts_hyp_search_create)
  int dump(uint32_t tree, double* records, uint32_t capacity_records, uint32_t* num_records) override {
    uint32_t nn = 0;
    MAMCTS_CUDA(e, cudaMemcpyAsync(&nn, d_tree_nodes + tree, sizeof(nn), cudaMemcpyDeviceToHost, e->stream));
    MAMCTS_CUDA(e, cudaStreamSynchronize(e->stream));
    std::vector<Node> arena(nn);
    MAMCTS_CUDA(e, cudaMemcpyAsync(arena.data(), static_cast<const char*>(d_nodes) + sizeof(Node) * (size_t)tree * p.max_nodes,
                                   sizeof(Node) * nn, cudaMemcpyDeviceToHost, e->stream));
    MAMCTS_CUDA(e, cudaStreamSynchronize(e->stream));
    const uint32_t H = hp.H, MA = hp.MA, nE = p.n_ego;
    const uint32_t M = std::max(nE, MA);
    const uint32_t per_other = 6 + H * (2 + MA) + 2 * H * MA;
    const uint32_t stride = 4 + (3 + 2 * nE) + J * per_other;
    std::vector<double> code(nn, -1.0);
    std::vector<std::vector<uint32_t>> kids(nn);
    for (uint32_t i = 1; i < nn; ++i) {
      double c = arena[i].ja[0], mul = M;
      for (int j = 1; j <= J; ++j) { c += mul * arena[i].ja[j]; mul *= M; }
      code[i] = c;
      kids[arena[i].parent].push_back(i);
    }
    for (auto& k : kids) std::sort(k.begin(), k.end(), [&](uint32_t a, uint32_t b) { return code[a] < code[b]; });
    uint32_t count = 0;
    auto emit = [&](uint32_t i) {
      if (count < capacity_records) {
        const Node& nd = arena[i];
        double* r = records + (size_t)count * stride;
        for (uint32_t k = 0; k < stride; ++k) r[k] = 0.0;
        r[0] = nd.depth; r[1] = code[i]; r[2] = nd.visit_count; r[3] = (double)kids[i].size();
        double* eo = r + 4;
        eo[0] = nd.ego.N; eo[1] = nd.ego.V; eo[2] = nd.ego.G;
        const uint32_t mask = tab.exp_mask[0][std::min<uint32_t>(nd.ego.nexp, nE)];
        for (uint32_t k = 0; k < nE; ++k) {
          const bool ex = (mask >> k) & 1u;
          eo[3 + k] = ex ? (double)nd.ego.n[k] : -1.0;
          eo[3 + nE + k] = ex ? nd.ego.Q[k] : 0.0;
        }
        for (int j = 0; j < J; ++j) {
          const HypCommon& o = nd.com[j];
          double* oo = r + 4 + (3 + 2 * nE) + j * per_other;
          oo[0] = o.N; oo[1] = o.cost_value; oo[2] = o.latest; oo[3] = o.nexp_total; oo[4] = o.visits_notset; oo[5] = 0.0;
          for (uint32_t h = 0; h < H; ++h) {
            // a hypothesis the agent never chose under at this node: its line was never initialised - nothing in it
            HypHyp empty;
            std::memset(&empty, 0, sizeof(empty));
            const HypHyp& q = o.seen[h] ? nd.hyp[j][h] : empty;
            double* oh = oo + 6 + h * (2 + MA);
            oh[0] = o.seen[h] ? (double)q.visits : -1.0;
            oh[1] = o.seen[h] ? (double)q.nexp : -1.0;
            for (uint32_t k = 0; k < MA; ++k) oh[2 + k] = k < q.nexp ? (double)q.order[q.nexp - 1u - k] : -1.0;
            double* oc = oo + 6 + H * (2 + MA) + 2 * h * MA;
            for (uint32_t k = 0; k < MA; ++k) {
              const bool pr = (q.present >> k) & 1u;
              oc[2 * k] = pr ? (double)q.count[k] : -1.0;
              oc[2 * k + 1] = pr ? q.cost[k] : 0.0;
            }
          }
        }
      }
      ++count;
    };
    std::vector<std::pair<uint32_t, size_t>> stack;
    stack.push_back({0u, 0});
    emit(0);
    while (!stack.empty()) {
      auto& top = stack.back();
      if (top.second < kids[top.first].size()) {
        const uint32_t c = kids[top.first][top.second++];
        emit(c);
        stack.push_back({c, 0});
      } else {
        stack.pop_back();
      }
    }
    *num_records = count;
    return MAMCTS_OK;
  }
};

static mamcts_search* make_hyp_search(uint32_t J, uint32_t n_ego) {
  if (n_ego <= 4) {
    switch (J) {
      case 1: return new SearchH<1, 4>();
      case 2: return new SearchH<2, 4>();
      case 3: return new SearchH<3, 4>();
      default: return nullptr;
    }
  }
  if (n_ego <= 8) {
    switch (J) {
      case 1: return new SearchH<1, 8>();
      case 2: return new SearchH<2, 8>();
      case 3: return new SearchH<3, 8>();
      default: return nullptr;
    }
  }
  return nullptr;
}

template <int J, int NE>
static SearchH<J, NE>* as_hyp(mamcts_search* s) { return dynamic_cast<SearchH<J, NE>*>(s); }

// calls f(SearchH<J,NE>*) for the concrete type of s; false if s is not a hypothesis search
template <class F>
static bool with_hyp(mamcts_search* s, F&& f) {
#define MAMCTS_TRY(JJ, EE) if (auto* h = as_hyp<JJ, EE>(s)) { f(h); return true; }
  MAMCTS_TRY(1, 4) MAMCTS_TRY(2, 4) MAMCTS_TRY(3, 4) MAMCTS_TRY(1, 8) MAMCTS_TRY(2, 8) MAMCTS_TRY(3, 8)
#undef MAMCTS_TRY
  return false;
}

}  // namespace mamcts

using namespace mamcts;

extern "C" {

void mamcts_default_hypothesis_params(mamcts_hypothesis_params* p) {
  if (!p) return;
  // mcts_default_parameters(), mcts/mcts_parameters.h:118-124; default_crossing_state_parameters: policy seed 1000
  p->upper_cost_bound = 100.0;
  p->lower_cost_bound = 0.0;
  p->pw_k = 1.0;
  p->pw_alpha = 0.5;
  p->exploration_constant = 0.7;
  p->cost_based_action_selection = 0;
  p->progressive_widening_hypothesis_based = 1;
  p->chance_constrained_update = 0;
  p->policy_random_seed = 1000;
}

int mamcts_reference_hypothesis_sequence(uint32_t seed, const float* beliefs, uint32_t num_others,
                                         uint32_t num_hypotheses, uint32_t skip_iterations, uint32_t n_iterations,
                                         uint8_t* out) {
  if (!beliefs || !out || num_others == 0 || num_others >= MAMCTS_MAX_AGENTS || num_hypotheses == 0 ||
      num_hypotheses > MAMCTS_MAX_HYPOTHESES)
    return fail(nullptr, MAMCTS_ERR_INVALID, "mamcts_reference_hypothesis_sequence: bad argument");
  // HypothesisBeliefTracker::sample_current_hypothesis (hypothesis_belief_tracker.h:136-149): the SAME containers
  // (an unordered_map keyed by the agent id, filled in get_other_agent_idx() order by belief_update) and the
  // same std::discrete_distribution, so the agents draw in the order the reference's map yields them
  std::unordered_map<unsigned, std::vector<float>> tracked;
  for (uint32_t j = 0; j < num_others; ++j) {
    auto& b = tracked[j + 1];
    b.assign(beliefs + (size_t)j * num_hypotheses, beliefs + (size_t)(j + 1) * num_hypotheses);
  }
  std::mt19937 gen(seed);
  for (uint32_t it = 0; it < skip_iterations + n_iterations; ++it) {
    for (const auto& kv : tracked) {
      std::discrete_distribution<unsigned> dist(kv.second.begin(), kv.second.end());
      const unsigned h = dist(gen);
      if (it >= skip_iterations) out[(size_t)(it - skip_iterations) * num_others + (kv.first - 1)] = (uint8_t)h;
    }
  }
  return MAMCTS_OK;
}

int mamcts_hyp_search_create(mamcts_engine* e, const mamcts_hyp_search_config* cfg, const int32_t* root_x,
                             const int32_t* root_last_action, const uint8_t* hypothesis_sequence,
                             mamcts_search** out) {
  if (!e) return fail(nullptr, MAMCTS_ERR_INVALID, "mamcts_hyp_search_create: null engine");
  std::lock_guard<std::mutex> lock(e->mu);
  if (!cfg || !root_x || !hypothesis_sequence || !out) return fail(e, MAMCTS_ERR_INVALID, "mamcts_hyp_search_create: null argument");
  const mamcts_params& mp = cfg->mcts;
  const mamcts_hypothesis_params& hpar = cfg->hypothesis;
  const mamcts_crossing_params& cp = cfg->crossing;
  if (cfg->num_trees == 0 || mp.max_number_of_iterations == 0) return fail(e, MAMCTS_ERR_INVALID, "mamcts_hyp_search_create: no trees / iterations");
  if (cfg->num_hypotheses == 0 || cfg->num_hypotheses > MAMCTS_MAX_HYPOTHESES)
    return fail(e, MAMCTS_ERR_UNSUPPORTED, "mamcts_hyp_search_create: 1..4 hypotheses");
  if (!hpar.progressive_widening_hypothesis_based)
    return fail(e, MAMCTS_ERR_UNSUPPORTED, "mamcts_hyp_search_create: built for PROGRESSIVE_WIDENING_HYPOTHESIS_BASED = true (the reference's default)");
  if (!(hpar.pw_k >= 0.0) || !(hpar.pw_alpha >= 0.0) || !(mp.uct_pw_k >= 0.0) || !(mp.uct_pw_alpha >= 0.0))
    return fail(e, MAMCTS_ERR_INVALID, "mamcts_hyp_search_create: progressive widening k, alpha must be >= 0");
  const uint32_t J = cp.num_other_agents;
  const uint32_t n_ego = (uint32_t)(cp.max_velocity_ego - cp.min_velocity_ego + 1);
  const int64_t ma64 = (int64_t)cp.max_velocity_other - cp.min_velocity_other + 1;
  if (ma64 < 1 || ma64 > MAMCTS_HYP_MAX_OTHER_ACTIONS)
    return fail(e, MAMCTS_ERR_UNSUPPORTED, "mamcts_hyp_search_create: at most 8 other-agent velocities");
  const uint32_t MA = (uint32_t)ma64;
  {
    // the action keys of the reference's std::unordered_map<ActionIdx, UcbPair> are the velocities bit-cast into a
    // size_t (crossing_state_common.h:19-27: negative v -> 2^32 + v); with every key in a bucket of its own
    // (13 buckets until a 14th element) the container iterates in reverse insertion order, what the kernel assumes
    bool used[13] = {false};
    for (int32_t v = cp.min_velocity_other; v <= cp.max_velocity_other; ++v) {
      const uint64_t key = v >= 0 ? (uint64_t)v : (uint64_t)(uint32_t)v;
      const uint32_t b = (uint32_t)(key % 13u);
      if (used[b]) return fail(e, MAMCTS_ERR_UNSUPPORTED, "mamcts_hyp_search_create: two other-agent velocities share a bucket of the reference's action map; its iteration order is not reproduced for this velocity range");
      used[b] = true;
    }
  }
  for (uint32_t h = 0; h < cfg->num_hypotheses; ++h)
    if (cfg->gap_lo[h] > cfg->gap_hi[h] || (int64_t)cfg->gap_hi[h] - cfg->gap_lo[h] > 200)
      return fail(e, MAMCTS_ERR_INVALID, "mamcts_hyp_search_create: bad desired-gap range");
  mamcts_search* s = make_hyp_search(J, n_ego);
  if (!s) return fail(e, MAMCTS_ERR_UNSUPPORTED, "mamcts_hyp_search_create: compiled for 1, 2 or 3 other agents and up to 8 ego actions");
  s->e = e;
  std::memset(&s->cfg, 0, sizeof(s->cfg));
  s->cfg.env = MAMCTS_ENV_CROSSING_I32;
  s->cfg.num_trees = cfg->num_trees;
  s->cfg.first_tree_id = cfg->first_tree_id;
  s->cfg.mcts = mp;
  s->cfg.crossing = cp;
  s->A = J + 1;
  s->state_ints = J + 1;
  {
    const cudaError_t derr = cudaSetDevice(e->device);
    if (derr != cudaSuccess) { delete s; return cuda_fail(e, derr, "mamcts_hyp_search_create: cudaSetDevice"); }
  }
  const uint32_t max_iter = mp.max_number_of_iterations;
  uint32_t max_nodes = cfg->max_nodes_per_tree;
  const uint64_t needed = std::min<uint64_t>((uint64_t)mp.max_number_of_nodes + 1, (uint64_t)max_iter + 1);
  if (max_nodes == 0) max_nodes = (uint32_t)needed;
  if (max_nodes < needed) { delete s; return fail(e, MAMCTS_ERR_INVALID, "mamcts_hyp_search_create: max_nodes_per_tree too small"); }

  // ---- host tables ----------------------------------------------------------------------------------------
  SearchTables& tab = s->tab;
  std::memset(&tab, 0, sizeof(tab));
  {
    uint32_t order[MAMCTS_MAX_ACTIONS];
    int rc = mamcts_reference_expand_order(mp.random_seed, n_ego, order);
    if (rc != MAMCTS_OK) { delete s; return fail(e, rc, mamcts_last_error(nullptr)); }
    uint32_t mask = 0;
    for (uint32_t k = 0; k < n_ego; ++k) {
      tab.order[0][k] = (uint8_t)order[k];
      mask |= 1u << order[k];
      tab.exp_mask[0][k + 1] = mask;
      uint32_t thr = 0xFFFFFFFFu;
      for (uint32_t N = 0; N <= max_iter + 1; ++N)
        if ((double)k <= mp.uct_pw_k * std::pow((double)N, mp.uct_pw_alpha)) { thr = N; break; }
      tab.pw_min_visits[0][k] = thr;
    }
  }
  HypTables htab;
  std::memset(&htab, 0, sizeof(htab));
  for (uint32_t k = 0; k <= MA; ++k) {   // require_progressive_widening_hypothesis_based (:270-276)
    uint32_t thr = 0xFFFFFFFFu;
    for (uint32_t N = 0; N <= max_iter + 1; ++N)
      if ((double)k <= hpar.pw_k * std::pow((double)N, hpar.pw_alpha)) { thr = N; break; }
    htab.pw_min_visits[k] = thr;
  }
  std::vector<double> logs((size_t)max_iter + 2);
  for (size_t N = 0; N < logs.size(); ++N) logs[N] = std::log((double)N);
  const uint32_t max_steps = std::min(mp.rh_max_number_of_iterations, mp.max_search_depth);
  if (max_steps > (1u << 20)) { delete s; return fail(e, MAMCTS_ERR_UNSUPPORTED, "mamcts_hyp_search_create: rollout cap above 2^20"); }
  // positions: a policy generator advances by the draws of the agents that share its hypothesis - along a node's
  // lineage at most one per agent and visit, plus one rollout; a statistic's by at most one word per visit
  // (a few more where a distribution rejects)
  const uint32_t off_room = cfg->decorrelate_trees ? 4096u : 0u;
  const uint32_t pol_len = max_iter * J + max_steps * J + 256u + off_room;
  const uint32_t sel_len = max_iter + 256u + off_room;
  std::vector<uint16_t> gap_tab((size_t)cfg->num_hypotheses * pol_len), sel_tab((size_t)kHypMA * sel_len, 0);
  {
    const auto words = mt_stream(hpar.policy_random_seed, (size_t)pol_len + 64);
    for (uint32_t h = 0; h < cfg->num_hypotheses; ++h)
      for (uint32_t pos = 0; pos < pol_len; ++pos) {
        ReplayUrng g{&words, pos};
        std::uniform_int_distribution<int> dis(cfg->gap_lo[h], cfg->gap_hi[h]);   // agent_policy.h:70-71
        const int gap = dis(g);
        const size_t used = std::min<size_t>(g.pos - pos, 255);
        gap_tab[(size_t)h * pol_len + pos] = (uint16_t)((uint32_t)(gap - cfg->gap_lo[h]) | (uint32_t)(used << 8));
      }
  }
  {
    const auto words = mt_stream(mp.random_seed, (size_t)sel_len + 64);
    for (uint32_t sz = 1; sz <= MA; ++sz)
      for (uint32_t pos = 0; pos < sel_len; ++pos) {
        ReplayUrng g{&words, pos};
        std::uniform_int_distribution<size_t> pick(0, sz - 1);   // hypothesis_statistic.h:209
        const size_t idx = pick(g);
        const size_t used = std::min<size_t>(g.pos - pos, 255);
        sel_tab[(size_t)(sz - 1) * sel_len + pos] = (uint16_t)((uint32_t)idx | (uint32_t)(used << 8));
      }
  }
  uint32_t hcap = 16;
  while (hcap < 2 * max_nodes) hcap <<= 1;
  const uint32_t T = cfg->num_trees;

  cudaError_t err = cudaSuccess;
  auto alloc = [&](void** ptr, size_t bytes) { if (err == cudaSuccess) err = cudaMalloc(ptr, bytes); };
  uint8_t* d_seq = nullptr; uint16_t *d_gap = nullptr, *d_sel = nullptr; int32_t* d_root_last = nullptr;
  alloc(&s->d_nodes, s->node_bytes() * (size_t)T * max_nodes);
  alloc((void**)&s->d_hash, sizeof(uint32_t) * (size_t)T * hcap);
  alloc((void**)&s->d_tree_nodes, sizeof(uint32_t) * T);
  alloc((void**)&s->d_tree_iters, sizeof(uint32_t) * T);
  alloc((void**)&s->d_tab, sizeof(SearchTables));
  alloc((void**)&s->d_log, sizeof(double) * logs.size());
  alloc((void**)&s->d_disc, sizeof(double) * 2 * ((size_t)max_steps + 1));
  alloc((void**)&s->d_counters, sizeof(unsigned long long) * 8);
  alloc((void**)&s->d_root_x, sizeof(int32_t) * MAMCTS_MAX_AGENTS);
  alloc((void**)&d_root_last, sizeof(int32_t) * MAMCTS_MAX_AGENTS);
  alloc((void**)&d_seq, (size_t)max_iter * J);
  alloc((void**)&d_gap, sizeof(uint16_t) * gap_tab.size());
  alloc((void**)&d_sel, sizeof(uint16_t) * sel_tab.size());
  HypParams hp;
  std::memset(&hp, 0, sizeof(hp));
  hp.nodes = s->d_nodes; hp.hash = s->d_hash; hp.seq = d_seq; hp.gap_tab = d_gap; hp.sel_tab = d_sel;
  hp.seq_len = max_iter; hp.pol_len = pol_len; hp.sel_len = sel_len;
  hp.H = cfg->num_hypotheses; hp.MA = MA; hp.n_oth_actions = cp.num_other_actions;
  for (uint32_t h = 0; h < cfg->num_hypotheses; ++h) hp.gap_lo[h] = cfg->gap_lo[h];
  hp.min_v_other = cp.min_velocity_other;
  hp.cost_based = hpar.cost_based_action_selection; hp.chance = hpar.chance_constrained_update;
  hp.decorrelate = cfg->decorrelate_trees;
  hp.hyp_two_c = 2 * hpar.exploration_constant; hp.hyp_lb = hpar.lower_cost_bound; hp.hyp_ub = hpar.upper_cost_bound;
  hp.tab = htab;
  const bool typed = with_hyp(s, [&](auto* h) {
    h->hcfg = *cfg; h->hp = hp; h->d_seq = d_seq; h->d_gap = d_gap; h->d_sel = d_sel; h->d_root_last = d_root_last;
  });
  if (err != cudaSuccess || !typed) {
    if (!typed) { cudaFree(d_seq); cudaFree(d_gap); cudaFree(d_sel); cudaFree(d_root_last); }
    free_search(s);
    cudaGetLastError();
    return fail(e, MAMCTS_ERR_NOMEM, std::string("mamcts_hyp_search_create: device allocation failed: ") + cudaGetErrorString(err));
  }
  SearchParams& p = s->p;
  std::memset(&p, 0, sizeof(p));
  p.nodes = s->d_nodes; p.hash = s->d_hash; p.tree_nodes = s->d_tree_nodes; p.tree_iters = s->d_tree_iters;
  p.tab = s->d_tab; p.log_table = s->d_log; p.counters = s->d_counters;
  p.disc = s->d_disc; p.disc_len = max_steps + 1;
  p.T = T; p.max_nodes = max_nodes; p.hash_mask = hcap - 1; p.first_tree_id = cfg->first_tree_id;
  p.n_ego = n_ego; p.n_oth = cp.num_other_actions;
  p.use_bound_est = mp.use_bound_estimation; p.max_depth = mp.max_search_depth;
  p.node_budget = mp.max_number_of_nodes; p.rh_cap = mp.rh_max_number_of_iterations;
  p.rollouts_per_leaf = 1;
  p.gamma = mp.discount_factor; p.two_c = 2 * mp.uct_exploration_constant;
  p.lb = mp.uct_lower_bound; p.ub = mp.uct_upper_bound;
  make_crossing_env(cp, p.env);
  {
    const uint64_t warp_slots = (uint64_t)std::max(e->num_sms, 1) * (2u * MAMCTS_HYP_MIN_BLOCKS);
    uint32_t lpw = 1;
    while (lpw < 32 && (uint64_t)lpw * warp_slots < (uint64_t)T) lpw <<= 1;
    p.lanes_per_warp = lpw;
    if (const char* v = std::getenv("MAMCTS_SEARCH_LPW")) {   // profiling knob, as for the UCT searches
      const int l = std::atoi(v);
      if (l >= 1 && l <= 32) p.lanes_per_warp = (uint32_t)l;
    }
  }
  err = cudaMemcpyAsync(s->d_tab, &tab, sizeof(tab), cudaMemcpyHostToDevice, e->stream);
  if (err == cudaSuccess) err = cudaMemcpyAsync(s->d_log, logs.data(), sizeof(double) * logs.size(), cudaMemcpyHostToDevice, e->stream);
  if (err == cudaSuccess) err = cudaMemcpyAsync(d_gap, gap_tab.data(), sizeof(uint16_t) * gap_tab.size(), cudaMemcpyHostToDevice, e->stream);
  if (err == cudaSuccess) err = cudaMemcpyAsync(d_sel, sel_tab.data(), sizeof(uint16_t) * sel_tab.size(), cudaMemcpyHostToDevice, e->stream);
  if (err == cudaSuccess) err = cudaMemcpyAsync(d_seq, hypothesis_sequence, (size_t)max_iter * J, cudaMemcpyHostToDevice, e->stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
  if (err != cudaSuccess) { free_search(s); return cuda_fail(e, err, "mamcts_hyp_search_create uploads"); }
  int rc = s->upload_root(root_x, root_last_action);
  if (rc == MAMCTS_OK) rc = fill_discount_table(e, mp.discount_factor, 1.0 / (double)n_ego, max_steps, s->d_disc);
  if (rc == MAMCTS_OK) rc = s->launch_reset();
  if (rc != MAMCTS_OK) { free_search(s); return rc; }
  *out = s;
  return MAMCTS_OK;
}

int mamcts_hyp_search_reset(mamcts_search* s, const int32_t* root_x, const int32_t* root_last_action,
                            const uint8_t* hypothesis_sequence) {
  if (!s) return fail(nullptr, MAMCTS_ERR_INVALID, "mamcts_hyp_search_reset: null handle");
  mamcts_engine* e = s->e;
  std::lock_guard<std::mutex> lock(e->mu);
  MAMCTS_CUDA(e, cudaSetDevice(e->device));
  int rc = MAMCTS_OK;
  const bool typed = with_hyp(s, [&](auto* h) {
    if (hypothesis_sequence) rc = h->upload_sequence(hypothesis_sequence);
  });
  if (!typed) return fail(e, MAMCTS_ERR_INVALID, "mamcts_hyp_search_reset: not a hypothesis search");
  if (rc != MAMCTS_OK) return rc;
  if (root_x) {
    rc = s->upload_root(root_x, root_last_action);
    if (rc != MAMCTS_OK) return rc;
  }
  return s->launch_reset();
}

}  // extern "C"
