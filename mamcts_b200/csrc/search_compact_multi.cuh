// search_compact_multi.cuh - two-tier ("compact") tree layout of the device-resident search for MORE THAN TWO
// agents whose other agents never receive a reward (CrossingState: rewards[1..] stay 0, reference
// environments/crossing_state.h:145, and a discount >= 0).
//
// Why: the classic layout (search_kernel.cuh) spends ~950 B per node at 8 agents - every other agent carries
// V, G and 7 action values that are +0.0 for the whole search - so 16 384 ten-thousand-iteration trees fill a
// B200 and the search runs at the latency of one dependent 950-B record per level with a sixth of the trees a
// full wave needs. What an other agent's statistic really holds is its visit counts:
//
//   leaf   8 B   the ego's heuristic value h (N = 1, V = G = h implied; the others' values are +0.0). A node
//                stays a leaf until the tree policy selects it a second time; its state is recomputed by
//                carrying the state down from the root (one execute() per level), not stored.
//   inner  64-B aligned record: N, visit_count, #expanded per agent, n[a] of the ego and of every other agent
//                (16- or 32-bit), then V, G and the action values of the ego, the parent and the joint action.
//                8 agents, 4 / 7 actions, 16-bit counts: 192 B.
//   children     one open-addressing table per tree keyed by (parent inner index, joint-action code); a slot is
//                one 64-bit word {key:48, entry:16}; entry (L + 1) << 1 = leaf L, (I << 1) | 1 = inner node I.
//
// Results are bit-identical to the classic layout and to the oracle (tests/test_gpu_search.py): the same
// selections (the ego's UCB through select_action, the others' integer first-minimum of the counts - exact
// for equal action values, search_kernel.cuh), the same Philox streams (stream_lo = creation index of the
// node), the same FP64 arithmetic for the ego in the same order.
#pragma once
#include <cstddef>

#include "search_kernel.cuh"

namespace mamcts {

constexpr int kPathMaxMulti = 64;

template <class Env, int NE, int NA, class CNT>
struct alignas(64) InnerRecM {
  static constexpr int A = Env::A;
  static constexpr int NO = A - 1;
  CNT N;                     // total_node_visits_ (the same for every agent: they are updated together)
  CNT visit_count;           // StageNode::visit_count_
  uint8_t nexp_e;            // ucb_statistics_.size() of the ego: the first nexp entries of the widening order
  uint8_t nexp_o[NO];
  CNT n_e[NE];               // action_count_
  CNT n_o[NO][NA];
  alignas(16) double vg_e[2];   // value_, latest_return_ of the ego
  double q_e[NE];            // action_value_ of the ego
  uint32_t parent;           // inner index of the parent (export only)
  uint8_t ja[A];             // joint action leading here (export only)
};

struct CompactMultiParams {
  void* inner;               // [T][max_inner]
  double* leaves;            // [T][max_nodes] h of the ego
  unsigned long long* table; // [T][table_size] {key:48, entry:16}
  uint32_t* tree_inner;      // [T] inner records in use
  uint32_t* tree_leaves;     // [T] leaf records in use
  uint32_t max_inner, table_mask;
  const int32_t* root_x;
};

__device__ __forceinline__ uint32_t mix_key(unsigned long long key) {
  key ^= key >> 23;
  key *= 0x2127599BF4325C37ull;
  key ^= key >> 29;
  return (uint32_t)key;
}

template <class Env, int NE, int NA, class CNT, int KIND, bool MULTI = false>
__global__ void __launch_bounds__(kSearchThreads)
search_compact_multi_kernel(const __grid_constant__ SearchParams p, const __grid_constant__ CompactMultiParams c) {
  using Inner = InnerRecM<Env, NE, NA, CNT>;
  constexpr int A = Env::A;
  constexpr int NO = A - 1;
  static_assert(Env::kOtherRewardZero, "built for environments whose other agents are reward-free");
  __shared__ SearchTables tab;
  stage_tables(tab, p.tab);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t tree = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * p.lanes_per_warp + lane;
  const unsigned live = __ballot_sync(0xFFFFFFFFu, lane < p.lanes_per_warp && tree < p.T);
  if (lane >= p.lanes_per_warp || tree >= p.T) return;
  Inner* inner = static_cast<Inner*>(c.inner) + (size_t)tree * c.max_inner;
  double* leaves = c.leaves + (size_t)tree * p.max_nodes;
  unsigned long long* table = c.table + (size_t)tree * ((size_t)c.table_mask + 1);
  uint32_t num_nodes = p.tree_nodes[tree];
  uint32_t num_inner = c.tree_inner[tree];
  uint32_t num_leaves = c.tree_leaves[tree];
  unsigned long long rollout_steps = 0, expand_steps = 0, backup_levels = 0, exact_evals = 0;
  const uint32_t nodes_before = num_nodes;
  uint8_t const_act[A];
#pragma unroll
  for (int a = 0; a < A; ++a) const_act[a] = tab.order[a == 0 ? 0 : 1][0];   // F1: first draw
  typename Env::State root;
#pragma unroll
  for (int i = 0; i < A; ++i) { root.x[i] = __ldg(c.root_x + i); root.last[i] = 0; }
  root.flags = 0;
  // joint-action code: sum_a ja[a] * 8^a (NE, NA <= 8) - 24 bits for 8 agents
  uint32_t done = 0;
  bool failed = false;

  uint32_t path_node[kPathMaxMulti];
  uint8_t path_ja[kPathMaxMulti][A];
  BackRegs path_back[kPathMaxMulti];   // the ego's values as the descent saw them
  double path_r0[kPathMaxMulti];
  CNT path_cnt_o[kPathMaxMulti][NO];   // n[a] of every other agent's chosen action as the descent saw it: the
                                       // backup does no loads (a tree is touched by its lane only)

#pragma unroll 1
  for (uint32_t iter = 0; iter < p.iterations; ++iter) {
    uint32_t node = 0, depth = 0;
    typename Env::State st = root;
    bool do_rollout = false;
    uint32_t new_leaf = 0, creation = 0;
    double G = 0.0;
    // ---- select & expand: mcts.h:300-305, stage_node.h:167-232 ------------------------------------------
#pragma unroll 1
    while (!failed) {
      Inner* nd = inner + node;
      // one memory round trip for the whole record: every line of it is requested before the first use (the
      // loop over the agents below would otherwise fetch them one after the other)
#pragma unroll
      for (size_t off = 64; off < sizeof(Inner); off += 64) prefetch_l1(reinterpret_cast<const char*>(nd) + off);
      const uint32_t vc = nd->visit_count;
      nd->visit_count = (CNT)(vc + 1);                                          // :180
      if (Env::terminal(p.env, st) || depth == p.max_depth || num_nodes > p.node_budget) {   // :183-187
        G = nd->vg_e[1];
        break;
      }
      if (depth >= (uint32_t)kPathMaxMulti) { failed = true; atomicAdd(p.counters + 7, 1ull); break; }
      uint8_t ja[A];
      const uint32_t N = nd->N;
      {
        StatRegs<NE> ego;
        ego.N = N; ego.nexp = nd->nexp_e; ego.V = nd->vg_e[0];
#pragma unroll
        for (int a = 0; a < NE; ++a) { ego.Q[a] = nd->q_e[a]; ego.n[a] = nd->n_e[a]; }
        BackRegs back;
        bool w;
        ja[0] = (uint8_t)select_action<NE>(ego, p, tab, 0, p.n_ego, exact_evals, back, w);   // :190
        if (w) nd->nexp_e = (uint8_t)(ego.nexp + 1);
        path_back[depth] = back;
      }
#pragma unroll 1
      for (int j = 0; j < NO; ++j) {   // reward-free agents: every value is +0.0, the counts decide
        StatRegs<NA> oth;
        oth.N = N; oth.nexp = nd->nexp_o[j]; oth.V = 0.0;
#pragma unroll
        for (int a = 0; a < NA; ++a) { oth.Q[a] = 0.0; oth.n[a] = nd->n_o[j][a]; }
        BackRegs back;
        bool w;
        ja[j + 1] = (uint8_t)select_action<NA>(oth, p, tab, 1, p.n_oth, exact_evals, back, w, true);
        if (w) nd->nexp_o[j] = (uint8_t)(oth.nexp + 1);
        path_cnt_o[depth][j] = (CNT)back.cnt0;
      }
      path_node[depth] = node;
      uint32_t code = 0;
#pragma unroll
      for (int a = 0; a < A; ++a) { path_ja[depth][a] = ja[a]; code |= (uint32_t)ja[a] << (3 * a); }
      // child lookup, :198
      const unsigned long long key = ((unsigned long long)node << 24) | code;
      uint32_t slot = mix_key(key) & c.table_mask;
      uint32_t entry = 0;
      while (true) {
        const unsigned long long w = table[slot];
        if (w == 0ull) break;
        if ((w >> 16) == key) { entry = (uint32_t)(w & 0xFFFFull); break; }
        slot = (slot + 1) & c.table_mask;
      }
      if (entry & 1u) prefetch_l1(inner + (entry >> 1));   // the child's record travels while the state is stepped
      // the child's state and the reward collected on the edge (collect(), node_statistic.h:115-121)
      int32_t act[A];
#pragma unroll
      for (int a = 0; a < A; ++a) act[a] = ja[a];   // index-valued: others move by +index (F5)
      double r_ego, r_other, c0;
      Env::step(p.env, st, act, r_ego, r_other, c0);
      path_r0[depth] = r_ego;
      depth += 1;
      if (entry & 1u) { node = entry >> 1; continue; }                          // existing inner node, :199-206
      if (entry == 0u) {
        // expand, :207-231: the new node is a leaf record
        new_leaf = num_leaves++;
        creation = num_nodes++;
        expand_steps += 1;
        table[slot] = (key << 16) | (unsigned long long)((new_leaf + 1u) << 1);
        do_rollout = true;
        break;
      }
      // second selection of a leaf: promote it to an inner record and go on from there
      {
        const uint32_t leaf = (entry >> 1) - 1u;
        const double h = leaves[leaf];
        if (num_inner >= c.max_inner) { failed = true; atomicAdd(p.counters + 6, 1ull); break; }
        const uint32_t id = num_inner++;
        Inner* cn = inner + id;
        uint4* wz = reinterpret_cast<uint4*>(cn);
#pragma unroll
        for (size_t i = 0; i < sizeof(Inner) / 16; ++i) wz[i] = make_uint4(0u, 0u, 0u, 0u);
        cn->N = 1;                       // update_from_heuristic left 0 + 1 visits
        cn->vg_e[0] = h; cn->vg_e[1] = h;   // V = G = h
        cn->parent = node;
#pragma unroll
        for (int a = 0; a < A; ++a) cn->ja[a] = ja[a];
        table[slot] = (key << 16) | (unsigned long long)((id << 1) | 1u);
        node = id;
      }
    }
    __syncwarp(live);   // the lanes leave the descent through different exits: reconverge before the rollout
    if (failed) continue;
    // ---- heuristic: mcts.h:309-312, random_heuristic.h:106-134 ----------------------------------------------
    if (do_rollout) {
      double other = 0.0;
      uint32_t steps = 0;
      const double ego_h = leaf_heuristic<Env, KIND, MULTI>(p, st, depth, p.first_tree_id + tree, creation, const_act,
                                                            other, steps);
      rollout_steps += steps;
      leaves[new_leaf] = ego_h;       // update_from_heuristic (uct_statistic.h:100-110): V = G = h, N = 0 + 1
      G = ego_h;
    }
    // ---- back-propagation: mcts.h:314-329, stage_node.h:259-271 -------------------------------------------
    backup_levels += depth;
    uint32_t level = depth;
#pragma unroll 1
    while (level > 0) {
      level -= 1;
      Inner* pn = inner + path_node[level];
      const BackOut e = back_compute(path_back[level], path_r0[level], p.gamma, G);
      const uint32_t ae = path_ja[level][0];
      pn->N = (CNT)e.nv;
      pn->n_e[ae] = (CNT)e.cnt;
      pn->vg_e[0] = e.v;
      pn->vg_e[1] = e.G;
      if (__double_as_longlong(e.q) != __double_as_longlong(path_back[level].q)) pn->q_e[ae] = e.q;
#pragma unroll
      for (int j = 0; j < NO; ++j) {   // reward-free agents: only the counts move
        const uint32_t ao = path_ja[level][j + 1];
        pn->n_o[j][ao] = (CNT)(path_cnt_o[level][j] + 1);
      }
      G = e.G;
    }
    done += 1;
  }
  p.tree_nodes[tree] = num_nodes;
  c.tree_inner[tree] = num_inner;
  c.tree_leaves[tree] = num_leaves;
  p.tree_iters[tree] += done;
  atomicAdd(p.counters + 0, (unsigned long long)done);
  atomicAdd(p.counters + 1, rollout_steps);
  atomicAdd(p.counters + 2, expand_steps);
  atomicAdd(p.counters + 3, (unsigned long long)(num_nodes - nodes_before));
  atomicAdd(p.counters + 4, backup_levels);
  atomicAdd(p.counters + 5, exact_evals);
}

template <class Env, int NE, int NA, class CNT>
__global__ void search_compact_multi_reset_kernel(void* inner_v, uint32_t* tree_nodes, uint32_t* tree_iters,
                                                  uint32_t* tree_inner, uint32_t* tree_leaves, uint32_t T,
                                                  uint32_t max_inner) {
  using Inner = InnerRecM<Env, NE, NA, CNT>;
  const uint32_t tree = blockIdx.x * blockDim.x + threadIdx.x;
  if (tree >= T) return;
  Inner* root = static_cast<Inner*>(inner_v) + (size_t)tree * max_inner;
  uint4* wz = reinterpret_cast<uint4*>(root);
  for (size_t i = 0; i < sizeof(Inner) / 16; ++i) wz[i] = make_uint4(0u, 0u, 0u, 0u);
  root->parent = kNoParent;
  tree_nodes[tree] = 1;
  tree_iters[tree] = 0;
  tree_inner[tree] = 1;
  tree_leaves[tree] = 0;
}

}  // namespace mamcts
