// search_compact.cuh - two-tier ("compact") tree layout of the device-resident search, for
// configurations with a direct child table (2 agents).
//
// Why: the search is a chain of dependent record visits inside one tree, so throughput is
// (resident trees) / (latency of one iteration), and resident trees are limited by HBM per tree.
// Three quarters of the nodes of a 10 000-iteration tree are never visited again after their
// creation (profiles/: 24-26 % of the nodes have visit_count > 0). Such a node only has to remember
// its heuristic value: everything else the reference's StageNode holds for it is either constant
// (N = 1, V = G = h, no expanded action) or a function of (parent state, joint action), which the
// descent recomputes by carrying the state from the root (one execute() per level, ~20 integer
// instructions) instead of storing it. So:
//
//   LeafRec   16 B for 2 agents: h per agent. Index = creation order among leaves.
//   InnerRec  256 B, 64-B aligned (2 agents, 4/7 actions): parent, visit_count, joint action, the
//             statistics of every agent and the direct child table. A leaf is PROMOTED to an inner
//             record the second time the tree policy selects it (its statistics start as
//             N = 1, V = G = h, exactly what update_from_heuristic left in the reference's node).
//   child table entry: 0 = absent, (L + 1) << 1 = leaf L, (I << 1) | 1 = inner node I.
//
// A 10 000-iteration C2 tree takes 10 001 x 16 B + ~2 600 x 256 B = 0.83 MB instead of 2.88 MB: three
// times more trees stay inside the range of HBM the GPU translates quickly (profiles/
// r01_randrec_microbench.txt), and a new node costs 16 B of write traffic instead of 288 B.
//
// Results are bit-identical to the classic layout (search_kernel.cuh) and to the reference: the same
// selections, the same Philox streams (stream_lo = creation index of the node), the same FP64
// arithmetic in the same order; only where the values live changes.
#pragma once
#include "search_kernel.cuh"

namespace mamcts {

constexpr int kPathMaxCompact = 64;  // deepest leaf the compact layout can back up (error beyond)

template <class Env, int NE, int NA, int NCHILD, class CIDX>
struct alignas(64) InnerRec {
  static constexpr int A = Env::A;
  static_assert(NCHILD > 0, "the compact layout needs a direct child table");
  uint32_t parent;       // inner index of the parent (export only)
  uint32_t visit_count;  // StageNode::visit_count_
  uint8_t ja[8];         // joint action leading here (export only)
  StatRec<NE> ego;
  StatRec<NA> oth[A - 1];
  CIDX child[NCHILD];
};

template <int A>
struct LeafRec {
  double h[A];  // value_ = latest_return_ of every agent after update_from_heuristic
};

struct CompactParams {
  void* inner;            // [T][max_inner] InnerRec
  void* leaves;           // [T][max_nodes] LeafRec
  uint32_t* tree_inner;   // [T] inner records in use
  uint32_t* tree_leaves;  // [T] leaf records in use
  uint32_t max_inner;
  const int32_t* root_x;  // the decision state
};

template <class CIDX>
__device__ __forceinline__ CIDX leaf_entry(uint32_t leaf) { return (CIDX)((leaf + 1u) << 1); }
template <class CIDX>
__device__ __forceinline__ CIDX inner_entry(uint32_t inner) { return (CIDX)((inner << 1) | 1u); }

// counters[6] = trees that ran out of inner records, counters[7] = trees with a path deeper than
// kPathMaxCompact; either makes mamcts_search_run's caller see an error at the next read-back.
template <class Env, int NE, int NA, int NCHILD, class CIDX, int KIND>
__global__ void __launch_bounds__(kSearchThreads)
search_compact_kernel(const __grid_constant__ SearchParams p, const __grid_constant__ CompactParams c) {
  using Inner = InnerRec<Env, NE, NA, NCHILD, CIDX>;
  constexpr int A = Env::A;
  using Leaf = LeafRec<A>;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t tree = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * p.lanes_per_warp + lane;
  // the lanes that own a tree; they re-converge explicitly once per iteration (see below)
  const unsigned live = __ballot_sync(0xFFFFFFFFu, lane < p.lanes_per_warp && tree < p.T);
  if (lane >= p.lanes_per_warp || tree >= p.T) return;
  Inner* inner = static_cast<Inner*>(c.inner) + (size_t)tree * c.max_inner;
  Leaf* leaves = static_cast<Leaf*>(c.leaves) + (size_t)tree * p.max_nodes;
  uint32_t num_nodes = p.tree_nodes[tree];   // nodes created (root included): the creation index
  uint32_t num_inner = c.tree_inner[tree];
  uint32_t num_leaves = c.tree_leaves[tree];
  unsigned long long rollout_steps = 0, expand_steps = 0, backup_levels = 0, exact_evals = 0;
  const uint32_t nodes_before = num_nodes;
  uint8_t const_act[A];
#pragma unroll
  for (int a = 0; a < A; ++a) const_act[a] = p.tab->order[a == 0 ? 0 : 1][0];  // F1: first draw
  typename Env::State root;
#pragma unroll
  for (int i = 0; i < A; ++i) { root.x[i] = (i < Env::kStateInts) ? __ldg(c.root_x + i) : 0; root.last[i] = 0; }
  root.flags = 0;
  root.flags = (uint32_t)Env::terminal(p.env, root);

  uint32_t path_node[kPathMaxCompact];
  uint8_t path_ja[kPathMaxCompact][A];
  double path_r0[kPathMaxCompact], path_r1[kPathMaxCompact];
  BackRegs path_back[kPathMaxCompact][A];
  uint32_t done = 0;
  bool failed = false;

#pragma unroll 1
  for (uint32_t iter = 0; iter < p.iterations; ++iter) {   // a failed lane idles through the barriers
    uint32_t node = 0;
    uint32_t depth = 0;
    typename Env::State st = root;
    double G[A];
    bool do_rollout = false;
    uint32_t new_leaf = 0, creation = 0;
    // ---- select & expand: mcts.h:300-305, stage_node.h:167-232 --------------------------------
#pragma unroll 1
    while (!failed) {
      Inner* nd = inner + node;
      // one round trip: everything the level needs is requested before anything is used
      const uint32_t vc = nd->visit_count;
      const StatRegs<NE> ego = load_stat<NE>(&nd->ego);
      StatRegs<NA> oth0;
      if (A > 1) oth0 = load_stat<NA>(&nd->oth[0]);
      prefetch_l1(&nd->child[0]);
      nd->visit_count = vc + 1;                                          // :180
      if (Env::terminal(p.env, st) || depth == p.max_depth || num_nodes > p.node_budget) {  // :183-187
        // no rollout: the node's own latest return is backed up (mcts.h:309-329)
        G[0] = nd->ego.G;
        if (A > 1) G[1] = nd->oth[0].G;
        break;
      }
      if (depth >= (uint32_t)kPathMaxCompact) { failed = true; atomicAdd(p.counters + 7, 1ull); break; }
      uint8_t ja[A];
      BackRegs back[A];
      ja[0] = (uint8_t)choose_action<NE>(&nd->ego, ego, p, 0, p.n_ego, exact_evals, back[0]);  // :190-195
      if (A > 1) ja[1] = (uint8_t)choose_action<NA>(&nd->oth[0], oth0, p, 1, p.n_oth, exact_evals, back[1]);
      const uint32_t slot = (uint32_t)ja[0] * NA + (A > 1 ? (uint32_t)ja[A > 1 ? 1 : 0] : 0u);
      const uint32_t entry = nd->child[slot];                            // :198
      // the child's state and the reward collected on the edge (collect(), node_statistic.h:115-121)
      int32_t act[A];
#pragma unroll
      for (int a = 0; a < A; ++a) act[a] = ja[a];  // index-valued: others move by +index (F5)
      double r_ego, r_other, c0;
      Env::step(p.env, st, act, r_ego, r_other, c0);
      st.flags = (st.flags & ~1u) | (uint32_t)Env::terminal(p.env, st);
      path_node[depth] = node;
      path_r0[depth] = r_ego;
      path_r1[depth] = r_other;
#pragma unroll
      for (int a = 0; a < A; ++a) { path_ja[depth][a] = ja[a]; path_back[depth][a] = back[a]; }
      depth += 1;
      if (entry & 1u) { node = entry >> 1; continue; }                   // existing inner node, :199-206
      if (entry == 0u) {
        // expand, :207-231: the new node is a leaf record. Its rollout runs after the loop, where the
        // lanes of the warp have reconverged (inside the loop every depth would run its own copy).
        new_leaf = num_leaves++;
        creation = num_nodes++;
        expand_steps += 1;
        nd->child[slot] = leaf_entry<CIDX>(new_leaf);
        do_rollout = true;
        break;
      }
      // second selection of a leaf: promote it to an inner record and go on from there
      {
        const uint32_t leaf = (entry >> 1) - 1u;
        double h[A];
#pragma unroll
        for (int a = 0; a < A; ++a) h[a] = leaves[leaf].h[a];
        if (num_inner >= c.max_inner) { failed = true; atomicAdd(p.counters + 6, 1ull); break; }
        const uint32_t id = num_inner++;
        Inner* cn = inner + id;
        cn->parent = node;
        cn->visit_count = 0;
#pragma unroll
        for (int a = 0; a < 8; ++a) cn->ja[a] = (a < A) ? ja[a] : 0;
        cn->ego.V = h[0]; cn->ego.G = h[0]; cn->ego.N = 1; cn->ego.nexp = 0;
#pragma unroll
        for (int j = 0; j < A - 1; ++j) {
          cn->oth[j].V = h[j + 1]; cn->oth[j].G = h[j + 1]; cn->oth[j].N = 1; cn->oth[j].nexp = 0;
        }
#pragma unroll
        for (int k = 0; k < NCHILD; ++k) cn->child[k] = 0;
        nd->child[slot] = inner_entry<CIDX>(id);
        node = id;
      }
    }
    // The descent leaves the loop through four different exits at different depths; without an explicit
    // barrier the compiler threads the rollout onto the "expand" exit and every group of lanes runs
    // its own copy of the step loop (measured: 4.9 active lanes instead of 10.9).
    __syncwarp(live);
    if (failed) continue;
    // ---- heuristic: mcts.h:309-312, random_heuristic.h:106-134 ------------------------------------
    if (do_rollout) {
      double other = 0.0;
      uint32_t steps = 0;
      const double ego_h = lane_rollout<Env, KIND>(p, st, depth, p.first_tree_id + tree, creation, const_act,
                                                   other, steps);
      rollout_steps += steps;
      // update_from_heuristic (uct_statistic.h:100-110): value_ = latest_return_ = h, visits 0 + 1
      leaves[new_leaf].h[0] = ego_h;
      G[0] = ego_h;
#pragma unroll
      for (int j = 0; j < A - 1; ++j) { leaves[new_leaf].h[j + 1] = other; G[j + 1] = other; }
    }
    // ---- backpropagation: mcts.h:314-329, stage_node.h:259-271; recorded path, no loads ---------
    backup_levels += depth;
    uint32_t level = depth;
#pragma unroll 1
    while (level > 0) {
      level -= 1;
      Inner* pn = inner + path_node[level];
      const double r0 = path_r0[level], r1 = path_r1[level];
      G[0] = back_apply<NE>(&pn->ego, path_ja[level][0], path_back[level][0], r0, p.gamma, G[0]);
#pragma unroll
      for (int j = 0; j < A - 1; ++j)
        G[j + 1] = back_apply<NA>(&pn->oth[j], path_ja[level][j + 1], path_back[level][j + 1], r1, p.gamma, G[j + 1]);
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

template <class Env, int NE, int NA, int NCHILD, class CIDX>
__global__ void search_compact_reset_kernel(void* inner_v, uint32_t* tree_nodes, uint32_t* tree_iters,
                                            uint32_t* tree_inner, uint32_t* tree_leaves, uint32_t T,
                                            uint32_t max_inner) {
  using Inner = InnerRec<Env, NE, NA, NCHILD, CIDX>;
  const uint32_t tree = blockIdx.x * blockDim.x + threadIdx.x;
  if (tree >= T) return;
  Inner* root = static_cast<Inner*>(inner_v) + (size_t)tree * max_inner;
  root->parent = kNoParent;
  root->visit_count = 0;
  for (int a = 0; a < 8; ++a) root->ja[a] = 0;
  root->ego.V = 0.0; root->ego.G = 0.0; root->ego.N = 0; root->ego.nexp = 0;
  for (int j = 0; j < Env::A - 1; ++j) { root->oth[j].V = 0.0; root->oth[j].G = 0.0; root->oth[j].N = 0; root->oth[j].nexp = 0; }
  for (int k = 0; k < NCHILD; ++k) root->child[k] = 0;
  tree_nodes[tree] = 1;
  tree_iters[tree] = 0;
  tree_inner[tree] = 1;
  tree_leaves[tree] = 0;
}

}  // namespace mamcts
