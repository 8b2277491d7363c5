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
#include <cstddef>
#include <type_traits>

#include "search_kernel.cuh"

namespace mamcts {

#ifndef MAMCTS_COMPACT_MIN_BLOCKS
#define MAMCTS_COMPACT_MIN_BLOCKS 8   // 124 registers used. Registers are granted in steps that make 96 the
                                      // next useful budget (112 or 104 still give 8 blocks per SM, r82); 96 and
                                      // 80 registers (10 / 12 blocks) spill and are slower (r49, r57)
#endif
#ifndef MAMCTS_PATH_SHARED_LEVELS
#define MAMCTS_PATH_SHARED_LEVELS 5
#endif
constexpr int kPathMaxCompact = 64;  // deepest leaf the compact layout can back up (error beyond)

// Inner record, laid out by 32-byte sector so that a visit touches as few sectors as possible (DRAM
// sectors of randomly placed records and the latency of fetching them are what the kernel's time goes
// into):
// (C2 offsets, 16-bit counts and child entries)
//   [  0, 32) counters  every integer a visit reads or writes: N per agent, visit_count, #expanded per
//                       agent, n[a] of both agents, the joint action leading here - ONE sector with
//                       16-bit counts (searches of < 65 536 iterations), two with 32-bit counts
//   [ 32, 48) vg_e      V and G (value_, latest_return_) of the ego agent
//   [ 48, 80) q_e       the ego's action values
//   [ 80,136) child     the direct child table, then [136,140) the parent index
//   [144,160) vg_o      V and G of the other agent
//   [160,216) q_o       the other agent's action values
// For an environment whose other agent is reward-free (its values are +0.0 for the whole search and are
// neither read nor rewritten) a visit reads sectors 0, 1, 2 and the sector of the chosen child entry
// (sector 2 again for the first 8 slots) and dirties sectors 0 and 1, plus sector 2 when the chosen ego
// action is one of the last two: 3.7 sectors read and 2.3 written on average, against 4.4 and 3 with V / G
// of both agents in a sector of their own (the previous layout) and 8 / 6-7 with the first, AoS-by-agent,
// record.
template <class Env, int NE, int NA, int NCHILD, class CIDX, class CNT>
struct alignas(64) InnerRec {
  static constexpr int A = Env::A;
  static_assert(NCHILD > 0, "the compact layout needs a direct child table");
  static_assert(A == 2, "the compact layout is built for two agents");
  // the part of the counters every backup of a visit rewrites, sized for ONE 8- or 16-byte store
  struct alignas(4 * sizeof(CNT)) Head {
    CNT N[2];              // total_node_visits_ (ego, other)
    CNT visit_count;       // StageNode::visit_count_
    uint8_t nexp[2];       // ucb_statistics_.size(): the first nexp entries of the widening order
  };
  struct alignas(32) Counters {
    Head h;
    CNT n_e[NE];           // action_count_
    CNT n_o[NA];
    uint8_t ja[2];         // joint action leading here (export only)
  } c;
  alignas(16) double vg_e[2];   // value_, latest_return_ of the ego agent
  double q_e[NE];               // action_value_ of the ego agent
  CIDX child[NCHILD];
  uint32_t parent;              // inner index of the parent (export only)
  alignas(16) double vg_o[2];   // value_, latest_return_ of the other agent
  double q_o[NA];               // action_value_ of the other agent
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

// One 32-byte sector per lane in one instruction (LDG.256 / STG.256, sm_100): a record visit is a handful
// of whole-sector accesses, and every load instruction of 32 different records costs 32 L1 wavefronts
// whatever its width.
struct alignas(32) Sector { uint32_t w[8]; };
__device__ __forceinline__ Sector ld_sector(const void* ptr) {
  Sector s;
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(s.w[0]), "=r"(s.w[1]), "=r"(s.w[2]), "=r"(s.w[3]), "=r"(s.w[4]), "=r"(s.w[5]),
                 "=r"(s.w[6]), "=r"(s.w[7]) : "l"(ptr));
  return s;
}
__device__ __forceinline__ void st_sector(void* ptr, const Sector& s) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "l"(ptr), "r"(s.w[0]), "r"(s.w[1]), "r"(s.w[2]), "r"(s.w[3]), "r"(s.w[4]), "r"(s.w[5]),
                  "r"(s.w[6]), "r"(s.w[7]) : "memory");
}

template <class CIDX>
__device__ __forceinline__ CIDX leaf_entry(uint32_t leaf) { return (CIDX)((leaf + 1u) << 1); }
template <class CIDX>
__device__ __forceinline__ CIDX inner_entry(uint32_t inner) { return (CIDX)((inner << 1) | 1u); }

// counters[6] = trees that ran out of inner records, counters[7] = trees with a path deeper than
// kPathMaxCompact; either makes mamcts_search_run's caller see an error at the next read-back.
// HELP (few trees per GPU: at most 4 per warp, reward-free others, PHILOX, one rollout per leaf): a tree gets a group
// of 8 lanes instead of one. Lane 0 of the group owns the tree and runs the descent as ever; the other seven would
// have been idle and now shorten the two serial parts that do not touch the tree's records in order: the rollout
// (group_rollout: the Philox blocks of 32 steps computed side by side) and the backup of the shared-memory levels
// (one lane per level after the short chain of returns). Same trees bit for bit.
template <class Env, int NE, int NA, int NCHILD, class CIDX, class CNT, int KIND, bool MULTI = false, bool OVZ = false,
          bool HELP = false>
__global__ void __launch_bounds__(kSearchThreads, MAMCTS_COMPACT_MIN_BLOCKS)
search_compact_kernel(const __grid_constant__ SearchParams p, const __grid_constant__ CompactParams c) {
  using Inner = InnerRec<Env, NE, NA, NCHILD, CIDX, CNT>;
  constexpr int A = Env::A;
  using Leaf = LeafRec<A>;
  static_assert(!HELP || (OVZ && !MULTI && KIND == MAMCTS_SRC_PHILOX && A == 2), "see the comment above");
  __shared__ SearchTables tab;
  stage_tables(tab, p.tab);   // before any lane leaves: contains a block barrier
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t gl = HELP ? (lane & 7u) : 0u;               // position inside the tree's group of lanes
  const bool owner = gl == 0u;                                // the lane that owns the tree
  const uint32_t slot_in_warp = HELP ? (lane >> 3) : lane;
  const unsigned gmask = HELP ? (0xFFu << (lane & 24u)) : 0u; // the group's lanes
  const uint32_t tree = warp * p.lanes_per_warp + slot_in_warp;
  // the lanes that work on a tree; they re-converge explicitly once per iteration (see below)
  const unsigned live = __ballot_sync(0xFFFFFFFFu, slot_in_warp < p.lanes_per_warp && tree < p.T);
  if (slot_in_warp >= p.lanes_per_warp || tree >= p.T) return;
  Inner* inner = static_cast<Inner*>(c.inner) + (size_t)tree * c.max_inner;
  Leaf* leaves = static_cast<Leaf*>(c.leaves) + (size_t)tree * p.max_nodes;
  uint32_t num_nodes = p.tree_nodes[tree];   // nodes created (root included): the creation index
  uint32_t num_inner = c.tree_inner[tree];
  uint32_t num_leaves = c.tree_leaves[tree];
  // Launch totals per lane. With 16-bit counts a launch has fewer than 65 535 iterations of at most 64 levels
  // and min(rollout cap, 2^24) steps... the sums of one lane fit 32 bits whenever iterations * cap does; the
  // expansions of a launch are the nodes it created (num_nodes - p.tree_nodes[tree], read again at the end).
  using Acc = typename std::conditional<sizeof(CNT) == 2, uint32_t, unsigned long long>::type;
  Acc rollout_steps = 0, backup_levels = 0, exact_evals = 0;
  uint8_t const_act[A];
#pragma unroll
  for (int a = 0; a < A; ++a) const_act[a] = tab.order[a == 0 ? 0 : 1][0];  // F1: first draw

  // The path of the current iteration (collect(), node_statistic.h:115-121) and everything the backup
  // of a level needs: the head of the counters as loaded, the counts and values of the chosen actions
  // and the node values. These arrays live in local memory, which is interleaved by lane (one
  // wavefront serves all 32 lanes, unlike a global access of 32 records) - but every word of them is
  // L2 traffic, as much as the tree itself, so a level records 32 bytes (48 when the other agent's
  // values are not known to be zero) plus its rewards when they are not +0.0 (a bit mask in a
  // register says which levels have one), not the whole counters sector.
  using Counters = typename Inner::Counters;
  using Head = typename Inner::Head;
  using HeadWord = typename std::conditional<sizeof(Head) == 8, uint2, uint4>::type;
  static_assert(sizeof(Head) == sizeof(HeadWord), "the head is written with one store");
  using CountPair = typename std::conditional<sizeof(CNT) == 2, uint32_t, uint2>::type;
  uint32_t path_node[kPathMaxCompact];   // inner index | ego action << 24 | other action << 28
  double path_r0[kPathMaxCompact], path_q0[kPathMaxCompact], path_v0[kPathMaxCompact];
  double path_r1[kPathMaxCompact], path_q1[kPathMaxCompact], path_v1[kPathMaxCompact];  // unused if zero
  HeadWord path_head[kPathMaxCompact];
  CountPair path_n[kPathMaxCompact];     // n[a] of the chosen actions as loaded
  // The first kSharedLevels levels (nearly every visit) keep their record in shared memory instead:
  // no L1/L2 traffic at all. Column = thread, so a warp's access is conflict-free.
  constexpr int kSharedLevels = MAMCTS_PATH_SHARED_LEVELS;
  constexpr int kSL = kSharedLevels > 0 ? kSharedLevels : 1;
  __shared__ uint32_t s_node[kSL][kSearchThreads];
  __shared__ CountPair s_n[kSL][kSearchThreads];
  __shared__ HeadWord s_head[kSL][kSearchThreads];
  __shared__ double s_q0[kSL][kSearchThreads], s_v0[kSL][kSearchThreads];
  __shared__ double s_r0[HELP ? kSL : 1][HELP ? kSearchThreads : 1];   // HELP: the rewards of the shared levels
  static_assert(!HELP || (kSharedLevels >= 1 && kSharedLevels <= 8), "HELP: one lane of the group per shared level");
  const uint32_t tid = threadIdx.x;
  const uint32_t otid = HELP ? (tid & ~7u) : tid;            // the owner's column of the shared arrays
  uint32_t done = 0;
  bool failed = false;
  // Environments whose other agents never receive a reward (CrossingState: rewards[1..] stay 0,
  // crossing_state.h:145): with a discount >= 0 every return of theirs is +0.0 (m >= 0, m * 0.0 = +0.0),
  // hence every G = 0.0 + gamma * (+0.0), Q and V of theirs is +0.0 for the whole search - their values
  // are neither loaded nor recomputed nor rewritten: a record is created with them zero-filled, which
  // is what the reference's statistic holds, and their selection is integer-only (select_action).
  // A template parameter (the launch picks OVZ = Env::kOtherRewardZero && gamma >= 0): the other agent's V, G
  // and seven action values drop out of the register allocation, not just out of the executed path.
  static_assert(!OVZ || Env::kOtherRewardZero, "OVZ needs a reward-free other agent");
  constexpr bool other_values_zero = OVZ;

#pragma unroll 1
  for (uint32_t iter = 0; iter < p.iterations; ++iter) {   // a failed lane idles through the barriers
    uint32_t node = 0;
    uint32_t depth = 0;
    // the decision state, read again every iteration (two L1 hits next to the root record's fetch) rather
    // than held in registers for the whole launch
    typename Env::State st;
#pragma unroll
    for (int i = 0; i < A; ++i) { st.x[i] = (i < Env::kStateInts) ? __ldg(c.root_x + i) : 0; st.last[i] = 0; }
    st.flags = 0;
    st.flags = (uint32_t)Env::terminal(p.env, st);
    double G[A];
    bool do_rollout = false;
    unsigned long long reward_levels = 0;   // bit d: the edge leaving level d carries a reward != +0.0
    uint32_t new_leaf = 0, creation = 0;
    // ---- select & expand: mcts.h:300-305, stage_node.h:167-232 --------------------------------
#pragma unroll 1
    while (!failed && (!HELP || owner)) {
      Inner* nd = inner + node;
      // One round trip, and as few load instructions as possible: each lane reads its own record, so
      // every global load costs 32 L1 wavefronts whatever its width (the L1/LSU pipe, not DRAM, was the
      // limiter with scalar loads; profiles/) - the record's prefix is pulled with 256-bit loads.
      // the counters, the ego's V / G and its action values: the first kPrefix sectors of the record
      constexpr int kPrefix = (int)((offsetof(Inner, q_e) + sizeof(double) * NE + 31) / 32);
      Sector pre[kPrefix];
#pragma unroll
      for (int i = 0; i < kPrefix; ++i) pre[i] = ld_sector(reinterpret_cast<const char*>(nd) + 32 * i);
      double2 vge;                                                            // V, G of the ego
      __builtin_memcpy(&vge, reinterpret_cast<const char*>(pre) + offsetof(Inner, vg_e), sizeof(vge));
      const double2 vgo = other_values_zero ? make_double2(0.0, 0.0) : *reinterpret_cast<const double2*>(&nd->vg_o[0]);
      double2 qe[NE / 2 + (NE & 1)], qo[NA / 2 + (NA & 1)];
      __builtin_memcpy(qe, reinterpret_cast<const char*>(pre) + offsetof(Inner, q_e), sizeof(double) * NE);
      if (NE & 1) qe[NE / 2].y = 0.0;
#pragma unroll
      for (int i = 0; i < NA / 2 + (NA & 1); ++i)
        qo[i] = other_values_zero ? make_double2(0.0, 0.0) : reinterpret_cast<const double2*>(&nd->q_o[0])[i];
      // child table: its first entries came with the prefix, the sector behind the prefix is requested
      // up front, and the entry load below fetches the last one when the chosen slot lies there
      // (requesting all of it up front is neutral or costs DRAM sectors: r01_search_sweeps.txt r57, r86)
      prefetch_l1(reinterpret_cast<const char*>(nd) + ((offsetof(Inner, child) + 31) & ~size_t(31)));
      Counters cc;
      __builtin_memcpy(&cc, pre, sizeof(cc));
      if (Env::terminal(p.env, st) || depth == p.max_depth || num_nodes > p.node_budget) {  // :183-187
        // no rollout: the node's own latest return is backed up (mcts.h:309-329)
        nd->c.h.visit_count = (CNT)(cc.h.visit_count + 1);               // :180
        G[0] = vge.y;
        G[1] = vgo.y;
        break;
      }
      if (depth >= (uint32_t)kPathMaxCompact) { failed = true; atomicAdd(p.counters + 7, 1ull); break; }
      StatRegs<NE> ego;
      StatRegs<NA> oth;
      ego.N = cc.h.N[0]; ego.nexp = cc.h.nexp[0]; ego.V = vge.x;
      oth.N = cc.h.N[1]; oth.nexp = cc.h.nexp[1]; oth.V = vgo.x;
#pragma unroll
      for (int a = 0; a < NE; ++a) { ego.Q[a] = (a & 1) ? qe[a / 2].y : qe[a / 2].x; ego.n[a] = cc.n_e[a]; }
#pragma unroll
      for (int a = 0; a < NA; ++a) { oth.Q[a] = (a & 1) ? qo[a / 2].y : qo[a / 2].x; oth.n[a] = cc.n_o[a]; }
      uint8_t ja[A];
      BackRegs back[A];
      bool w0, w1;
      ja[0] = (uint8_t)select_action<NE>(ego, p, tab, 0, p.n_ego, exact_evals, back[0], w0);  // :190-195
      ja[1] = (uint8_t)select_action<NA>(oth, p, tab, 1, p.n_oth, exact_evals, back[1], w1, other_values_zero);
      // Nothing is stored during the descent: visit_count + 1 (:180) and a newly expanded action
      // (ucb_statistics_[a] = UcbPair()) go into the image of the head, which this iteration's backup
      // of the level writes back together with N, followed by the two n[a].
      cc.h.visit_count = (CNT)(cc.h.visit_count + 1);
      if (w0) cc.h.nexp[0] = (uint8_t)(cc.h.nexp[0] + 1);
      if (w1) cc.h.nexp[1] = (uint8_t)(cc.h.nexp[1] + 1);
      HeadWord head_w;
      __builtin_memcpy(&head_w, &cc.h, sizeof(Head));
      CountPair pair_w;
      if (sizeof(CNT) == 2) {
        const uint32_t pair = back[0].cnt0 | (back[1].cnt0 << 16);
        __builtin_memcpy(&pair_w, &pair, sizeof(pair));
      } else {
        const uint2 pair = make_uint2(back[0].cnt0, back[1].cnt0);
        __builtin_memcpy(&pair_w, &pair, sizeof(CountPair));
      }
      const uint32_t node_w = node | ((uint32_t)ja[0] << 24) | ((uint32_t)ja[1] << 28);
      if (depth < (uint32_t)kSharedLevels) {
        s_head[depth][tid] = head_w; s_n[depth][tid] = pair_w; s_node[depth][tid] = node_w;
        s_q0[depth][tid] = back[0].q; s_v0[depth][tid] = back[0].v;
      } else {
        path_head[depth] = head_w; path_n[depth] = pair_w; path_node[depth] = node_w;
        path_q0[depth] = back[0].q; path_v0[depth] = back[0].v;
      }
      if (!other_values_zero) { path_q1[depth] = back[1].q; path_v1[depth] = back[1].v; }
      const uint32_t slot = (uint32_t)ja[0] * NA + (uint32_t)ja[1];
      const uint32_t entry = nd->child[slot];                            // :198
      // the child's state and the reward collected on the edge (collect(), node_statistic.h:115-121)
      int32_t act[A];
#pragma unroll
      for (int a = 0; a < A; ++a) act[a] = ja[a];  // index-valued: others move by +index (F5)
      double r_ego, r_other, c0;
      Env::step(p.env, st, act, r_ego, r_other, c0);
      st.flags = (st.flags & ~1u) | (uint32_t)Env::terminal(p.env, st);
      if (__double_as_longlong(r_ego) != 0ll || (!other_values_zero && __double_as_longlong(r_other) != 0ll)) {
        path_r0[depth] = r_ego;
        if (HELP && depth < (uint32_t)kSharedLevels) s_r0[depth][tid] = r_ego;
        if (!other_values_zero) path_r1[depth] = r_other;
        reward_levels |= 1ull << depth;
      }
      depth += 1;
      if (entry & 1u) { node = entry >> 1; continue; }                   // existing inner node, :199-206
      if (entry == 0u) {
        // expand, :207-231: the new node is a leaf record. Its rollout runs after the loop, where the
        // lanes of the warp have reconverged (inside the loop every depth would run its own copy).
        new_leaf = num_leaves++;
        creation = num_nodes++;
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
        // The WHOLE record is written, every sector completely (zeros where the reference's statistic
        // has nothing yet: action values of unexpanded actions, padding): a sector that is only partly
        // written has to be read from HBM first when it leaves the L2, and one that is never written
        // would be read as is by the visit that follows. (A new action value is 0 = what is stored, so
        // the backup of a first visit needs no store for it - no partial write of the q_o sectors.)
        {
          uint4 img[sizeof(Inner) / 16];
#pragma unroll
          for (size_t i = 0; i < sizeof(Inner) / 16; ++i) img[i] = make_uint4(0u, 0u, 0u, 0u);
          Counters c0;
          __builtin_memcpy(&c0, img, sizeof(Counters));
          c0.h.N[0] = 1; c0.h.N[1] = 1;          // update_from_heuristic left 0 + 1 visits
          c0.ja[0] = ja[0]; c0.ja[1] = ja[1];
          __builtin_memcpy(reinterpret_cast<char*>(img) + offsetof(Inner, c), &c0, sizeof(Counters));
          const double ve[2] = {h[0], h[0]}, vo[2] = {h[1], h[1]};   // V = G = h
          __builtin_memcpy(reinterpret_cast<char*>(img) + offsetof(Inner, vg_e), ve, sizeof(ve));
          __builtin_memcpy(reinterpret_cast<char*>(img) + offsetof(Inner, vg_o), vo, sizeof(vo));
          __builtin_memcpy(reinterpret_cast<char*>(img) + offsetof(Inner, parent), &node, sizeof(uint32_t));
          // whole sectors; a reward-free agent's action values (+0.0 for the whole search) are never read by
          // the kernel and the export reports them as the zeros they are (SearchC::read_stat), so the sectors
          // that hold nothing else are not written at all
          constexpr size_t used = other_values_zero ? (offsetof(Inner, q_o) + 31) / 32
                                                    : (offsetof(Inner, q_o) + sizeof(double) * NA + 31) / 32;
#pragma unroll
          for (size_t i = 0; i < used; ++i) {
            Sector sec;
            __builtin_memcpy(&sec, &img[2 * i], sizeof(sec));
            st_sector(reinterpret_cast<char*>(cn) + 32 * i, sec);
          }
        }
        nd->child[slot] = inner_entry<CIDX>(id);
        node = id;
      }
    }
    // The descent leaves the loop through four different exits at different depths; without an explicit
    // barrier the compiler threads the rollout onto the "expand" exit and every group of lanes runs
    // its own copy of the step loop (measured: 4.9 active lanes instead of 10.9).
    __syncwarp(live);
    if constexpr (HELP) {
      // what the group needs from the owner: where the descent ended
      uint32_t fl = (failed ? 1u : 0u) | (do_rollout ? 2u : 0u) | ((st.flags & 7u) << 2) | (depth << 8);
      fl = __shfl_sync(gmask, fl, 0, 8);
      failed = fl & 1u; do_rollout = fl & 2u; st.flags = (fl >> 2) & 7u; depth = fl >> 8;
#pragma unroll
      for (int i = 0; i < A; ++i) st.x[i] = __shfl_sync(gmask, st.x[i], 0, 8);
      creation = __shfl_sync(gmask, creation, 0, 8);
    }
    if (failed) continue;
    // ---- heuristic: mcts.h:309-312, random_heuristic.h:106-134 ------------------------------------
    if (do_rollout) {
      double other = 0.0;
      uint32_t steps = 0;
      double ego_h;
      if constexpr (HELP) ego_h = group_rollout<Env>(p, st, depth, p.first_tree_id + tree, creation, gl, gmask, other, steps);
      else ego_h = leaf_heuristic<Env, KIND, MULTI>(p, st, depth, p.first_tree_id + tree, creation, const_act, other, steps);
      if (!owner) steps = 0;
      rollout_steps += steps;
      if (sizeof(Acc) == 4 && rollout_steps >= 0x80000000u) {   // never with realistic rollout caps; keeps the sum exact
        atomicAdd(p.counters + 1, (unsigned long long)rollout_steps);
        rollout_steps = 0;
      }
      // update_from_heuristic (uct_statistic.h:100-110): value_ = latest_return_ = h, visits 0 + 1
      if (owner) {
        leaves[new_leaf].h[0] = ego_h;
        leaves[new_leaf].h[1] = other;
      }
      G[0] = ego_h;
      G[1] = other;
    }
    // ---- backpropagation: mcts.h:314-329, stage_node.h:259-271; recorded path, no loads ---------
    if (owner) backup_levels += depth;
    uint32_t level = depth;
    if constexpr (HELP) {
      // The levels kept in shared memory (nearly all of a path), one lane of the group each: the return that
      // reaches a level is a short chain (G = r + gamma G' per level below it, r != 0 on flagged levels only)
      // that every lane walks; the update of the level's statistic (two FP64 divisions) and its stores then
      // run side by side. Deeper levels stay with the owner (the loop below), before the shared ones.
      const uint32_t top = depth < (uint32_t)kSharedLevels ? depth : (uint32_t)kSharedLevels;
      {
        uint32_t lo = (uint32_t)reward_levels, hi = (uint32_t)(reward_levels >> 32);
        lo = __shfl_sync(gmask, lo, 0, 8); hi = __shfl_sync(gmask, hi, 0, 8);
        reward_levels = ((unsigned long long)hi << 32) | lo;
      }
      level = owner ? depth : top;      // the owner's loop below stops at the shared levels; the others skip it
    }
#pragma unroll 1
    while (level > (HELP ? (depth < (uint32_t)kSharedLevels ? depth : (uint32_t)kSharedLevels) : 0u)) {
      level -= 1;
      uint32_t word;
      HeadWord head_w;
      CountPair pair_w;
      double q0, v0;
      if (level < (uint32_t)kSharedLevels) {
        word = s_node[level][tid]; head_w = s_head[level][tid]; pair_w = s_n[level][tid];
        q0 = s_q0[level][tid]; v0 = s_v0[level][tid];
      } else {
        word = path_node[level]; head_w = path_head[level]; pair_w = path_n[level];
        q0 = path_q0[level]; v0 = path_v0[level];
      }
      Inner* pn = inner + (word & 0xFFFFFFu);
      Head hd;
      __builtin_memcpy(&hd, &head_w, sizeof(Head));
      const uint32_t ae = (word >> 24) & 15u, ao = word >> 28;
      uint32_t n_ego, n_oth;
      if (sizeof(CNT) == 2) {
        uint32_t pair;
        __builtin_memcpy(&pair, &pair_w, sizeof(pair));
        n_ego = pair & 0xFFFFu; n_oth = pair >> 16;
      } else {
        uint2 pair;
        __builtin_memcpy(&pair, &pair_w, sizeof(CountPair));
        n_ego = pair.x; n_oth = pair.y;
      }
      const bool has_reward = (reward_levels >> level) & 1ull;
      BackRegs be, bo;
      be.q = q0; be.v = v0; be.cnt0 = n_ego; be.nv0 = hd.N[0];
      bo.q = other_values_zero ? 0.0 : path_q1[level];
      bo.v = other_values_zero ? 0.0 : path_v1[level];
      bo.cnt0 = n_oth; bo.nv0 = hd.N[1];
      const BackOut e = back_compute(be, has_reward ? path_r0[level] : 0.0, p.gamma, G[0]);
      BackOut o;   // a reward-free agent: every value stays +0.0 (see other_values_zero), only the counts move
      if (other_values_zero) { o.G = 0.0; o.q = 0.0; o.v = 0.0; o.cnt = bo.cnt0 + 1; o.nv = bo.nv0 + 1; }
      else o = back_compute(bo, has_reward ? path_r1[level] : 0.0, p.gamma, G[1]);
      hd.N[0] = (CNT)e.nv; hd.N[1] = (CNT)o.nv;
      // the head of the counters goes back with one store, then the two action counts; V and G of an
      // agent with one 16-byte store (a reward-free agent's stay what the record was created with)
      {
        HeadWord hw;
        __builtin_memcpy(&hw, &hd, sizeof(Head));
        *reinterpret_cast<HeadWord*>(&pn->c.h) = hw;
        pn->c.n_e[ae] = (CNT)e.cnt;
        pn->c.n_o[ao] = (CNT)o.cnt;
        *reinterpret_cast<double2*>(&pn->vg_e[0]) = make_double2(e.v, e.G);
        if (!other_values_zero) *reinterpret_cast<double2*>(&pn->vg_o[0]) = make_double2(o.v, o.G);
      }
      // an action value that did not change (always the case for CrossingState's reward-free other
      // agents) is not rewritten: its sector stays clean and is never written back to HBM
      // (bit comparison: +0.0 and -0.0 are different values to keep; a newly expanded action's slot
      // already holds its initial 0.0 - records are created zero-filled)
      if (__double_as_longlong(e.q) != __double_as_longlong(be.q)) pn->q_e[ae] = e.q;
      if (__double_as_longlong(o.q) != __double_as_longlong(bo.q)) pn->q_o[ao] = o.q;
      G[0] = e.G;
      G[1] = o.G;
    }
    if constexpr (HELP) {
      const uint32_t top = depth < (uint32_t)kSharedLevels ? depth : (uint32_t)kSharedLevels;
      // the return that enters the shared levels, from the owner (its own rollout value or the deeper levels')
      double Gc;
      {
        int lo = __double2loint(G[0]), hi = __double2hiint(G[0]);
        lo = __shfl_sync(gmask, lo, 0, 8); hi = __shfl_sync(gmask, hi, 0, 8);
        Gc = __hiloint2double(hi, lo);
      }
      double child_G = 0.0, my_r = 0.0;
#pragma unroll 1
      for (uint32_t l = top; l-- > 0u;) {
        const double r = ((reward_levels >> l) & 1ull) ? s_r0[l][otid] : 0.0;
        if (l == gl) { child_G = Gc; my_r = r; }
        Gc = __dadd_rn(r, __dmul_rn(p.gamma, Gc));     // back_compute's o.G of level l
      }
      if (gl < top) {
        const uint32_t word = s_node[gl][otid];
        const HeadWord head_w = s_head[gl][otid];
        const CountPair pair_w = s_n[gl][otid];
        Inner* pn = inner + (word & 0xFFFFFFu);
        Head hd;
        __builtin_memcpy(&hd, &head_w, sizeof(Head));
        const uint32_t ae = (word >> 24) & 15u, ao = word >> 28;
        uint32_t n_ego, n_oth;
        if (sizeof(CNT) == 2) {
          uint32_t pair;
          __builtin_memcpy(&pair, &pair_w, sizeof(pair));
          n_ego = pair & 0xFFFFu; n_oth = pair >> 16;
        } else {
          uint2 pair;
          __builtin_memcpy(&pair, &pair_w, sizeof(CountPair));
          n_ego = pair.x; n_oth = pair.y;
        }
        BackRegs be;
        be.q = s_q0[gl][otid]; be.v = s_v0[gl][otid]; be.cnt0 = n_ego; be.nv0 = hd.N[0];
        const BackOut e = back_compute(be, my_r, p.gamma, child_G);
        hd.N[0] = (CNT)e.nv; hd.N[1] = (CNT)(hd.N[1] + 1);   // a reward-free agent: only the counts move
        HeadWord hw;
        __builtin_memcpy(&hw, &hd, sizeof(Head));
        *reinterpret_cast<HeadWord*>(&pn->c.h) = hw;
        pn->c.n_e[ae] = (CNT)e.cnt;
        pn->c.n_o[ao] = (CNT)(n_oth + 1u);
        *reinterpret_cast<double2*>(&pn->vg_e[0]) = make_double2(e.v, e.G);
        if (__double_as_longlong(e.q) != __double_as_longlong(be.q)) pn->q_e[ae] = e.q;
      }
      __syncwarp(gmask);   // the shared levels are read before the owner's next descent overwrites them
    }
    done += 1;
  }
  if (HELP && !owner) return;
  const uint32_t nodes_before = p.tree_nodes[tree];
  p.tree_nodes[tree] = num_nodes;
  c.tree_inner[tree] = num_inner;
  c.tree_leaves[tree] = num_leaves;
  p.tree_iters[tree] += done;
  atomicAdd(p.counters + 0, (unsigned long long)done);
  atomicAdd(p.counters + 1, (unsigned long long)rollout_steps);
  atomicAdd(p.counters + 2, (unsigned long long)(num_nodes - nodes_before));   // one execute() per expansion
  atomicAdd(p.counters + 3, (unsigned long long)(num_nodes - nodes_before));
  atomicAdd(p.counters + 4, (unsigned long long)backup_levels);
  atomicAdd(p.counters + 5, (unsigned long long)exact_evals);
}

template <class Env, int NE, int NA, int NCHILD, class CIDX, class CNT>
__global__ void search_compact_reset_kernel(void* inner_v, uint32_t* tree_nodes, uint32_t* tree_iters,
                                            uint32_t* tree_inner, uint32_t* tree_leaves, uint32_t T,
                                            uint32_t max_inner) {
  using Inner = InnerRec<Env, NE, NA, NCHILD, CIDX, CNT>;
  const uint32_t tree = blockIdx.x * blockDim.x + threadIdx.x;
  if (tree >= T) return;
  Inner* root = static_cast<Inner*>(inner_v) + (size_t)tree * max_inner;
  root->parent = kNoParent;
  root->c.h.visit_count = 0;
  root->c.h.N[0] = 0; root->c.h.N[1] = 0; root->c.h.nexp[0] = 0; root->c.h.nexp[1] = 0;
  root->c.ja[0] = 0; root->c.ja[1] = 0;
  for (int a = 0; a < NE; ++a) { root->c.n_e[a] = 0; root->q_e[a] = 0.0; }  // the root merge reads n[a] > 0 as "expanded"
  for (int a = 0; a < NA; ++a) { root->c.n_o[a] = 0; root->q_o[a] = 0.0; }
  root->vg_e[0] = 0.0; root->vg_e[1] = 0.0; root->vg_o[0] = 0.0; root->vg_o[1] = 0.0;
  for (int k = 0; k < NCHILD; ++k) root->child[k] = 0;
  tree_nodes[tree] = 1;
  tree_iters[tree] = 0;
  tree_inner[tree] = 1;
  tree_leaves[tree] = 0;
}

}  // namespace mamcts
