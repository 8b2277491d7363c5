// search_async.cuh - asynchronous ("tree pool") scheduling of the compact two-tier search (search_compact.cuh).
//
// Why: with one lane per tree (search_compact_kernel) the 32 trees of a warp wait for the deepest descent
// and the longest rollout among them - a rollout has 11 steps on average, the longest of 32 has ~30, so the
// rollout part of an iteration runs with 11 of 32 lanes and the kernel as a whole with 19
// (profiles/r01_search_compact_v11_final.txt). And the kernel needs 124 registers with every phase's state
// live at once, which caps the SM at 16 warps.
//
// Here a tree is not bound to a lane. A warp owns a POOL of 32 * SPL trees whose per-tree state (current node,
// depth, environment state, rollout position, path record) lives in shared memory. An iteration is cut into
// homogeneous work items -
//     DESC   one level of select / expand              (mcts.h:300-305, stage_node.h:167-232)
//     ROLL   one Philox block (4 steps) of the rollout  (random_heuristic.h:27-104)
//     BACK   the whole backup of a finished iteration   (mcts.h:314-329, uct_statistic.h:134-144)
// - and the trees wait in one FIFO per kind. A pass of the warp takes up to 32 trees from the longest FIFO and
// runs that one kind of item for them with (nearly) all lanes active, then files every tree under what it
// needs next. Trees advance independently of each other: nobody waits for somebody else's rollout.
//
// Results are bit-identical to search_compact_kernel (same record layout, same arithmetic, same Philox stream
// positions, same per-tree order of operations); the same host code resets, merges and exports the trees.
// Built for the configurations the bench and BASELINE.json name: two agents, a reward-free other agent, a
// discount >= 0 and a zero per-step reward (CrossingState's defaults) - anything else runs the lane-per-tree
// kernel.
#pragma once
#include "search_compact.cuh"

namespace mamcts {

constexpr int kAsyncWarps = 4;                 // warps per block; every warp schedules its own pool
constexpr int kAsyncThreads = kAsyncWarps * 32;
#ifndef MAMCTS_ASYNC_SHARED_LEVELS
#define MAMCTS_ASYNC_SHARED_LEVELS 4           // path levels kept in shared memory (deeper ones: global scratch)
#endif
#ifndef MAMCTS_ASYNC_MIN_BLOCKS
#define MAMCTS_ASYNC_MIN_BLOCKS 2
#endif
#ifndef MAMCTS_ASYNC_PREFETCH
#define MAMCTS_ASYNC_PREFETCH 1                // request the next record from HBM when the child becomes known
#endif

enum : uint32_t { kPhDesc = 0, kPhRoll = 1, kPhBack = 2, kPhNone = 3 };

template <class CNT>
struct AsyncPathRec {   // one level of a path below the shared-memory levels
  using HeadWord = typename std::conditional<sizeof(CNT) == 2, uint2, uint4>::type;
  using CountPair = typename std::conditional<sizeof(CNT) == 2, uint32_t, uint2>::type;
  HeadWord head;
  double q0, v0;
  CountPair n;
  uint32_t node;
};

struct AsyncParams {
  void* path;               // [T][kPathMaxCompact] AsyncPathRec
  double* path_r;           // [T][kPathMaxCompact] reward of the ego agent on the edge leaving a level (where != +0.0)
  uint32_t slots_per_warp;  // trees per warp, <= 32 * SPL
};

template <int M, int SL, class HeadWord, class CountPair>
struct AsyncPool {
  double G0[M];                  // the return the backup starts from
  unsigned long long rlev[M];    // bit d: the edge leaving level d carries a reward != +0.0
  double p_q0[SL][M], p_v0[SL][M];
  HeadWord p_head[SL][M];
  CountPair p_n[SL][M];
  uint32_t p_node[SL][M];
  uint32_t word[M];              // state flags (3 bits) | depth << 8
  uint32_t node[M];              // current inner node
  int32_t x[2][M];               // environment state (positions; two agents)
  uint32_t it[M];                // rollout steps done
  uint32_t nn[M], ni[M], nl[M];  // nodes / inner records / leaf records of the tree
  uint32_t done[M];              // iterations finished in this launch
  uint8_t ring[3][M];            // one FIFO of slots per kind of work item
};

template <class Env, int NE, int NA, int NCHILD, class CIDX, class CNT, int KIND, int SPL>
__global__ void __launch_bounds__(kAsyncThreads, MAMCTS_ASYNC_MIN_BLOCKS)
search_async_kernel(const __grid_constant__ SearchParams p, const __grid_constant__ CompactParams c,
                    const __grid_constant__ AsyncParams ap) {
  using Inner = InnerRec<Env, NE, NA, NCHILD, CIDX, CNT>;
  constexpr int A = Env::A;
  static_assert(A == 2 && Env::kOtherRewardZero && Env::kOutcomeRewards, "see the header comment");
  using Leaf = LeafRec<A>;
  using Counters = typename Inner::Counters;
  using Head = typename Inner::Head;
  using PathRec = AsyncPathRec<CNT>;
  using HeadWord = typename PathRec::HeadWord;
  using CountPair = typename PathRec::CountPair;
  static_assert(sizeof(Head) == sizeof(HeadWord), "the head is written with one store");
  constexpr int M = 32 * SPL;
  constexpr int SL = MAMCTS_ASYNC_SHARED_LEVELS;
  using Pool = AsyncPool<M, SL, HeadWord, CountPair>;
  static_assert(M <= 256, "slot ids are bytes");

  __shared__ SearchTables tab;
  extern __shared__ __align__(16) unsigned char async_smem[];
  stage_tables(tab, p.tab);
  Pool& P = reinterpret_cast<Pool*>(async_smem)[threadIdx.x >> 5];
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t spw = ap.slots_per_warp;
  const uint32_t tree0 = (blockIdx.x * kAsyncWarps + (threadIdx.x >> 5)) * spw;

  uint8_t const_act[A];
#pragma unroll
  for (int a = 0; a < A; ++a) const_act[a] = tab.order[a == 0 ? 0 : 1][0];  // F1: first draw
  typename Env::State root;
#pragma unroll
  for (int i = 0; i < A; ++i) { root.x[i] = (i < Env::kStateInts) ? __ldg(c.root_x + i) : 0; root.last[i] = 0; }
  root.flags = 0;
  root.flags = (uint32_t)Env::terminal(p.env, root);

  // FIFO heads and lengths: the same values in every lane (they only change by ballot counts)
  uint32_t hD = 0, hR = 0, hB = 0, cD = 0, cR = 0, cB = 0;
#pragma unroll
  for (int k = 0; k < SPL; ++k) {
    const uint32_t j = lane + 32u * k;
    const uint32_t tree = tree0 + j;
    const bool valid = j < spw && tree < p.T && p.iterations > 0;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, valid);
    if (valid) {
      P.word[j] = root.flags & 7u;
      P.node[j] = 0;
      P.x[0][j] = root.x[0]; P.x[1][j] = root.x[1];
      P.it[j] = 0;
      P.rlev[j] = 0ull;
      P.nn[j] = p.tree_nodes[tree];
      P.ni[j] = c.tree_inner[tree];
      P.nl[j] = c.tree_leaves[tree];
      P.done[j] = 0;
      P.ring[kPhDesc][cD + __popc(m & lt_mask)] = (uint8_t)j;
    }
    cD += __popc(m);
  }
  __syncwarp();

  uint32_t acc_roll = 0, acc_expand = 0, acc_levels = 0, acc_exact = 0, acc_nodes = 0;

#pragma unroll 1
  for (;;) {
    // the longest FIFO (ties: backup first - it frees a tree for its next iteration)
    uint32_t ph = kPhBack, n = cB, head = hB;
    if (cR > n) { ph = kPhRoll; n = cR; head = hR; }
    if (cD > n) { ph = kPhDesc; n = cD; head = hD; }
    if (n == 0) break;
    n = n < 32u ? n : 32u;
    const bool active = lane < n;
    uint32_t j = 0;
    {
      uint32_t idx = head + lane;
      idx = idx >= (uint32_t)M ? idx - (uint32_t)M : idx;
      if (active) j = P.ring[ph][idx];
      head += n;
      head = head >= (uint32_t)M ? head - (uint32_t)M : head;
    }
    if (ph == kPhDesc) { hD = head; cD -= n; } else if (ph == kPhRoll) { hR = head; cR -= n; } else { hB = head; cB -= n; }
    uint32_t next = kPhNone;
    const uint32_t tree = tree0 + j;

    if (ph == kPhDesc) {
      // ---- one level of select & expand: mcts.h:300-305, stage_node.h:167-232 -------------------------
      if (active) {
        Inner* inner = static_cast<Inner*>(c.inner) + (size_t)tree * c.max_inner;
        const uint32_t w = P.word[j];
        uint32_t depth = w >> 8;
        uint32_t node = P.node[j];
        typename Env::State st;
        st.x[0] = P.x[0][j]; st.x[1] = P.x[1][j]; st.last[0] = 0; st.last[1] = 0; st.flags = w & 7u;
        const uint32_t nn = P.nn[j];
        Inner* nd = inner + node;
        constexpr int kPrefix = (int)((offsetof(Inner, q_e) + sizeof(double) * NE + 31) / 32);
        Sector pre[kPrefix];
#pragma unroll
        for (int i = 0; i < kPrefix; ++i) pre[i] = ld_sector(reinterpret_cast<const char*>(nd) + 32 * i);
        prefetch_l1(reinterpret_cast<const char*>(nd) + ((offsetof(Inner, child) + 31) & ~size_t(31)));
        double2 vge;
        __builtin_memcpy(&vge, reinterpret_cast<const char*>(pre) + offsetof(Inner, vg_e), sizeof(vge));
        double2 qe[NE / 2 + (NE & 1)];
        __builtin_memcpy(qe, reinterpret_cast<const char*>(pre) + offsetof(Inner, q_e), sizeof(double) * NE);
        if (NE & 1) qe[NE / 2].y = 0.0;
        Counters cc;
        __builtin_memcpy(&cc, pre, sizeof(cc));
        if (Env::terminal(p.env, st) || depth == p.max_depth || nn > p.node_budget) {   // :183-187
          nd->c.h.visit_count = (CNT)(cc.h.visit_count + 1);                               // :180
          P.G0[j] = vge.y;      // no rollout: the node's own latest return is backed up (mcts.h:309-329)
          next = kPhBack;
        } else if (depth >= (uint32_t)kPathMaxCompact) {
          atomicAdd(p.counters + 7, 1ull);
        } else {
          StatRegs<NE> ego;
          StatRegs<NA> oth;
          ego.N = cc.h.N[0]; ego.nexp = cc.h.nexp[0]; ego.V = vge.x;
          oth.N = cc.h.N[1]; oth.nexp = cc.h.nexp[1]; oth.V = 0.0;
#pragma unroll
          for (int a = 0; a < NE; ++a) { ego.Q[a] = (a & 1) ? qe[a / 2].y : qe[a / 2].x; ego.n[a] = cc.n_e[a]; }
#pragma unroll
          for (int a = 0; a < NA; ++a) { oth.Q[a] = 0.0; oth.n[a] = cc.n_o[a]; }
          uint8_t ja[A];
          BackRegs back[A];
          bool w0, w1;
          unsigned long long exact = 0;
          ja[0] = (uint8_t)select_action<NE>(ego, p, tab, 0, p.n_ego, exact, back[0], w0);   // :190-195
          ja[1] = (uint8_t)select_action<NA>(oth, p, tab, 1, p.n_oth, exact, back[1], w1, true);
          acc_exact += (uint32_t)exact;
          // nothing is stored during the descent: visit_count + 1 (:180) and a newly expanded action go into
          // the image of the head, which the backup of this level writes back (search_compact.cuh)
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
          if (depth < (uint32_t)SL) {
            P.p_head[depth][j] = head_w; P.p_n[depth][j] = pair_w; P.p_node[depth][j] = node_w;
            P.p_q0[depth][j] = back[0].q; P.p_v0[depth][j] = back[0].v;
          } else {
            PathRec* pr = static_cast<PathRec*>(ap.path) + (size_t)tree * kPathMaxCompact + depth;
            pr->head = head_w; pr->n = pair_w; pr->node = node_w; pr->q0 = back[0].q; pr->v0 = back[0].v;
          }
          const uint32_t slot = (uint32_t)ja[0] * NA + (uint32_t)ja[1];
          const uint32_t entry = nd->child[slot];                            // :198
          int32_t act[A];
#pragma unroll
          for (int a = 0; a < A; ++a) act[a] = ja[a];  // index-valued: others move by +index (F5)
          double r_ego, r_other, c0;
          Env::step(p.env, st, act, r_ego, r_other, c0);
          st.flags = (st.flags & ~1u) | (uint32_t)Env::terminal(p.env, st);
          if (__double_as_longlong(r_ego) != 0ll) {
            ap.path_r[(size_t)tree * kPathMaxCompact + depth] = r_ego;
            P.rlev[j] |= 1ull << depth;
          }
          depth += 1;
          next = kPhDesc;
          if (entry & 1u) {                                                  // existing inner node, :199-206
            node = entry >> 1;
#if MAMCTS_ASYNC_PREFETCH
            {
              const char* nx = reinterpret_cast<const char*>(inner + node);
              asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 64));
            }
#endif
          } else if (entry == 0u) {
            // expand, :207-231: the new node is a leaf record; its rollout is the tree's next work item
            const uint32_t nl = P.nl[j];
            P.nl[j] = nl + 1;
            P.nn[j] = nn + 1;
            acc_expand += 1;
            acc_nodes += 1;
            nd->child[slot] = leaf_entry<CIDX>(nl);
            P.it[j] = 0;
            next = kPhRoll;
          } else {
            // second selection of a leaf: promote it to an inner record and go on from there
            const uint32_t leaf = (entry >> 1) - 1u;
            const Leaf* leaves = static_cast<const Leaf*>(c.leaves) + (size_t)tree * p.max_nodes;
            double h[A];
#pragma unroll
            for (int a = 0; a < A; ++a) h[a] = leaves[leaf].h[a];
            const uint32_t id = P.ni[j];
            if (id >= c.max_inner) {
              atomicAdd(p.counters + 6, 1ull);
              next = kPhNone;
            } else {
              P.ni[j] = id + 1;
              Inner* cn = inner + id;
              uint4 img[sizeof(Inner) / 16];
#pragma unroll
              for (size_t i = 0; i < sizeof(Inner) / 16; ++i) img[i] = make_uint4(0u, 0u, 0u, 0u);
              Counters c1;
              __builtin_memcpy(&c1, img, sizeof(Counters));
              c1.h.N[0] = 1; c1.h.N[1] = 1;          // update_from_heuristic left 0 + 1 visits
              c1.ja[0] = ja[0]; c1.ja[1] = ja[1];
              __builtin_memcpy(reinterpret_cast<char*>(img) + offsetof(Inner, c), &c1, sizeof(Counters));
              const double ve[2] = {h[0], h[0]}, vo[2] = {h[1], h[1]};   // V = G = h
              __builtin_memcpy(reinterpret_cast<char*>(img) + offsetof(Inner, vg_e), ve, sizeof(ve));
              __builtin_memcpy(reinterpret_cast<char*>(img) + offsetof(Inner, vg_o), vo, sizeof(vo));
              __builtin_memcpy(reinterpret_cast<char*>(img) + offsetof(Inner, parent), &node, sizeof(uint32_t));
              constexpr size_t used = (offsetof(Inner, q_o) + 31) / 32;   // whole sectors; the reward-free agent's action values are never stored (search_compact.cuh)
#pragma unroll
              for (size_t i = 0; i < used; ++i) {
                Sector sec;
                __builtin_memcpy(&sec, &img[2 * i], sizeof(sec));
                st_sector(reinterpret_cast<char*>(cn) + 32 * i, sec);
              }
              nd->child[slot] = inner_entry<CIDX>(id);
              node = id;
            }
          }
          P.word[j] = (st.flags & 7u) | (depth << 8);
          P.node[j] = node;
          P.x[0][j] = st.x[0]; P.x[1][j] = st.x[1];
        }
      }
    } else if (ph == kPhRoll) {
      // ---- one Philox block of the rollout: random_heuristic.h:27-104 (lane_rollout, search_kernel.cuh) ----
      if (active) {
        constexpr int S = (KIND == MAMCTS_SRC_PHILOX) ? 8 / A : 4;
        const uint32_t w = P.word[j];
        const uint32_t depth = w >> 8;
        typename Env::State st;
        st.x[0] = P.x[0][j]; st.x[1] = P.x[1][j]; st.last[0] = 0; st.last[1] = 0; st.flags = w & 7u;
        uint32_t it = P.it[j];
        const uint32_t nn = P.nn[j];
        const uint32_t s_lo = nn - 1u;                      // creation index of the new leaf
        const uint32_t s_hi = p.first_tree_id + tree;
        const uint32_t depth_room = p.max_depth > depth ? p.max_depth - depth : 0u;
        const uint32_t limit = p.rh_cap < depth_room ? p.rh_cap : depth_room;
        bool fin = Env::terminal(p.env, st) | (it >= limit);
        uint32_t f = 0;
        if (!fin) {
          Philox4 blk;
          if (KIND == MAMCTS_SRC_PHILOX) blk = philox4x32_10(it / S, s_lo, s_hi, 0u, p.key0, p.key1);
#pragma unroll
          for (int sub = 0; sub < S; ++sub) {
            int32_t act[A];
            if (KIND == MAMCTS_SRC_PHILOX) {
              uint32_t h[A], nact[A];
#pragma unroll
              for (int a = 0; a < A; ++a) { h[a] = philox_half(blk, sub * A + a); nact[a] = Env::num_actions(p.env, a); }
              philox_bounded_all<A>(h, nact, it * A, s_lo, s_hi, p.key0, p.key1);
#pragma unroll
              for (int a = 0; a < A; ++a) act[a] = (int32_t)h[a];
            } else {
#pragma unroll
              for (int a = 0; a < A; ++a) act[a] = (int32_t)const_act[a];
            }
            f = Env::step_outcome(p.env, st, act);
            it += 1;
            fin = (f != 0u) | (it >= limit);
            if (fin) break;
          }
        }
        if (fin) {
          // the single nonzero term of the reference's sum (lane_rollout); update_from_heuristic
          // (uct_statistic.h:100-110): value_ = latest_return_ = h, visits 0 + 1
          double ego = 0.0;
          const double r_last = it ? Env::outcome_reward(p.env, f) : 0.0;
          if (r_last != 0.0) ego = __dadd_rn(ego, __dmul_rn(__ldg(p.disc + it - 1), r_last));
          const double other = it ? __dmul_rn(__ldg(p.disc + it - 1), 0.0) : 0.0;
          const double ego_h = __dmul_rn(ego, __ldg(p.disc + p.disc_len + it));
          Leaf* leaves = static_cast<Leaf*>(c.leaves) + (size_t)tree * p.max_nodes;
          const uint32_t new_leaf = P.nl[j] - 1u;
          leaves[new_leaf].h[0] = ego_h;
          leaves[new_leaf].h[1] = other;
          P.G0[j] = ego_h;
          acc_roll += it;
          next = kPhBack;
        } else {
          P.it[j] = it;
          P.x[0][j] = st.x[0]; P.x[1][j] = st.x[1];
          next = kPhRoll;
        }
      }
    } else {
      // ---- backpropagation: mcts.h:314-329, stage_node.h:259-271; recorded path, no loads from the tree ----
      if (active) {
        Inner* inner = static_cast<Inner*>(c.inner) + (size_t)tree * c.max_inner;
        const uint32_t depth = P.word[j] >> 8;
        double G = P.G0[j];
        const unsigned long long reward_levels = P.rlev[j];
        acc_levels += depth;
        uint32_t level = depth;
#pragma unroll 1
        while (level > 0) {
          level -= 1;
          uint32_t word;
          HeadWord head_w;
          CountPair pair_w;
          double q0, v0;
          if (level < (uint32_t)SL) {
            word = P.p_node[level][j]; head_w = P.p_head[level][j]; pair_w = P.p_n[level][j];
            q0 = P.p_q0[level][j]; v0 = P.p_v0[level][j];
          } else {
            const PathRec* pr = static_cast<const PathRec*>(ap.path) + (size_t)tree * kPathMaxCompact + level;
            word = pr->node; head_w = pr->head; pair_w = pr->n; q0 = pr->q0; v0 = pr->v0;
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
          BackRegs be;
          be.q = q0; be.v = v0; be.cnt0 = n_ego; be.nv0 = hd.N[0];
          const double r0 = has_reward ? ap.path_r[(size_t)tree * kPathMaxCompact + level] : 0.0;
          const BackOut e = back_compute(be, r0, p.gamma, G);
          hd.N[0] = (CNT)e.nv; hd.N[1] = (CNT)(hd.N[1] + 1);   // a reward-free agent: only the counts move
          HeadWord hw;
          __builtin_memcpy(&hw, &hd, sizeof(Head));
          *reinterpret_cast<HeadWord*>(&pn->c.h) = hw;
          pn->c.n_e[ae] = (CNT)e.cnt;
          pn->c.n_o[ao] = (CNT)(n_oth + 1u);
          *reinterpret_cast<double2*>(&pn->vg_e[0]) = make_double2(e.v, e.G);
          if (__double_as_longlong(e.q) != __double_as_longlong(be.q)) pn->q_e[ae] = e.q;
          G = e.G;
        }
        const uint32_t done = P.done[j] + 1u;
        P.done[j] = done;
        if (done < p.iterations) {
          P.word[j] = root.flags & 7u;
          P.node[j] = 0;
          P.x[0][j] = root.x[0]; P.x[1][j] = root.x[1];
          P.rlev[j] = 0ull;
          next = kPhDesc;
        }
      }
    }
    __syncwarp();
    // file every tree under what it needs next
    {
      const uint32_t mD = __ballot_sync(0xFFFFFFFFu, next == kPhDesc);
      const uint32_t mR = __ballot_sync(0xFFFFFFFFu, next == kPhRoll);
      const uint32_t mB = __ballot_sync(0xFFFFFFFFu, next == kPhBack);
      if (next != kPhNone) {
        const uint32_t m = next == kPhDesc ? mD : next == kPhRoll ? mR : mB;
        const uint32_t h = next == kPhDesc ? hD : next == kPhRoll ? hR : hB;
        const uint32_t cn = next == kPhDesc ? cD : next == kPhRoll ? cR : cB;
        uint32_t idx = h + cn + __popc(m & lt_mask);
        idx = idx >= (uint32_t)M ? idx - (uint32_t)M : idx;
        P.ring[next][idx] = (uint8_t)j;
      }
      cD += __popc(mD); cR += __popc(mR); cB += __popc(mB);
    }
    __syncwarp();
  }

  // the trees' bookkeeping goes back to global memory; counters
  uint32_t done_sum = 0;
#pragma unroll
  for (int k = 0; k < SPL; ++k) {
    const uint32_t j = lane + 32u * k;
    const uint32_t tree = tree0 + j;
    if (j < spw && tree < p.T && p.iterations > 0) {
      p.tree_nodes[tree] = P.nn[j];
      c.tree_inner[tree] = P.ni[j];
      c.tree_leaves[tree] = P.nl[j];
      p.tree_iters[tree] += P.done[j];
      done_sum += P.done[j];
    }
  }
  atomicAdd(p.counters + 0, (unsigned long long)done_sum);
  atomicAdd(p.counters + 1, (unsigned long long)acc_roll);
  atomicAdd(p.counters + 2, (unsigned long long)acc_expand);
  atomicAdd(p.counters + 3, (unsigned long long)acc_nodes);
  atomicAdd(p.counters + 4, (unsigned long long)acc_levels);
  atomicAdd(p.counters + 5, (unsigned long long)acc_exact);
}

}  // namespace mamcts
