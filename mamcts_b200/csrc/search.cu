// search.cu - device-resident root-parallel MCTS: T independent trees live in HBM, one lane owns one
// tree and runs whole iterations (select/expand -> rollout -> backup) without ever returning to the
// host. Replaces, for T trees at once, Mcts<S,UctStatistic,UctStatistic,RandomHeuristic>::search
// (reference mcts/mcts.h:166-186,294-333), StageNode::select_or_expand / update_statistics
// (mcts/stage_node.h:167-271) and UctStatistic (mcts/statistics/uct_statistic.h).
//
// HBM layout: per tree a contiguous arena of fixed-size node records (AoS: a lane
// that touches a node uses every byte of the lines it pulls) and an open-addressing child table
// (uint32 child index per slot, keyed by hash(parent, joint action), verified against the child's
// own header). Statistics are FP64 and are updated in the reference's operation order; log(N) comes
// from a host-computed table and the widening thresholds from host libm, so a REFERENCE_CONST
// search reproduces the reference's tree bit for bit.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include "engine.h"
#include "env.cuh"
#include "philox.cuh"
#include "root_merge.h"

namespace mamcts {

constexpr uint32_t kNoParent = 0xFFFFFFFFu;
constexpr int kSearchThreads = 128;

template <int NACT>
struct alignas(8) StatRec {
  double V;        // value_
  double G;        // latest_return_
  uint32_t N;      // total_node_visits_
  uint32_t nexp;   // ucb_statistics_.size(): the first nexp entries of the widening order
  double Q[NACT];  // action_value_
  uint32_t n[NACT + (NACT & 1)];  // action_count_
};

template <class Env, int NE, int NA>
struct alignas(8) NodeRec {
  static constexpr int A = Env::A;
  uint32_t parent;
  uint32_t visit_count;              // StageNode::visit_count_
  uint32_t depth;
  uint8_t flags;                     // bit0 terminal, bit1 goal, bit2 collided
  uint8_t ja[(A + 1 + 3) / 4 * 4 - 1];  // joint action leading here (pads the header to 4 bytes)
  int32_t x[Env::kStateInts + (Env::kStateInts & 1)];
  double r_in[2];                    // reward collected on the incoming edge: ego / every other
  StatRec<NE> ego;
  StatRec<NA> oth[A - 1];
};

struct SearchTables {
  // class 0 = ego, class 1 = others
  uint8_t order[2][MAMCTS_MAX_ACTIONS];        // widening order (mamcts_reference_expand_order)
  uint32_t exp_mask[2][MAMCTS_MAX_ACTIONS + 1];// bitmask of the first k actions of the order
  uint32_t pw_min_visits[2][MAMCTS_MAX_ACTIONS];// smallest N with k <= pw_k * pow(N, pw_alpha)
};

struct SearchParams {
  void* nodes;
  uint32_t* hash;
  uint32_t* tree_nodes;   // [T] nodes in use (num_nodes_ of the tree)
  uint32_t* tree_iters;   // [T] iterations done
  const SearchTables* tab;
  const double* log_table;  // log((double)N), host libm
  unsigned long long* counters;  // iterations, rollout steps, expand steps, nodes created
  uint32_t T, max_nodes, hash_mask, iterations, first_tree_id;
  uint32_t n_ego, n_oth;
  uint32_t use_bound_est, max_depth, node_budget, rh_cap;
  uint32_t key0, key1;
  double gamma, two_c, lb, ub;
  EnvParams env;
};

// UctStatistic::choose_next_action, reference uct_statistic.h:61-76 (+ :122-132, :223-244, :256-262).
template <int NACT>
__device__ __forceinline__ uint32_t choose_action(StatRec<NACT>* s, const SearchParams& p, int cls,
                                                  uint32_t nact) {
  const uint32_t N = s->N;
  const uint32_t nexp = s->nexp;
  if (nexp < nact && N >= p.tab->pw_min_visits[cls][nexp]) {  // require_progressive_widening_total
    const uint32_t a = p.tab->order[cls][nexp];               // :64-68 (deterministic order, F2)
    s->nexp = nexp + 1;
    s->Q[a] = 0.0;                                            // ucb_statistics_[a] = UcbPair()
    s->n[a] = 0;
    return a;
  }
  const uint32_t mask = p.tab->exp_mask[cls][nexp];
  double q[NACT];
  uint32_t cnt[NACT];
#pragma unroll
  for (int a = 0; a < NACT; ++a) {
    const bool ex = (mask >> a) & 1u;
    q[a] = ex ? s->Q[a] : 0.0;
    cnt[a] = ex ? s->n[a] : 1u;
  }
  double lo = p.lb, hi = p.ub;
  if (p.use_bound_est) {  // calculate_bounds_based_on_stat :122-132
    double mn = 0.0, mx = 0.0;
    bool first = true;
#pragma unroll
    for (int a = 0; a < NACT; ++a) {
      if ((mask >> a) & 1u) {
        if (first) { mn = q[a]; mx = q[a]; first = false; }
        else { mn = q[a] < mn ? q[a] : mn; mx = q[a] < mx ? mx : q[a]; }
      }
    }
    if (mn == mx) { lo = 0.0; hi = (mx != 0.0) ? mx : 1.0; }
    else { lo = mn; hi = mx; }
  }
  const double log2N = __dmul_rn(2.0, __ldg(p.log_table + N));  // 2 * log(total_node_visits_)
  const double range = __dsub_rn(hi, lo);

  // FP32 screen. The FP64 UCB values cost two divisions and a square root per action; the arg-max
  // almost never needs that precision. Scores are first evaluated in FP32 (every operation IEEE
  // rounded: error <= 2e-7 * (|score| + 1) per score); when the first maximum beats every action
  // with different inputs by more than `delta` (>= 5x the worst-case error of two scores), the
  // exact arg-max is provably the same and is returned. Actions with bit-identical (Q, n) have
  // identical exact scores, so the strict-> first-max rule picks the same one either way. Anything
  // else falls through to the exact FP64 evaluation below, so the selected action is ALWAYS the
  // reference's.
  {
    const float rangef = (float)range, Lf = (float)log2N, c2f = (float)p.two_c;
    float sc[NACT];
    float best_sc = -1.0f;
    int bi = -1;
    bool ok = isfinite(rangef) && rangef != 0.0f;
#pragma unroll
    for (int a = 0; a < NACT; ++a) {
      sc[a] = -2.0f;
      if ((mask >> a) & 1u) {
        const float normf = (float)__dsub_rn(q[a], lo) / rangef;
        sc[a] = fmaf(c2f, sqrtf(Lf / (float)cnt[a]), normf);
        ok &= (cnt[a] != 0u) & isfinite(sc[a]);
        if (sc[a] > best_sc) { best_sc = sc[a]; bi = a; }
      }
    }
    // best_sc must be safely above DBL_MIN's role as the initial maximum (:225)
    ok &= (bi >= 0) & (best_sc > 1e-3f);
    if (ok) {
      const float delta = 4e-6f * (best_sc + 1.0f);
      double qb = 0.0;
      uint32_t nb = 0;
#pragma unroll
      for (int a = 0; a < NACT; ++a) if (a == bi) { qb = q[a]; nb = cnt[a]; }
#pragma unroll
      for (int a = 0; a < NACT; ++a) {
        if (((mask >> a) & 1u) && a != bi && sc[a] >= best_sc - delta)
          ok &= (q[a] == qb) & (cnt[a] == nb);
      }
      if (ok) return (uint32_t)bi;
    }
  }

  // exact evaluation, calculate_ucb_and_max_action :223-244
  uint32_t best = 0;
  double max_value = DBL_MIN;  // std::numeric_limits<double>::min(), :225
#pragma unroll
  for (int a = 0; a < NACT; ++a) {
    if ((mask >> a) & 1u) {
      const double norm = __ddiv_rn(__dsub_rn(q[a], lo), range);
      const double v = __dadd_rn(norm, __dmul_rn(p.two_c, __dsqrt_rn(__ddiv_rn(log2N, (double)cnt[a]))));
      if (v > max_value) { max_value = v; best = a; }
    }
  }
  return best;
}

// UctStatistic::update_statistics_from_backpropagated, reference uct_statistic.h:134-144.
template <int NACT>
__device__ __forceinline__ double backprop_stat(StatRec<NACT>* s, uint32_t act, double r, double gamma,
                                                double child_G) {
  const double G = __dadd_rn(r, __dmul_rn(gamma, child_G));
  const uint32_t cnt = s->n[act] + 1;
  s->n[act] = cnt;
  // x + (d / k) with d == +-0 is x + d for any k > 0 (the quotient keeps d's signed zero), so the
  // division is skipped when the running mean already equals the return - always the case for the
  // reward-free other agents of CrossingState.
  const double q = s->Q[act];
  const double dq = __dsub_rn(G, q);
  s->Q[act] = (dq == 0.0) ? __dadd_rn(q, dq) : __dadd_rn(q, __ddiv_rn(dq, (double)cnt));
  const uint32_t nv = s->N + 1;
  s->N = nv;
  const double v = s->V;
  const double dv = __dsub_rn(G, v);
  s->V = (dv == 0.0) ? __dadd_rn(v, dv) : __dadd_rn(v, __ddiv_rn(dv, (double)nv));
  s->G = G;
  return G;
}

template <int A>
__device__ __forceinline__ uint32_t hash_ja(uint32_t parent, const uint8_t (&ja)[A]) {
  uint32_t h = parent * 0x9E3779B1u;
#pragma unroll
  for (int a = 0; a < A; ++a) h = (h ^ ja[a]) * 0x01000193u;
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 12;
  return h;
}

// RandomHeuristic::rollout for one lane (reference random_heuristic.h:27-104); same arithmetic as
// rollout_kernel. Returns the ego return; `other` = the (shared) other-agent return.
template <class Env, int KIND>
__device__ __forceinline__ double lane_rollout(const SearchParams& p, typename Env::State st,
                                               uint32_t depth, uint32_t s_hi, uint32_t s_lo,
                                               const uint8_t* const_act, double& other,
                                               uint32_t& steps) {
  constexpr int A = Env::A;
  constexpr int S = (KIND == MAMCTS_SRC_PHILOX && A <= 8) ? 8 / A : 1;
  const double inv_n_ego = 1.0 / (double)Env::num_actions(p.env, 0);
  double m = p.gamma, ego = 0.0, prob = 1.0;
  other = 0.0;
  uint32_t it = 0;
  bool fin = Env::terminal(p.env, st) | (it >= p.rh_cap) | (depth >= p.max_depth);
  while (!fin) {
    Philox4 blk[(A + 7) / 8];
    if (KIND == MAMCTS_SRC_PHILOX) {
      if (A <= 8) {
        blk[0] = philox4x32_10(it / S, s_lo, s_hi, 0u, p.key0, p.key1);
      } else {
#pragma unroll
        for (int b = 0; b < (A + 7) / 8; ++b)
          blk[b] = philox4x32_10(it * ((A + 7) / 8) + b, s_lo, s_hi, 0u, p.key0, p.key1);
      }
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
      if (!fin) {
        int32_t act[A];
#pragma unroll
        for (int a = 0; a < A; ++a) {
          if (KIND == MAMCTS_SRC_PHILOX) {
            const uint32_t h = (A <= 8) ? philox_half(blk[0], s * A + a) : philox_half(blk[a / 8], a % 8);
            act[a] = (int32_t)philox_bounded(h, Env::num_actions(p.env, a), it * A + a, s_lo, s_hi,
                                             p.key0, p.key1);
          } else {
            act[a] = (int32_t)const_act[a];
          }
        }
        double r_ego, r_other, c;
        Env::step(p.env, st, act, r_ego, r_other, c);
        m = __dmul_rn(p.gamma, m);
        ego = __dadd_rn(ego, __dmul_rn(m, r_ego));
        other = __dmul_rn(m, r_other);
        m = __dmul_rn(m, p.gamma);
        prob = __dmul_rn(prob, inv_n_ego);
        it += 1;
        depth += 1;
        fin = Env::terminal(p.env, st) | (it >= p.rh_cap) | (depth >= p.max_depth);
      }
    }
  }
  steps = it;
  return __dmul_rn(ego, prob);
}

template <class Env, int NE, int NA, int KIND>
__global__ void __launch_bounds__(kSearchThreads) search_kernel(const __grid_constant__ SearchParams p) {
  using Node = NodeRec<Env, NE, NA>;
  constexpr int A = Env::A;
  const uint32_t tree = blockIdx.x * blockDim.x + threadIdx.x;
  if (tree >= p.T) return;
  Node* nodes = static_cast<Node*>(p.nodes) + (size_t)tree * p.max_nodes;
  uint32_t* hash = p.hash + (size_t)tree * (p.hash_mask + 1);
  uint32_t num_nodes = p.tree_nodes[tree];
  unsigned long long rollout_steps = 0, expand_steps = 0, backup_levels = 0;
  const uint32_t nodes_before = num_nodes;
  uint8_t const_act[A];
#pragma unroll
  for (int a = 0; a < A; ++a) const_act[a] = p.tab->order[a == 0 ? 0 : 1][0];  // F1: first draw

  for (uint32_t iter = 0; iter < p.iterations; ++iter) {
    uint32_t node = 0;
    bool do_rollout = false;
    typename Env::State st;
    uint32_t depth = 0;
    // ---- select & expand: mcts.h:300-305, stage_node.h:167-232 --------------------------------
    while (true) {
      Node* nd = nodes + node;
      nd->visit_count += 1;                                              // :180
      depth = nd->depth;
#pragma unroll
      for (int i = 0; i < A; ++i) { st.x[i] = (i < Env::kStateInts) ? nd->x[i] : 0; st.last[i] = 0; }
      st.flags = nd->flags;
      if (Env::terminal(p.env, st) || depth == p.max_depth || num_nodes > p.node_budget) break;  // :183-187
      uint8_t ja[A];
      ja[0] = (uint8_t)choose_action<NE>(&nd->ego, p, 0, p.n_ego);       // :190-195
#pragma unroll
      for (int j = 0; j < A - 1; ++j) ja[j + 1] = (uint8_t)choose_action<NA>(&nd->oth[j], p, 1, p.n_oth);
      // child lookup, :198
      uint32_t h = hash_ja<A>(node, ja) & p.hash_mask;
      uint32_t child = 0;
      while (true) {
        const uint32_t c = hash[h];
        if (c == 0) break;
        const Node* cn = nodes + c;
        bool same = cn->parent == node;
#pragma unroll
        for (int a = 0; a < A; ++a) same &= cn->ja[a] == ja[a];
        if (same) { child = c; break; }
        h = (h + 1) & p.hash_mask;
      }
      if (child) { node = child; continue; }                             // :199-206
      // expand, :207-231
      int32_t act[A];
#pragma unroll
      for (int a = 0; a < A; ++a) act[a] = ja[a];  // index-valued: others move by +index (F5)
      double r_ego, r_other, c0;
      Env::step(p.env, st, act, r_ego, r_other, c0);
      expand_steps += 1;
      child = num_nodes++;
      Node* cn = nodes + child;
      cn->parent = node;
      cn->visit_count = 0;
      cn->depth = depth + 1;
      cn->flags = (uint8_t)((st.flags & ~1u) | (uint32_t)Env::terminal(p.env, st));
#pragma unroll
      for (int a = 0; a < A; ++a) cn->ja[a] = ja[a];
#pragma unroll
      for (int i = 0; i < Env::kStateInts; ++i) cn->x[i] = st.x[i];
      cn->r_in[0] = r_ego;
      cn->r_in[1] = r_other;
      cn->ego.V = 0.0; cn->ego.G = 0.0; cn->ego.N = 0; cn->ego.nexp = 0;
#pragma unroll
      for (int j = 0; j < A - 1; ++j) { cn->oth[j].V = 0.0; cn->oth[j].G = 0.0; cn->oth[j].N = 0; cn->oth[j].nexp = 0; }
      hash[h] = child;
      node = child;
      depth += 1;
      st.flags = cn->flags;
      do_rollout = true;
      break;
    }
    // ---- heuristic: mcts.h:309-312, random_heuristic.h:106-134, uct_statistic.h:100-110 --------
    Node* leaf = nodes + node;
    double G[A];
    if (do_rollout) {
      double other = 0.0;
      uint32_t steps = 0;
      const double ego = lane_rollout<Env, KIND>(p, st, depth, p.first_tree_id + tree, node, const_act,
                                                 other, steps);
      rollout_steps += steps;
      leaf->ego.V = ego; leaf->ego.G = ego; leaf->ego.N += 1;
      G[0] = ego;
#pragma unroll
      for (int j = 0; j < A - 1; ++j) {
        leaf->oth[j].V = other; leaf->oth[j].G = other; leaf->oth[j].N += 1;
        G[j + 1] = other;
      }
    } else {
      G[0] = leaf->ego.G;
#pragma unroll
      for (int j = 0; j < A - 1; ++j) G[j + 1] = leaf->oth[j].G;
    }
    // ---- backpropagation: mcts.h:314-329, stage_node.h:259-271 ---------------------------------
    uint32_t child = node;
    while (true) {
      const Node* cn = nodes + child;
      const uint32_t par = cn->parent;
      if (par == kNoParent) break;
      Node* pn = nodes + par;
      const double r0 = cn->r_in[0], r1 = cn->r_in[1];
      G[0] = backprop_stat<NE>(&pn->ego, cn->ja[0], r0, p.gamma, G[0]);
#pragma unroll
      for (int j = 0; j < A - 1; ++j) G[j + 1] = backprop_stat<NA>(&pn->oth[j], cn->ja[j + 1], r1, p.gamma, G[j + 1]);
      backup_levels += 1;
      child = par;
    }
  }
  p.tree_nodes[tree] = num_nodes;
  p.tree_iters[tree] += p.iterations;
  atomicAdd(p.counters + 0, (unsigned long long)p.iterations);
  atomicAdd(p.counters + 1, rollout_steps);
  atomicAdd(p.counters + 2, expand_steps);
  atomicAdd(p.counters + 3, (unsigned long long)(num_nodes - nodes_before));
  atomicAdd(p.counters + 4, backup_levels);
}

template <class Env, int NE, int NA>
__global__ void search_reset_kernel(void* nodes_v, uint32_t* tree_nodes, uint32_t* tree_iters, uint32_t T,
                                    uint32_t max_nodes, const int32_t* root_x, EnvParams env) {
  using Node = NodeRec<Env, NE, NA>;
  const uint32_t tree = blockIdx.x * blockDim.x + threadIdx.x;
  if (tree >= T) return;
  Node* root = static_cast<Node*>(nodes_v) + (size_t)tree * max_nodes;
  root->parent = kNoParent;
  root->visit_count = 0;
  root->depth = 0;
  typename Env::State st;
  for (int i = 0; i < Env::A; ++i) { st.x[i] = (i < Env::kStateInts) ? root_x[i] : 0; st.last[i] = 0; }
  st.flags = 0;
  root->flags = (uint8_t)Env::terminal(env, st);
  for (int a = 0; a < Env::A; ++a) root->ja[a] = 0;
  for (int i = 0; i < Env::kStateInts; ++i) root->x[i] = root_x[i];
  root->r_in[0] = 0.0;
  root->r_in[1] = 0.0;
  root->ego.V = 0.0; root->ego.G = 0.0; root->ego.N = 0; root->ego.nexp = 0;
  for (int j = 0; j < Env::A - 1; ++j) { root->oth[j].V = 0.0; root->oth[j].G = 0.0; root->oth[j].N = 0; root->oth[j].nexp = 0; }
  tree_nodes[tree] = 1;
  tree_iters[tree] = 0;
}

// ---- host side ---------------------------------------------------------------------------------

}  // namespace mamcts

// Type-erased handle behind the C ABI; one concrete SearchT per compiled (environment, action-count)
// combination.
struct mamcts_search {
  mamcts_engine* e = nullptr;
  mamcts_search_config cfg;
  mamcts::SearchParams p;
  mamcts::SearchTables tab;  // host copy
  uint32_t A = 0, state_ints = 0;
  void* d_nodes = nullptr;
  uint32_t* d_hash = nullptr;
  uint32_t* d_tree_nodes = nullptr;
  uint32_t* d_tree_iters = nullptr;
  mamcts::SearchTables* d_tab = nullptr;
  double* d_log = nullptr;
  unsigned long long* d_counters = nullptr;
  int32_t* d_root_x = nullptr;
  uint32_t iters_done = 0;

  virtual ~mamcts_search() {}
  virtual size_t node_bytes() const = 0;
  virtual int launch_reset() = 0;
  virtual int launch_run(uint32_t iterations) = 0;
  // host-side views of a node record copied back from the device
  virtual void read_header(const void* node, uint32_t* parent, uint32_t* depth, uint32_t* visits,
                           uint8_t* ja) const = 0;
  virtual void read_stat(const void* node, uint32_t agent, double* value, double* latest,
                         uint32_t* visits, uint32_t* nexp, double* q, uint32_t* n) const = 0;
  virtual void root_gather(mamcts::RootGather* g) const = 0;
};

namespace mamcts {

template <class Env, int NE, int NA>
struct SearchT final : public mamcts_search {
  using Node = NodeRec<Env, NE, NA>;

  size_t node_bytes() const override { return sizeof(Node); }

  int launch_reset() override {
    const uint32_t T = cfg.num_trees;
    MAMCTS_CUDA(e, cudaMemsetAsync(d_hash, 0, sizeof(uint32_t) * (size_t)T * (p.hash_mask + 1), e->stream));
    MAMCTS_CUDA(e, cudaMemsetAsync(d_counters, 0, sizeof(unsigned long long) * 8, e->stream));
    search_reset_kernel<Env, NE, NA><<<(T + 255) / 256, 256, 0, e->stream>>>(
        d_nodes, d_tree_nodes, d_tree_iters, T, p.max_nodes, d_root_x, p.env);
    e->launches++;
    MAMCTS_CUDA(e, cudaGetLastError());
    iters_done = 0;
    return MAMCTS_OK;
  }

  int launch_run(uint32_t iterations) override {
    SearchParams q = p;
    q.iterations = iterations;
    const uint32_t T = cfg.num_trees;
    const int grid = (int)((T + kSearchThreads - 1) / kSearchThreads);
    if (cfg.action_source == MAMCTS_SRC_PHILOX)
      search_kernel<Env, NE, NA, MAMCTS_SRC_PHILOX><<<grid, kSearchThreads, 0, e->stream>>>(q);
    else
      search_kernel<Env, NE, NA, MAMCTS_SRC_REFERENCE_CONST><<<grid, kSearchThreads, 0, e->stream>>>(q);
    e->launches++;
    MAMCTS_CUDA(e, cudaGetLastError());
    return MAMCTS_OK;
  }

  void read_header(const void* node, uint32_t* parent, uint32_t* depth, uint32_t* visits,
                   uint8_t* ja) const override {
    const Node* nd = static_cast<const Node*>(node);
    *parent = nd->parent;
    *depth = nd->depth;
    *visits = nd->visit_count;
    for (int a = 0; a < Env::A; ++a) ja[a] = nd->ja[a];
  }

  template <int NACT>
  static void read_stat_rec(const StatRec<NACT>& st, double* value, double* latest, uint32_t* visits,
                            uint32_t* nexp, double* q, uint32_t* n) {
    *value = st.V;
    *latest = st.G;
    *visits = st.N;
    *nexp = st.nexp;
    for (int a = 0; a < NACT; ++a) { q[a] = st.Q[a]; n[a] = st.n[a]; }
  }

  void read_stat(const void* node, uint32_t agent, double* value, double* latest, uint32_t* visits,
                 uint32_t* nexp, double* q, uint32_t* n) const override {
    const Node* nd = static_cast<const Node*>(node);
    if (agent == 0) read_stat_rec<NE>(nd->ego, value, latest, visits, nexp, q, n);
    else read_stat_rec<NA>(nd->oth[agent - 1], value, latest, visits, nexp, q, n);
  }

  void root_gather(RootGather* g) const override {
    const char* base = static_cast<const char*>(d_nodes);
    const size_t tree_bytes = sizeof(Node) * (size_t)p.max_nodes;
    g->q = reinterpret_cast<const double*>(base + offsetof(Node, ego) + offsetof(StatRec<NE>, Q));
    g->n = reinterpret_cast<const uint32_t*>(base + offsetof(Node, ego) + offsetof(StatRec<NE>, n));
    g->visits = reinterpret_cast<const uint32_t*>(base + offsetof(Node, ego) + offsetof(StatRec<NE>, N));
    g->root_slot = nullptr;
    g->stride_q = tree_bytes / sizeof(double);
    g->stride_n = tree_bytes / sizeof(uint32_t);
    g->stride_v = tree_bytes / sizeof(uint32_t);
    g->max_actions = NE;
  }
};

static mamcts_search* make_search(const mamcts_search_config& cfg, uint32_t n_ego, uint32_t n_oth) {
  if (cfg.env == MAMCTS_ENV_SIMPLE) return new SearchT<SimpleEnv, 2, 2>();
  if (n_ego == 4 && n_oth == 7) {
    switch (cfg.crossing.num_other_agents) {
      case 1: return new SearchT<CrossingEnv<1>, 4, 7>();
      case 2: return new SearchT<CrossingEnv<2>, 4, 7>();
      case 3: return new SearchT<CrossingEnv<3>, 4, 7>();
      case 7: return new SearchT<CrossingEnv<7>, 4, 7>();
      default: return nullptr;
    }
  }
  return nullptr;
}

static void free_search(mamcts_search* s) {
  if (!s) return;
  cudaSetDevice(s->e->device);
  cudaStreamSynchronize(s->e->stream);
  cudaFree(s->d_nodes); cudaFree(s->d_hash); cudaFree(s->d_tree_nodes); cudaFree(s->d_tree_iters);
  cudaFree(s->d_tab); cudaFree(s->d_log); cudaFree(s->d_counters); cudaFree(s->d_root_x);
  delete s;
}

}  // namespace mamcts

using namespace mamcts;

extern "C" {

int mamcts_search_create(mamcts_engine* e, const mamcts_search_config* cfg, const int32_t* root_x,
                         const int32_t* root_last_action, mamcts_search** out) {
  (void)root_last_action;  // last_action does not influence UctStatistic searches (no hypothesis policy)
  if (!e) return fail(nullptr, MAMCTS_ERR_INVALID, "mamcts_search_create: null engine");
  std::lock_guard<std::mutex> lock(e->mu);
  if (!cfg || !root_x || !out) return fail(e, MAMCTS_ERR_INVALID, "mamcts_search_create: null argument");
  if (cfg->num_trees == 0) return fail(e, MAMCTS_ERR_INVALID, "mamcts_search_create: num_trees == 0");
  if (cfg->action_source != MAMCTS_SRC_PHILOX && cfg->action_source != MAMCTS_SRC_REFERENCE_CONST)
    return fail(e, MAMCTS_ERR_INVALID, "mamcts_search_create: action source must be REFERENCE_CONST or PHILOX");
  if (cfg->rollouts_per_leaf != 1)
    return fail(e, MAMCTS_ERR_UNSUPPORTED, "mamcts_search_create: rollouts_per_leaf must be 1 in this build");
  const mamcts_params& mp = cfg->mcts;
  if (!(mp.uct_pw_k >= 0.0) || !(mp.uct_pw_alpha >= 0.0))
    return fail(e, MAMCTS_ERR_INVALID, "mamcts_search_create: progressive widening k, alpha must be >= 0");
  if (mp.max_number_of_iterations == 0) return fail(e, MAMCTS_ERR_INVALID, "mamcts_search_create: zero iterations");
  uint32_t n_ego, n_oth, A;
  EnvParams ep;
  std::memset(&ep, 0, sizeof(ep));
  if (cfg->env == MAMCTS_ENV_SIMPLE) {
    n_ego = n_oth = 2; A = 2;
    ep.win_len = cfg->simple.winning_state_length;
    ep.lose_len = cfg->simple.loosing_state_length;
    ep.n_ego_actions = ep.n_other_actions = 2;
  } else if (cfg->env == MAMCTS_ENV_CROSSING_I32) {
    const mamcts_crossing_params& cp = cfg->crossing;
    n_ego = (uint32_t)(cp.max_velocity_ego - cp.min_velocity_ego + 1);
    n_oth = cp.num_other_actions;
    A = cp.num_other_agents + 1;
    ep.r_goal = cp.reward_goal_reached; ep.r_coll = cp.reward_collision; ep.r_step = cp.reward_step;
    ep.min_v_ego = cp.min_velocity_ego;
    ep.crossing_point = (cp.chain_length - 1) / 2 + 1;
    ep.goal_pos = cp.ego_goal_pos;
    ep.cost_only_collision = cp.cost_only_collision;
    ep.n_ego_actions = n_ego; ep.n_other_actions = n_oth;
  } else {
    return fail(e, MAMCTS_ERR_INVALID, "mamcts_search_create: unknown environment");
  }
  mamcts_search* s = make_search(*cfg, n_ego, n_oth);
  if (!s)
    return fail(e, MAMCTS_ERR_UNSUPPORTED,
                "mamcts_search_create: compiled for SimpleState and CrossingState<int> with 4 ego / 7 other "
                "actions and 1, 2, 3 or 7 other agents");
  s->e = e;
  s->cfg = *cfg;
  s->A = A;
  MAMCTS_CUDA(e, cudaSetDevice(e->device));

  // host tables: widening order (libstdc++), expanded masks, widening thresholds (libm pow), log table
  SearchTables& tab = s->tab;
  std::memset(&tab, 0, sizeof(tab));
  const uint32_t nact[2] = {n_ego, n_oth};
  const uint32_t max_iter = mp.max_number_of_iterations;
  for (int cls = 0; cls < 2; ++cls) {
    uint32_t order[MAMCTS_MAX_ACTIONS];
    int rc = mamcts_reference_expand_order(mp.random_seed, nact[cls], order);
    if (rc != MAMCTS_OK) { delete s; return rc; }
    uint32_t mask = 0;
    tab.exp_mask[cls][0] = 0;
    for (uint32_t k = 0; k < nact[cls]; ++k) {
      tab.order[cls][k] = (uint8_t)order[k];
      mask |= 1u << order[k];
      tab.exp_mask[cls][k + 1] = mask;
      // require_progressive_widening_total (uct_statistic.h:256-262): k <= pw_k * pow(N, pw_alpha);
      // non-decreasing in N for k, alpha >= 0, so the first N that satisfies it is a threshold.
      uint32_t thr = 0xFFFFFFFFu;
      for (uint32_t N = 0; N <= max_iter + 1; ++N) {
        const double term = mp.uct_pw_k * std::pow((double)N, mp.uct_pw_alpha);
        if ((double)k <= term) { thr = N; break; }
      }
      tab.pw_min_visits[cls][k] = thr;
    }
  }
  std::vector<double> logs((size_t)max_iter + 2);
  for (size_t N = 0; N < logs.size(); ++N) logs[N] = std::log((double)N);  // log(total_node_visits_)

  uint32_t max_nodes = cfg->max_nodes_per_tree;
  const uint64_t needed = std::min<uint64_t>((uint64_t)mp.max_number_of_nodes + 1, (uint64_t)max_iter + 1);
  if (max_nodes == 0) max_nodes = (uint32_t)needed;
  if (max_nodes < needed) {
    delete s;
    return fail(e, MAMCTS_ERR_INVALID, "mamcts_search_create: max_nodes_per_tree below min(MAX_NUMBER_OF_NODES, iterations) + 1");
  }
  uint32_t hcap = 16;
  while (hcap < 2 * max_nodes) hcap <<= 1;

  const uint32_t T = cfg->num_trees;
  cudaError_t err = cudaSuccess;
  auto alloc = [&](void** ptr, size_t bytes) { if (err == cudaSuccess) err = cudaMalloc(ptr, bytes); };
  alloc(&s->d_nodes, s->node_bytes() * (size_t)T * max_nodes);
  alloc((void**)&s->d_hash, sizeof(uint32_t) * (size_t)T * hcap);
  alloc((void**)&s->d_tree_nodes, sizeof(uint32_t) * T);
  alloc((void**)&s->d_tree_iters, sizeof(uint32_t) * T);
  alloc((void**)&s->d_tab, sizeof(SearchTables));
  alloc((void**)&s->d_log, sizeof(double) * logs.size());
  alloc((void**)&s->d_counters, sizeof(unsigned long long) * 8);
  alloc((void**)&s->d_root_x, sizeof(int32_t) * MAMCTS_MAX_AGENTS);
  if (err != cudaSuccess) {
    free_search(s);
    cudaGetLastError();
    return fail(e, MAMCTS_ERR_NOMEM, std::string("mamcts_search_create: device allocation failed: ") + cudaGetErrorString(err));
  }
  SearchParams& p = s->p;
  std::memset(&p, 0, sizeof(p));
  p.nodes = s->d_nodes; p.hash = s->d_hash; p.tree_nodes = s->d_tree_nodes; p.tree_iters = s->d_tree_iters;
  p.tab = s->d_tab; p.log_table = s->d_log; p.counters = s->d_counters;
  p.T = T; p.max_nodes = max_nodes; p.hash_mask = hcap - 1; p.first_tree_id = cfg->first_tree_id;
  p.n_ego = n_ego; p.n_oth = n_oth;
  p.use_bound_est = mp.use_bound_estimation; p.max_depth = mp.max_search_depth;
  p.node_budget = mp.max_number_of_nodes; p.rh_cap = mp.rh_max_number_of_iterations;
  p.key0 = (uint32_t)cfg->philox_seed; p.key1 = (uint32_t)(cfg->philox_seed >> 32);
  p.gamma = mp.discount_factor; p.two_c = 2 * mp.uct_exploration_constant;
  p.lb = mp.uct_lower_bound; p.ub = mp.uct_upper_bound;
  p.env = ep;
  s->state_ints = (cfg->env == MAMCTS_ENV_SIMPLE) ? 1 : A;
  int32_t rx[MAMCTS_MAX_AGENTS] = {0};
  for (uint32_t i = 0; i < s->state_ints; ++i) rx[i] = root_x[i];
  err = cudaMemcpyAsync(s->d_tab, &tab, sizeof(tab), cudaMemcpyHostToDevice, e->stream);
  if (err == cudaSuccess) err = cudaMemcpyAsync(s->d_log, logs.data(), sizeof(double) * logs.size(), cudaMemcpyHostToDevice, e->stream);
  if (err == cudaSuccess) err = cudaMemcpyAsync(s->d_root_x, rx, sizeof(rx), cudaMemcpyHostToDevice, e->stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);  // host sources go out of scope
  if (err != cudaSuccess) { free_search(s); return cuda_fail(e, err, "mamcts_search_create uploads"); }
  int rc = s->launch_reset();
  if (rc != MAMCTS_OK) { free_search(s); return rc; }
  *out = s;
  return MAMCTS_OK;
}

int mamcts_search_destroy(mamcts_search* s) {
  if (!s) return MAMCTS_OK;
  std::lock_guard<std::mutex> lock(s->e->mu);
  free_search(s);
  return MAMCTS_OK;
}

int mamcts_search_reset(mamcts_search* s, const int32_t* root_x, const int32_t* root_last_action) {
  (void)root_last_action;
  if (!s) return fail(nullptr, MAMCTS_ERR_INVALID, "mamcts_search_reset: null handle");
  mamcts_engine* e = s->e;
  std::lock_guard<std::mutex> lock(e->mu);
  MAMCTS_CUDA(e, cudaSetDevice(e->device));
  if (root_x) {  // NULL: re-root at the state already resident on the device (no copy)
    int32_t rx[MAMCTS_MAX_AGENTS] = {0};
    for (uint32_t i = 0; i < s->state_ints; ++i) rx[i] = root_x[i];
    MAMCTS_CUDA(e, cudaMemcpyAsync(s->d_root_x, rx, sizeof(rx), cudaMemcpyHostToDevice, e->stream));
    MAMCTS_CUDA(e, cudaStreamSynchronize(e->stream));
  }
  return s->launch_reset();
}

int mamcts_search_run(mamcts_search* s, uint32_t iterations) {
  if (!s) return fail(nullptr, MAMCTS_ERR_INVALID, "mamcts_search_run: null handle");
  mamcts_engine* e = s->e;
  std::lock_guard<std::mutex> lock(e->mu);
  if ((uint64_t)s->iters_done + iterations > s->cfg.mcts.max_number_of_iterations)
    return fail(e, MAMCTS_ERR_INVALID, "mamcts_search_run: exceeds MAX_NUMBER_OF_ITERATIONS of the configuration");
  if (iterations == 0) return MAMCTS_OK;
  MAMCTS_CUDA(e, cudaSetDevice(e->device));
  int rc = s->launch_run(iterations);
  if (rc == MAMCTS_OK) s->iters_done += iterations;
  return rc;
}

int mamcts_search_counters(mamcts_search* s, uint64_t* iterations, uint64_t* rollout_steps,
                           uint64_t* expand_steps, uint64_t* nodes, uint64_t* backup_levels) {
  if (!s) return fail(nullptr, MAMCTS_ERR_INVALID, "mamcts_search_counters: null handle");
  mamcts_engine* e = s->e;
  std::lock_guard<std::mutex> lock(e->mu);
  MAMCTS_CUDA(e, cudaSetDevice(e->device));
  unsigned long long h[8];
  MAMCTS_CUDA(e, cudaMemcpyAsync(h, s->d_counters, sizeof(h), cudaMemcpyDeviceToHost, e->stream));
  MAMCTS_CUDA(e, cudaStreamSynchronize(e->stream));
  if (iterations) *iterations = h[0];
  if (rollout_steps) *rollout_steps = h[1];
  if (expand_steps) *expand_steps = h[2];
  if (nodes) *nodes = h[3] + s->cfg.num_trees;  // + the roots
  if (backup_levels) *backup_levels = h[4];
  return MAMCTS_OK;
}

int mamcts_search_root_stats(mamcts_search* s, uint32_t tree, uint32_t agent, double* value,
                             uint32_t* total_visits, double* q, uint32_t* n, uint32_t* expanded_mask,
                             uint32_t* num_nodes) {
  if (!s) return fail(nullptr, MAMCTS_ERR_INVALID, "mamcts_search_root_stats: null handle");
  mamcts_engine* e = s->e;
  std::lock_guard<std::mutex> lock(e->mu);
  if (tree >= s->cfg.num_trees || agent >= s->A) return fail(e, MAMCTS_ERR_INVALID, "mamcts_search_root_stats: index out of range");
  MAMCTS_CUDA(e, cudaSetDevice(e->device));
  std::vector<char> node(s->node_bytes());
  const char* src = static_cast<const char*>(s->d_nodes) + s->node_bytes() * (size_t)tree * s->p.max_nodes;
  MAMCTS_CUDA(e, cudaMemcpyAsync(node.data(), src, node.size(), cudaMemcpyDeviceToHost, e->stream));
  uint32_t nn = 0;
  MAMCTS_CUDA(e, cudaMemcpyAsync(&nn, s->d_tree_nodes + tree, sizeof(nn), cudaMemcpyDeviceToHost, e->stream));
  MAMCTS_CUDA(e, cudaStreamSynchronize(e->stream));
  double v, l, qq[MAMCTS_MAX_ACTIONS];
  uint32_t vis, nexp, nn_a[MAMCTS_MAX_ACTIONS + 1];
  s->read_stat(node.data(), agent, &v, &l, &vis, &nexp, qq, nn_a);
  const int cls = agent == 0 ? 0 : 1;
  const uint32_t nact = agent == 0 ? s->p.n_ego : s->p.n_oth;
  const uint32_t mask = s->tab.exp_mask[cls][std::min(nexp, nact)];
  if (value) *value = v;
  if (total_visits) *total_visits = vis;
  for (uint32_t a = 0; a < nact; ++a) {
    const bool ex = (mask >> a) & 1u;
    if (q) q[a] = ex ? qq[a] : 0.0;
    if (n) n[a] = ex ? nn_a[a] : 0u;
  }
  if (expanded_mask) *expanded_mask = mask;
  if (num_nodes) *num_nodes = nn;
  return MAMCTS_OK;
}

int mamcts_search_merge_roots(mamcts_search* s, double* merged_q, uint64_t* merged_n,
                              uint32_t* merged_occurrences, uint64_t* merged_total_visits,
                              double* merged_value, uint32_t* best_action) {
  if (!s) return fail(nullptr, MAMCTS_ERR_INVALID, "mamcts_search_merge_roots: null handle");
  mamcts_engine* e = s->e;
  std::lock_guard<std::mutex> lock(e->mu);
  MAMCTS_CUDA(e, cudaSetDevice(e->device));
  RootGather g;
  std::memset(&g, 0, sizeof(g));
  s->root_gather(&g);
  return root_merge_device(e, g, s->cfg.num_trees, s->p.n_ego, merged_q, merged_n, merged_occurrences,
                           merged_total_visits, merged_value, best_action);
}

int mamcts_search_dump_tree(mamcts_search* s, uint32_t tree, double* records,
                            uint32_t capacity_records, uint32_t* num_records) {
  if (!s || !records || !num_records) return fail(s ? s->e : nullptr, MAMCTS_ERR_INVALID, "mamcts_search_dump_tree: null argument");
  mamcts_engine* e = s->e;
  std::lock_guard<std::mutex> lock(e->mu);
  if (tree >= s->cfg.num_trees) return fail(e, MAMCTS_ERR_INVALID, "mamcts_search_dump_tree: tree out of range");
  MAMCTS_CUDA(e, cudaSetDevice(e->device));
  uint32_t nn = 0;
  MAMCTS_CUDA(e, cudaMemcpyAsync(&nn, s->d_tree_nodes + tree, sizeof(nn), cudaMemcpyDeviceToHost, e->stream));
  MAMCTS_CUDA(e, cudaStreamSynchronize(e->stream));
  const size_t nb = s->node_bytes();
  std::vector<char> arena(nb * nn);
  const char* src = static_cast<const char*>(s->d_nodes) + nb * (size_t)tree * s->p.max_nodes;
  MAMCTS_CUDA(e, cudaMemcpyAsync(arena.data(), src, arena.size(), cudaMemcpyDeviceToHost, e->stream));
  MAMCTS_CUDA(e, cudaStreamSynchronize(e->stream));
  // Export only (like StageNode::printTree, reference stage_node.h:322-391): children lists from the
  // parent pointers, then a depth-first walk with children in ascending joint-action code.
  const uint32_t A = s->A;
  const uint32_t max_actions = std::max(s->p.n_ego, s->p.n_oth);
  const uint32_t stride = 4 + A * (3 + 2 * max_actions);
  std::vector<double> code(nn, -1.0);
  std::vector<std::vector<uint32_t>> kids(nn);
  for (uint32_t i = 1; i < nn; ++i) {
    uint32_t parent, depth, visits;
    uint8_t ja[MAMCTS_MAX_AGENTS];
    s->read_header(arena.data() + nb * i, &parent, &depth, &visits, ja);
    double c = 0.0, mul = 1.0;
    for (uint32_t a = 0; a < A; ++a) { c += mul * ja[a]; mul *= max_actions; }
    code[i] = c;
    kids[parent].push_back(i);
  }
  for (auto& k : kids) std::sort(k.begin(), k.end(), [&](uint32_t a, uint32_t b) { return code[a] < code[b]; });
  uint32_t count = 0;
  std::vector<std::pair<uint32_t, size_t>> stack;  // (node, next child position)
  stack.push_back({0u, 0});
  auto emit = [&](uint32_t i) {
    if (count < capacity_records) {
      double* r = records + (size_t)count * stride;
      uint32_t parent, depth, visits;
      uint8_t ja[MAMCTS_MAX_AGENTS];
      s->read_header(arena.data() + nb * i, &parent, &depth, &visits, ja);
      r[0] = depth; r[1] = code[i]; r[2] = visits; r[3] = (double)kids[i].size();
      for (uint32_t a = 0; a < A; ++a) {
        double* o = r + 4 + a * (3 + 2 * max_actions);
        double v, l, qq[MAMCTS_MAX_ACTIONS];
        uint32_t vis, nexp, cnt[MAMCTS_MAX_ACTIONS + 1];
        s->read_stat(arena.data() + nb * i, a, &v, &l, &vis, &nexp, qq, cnt);
        const uint32_t nact = a == 0 ? s->p.n_ego : s->p.n_oth;
        const uint32_t mask = s->tab.exp_mask[a == 0 ? 0 : 1][std::min(nexp, nact)];
        o[0] = vis; o[1] = v; o[2] = l;
        for (uint32_t k = 0; k < max_actions; ++k) {
          const bool ex = k < nact && ((mask >> k) & 1u);
          o[3 + k] = ex ? (double)cnt[k] : -1.0;
          o[3 + max_actions + k] = ex ? qq[k] : 0.0;
        }
      }
    }
    ++count;
  };
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

}  // extern "C"
