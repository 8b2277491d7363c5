// search_kernel.cuh - device side of the device-resident root-parallel MCTS (see search.cu).
//
// One lane owns one tree and runs whole iterations without returning to the host:
//   select/expand  mcts.h:300-305, stage_node.h:167-232, uct_statistic.h:61-76,122-132,223-262
//   rollout        random_heuristic.h:27-104 (+ leaf update uct_statistic.h:100-110)
//   backup         mcts.h:314-329, stage_node.h:259-271, uct_statistic.h:134-144
//
// Latency design (the kernel is bound by the dependent memory round trips of one tree, not by
// bandwidth): every level of the descent costs ONE round trip - all fields of a node record are
// loaded together into registers, the child index comes from a direct per-node child table when
// the joint-action space is small (2 agents) or from a hash probe whose target node is prefetched;
// the descent records its path so that the backup needs no pointer chasing and its statistic lines
// are prefetched in one batch.
#pragma once
#include <cfloat>
#include <stdint.h>

#include "env.cuh"
#include "mamcts_b200.h"
#include "philox.cuh"

namespace mamcts {

constexpr uint32_t kNoParent = 0xFFFFFFFFu;
constexpr int kSearchThreads = 64;
constexpr int kPathMax = 32;  // recorded path levels; deeper leaves fall back to parent pointers

template <int NACT>
struct alignas(8) StatRec {
  double V;        // value_
  double G;        // latest_return_
  uint32_t N;      // total_node_visits_
  uint32_t nexp;   // ucb_statistics_.size(): the first nexp entries of the widening order
  double Q[NACT];  // action_value_
  uint32_t n[NACT + (NACT & 1)];  // action_count_
};

// NCHILD > 0: direct child table indexed by the joint-action code (ego * NA + other), CIDX-typed
// node indices (0 = absent). NCHILD == 0: children are found through the per-tree hash table.
template <class Env, int NE, int NA, int NCHILD, class CIDX>
struct alignas(32) NodeRec {
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
  CIDX child[NCHILD > 0 ? NCHILD : 1];
};

struct SearchTables {
  // class 0 = ego, class 1 = others
  uint8_t order[2][MAMCTS_MAX_ACTIONS];         // widening order (mamcts_reference_expand_order)
  uint32_t exp_mask[2][MAMCTS_MAX_ACTIONS + 1]; // bitmask of the first k actions of the order
  uint32_t pw_min_visits[2][MAMCTS_MAX_ACTIONS];// smallest N with k <= pw_k * pow(N, pw_alpha)
};

// The selection tables are read at every level with an index that comes out of the node record:
// staged in shared memory once per block, the lookup is a fixed ~25-cycle access instead of a global
// load that competes with the tree records for the L1.
__device__ __forceinline__ void stage_tables(SearchTables& dst, const SearchTables* src) {
  static_assert(sizeof(SearchTables) % 4 == 0, "copied by words");
  const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
  uint32_t* d = reinterpret_cast<uint32_t*>(&dst);
  for (uint32_t i = threadIdx.x; i < sizeof(SearchTables) / 4; i += blockDim.x) d[i] = s[i];
  __syncthreads();
}

struct SearchParams {
  void* nodes;
  uint32_t* hash;
  uint32_t* tree_nodes;   // [T] nodes in use (num_nodes_ of the tree)
  uint32_t* tree_iters;   // [T] iterations done
  const SearchTables* tab;
  const double* log_table;  // log((double)N), host libm
  const double* disc;       // rollout discount / probability table (engine.h: ensure_discount_table)
  uint32_t disc_len;
  unsigned long long* counters;  // iterations, rollout steps, expand steps, nodes, backup levels, exact UCB
  uint32_t T, max_nodes, hash_mask, iterations, first_tree_id;
  uint32_t lanes_per_warp;  // trees per warp: lanes >= this exit at once (see search.cu: divergence)
  uint32_t n_ego, n_oth;
  uint32_t use_bound_est, max_depth, node_budget, rh_cap;
  uint32_t key0, key1;
  uint32_t rollouts_per_leaf;   // K >= 1 (power of two <= 16): the leaf value is the mean of K rollouts
  double gamma, two_c, lb, ub;
  EnvParams env;
};

__device__ __forceinline__ void prefetch_l1(const void* ptr) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr));
}

// Register image of the part of a statistic the selection reads. All loads are unconditional and
// independent of each other, so one node costs one memory round trip.
template <int NACT>
struct StatRegs {
  uint32_t N, nexp;
  double V;
  double Q[NACT];
  uint32_t n[NACT];
};

// What one backup level needs from a statistic. The descent has all four values in registers (a tree
// is touched by one lane only, so they are still current when the backup reaches the level); they are
// recorded with the path and the backup does no loads at all.
struct BackRegs {
  double q, v;
  uint32_t cnt0, nv0;
};

// The statistic of an agent whose values are known to be +0.0 (a reward-free agent, see
// Env::kOtherRewardZero): only the counts are read.
template <int NACT>
__device__ __forceinline__ StatRegs<NACT> load_stat_counts(const StatRec<NACT>* s) {
  StatRegs<NACT> r;
  r.N = s->N;
  r.nexp = s->nexp;
  r.V = 0.0;
#pragma unroll
  for (int a = 0; a < NACT; ++a) { r.Q[a] = 0.0; r.n[a] = s->n[a]; }
  return r;
}

template <int NACT>
__device__ __forceinline__ StatRegs<NACT> load_stat(const StatRec<NACT>* s) {
  StatRegs<NACT> r;
  r.N = s->N;
  r.nexp = s->nexp;
  r.V = s->V;
#pragma unroll
  for (int a = 0; a < NACT; ++a) { r.Q[a] = s->Q[a]; r.n[a] = s->n[a]; }
  return r;
}

// 1 / x and 1 / sqrt(x) in FP64 to ~2^-50 relative: the hardware's approximations (rcp.approx.ftz.f64 and
// rsqrt.approx.ftz.f64: relative error <= 2^-22, PTX ISA) refined by two Newton steps each (error e -> ~1.5 e^2 per
// step, plus a few roundings). For normal, finite x only (the callers check). Used by the second screening stage
// below, never for a value that is stored.
__device__ __forceinline__ double fast_rcp64(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#pragma unroll
  for (int k = 0; k < 2; ++k) { const double e = __fma_rn(-x, y, 1.0); y = __fma_rn(y, e, y); }
  return y;
}
__device__ __forceinline__ double fast_rsqrt64(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double h = 0.5 * x;
#pragma unroll
  for (int k = 0; k < 2; ++k) { const double e = __fma_rn(-h * y, y, 0.5); y = __fma_rn(y, e, y); }
  return y;
}

// Exact UCB arg-max, calculate_ucb_and_max_action (reference uct_statistic.h:223-244). Cold: only
// reached when the FP32 screen of choose_action cannot prove the winner, so it is kept out of line
// (the hot loop must stay small enough for the instruction cache).
//
// Second screening stage (try_fast: the FP32 screen's preconditions held and `mask` is its set of near-maximal
// candidates, every other action provably below the winner). The FP32 screen cannot separate scores closer than
// ~1e-5 - at the root of a converged tree, where two actions have nearly the same value and 2 000 visits each, that
// is most of its failures - but FP64 with approximate reciprocal / square root can, at a quarter of the cost of the
// correctly rounded divisions and square roots (this path runs with one or two lanes of a warp while the others
// wait): score' = (Q - lo) * rcp(range) + 2c * sqrt(L) * rsqrt(n), |score' - score| <= 1e-14 (|norm| + |bonus|) (two
// Newton steps from 2^-22: 2^-50 per factor, a handful of roundings), the reference's own five roundings stay within
// 1e-15 (|norm| + |bonus|). A candidate whose score' exceeds every other candidate's by more than
// 1e-11 (max(|norm| + |bonus|) + 1) - a thousand times those bounds - therefore has the strictly largest reference
// score (and a positive one, above the DBL_MIN the reference's loop starts from) and is what that loop returns.
// Anything closer, or not finite, goes on to the exact evaluation. Bit 31 of the result says the exact one ran.
template <int NACT>
__device__ __noinline__ uint32_t exact_ucb_argmax(StatRegs<NACT> r, uint32_t mask, double lo, double range,
                                                  double log2N, double two_c, bool try_fast) {
  if (try_fast) {
    const double ir = fast_rcp64(range);
    const double sq = two_c * (log2N * fast_rsqrt64(log2N));
    double s_best = -1.0, s_next = -1.0, mag = 0.0;   // mag: the largest |norm| + |bonus| - what the error bounds scale with
    int bi = -1;
    bool clean = true;
#pragma unroll
    for (int a = 0; a < NACT; ++a) {
      if ((mask >> a) & 1u) {
        const double nrm = (r.Q[a] - lo) * ir, bonus = sq * fast_rsqrt64((double)r.n[a]);
        const double s = nrm + bonus, m = fabs(nrm) + fabs(bonus);
        clean &= (s == s) & (m < 1e300);               // not NaN, not huge
        mag = m > mag ? m : mag;
        if (s > s_best) { s_next = s_best; s_best = s; bi = a; }
        else if (s > s_next) s_next = s;
      }
    }
    if (clean && bi >= 0 && s_best > 0.0 && s_best - s_next > 1e-11 * (mag + 1.0)) return (uint32_t)bi;
  }
  // the scores of all candidates first, as independent chains (two divisions and a square root each: the
  // latency of this path is what the rest of the warp waits for), then the reference's loop over them
  double v[NACT];
#pragma unroll
  for (int a = 0; a < NACT; ++a) {
    v[a] = 0.0;
    if ((mask >> a) & 1u) {
      const double norm = __ddiv_rn(__dsub_rn(r.Q[a], lo), range);
      v[a] = __dadd_rn(norm, __dmul_rn(two_c, __dsqrt_rn(__ddiv_rn(log2N, (double)r.n[a]))));
    }
  }
  uint32_t best = 0;
  double max_value = DBL_MIN;  // std::numeric_limits<double>::min(), :225
#pragma unroll
  for (int a = 0; a < NACT; ++a) {
    if (((mask >> a) & 1u) && v[a] > max_value) { max_value = v[a]; best = a; }
  }
  return best | 0x80000000u;
}

// UctStatistic::choose_next_action, reference uct_statistic.h:61-76 (+ :122-132, :223-244, :256-262),
// on the register image of a statistic; no memory is written. `widened` = the action was newly
// expanded (ucb_statistics_[a] = UcbPair(), the caller records it in its own record layout).
// Also returns (in `back`) the values the backup of this level will need for the chosen action.
template <int NACT, class CTR = unsigned long long>
__device__ __forceinline__ uint32_t select_action(const StatRegs<NACT>& r, const SearchParams& p,
                                                  const SearchTables& tab, int cls,
                                                  uint32_t nact, CTR& exact_evals,
                                                  BackRegs& back, bool& widened,
                                                  bool values_known_equal = false) {
  const uint32_t N = r.N;
  const uint32_t nexp = r.nexp;
  back.nv0 = N;
  back.v = r.V;
  widened = false;
  if (nexp < nact && N >= tab.pw_min_visits[cls][nexp]) {  // require_progressive_widening_total
    widened = true;
    back.cnt0 = 0;
    back.q = 0.0;
    return tab.order[cls][nexp];                              // :64-68 (deterministic order, F2)
  }
  const uint32_t mask = tab.exp_mask[cls][nexp];
  // Equal action values (always the case for an agent that never receives a reward, e.g. the other
  // agents of CrossingState): with bound estimation the normalised value is the same for every action
  // (0 when the common value is 0, else exactly q / q = 1), so the UCB score is
  // norm + fl(2c * fl(sqrt(fl(2 ln N / n_a)))), a non-increasing function of n_a. It is STRICTLY
  // decreasing under the conditions checked here - N >= 2 (2 ln N >= 1.38), 0.01 <= 2c <= 1e6 (every bonus
  // finite: with a huge exploration constant all bonuses overflow to +inf, the scores tie and the
  // reference's strict > keeps the FIRST expanded index, not the least visited), n_a <= 2^20: two
  // counts differ by a factor >= 1 + 2^-20, their bonuses by a relative 4.7e-7 and an absolute
  // >= 5e-12 * max(1, bonus), far beyond the four roundings (1.1e-16 relative each) and the ulp of the
  // sum - and every score exceeds DBL_MIN, the initial maximum of the reference's loop. So the first
  // maximum of the exact scores is the first minimum of the counts; no floating-point work is needed.
  if (p.use_bound_est && N >= 2u && N <= (1u << 20) && p.two_c >= 0.01 && p.two_c <= 1e6) {   // every n_a <= N
    long long q_first = 0;
    bool first = true, same = true;
    if (!values_known_equal) {
#pragma unroll
      for (int a = 0; a < NACT; ++a) {
        if ((mask >> a) & 1u) {
          const long long bits = __double_as_longlong(r.Q[a]);
          if (first) { q_first = bits; first = false; }
          same &= (bits == q_first);
        }
      }
    } else {
      q_first = __double_as_longlong(r.Q[0]);
    }
    if (same) {
      // first minimum of the counts: minimum of the keys n_a * 8 + a (NACT <= 8, n_a <= 2^20)
      static_assert(NACT <= 8, "three key bits for the action");
      uint32_t key = 0xFFFFFFFFu;
#pragma unroll
      for (int a = 0; a < NACT; ++a) {
        const uint32_t k = ((mask >> a) & 1u) ? ((r.n[a] << 3) | (uint32_t)a) : 0xFFFFFFFFu;
        key = k < key ? k : key;
      }
      const uint32_t min_cnt = key >> 3;
      if (key != 0xFFFFFFFFu && min_cnt >= 1u && isfinite(__longlong_as_double(q_first))) {
        back.q = __longlong_as_double(q_first);
        back.cnt0 = min_cnt;
        return key & 7u;
      }
    }
  }
  double q[NACT];
  uint32_t cnt[NACT];
#pragma unroll
  for (int a = 0; a < NACT; ++a) {
    const bool ex = (mask >> a) & 1u;
    q[a] = ex ? r.Q[a] : 0.0;
    cnt[a] = ex ? r.n[a] : 1u;
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
  const double range = __dsub_rn(hi, lo);

  // FP32 screen. The FP64 UCB values cost two divisions and a square root per action; the arg-max
  // almost never needs that precision. Scores are first evaluated in FP32 with one correctly
  // rounded reciprocal and square root per statistic and one MUFU.RSQ (<= 2 ulp) per action:
  //   norm  = float(q - lo) * rcp(float(range))                     |error| <= 4 u          (u = 2^-24)
  //   bonus = (float(2c) * sqrt(2 * logf(N))) * rsqrt(float(n))     |error| <= 11 u * bonus
  //           (__logf: absolute error <= 2^-21.41 on [0.5, 2] and <= 2 ulp elsewhere (CUDA C Programming Guide);
  //            the screen only runs for N >= 3, where ln N >= 1.09 makes that <= 3 ulp = 6 u relative,
  //            halved by the square root; N = 2 goes to the exact evaluation; the host's exact
  //            log table is only read by the exact evaluation - a table lookup here would be one
  //            more dependent memory access on the critical path of every level)
  //   score = fma(...)                                              |error| <= 12 u * (score + 1)
  // so two scores compare wrongly only if they are closer than 1.45e-6 * (score + 1). When the first
  // maximum beats every action with different inputs by more than `delta` = 6e-6 * (score + 1)
  // (4x that bound), the exact arg-max is provably the same and is returned. Actions with
  // bit-identical (Q, n) have identical exact scores, so the strict-> first-max rule picks the same
  // one either way. Anything else (including non-finite intermediates) falls through to the exact
  // FP64 evaluation, so the selected action is ALWAYS the reference's.
  // The exact evaluation then only has to look at the actions within `delta` of the FP32 maximum:
  // every other action's exact score is provably below the exact score of the FP32 winner, so it
  // cannot be the (first) exact maximum.
  int bi = -1;
  bool ok;
  bool narrowed = false;   // `candidates` is the FP32 screen's near-maximal set (its preconditions held)
  uint32_t candidates = mask;
  {
    const float rangef = (float)range;
    const float inv_range = __frcp_rn(rangef);
    const float csl = (float)p.two_c * __fsqrt_rn(2.0f * __logf((float)N));
    float sc[NACT];
    float best_sc = -1.0f;
    ok = isfinite(rangef) && rangef != 0.0f && N >= 3u;
#pragma unroll
    for (int a = 0; a < NACT; ++a) {
      sc[a] = -2.0f;
      if ((mask >> a) & 1u) {
        const float normf = (float)__dsub_rn(q[a], lo) * inv_range;
        sc[a] = fmaf(csl, rsqrtf((float)cnt[a]), normf);
        ok &= (cnt[a] != 0u) & isfinite(sc[a]);
        if (sc[a] > best_sc) { best_sc = sc[a]; bi = a; }
      }
    }
    // best_sc must be safely above DBL_MIN's role as the initial maximum (:225)
    ok &= (bi >= 0) & (best_sc > 1e-3f);
    double qb = 0.0;
    uint32_t nb = 0;
#pragma unroll
    for (int a = 0; a < NACT; ++a) if (a == bi) { qb = q[a]; nb = cnt[a]; }
    if (ok) {
      const float delta = 6e-6f * (best_sc + 1.0f);
      narrowed = true;
      candidates = 0u;
#pragma unroll
      for (int a = 0; a < NACT; ++a) {
        if (((mask >> a) & 1u) && sc[a] >= best_sc - delta) {
          candidates |= 1u << a;
          if (a != bi) ok &= (q[a] == qb) & (cnt[a] == nb);
        }
      }
    }
    back.q = qb;
    back.cnt0 = nb;
  }
  if (ok) return (uint32_t)bi;

  const double log2N = __dmul_rn(2.0, __ldg(p.log_table + N));  // 2 * log(total_node_visits_), host libm
  // the near-maximal candidates in FP64 with approximate reciprocals, then (rarely) the reference's exact evaluation
  const bool try_fast = narrowed && range > 1e-300 && range < 1e300 && p.two_c < 1e6;
  uint32_t best = exact_ucb_argmax<NACT>(r, candidates, lo, range, log2N, p.two_c, try_fast);
  exact_evals += best >> 31;
  best &= 0x7FFFFFFFu;
#pragma unroll
  for (int a = 0; a < NACT; ++a) if ((uint32_t)a == best) { back.q = r.Q[a]; back.cnt0 = r.n[a]; }
  return best;
}

// select_action on a classic StatRec: records a newly expanded action in place.
template <int NACT>
__device__ __forceinline__ uint32_t choose_action(StatRec<NACT>* s, const StatRegs<NACT>& r,
                                                  const SearchParams& p, const SearchTables& tab, int cls, uint32_t nact,
                                                  unsigned long long& exact_evals, BackRegs& back,
                                                  bool values_known_equal = false) {
  bool widened;
  const uint32_t a = select_action<NACT>(r, p, tab, cls, nact, exact_evals, back, widened, values_known_equal);
  if (widened) {
    s->nexp = r.nexp + 1;
    s->Q[a] = 0.0;  // ucb_statistics_[a] = UcbPair()
    s->n[a] = 0;
  }
  return a;
}

// Loads the backup inputs of one statistic (only used below the recorded path, see deep_backup).
template <int NACT>
__device__ __forceinline__ BackRegs back_load(const StatRec<NACT>* s, uint32_t act) {
  BackRegs b;
  b.cnt0 = s->n[act];
  b.q = s->Q[act];
  b.nv0 = s->N;
  b.v = s->V;
  return b;
}

// UctStatistic::update_statistics_from_backpropagated, reference uct_statistic.h:134-144, as pure
// arithmetic: G = r + gamma G_child, n_a += 1, Q_a += (G - Q_a) / n_a, N += 1, V += (G - V) / N.
struct BackOut {
  double G, q, v;
  uint32_t cnt, nv;
};

__device__ __forceinline__ BackOut back_compute(const BackRegs& b, double r, double gamma, double child_G) {
  BackOut o;
  o.G = __dadd_rn(r, __dmul_rn(gamma, child_G));
  o.cnt = b.cnt0 + 1;
  o.nv = b.nv0 + 1;
  // x + (d / k) with d == +-0 is x + d for any k > 0 (the quotient keeps d's signed zero), so the
  // division is skipped when the running mean already equals the return - always the case for the
  // reward-free other agents of CrossingState.
  const double dq = __dsub_rn(o.G, b.q);
  const double dv = __dsub_rn(o.G, b.v);
  o.q = (dq == 0.0) ? __dadd_rn(b.q, dq) : __dadd_rn(b.q, __ddiv_rn(dq, (double)o.cnt));
  o.v = (dv == 0.0) ? __dadd_rn(b.v, dv) : __dadd_rn(b.v, __ddiv_rn(dv, (double)o.nv));
  return o;
}

template <int NACT>
__device__ __forceinline__ double back_apply(StatRec<NACT>* s, uint32_t act, const BackRegs& b, double r,
                                             double gamma, double child_G) {
  const BackOut o = back_compute(b, r, gamma, child_G);
  s->n[act] = o.cnt;
  // an unchanged value is not rewritten (bit comparison: +0.0 and -0.0 differ), so its sector stays
  // clean - always the case for CrossingState's reward-free other agents
  if (__double_as_longlong(o.q) != __double_as_longlong(b.q)) s->Q[act] = o.q;
  s->N = o.nv;
  if (__double_as_longlong(o.v) != __double_as_longlong(b.v)) s->V = o.v;
  s->G = o.G;
  return o.G;
}

// back_apply for an agent whose values are known to stay +0.0: only the counts move (V, G and Q hold
// +0.0 since the node was created or the action expanded, exactly what the recurrence would rewrite).
template <int NACT>
__device__ __forceinline__ void back_apply_counts(StatRec<NACT>* s, uint32_t act, const BackRegs& b) {
  s->n[act] = b.cnt0 + 1;
  s->N = b.nv0 + 1;
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

// RandomHeuristic::rollout for one lane (reference random_heuristic.h:27-104); same arithmetic and
// the same Philox stream positions as rollout_kernel. Returns the ego return; `other` = the (shared)
// other-agent return. All lanes of a warp start their rollout together, so the position inside a
// Philox block is uniform among the active lanes.
template <class Env, int KIND>
__device__ __forceinline__ double lane_rollout(const SearchParams& p, typename Env::State st,
                                               uint32_t depth, uint32_t s_hi, uint32_t s_lo,
                                               const uint8_t* const_act, double& other,
                                               uint32_t& steps) {
  constexpr int A = Env::A;
  constexpr int S = (KIND == MAMCTS_SRC_PHILOX && A <= 8) ? 8 / A : 1;
  constexpr int NB = (A + 7) / 8;
  // accumulation without per-step FP64 work: see rollout_kernel (rollout.cu)
  double ego = 0.0, r_other_last = 0.0;
  uint32_t it = 0;
  // it >= rh_cap or depth >= max_depth (random_heuristic.h:45-47) as one limit on `it`
  const uint32_t depth_room = p.max_depth > depth ? p.max_depth - depth : 0u;
  const uint32_t limit = p.rh_cap < depth_room ? p.rh_cap : depth_room;
  bool fin = Env::terminal(p.env, st) | (it >= limit);
  if constexpr (Env::kOutcomeRewards) {
    // Environments whose step reward is a function of the step's outcome code, nonzero only on the
    // step that ends the episode (CrossingState with a zero per-step reward, the default): the loop
    // carries no reward at all, the single nonzero term is looked up afterwards. Same value: the
    // reference's sum is 0.0 + ... + 0.0 + d * r_last, and 0.0 + x = x.
    if (__double_as_longlong(Env::outcome_reward(p.env, 0u)) == 0ll) {
      uint32_t f = 0;
#pragma unroll 1
      while (!fin) {
        Philox4 blk[NB];
        if (KIND == MAMCTS_SRC_PHILOX) {
#pragma unroll
          for (int b = 0; b < NB; ++b)
            blk[b] = philox4x32_10((A <= 8 ? it / S : it * NB + b), s_lo, s_hi, 0u, p.key0, p.key1);
        }
#pragma unroll
        for (int sub = 0; sub < S; ++sub) {
          int32_t act[A];
          if (KIND == MAMCTS_SRC_PHILOX) {
            uint32_t h[A], nact[A];
#pragma unroll
            for (int a = 0; a < A; ++a) {
              h[a] = (A <= 8) ? philox_half(blk[0], sub * A + a) : philox_half(blk[a / 8], a % 8);
              nact[a] = Env::num_actions(p.env, a);
            }
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
      steps = it;
      const double r_last = it ? Env::outcome_reward(p.env, f) : 0.0;
      if (r_last != 0.0) ego = __dadd_rn(ego, __dmul_rn(__ldg(p.disc + it - 1), r_last));
      other = it ? __dmul_rn(__ldg(p.disc + it - 1), 0.0) : 0.0;   // rewards[1..] are 0
      return __dmul_rn(ego, __ldg(p.disc + p.disc_len + it));
    }
  }
#pragma unroll 1
  while (!fin) {
    // one Philox block feeds S steps; the S steps are unrolled so that every step reads its 16-bit
    // lanes from fixed registers (a dynamic pick cost 8 % of the kernel's instructions)
    Philox4 blk[NB];
    if (KIND == MAMCTS_SRC_PHILOX) {
#pragma unroll
      for (int b = 0; b < NB; ++b)
        blk[b] = philox4x32_10((A <= 8 ? it / S : it * NB + b), s_lo, s_hi, 0u, p.key0, p.key1);
    }
#pragma unroll
    for (int sub = 0; sub < S; ++sub) {
      int32_t act[A];
      if (KIND == MAMCTS_SRC_PHILOX) {
        uint32_t h[A], nact[A];
#pragma unroll
        for (int a = 0; a < A; ++a) {
          h[a] = (A <= 8) ? philox_half(blk[0], sub * A + a) : philox_half(blk[a / 8], a % 8);
          nact[a] = Env::num_actions(p.env, a);
        }
        philox_bounded_all<A>(h, nact, it * A, s_lo, s_hi, p.key0, p.key1);
#pragma unroll
        for (int a = 0; a < A; ++a) act[a] = (int32_t)h[a];
      } else {
#pragma unroll
        for (int a = 0; a < A; ++a) act[a] = (int32_t)const_act[a];
      }
      double r_ego, r_other, c;
      Env::step(p.env, st, act, r_ego, r_other, c);
      if (r_ego != 0.0) ego = __dadd_rn(ego, __dmul_rn(__ldg(p.disc + it), r_ego));
      r_other_last = r_other;
      it += 1;
      fin = Env::terminal(p.env, st) | (it >= limit);
      if (fin) break;
    }
  }
  steps = it;
  other = it ? __dmul_rn(__ldg(p.disc + it - 1), r_other_last) : 0.0;
  return __dmul_rn(ego, __ldg(p.disc + p.disc_len + it));
}

// lane_rollout for a GROUP of 8 lanes that work on one tree (search_compact_kernel with HELP: few trees per GPU,
// where a decision's time is the serial chain of one tree's iteration and most lanes would be idle). Philox is
// counter-based, so the blocks of the next 32 steps are computed side by side, one per lane; the step loop then
// runs on every lane of the group alike (same state, same control flow) and takes the 16-bit words of step s from
// lane s / 4 with one shuffle. Same stream positions, same arithmetic, same result as lane_rollout - for
// environments with outcome rewards and a zero per-step reward (the caller checks), two agents, PHILOX.
template <class Env>
__device__ __forceinline__ double group_rollout(const SearchParams& p, typename Env::State st, uint32_t depth,
                                                uint32_t s_hi, uint32_t s_lo, uint32_t gl, unsigned gmask,
                                                double& other, uint32_t& steps) {
  constexpr int A = Env::A;
  static_assert(A == 2 && Env::kOutcomeRewards, "group_rollout: two agents, outcome rewards");
  uint32_t it = 0, f = 0;
  const uint32_t depth_room = p.max_depth > depth ? p.max_depth - depth : 0u;
  const uint32_t limit = p.rh_cap < depth_room ? p.rh_cap : depth_room;
  bool fin = Env::terminal(p.env, st) | (it >= limit);
#pragma unroll 1
  while (!fin) {
    const Philox4 blk = philox4x32_10(it / 4u + gl, s_lo, s_hi, 0u, p.key0, p.key1);   // steps [it + 4 gl, it + 4 gl + 4)
#pragma unroll 1
    for (int src = 0; src < 8 && !fin; ++src) {
#pragma unroll
      for (int sub = 0; sub < 4; ++sub) {
        const uint32_t w = __shfl_sync(gmask, blk.w[sub], src, 8);   // halves 2 sub, 2 sub + 1: the two agents of the step
        uint32_t h[A] = {w & 0xFFFFu, w >> 16};
        const uint32_t nact[A] = {Env::num_actions(p.env, 0), Env::num_actions(p.env, 1)};
        philox_bounded_all<A>(h, nact, it * A, s_lo, s_hi, p.key0, p.key1);
        const int32_t act[A] = {(int32_t)h[0], (int32_t)h[1]};
        f = Env::step_outcome(p.env, st, act);
        it += 1;
        fin = (f != 0u) | (it >= limit);
        if (fin) break;
      }
    }
  }
  steps = it;
  double ego = 0.0;
  const double r_last = it ? Env::outcome_reward(p.env, f) : 0.0;
  if (r_last != 0.0) ego = __dadd_rn(ego, __dmul_rn(__ldg(p.disc + it - 1), r_last));
  other = it ? __dmul_rn(__ldg(p.disc + it - 1), 0.0) : 0.0;   // rewards[1..] are 0
  return __dmul_rn(ego, __ldg(p.disc + p.disc_len + it));
}

// The heuristic value of a new leaf (RandomHeuristic::calculate_heuristic_values, random_heuristic.h:106-134).
// MULTI = false is the reference: one rollout. MULTI = true (K = rollouts_per_leaf > 1, a power of two <= 16,
// PHILOX): the mean of K rollouts from the same state with stream_lo = creation index * K + k, summed in the
// fixed pairwise order of DESIGN.md section 6 (what reduce_mean_kernel and the oracle's orc_reduce_mean
// compute for K < 32), divided by K. A template parameter of the kernels, so that the K = 1 kernel is exactly
// what it is without this (a run-time branch around the rollout cost it 2-6 %).
template <class Env, int KIND, bool MULTI>
__device__ __forceinline__ double leaf_heuristic(const SearchParams& p, const typename Env::State& st,
                                                 uint32_t depth, uint32_t s_hi, uint32_t creation,
                                                 const uint8_t* const_act, double& other, uint32_t& steps) {
  if constexpr (!MULTI) {
    return lane_rollout<Env, KIND>(p, st, depth, s_hi, creation, const_act, other, steps);
  } else {
    const uint32_t K = p.rollouts_per_leaf;
    double ev[16], ov[16];
    steps = 0;
#pragma unroll 1
    for (uint32_t k = 0; k < K; ++k) {
      uint32_t n = 0;
      double o = 0.0;
      ev[k] = lane_rollout<Env, KIND>(p, st, depth, s_hi, creation * K + k, const_act, o, n);
      ov[k] = o;
      steps += n;
    }
#pragma unroll 1
    for (uint32_t off = K / 2; off >= 1; off /= 2) {   // v[l] + v[l ^ off], the entries that reach index 0
#pragma unroll 1
      for (uint32_t l = 0; l < off; ++l) { ev[l] = __dadd_rn(ev[l], ev[l + off]); ov[l] = __dadd_rn(ov[l], ov[l + off]); }
    }
    other = __ddiv_rn(ov[0], (double)K);
    return __ddiv_rn(ev[0], (double)K);
  }
}

// Backup below the recorded path (leaves deeper than kPathMax): follows parent pointers and loads the
// statistics. Cold for every configuration of the reference; kept out of line.
template <class Node, int NE, int NA, int A>
__device__ __noinline__ uint32_t deep_backup(Node* nodes, uint32_t child, uint32_t level, double gamma, double* G) {
  while (level > (uint32_t)kPathMax) {
    const Node* cn = nodes + child;
    const uint32_t par = cn->parent;
    Node* pn = nodes + par;
    const double r0 = cn->r_in[0], r1 = cn->r_in[1];
    const BackRegs be = back_load<NE>(&pn->ego, cn->ja[0]);
    G[0] = back_apply<NE>(&pn->ego, cn->ja[0], be, r0, gamma, G[0]);
#pragma unroll 1
    for (int j = 0; j < A - 1; ++j) {
      const BackRegs bo = back_load<NA>(&pn->oth[j], cn->ja[j + 1]);
      G[j + 1] = back_apply<NA>(&pn->oth[j], cn->ja[j + 1], bo, r1, gamma, G[j + 1]);
    }
    child = par;
    level -= 1;
  }
  return child;
}

template <class Env, int NE, int NA, int NCHILD, class CIDX, int KIND, bool MULTI = false>
__global__ void __launch_bounds__(kSearchThreads) search_kernel(const __grid_constant__ SearchParams p) {
  using Node = NodeRec<Env, NE, NA, NCHILD, CIDX>;
  constexpr int A = Env::A;
  constexpr bool kDirect = NCHILD > 0;
  __shared__ SearchTables tab;
  stage_tables(tab, p.tab);   // before any lane leaves: contains a block barrier
  const uint32_t lane = threadIdx.x & 31u;
  if (lane >= p.lanes_per_warp) return;
  const uint32_t tree = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * p.lanes_per_warp + lane;
  if (tree >= p.T) return;
  Node* nodes = static_cast<Node*>(p.nodes) + (size_t)tree * p.max_nodes;
  uint32_t* hash = kDirect ? nullptr : p.hash + (size_t)tree * (p.hash_mask + 1);
  uint32_t num_nodes = p.tree_nodes[tree];
  unsigned long long rollout_steps = 0, expand_steps = 0, backup_levels = 0, exact_evals = 0;
  const uint32_t nodes_before = num_nodes;
  uint8_t const_act[A];
#pragma unroll
  for (int a = 0; a < A; ++a) const_act[a] = tab.order[a == 0 ? 0 : 1][0];  // F1: first draw
  // Reward-free other agents (CrossingState: rewards[1..] stay 0, crossing_state.h:145) with a discount
  // >= 0: every return, V, G and Q of theirs is +0.0 for the whole search (m >= 0, m * 0.0 = +0.0), so
  // their values are neither loaded nor recomputed nor rewritten, and their selection is the integer
  // first-minimum of the counts (select_action).
  const bool other_values_zero = Env::kOtherRewardZero && p.gamma >= 0.0;

  // the path of the current iteration (collect(), node_statistic.h:115-121): per edge the parent
  // node, the joint action taken, the rewards collected on it and, per agent, the four statistic
  // values the backup of that level reads (already in registers during the descent)
  uint32_t path_node[kPathMax];
  uint8_t path_ja[kPathMax][A];
  double path_r0[kPathMax], path_r1[kPathMax];
  BackRegs path_back[kPathMax][A];

#pragma unroll 1
  for (uint32_t iter = 0; iter < p.iterations; ++iter) {
    uint32_t node = 0;
    uint32_t depth = 0;
    bool do_rollout = false;
    typename Env::State st;
    // ---- select & expand: mcts.h:300-305, stage_node.h:167-232 --------------------------------
#pragma unroll 1
    while (true) {
      Node* nd = nodes + node;
      // one round trip: everything the level needs is requested before anything is used
      const uint32_t vc = nd->visit_count;
      const uint32_t flags = nd->flags;
      const StatRegs<NE> ego = load_stat<NE>(&nd->ego);
      StatRegs<NA> oth0;
      if (A > 1) oth0 = other_values_zero ? load_stat_counts<NA>(&nd->oth[0]) : load_stat<NA>(&nd->oth[0]);
      if (A > 2) {
#pragma unroll 1
        for (int j = 1; j < A - 1; ++j) { prefetch_l1(&nd->oth[j]); prefetch_l1(&nd->oth[j].Q[NA - 1]); }
      }
      if (kDirect) prefetch_l1(&nd->child[0]);
      double r_in0 = 0.0, r_in1 = 0.0;
      if (depth > 0) { r_in0 = nd->r_in[0]; r_in1 = nd->r_in[1]; }
      nd->visit_count = vc + 1;                                          // :180
      if (depth > 0 && depth <= kPathMax) { path_r0[depth - 1] = r_in0; path_r1[depth - 1] = r_in1; }
      st.flags = flags;
      if (Env::kStateInts == 1) st.x[0] = nd->x[0];  // SimpleState: terminal() reads the state
      if (Env::terminal(p.env, st) || depth == p.max_depth || num_nodes > p.node_budget) break;  // :183-187
      uint8_t ja[A];
      BackRegs back[A];
      ja[0] = (uint8_t)choose_action<NE>(&nd->ego, ego, p, tab, 0, p.n_ego, exact_evals, back[0]);  // :190-195
      if (A > 1) ja[1] = (uint8_t)choose_action<NA>(&nd->oth[0], oth0, p, tab, 1, p.n_oth, exact_evals, back[1], other_values_zero);
      if (A > 2) {
        // further other agents: one copy of the selection code, statistics come from L1 (prefetched)
#pragma unroll 1
        for (int j = 1; j < A - 1; ++j) {
          const StatRegs<NA> o = other_values_zero ? load_stat_counts<NA>(&nd->oth[j]) : load_stat<NA>(&nd->oth[j]);
          BackRegs b;
          ja[j + 1] = (uint8_t)choose_action<NA>(&nd->oth[j], o, p, tab, 1, p.n_oth, exact_evals, b, other_values_zero);
          back[j + 1] = b;
        }
      }
      if (depth < kPathMax) {
        path_node[depth] = node;
#pragma unroll
        for (int a = 0; a < A; ++a) { path_ja[depth][a] = ja[a]; path_back[depth][a] = back[a]; }
      }
      // child lookup, :198
      uint32_t child = 0;
      uint32_t slot = 0;
      if (kDirect) {
        slot = (uint32_t)ja[0] * NA + (A > 1 ? (uint32_t)ja[A > 1 ? 1 : 0] : 0u);
        child = nd->child[slot];
      } else {
        slot = hash_ja<A>(node, ja) & p.hash_mask;
        while (true) {
          const uint32_t c = hash[slot];
          if (c == 0) break;
          const Node* cn = nodes + c;
          // the candidate is almost always the child: start pulling its record now
          prefetch_l1(&cn->ego);
          prefetch_l1(&cn->oth[0]);
          bool same = cn->parent == node;
#pragma unroll
          for (int a = 0; a < A; ++a) same &= cn->ja[a] == ja[a];
          if (same) { child = c; break; }
          slot = (slot + 1) & p.hash_mask;
        }
      }
      if (child) { node = child; depth += 1; continue; }                 // :199-206
      // expand, :207-231
#pragma unroll
      for (int i = 0; i < A; ++i) { st.x[i] = (i < Env::kStateInts) ? nd->x[i] : 0; st.last[i] = 0; }
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
      st.flags = (st.flags & ~1u) | (uint32_t)Env::terminal(p.env, st);
      cn->flags = (uint8_t)st.flags;
#pragma unroll
      for (int a = 0; a < A; ++a) cn->ja[a] = ja[a];
#pragma unroll
      for (int i = 0; i < Env::kStateInts; ++i) cn->x[i] = st.x[i];
      cn->r_in[0] = r_ego;
      cn->r_in[1] = r_other;
      cn->ego.nexp = 0;
#pragma unroll
      for (int j = 0; j < A - 1; ++j) cn->oth[j].nexp = 0;
      if (kDirect) {
#pragma unroll
        for (int k = 0; k < (NCHILD > 0 ? NCHILD : 1); ++k) cn->child[k] = 0;
        nd->child[slot] = (CIDX)child;
      } else {
        hash[slot] = child;
      }
      if (depth < kPathMax) { path_r0[depth] = r_ego; path_r1[depth] = r_other; }
      node = child;
      depth += 1;
      do_rollout = true;
      break;
    }
    // ---- heuristic: mcts.h:309-312, random_heuristic.h:106-134, uct_statistic.h:100-110 --------
    Node* leaf = nodes + node;
    double G[A];
    if (do_rollout) {
      double other = 0.0;
      uint32_t steps = 0;
      const double ego = leaf_heuristic<Env, KIND, MULTI>(p, st, depth, p.first_tree_id + tree, node, const_act,
                                                   other, steps);
      rollout_steps += steps;
      leaf->ego.V = ego; leaf->ego.G = ego; leaf->ego.N = 1;   // fresh node: 0 + 1
      G[0] = ego;
#pragma unroll
      for (int j = 0; j < A - 1; ++j) {
        leaf->oth[j].V = other; leaf->oth[j].G = other; leaf->oth[j].N = 1;
        G[j + 1] = other;
      }
    } else {
      G[0] = leaf->ego.G;
#pragma unroll
      for (int j = 0; j < A - 1; ++j) G[j + 1] = leaf->oth[j].G;
    }
    // ---- backpropagation: mcts.h:314-329, stage_node.h:259-271 ---------------------------------
    uint32_t level = depth;  // number of edges between the root and the leaf
    backup_levels += depth;
    if (level > (uint32_t)kPathMax) {
      deep_backup<Node, NE, NA, A>(nodes, node, level, p.gamma, G);
      level = kPathMax;
    }
    // recorded path: no loads, only the arithmetic and the stores
#pragma unroll 1
    while (level > 0) {
      level -= 1;
      Node* pn = nodes + path_node[level];
      const double r0 = path_r0[level], r1 = path_r1[level];
      G[0] = back_apply<NE>(&pn->ego, path_ja[level][0], path_back[level][0], r0, p.gamma, G[0]);
      if (other_values_zero) {
#pragma unroll
        for (int j = 0; j < A - 1; ++j) back_apply_counts<NA>(&pn->oth[j], path_ja[level][j + 1], path_back[level][j + 1]);
      } else {
#pragma unroll
        for (int j = 0; j < A - 1; ++j)
          G[j + 1] = back_apply<NA>(&pn->oth[j], path_ja[level][j + 1], path_back[level][j + 1], r1, p.gamma, G[j + 1]);
      }
    }
  }
  p.tree_nodes[tree] = num_nodes;
  p.tree_iters[tree] += p.iterations;
  atomicAdd(p.counters + 0, (unsigned long long)p.iterations);
  atomicAdd(p.counters + 1, rollout_steps);
  atomicAdd(p.counters + 2, expand_steps);
  atomicAdd(p.counters + 3, (unsigned long long)(num_nodes - nodes_before));
  atomicAdd(p.counters + 4, backup_levels);
  atomicAdd(p.counters + 5, exact_evals);
}

template <class Env, int NE, int NA, int NCHILD, class CIDX>
__global__ void search_reset_kernel(void* nodes_v, uint32_t* tree_nodes, uint32_t* tree_iters, uint32_t T,
                                    uint32_t max_nodes, const int32_t* root_x, EnvParams env) {
  using Node = NodeRec<Env, NE, NA, NCHILD, CIDX>;
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
  for (int a = 0; a < NE; ++a) { root->ego.Q[a] = 0.0; root->ego.n[a] = 0; }  // the root merge reads n[a] > 0 as "expanded"
  for (int j = 0; j < Env::A - 1; ++j) { root->oth[j].V = 0.0; root->oth[j].G = 0.0; root->oth[j].N = 0; root->oth[j].nexp = 0; }
  for (int k = 0; k < (NCHILD > 0 ? NCHILD : 1); ++k) root->child[k] = 0;
  tree_nodes[tree] = 1;
  tree_iters[tree] = 0;
}

}  // namespace mamcts
