// rollout.cu - the batched random-heuristic rollout (reference
// mcts/heuristics/random_heuristic.h:27-104) as a persistent CUDA kernel for sm_100a.
//
// Mapping: one lane = one episode, state in registers; a warp keeps 32 episodes in flight and
// refills finished lanes from a global work list (one atomicAdd per refill of a whole warp) in a
// warp-uniform "service" step, so the environment step always runs converged and no lane idles while
// rollouts are left - episode lengths vary by 5x, a static assignment left 40 % of the lanes idle. PHILOX draws are laid out so that one Philox4x32-10 block
// feeds S = 8/A consecutive steps and all lanes of a warp regenerate their block at the same loop
// iteration. K > 1 rollouts per leaf go through a per-rollout scratch and a fixed-
// order warp-shuffle segmented reduction (reduce.cu), so the means are deterministic.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "engine.h"
#include "env.cuh"
#include "philox.cuh"
#include "staging.h"

namespace mamcts {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int kRolloutThreads = 256;
constexpr int kServiceThreshold = 8;  // finished lanes that trigger a refill while others still run

struct RolloutParams {
  uint32_t n_leaves, K, total;
  // inputs (device); x / last are Env::Domain-typed (int32 or float)
  const void* x;
  const void* last;
  const uint8_t* flags;
  const uint32_t* depth;
  // action source
  const int8_t* replay;
  const float* replay_f;   // CrossingState<float>
  const uint32_t* replay_steps;
  uint32_t replay_stride;
  uint32_t const_actions[MAMCTS_MAX_AGENTS];
  uint32_t key0, key1;
  const uint32_t* stream_hi;
  const uint32_t* stream_lo;
  uint32_t stream_hi_base, stream_lo_base;
  const int32_t* hyp_lo;   // HYPOTHESIS: desired-gap range of other j of leaf i at [j * n_leaves + i]
  const int32_t* hyp_hi;
  const float* hyp_lo_f;   // the same for CrossingState<float>
  const float* hyp_hi_f;
  // mcts
  const double* disc;      // [0, disc_len): reward weight of step k; [disc_len, 2 disc_len): (1/nA_ego)^k
  uint32_t disc_len;
  uint32_t rh_cap, max_depth;
  EnvParams env;
  // per-rollout outputs (device); index r = leaf*K + k; other_ret / final_* are SoA with stride total
  double* ego_ret;
  double* other_ret;
  double* cost;
  double* step_len;
  uint32_t* steps;
  void* final_x;           // Env::Domain-typed
  void* final_last;
  uint8_t* final_flags;
  double* trace;
  uint32_t trace_stride;
  unsigned long long* total_steps;
  uint32_t service_threshold;  // finished lanes that trigger a refill while others still run
  unsigned int* work_cursor;   // rollouts handed out beyond the statically assigned first grid*threads
};

template <class Env, int KIND>
struct StepsPerBlock {
  static constexpr bool kPhilox = KIND == MAMCTS_SRC_PHILOX || KIND == MAMCTS_SRC_HYPOTHESIS;
  static constexpr int value = (kPhilox && Env::A <= 8) ? 8 / Env::A : 1;
};

template <class Env, int KIND, bool PARITY>
__global__ void __launch_bounds__(kRolloutThreads)
rollout_kernel(const __grid_constant__ RolloutParams p) {
  constexpr int A = Env::A;
  constexpr int S = StepsPerBlock<Env, KIND>::value;
  constexpr bool kSimple = std::is_same<Env, SimpleEnv>::value;
  using D = typename Env::Domain;   // int32_t, or float for CrossingState<float>
  constexpr bool kFloat = std::is_same<D, float>::value;
  const D* const in_x = static_cast<const D*>(p.x);
  const D* const in_last = static_cast<const D*>(p.last);
  D* const out_x = static_cast<D*>(p.final_x);
  D* const out_last = static_cast<D*>(p.final_last);
  constexpr bool kPhilox = StepsPerBlock<Env, KIND>::kPhilox;   // draws come from the Philox stream
  constexpr bool kHyp = KIND == MAMCTS_SRC_HYPOTHESIS;
  constexpr int XR = kSimple ? 1 : A;  // rows of the x arrays (SimpleState: only state_length_)
  uint32_t rid = blockIdx.x * blockDim.x + threadIdx.x;   // first rollout: static; later ones from the cursor

  typename Env::State st;
  // Return accumulation (random_heuristic.h:71-88) without per-step FP64 work: the weight of step k
  // and (1/nA_ego)^k come from the table (engine.h: ensure_discount_table), and adding a zero reward
  // or cost is the identity on an accumulator that started at +0.0 (it can never hold -0.0), so those
  // additions are skipped; other agents' returns are ASSIGNED (:75), so only the last step matters.
  double ego = 0.0, cost = 0.0, r_other_last = 0.0;
  uint32_t it = 0, depth = 0, cap = 0, leaf = 0, s_lo = 0, s_hi = 0;
  // Environments whose reward and cost are functions of the step's outcome code, which is nonzero exactly
  // on the step that ends the episode (CrossingState<int>): when the per-step reward and cost are zero
  // (the default), the loop carries neither; the one nonzero term of each sum is looked up from the last
  // outcome afterwards (0.0 + x = x, so the sums are the reference's bit for bit).
  uint32_t last_outcome = 0;
  bool lazy = false;
  if constexpr (Env::kOutcomeRewards) {
    lazy = __double_as_longlong(Env::outcome_reward(p.env, 0u)) == 0ll &&
           __double_as_longlong(Env::outcome_cost(p.env, 0u)) == 0ll && !(PARITY && p.trace);
  }
  bool has = false, fin = true;
  unsigned long long lane_steps = 0;
  int32_t gap_lo[kHyp ? A : 1], gap_span[kHyp ? A : 1];   // [a]: other agent at position a (a >= 1)
  float gap_lo_f[(kHyp && kFloat) ? A : 1], gap_span_f[(kHyp && kFloat) ? A : 1];   // CrossingState<float>: lo, hi - lo

  auto load = [&]() {
    has = rid < p.total;
    if (!has) return;
    leaf = (p.K == 1) ? rid : rid / p.K;
#pragma unroll
    for (int a = 0; a < A; ++a) {
      st.x[a] = (a < XR) ? __ldg(in_x + (size_t)a * p.n_leaves + leaf) : (D)0;
      st.last[a] = (in_last && !kSimple) ? __ldg(in_last + (size_t)a * p.n_leaves + leaf) : (D)2;
    }
    st.flags = p.flags ? (uint32_t)(__ldg(p.flags + leaf) & 1u) : 0u;
    depth = p.depth ? __ldg(p.depth + leaf) : 0u;
    cap = p.rh_cap;
    if (KIND == MAMCTS_SRC_REPLAY) cap = min(cap, __ldg(p.replay_steps + leaf));
    // it >= cap or depth >= max_depth (random_heuristic.h:45-47, :51-54) as one limit on `it`
    cap = min(cap, p.max_depth > depth ? p.max_depth - depth : 0u);
    if (kPhilox) {
      s_hi = p.stream_hi ? __ldg(p.stream_hi + leaf) : p.stream_hi_base;
      s_lo = p.stream_lo ? __ldg(p.stream_lo + leaf) + (rid - leaf * p.K) : p.stream_lo_base + rid;
    }
    if (kHyp && !kFloat) {
#pragma unroll
      for (int a = 1; a < A; ++a) {
        gap_lo[a] = __ldg(p.hyp_lo + (size_t)(a - 1) * p.n_leaves + leaf);
        gap_span[a] = max(__ldg(p.hyp_hi + (size_t)(a - 1) * p.n_leaves + leaf) - gap_lo[a] + 1, 1);  // lo <= hi (agent_policy.h:28)
      }
    }
    if constexpr (kHyp && kFloat) {
#pragma unroll
      for (int a = 1; a < A; ++a) {
        gap_lo_f[a] = __ldg(p.hyp_lo_f + (size_t)(a - 1) * p.n_leaves + leaf);
        gap_span_f[a] = __fsub_rn(__ldg(p.hyp_hi_f + (size_t)(a - 1) * p.n_leaves + leaf), gap_lo_f[a]);
      }
    }
    ego = 0.0; cost = 0.0; r_other_last = 0.0; it = 0;   // random_heuristic.h:36-49
    last_outcome = 0;
    fin = Env::terminal(p.env, st) | (it >= cap);        // :51-54
  };

  auto finalize = [&]() {
    if constexpr (Env::kOutcomeRewards) {
      if (lazy && it) {
        const double r_last = Env::outcome_reward(p.env, last_outcome);
        const double c_last = Env::outcome_cost(p.env, last_outcome);
        if (r_last != 0.0) ego = __dadd_rn(ego, __dmul_rn(__ldg(p.disc + it - 1), r_last));
        if (c_last != 0.0) cost = __dadd_rn(cost, c_last);
      }
    }
    // random_heuristic.h:95-101
    const double prob = __ldg(p.disc + p.disc_len + it);
    const double other = it ? __dmul_rn(__ldg(p.disc + it - 1), r_other_last) : 0.0;  // :73-77
    const double ego_out = __dmul_rn(ego, prob);
    const double cost_out = Env::kHasCost ? __dmul_rn(cost, prob) : 0.0;
    if (p.ego_ret) p.ego_ret[rid] = ego_out;
    if (p.other_ret) {
#pragma unroll
      for (int j = 0; j < A - 1; ++j) p.other_ret[(size_t)j * p.total + rid] = other;
    }
    if (p.cost) p.cost[rid] = cost_out;
    if (p.step_len) p.step_len[rid] = (double)it;
    lane_steps += it;
    if (PARITY) {
      if (p.steps) p.steps[rid] = it;
      if (p.final_x) {
#pragma unroll
        for (int a = 0; a < XR; ++a) out_x[(size_t)a * p.total + rid] = st.x[a];
      }
      if (p.final_last) {
#pragma unroll
        for (int a = 0; a < A; ++a) out_last[(size_t)a * p.total + rid] = st.last[a];
      }
      if (p.final_flags) p.final_flags[rid] = (uint8_t)((st.flags & ~1u) | (uint32_t)Env::terminal(p.env, st));
    }
  };

  load();

  while (true) {
    const unsigned need = __ballot_sync(FULL, has && fin);
    const unsigned running = __ballot_sync(FULL, has && !fin);
    if (need != 0 && (running == 0 || __popc(need) >= (int)p.service_threshold)) {
      // the finished lanes take the next popc(need) rollouts of the global list
      uint32_t base = 0;
      const uint32_t leader = __ffs(need) - 1;
      if ((threadIdx.x & 31u) == leader) base = atomicAdd(p.work_cursor, (unsigned int)__popc(need));
      base = __shfl_sync(FULL, base, leader);
      if (has && fin) {
        finalize();
        const uint32_t rank = __popc(need & ((1u << (threadIdx.x & 31u)) - 1u));
        // lanes past the end stop asking, so the cursor never exceeds total + launched lanes < 2^32
        rid = gridDim.x * blockDim.x + base + rank;
        load();
      }
      continue;
    }
    if (running == 0) break;

    const bool active0 = has && !fin;
    Philox4 blk[(A + 7) / 8];
    if (kPhilox && active0) {
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
      if (has && !fin) {
        D act[A];
        if (kPhilox) {
          // all draws of the step: 16-bit words of the block(s), then one branch-free bounded reduction
          uint32_t h[A], span[A];
#pragma unroll
          for (int a = 0; a < A; ++a) {
            h[a] = (A <= 8) ? philox_half(blk[0], s * A + a) : philox_half(blk[a / 8], a % 8);
            // HYPOTHESIS: the other agents draw their gap for this step
            // (AgentPolicyCrossingState<int>::act, agent_policy.h:68-75)
            span[a] = (kHyp && a > 0) ? (uint32_t)gap_span[kHyp ? a : 0] : Env::num_actions(p.env, a);
          }
          // CrossingState<float> hypotheses: the desired gap is lo + (hi - lo) * (w / 65536) of the raw 16-bit word
          // w (uniform_real_distribution's stand-in, agent_policy.h:78-84); only the ego's draw is a bounded index
          float gap_f[(kHyp && kFloat) ? A : 1];
          if constexpr (kHyp && kFloat) {
#pragma unroll
            for (int a = 1; a < A; ++a) {
              gap_f[a] = __fadd_rn(gap_lo_f[a], __fmul_rn(gap_span_f[a], __fmul_rn((float)h[a], 1.0f / 65536.0f)));
              span[a] = 65536u;   // no rejection: the word itself is the draw
            }
          }
          philox_bounded_all<A>(h, span, it * A, s_lo, s_hi, p.key0, p.key1);
#pragma unroll
          for (int a = 0; a < A; ++a) {
            if constexpr (kHyp && kFloat) act[a] = a > 0 ? (D)gap_f[a] : Env::action_from_index(a, h[a]);
            else act[a] = (kHyp && a > 0) ? (D)(gap_lo[kHyp ? a : 0] + (int32_t)h[a]) : Env::action_from_index(a, h[a]);
          }
        } else {
#pragma unroll
          for (int a = 0; a < A; ++a) {
            if (KIND == MAMCTS_SRC_REPLAY) {
              const size_t ri = ((size_t)leaf * p.replay_stride + it) * A + a;
              act[a] = kFloat ? (D)p.replay_f[kFloat ? ri : 0] : (D)p.replay[kFloat ? 0 : ri];
            } else {
              act[a] = Env::action_from_index(a, p.const_actions[a]);
            }
          }
        }
        // the hypothesis policy reads the PRE-step state of every agent, so it runs after all draws
        if constexpr (kHyp && !kSimple && !kFloat) {
#pragma unroll
          for (int a = 1; a < A; ++a)
            act[a] = crossing_policy_action(p.env, st.x[a], st.last[a], st.x[0], st.last[0], act[a]);
        }
        if constexpr (kHyp && kFloat) {
          D v[A];
#pragma unroll
          for (int a = 1; a < A; ++a)
            v[a] = crossing_policy_action_f32(p.env, st.x[a], st.last[a], st.x[0], st.last[0], act[a]);
#pragma unroll
          for (int a = 1; a < A; ++a) act[a] = v[a];
        }
        bool ended;
        if constexpr (Env::kOutcomeRewards) {
          const uint32_t f = Env::step_outcome(p.env, st, act);
          last_outcome = f;
          ended = f != 0u;
          if (!lazy) {
            const double r_ego = Env::outcome_reward(p.env, f), c = Env::outcome_cost(p.env, f);
            if (PARITY && p.trace && it < p.trace_stride) p.trace[(size_t)rid * p.trace_stride + it] = r_ego;
            if (r_ego != 0.0) ego = __dadd_rn(ego, __dmul_rn(__ldg(p.disc + it), r_ego));  // :71-72
            if (c != 0.0) cost = __dadd_rn(cost, c);                                       // :79-84
          }
          r_other_last = 0.0;                      // rewards[1..] stay 0 (crossing_state.h:145)
        } else {
          double r_ego, r_other, c;
          Env::step(p.env, st, act, r_ego, r_other, c);
          if (PARITY && p.trace && it < p.trace_stride) p.trace[(size_t)rid * p.trace_stride + it] = r_ego;
          if (r_ego != 0.0) ego = __dadd_rn(ego, __dmul_rn(__ldg(p.disc + it), r_ego));  // :71-72
          r_other_last = r_other;                                                        // :73-77
          if (Env::kHasCost && c != 0.0) cost = __dadd_rn(cost, c);                      // :79-84
          ended = Env::terminal(p.env, st);
        }
        it += 1;                                   // :89 (depth advances with it, :90: folded into cap)
        fin = ended | (it >= cap);
      }
    }
  }

  // env-steps counter: one atomic per warp
  if (p.total_steps) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) lane_steps += __shfl_xor_sync(FULL, lane_steps, off);
    if ((threadIdx.x & 31) == 0 && lane_steps) atomicAdd(p.total_steps, lane_steps);
  }
}

// K-rollout mean per leaf in the fixed order of DESIGN.md §6 (oracle: orc_reduce_mean).
//   K < 32 : K-lane segments, xor-butterfly K/2..1, mean = total / K
//   K >= 32: one warp per leaf; lane l sums vals[c*32R + j*32 + l], j < R = min(K/32, 8), then a
//            32-lane xor-butterfly; chunk partials are added in chunk order.
// blockIdx.y selects the array: per-rollout values at src + y*src_stride -> per-leaf means at dst.p[y]
// (null = not requested), so one launch reduces ego, every other agent, cost and step length.
struct ReduceDst {
  double* p[MAMCTS_MAX_AGENTS + 2];
};

__global__ void __launch_bounds__(256)
reduce_mean_kernel(const double* __restrict__ src, const __grid_constant__ ReduceDst dsts, uint32_t n, uint32_t K,
                   size_t src_stride) {
  double* __restrict__ dst = dsts.p[blockIdx.y];
  if (!dst) return;
  src += (size_t)blockIdx.y * src_stride;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t num_warps = (gridDim.x * blockDim.x) >> 5;
  if (K < 32) {
    const uint32_t per_warp = 32 / K;
    const uint32_t groups = (n + per_warp - 1) / per_warp;
    for (uint32_t g = warp; g < groups; g += num_warps) {
      const uint32_t leaf = g * per_warp + lane / K;
      double v = leaf < n ? src[(size_t)leaf * K + (lane % K)] : 0.0;
      for (uint32_t off = K / 2; off >= 1; off >>= 1) v = __dadd_rn(v, __shfl_xor_sync(FULL, v, off));
      if (leaf < n && (lane % K) == 0) dst[leaf] = v / (double)K;
    }
  } else {
    const uint32_t R = min(K / 32, 8u);
    const uint32_t chunk = 32 * R;
    const uint32_t nchunks = K / chunk;
    for (uint32_t leaf = warp; leaf < n; leaf += num_warps) {
      double total = 0.0;
      for (uint32_t c = 0; c < nchunks; ++c) {
        double acc = 0.0;
        for (uint32_t j = 0; j < R; ++j)
          acc = __dadd_rn(acc, src[(size_t)leaf * K + (size_t)c * chunk + j * 32 + lane]);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(FULL, acc, off));
        total = __dadd_rn(total, acc);
      }
      if (lane == 0) dst[leaf] = total / (double)K;
    }
  }
}

static bool valid_k(uint32_t K) {
  if (K == 1 || K == 2 || K == 4 || K == 8 || K == 16) return true;
  if (K >= 32 && K <= 256 && K % 32 == 0) return true;
  return K > 256 && K % 256 == 0;
}

// Persistent grid: as many blocks as stay resident (occupancy x SM count), never more than the work.
template <class Kernel>
static int launch_persistent(mamcts_engine* e, Kernel kernel, RolloutParams p) {
  int per_sm = 0;
  MAMCTS_CUDA(e, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kRolloutThreads, 0));
  const int max_blocks = e->num_sms * std::max(per_sm, 1);
  const int need_blocks = (int)(((uint64_t)p.total + kRolloutThreads - 1) / kRolloutThreads);
  const int grid = std::max(1, std::min(max_blocks, need_blocks));
  // the first grid*threads rollouts are assigned statically, the cursor hands out the rest
  if (!e->work_cursor) MAMCTS_CUDA(e, cudaMalloc(reinterpret_cast<void**>(&e->work_cursor), sizeof(unsigned int)));
  MAMCTS_CUDA(e, cudaMemsetAsync(e->work_cursor, 0, sizeof(unsigned int), e->stream));
  p.work_cursor = e->work_cursor;
  p.service_threshold = kServiceThreshold;
  kernel<<<dim3(grid), dim3(kRolloutThreads), 0, e->stream>>>(p);
  return MAMCTS_OK;
}

template <class Env>
static int launch_rollout(mamcts_engine* e, const RolloutParams& p, int kind, bool parity) {
#define MAMCTS_LAUNCH(KIND)                                                                   \
  do {                                                                                        \
    int rc__ = parity ? launch_persistent(e, rollout_kernel<Env, KIND, true>, p)              \
                      : launch_persistent(e, rollout_kernel<Env, KIND, false>, p);            \
    if (rc__ != MAMCTS_OK) return rc__;                                                       \
  } while (0)
  if (kind == MAMCTS_SRC_REPLAY) MAMCTS_LAUNCH(MAMCTS_SRC_REPLAY);
  else if (kind == MAMCTS_SRC_REFERENCE_CONST) MAMCTS_LAUNCH(MAMCTS_SRC_REFERENCE_CONST);
  else if (kind == MAMCTS_SRC_HYPOTHESIS) {
    if constexpr (std::is_same<Env, SimpleEnv>::value) return fail(e, MAMCTS_ERR_INVALID, "rollout: HYPOTHESIS is a CrossingState action source");
    else MAMCTS_LAUNCH(MAMCTS_SRC_HYPOTHESIS);
  }
  else MAMCTS_LAUNCH(MAMCTS_SRC_PHILOX);
#undef MAMCTS_LAUNCH
  e->launches += 1;
  MAMCTS_CUDA(e, cudaGetLastError());
  return MAMCTS_OK;
}

static int dispatch_crossing(mamcts_engine* e, const RolloutParams& p, int num_others, int kind,
                             bool parity) {
  switch (num_others) {
    case 1: return launch_rollout<CrossingEnv<1>>(e, p, kind, parity);
    case 2: return launch_rollout<CrossingEnv<2>>(e, p, kind, parity);
    case 3: return launch_rollout<CrossingEnv<3>>(e, p, kind, parity);
    case 4: return launch_rollout<CrossingEnv<4>>(e, p, kind, parity);
    case 5: return launch_rollout<CrossingEnv<5>>(e, p, kind, parity);
    case 6: return launch_rollout<CrossingEnv<6>>(e, p, kind, parity);
    case 7: return launch_rollout<CrossingEnv<7>>(e, p, kind, parity);
    case 11: return launch_rollout<CrossingEnv<11>>(e, p, kind, parity);
    case 15: return launch_rollout<CrossingEnv<15>>(e, p, kind, parity);
    default:
      return fail(e, MAMCTS_ERR_UNSUPPORTED,
                  "rollout: compiled for 1..7, 11 or 15 other agents (CrossingState)");
  }
}

// CrossingState<float>: the same kernel over CrossingEnvF (fewer agent counts are compiled: the
// reference's float tests use 2 other agents, crossing_state_float_test.cc)
constexpr int kEnvCrossingF32 = 100;   // internal environment kind of rollout_common

static int dispatch_crossing_f32(mamcts_engine* e, const RolloutParams& p, int num_others, int kind, bool parity) {
  switch (num_others) {
    case 1: return launch_rollout<CrossingEnvF<1>>(e, p, kind, parity);
    case 2: return launch_rollout<CrossingEnvF<2>>(e, p, kind, parity);
    case 3: return launch_rollout<CrossingEnvF<3>>(e, p, kind, parity);
    case 7: return launch_rollout<CrossingEnvF<7>>(e, p, kind, parity);
    default:
      return fail(e, MAMCTS_ERR_UNSUPPORTED, "rollout: CrossingState<float> is compiled for 1, 2, 3 or 7 other agents");
  }
}

// Shared body of the entry points. A = number of agents; x_rows = rows of the state array; the state
// arrays are int32 or (kEnvCrossingF32) float - both 4 bytes per element.
// (the caller holds e->mu)
static int rollout_locked(mamcts_engine* e, int env_kind, const mamcts_params* mp, const EnvParams& ep,
                          uint32_t A, uint32_t n, uint32_t K, int32_t mem, const void* state_x,
                          const void* state_last, const uint8_t* state_flags,
                          const uint32_t* start_depth, const mamcts_action_source* src,
                          mamcts_rollout_out* out) {
  if (!mp || !src || !out || !state_x) return fail(e, MAMCTS_ERR_INVALID, "rollout: null argument");
  if (n == 0) return MAMCTS_OK;
  if (!valid_k(K))
    return fail(e, MAMCTS_ERR_INVALID,
                "rollout: rollouts_per_leaf must be 1,2,4,8,16, a multiple of 32 up to 256, or a multiple of 256");
  if ((uint64_t)n * K > 0xF0000000ull) return fail(e, MAMCTS_ERR_INVALID, "rollout: n_leaves * K too large for one call");
  const bool f32 = env_kind == kEnvCrossingF32;
  if (src->kind == MAMCTS_SRC_REPLAY &&
      (K != 1 || !(f32 ? (const void*)src->replay_actions_f32 : (const void*)src->replay_actions) || !src->replay_steps))
    return fail(e, MAMCTS_ERR_INVALID, "rollout: REPLAY needs K == 1 and replay buffers");
  if (src->kind < MAMCTS_SRC_REPLAY || src->kind > MAMCTS_SRC_HYPOTHESIS)
    return fail(e, MAMCTS_ERR_INVALID, "rollout: unknown action source");
  if (src->kind == MAMCTS_SRC_HYPOTHESIS &&
      !((env_kind == MAMCTS_ENV_CROSSING_I32 && src->hyp_gap_lo && src->hyp_gap_hi) ||
        (f32 && src->hyp_gap_lo_f32 && src->hyp_gap_hi_f32)))
    return fail(e, MAMCTS_ERR_INVALID, "rollout: HYPOTHESIS needs CrossingState and the per-leaf gap ranges");
  if (src->kind == MAMCTS_SRC_REFERENCE_CONST) {
    for (uint32_t a = 0; a < A; ++a) {
      const uint32_t na = (env_kind == MAMCTS_ENV_SIMPLE) ? 2u : (a == 0 ? ep.n_ego_actions : ep.n_other_actions);
      if (src->const_actions[a] >= na) return fail(e, MAMCTS_ERR_INVALID, "rollout: const action out of range");
    }
  }
  MAMCTS_CUDA(e, cudaSetDevice(e->device));
  const uint32_t total = n * K;
  const uint32_t x_rows = (env_kind == MAMCTS_ENV_SIMPLE) ? 1 : A;

  ArgStager stg(e);
  const int t_x = stg.plan(state_x, sizeof(int32_t) * x_rows * n, mem, false);
  const int t_last = stg.plan(state_last, sizeof(int32_t) * x_rows * n, mem, false);
  const int t_flags = stg.plan(state_flags, n, mem, false);
  const int t_depth = stg.plan(start_depth, sizeof(uint32_t) * n, mem, false);
  const int t_rep = stg.plan(src->kind == MAMCTS_SRC_REPLAY && !f32 ? src->replay_actions : nullptr,
                             (size_t)n * src->replay_stride * A, src->mem, false);
  const int t_repf = stg.plan(src->kind == MAMCTS_SRC_REPLAY && f32 ? src->replay_actions_f32 : nullptr,
                              sizeof(float) * (size_t)n * src->replay_stride * A, src->mem, false);
  const int t_reps = stg.plan(src->kind == MAMCTS_SRC_REPLAY ? src->replay_steps : nullptr,
                              sizeof(uint32_t) * n, src->mem, false);
  const int t_shi = stg.plan((src->kind == MAMCTS_SRC_PHILOX || src->kind == MAMCTS_SRC_HYPOTHESIS) ? src->stream_hi : nullptr,
                             sizeof(uint32_t) * n, src->mem, false);
  const bool philox_like = src->kind == MAMCTS_SRC_PHILOX || src->kind == MAMCTS_SRC_HYPOTHESIS;
  const int t_slo = stg.plan(philox_like ? src->stream_lo : nullptr, sizeof(uint32_t) * n, src->mem, false);
  const bool hyp = src->kind == MAMCTS_SRC_HYPOTHESIS;
  const int t_hlo = stg.plan(hyp ? (f32 ? (const void*)src->hyp_gap_lo_f32 : (const void*)src->hyp_gap_lo) : nullptr,
                             sizeof(int32_t) * (A - 1) * n, src->mem, false);
  const int t_hhi = stg.plan(hyp ? (f32 ? (const void*)src->hyp_gap_hi_f32 : (const void*)src->hyp_gap_hi) : nullptr,
                             sizeof(int32_t) * (A - 1) * n, src->mem, false);
  const int o_ego = stg.plan(out->ego_return, sizeof(double) * n, out->mem, true);
  const int o_oth = stg.plan(out->other_return, sizeof(double) * (A - 1) * n, out->mem, true);
  const int o_cost = stg.plan(out->ego_cost, sizeof(double) * n, out->mem, true);
  const int o_len = stg.plan(out->step_length, sizeof(double) * n, out->mem, true);
  const int o_tot = stg.plan(out->total_steps, sizeof(uint64_t), out->mem, true);
  const int o_steps = stg.plan(out->steps, sizeof(uint32_t) * total, out->mem, true);
  const int o_fx = stg.plan(f32 ? (void*)out->final_x_f32 : (void*)out->final_x, sizeof(int32_t) * x_rows * total, out->mem, true);
  const int o_fl = stg.plan(f32 ? (void*)out->final_last_action_f32 : (void*)out->final_last_action,
                            sizeof(int32_t) * A * total, out->mem, true);
  const int o_ff = stg.plan(out->final_flags, total, out->mem, true);
  const int o_tr = stg.plan(out->reward_trace, sizeof(double) * (size_t)total * out->trace_stride, out->mem, true);
  const int o_mom = stg.plan(out->moments, sizeof(double) * (2 * (size_t)A + 1), out->mem, true);
  if (out->moments && K == 1 && (!out->ego_return || (A > 1 && !out->other_return)))
    return fail(e, MAMCTS_ERR_INVALID, "rollout: moments with K == 1 need ego_return and other_return as well");
  int rc = stg.commit();
  if (rc != MAMCTS_OK) return rc;

  RolloutParams p;
  std::memset(&p, 0, sizeof(p));
  p.n_leaves = n; p.K = K; p.total = total;
  p.x = stg.dev<const void>(t_x);
  p.last = stg.dev<const void>(t_last);
  p.flags = stg.dev<const uint8_t>(t_flags);
  p.depth = stg.dev<const uint32_t>(t_depth);
  p.replay = stg.dev<const int8_t>(t_rep);
  p.replay_f = stg.dev<const float>(t_repf);
  p.replay_steps = stg.dev<const uint32_t>(t_reps);
  p.replay_stride = src->replay_stride;
  for (int a = 0; a < MAMCTS_MAX_AGENTS; ++a) p.const_actions[a] = src->const_actions[a];
  p.key0 = (uint32_t)src->philox_seed;
  p.key1 = (uint32_t)(src->philox_seed >> 32);
  p.stream_hi = stg.dev<const uint32_t>(t_shi);
  p.stream_lo = stg.dev<const uint32_t>(t_slo);
  p.hyp_lo = stg.dev<const int32_t>(t_hlo);
  p.hyp_hi = stg.dev<const int32_t>(t_hhi);
  p.hyp_lo_f = stg.dev<const float>(t_hlo);
  p.hyp_hi_f = stg.dev<const float>(t_hhi);
  p.stream_hi_base = src->stream_hi_base;
  p.stream_lo_base = src->stream_lo_base;
  {
    const uint32_t n_ego = (env_kind == MAMCTS_ENV_SIMPLE) ? 2u : ep.n_ego_actions;
    const uint32_t max_steps = std::min(mp->rh_max_number_of_iterations, mp->max_search_depth);
    if (max_steps > (1u << 24))
      return fail(e, MAMCTS_ERR_UNSUPPORTED, "rollout: min(random_heuristic.MAX_NUMBER_OF_ITERATIONS, MAX_SEARCH_DEPTH) above 2^24");
    rc = ensure_discount_table(e, mp->discount_factor, 1.0 / (double)n_ego, max_steps, &p.disc, &p.disc_len);
    if (rc != MAMCTS_OK) return rc;
  }
  p.rh_cap = mp->rh_max_number_of_iterations;
  p.max_depth = mp->max_search_depth;
  p.env = ep;
  p.steps = stg.dev<uint32_t>(o_steps);
  p.final_x = stg.dev<void>(o_fx);
  p.final_last = stg.dev<void>(o_fl);
  p.final_flags = stg.dev<uint8_t>(o_ff);
  p.trace = stg.dev<double>(o_tr);
  p.trace_stride = out->trace_stride;
  p.total_steps = reinterpret_cast<unsigned long long*>(stg.dev<uint64_t>(o_tot));
  if (p.total_steps) MAMCTS_CUDA(e, cudaMemsetAsync(p.total_steps, 0, sizeof(uint64_t), e->stream));
  if (p.trace) MAMCTS_CUDA(e, cudaMemsetAsync(p.trace, 0, sizeof(double) * (size_t)total * out->trace_stride, e->stream));

  double* d_ego = stg.dev<double>(o_ego);
  double* d_oth = stg.dev<double>(o_oth);
  double* d_cost = stg.dev<double>(o_cost);
  double* d_len = stg.dev<double>(o_len);
  const uint32_t n_arrays = A + 2;  // ego, A-1 others, cost, step_len
  if (K == 1) {
    p.ego_ret = d_ego; p.other_ret = d_oth; p.cost = d_cost; p.step_len = d_len;
  } else {
    rc = ensure_scratch(e, sizeof(double) * (size_t)n_arrays * total);
    if (rc != MAMCTS_OK) return rc;
    double* s = static_cast<double*>(e->scratch);
    p.ego_ret = s;
    p.other_ret = s + (size_t)total;
    p.cost = s + (size_t)A * total;
    p.step_len = s + (size_t)(A + 1) * total;
  }
  const bool parity = p.steps || p.final_x || p.final_last || p.final_flags || p.trace;

  if (env_kind == MAMCTS_ENV_SIMPLE) rc = launch_rollout<SimpleEnv>(e, p, src->kind, parity);
  else if (f32) rc = dispatch_crossing_f32(e, p, (int)A - 1, src->kind, parity);
  else rc = dispatch_crossing(e, p, (int)A - 1, src->kind, parity);
  if (rc != MAMCTS_OK) return rc;

  if (out->moments) {   // over the per-rollout values: the outputs themselves (K = 1) or the scratch rows
    rc = rollout_moments_device(e, p.ego_ret, p.other_ret, A - 1, total, stg.dev<double>(o_mom));
    if (rc != MAMCTS_OK) return rc;
  }
  if (K > 1) {
    const double* s = static_cast<const double*>(e->scratch);
    const int rgrid = std::max(1, std::min(grid_for(e, 4), (int)((n + 7) / 8)));
    // scratch rows: 0 ego, 1..A-1 others, A cost, A+1 step length (all `total` apart): one launch
    ReduceDst dsts;
    std::memset(&dsts, 0, sizeof(dsts));
    dsts.p[0] = d_ego;
    for (uint32_t j = 0; j + 1 < A; ++j) dsts.p[1 + j] = d_oth ? d_oth + (size_t)j * n : nullptr;
    dsts.p[A] = d_cost;
    dsts.p[A + 1] = d_len;
    reduce_mean_kernel<<<dim3(rgrid, n_arrays), 256, 0, e->stream>>>(s, dsts, n, K, total);
    e->launches++;
    MAMCTS_CUDA(e, cudaGetLastError());
  }
  return stg.finish();
}

static int rollout_common(mamcts_engine* e, int env_kind, const mamcts_params* mp, const EnvParams& ep,
                          uint32_t A, uint32_t n, uint32_t K, int32_t mem, const void* state_x,
                          const void* state_last, const uint8_t* state_flags,
                          const uint32_t* start_depth, const mamcts_action_source* src,
                          mamcts_rollout_out* out) {
  if (!e) return fail(nullptr, MAMCTS_ERR_INVALID, "rollout: null engine");
  std::lock_guard<std::mutex> lock(e->mu);
  return rollout_locked(e, env_kind, mp, ep, A, n, K, mem, state_x, state_last, state_flags, start_depth, src, out);
}

// ---- per-ego-action heuristic: ActionValueRandomHeuristic (action_value_random_heuristic.h:27-112) -----------
// For every leaf and every ego action a: one first step (ego = a, the other agents as a fresh statistic makes
// them act) and one RandomHeuristic::rollout from the resulting state at depth + 1 (:37-61). The rollouts are
// exactly the batched rollout kernel above over n * nA start states; two small kernels sit around it:
//   av_first_step_kernel   item (leaf, a) -> the stepped state, its depth, the first step's reward and cost
//   av_combine_kernel      leaf -> action_return / action_cost / action_step_length per action (:64-84), the
//                          other agents' returns (of the LAST action: :58-60 ties the rollout's map to the
//                          accumulator) and the mean cost in the reference's summation order (:97-103)
struct ActionValueParams {
  uint32_t n_leaves, nA;
  const int32_t* x;          // [A][n]
  const int32_t* last;       // [A][n] or null (2)
  const uint8_t* flags;      // [n] or null
  const uint32_t* depth;     // [n] or null
  int kind;
  uint32_t const_actions[MAMCTS_MAX_AGENTS];
  uint32_t key0, key1;
  const uint32_t* stream_hi;
  const uint32_t* stream_lo;
  uint32_t stream_hi_base, stream_lo_base;
  EnvParams env;
  // per item (leaf * nA + a)
  int32_t* x2;               // [A][n * nA]
  int32_t* last2;            // [A][n * nA]
  uint8_t* flags2;           // [n * nA]
  uint32_t* depth2;          // [n * nA]
  uint32_t* s_hi2;           // [n * nA] Philox stream of the item's rollout
  uint32_t* s_lo2;
  double* r0;                // [n * nA] first-step ego reward
  double* c0;                // [n * nA] first-step cost[0]
};

template <class Env>
__global__ void __launch_bounds__(256) av_first_step_kernel(const __grid_constant__ ActionValueParams p) {
  constexpr int A = Env::A;
  const uint32_t total = p.n_leaves * p.nA;
  for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < total; item += gridDim.x * blockDim.x) {
    const uint32_t leaf = item / p.nA, a = item - leaf * p.nA;
    typename Env::State st;
#pragma unroll
    for (int i = 0; i < A; ++i) {
      st.x[i] = __ldg(p.x + (size_t)i * p.n_leaves + leaf);
      st.last[i] = p.last ? __ldg(p.last + (size_t)i * p.n_leaves + leaf) : 2;
    }
    st.flags = p.flags ? (uint32_t)(__ldg(p.flags + leaf) & 1u) : 0u;
    const uint32_t depth = p.depth ? __ldg(p.depth + leaf) : 0u;
    const uint32_t s_hi = p.stream_hi ? __ldg(p.stream_hi + leaf) : p.stream_hi_base;
    const uint32_t s_lo = (p.stream_lo ? __ldg(p.stream_lo + leaf) : p.stream_lo_base + leaf) * p.nA + a;
    double r0 = 0.0, c0 = 0.0;
    if (!Env::terminal(p.env, st)) {                      // :38: a terminal state tries no action
      int32_t act[A];
      act[0] = (int32_t)a;                                // :41
      if (p.kind == MAMCTS_SRC_PHILOX) {
        // (step 0, agent j) of stream (s_hi ^ 0x80000000, s_lo): same layout as the rollout's draws
        const uint32_t hi2 = s_hi ^ 0x80000000u;
        Philox4 blk[(A + 7) / 8];
#pragma unroll
        for (int b = 0; b < (A + 7) / 8; ++b) blk[b] = philox4x32_10((A <= 8 ? 0 : b), s_lo, hi2, 0u, p.key0, p.key1);
        uint32_t h[A], span[A];
#pragma unroll
        for (int i = 0; i < A; ++i) {
          h[i] = (A <= 8) ? philox_half(blk[0], i) : philox_half(blk[i / 8], i % 8);
          span[i] = Env::num_actions(p.env, i);
        }
        philox_bounded_all<A>(h, span, 0u, s_lo, hi2, p.key0, p.key1);
#pragma unroll
        for (int i = 1; i < A; ++i) act[i] = Env::action_from_index(i, h[i]);
      } else {
#pragma unroll
        for (int i = 1; i < A; ++i) act[i] = Env::action_from_index(i, p.const_actions[i]);   // :43-47
      }
      double r_other;
      Env::step(p.env, st, act, r0, r_other, c0);         // :52
    }
#pragma unroll
    for (int i = 0; i < A; ++i) {
      p.x2[(size_t)i * total + item] = st.x[i];
      p.last2[(size_t)i * total + item] = st.last[i];
    }
    p.flags2[item] = (uint8_t)(st.flags & 1u);
    p.depth2[item] = depth + 1u;
    p.s_hi2[item] = s_hi;
    p.s_lo2[item] = s_lo;
    p.r0[item] = r0;
    p.c0[item] = c0;
  }
}

struct ActionValueCombine {
  uint32_t n_leaves, nA, num_others;
  double gamma;
  const uint8_t* flags;      // leaf flags or null
  const double* r0;          // per item
  const double* c0;
  const double* ego_roll;    // per item: the rollout's outputs (K = 1)
  const double* other_roll;  // [num_others][n * nA]
  const double* cost_roll;
  const double* len_roll;
  double* action_return;     // [n][nA]
  double* action_cost;
  double* action_len;
  double* other_return;      // [num_others][n]
  double* mean_cost;         // [n]
  unsigned long long* total_steps;   // += one first step per item of a non-terminal leaf
};

__global__ void __launch_bounds__(256) av_combine_kernel(const __grid_constant__ ActionValueCombine p) {
  unsigned long long first_steps = 0;
  for (uint32_t leaf = blockIdx.x * blockDim.x + threadIdx.x; leaf < p.n_leaves; leaf += gridDim.x * blockDim.x) {
    const bool terminal = p.flags && (p.flags[leaf] & 1u);
    const size_t total = (size_t)p.n_leaves * p.nA;
    double mean = 0.0;
    for (uint32_t k = 0; k < p.nA; ++k) {                  // :97-103 in the reference container's order: last first
      const uint32_t a = p.nA - 1u - k;
      const size_t item = (size_t)leaf * p.nA + a;
      double ret = 0.0, cost = 0.0, len = 0.0;
      if (!terminal) {
        ret = __dadd_rn(0.0, __dadd_rn(__dmul_rn(p.gamma, p.r0[item]), __dmul_rn(p.gamma, p.ego_roll[item])));   // :64-65
        cost = __dadd_rn(0.0, __dadd_rn(p.c0[item], p.cost_roll[item]));                                          // :73-81
        len = __dadd_rn(0.0, __dadd_rn(1.0, p.len_roll[item]));                                                   // :83-84
      }
      if (p.action_return) p.action_return[item] = ret;
      if (p.action_cost) p.action_cost[item] = cost;
      if (p.action_len) p.action_len[item] = len;
      mean = __dadd_rn(mean, cost);
    }
    if (p.mean_cost) p.mean_cost[leaf] = __ddiv_rn(mean, (double)p.nA);
    if (p.other_return) {
      const size_t last_item = (size_t)leaf * p.nA + (p.nA - 1u);
      for (uint32_t j = 0; j < p.num_others; ++j)   // :66-71 on the tied map; rewards[1..] of CrossingState are 0
        p.other_return[(size_t)j * p.n_leaves + leaf] =
            terminal ? 0.0 : __dadd_rn(p.other_roll[(size_t)j * total + last_item],
                                       __dadd_rn(__dmul_rn(p.gamma, 0.0), __dmul_rn(p.gamma, 0.0)));
    }
    if (!terminal) first_steps += p.nA;
  }
  if (p.total_steps && first_steps) atomicAdd(p.total_steps, first_steps);
}

template <class Env>
static void launch_av_first_step(mamcts_engine* e, const ActionValueParams& p) {
  const uint32_t total = p.n_leaves * p.nA;
  const int grid = (int)std::max<uint32_t>(1u, std::min<uint32_t>((total + 255) / 256, (uint32_t)e->num_sms * 8u));
  av_first_step_kernel<Env><<<grid, 256, 0, e->stream>>>(p);
}

}  // namespace mamcts

using namespace mamcts;

extern "C" {

int mamcts_action_values_crossing_i32(mamcts_engine* e, const mamcts_params* mp, const mamcts_crossing_params* cp,
                                      uint32_t n_leaves, int32_t mem, const int32_t* state_x,
                                      const int32_t* state_last_action, const uint8_t* state_flags,
                                      const uint32_t* start_depth, const mamcts_action_source* src,
                                      mamcts_action_values_out* out) {
  if (!e) return fail(nullptr, MAMCTS_ERR_INVALID, "mamcts_action_values_crossing_i32: null engine");
  std::lock_guard<std::mutex> lock(e->mu);
  if (!mp || !cp || !src || !out || !state_x) return fail(e, MAMCTS_ERR_INVALID, "mamcts_action_values_crossing_i32: null argument");
  if (cp->num_other_agents + 1 > MAMCTS_MAX_AGENTS || cp->num_other_agents == 0)
    return fail(e, MAMCTS_ERR_UNSUPPORTED, "mamcts_action_values_crossing_i32: 1..15 other agents supported");
  if (src->kind != MAMCTS_SRC_REFERENCE_CONST && src->kind != MAMCTS_SRC_PHILOX)
    return fail(e, MAMCTS_ERR_INVALID, "mamcts_action_values_crossing_i32: action source must be REFERENCE_CONST or PHILOX");
  if (n_leaves == 0) return MAMCTS_OK;
  EnvParams ep;
  std::memset(&ep, 0, sizeof(ep));
  make_crossing_env(*cp, ep);
  if (ep.n_ego_actions == 0 || ep.n_ego_actions > MAMCTS_MAX_ACTIONS || ep.n_other_actions == 0 ||
      ep.n_other_actions > MAMCTS_MAX_ACTIONS)
    return fail(e, MAMCTS_ERR_INVALID, "mamcts_action_values_crossing_i32: action counts out of range");
  const uint32_t A = cp->num_other_agents + 1, nA = ep.n_ego_actions, n = n_leaves;
  if ((uint64_t)n * nA > 0x40000000ull) return fail(e, MAMCTS_ERR_INVALID, "mamcts_action_values_crossing_i32: too many leaves for one call");
  if (src->kind == MAMCTS_SRC_REFERENCE_CONST)
    for (uint32_t a = 1; a < A; ++a)
      if (src->const_actions[a] >= ep.n_other_actions) return fail(e, MAMCTS_ERR_INVALID, "mamcts_action_values_crossing_i32: const action out of range");
  MAMCTS_CUDA(e, cudaSetDevice(e->device));
  const uint32_t total = n * nA;

  ArgStager stg(e);
  const int t_x = stg.plan(state_x, sizeof(int32_t) * A * n, mem, false);
  const int t_last = stg.plan(state_last_action, sizeof(int32_t) * A * n, mem, false);
  const int t_flags = stg.plan(state_flags, n, mem, false);
  const int t_depth = stg.plan(start_depth, sizeof(uint32_t) * n, mem, false);
  const int t_shi = stg.plan(src->kind == MAMCTS_SRC_PHILOX ? src->stream_hi : nullptr, sizeof(uint32_t) * n, src->mem, false);
  const int t_slo = stg.plan(src->kind == MAMCTS_SRC_PHILOX ? src->stream_lo : nullptr, sizeof(uint32_t) * n, src->mem, false);
  const int o_ret = stg.plan(out->action_return, sizeof(double) * total, out->mem, true);
  const int o_cost = stg.plan(out->action_cost, sizeof(double) * total, out->mem, true);
  const int o_len = stg.plan(out->action_step_length, sizeof(double) * total, out->mem, true);
  const int o_oth = stg.plan(out->other_return, sizeof(double) * (A - 1) * n, out->mem, true);
  const int o_mean = stg.plan(out->mean_cost, sizeof(double) * n, out->mem, true);
  const int o_tot = stg.plan(out->total_steps, sizeof(uint64_t), out->mem, true);
  int rc = stg.commit();
  if (rc != MAMCTS_OK) return rc;

  // scratch: the stepped states and the rollouts' per-item results
  auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
  const size_t b_i32 = up(sizeof(int32_t) * (size_t)A * total), b_u32 = up(sizeof(uint32_t) * (size_t)total),
               b_u8 = up(total), b_f64 = up(sizeof(double) * (size_t)total);
  rc = ensure_scratch(e, 2 * b_i32 + 3 * b_u32 + b_u8 + (5 + (size_t)(A - 1)) * b_f64);
  if (rc != MAMCTS_OK) return rc;
  char* base = static_cast<char*>(e->scratch);
  auto take = [&](size_t bytes) { char* ptr = base; base += bytes; return ptr; };
  ActionValueParams ap;
  std::memset(&ap, 0, sizeof(ap));
  ap.n_leaves = n; ap.nA = nA;
  ap.x = stg.dev<const int32_t>(t_x); ap.last = stg.dev<const int32_t>(t_last);
  ap.flags = stg.dev<const uint8_t>(t_flags); ap.depth = stg.dev<const uint32_t>(t_depth);
  ap.kind = src->kind;
  for (int a = 0; a < MAMCTS_MAX_AGENTS; ++a) ap.const_actions[a] = src->const_actions[a];
  ap.key0 = (uint32_t)src->philox_seed; ap.key1 = (uint32_t)(src->philox_seed >> 32);
  ap.stream_hi = stg.dev<const uint32_t>(t_shi); ap.stream_lo = stg.dev<const uint32_t>(t_slo);
  ap.stream_hi_base = src->stream_hi_base; ap.stream_lo_base = src->stream_lo_base;
  ap.env = ep;
  ap.x2 = reinterpret_cast<int32_t*>(take(b_i32));
  ap.last2 = reinterpret_cast<int32_t*>(take(b_i32));
  ap.depth2 = reinterpret_cast<uint32_t*>(take(b_u32));
  ap.s_hi2 = reinterpret_cast<uint32_t*>(take(b_u32));
  ap.s_lo2 = reinterpret_cast<uint32_t*>(take(b_u32));
  ap.flags2 = reinterpret_cast<uint8_t*>(take(b_u8));
  ap.r0 = reinterpret_cast<double*>(take(b_f64));
  ap.c0 = reinterpret_cast<double*>(take(b_f64));
  double* ego_roll = reinterpret_cast<double*>(take(b_f64));
  double* cost_roll = reinterpret_cast<double*>(take(b_f64));
  double* len_roll = reinterpret_cast<double*>(take(b_f64));
  double* other_roll = reinterpret_cast<double*>(take((size_t)(A - 1) * b_f64));   // rows `total` apart
  switch (A - 1) {
    case 1: launch_av_first_step<CrossingEnv<1>>(e, ap); break;
    case 2: launch_av_first_step<CrossingEnv<2>>(e, ap); break;
    case 3: launch_av_first_step<CrossingEnv<3>>(e, ap); break;
    case 4: launch_av_first_step<CrossingEnv<4>>(e, ap); break;
    case 5: launch_av_first_step<CrossingEnv<5>>(e, ap); break;
    case 6: launch_av_first_step<CrossingEnv<6>>(e, ap); break;
    case 7: launch_av_first_step<CrossingEnv<7>>(e, ap); break;
    case 11: launch_av_first_step<CrossingEnv<11>>(e, ap); break;
    case 15: launch_av_first_step<CrossingEnv<15>>(e, ap); break;
    default: return fail(e, MAMCTS_ERR_UNSUPPORTED, "mamcts_action_values_crossing_i32: compiled for 1..7, 11 or 15 other agents");
  }
  e->launches++;
  MAMCTS_CUDA(e, cudaGetLastError());

  // the rollouts: n * nA start states, one rollout each, on the same stream (device buffers throughout)
  mamcts_action_source rs = *src;
  rs.mem = MAMCTS_MEM_DEVICE;
  rs.stream_hi = ap.s_hi2;
  rs.stream_lo = ap.s_lo2;
  mamcts_rollout_out ro;
  std::memset(&ro, 0, sizeof(ro));
  ro.mem = MAMCTS_MEM_DEVICE;
  ro.ego_return = ego_roll; ro.other_return = other_roll; ro.ego_cost = cost_roll; ro.step_length = len_roll;
  uint64_t* d_tot = stg.dev<uint64_t>(o_tot);
  ro.total_steps = d_tot;
  // (the rollout kernel writes other j of item r at other_roll[j * total + r]: rows `total` doubles apart)
  rc = rollout_locked(e, MAMCTS_ENV_CROSSING_I32, mp, ep, A, total, 1, MAMCTS_MEM_DEVICE, ap.x2, ap.last2, ap.flags2,
                      ap.depth2, &rs, &ro);
  if (rc != MAMCTS_OK) return rc;
  ActionValueCombine cb;
  std::memset(&cb, 0, sizeof(cb));
  cb.n_leaves = n; cb.nA = nA; cb.num_others = A - 1; cb.gamma = mp->discount_factor;
  cb.flags = ap.flags; cb.r0 = ap.r0; cb.c0 = ap.c0;
  cb.ego_roll = ego_roll; cb.other_roll = other_roll; cb.cost_roll = cost_roll; cb.len_roll = len_roll;
  cb.action_return = stg.dev<double>(o_ret); cb.action_cost = stg.dev<double>(o_cost);
  cb.action_len = stg.dev<double>(o_len); cb.other_return = stg.dev<double>(o_oth);
  cb.mean_cost = stg.dev<double>(o_mean);
  cb.total_steps = reinterpret_cast<unsigned long long*>(d_tot);
  const int grid = (int)std::max<uint32_t>(1u, std::min<uint32_t>((n + 255) / 256, (uint32_t)e->num_sms * 8u));
  av_combine_kernel<<<grid, 256, 0, e->stream>>>(cb);
  e->launches++;
  MAMCTS_CUDA(e, cudaGetLastError());
  return stg.finish();
}

int mamcts_rollout_crossing_i32(mamcts_engine* e, const mamcts_params* mp,
                                const mamcts_crossing_params* cp, uint32_t n_leaves,
                                uint32_t rollouts_per_leaf, int32_t mem, const int32_t* state_x,
                                const int32_t* state_last_action, const uint8_t* state_flags,
                                const uint32_t* start_depth, const mamcts_action_source* src,
                                mamcts_rollout_out* out) {
  if (!cp) return fail(e, MAMCTS_ERR_INVALID, "mamcts_rollout_crossing_i32: null crossing params");
  if (cp->num_other_agents + 1 > MAMCTS_MAX_AGENTS || cp->num_other_agents == 0)
    return fail(e, MAMCTS_ERR_UNSUPPORTED, "mamcts_rollout_crossing_i32: 1..15 other agents supported");
  EnvParams ep;
  std::memset(&ep, 0, sizeof(ep));
  make_crossing_env(*cp, ep);
  if (ep.n_ego_actions == 0 || ep.n_ego_actions > MAMCTS_MAX_ACTIONS || ep.n_other_actions == 0 ||
      ep.n_other_actions > MAMCTS_MAX_ACTIONS)
    return fail(e, MAMCTS_ERR_INVALID, "mamcts_rollout_crossing_i32: action counts out of range");
  return rollout_common(e, MAMCTS_ENV_CROSSING_I32, mp, ep, cp->num_other_agents + 1, n_leaves,
                        rollouts_per_leaf, mem, state_x, state_last_action, state_flags, start_depth,
                        src, out);
}

int mamcts_rollout_crossing_f32(mamcts_engine* e, const mamcts_params* mp,
                                const mamcts_crossing_params_f32* cp, uint32_t n_leaves,
                                uint32_t rollouts_per_leaf, int32_t mem, const float* state_x,
                                const float* state_last_action, const uint8_t* state_flags,
                                const uint32_t* start_depth, const mamcts_action_source* src,
                                mamcts_rollout_out* out) {
  if (!cp) return fail(e, MAMCTS_ERR_INVALID, "mamcts_rollout_crossing_f32: null crossing params");
  if (cp->num_other_agents + 1 > MAMCTS_MAX_AGENTS || cp->num_other_agents == 0)
    return fail(e, MAMCTS_ERR_UNSUPPORTED, "mamcts_rollout_crossing_f32: 1..15 other agents supported");
  EnvParams ep;
  std::memset(&ep, 0, sizeof(ep));
  make_crossing_env_f32(*cp, ep);
  if (ep.n_ego_actions == 0 || ep.n_ego_actions > MAMCTS_MAX_ACTIONS || ep.n_other_actions == 0 ||
      ep.n_other_actions > MAMCTS_MAX_ACTIONS)
    return fail(e, MAMCTS_ERR_INVALID, "mamcts_rollout_crossing_f32: action counts out of range");
  return rollout_common(e, kEnvCrossingF32, mp, ep, cp->num_other_agents + 1, n_leaves, rollouts_per_leaf, mem,
                        state_x, state_last_action, state_flags, start_depth, src, out);
}

int mamcts_rollout_simple(mamcts_engine* e, const mamcts_params* mp, const mamcts_simple_params* sp,
                          uint32_t n_leaves, uint32_t rollouts_per_leaf, int32_t mem,
                          const int32_t* state_len, const uint32_t* start_depth,
                          const mamcts_action_source* src, mamcts_rollout_out* out) {
  if (!sp) return fail(e, MAMCTS_ERR_INVALID, "mamcts_rollout_simple: null simple params");
  EnvParams ep;
  std::memset(&ep, 0, sizeof(ep));
  ep.win_len = sp->winning_state_length;
  ep.lose_len = sp->loosing_state_length;
  ep.n_ego_actions = 2;
  ep.n_other_actions = 2;
  return rollout_common(e, MAMCTS_ENV_SIMPLE, mp, ep, 2, n_leaves, rollouts_per_leaf, mem, state_len,
                        nullptr, nullptr, start_depth, src, out);
}

}  // extern "C"
