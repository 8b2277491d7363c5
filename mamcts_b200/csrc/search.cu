// search.cu - device-resident root-parallel MCTS: T independent trees live in HBM, one lane owns one
// tree and runs whole iterations (select/expand -> rollout -> backup) without ever returning to the
// host. Replaces, for T trees at once, Mcts<S,UctStatistic,UctStatistic,RandomHeuristic>::search
// (reference mcts/mcts.h:166-186,294-333), StageNode::select_or_expand / update_statistics
// (mcts/stage_node.h:167-271) and UctStatistic (mcts/statistics/uct_statistic.h).
//
// HBM layout: per tree a contiguous arena of fixed-size node records (AoS: a lane
// that touches a node uses every byte of the lines it pulls). Children are found through a direct
// per-node table when the joint-action space is small (2 agents) and otherwise through a per-tree
// open-addressing table (uint32 child index per slot, keyed by hash(parent, joint action), verified
// against the child's own header). Statistics are FP64 and are updated in the reference's operation
// order; log(N) comes from a host-computed table and the widening thresholds from host libm, so a
// REFERENCE_CONST search reproduces the reference's tree bit for bit.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "engine.h"
#include "root_merge.h"
#include "search_impl.cuh"


namespace mamcts {

// CrossingState with 4, 5 or 6 other agents: instantiated in search_more.cu
mamcts_search* make_search_more(const mamcts_search_config& cfg, uint32_t n_ego, uint32_t n_oth, uint32_t max_nodes);

static mamcts_search* make_search(const mamcts_search_config& cfg, uint32_t n_ego, uint32_t n_oth,
                                  uint32_t max_nodes) {
  // layout: 0 = automatic (compact wherever a direct child table exists), 1 = classic, 2 = compact
  const bool compact = cfg.layout != MAMCTS_LAYOUT_CLASSIC;
  if (cfg.env == MAMCTS_ENV_SIMPLE) return make_search_cidx<SimpleEnv, 2, 2, 4>(max_nodes, compact, cfg.mcts.max_number_of_iterations);
  // more than two agents: the counts-only compact layout (search_compact_multi.cuh) wherever the other agents'
  // values are provably +0.0 (CrossingState, discount >= 0) and leaf / inner indices fit 15 bits
  const bool multi = compact && cfg.env == MAMCTS_ENV_CROSSING_I32 && cfg.mcts.discount_factor >= 0.0 && max_nodes < 32767;
  if (multi && n_ego == 4 && n_oth == 7) {
    switch (cfg.crossing.num_other_agents) {
      case 2: return make_search_multi<CrossingEnv<2>, 4, 7>(cfg.mcts.max_number_of_iterations);
      case 3: return make_search_multi<CrossingEnv<3>, 4, 7>(cfg.mcts.max_number_of_iterations);
      case 7: return make_search_multi<CrossingEnv<7>, 4, 7>(cfg.mcts.max_number_of_iterations);
      default: break;
    }
  } else if (multi && n_ego <= 8 && n_oth <= 8) {
    switch (cfg.crossing.num_other_agents) {
      case 2: return make_search_multi<CrossingEnv<2>, 8, 8>(cfg.mcts.max_number_of_iterations);
      case 3: return make_search_multi<CrossingEnv<3>, 8, 8>(cfg.mcts.max_number_of_iterations);
      default: break;
    }
  }
  if (cfg.env == MAMCTS_ENV_CROSSING_I32 && cfg.crossing.num_other_agents >= 4 && cfg.crossing.num_other_agents <= 6)
    return make_search_more(cfg, n_ego, n_oth, max_nodes);
  if (n_ego == 4 && n_oth == 7) {
    switch (cfg.crossing.num_other_agents) {
      case 1: return make_search_cidx<CrossingEnv<1>, 4, 7, 28>(max_nodes, compact, cfg.mcts.max_number_of_iterations);
      case 2: return make_search_cidx<CrossingEnv<2>, 4, 7, 0>(max_nodes, false, cfg.mcts.max_number_of_iterations);
      case 3: return make_search_cidx<CrossingEnv<3>, 4, 7, 0>(max_nodes, false, cfg.mcts.max_number_of_iterations);
      case 7: return make_search_cidx<CrossingEnv<7>, 4, 7, 0>(max_nodes, false, cfg.mcts.max_number_of_iterations);
      default: return nullptr;
    }
  }
  // any other action counts up to 8 per agent (CrossingStateParameters lets the velocity ranges be
  // chosen freely, crossing_state_parameters.h:15-33): records sized for 8 actions, the real counts are
  // run-time values (widening order, expansion masks)
  if (n_ego <= 8 && n_oth <= 8) {
    switch (cfg.crossing.num_other_agents) {
      case 1: return make_search_cidx<CrossingEnv<1>, 8, 8, 64>(max_nodes, compact, cfg.mcts.max_number_of_iterations);
      case 2: return make_search_cidx<CrossingEnv<2>, 8, 8, 0>(max_nodes, false, cfg.mcts.max_number_of_iterations);
      case 3: return make_search_cidx<CrossingEnv<3>, 8, 8, 0>(max_nodes, false, cfg.mcts.max_number_of_iterations);
      default: return nullptr;
    }
  }
  return nullptr;
}

void free_search(mamcts_search* s) {
  if (!s) return;
  cudaSetDevice(s->e->device);
  cudaStreamSynchronize(s->e->stream);
  s->free_extra();
  cudaFree(s->d_inner); cudaFree(s->d_leaves); cudaFree(s->d_tree_inner); cudaFree(s->d_tree_leaves);
  cudaFree(s->d_nodes); cudaFree(s->d_hash); cudaFree(s->d_tree_nodes); cudaFree(s->d_tree_iters);
  cudaFree(s->d_tab); cudaFree(s->d_log); cudaFree(s->d_disc); cudaFree(s->d_counters); cudaFree(s->d_root_x);
  delete s;
}

// Device-side failures of the compact layout are counted (search_compact.cuh); every read-back
// reports them instead of returning statistics of an unfinished search.
static int check_search_health(mamcts_search* s) { return s->health(); }

}  // namespace mamcts

int mamcts_search::upload_root(const int32_t* root_x, const int32_t* /*root_last_action*/) {
  int32_t rx[MAMCTS_MAX_AGENTS] = {0};
  for (uint32_t i = 0; i < state_ints; ++i) rx[i] = root_x[i];
  MAMCTS_CUDA(e, cudaMemcpyAsync(d_root_x, rx, sizeof(rx), cudaMemcpyHostToDevice, e->stream));
  MAMCTS_CUDA(e, cudaStreamSynchronize(e->stream));
  return MAMCTS_OK;
}

int mamcts_search::health() {
  using namespace mamcts;
  mamcts_search* s = this;
  if (!s->is_compact()) return MAMCTS_OK;
  unsigned long long h[2] = {0, 0};
  MAMCTS_CUDA(e, cudaMemcpyAsync(h, s->d_counters + 6, sizeof(h), cudaMemcpyDeviceToHost, e->stream));
  MAMCTS_CUDA(e, cudaStreamSynchronize(e->stream));
  if (h[0])
    return fail(e, MAMCTS_ERR_NOMEM, "mamcts_search: " + std::to_string(h[0]) + " tree(s) ran out of inner node records; "
                "raise mamcts_search_config.max_inner_nodes_per_tree (0 = one per node, always enough)");
  if (h[1])
    return fail(e, MAMCTS_ERR_UNSUPPORTED, "mamcts_search: " + std::to_string(h[1]) + " tree(s) grew deeper than 64 levels; "
                "use layout = MAMCTS_LAYOUT_CLASSIC for such searches");
  return MAMCTS_OK;
}

namespace mamcts {
int dump_classic(mamcts_search* s, uint32_t tree, double* records, uint32_t capacity_records,
                 uint32_t* num_records) {
  mamcts_engine* e = s->e;
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
  if (cfg->layout > MAMCTS_LAYOUT_COMPACT)
    return fail(e, MAMCTS_ERR_INVALID, "mamcts_search_create: unknown layout");
  {
    const uint32_t K = cfg->rollouts_per_leaf;
    if (K == 0 || K > 16 || (K & (K - 1)) != 0)
      return fail(e, MAMCTS_ERR_UNSUPPORTED, "mamcts_search_create: rollouts_per_leaf must be 1, 2, 4, 8 or 16");
    if (K > 1 && cfg->action_source != MAMCTS_SRC_PHILOX)
      return fail(e, MAMCTS_ERR_INVALID, "mamcts_search_create: rollouts_per_leaf > 1 needs MAMCTS_SRC_PHILOX "
                                         "(the reference's rollouts are all the same, SURVEY.md F1)");
  }
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
    make_crossing_env(cp, ep);
  } else {
    return fail(e, MAMCTS_ERR_INVALID, "mamcts_search_create: unknown environment");
  }
  const uint32_t max_iter = mp.max_number_of_iterations;
  uint32_t max_nodes = cfg->max_nodes_per_tree;
  const uint64_t needed = std::min<uint64_t>((uint64_t)mp.max_number_of_nodes + 1, (uint64_t)max_iter + 1);
  if (max_nodes == 0) max_nodes = (uint32_t)needed;
  if (max_nodes < needed)
    return fail(e, MAMCTS_ERR_INVALID, "mamcts_search_create: max_nodes_per_tree below min(MAX_NUMBER_OF_NODES, iterations) + 1");

  mamcts_search* s = make_search(*cfg, n_ego, n_oth, max_nodes);
  if (!s)
    return fail(e, MAMCTS_ERR_UNSUPPORTED,
                "mamcts_search_create: compiled for SimpleState and CrossingState<int> with 1 to 7 other agents "
                "at 4 ego / 7 other actions, and 1, 2 or 3 other agents at any action counts up to 8");
  if (cfg->layout == MAMCTS_LAYOUT_COMPACT && !s->is_compact()) {
    delete s;
    return fail(e, MAMCTS_ERR_UNSUPPORTED, "mamcts_search_create: the compact layouts are built for SimpleState, for CrossingState with 2 agents, "
                "and for 3 to 8 agents with a discount >= 0 and fewer than 32 767 nodes per tree");
  }
  s->e = e;
  s->cfg = *cfg;
  s->A = A;
  {
    const cudaError_t derr = cudaSetDevice(e->device);
    if (derr != cudaSuccess) { delete s; return cuda_fail(e, derr, "mamcts_search_create: cudaSetDevice"); }
  }

  // host tables: widening order (libstdc++), expanded masks, widening thresholds (libm pow), log table
  SearchTables& tab = s->tab;
  std::memset(&tab, 0, sizeof(tab));
  const uint32_t nact[2] = {n_ego, n_oth};
  for (int cls = 0; cls < 2; ++cls) {
    uint32_t order[MAMCTS_MAX_ACTIONS];
    int rc = mamcts_reference_expand_order(mp.random_seed, nact[cls], order);
    if (rc != MAMCTS_OK) { delete s; return fail(e, rc, mamcts_last_error(nullptr)); }   // the message of the failing call
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

  uint32_t hcap = 16;
  if (s->uses_hash()) while (hcap < 2 * max_nodes) hcap <<= 1;

  const uint32_t T = cfg->num_trees;
  cudaError_t err = cudaSuccess;
  auto alloc = [&](void** ptr, size_t bytes) { if (err == cudaSuccess) err = cudaMalloc(ptr, bytes); };
  if (s->is_compact()) {
    s->max_inner = cfg->max_inner_nodes_per_tree ? std::min(cfg->max_inner_nodes_per_tree, max_nodes) : max_nodes;
    s->max_inner = std::min(s->max_inner, 1u << 24);   // the path record keeps an inner index in 24 bits
    alloc(&s->d_inner, s->node_bytes() * (size_t)T * s->max_inner);
    alloc(&s->d_leaves, s->leaf_bytes() * (size_t)T * max_nodes);
    alloc((void**)&s->d_tree_inner, sizeof(uint32_t) * T);
    alloc((void**)&s->d_tree_leaves, sizeof(uint32_t) * T);
  } else {
    alloc(&s->d_nodes, s->node_bytes() * (size_t)T * max_nodes);
  }
  alloc((void**)&s->d_hash, s->uses_hash() ? sizeof(uint32_t) * (size_t)T * hcap : 256);
  alloc((void**)&s->d_tree_nodes, sizeof(uint32_t) * T);
  alloc((void**)&s->d_tree_iters, sizeof(uint32_t) * T);
  alloc((void**)&s->d_tab, sizeof(SearchTables));
  alloc((void**)&s->d_log, sizeof(double) * logs.size());
  const uint32_t max_steps = std::min(mp.rh_max_number_of_iterations, mp.max_search_depth);
  alloc((void**)&s->d_disc, sizeof(double) * 2 * ((size_t)std::min(max_steps, 1u << 24) + 1));
  alloc((void**)&s->d_counters, sizeof(unsigned long long) * 8);
  alloc((void**)&s->d_root_x, sizeof(int32_t) * MAMCTS_MAX_AGENTS);
  if (err == cudaSuccess && s->alloc_extra(max_nodes) != MAMCTS_OK) err = cudaErrorMemoryAllocation;
  if (err == cudaSuccess && max_steps > (1u << 24)) {
    free_search(s);
    return fail(e, MAMCTS_ERR_UNSUPPORTED, "mamcts_search_create: min(random_heuristic.MAX_NUMBER_OF_ITERATIONS, MAX_SEARCH_DEPTH) above 2^24");
  }
  if (err != cudaSuccess) {
    free_search(s);
    cudaGetLastError();
    return fail(e, MAMCTS_ERR_NOMEM, std::string("mamcts_search_create: device allocation failed: ") + cudaGetErrorString(err));
  }
  SearchParams& p = s->p;
  std::memset(&p, 0, sizeof(p));
  p.nodes = s->d_nodes; p.hash = s->d_hash; p.tree_nodes = s->d_tree_nodes; p.tree_iters = s->d_tree_iters;
  p.tab = s->d_tab; p.log_table = s->d_log; p.counters = s->d_counters;
  p.disc = s->d_disc; p.disc_len = max_steps + 1;
  p.T = T; p.max_nodes = max_nodes; p.hash_mask = hcap - 1; p.first_tree_id = cfg->first_tree_id;
  p.n_ego = n_ego; p.n_oth = n_oth;
  p.use_bound_est = mp.use_bound_estimation; p.max_depth = mp.max_search_depth;
  p.node_budget = mp.max_number_of_nodes; p.rh_cap = mp.rh_max_number_of_iterations;
  p.key0 = (uint32_t)cfg->philox_seed; p.key1 = (uint32_t)(cfg->philox_seed >> 32);
  p.rollouts_per_leaf = cfg->rollouts_per_leaf;
  p.gamma = mp.discount_factor; p.two_c = 2 * mp.uct_exploration_constant;
  p.lb = mp.uct_lower_bound; p.ub = mp.uct_upper_bound;
  p.env = ep;
  // Trees per warp: lanes of a warp wait for the deepest path and the longest rollout among them, so a
  // warp should carry as few trees as the GPU has room for - the smallest power of two that still fits
  // all trees into the resident warps (16 per SM). 75 776 trees and more: 32; 8 192 trees (C5 per GPU): 4,
  // +21 % over 32; 1 024 trees: 1, +77 % (profiles/r01_search_sweeps.txt r60). The environment variables
  // are profiling knobs (scripts/sweep_search.py).
  {
    const uint64_t warp_slots = (uint64_t)std::max(e->num_sms, 1) * 16u;
    uint32_t lpw = 1;
    while (lpw < 32 && (uint64_t)lpw * warp_slots < (uint64_t)cfg->num_trees) lpw <<= 1;
    p.lanes_per_warp = lpw;
  }
  if (const char* v = std::getenv("MAMCTS_SEARCH_LPW")) {
    const int l = std::atoi(v);
    if (l >= 1 && l <= 32) p.lanes_per_warp = (uint32_t)l;
  }
  if (const char* v = std::getenv("MAMCTS_SEARCH_THREADS")) {
    const int t = std::atoi(v);
    if (t >= 32 && t <= kSearchThreads && t % 32 == 0) s->block_threads = t;
  }
  s->state_ints = (cfg->env == MAMCTS_ENV_SIMPLE) ? 1 : A;
  int32_t rx[MAMCTS_MAX_AGENTS] = {0};
  for (uint32_t i = 0; i < s->state_ints; ++i) rx[i] = root_x[i];
  err = cudaMemcpyAsync(s->d_tab, &tab, sizeof(tab), cudaMemcpyHostToDevice, e->stream);
  if (err == cudaSuccess) err = cudaMemcpyAsync(s->d_log, logs.data(), sizeof(double) * logs.size(), cudaMemcpyHostToDevice, e->stream);
  if (err == cudaSuccess) err = cudaMemcpyAsync(s->d_root_x, rx, sizeof(rx), cudaMemcpyHostToDevice, e->stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);  // host sources go out of scope
  if (err != cudaSuccess) { free_search(s); return cuda_fail(e, err, "mamcts_search_create uploads"); }
  int rc = fill_discount_table(e, mp.discount_factor, 1.0 / (double)n_ego, max_steps, s->d_disc);
  if (rc == MAMCTS_OK) rc = s->launch_reset();
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
  if (!s) return fail(nullptr, MAMCTS_ERR_INVALID, "mamcts_search_reset: null handle");
  mamcts_engine* e = s->e;
  std::lock_guard<std::mutex> lock(e->mu);
  MAMCTS_CUDA(e, cudaSetDevice(e->device));
  if (root_x) {  // NULL: re-root at the state already resident on the device (no copy)
    const int rc = s->upload_root(root_x, root_last_action);
    if (rc != MAMCTS_OK) return rc;
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
  { int rc = check_search_health(s); if (rc != MAMCTS_OK) return rc; }
  if (iterations) *iterations = h[0];
  if (rollout_steps) *rollout_steps = h[1];
  if (expand_steps) *expand_steps = h[2];
  if (nodes) *nodes = h[3] + s->cfg.num_trees;  // + the roots
  if (backup_levels) *backup_levels = h[4];
  return MAMCTS_OK;
}

int mamcts_search_exact_ucb_evaluations(mamcts_search* s, uint64_t* count) {
  if (!s || !count) return fail(s ? s->e : nullptr, MAMCTS_ERR_INVALID, "mamcts_search_exact_ucb_evaluations: null argument");
  mamcts_engine* e = s->e;
  std::lock_guard<std::mutex> lock(e->mu);
  MAMCTS_CUDA(e, cudaSetDevice(e->device));
  unsigned long long h = 0;
  MAMCTS_CUDA(e, cudaMemcpyAsync(&h, s->d_counters + 5, sizeof(h), cudaMemcpyDeviceToHost, e->stream));
  MAMCTS_CUDA(e, cudaStreamSynchronize(e->stream));
  *count = h;
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
  { int rc = check_search_health(s); if (rc != MAMCTS_OK) return rc; }
  std::vector<char> node(s->node_bytes());
  const char* src = s->root_record(tree);
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
  { int rc = check_search_health(s); if (rc != MAMCTS_OK) return rc; }
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
  { int rc = check_search_health(s); if (rc != MAMCTS_OK) return rc; }
  return s->dump(tree, records, capacity_records, num_records);
}

}  // extern "C"
