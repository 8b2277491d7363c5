// backup.cu - batched UctStatistic back-propagation and the root-parallel merge.
//
// backup_uct_kernel : one lane per (path, agent); every agent's statistic is an independent
//   sequential recurrence along the path (reference mcts/mcts.h:309-329,
//   mcts/statistics/uct_statistic.h:100-110,134-144), so A adjacent lanes walk one path together
//   and read the edge records coalesced.
// (The root-parallel merge, reference uct_statistic.h:165-184, lives in reduce.cu.)
#include <algorithm>
#include <cstring>

#include "engine.h"
#include "staging.h"

namespace mamcts {

struct BackupParams {
  uint32_t n_paths, A, max_actions;
  double gamma;
  const uint32_t* leaf_node;
  const uint8_t* leaf_has_heuristic;
  const double* ego_return;
  const double* other_return;  // [(A-1)*n_paths]
  const uint32_t* path_offset;
  const uint32_t* path_len;
  const uint32_t* edge_node;
  const uint8_t* edge_action;
  const double* edge_reward;
  double* value;
  double* latest;
  uint32_t* visits;
  double* q;
  uint32_t* n;
};

__global__ void __launch_bounds__(256) backup_uct_kernel(const __grid_constant__ BackupParams p) {
  const uint64_t total = (uint64_t)p.n_paths * p.A;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t path = (uint32_t)(t / p.A);
    const uint32_t a = (uint32_t)(t % p.A);
    const uint32_t leaf = p.leaf_node[path];
    const size_t leaf_slot = (size_t)leaf * p.A + a;
    double G;
    if (p.leaf_has_heuristic[path]) {
      // UctStatistic::update_from_heuristic, uct_statistic.h:100-110
      const double h = (a == 0) ? p.ego_return[path] : p.other_return[(size_t)(a - 1) * p.n_paths + path];
      p.value[leaf_slot] = h;
      p.latest[leaf_slot] = h;
      p.visits[leaf_slot] += 1;
      G = h;
    } else {
      G = p.latest[leaf_slot];  // re-visited terminal node: no update (stage_node.h:183-187)
    }
    const uint32_t off = p.path_offset[path];
    const uint32_t len = p.path_len[path];
    for (uint32_t e = 0; e < len; ++e) {
      const size_t idx = (size_t)off + e;
      const size_t slot = (size_t)p.edge_node[idx] * p.A + a;
      const uint32_t act = p.edge_action[idx * p.A + a];
      const double r = p.edge_reward[idx * p.A + a];
      // UctStatistic::update_statistics_from_backpropagated, uct_statistic.h:134-144
      G = __dadd_rn(r, __dmul_rn(p.gamma, G));
      const size_t qa = slot * p.max_actions + act;
      const uint32_t cnt = p.n[qa] + 1;
      p.n[qa] = cnt;
      const double qv = p.q[qa];
      p.q[qa] = __dadd_rn(qv, __ddiv_rn(__dsub_rn(G, qv), (double)cnt));
      const uint32_t nv = p.visits[slot] + 1;
      p.visits[slot] = nv;
      const double v = p.value[slot];
      p.value[slot] = __dadd_rn(v, __ddiv_rn(__dsub_rn(G, v), (double)nv));
      p.latest[slot] = G;
    }
  }
}

// ---- HypothesisStatistic back-propagation (C4) ------------------------------------------------

struct HypBackupParams {
  uint32_t n_paths, J, H, MA, C;
  int32_t chance;
  double gamma;
  const uint32_t* leaf_node;
  const uint8_t* leaf_has_heuristic;
  const double* leaf_cost;
  const uint32_t* path_offset;
  const uint32_t* path_len;
  const uint32_t* edge_node;
  const uint8_t* edge_hyp;
  const uint8_t* edge_action;
  const double* edge_cost;
  double* value;
  double* latest;
  uint32_t* visits;
  uint32_t* expanded;
  uint32_t* visits_hyp;
  double* a_cost;
  uint32_t* a_count;
};

// mean of a cost vector as std::accumulate(begin, end, 0.0) / size() forms it
// (hypothesis_statistic.h:120, :159)
__device__ __forceinline__ double cost_mean(const double* c, uint32_t C) {
  double acc = 0.0;
  for (uint32_t i = 0; i < C; ++i) acc = __dadd_rn(acc, c[i]);
  return __ddiv_rn(acc, (double)C);
}

// one lane per (path, other agent): a sequential recurrence along the path, like backup_uct_kernel
__global__ void __launch_bounds__(256) backup_hypothesis_kernel(const __grid_constant__ HypBackupParams p) {
  const uint64_t total = (uint64_t)p.n_paths * p.J;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t path = (uint32_t)(t / p.J);
    const uint32_t j = (uint32_t)(t % p.J);
    const size_t leaf_slot = (size_t)p.leaf_node[path] * p.J + j;
    double L;  // latest_ego_cost_ of the child
    if (p.leaf_has_heuristic[path]) {
      // update_from_heuristic, hypothesis_statistic.h:100-110 (+ set_heuristic_estimate :158-161)
      const double h = cost_mean(p.leaf_cost + (size_t)path * p.C, p.C);
      p.value[leaf_slot] = h;
      p.latest[leaf_slot] = h;
      p.visits_hyp[leaf_slot * (p.H + 1) + p.H] += 1;   // key HYPOTHESIS_ID_NOT_SET
      p.visits[leaf_slot] += 1;
      L = h;
    } else {
      L = p.latest[leaf_slot];
    }
    const uint32_t off = p.path_offset[path];
    const uint32_t len = p.path_len[path];
    for (uint32_t e = 0; e < len; ++e) {
      const size_t idx = (size_t)off + e;
      const size_t slot = (size_t)p.edge_node[idx] * p.J + j;
      const uint32_t hyp = p.edge_hyp[idx * p.J + j];
      const uint32_t act = p.edge_action[idx * p.J + j];
      // update_statistic, hypothesis_statistic.h:112-134
      const double cw = cost_mean(p.edge_cost + idx * p.C, p.C);
      L = p.chance ? (cw < L ? L : cw) : __dadd_rn(cw, __dmul_rn(p.gamma, L));
      p.latest[slot] = L;
      const size_t ia = (slot * p.H + hyp) * p.MA + act;
      const uint32_t cnt = p.a_count[ia] + 1;
      p.a_count[ia] = cnt;
      const double ac = p.a_cost[ia];
      p.a_cost[ia] = __dadd_rn(ac, __ddiv_rn(__dsub_rn(L, ac), (double)cnt));
      const size_t ih = slot * (p.H + 1) + hyp;
      const uint32_t nh = p.visits_hyp[ih] + 1;
      p.visits_hyp[ih] = nh;
      const double v = p.value[slot];
      p.value[slot] = __dadd_rn(v, __ddiv_rn(__dsub_rn(L, v), (double)nh));
      p.visits[slot] += 1;
      p.expanded[slot] += 1;
    }
  }
}

}  // namespace mamcts

using namespace mamcts;

extern "C" {

int mamcts_backup_uct(mamcts_engine* e, const mamcts_params* mp, uint32_t num_agents,
                      uint32_t n_paths, int32_t mem, const uint32_t* leaf_node,
                      const uint8_t* leaf_has_heuristic, const double* ego_return,
                      const double* other_return, const uint32_t* path_offset,
                      const uint32_t* path_len, const uint32_t* edge_node,
                      const uint8_t* edge_action, const double* edge_reward,
                      mamcts_uct_stats* stats) {
  if (!e) return fail(nullptr, MAMCTS_ERR_INVALID, "mamcts_backup_uct: null engine");
  std::lock_guard<std::mutex> lock(e->mu);
  if (!mp || !stats || !leaf_node || !leaf_has_heuristic || !ego_return || !path_offset ||
      !path_len || (num_agents > 1 && !other_return))
    return fail(e, MAMCTS_ERR_INVALID, "mamcts_backup_uct: null argument");
  if (num_agents == 0 || num_agents > MAMCTS_MAX_AGENTS || stats->max_actions == 0 ||
      stats->max_actions > MAMCTS_MAX_ACTIONS || stats->num_slots % num_agents != 0)
    return fail(e, MAMCTS_ERR_INVALID, "mamcts_backup_uct: bad agent / action / slot counts");
  if (n_paths == 0) return MAMCTS_OK;
  MAMCTS_CUDA(e, cudaSetDevice(e->device));
  // total number of edges: needs the host view of offsets/lengths when they are host arrays;
  // for device arrays the caller guarantees edge arrays cover max(offset+len).
  size_t n_edges = 0;
  if (mem == MAMCTS_MEM_HOST) {
    for (uint32_t i = 0; i < n_paths; ++i) n_edges = std::max<size_t>(n_edges, (size_t)path_offset[i] + path_len[i]);
    for (uint32_t i = 0; i < n_paths; ++i)
      if ((size_t)leaf_node[i] * num_agents >= stats->num_slots)
        return fail(e, MAMCTS_ERR_INVALID, "mamcts_backup_uct: leaf node out of range");
  }
  if (mem == MAMCTS_MEM_HOST && n_edges > 0 && (!edge_node || !edge_action || !edge_reward))
    return fail(e, MAMCTS_ERR_INVALID, "mamcts_backup_uct: null edge arrays");
  if (mem == MAMCTS_MEM_HOST) {
    for (size_t i = 0; i < n_edges; ++i) {
      if ((size_t)edge_node[i] * num_agents >= stats->num_slots)
        return fail(e, MAMCTS_ERR_INVALID, "mamcts_backup_uct: edge node out of range");
      for (uint32_t a = 0; a < num_agents; ++a)
        if (edge_action[i * num_agents + a] >= stats->max_actions)
          return fail(e, MAMCTS_ERR_INVALID, "mamcts_backup_uct: action index out of range");
    }
  }
  const size_t A = num_agents, S = stats->num_slots, MA = stats->max_actions;
  ArgStager stg(e);
  const int t_leaf = stg.plan(leaf_node, sizeof(uint32_t) * n_paths, mem, false);
  const int t_has = stg.plan(leaf_has_heuristic, n_paths, mem, false);
  const int t_ego = stg.plan(ego_return, sizeof(double) * n_paths, mem, false);
  const int t_oth = stg.plan(other_return, sizeof(double) * (A - 1) * n_paths, mem, false);
  const int t_off = stg.plan(path_offset, sizeof(uint32_t) * n_paths, mem, false);
  const int t_len = stg.plan(path_len, sizeof(uint32_t) * n_paths, mem, false);
  const int t_en = stg.plan(n_edges || mem == MAMCTS_MEM_DEVICE ? edge_node : nullptr, sizeof(uint32_t) * n_edges, mem, false);
  const int t_ea = stg.plan(n_edges || mem == MAMCTS_MEM_DEVICE ? edge_action : nullptr, n_edges * A, mem, false);
  const int t_er = stg.plan(n_edges || mem == MAMCTS_MEM_DEVICE ? edge_reward : nullptr, sizeof(double) * n_edges * A, mem, false);
  // statistics are in/out: staged in as inputs, copied back explicitly below
  const int s_val = stg.plan(stats->value, sizeof(double) * S, stats->mem, false);
  const int s_lat = stg.plan(stats->latest_return, sizeof(double) * S, stats->mem, false);
  const int s_vis = stg.plan(stats->total_visits, sizeof(uint32_t) * S, stats->mem, false);
  const int s_q = stg.plan(stats->q, sizeof(double) * S * MA, stats->mem, false);
  const int s_n = stg.plan(stats->n, sizeof(uint32_t) * S * MA, stats->mem, false);
  int rc = stg.commit();
  if (rc != MAMCTS_OK) return rc;

  BackupParams p;
  p.n_paths = n_paths; p.A = num_agents; p.max_actions = stats->max_actions;
  p.gamma = mp->discount_factor;
  p.leaf_node = stg.dev<const uint32_t>(t_leaf);
  p.leaf_has_heuristic = stg.dev<const uint8_t>(t_has);
  p.ego_return = stg.dev<const double>(t_ego);
  p.other_return = stg.dev<const double>(t_oth);
  p.path_offset = stg.dev<const uint32_t>(t_off);
  p.path_len = stg.dev<const uint32_t>(t_len);
  p.edge_node = stg.dev<const uint32_t>(t_en);
  p.edge_action = stg.dev<const uint8_t>(t_ea);
  p.edge_reward = stg.dev<const double>(t_er);
  p.value = stg.dev<double>(s_val);
  p.latest = stg.dev<double>(s_lat);
  p.visits = stg.dev<uint32_t>(s_vis);
  p.q = stg.dev<double>(s_q);
  p.n = stg.dev<uint32_t>(s_n);
  const uint64_t total = (uint64_t)n_paths * num_agents;
  const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((total + 255) / 256, (uint64_t)e->num_sms * 8));
  backup_uct_kernel<<<grid, 256, 0, e->stream>>>(p);
  e->launches++;
  MAMCTS_CUDA(e, cudaGetLastError());
  if (stats->mem == MAMCTS_MEM_HOST) {
    MAMCTS_CUDA(e, cudaMemcpyAsync(stats->value, p.value, sizeof(double) * S, cudaMemcpyDeviceToHost, e->stream));
    MAMCTS_CUDA(e, cudaMemcpyAsync(stats->latest_return, p.latest, sizeof(double) * S, cudaMemcpyDeviceToHost, e->stream));
    MAMCTS_CUDA(e, cudaMemcpyAsync(stats->total_visits, p.visits, sizeof(uint32_t) * S, cudaMemcpyDeviceToHost, e->stream));
    MAMCTS_CUDA(e, cudaMemcpyAsync(stats->q, p.q, sizeof(double) * S * MA, cudaMemcpyDeviceToHost, e->stream));
    MAMCTS_CUDA(e, cudaMemcpyAsync(stats->n, p.n, sizeof(uint32_t) * S * MA, cudaMemcpyDeviceToHost, e->stream));
  }
  return stg.finish();
}

int mamcts_backup_hypothesis(mamcts_engine* e, const mamcts_params* mp, uint32_t num_others,
                             uint32_t n_paths, int32_t mem, const uint32_t* leaf_node,
                             const uint8_t* leaf_has_heuristic, const double* leaf_cost,
                             uint32_t num_costs, const uint32_t* path_offset, const uint32_t* path_len,
                             const uint32_t* edge_node, const uint8_t* edge_hypothesis,
                             const uint8_t* edge_action, const double* edge_cost,
                             int32_t chance_constrained_update, mamcts_hyp_stats* stats) {
  if (!e) return fail(nullptr, MAMCTS_ERR_INVALID, "mamcts_backup_hypothesis: null engine");
  std::lock_guard<std::mutex> lock(e->mu);
  if (!mp || !stats || !leaf_node || !leaf_has_heuristic || !leaf_cost || !path_offset || !path_len)
    return fail(e, MAMCTS_ERR_INVALID, "mamcts_backup_hypothesis: null argument");
  if (num_others == 0 || num_others >= MAMCTS_MAX_AGENTS || num_costs == 0 || stats->num_hypotheses == 0 ||
      stats->num_hypotheses > 255 || stats->max_actions == 0 || stats->max_actions > 255 ||
      stats->num_slots % num_others != 0)
    return fail(e, MAMCTS_ERR_INVALID, "mamcts_backup_hypothesis: bad agent / hypothesis / action / slot counts");
  if (n_paths == 0) return MAMCTS_OK;
  MAMCTS_CUDA(e, cudaSetDevice(e->device));
  const size_t J = num_others, S = stats->num_slots, H = stats->num_hypotheses, MA = stats->max_actions, C = num_costs;
  size_t n_edges = 0;
  if (mem == MAMCTS_MEM_HOST) {
    for (uint32_t i = 0; i < n_paths; ++i) n_edges = std::max<size_t>(n_edges, (size_t)path_offset[i] + path_len[i]);
    for (uint32_t i = 0; i < n_paths; ++i)
      if ((size_t)leaf_node[i] * J >= S) return fail(e, MAMCTS_ERR_INVALID, "mamcts_backup_hypothesis: leaf node out of range");
    if (n_edges > 0 && (!edge_node || !edge_hypothesis || !edge_action || !edge_cost))
      return fail(e, MAMCTS_ERR_INVALID, "mamcts_backup_hypothesis: null edge arrays");
    for (size_t i = 0; i < n_edges; ++i) {
      if ((size_t)edge_node[i] * J >= S) return fail(e, MAMCTS_ERR_INVALID, "mamcts_backup_hypothesis: edge node out of range");
      for (size_t j = 0; j < J; ++j)
        if (edge_hypothesis[i * J + j] >= H || edge_action[i * J + j] >= MA)
          return fail(e, MAMCTS_ERR_INVALID, "mamcts_backup_hypothesis: hypothesis / action index out of range");
    }
  }
  const bool edges = n_edges || mem == MAMCTS_MEM_DEVICE;
  ArgStager stg(e);
  const int t_leaf = stg.plan(leaf_node, sizeof(uint32_t) * n_paths, mem, false);
  const int t_has = stg.plan(leaf_has_heuristic, n_paths, mem, false);
  const int t_lc = stg.plan(leaf_cost, sizeof(double) * n_paths * C, mem, false);
  const int t_off = stg.plan(path_offset, sizeof(uint32_t) * n_paths, mem, false);
  const int t_len = stg.plan(path_len, sizeof(uint32_t) * n_paths, mem, false);
  const int t_en = stg.plan(edges ? edge_node : nullptr, sizeof(uint32_t) * n_edges, mem, false);
  const int t_eh = stg.plan(edges ? edge_hypothesis : nullptr, n_edges * J, mem, false);
  const int t_ea = stg.plan(edges ? edge_action : nullptr, n_edges * J, mem, false);
  const int t_ec = stg.plan(edges ? edge_cost : nullptr, sizeof(double) * n_edges * C, mem, false);
  // statistics are in/out: staged in as inputs, copied back explicitly below
  const int s_val = stg.plan(stats->ego_cost_value, sizeof(double) * S, stats->mem, false);
  const int s_lat = stg.plan(stats->latest_ego_cost, sizeof(double) * S, stats->mem, false);
  const int s_vis = stg.plan(stats->total_visits, sizeof(uint32_t) * S, stats->mem, false);
  const int s_exp = stg.plan(stats->num_expanded, sizeof(uint32_t) * S, stats->mem, false);
  const int s_vh = stg.plan(stats->visits_hypothesis, sizeof(uint32_t) * S * (H + 1), stats->mem, false);
  const int s_ac = stg.plan(stats->action_cost, sizeof(double) * S * H * MA, stats->mem, false);
  const int s_an = stg.plan(stats->action_count, sizeof(uint32_t) * S * H * MA, stats->mem, false);
  int rc = stg.commit();
  if (rc != MAMCTS_OK) return rc;
  HypBackupParams p;
  p.n_paths = n_paths; p.J = num_others; p.H = stats->num_hypotheses; p.MA = stats->max_actions; p.C = num_costs;
  p.chance = chance_constrained_update;
  p.gamma = mp->discount_factor;
  p.leaf_node = stg.dev<const uint32_t>(t_leaf);
  p.leaf_has_heuristic = stg.dev<const uint8_t>(t_has);
  p.leaf_cost = stg.dev<const double>(t_lc);
  p.path_offset = stg.dev<const uint32_t>(t_off);
  p.path_len = stg.dev<const uint32_t>(t_len);
  p.edge_node = stg.dev<const uint32_t>(t_en);
  p.edge_hyp = stg.dev<const uint8_t>(t_eh);
  p.edge_action = stg.dev<const uint8_t>(t_ea);
  p.edge_cost = stg.dev<const double>(t_ec);
  p.value = stg.dev<double>(s_val);
  p.latest = stg.dev<double>(s_lat);
  p.visits = stg.dev<uint32_t>(s_vis);
  p.expanded = stg.dev<uint32_t>(s_exp);
  p.visits_hyp = stg.dev<uint32_t>(s_vh);
  p.a_cost = stg.dev<double>(s_ac);
  p.a_count = stg.dev<uint32_t>(s_an);
  const uint64_t total = (uint64_t)n_paths * num_others;
  const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((total + 255) / 256, (uint64_t)e->num_sms * 8));
  backup_hypothesis_kernel<<<grid, 256, 0, e->stream>>>(p);
  e->launches++;
  MAMCTS_CUDA(e, cudaGetLastError());
  if (stats->mem == MAMCTS_MEM_HOST) {
    MAMCTS_CUDA(e, cudaMemcpyAsync(stats->ego_cost_value, p.value, sizeof(double) * S, cudaMemcpyDeviceToHost, e->stream));
    MAMCTS_CUDA(e, cudaMemcpyAsync(stats->latest_ego_cost, p.latest, sizeof(double) * S, cudaMemcpyDeviceToHost, e->stream));
    MAMCTS_CUDA(e, cudaMemcpyAsync(stats->total_visits, p.visits, sizeof(uint32_t) * S, cudaMemcpyDeviceToHost, e->stream));
    MAMCTS_CUDA(e, cudaMemcpyAsync(stats->num_expanded, p.expanded, sizeof(uint32_t) * S, cudaMemcpyDeviceToHost, e->stream));
    MAMCTS_CUDA(e, cudaMemcpyAsync(stats->visits_hypothesis, p.visits_hyp, sizeof(uint32_t) * S * (H + 1), cudaMemcpyDeviceToHost, e->stream));
    MAMCTS_CUDA(e, cudaMemcpyAsync(stats->action_cost, p.a_cost, sizeof(double) * S * H * MA, cudaMemcpyDeviceToHost, e->stream));
    MAMCTS_CUDA(e, cudaMemcpyAsync(stats->action_count, p.a_count, sizeof(uint32_t) * S * H * MA, cudaMemcpyDeviceToHost, e->stream));
  }
  return stg.finish();
}

}  // extern "C"
