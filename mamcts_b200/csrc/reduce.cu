// reduce.cu - the two reductions that precede the path's only collective:
//
//   root_merge_kernel  the local part of UctStatistic::merge_node_statistics (reference
//     mcts/statistics/uct_statistic.h:165-184 via mcts/mcts.h:270-292): every thread reads ONE tree's ego
//     root statistic once and carries {sum_q[a], sum_n[a], occ[a]} + sum_visits through a warp-shuffle
//     butterfly, a per-block shared-memory step and a last-block pass over the block partials.
//   moments_kernel     C3: {sum, sum of squares} of the per-rollout returns of every agent.
//
// Both are deterministic: the grid is a function of the element count only, every partial is formed in a
// fixed order (strided per-thread sums -> xor butterfly -> warps in index order -> blocks in index order by
// the block that takes the last ticket), so the result does not depend on scheduling or on the number of
// GPUs a caller later all-reduces over. One launch per decision; with a communicator attached the
// ncclAllReduce of the 3 nA + 1 (or 2 A + 1) doubles follows and a one-thread kernel finishes.
#include <algorithm>
#include <cstring>

#include "engine.h"
#include "root_merge.h"
#include "staging.h"

namespace mamcts {

constexpr int kReduceThreads = 256;
constexpr int kReduceWarps = kReduceThreads / 32;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xFFFFFFFFu, v, off));
  return v;
}

// The block that finishes last adds the block partials (index order: lane-strided, then the butterfly)
// into acc[0..ncomp). Returns true in that block, after acc is written.
__device__ __forceinline__ bool last_block_reduce(const double* sm_block /* [ncomp] */, uint32_t ncomp,
                                                  double* partials, unsigned int* ticket, double* acc) {
  __shared__ bool s_last;
  for (uint32_t c = threadIdx.x; c < ncomp; c += blockDim.x) partials[(size_t)blockIdx.x * ncomp + c] = sm_block[c];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  for (uint32_t c = warp; c < ncomp; c += kReduceWarps) {
    double part = 0.0;
    for (uint32_t b = lane; b < gridDim.x; b += 32) part = __dadd_rn(part, __ldcg(partials + (size_t)b * ncomp + c));
    part = warp_sum(part);
    if (lane == 0) acc[c] = part;
  }
  if (threadIdx.x == 0) *ticket = 0;   // ready for the next launch on this stream
  __syncthreads();
  return true;
}

// merge_node_statistics' division and mean (uct_statistic.h:173-183) + get_best_action (:78-89).
// out (doubles): [0,nA) merged q, [nA,2nA) merged n, [2nA,3nA) occurrences, [3nA] visits,
// [3nA+1] value, [3nA+2] best action.
__device__ void root_finish(const double* acc, uint32_t nA, double* out) {
  double value = 0.0;
  uint32_t present = 0;
  double best_q = 0.0;
  uint32_t best = 0;
  bool first = true;
  for (uint32_t a = 0; a < nA; ++a) {
    const double occ = acc[2 * nA + a];
    double q = 0.0;
    if (occ > 0.0) {
      q = acc[a] / occ;
      value = __dadd_rn(value, q);
      present += 1;
      if (first) { best_q = q; best = a; first = false; }
      else if (q > best_q) { best_q = q; best = a; }
    }
    out[a] = q;
    out[nA + a] = acc[nA + a];
    out[2 * nA + a] = occ;
  }
  out[3 * nA] = acc[3 * nA];
  out[3 * nA + 1] = present ? value / (double)present : 0.0;
  out[3 * nA + 2] = (double)best;
}

__global__ void root_finish_kernel(const double* __restrict__ acc, uint32_t nA, double* __restrict__ out) {
  if (blockIdx.x == 0 && threadIdx.x == 0) root_finish(acc, nA, out);
}

__global__ void __launch_bounds__(kReduceThreads)
root_merge_kernel(RootGather g, uint32_t n_trees, uint32_t nA, double* partials, unsigned int* ticket,
                  double* acc, double* out /* null: a collective follows, root_finish_kernel finishes */) {
  __shared__ double sm[kReduceWarps][3 * MAMCTS_MAX_ACTIONS + 1];
  __shared__ double sm_block[3 * MAMCTS_MAX_ACTIONS + 1];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = t < n_trees;
  size_t qbase = 0, nbase = 0, vbase = 0;
  if (live) {
    if (g.root_slot) {
      const uint32_t slot = g.root_slot[t];
      qbase = nbase = (size_t)slot * g.max_actions;
      vbase = slot;
    } else {
      qbase = (size_t)t * g.stride_q;
      nbase = (size_t)t * g.stride_n;
      vbase = (size_t)t * g.stride_v;
    }
  }
  auto count_at = [&](const void* ptr, size_t idx) -> uint32_t {
    return g.count_bytes == 2 ? (uint32_t)static_cast<const uint16_t*>(ptr)[idx]
                              : static_cast<const uint32_t*>(ptr)[idx];
  };
  // one pass over the tree's root record: an action is present in a tree iff it was expanded there, and an
  // expanded action has n > 0 at iteration boundaries
  for (uint32_t a = 0; a < nA; ++a) {
    const uint32_t n = live ? count_at(g.n, nbase + a) : 0u;
    const double q = n ? g.q[qbase + a] : 0.0;
    const double sq = warp_sum(q), sn = warp_sum((double)n), so = warp_sum(n ? 1.0 : 0.0);
    if (lane == 0) { sm[warp][a] = sq; sm[warp][nA + a] = sn; sm[warp][2 * nA + a] = so; }
  }
  {
    const double sv = warp_sum(live ? (double)count_at(g.visits, vbase) : 0.0);
    if (lane == 0) sm[warp][3 * nA] = sv;
  }
  __syncthreads();
  const uint32_t ncomp = 3 * nA + 1;
  for (uint32_t c = threadIdx.x; c < ncomp; c += blockDim.x) {
    double v = sm[0][c];
#pragma unroll
    for (int w = 1; w < kReduceWarps; ++w) v = __dadd_rn(v, sm[w][c]);
    sm_block[c] = v;
  }
  __syncthreads();
  if (last_block_reduce(sm_block, ncomp, partials, ticket, acc) && out && threadIdx.x == 0) root_finish(acc, nA, out);
}

// values of row r: r == 0 ? ego : other + (r - 1) * n. acc: per row {sum, sum of squares}, then the count.
__global__ void __launch_bounds__(kReduceThreads)
moments_kernel(const double* __restrict__ ego, const double* __restrict__ other, uint32_t rows, uint32_t n,
               uint32_t per_block, double* partials, unsigned int* ticket, double* acc) {
  __shared__ double sm[kReduceWarps][2 * MAMCTS_MAX_AGENTS + 1];
  __shared__ double sm_block[2 * MAMCTS_MAX_AGENTS + 1];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t begin = blockIdx.x * per_block;
  const uint32_t end = min(n, begin + per_block);
  for (uint32_t r = 0; r < rows; ++r) {
    const double* v = r == 0 ? ego : other + (size_t)(r - 1) * n;
    double s = 0.0, ss = 0.0;
    for (uint32_t i = begin + threadIdx.x; i < end; i += blockDim.x) {
      const double x = __ldcs(v + i);
      s = __dadd_rn(s, x);
      ss = __dadd_rn(ss, __dmul_rn(x, x));
    }
    s = warp_sum(s);
    ss = warp_sum(ss);
    if (lane == 0) { sm[warp][2 * r] = s; sm[warp][2 * r + 1] = ss; }
  }
  if (lane == 0) sm[warp][2 * rows] = 0.0;
  if (threadIdx.x == 0) sm[0][2 * rows] = (double)(end > begin ? end - begin : 0u);
  __syncthreads();
  const uint32_t ncomp = 2 * rows + 1;
  for (uint32_t c = threadIdx.x; c < ncomp; c += blockDim.x) {
    double x = sm[0][c];
#pragma unroll
    for (int w = 1; w < kReduceWarps; ++w) x = __dadd_rn(x, sm[w][c]);
    sm_block[c] = x;
  }
  __syncthreads();
  last_block_reduce(sm_block, ncomp, partials, ticket, acc);
}

// scratch layout of both reductions: [ticket (256 B)] [acc] [out] [partials]
struct ReduceScratch {
  unsigned int* ticket;
  double *acc, *out, *partials;
};

static int reduce_scratch(mamcts_engine* e, size_t nacc, size_t nout, size_t blocks, ReduceScratch* rs) {
  const size_t bytes = 256 + sizeof(double) * (nacc + nout + blocks * nacc);
  int rc = ensure_reduce_buf(e, bytes);
  if (rc != MAMCTS_OK) return rc;
  char* base = static_cast<char*>(e->reduce_buf);
  rs->ticket = reinterpret_cast<unsigned int*>(base);
  rs->acc = reinterpret_cast<double*>(base + 256);
  rs->out = rs->acc + nacc;
  rs->partials = rs->out + nout;
  MAMCTS_CUDA(e, cudaMemsetAsync(rs->ticket, 0, sizeof(unsigned int), e->stream));
  return MAMCTS_OK;
}

// Shared by mamcts_root_merge and mamcts_search_merge_roots (search.cu).
int root_merge_device(mamcts_engine* e, const RootGather& g, uint32_t n_trees, uint32_t nA,
                      double* merged_q, uint64_t* merged_n, uint32_t* merged_occ,
                      uint64_t* merged_total_visits, double* merged_value, uint32_t* best_action) {
  if (nA == 0 || nA > MAMCTS_MAX_ACTIONS) return fail(e, MAMCTS_ERR_INVALID, "root merge: bad num_actions");
  const size_t nacc = 3 * (size_t)nA + 1, nout = 3 * (size_t)nA + 3;
  const uint32_t blocks = (n_trees + kReduceThreads - 1) / kReduceThreads;
  ReduceScratch rs;
  int rc = reduce_scratch(e, nacc, nout, blocks, &rs);
  if (rc != MAMCTS_OK) return rc;
  const bool collective = e->nccl_comm != nullptr;
  root_merge_kernel<<<blocks, kReduceThreads, 0, e->stream>>>(g, n_trees, nA, rs.partials, rs.ticket, rs.acc,
                                                              collective ? nullptr : rs.out);
  e->launches++;
  MAMCTS_CUDA(e, cudaGetLastError());
  if (collective) {
    rc = mamcts_allreduce_f64(e, rs.acc, nacc);  // one all-reduce per decision
    if (rc != MAMCTS_OK) return rc;
    root_finish_kernel<<<1, 32, 0, e->stream>>>(rs.acc, nA, rs.out);
    e->launches++;
    MAMCTS_CUDA(e, cudaGetLastError());
  }
  double host[3 * MAMCTS_MAX_ACTIONS + 3];
  MAMCTS_CUDA(e, cudaMemcpyAsync(host, rs.out, sizeof(double) * nout, cudaMemcpyDeviceToHost, e->stream));
  MAMCTS_CUDA(e, cudaStreamSynchronize(e->stream));
  for (uint32_t a = 0; a < nA; ++a) {
    if (merged_q) merged_q[a] = host[a];
    if (merged_n) merged_n[a] = (uint64_t)host[nA + a];
    if (merged_occ) merged_occ[a] = (uint32_t)host[2 * nA + a];
  }
  if (merged_total_visits) *merged_total_visits = (uint64_t)host[3 * nA];
  if (merged_value) *merged_value = host[3 * nA + 1];
  if (best_action) *best_action = (uint32_t)host[3 * nA + 2];
  return MAMCTS_OK;
}

int rollout_moments_device(mamcts_engine* e, const double* d_ego, const double* d_other, uint32_t num_others,
                           uint32_t n, double* d_dst) {
  const uint32_t rows = num_others + 1;
  const size_t nacc = 2 * (size_t)rows + 1;
  // 4096 values per block and row, at most 1024 blocks: the geometry is a function of n alone
  uint32_t per_block = 4096;
  while ((uint64_t)per_block * 1024u < n) per_block *= 2;
  const uint32_t blocks = std::max(1u, (n + per_block - 1) / per_block);
  ReduceScratch rs;
  int rc = reduce_scratch(e, nacc, 0, blocks, &rs);
  if (rc != MAMCTS_OK) return rc;
  moments_kernel<<<blocks, kReduceThreads, 0, e->stream>>>(d_ego, d_other, rows, n, per_block, rs.partials, rs.ticket, rs.acc);
  e->launches++;
  MAMCTS_CUDA(e, cudaGetLastError());
  rc = mamcts_allreduce_f64(e, rs.acc, nacc);   // no communicator: the local sums
  if (rc != MAMCTS_OK) return rc;
  MAMCTS_CUDA(e, cudaMemcpyAsync(d_dst, rs.acc, sizeof(double) * nacc, cudaMemcpyDeviceToDevice, e->stream));
  return MAMCTS_OK;
}

}  // namespace mamcts

using namespace mamcts;

extern "C" {

int mamcts_root_merge(mamcts_engine* e, const mamcts_uct_stats* stats, uint32_t n_trees,
                      int32_t mem, const uint32_t* root_slot, uint32_t num_actions,
                      double* merged_q, uint64_t* merged_n, uint32_t* merged_occurrences,
                      uint64_t* merged_total_visits, double* merged_value,
                      uint32_t* best_action) {
  if (!e) return fail(nullptr, MAMCTS_ERR_INVALID, "mamcts_root_merge: null engine");
  std::lock_guard<std::mutex> lock(e->mu);
  if (!stats || !root_slot || n_trees == 0) return fail(e, MAMCTS_ERR_INVALID, "mamcts_root_merge: null argument");
  if (num_actions > stats->max_actions) return fail(e, MAMCTS_ERR_INVALID, "mamcts_root_merge: num_actions > max_actions");
  if (mem == MAMCTS_MEM_HOST) {
    for (uint32_t t = 0; t < n_trees; ++t)
      if (root_slot[t] >= stats->num_slots) return fail(e, MAMCTS_ERR_INVALID, "mamcts_root_merge: root slot out of range");
  }
  MAMCTS_CUDA(e, cudaSetDevice(e->device));
  const size_t S = stats->num_slots, MA = stats->max_actions;
  ArgStager stg(e);
  const int t_slot = stg.plan(root_slot, sizeof(uint32_t) * n_trees, mem, false);
  const int s_vis = stg.plan(stats->total_visits, sizeof(uint32_t) * S, stats->mem, false);
  const int s_q = stg.plan(stats->q, sizeof(double) * S * MA, stats->mem, false);
  const int s_n = stg.plan(stats->n, sizeof(uint32_t) * S * MA, stats->mem, false);
  int rc = stg.commit();
  if (rc != MAMCTS_OK) return rc;
  RootGather g;
  std::memset(&g, 0, sizeof(g));
  g.q = stg.dev<const double>(s_q);
  g.n = stg.dev<const uint32_t>(s_n);
  g.visits = stg.dev<const uint32_t>(s_vis);
  g.count_bytes = 4;
  g.root_slot = stg.dev<const uint32_t>(t_slot);
  g.max_actions = stats->max_actions;
  return root_merge_device(e, g, n_trees, num_actions, merged_q, merged_n, merged_occurrences,
                           merged_total_visits, merged_value, best_action);
}

int mamcts_rollout_moments(mamcts_engine* e, int32_t mem, const double* ego_return, const double* other_return,
                           uint32_t num_others, uint32_t n, double* moments) {
  if (!e) return fail(nullptr, MAMCTS_ERR_INVALID, "mamcts_rollout_moments: null engine");
  std::lock_guard<std::mutex> lock(e->mu);
  if (!ego_return || !moments || (num_others && !other_return))
    return fail(e, MAMCTS_ERR_INVALID, "mamcts_rollout_moments: null argument");
  if (num_others >= MAMCTS_MAX_AGENTS) return fail(e, MAMCTS_ERR_INVALID, "mamcts_rollout_moments: too many agents");
  MAMCTS_CUDA(e, cudaSetDevice(e->device));
  ArgStager stg(e);
  const int t_ego = stg.plan(ego_return, sizeof(double) * n, mem, false);
  const int t_oth = stg.plan(num_others ? other_return : nullptr, sizeof(double) * (size_t)num_others * n, mem, false);
  const int t_out = stg.plan(moments, sizeof(double) * (2 * ((size_t)num_others + 1) + 1), mem, true);
  int rc = stg.commit();
  if (rc != MAMCTS_OK) return rc;
  rc = rollout_moments_device(e, stg.dev<const double>(t_ego), stg.dev<const double>(t_oth), num_others, n,
                              stg.dev<double>(t_out));
  if (rc != MAMCTS_OK) return rc;
  return stg.finish();
}

}  // extern "C"
