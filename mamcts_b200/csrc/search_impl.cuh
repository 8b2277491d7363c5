// search_impl.cuh - the concrete searches behind mamcts_search (search_host.h): one class template per tree layout
// (SearchT classic, SearchC compact two-agent, SearchCM compact multi-agent), instantiated per (environment, action
// counts, index widths) by search.cu and search_more.cu (the translation units split the instantiations so that
// they compile side by side).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "engine.h"
#include "root_merge.h"
#include "search_async.cuh"
#include "search_compact.cuh"
#include "search_compact_multi.cuh"
#include "search_host.h"
#include "search_kernel.cuh"

namespace mamcts {

// canonical export of a classic-layout tree (search.cu)
int dump_classic(mamcts_search* s, uint32_t tree, double* records, uint32_t capacity_records,
                 uint32_t* num_records);

template <class Env, int NE, int NA, int NCHILD, class CIDX>
struct SearchT final : public mamcts_search {
  int dump(uint32_t tree, double* records, uint32_t capacity_records, uint32_t* num_records) override {
    return dump_classic(this, tree, records, capacity_records, num_records);
  }
  using Node = NodeRec<Env, NE, NA, NCHILD, CIDX>;

  size_t node_bytes() const override { return sizeof(Node); }
  bool uses_hash() const override { return NCHILD == 0; }

  int launch_reset() override {
    const uint32_t T = cfg.num_trees;
    if (NCHILD == 0)
      MAMCTS_CUDA(e, cudaMemsetAsync(d_hash, 0, sizeof(uint32_t) * (size_t)T * (p.hash_mask + 1), e->stream));
    MAMCTS_CUDA(e, cudaMemsetAsync(d_counters, 0, sizeof(unsigned long long) * 8, e->stream));
    search_reset_kernel<Env, NE, NA, NCHILD, CIDX><<<(T + 255) / 256, 256, 0, e->stream>>>(
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
    const uint32_t trees_per_block = (uint32_t)(block_threads / 32) * q.lanes_per_warp;
    const int grid = (int)((T + trees_per_block - 1) / trees_per_block);
    if (cfg.action_source == MAMCTS_SRC_PHILOX && cfg.rollouts_per_leaf > 1)
      search_kernel<Env, NE, NA, NCHILD, CIDX, MAMCTS_SRC_PHILOX, true><<<grid, block_threads, 0, e->stream>>>(q);
    else if (cfg.action_source == MAMCTS_SRC_PHILOX)
      search_kernel<Env, NE, NA, NCHILD, CIDX, MAMCTS_SRC_PHILOX><<<grid, block_threads, 0, e->stream>>>(q);
    else
      search_kernel<Env, NE, NA, NCHILD, CIDX, MAMCTS_SRC_REFERENCE_CONST><<<grid, block_threads, 0, e->stream>>>(q);
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
    g->n = base + offsetof(Node, ego) + offsetof(StatRec<NE>, n);
    g->visits = base + offsetof(Node, ego) + offsetof(StatRec<NE>, N);
    g->count_bytes = 4;
    g->root_slot = nullptr;
    g->stride_q = tree_bytes / sizeof(double);
    g->stride_n = tree_bytes / sizeof(uint32_t);
    g->stride_v = tree_bytes / sizeof(uint32_t);
    g->max_actions = NE;
  }
};

// Compact (two-tier) layout: search_compact.cuh.
static_assert(sizeof(InnerRec<CrossingEnv<1>, 4, 7, 28, uint16_t, uint16_t>) == 256 &&
                  sizeof(InnerRec<CrossingEnv<1>, 4, 7, 28, uint16_t, uint16_t>::Counters) == 32 &&
                  sizeof(InnerRec<CrossingEnv<1>, 4, 7, 28, uint16_t, uint32_t>::Counters) == 64,
              "C2 inner record: 256 bytes, counters in one sector (16-bit counts) or two (32-bit)");
template <class Env, int NE, int NA, int NCHILD, class CIDX, class CNT>
struct SearchC final : public mamcts_search {
  using Inner = InnerRec<Env, NE, NA, NCHILD, CIDX, CNT>;
  using Leaf = LeafRec<Env::A>;

  size_t node_bytes() const override { return sizeof(Inner); }
  size_t leaf_bytes() const override { return sizeof(Leaf); }
  bool uses_hash() const override { return false; }
  bool is_compact() const override { return true; }
  const char* root_record(uint32_t tree) const override {
    return static_cast<const char*>(d_inner) + sizeof(Inner) * (size_t)tree * max_inner;
  }

  int launch_reset() override {
    const uint32_t T = cfg.num_trees;
    MAMCTS_CUDA(e, cudaMemsetAsync(d_counters, 0, sizeof(unsigned long long) * 8, e->stream));
    search_compact_reset_kernel<Env, NE, NA, NCHILD, CIDX, CNT><<<(T + 255) / 256, 256, 0, e->stream>>>(
        d_inner, d_tree_nodes, d_tree_iters, d_tree_inner, d_tree_leaves, T, max_inner);
    e->launches++;
    MAMCTS_CUDA(e, cudaGetLastError());
    iters_done = 0;
    return MAMCTS_OK;
  }

  // Asynchronous tree-pool scheduling (search_async.cuh): the same trees, the same results, trees not bound to lanes.
  void* d_async_path = nullptr;
  double* d_async_path_r = nullptr;
  void free_extra() override { cudaFree(d_async_path); cudaFree(d_async_path_r); }
  static constexpr bool kAsyncBuilt = Env::A == 2 && Env::kOtherRewardZero && Env::kOutcomeRewards &&
                                      sizeof(CNT) == 2 && sizeof(CIDX) == 2;
  // slots per lane of the pool (0 = lane-per-tree kernel)
  int async_spl() const {
    if (!kAsyncBuilt) return 0;
    if (!(p.gamma >= 0.0) || cfg.rollouts_per_leaf != 1) return 0;
    unsigned long long bits;
    std::memcpy(&bits, &p.env.r_tab[0], sizeof(bits));
    if (bits != 0ull) return 0;   // a per-step reward: the rollout accumulates in FP64 (lane-per-tree kernel)
    // opt-in (profiling / comparison): the lane-per-tree kernel is faster at every size measured
    // (profiles/r02_search_async_pool.txt), MAMCTS_SEARCH_ASYNC = 2 or 4 tree slots per lane
    int spl = 0;
    if (const char* v = std::getenv("MAMCTS_SEARCH_ASYNC")) spl = std::atoi(v);
    return (spl == 2 || spl == 4) ? spl : 0;
  }
  template <int KIND, int SPL>
  int launch_async(const SearchParams& q, const CompactParams& c) {
    if constexpr (kAsyncBuilt) {
      using Pool = AsyncPool<32 * SPL, MAMCTS_ASYNC_SHARED_LEVELS, typename AsyncPathRec<CNT>::HeadWord,
                             typename AsyncPathRec<CNT>::CountPair>;
      const uint32_t T = cfg.num_trees;
      if (!d_async_path) {
        MAMCTS_CUDA(e, cudaMalloc(&d_async_path, sizeof(AsyncPathRec<CNT>) * (size_t)T * kPathMaxCompact));
        MAMCTS_CUDA(e, cudaMalloc(reinterpret_cast<void**>(&d_async_path_r), sizeof(double) * (size_t)T * kPathMaxCompact));
      }
      AsyncParams ap;
      ap.path = d_async_path; ap.path_r = d_async_path_r;
      ap.slots_per_warp = 32 * SPL;
      if (const char* v = std::getenv("MAMCTS_SEARCH_ASYNC_SPW")) {
        const int w = std::atoi(v);
        if (w >= 1 && w <= 32 * SPL) ap.slots_per_warp = (uint32_t)w;
      }
      const size_t smem = sizeof(Pool) * kAsyncWarps;
      auto kern = search_async_kernel<Env, NE, NA, NCHILD, CIDX, CNT, KIND, SPL>;
      MAMCTS_CUDA(e, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const uint32_t trees_per_block = ap.slots_per_warp * kAsyncWarps;
      const int grid = (int)((T + trees_per_block - 1) / trees_per_block);
      if (std::getenv("MAMCTS_SEARCH_DEBUG")) {
        int nb = 0;
        cudaFuncAttributes fa;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kAsyncThreads, smem);
        cudaFuncGetAttributes(&fa, kern);
        std::fprintf(stderr, "search_async_kernel<SPL=%d>: %d blocks/SM of %d threads, %d regs, %zu B dynamic smem, %zu B local, grid %d, trees/warp %u\n",
                     SPL, nb, kAsyncThreads, fa.numRegs, smem, fa.localSizeBytes, grid, ap.slots_per_warp);
      }
      kern<<<grid, kAsyncThreads, smem, e->stream>>>(q, c, ap);
      e->launches++;
      MAMCTS_CUDA(e, cudaGetLastError());
    }
    return MAMCTS_OK;
  }
  template <int KIND>
  int launch_async_spl(int spl, const SearchParams& q, const CompactParams& c) {
    return spl >= 4 ? launch_async<KIND, 4>(q, c) : launch_async<KIND, 2>(q, c);
  }

  int launch_run(uint32_t iterations) override {
    SearchParams q = p;
    q.iterations = iterations;
    CompactParams c;
    c.inner = d_inner; c.leaves = d_leaves; c.tree_inner = d_tree_inner; c.tree_leaves = d_tree_leaves;
    c.max_inner = max_inner; c.root_x = d_root_x;
    const uint32_t T = cfg.num_trees;
    if (const int spl = async_spl()) {
      return cfg.action_source == MAMCTS_SRC_PHILOX ? launch_async_spl<MAMCTS_SRC_PHILOX>(spl, q, c)
                                                    : launch_async_spl<MAMCTS_SRC_REFERENCE_CONST>(spl, q, c);
    }
    const uint32_t trees_per_block = (uint32_t)(block_threads / 32) * q.lanes_per_warp;
    const int grid = (int)((T + trees_per_block - 1) / trees_per_block);
    if (const char* v = std::getenv("MAMCTS_SEARCH_CARVEOUT")) {   // profiling knob: shared-memory carve-out in percent
      auto kern = search_compact_kernel<Env, NE, NA, NCHILD, CIDX, CNT, MAMCTS_SRC_PHILOX>;
      cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, std::atoi(v));
    }
    if (std::getenv("MAMCTS_SEARCH_DEBUG")) {   // resident blocks per SM as the runtime computes them
      int nb = 0;
      cudaFuncAttributes fa;
      auto kern = search_compact_kernel<Env, NE, NA, NCHILD, CIDX, CNT, MAMCTS_SRC_PHILOX>;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, block_threads, 0);
      cudaFuncGetAttributes(&fa, kern);
      std::fprintf(stderr, "search_compact_kernel: %d blocks/SM of %d threads, %d regs, %zu B smem, %zu B local, grid %d, lanes/warp %u\n",
                   nb, block_threads, fa.numRegs, fa.sharedSizeBytes, fa.localSizeBytes, grid, q.lanes_per_warp);
    }
    // a reward-free other agent with a discount >= 0: its values are +0.0 for the whole search (search_compact.cuh)
    const bool ovz = Env::kOtherRewardZero && p.gamma >= 0.0;
    if constexpr (Env::kOtherRewardZero) {
      if (ovz) {
        // few trees (at most 4 per warp): 8 lanes per tree, search_compact.cuh HELP
        if constexpr (Env::A == 2 && Env::kOutcomeRewards) {
          unsigned long long r0_bits;
          std::memcpy(&r0_bits, &p.env.r_tab[0], sizeof(r0_bits));
          const char* knob = std::getenv("MAMCTS_SEARCH_HELPERS");
          if (q.lanes_per_warp <= 4 && cfg.action_source == MAMCTS_SRC_PHILOX && cfg.rollouts_per_leaf == 1 &&
              r0_bits == 0ull && !(knob && std::atoi(knob) == 0)) {
            search_compact_kernel<Env, NE, NA, NCHILD, CIDX, CNT, MAMCTS_SRC_PHILOX, false, true, true>
                <<<grid, block_threads, 0, e->stream>>>(q, c);
            e->launches++;
            MAMCTS_CUDA(e, cudaGetLastError());
            return MAMCTS_OK;
          }
        }
        if (cfg.action_source == MAMCTS_SRC_PHILOX && cfg.rollouts_per_leaf > 1)
          search_compact_kernel<Env, NE, NA, NCHILD, CIDX, CNT, MAMCTS_SRC_PHILOX, true, true><<<grid, block_threads, 0, e->stream>>>(q, c);
        else if (cfg.action_source == MAMCTS_SRC_PHILOX)
          search_compact_kernel<Env, NE, NA, NCHILD, CIDX, CNT, MAMCTS_SRC_PHILOX, false, true><<<grid, block_threads, 0, e->stream>>>(q, c);
        else
          search_compact_kernel<Env, NE, NA, NCHILD, CIDX, CNT, MAMCTS_SRC_REFERENCE_CONST, false, true><<<grid, block_threads, 0, e->stream>>>(q, c);
        e->launches++;
        MAMCTS_CUDA(e, cudaGetLastError());
        return MAMCTS_OK;
      }
    }
    if (cfg.action_source == MAMCTS_SRC_PHILOX && cfg.rollouts_per_leaf > 1)
      search_compact_kernel<Env, NE, NA, NCHILD, CIDX, CNT, MAMCTS_SRC_PHILOX, true><<<grid, block_threads, 0, e->stream>>>(q, c);
    else if (cfg.action_source == MAMCTS_SRC_PHILOX)
      search_compact_kernel<Env, NE, NA, NCHILD, CIDX, CNT, MAMCTS_SRC_PHILOX><<<grid, block_threads, 0, e->stream>>>(q, c);
    else
      search_compact_kernel<Env, NE, NA, NCHILD, CIDX, CNT, MAMCTS_SRC_REFERENCE_CONST><<<grid, block_threads, 0, e->stream>>>(q, c);
    e->launches++;
    MAMCTS_CUDA(e, cudaGetLastError());
    return MAMCTS_OK;
  }

  void read_header(const void* node, uint32_t* parent, uint32_t* depth, uint32_t* visits,
                   uint8_t* ja) const override {
    const Inner* nd = static_cast<const Inner*>(node);
    *parent = nd->parent;
    *depth = 0;  // not stored: the export derives it
    *visits = nd->c.h.visit_count;
    for (int a = 0; a < Env::A; ++a) ja[a] = nd->c.ja[a];
  }

  void read_stat(const void* node, uint32_t agent, double* value, double* latest, uint32_t* visits,
                 uint32_t* nexp, double* q, uint32_t* n) const override {
    const Inner* nd = static_cast<const Inner*>(node);
    const uint32_t k = agent ? 1 : 0;
    *value = k ? nd->vg_o[0] : nd->vg_e[0];
    *latest = k ? nd->vg_o[1] : nd->vg_e[1];
    *visits = nd->c.h.N[k];
    *nexp = nd->c.h.nexp[k];
    if (agent == 0) { for (int a = 0; a < NE; ++a) { q[a] = nd->q_e[a]; n[a] = nd->c.n_e[a]; } }
    else {
      // a reward-free agent with a discount >= 0: every action value is +0.0 and the kernels do not store them
      const bool zero = Env::kOtherRewardZero && p.gamma >= 0.0;
      for (int a = 0; a < NA; ++a) { q[a] = zero ? 0.0 : nd->q_o[a]; n[a] = nd->c.n_o[a]; }
    }
  }

  void root_gather(RootGather* g) const override {
    const char* base = static_cast<const char*>(d_inner);
    const size_t tree_bytes = sizeof(Inner) * (size_t)max_inner;
    g->q = reinterpret_cast<const double*>(base + offsetof(Inner, q_e));
    g->n = base + offsetof(Inner, c) + offsetof(typename Inner::Counters, n_e);
    g->visits = base + offsetof(Inner, c) + offsetof(typename Inner::Counters, h) + offsetof(typename Inner::Head, N);
    g->count_bytes = sizeof(CNT);
    g->root_slot = nullptr;
    g->stride_q = tree_bytes / sizeof(double);
    g->stride_n = tree_bytes / sizeof(CNT);
    g->stride_v = tree_bytes / sizeof(CNT);
    g->max_actions = NE;
  }

  // Export (like StageNode::printTree, reference stage_node.h:322-391): depth-first from the root through
  // the child tables, children in ascending joint-action code; a leaf record is written out as the
  // node the reference would hold for it (visit_count 0, N = 1, V = G = h, nothing expanded).
  int dump(uint32_t tree, double* records, uint32_t capacity_records, uint32_t* num_records) override {
    uint32_t ni = 0, nl = 0;
    MAMCTS_CUDA(e, cudaMemcpyAsync(&ni, d_tree_inner + tree, sizeof(ni), cudaMemcpyDeviceToHost, e->stream));
    MAMCTS_CUDA(e, cudaMemcpyAsync(&nl, d_tree_leaves + tree, sizeof(nl), cudaMemcpyDeviceToHost, e->stream));
    MAMCTS_CUDA(e, cudaStreamSynchronize(e->stream));
    std::vector<Inner> in(ni);
    std::vector<Leaf> lf(nl ? nl : 1);
    MAMCTS_CUDA(e, cudaMemcpyAsync(in.data(), root_record(tree), sizeof(Inner) * ni, cudaMemcpyDeviceToHost, e->stream));
    if (nl)
      MAMCTS_CUDA(e, cudaMemcpyAsync(lf.data(), static_cast<const char*>(d_leaves) + sizeof(Leaf) * (size_t)tree * p.max_nodes,
                                     sizeof(Leaf) * nl, cudaMemcpyDeviceToHost, e->stream));
    MAMCTS_CUDA(e, cudaStreamSynchronize(e->stream));
    constexpr uint32_t A = Env::A;
    const uint32_t max_actions = std::max(p.n_ego, p.n_oth);
    const uint32_t stride = 4 + A * (3 + 2 * max_actions);
    uint32_t count = 0;
    auto code_of = [&](uint32_t slot) {   // slot = ja[0] * NA + ja[1]  ->  sum_a ja[a] * max_actions^a
      const uint32_t j0 = slot / NA, j1 = slot % NA;
      return (double)j0 + (double)max_actions * (double)j1;
    };
    struct Kid { double code; uint32_t entry; };
    auto kids_of = [&](const Inner& nd) {
      std::vector<Kid> k;
      for (int sl = 0; sl < NCHILD; ++sl) if (nd.child[sl]) k.push_back({code_of((uint32_t)sl), (uint32_t)nd.child[sl]});
      std::sort(k.begin(), k.end(), [](const Kid& a, const Kid& b) { return a.code < b.code; });
      return k;
    };
    auto emit_inner = [&](uint32_t i, uint32_t depth, double code, size_t nkids) {
      if (count < capacity_records) {
        double* r = records + (size_t)count * stride;
        r[0] = depth; r[1] = code; r[2] = in[i].c.h.visit_count; r[3] = (double)nkids;
        for (uint32_t a = 0; a < A; ++a) {
          double* o = r + 4 + a * (3 + 2 * max_actions);
          double v, l, qq[MAMCTS_MAX_ACTIONS];
          uint32_t vis, nexp, cnt[MAMCTS_MAX_ACTIONS + 1];
          read_stat(&in[i], a, &v, &l, &vis, &nexp, qq, cnt);
          const uint32_t nact = a == 0 ? p.n_ego : p.n_oth;
          const uint32_t mask = tab.exp_mask[a == 0 ? 0 : 1][std::min(nexp, nact)];
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
    auto emit_leaf = [&](uint32_t l, uint32_t depth, double code) {
      if (count < capacity_records) {
        double* r = records + (size_t)count * stride;
        r[0] = depth; r[1] = code; r[2] = 0.0; r[3] = 0.0;
        for (uint32_t a = 0; a < A; ++a) {
          double* o = r + 4 + a * (3 + 2 * max_actions);
          o[0] = 1.0; o[1] = lf[l].h[a]; o[2] = lf[l].h[a];
          for (uint32_t k = 0; k < max_actions; ++k) { o[3 + k] = -1.0; o[3 + max_actions + k] = 0.0; }
        }
      }
      ++count;
    };
    struct Frame { std::vector<Kid> kids; size_t next; uint32_t depth; };
    std::vector<Frame> stack;
    {
      auto k = kids_of(in[0]);
      emit_inner(0, 0, -1.0, k.size());
      stack.push_back({std::move(k), 0, 0});
    }
    while (!stack.empty()) {
      Frame& top = stack.back();
      if (top.next >= top.kids.size()) { stack.pop_back(); continue; }
      const Kid kid = top.kids[top.next++];
      const uint32_t d = top.depth + 1;
      if (kid.entry & 1u) {
        const uint32_t i = kid.entry >> 1;
        if (i >= ni) return fail(e, MAMCTS_ERR_INVALID, "mamcts_search_dump_tree: corrupt child table");
        auto k = kids_of(in[i]);
        emit_inner(i, d, kid.code, k.size());
        stack.push_back({std::move(k), 0, d});
      } else {
        const uint32_t l = (kid.entry >> 1) - 1u;
        if (l >= nl) return fail(e, MAMCTS_ERR_INVALID, "mamcts_search_dump_tree: corrupt child table");
        emit_leaf(l, d, kid.code);
      }
    }
    *num_records = count;
    return MAMCTS_OK;
  }
};

// Compact layout for more than two agents with reward-free other agents: search_compact_multi.cuh.
static_assert(sizeof(InnerRecM<CrossingEnv<7>, 4, 7, uint16_t>) == 192, "8-agent inner record: 192 bytes with 16-bit counts");
template <class Env, int NE, int NA, class CNT>
struct SearchCM final : public mamcts_search {
  using Inner = InnerRecM<Env, NE, NA, CNT>;
  static constexpr int A = Env::A;
  unsigned long long* d_table = nullptr;
  uint32_t table_size = 0;

  size_t node_bytes() const override { return sizeof(Inner); }
  size_t leaf_bytes() const override { return sizeof(double); }
  bool uses_hash() const override { return false; }   // its own 64-bit table, not d_hash
  bool is_compact() const override { return true; }
  void free_extra() override { cudaFree(d_table); }
  const char* root_record(uint32_t tree) const override {
    return static_cast<const char*>(d_inner) + sizeof(Inner) * (size_t)tree * max_inner;
  }
  int alloc_extra(uint32_t max_nodes) override {
    table_size = 16;
    while (table_size * 5u < max_nodes * 8u) table_size <<= 1;   // load factor <= 0.625
    MAMCTS_CUDA(e, cudaMalloc(reinterpret_cast<void**>(&d_table), sizeof(unsigned long long) * (size_t)cfg.num_trees * table_size));
    return MAMCTS_OK;
  }
  int launch_reset() override {
    const uint32_t T = cfg.num_trees;
    MAMCTS_CUDA(e, cudaMemsetAsync(d_counters, 0, sizeof(unsigned long long) * 8, e->stream));
    MAMCTS_CUDA(e, cudaMemsetAsync(d_table, 0, sizeof(unsigned long long) * (size_t)T * table_size, e->stream));
    search_compact_multi_reset_kernel<Env, NE, NA, CNT><<<(T + 255) / 256, 256, 0, e->stream>>>(
        d_inner, d_tree_nodes, d_tree_iters, d_tree_inner, d_tree_leaves, T, max_inner);
    e->launches++;
    MAMCTS_CUDA(e, cudaGetLastError());
    iters_done = 0;
    return MAMCTS_OK;
  }
  int launch_run(uint32_t iterations) override {
    SearchParams q = p;
    q.iterations = iterations;
    CompactMultiParams c;
    c.inner = d_inner; c.leaves = static_cast<double*>(d_leaves); c.table = d_table;
    c.tree_inner = d_tree_inner; c.tree_leaves = d_tree_leaves;
    c.max_inner = max_inner; c.table_mask = table_size - 1; c.root_x = d_root_x;
    const uint32_t T = cfg.num_trees;
    const uint32_t trees_per_block = (uint32_t)(block_threads / 32) * q.lanes_per_warp;
    const int grid = (int)((T + trees_per_block - 1) / trees_per_block);
    if (cfg.action_source == MAMCTS_SRC_PHILOX && cfg.rollouts_per_leaf > 1)
      search_compact_multi_kernel<Env, NE, NA, CNT, MAMCTS_SRC_PHILOX, true><<<grid, block_threads, 0, e->stream>>>(q, c);
    else if (cfg.action_source == MAMCTS_SRC_PHILOX)
      search_compact_multi_kernel<Env, NE, NA, CNT, MAMCTS_SRC_PHILOX><<<grid, block_threads, 0, e->stream>>>(q, c);
    else
      search_compact_multi_kernel<Env, NE, NA, CNT, MAMCTS_SRC_REFERENCE_CONST><<<grid, block_threads, 0, e->stream>>>(q, c);
    e->launches++;
    MAMCTS_CUDA(e, cudaGetLastError());
    return MAMCTS_OK;
  }
  void read_header(const void* node, uint32_t* parent, uint32_t* depth, uint32_t* visits, uint8_t* ja) const override {
    const Inner* nd = static_cast<const Inner*>(node);
    *parent = nd->parent; *depth = 0; *visits = nd->visit_count;
    for (int a = 0; a < A; ++a) ja[a] = nd->ja[a];
  }
  void read_stat(const void* node, uint32_t agent, double* value, double* latest, uint32_t* visits, uint32_t* nexp,
                 double* q, uint32_t* n) const override {
    const Inner* nd = static_cast<const Inner*>(node);
    *visits = nd->N;
    if (agent == 0) {
      *value = nd->vg_e[0]; *latest = nd->vg_e[1]; *nexp = nd->nexp_e;
      for (int a = 0; a < NE; ++a) { q[a] = nd->q_e[a]; n[a] = nd->n_e[a]; }
    } else {   // a reward-free agent: every value is +0.0
      *value = 0.0; *latest = 0.0; *nexp = nd->nexp_o[agent - 1];
      for (int a = 0; a < NA; ++a) { q[a] = 0.0; n[a] = nd->n_o[agent - 1][a]; }
    }
  }
  void root_gather(RootGather* g) const override {
    const char* base = static_cast<const char*>(d_inner);
    const size_t tree_bytes = sizeof(Inner) * (size_t)max_inner;
    g->q = reinterpret_cast<const double*>(base + offsetof(Inner, q_e));
    g->n = base + offsetof(Inner, n_e);
    g->visits = base + offsetof(Inner, N);
    g->count_bytes = sizeof(CNT);
    g->root_slot = nullptr;
    g->stride_q = tree_bytes / sizeof(double);
    g->stride_n = tree_bytes / sizeof(CNT);
    g->stride_v = tree_bytes / sizeof(CNT);
    g->max_actions = NE;
  }
  // Export: depth-first from the root; the children of an inner node are the slots of the tree's table whose
  // key names it, in ascending joint-action code; a leaf is written out as the node the reference would hold.
  int dump(uint32_t tree, double* records, uint32_t capacity_records, uint32_t* num_records) override {
    uint32_t ni = 0, nl = 0;
    MAMCTS_CUDA(e, cudaMemcpyAsync(&ni, d_tree_inner + tree, sizeof(ni), cudaMemcpyDeviceToHost, e->stream));
    MAMCTS_CUDA(e, cudaMemcpyAsync(&nl, d_tree_leaves + tree, sizeof(nl), cudaMemcpyDeviceToHost, e->stream));
    MAMCTS_CUDA(e, cudaStreamSynchronize(e->stream));
    std::vector<Inner> in(ni);
    std::vector<double> lf(nl ? nl : 1);
    std::vector<unsigned long long> tb(table_size);
    MAMCTS_CUDA(e, cudaMemcpyAsync(in.data(), root_record(tree), sizeof(Inner) * ni, cudaMemcpyDeviceToHost, e->stream));
    if (nl)
      MAMCTS_CUDA(e, cudaMemcpyAsync(lf.data(), static_cast<const double*>(d_leaves) + (size_t)tree * p.max_nodes,
                                     sizeof(double) * nl, cudaMemcpyDeviceToHost, e->stream));
    MAMCTS_CUDA(e, cudaMemcpyAsync(tb.data(), d_table + (size_t)tree * table_size, sizeof(unsigned long long) * table_size,
                                   cudaMemcpyDeviceToHost, e->stream));
    MAMCTS_CUDA(e, cudaStreamSynchronize(e->stream));
    const uint32_t max_actions = std::max(p.n_ego, p.n_oth);
    const uint32_t stride = 4 + A * (3 + 2 * max_actions);
    struct Kid { double code; uint32_t entry; };
    std::vector<std::vector<Kid>> kids(ni);
    for (unsigned long long w : tb) {
      if (!w) continue;
      const unsigned long long key = w >> 16;
      const uint32_t parent = (uint32_t)(key >> 24), packed = (uint32_t)(key & 0xFFFFFFull);
      if (parent >= ni) return fail(e, MAMCTS_ERR_INVALID, "mamcts_search_dump_tree: corrupt child table");
      double code = 0.0, mul = 1.0;
      for (int a = 0; a < A; ++a) { code += mul * (double)((packed >> (3 * a)) & 7u); mul *= max_actions; }
      kids[parent].push_back({code, (uint32_t)(w & 0xFFFFull)});
    }
    for (auto& k : kids) std::sort(k.begin(), k.end(), [](const Kid& a, const Kid& b) { return a.code < b.code; });
    uint32_t count = 0;
    auto emit_inner = [&](uint32_t i, uint32_t depth, double code) {
      if (count < capacity_records) {
        double* r = records + (size_t)count * stride;
        r[0] = depth; r[1] = code; r[2] = in[i].visit_count; r[3] = (double)kids[i].size();
        for (uint32_t a = 0; a < (uint32_t)A; ++a) {
          double* o = r + 4 + a * (3 + 2 * max_actions);
          double v, l, qq[MAMCTS_MAX_ACTIONS];
          uint32_t vis, nexp, cnt[MAMCTS_MAX_ACTIONS + 1];
          read_stat(&in[i], a, &v, &l, &vis, &nexp, qq, cnt);
          const uint32_t nact = a == 0 ? p.n_ego : p.n_oth;
          const uint32_t mask = tab.exp_mask[a == 0 ? 0 : 1][std::min(nexp, nact)];
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
    auto emit_leaf = [&](uint32_t l, uint32_t depth, double code) {
      if (count < capacity_records) {
        double* r = records + (size_t)count * stride;
        r[0] = depth; r[1] = code; r[2] = 0.0; r[3] = 0.0;
        for (uint32_t a = 0; a < (uint32_t)A; ++a) {
          double* o = r + 4 + a * (3 + 2 * max_actions);
          o[0] = 1.0; o[1] = a == 0 ? lf[l] : 0.0; o[2] = o[1];
          for (uint32_t k = 0; k < max_actions; ++k) { o[3 + k] = -1.0; o[3 + max_actions + k] = 0.0; }
        }
      }
      ++count;
    };
    struct Frame { uint32_t node; size_t next; uint32_t depth; };
    std::vector<Frame> stack;
    emit_inner(0, 0, -1.0);
    stack.push_back({0u, 0, 0});
    while (!stack.empty()) {
      Frame& top = stack.back();
      if (top.next >= kids[top.node].size()) { stack.pop_back(); continue; }
      const Kid kid = kids[top.node][top.next++];
      const uint32_t d = top.depth + 1;
      if (kid.entry & 1u) {
        const uint32_t i = kid.entry >> 1;
        if (i >= ni) return fail(e, MAMCTS_ERR_INVALID, "mamcts_search_dump_tree: corrupt child table");
        emit_inner(i, d, kid.code);
        stack.push_back({i, 0, d});
      } else {
        const uint32_t l = (kid.entry >> 1) - 1u;
        if (l >= nl) return fail(e, MAMCTS_ERR_INVALID, "mamcts_search_dump_tree: corrupt child table");
        emit_leaf(l, d, kid.code);
      }
    }
    *num_records = count;
    return MAMCTS_OK;
  }
};

template <class Env, int NE, int NA>
inline mamcts_search* make_search_multi(uint32_t max_iterations) {
  if (max_iterations < 65535) return new SearchCM<Env, NE, NA, uint16_t>();
  return new SearchCM<Env, NE, NA, uint32_t>();
}

template <class Env, int NE, int NA, int NCHILD>
inline mamcts_search* make_search_cidx(uint32_t max_nodes, bool compact, uint32_t max_iterations) {
  if (NCHILD > 0 && compact) {
    if constexpr (NCHILD > 0) {
      // 16-bit child entries / counts where everything they hold stays below 2^15 / 2^16
      if (max_nodes + 1 < 32768 && max_iterations < 65535) return new SearchC<Env, NE, NA, NCHILD, uint16_t, uint16_t>();
      if (max_nodes + 1 < 32768) return new SearchC<Env, NE, NA, NCHILD, uint16_t, uint32_t>();
      return new SearchC<Env, NE, NA, NCHILD, uint32_t, uint32_t>();
    }
  }
  if (NCHILD > 0 && max_nodes <= 65535) return new SearchT<Env, NE, NA, NCHILD, uint16_t>();
  return new SearchT<Env, NE, NA, NCHILD, uint32_t>();
}

}  // namespace mamcts
