// batched_mcts.h - BatchedRootParallelMcts: the leaf-parallel batcher for HOST trees.
//
// The reference advances one tree by one leaf per Mcts::iterate call (mcts/mcts.h:294-333) and runs
// its root-parallel trees one after the other or one std::thread each (mcts.h:188-221). This driver
// keeps the reference's own tree code - StageNode::select_or_expand / update_statistics, the
// statistics SE / SO - but advances T independent trees in lock step: every wave gathers one leaf
// per tree, evaluates ALL leaves with ONE mamcts_rollout_* launch, then feeds each tree's
// update_statistics and back-propagation loop exactly as Mcts::iterate does. Every tree therefore
// sees the same sequence of operations as under the reference (batch across trees,
// never several leaves of one tree).
//
// It only uses public members of StageNode (mcts/stage_node.h:82-111). With SE = SO = UctStatistic
// the trees equal the reference's own tree code driven by the same Philox streams; with
// CudaUctStatistic the T back-propagation paths of a wave are additionally executed as one
// mamcts_backup_uct call.
//
// Philox streams: the c-th heuristic evaluation (c = 1, 2, ...) of tree t uses rollout k's stream
// (first_tree_id + t, c * K + k) - the same definition as the device-resident search, whose c is
// the index of the node being evaluated.
#ifndef MAMCTS_B200_HOST_BATCHED_MCTS_H_
#define MAMCTS_B200_HOST_BATCHED_MCTS_H_

#include <algorithm>
#include <atomic>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "cuda_random_heuristic.h"
#include "mcts/mcts.h"

namespace mcts {

template <class S, class SE, class SO>
class BatchedRootParallelMcts {
 public:
  using H = CudaRandomHeuristic;
  using Node = StageNode<S, SE, SO, H>;
  using NodePtr = std::shared_ptr<Node>;

  // host_threads: the trees' select/expand and back-propagation phases of a wave are split over this
  // many std::threads (0 = hardware concurrency), as the reference's parallel_search does with
  // USE_MULTI_THREADING (mcts.h:198-210) - including its caveat: StageNode::num_nodes_ is one static
  // counter shared by all trees (stage_node.h:77), so node ids and the MAX_NUMBER_OF_NODES cut-off are
  // only exact with host_threads = 1. Statistics of different trees never alias.
  // env_parameters: the environment's parameters in the C ABI's form (CudaState<S>::to_params(...)); nullptr =
  // whatever is registered with CudaState<S>::set_parameters at the time of each search().
  BatchedRootParallelMcts(const MctsParameters& mcts_parameters, unsigned num_trees,
                          uint64_t philox_seed = 1000, unsigned first_tree_id = 0,
                          unsigned rollouts_per_leaf = 1, unsigned host_threads = 1,
                          const typename CudaState<S>::EnvParams* env_parameters = nullptr)
      : mcts_parameters_(mcts_parameters), num_trees_(num_trees), philox_seed_(philox_seed),
        first_tree_id_(first_tree_id), K_(rollouts_per_leaf), num_iterations_(0), rollout_steps_(0),
        host_threads_(host_threads ? host_threads : std::max(1u, std::thread::hardware_concurrency())),
        own_env_(env_parameters != nullptr) {
    if (env_parameters) env_ = *env_parameters;
    CudaEngine::get();
  }

  // Mcts::search for every tree (mcts.h:166-186 without the wall-clock limit).
  // NB StageNode::num_nodes_ is one static counter per instantiation (stage_node.h:77): as in the
  // reference's parallel_search, MAX_NUMBER_OF_NODES is a budget over ALL trees.
  void search(const S& current_state) {
    if (!own_env_) env_ = CudaState<S>::params();
    CudaState<S>::check_state(current_state, env_);
    Node::reset_counter();
    roots_.clear();
    calls_.assign(num_trees_, 0);
    for (unsigned t = 0; t < num_trees_; ++t)
      roots_.push_back(std::make_shared<Node, NodePtr, std::shared_ptr<S>, const JointAction&, const unsigned int&>(
          nullptr, current_state.clone(), JointAction(), 0, mcts_parameters_));
    run_waves(mcts_parameters_.MAX_NUMBER_OF_ITERATIONS);
  }

  const Node& get_root(unsigned tree) const { return *roots_.at(tree); }
  NodePtr get_root_ptr(unsigned tree) const { return roots_.at(tree); }
  unsigned numIterations() const { return num_iterations_; }
  uint64_t rolloutSteps() const { return rollout_steps_; }

  // Root-parallel decision: SE::merge_node_statistics over the ego roots (mcts.h:270-292,
  // stage_node.h:280-289), then get_best_action (uct_statistic.h:78-89).
  ActionIdx returnBestAction() {
    std::vector<SE> ego_statistics;
    for (const auto& r : roots_) ego_statistics.push_back(r->get_ego_int_node());
    const S* st = roots_[0]->get_state();
    SE merged(st->get_num_actions(st->get_ego_agent_idx()), st->get_ego_agent_idx(), mcts_parameters_);
    merged.merge_node_statistics(ego_statistics);
    return merged.get_best_action();
  }

 private:
  // Per-wave buffers, allocated once per search: slot t belongs to tree t (no compaction - a tree whose
  // leaf needs no rollout is marked terminal, which makes its rollout a no-op), so every phase but the one
  // launch is independent per tree and is split over the host threads.
  struct WaveBuffers {
    std::vector<NodePtr> leaf;
    std::vector<uint8_t> needs_rollout, flags;
    std::vector<int32_t> x, last;
    std::vector<uint32_t> depth, stream_hi, stream_lo;
    std::vector<double> ego, other, cost, len;
  };
  // A reusable barrier for the worker threads (they spin: a phase is a few milliseconds, the launch ~100 us).
  struct SpinBarrier {
    explicit SpinBarrier(unsigned n) : n_(n) {}
    void wait() {
      const unsigned gen = gen_.load(std::memory_order_acquire);
      if (count_.fetch_add(1, std::memory_order_acq_rel) + 1 == n_) {
        count_.store(0, std::memory_order_relaxed);
        gen_.fetch_add(1, std::memory_order_release);
      } else {
        while (gen_.load(std::memory_order_acquire) == gen) std::this_thread::yield();
      }
    }
    const unsigned n_;
    std::atomic<unsigned> count_{0}, gen_{0};
  };

  void run_waves(unsigned iterations) {
    using CS = CudaState<S>;
    const uint32_t T = num_trees_;
    const uint32_t A = CS::num_agents(*roots_[0]->get_state());
    const uint32_t XR = CS::kEnv == MAMCTS_ENV_SIMPLE ? 1 : A;
    WaveBuffers w;
    w.leaf.resize(T); w.needs_rollout.assign(T, 0); w.flags.assign(T, 0);
    w.x.assign((size_t)XR * T, 0); w.last.assign((size_t)A * T, 0);
    w.depth.assign(T, 0); w.stream_hi.assign(T, 0); w.stream_lo.assign(T, 0);
    w.ego.assign(T, 0.0); w.other.assign((size_t)(A > 1 ? A - 1 : 1) * T, 0.0); w.cost.assign(T, 0.0); w.len.assign(T, 0.0);
    // statistics that record their updates per host thread (CudaUctStatistic) stay on one thread
    const unsigned nt = kDeferredStatistic ? 1u : std::min<unsigned>(host_threads_, std::max(1u, T / 64));
    SpinBarrier barrier(nt);
    auto worker = [&](unsigned k) {
      const uint32_t t0 = (uint32_t)((uint64_t)T * k / nt), t1 = (uint32_t)((uint64_t)T * (k + 1) / nt);
      for (unsigned it = 0; it < iterations; ++it) {
        for (uint32_t t = t0; t < t1; ++t) select_and_pack(w, t, T, A, XR);
        barrier.wait();
        if (k == 0) { launch_rollouts(w, T, A); num_iterations_ = it + 1; }
        barrier.wait();
        for (uint32_t t = t0; t < t1; ++t) finish_tree(w, t, T);
        // deferred statistics (CudaUctStatistic, one thread): all T paths of the wave in one mamcts_backup_uct
        if (k == 0) flush_if_deferred<SE>(0);
      }
    };
    num_iterations_ = 0;
    if (nt <= 1) { worker(0); return; }
    std::vector<std::thread> pool;
    for (unsigned k = 1; k < nt; ++k) pool.emplace_back(worker, k);
    worker(0);
    for (auto& th : pool) th.join();
  }

  // ---- select & expand, one leaf per tree (mcts.h:300-305), and its state into slot t ----------------
  void select_and_pack(WaveBuffers& w, uint32_t t, uint32_t T, uint32_t A, uint32_t XR) {
    using CS = CudaState<S>;
    NodePtr node = roots_[t];
    std::pair<bool, bool> res(true, true);
    while (res.first) res = node->select_or_expand(node);
    w.leaf[t] = node;
    w.needs_rollout[t] = res.second ? 1 : 0;
    int32_t xi[MAMCTS_MAX_AGENTS] = {0}, li[MAMCTS_MAX_AGENTS] = {0};
    uint8_t fl = 0;
    CS::pack(*node->get_state(), xi, li, &fl);
    for (uint32_t a = 0; a < XR; ++a) w.x[(size_t)a * T + t] = xi[a];
    for (uint32_t a = 0; a < A; ++a) w.last[(size_t)a * T + t] = li[a];
    w.flags[t] = res.second ? fl : (uint8_t)1;   // no rollout wanted: a terminal start state skips the loop
    w.depth[t] = node->get_depth();
    w.stream_hi[t] = first_tree_id_ + t;
    if (res.second) w.stream_lo[t] = ++calls_[t] * K_;
  }

  // ---- ONE launch for all leaves (replaces T x heuristic_.calculate_heuristic_values, mcts.h:310) ------
  void launch_rollouts(WaveBuffers& w, uint32_t T, uint32_t A) {
    using CS = CudaState<S>;
    (void)A;
    const mamcts_params mp = to_mamcts_params(mcts_parameters_);
    mamcts_action_source src;
    std::memset(&src, 0, sizeof(src));
    src.kind = MAMCTS_SRC_PHILOX;
    src.mem = MAMCTS_MEM_HOST;
    src.philox_seed = philox_seed_;
    src.stream_hi = w.stream_hi.data();
    src.stream_lo = w.stream_lo.data();
    mamcts_rollout_out out;
    std::memset(&out, 0, sizeof(out));
    out.mem = MAMCTS_MEM_HOST;
    out.ego_return = w.ego.data();
    out.other_return = w.other.data();
    out.ego_cost = w.cost.data();
    out.step_length = w.len.data();
    uint64_t steps = 0;
    out.total_steps = &steps;
    if constexpr (CS::kEnv == MAMCTS_ENV_CROSSING_I32) {
      CudaEngine::check(mamcts_rollout_crossing_i32(CudaEngine::get(), &mp, &env_, T, K_, MAMCTS_MEM_HOST, w.x.data(),
                                                    w.last.data(), w.flags.data(), w.depth.data(), &src, &out),
                        "mamcts_rollout_crossing_i32");
    } else {
      // SimpleState: is_terminal() is a function of the state length; a tree that wants no rollout gets the
      // depth limit instead (random_heuristic.h:45-47)
      for (uint32_t t = 0; t < T; ++t) if (!w.needs_rollout[t]) w.depth[t] = mcts_parameters_.MAX_SEARCH_DEPTH;
      CudaEngine::check(mamcts_rollout_simple(CudaEngine::get(), &mp, &env_, T, K_, MAMCTS_MEM_HOST, w.x.data(),
                                              w.depth.data(), &src, &out),
                        "mamcts_rollout_simple");
    }
    rollout_steps_ += steps;
  }

  // ---- update_statistics + back-propagation of one tree (mcts.h:309-329) -------------------------------
  void finish_tree(WaveBuffers& w, uint32_t t, uint32_t T) {
    NodePtr node = w.leaf[t];
    if (w.needs_rollout[t]) {
      const S* st = node->get_state();
      SE ego_h(0, st->get_ego_agent_idx(), mcts_parameters_);
      EgoCosts accum_cost(st->get_num_costs(), 0.0);
      if (!accum_cost.empty()) accum_cost[0] = w.cost[t];
      ego_h.set_heuristic_estimate(w.ego[t], accum_cost);
      std::unordered_map<AgentIdx, SO> others_h;
      uint32_t j = 0;
      for (auto ai : st->get_other_agent_idx()) {
        SO so(0, ai, mcts_parameters_);
        so.set_heuristic_estimate(w.other[(size_t)j * T + t], accum_cost);
        others_h.insert(std::pair<AgentIdx, SO>(ai, so));
        ++j;
      }
      node->update_statistics(ego_h, others_h);
    }
    NodePtr node_p = node->get_parent().lock();
    while (bool(node_p)) {
      node_p->update_statistics(node);
      if (node_p->is_root()) break;
      node = node->get_parent().lock();
      node_p = node_p->get_parent().lock();
    }
  }

  template <class Stat>
  static auto has_flush(int) -> decltype(Stat::flush(), std::true_type());
  template <class Stat>
  static std::false_type has_flush(...);
  static constexpr bool kDeferredStatistic = decltype(has_flush<SE>(0))::value;

  template <class Stat>
  static auto flush_if_deferred(int) -> decltype(Stat::flush(), void()) { Stat::flush(); }
  template <class Stat>
  static void flush_if_deferred(...) {}

  const MctsParameters mcts_parameters_;
  unsigned num_trees_;
  uint64_t philox_seed_;
  unsigned first_tree_id_;
  unsigned K_;
  unsigned num_iterations_;
  uint64_t rollout_steps_;
  unsigned host_threads_;
  bool own_env_;
  typename CudaState<S>::EnvParams env_{};
  std::vector<NodePtr> roots_;
  std::vector<uint32_t> calls_;
};

}  // namespace mcts
#endif
