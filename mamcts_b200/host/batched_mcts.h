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
    for (num_iterations_ = 0; num_iterations_ < mcts_parameters_.MAX_NUMBER_OF_ITERATIONS; ++num_iterations_) wave();
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
  void wave() {
    using CS = CudaState<S>;
    const uint32_t T = num_trees_;
    std::vector<NodePtr> leaf(T);
    std::vector<uint8_t> needs_rollout(T, 0);
    std::vector<uint32_t> batch;  // trees whose leaf needs a rollout
    // ---- select & expand, one leaf per tree (mcts.h:300-305) ------------------------------------
    parallel_over_trees([&](uint32_t t) {
      NodePtr node = roots_[t];
      std::pair<bool, bool> res(true, true);
      while (res.first) res = node->select_or_expand(node);
      leaf[t] = node;
      needs_rollout[t] = res.second ? 1 : 0;
    });
    for (uint32_t t = 0; t < T; ++t) if (needs_rollout[t]) batch.push_back(t);
    // ---- ONE launch for all leaves (replaces T x heuristic_.calculate_heuristic_values, mcts.h:310)
    const uint32_t n = (uint32_t)batch.size();
    const uint32_t A = n ? CS::num_agents(*leaf[batch[0]]->get_state()) : 0;
    std::vector<double> ego(n), other((size_t)(A > 1 ? A - 1 : 1) * (n ? n : 1)), cost(n), len(n);
    if (n) {
      const uint32_t XR = CS::kEnv == MAMCTS_ENV_SIMPLE ? 1 : A;
      std::vector<int32_t> x((size_t)XR * n), last((size_t)A * n);
      std::vector<uint8_t> flags(n);
      std::vector<uint32_t> depth(n), stream_hi(n), stream_lo(n);
      for (uint32_t i = 0; i < n; ++i) {
        int32_t xi[MAMCTS_MAX_AGENTS] = {0}, li[MAMCTS_MAX_AGENTS] = {0};
        CS::pack(*leaf[batch[i]]->get_state(), xi, li, &flags[i]);
        for (uint32_t a = 0; a < XR; ++a) x[(size_t)a * n + i] = xi[a];
        for (uint32_t a = 0; a < A; ++a) last[(size_t)a * n + i] = li[a];
        depth[i] = leaf[batch[i]]->get_depth();
        stream_hi[i] = first_tree_id_ + batch[i];
        stream_lo[i] = ++calls_[batch[i]] * K_;
      }
      const mamcts_params mp = to_mamcts_params(mcts_parameters_);
      mamcts_action_source src;
      std::memset(&src, 0, sizeof(src));
      src.kind = MAMCTS_SRC_PHILOX;
      src.mem = MAMCTS_MEM_HOST;
      src.philox_seed = philox_seed_;
      src.stream_hi = stream_hi.data();
      src.stream_lo = stream_lo.data();
      mamcts_rollout_out out;
      std::memset(&out, 0, sizeof(out));
      out.mem = MAMCTS_MEM_HOST;
      out.ego_return = ego.data();
      out.other_return = other.data();
      out.ego_cost = cost.data();
      out.step_length = len.data();
      uint64_t steps = 0;
      out.total_steps = &steps;
      if constexpr (CS::kEnv == MAMCTS_ENV_CROSSING_I32) {
        CudaEngine::check(mamcts_rollout_crossing_i32(CudaEngine::get(), &mp, &env_, n, K_,
                                                      MAMCTS_MEM_HOST, x.data(), last.data(), flags.data(),
                                                      depth.data(), &src, &out),
                          "mamcts_rollout_crossing_i32");
      } else {
        CudaEngine::check(mamcts_rollout_simple(CudaEngine::get(), &mp, &env_, n, K_, MAMCTS_MEM_HOST,
                                                x.data(), depth.data(), &src, &out),
                          "mamcts_rollout_simple");
      }
      rollout_steps_ += steps;
    }
    // ---- update_statistics + back-propagation per tree (mcts.h:309-329) -------------------------
    std::vector<int32_t> batch_index(T, -1);
    for (uint32_t i = 0; i < n; ++i) batch_index[batch[i]] = (int32_t)i;
    auto finish_tree = [&](uint32_t t) {
      NodePtr node = leaf[t];
      if (batch_index[t] >= 0) {
        const uint32_t bi = (uint32_t)batch_index[t];
        const S* st = node->get_state();
        SE ego_h(0, st->get_ego_agent_idx(), mcts_parameters_);
        EgoCosts accum_cost(st->get_num_costs(), 0.0);
        if (!accum_cost.empty()) accum_cost[0] = cost[bi];
        ego_h.set_heuristic_estimate(ego[bi], accum_cost);
        std::unordered_map<AgentIdx, SO> others_h;
        uint32_t j = 0;
        for (auto ai : st->get_other_agent_idx()) {
          SO so(0, ai, mcts_parameters_);
          so.set_heuristic_estimate(other[(size_t)j * n + bi], accum_cost);
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
    };
    // statistics that record their updates per host thread (CudaUctStatistic) stay on one thread
    if (kDeferredStatistic) for (uint32_t t = 0; t < T; ++t) finish_tree(t);
    else parallel_over_trees(finish_tree);
    // deferred statistics (CudaUctStatistic): all T paths of the wave in one mamcts_backup_uct
    flush_if_deferred<SE>(0);
  }

  template <class F>
  void parallel_over_trees(F&& f) {
    const uint32_t T = num_trees_;
    const unsigned nt = std::min<unsigned>(host_threads_, std::max(1u, T / 64));
    if (nt <= 1) { for (uint32_t t = 0; t < T; ++t) f(t); return; }
    std::vector<std::thread> pool;
    for (unsigned k = 0; k < nt; ++k)
      pool.emplace_back([&, k] { for (uint32_t t = (uint32_t)((uint64_t)T * k / nt); t < (uint32_t)((uint64_t)T * (k + 1) / nt); ++t) f(t); });
    for (auto& th : pool) th.join();
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
