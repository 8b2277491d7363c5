// cuda_uct_statistic.h - CudaUctStatistic: drop-in for mcts::UctStatistic
// (reference mcts/statistics/uct_statistic.h) as the template parameters SE / SO of
// Mcts<S, SE, SO, H> (mcts/mcts.h:35-36).
//
// It HOLDS the reference's own UctStatistic (`core_`) and forwards everything that is not on the
// rollout-and-backup path to it unchanged - action selection (choose_next_action, progressive widening, UCB,
// bound estimation), get_best_action, get_policy, printing, accessors: none of that is restated here.
// What this class adds is the part that moves to the GPU:
//   update_from_heuristic  (uct_statistic.h:100-110)
//   update_statistic       (uct_statistic.h:117-120,134-144)
// are not executed on the host: the calls of one iteration are recorded (they arrive leaf-first, agent by
// agent, from Mcts::iterate, mcts.h:309-329) and executed as ONE path by mamcts_backup_uct the moment any
// statistic is read again, and merge_node_statistics (uct_statistic.h:165-184) is mamcts_root_merge.
//
// Copyable like the reference's statistic (stage_node.h:283-287 copies them freely); owns no device
// memory.
#ifndef MAMCTS_B200_HOST_CUDA_UCT_STATISTIC_H_
#define MAMCTS_B200_HOST_CUDA_UCT_STATISTIC_H_

#include <algorithm>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "cuda_random_heuristic.h"  // CudaEngine, to_mamcts_params
#include "mamcts_b200.h"
#include "mcts/node_statistic.h"
#include "mcts/statistics/uct_statistic.h"

namespace mcts {

class CudaUctStatistic : public mcts::NodeStatistic<CudaUctStatistic> {
  // the reference's statistic with its protected state reachable for the scatter / gather around the kernels
  struct Core : public UctStatistic {
    using UctStatistic::UctStatistic;
    double& value() { return value_; }
    double& latest() { return latest_return_; }
    unsigned& visits() { return total_node_visits_; }
    UcbStatistics& ucb() { return ucb_statistics_; }
    double value() const { return value_; }
    double latest() const { return latest_return_; }
    unsigned visits() const { return total_node_visits_; }
    double discount() const { return k_discount_factor; }
  };

 public:
  using UcbPair = UctStatistic::UcbPair;
  using UcbStatistics = UctStatistic::UcbStatistics;

  CudaUctStatistic(ActionIdx num_actions, AgentIdx agent_idx, const MctsParameters& mcts_parameters)
      : NodeStatistic<CudaUctStatistic>(num_actions, agent_idx, mcts_parameters),
        core_(num_actions, agent_idx, mcts_parameters) {}
  // a copy must see every recorded update
  CudaUctStatistic(const CudaUctStatistic& o) : NodeStatistic<CudaUctStatistic>(o), core_((flush(), o.core_)) {}
  // (not assignable, like the reference's statistic: NodeStatistic holds a reference to the parameters)
  ~CudaUctStatistic() { forget(this); }

  // ---- not on the GPU path: the reference's own code, on up-to-date values -------------------------
  template <class S>
  ActionIdx choose_next_action(const S& state) { flush(); return core_.choose_next_action(state); }
  ActionIdx get_best_action() const { flush(); return core_.get_best_action(); }
  Policy get_policy() const { flush(); return core_.get_policy(); }
  std::string print_node_information() const { flush(); return core_.print_node_information(); }
  std::string print_edge_information(const ActionIdx& action) const { flush(); return core_.print_edge_information(action); }
  std::string sprintf() const { flush(); return core_.sprintf(); }
  const UcbStatistics& get_ucb_statistics() const { flush(); return core_.get_ucb_statistics(); }
  double get_value() const { flush(); return core_.value(); }
  double get_latest_return() const { flush(); return core_.latest(); }
  unsigned get_total_node_visits() const { flush(); return core_.visits(); }
  Reward get_reward_lower_bound() const { return core_.get_reward_lower_bound(); }
  Reward get_reward_upper_bound() const { return core_.get_reward_upper_bound(); }
  unsigned int num_expanded_actions() const { flush(); return core_.num_expanded_actions(); }
  void set_heuristic_estimate(const Reward& accum_rewards, const EgoCosts& accum_ego_cost) {
    core_.set_heuristic_estimate(accum_rewards, accum_ego_cost);
  }

  // recorded; executed by mamcts_backup_uct at the next read (see flush())
  void update_from_heuristic(const NodeStatistic<CudaUctStatistic>& heuristic_statistic) {
    const CudaUctStatistic& h = heuristic_statistic.impl();
    pending().push_back(Op{this, nullptr, h.core_.value(), 0, 0.0, true});
  }
  void update_statistic(const NodeStatistic<CudaUctStatistic>& changed_child_statistic) {
    const CudaUctStatistic& child = changed_child_statistic.impl();
    pending().push_back(Op{this, &child, 0.0, collected_reward_.first, collected_reward_.second, false});
  }

  // uct_statistic.h:165-184 through mamcts_root_merge (+ the NCCL all-reduce when the engine has a
  // communicator attached)
  void merge_node_statistics(const std::vector<CudaUctStatistic>& statistics) {
    flush();
    const uint32_t T = (uint32_t)statistics.size();
    uint32_t MA = 1;
    for (const auto& s : statistics) MA = std::max<uint32_t>(MA, (uint32_t)s.num_actions_);
    for (const auto& s : statistics)
      for (const auto& kv : s.core_.get_ucb_statistics()) MA = std::max<uint32_t>(MA, (uint32_t)kv.first + 1);
    std::vector<double> value(T, 0.0), latest(T, 0.0), q((size_t)T * MA, 0.0);
    std::vector<uint32_t> visits(T, 0), n((size_t)T * MA, 0), slots(T);
    for (uint32_t t = 0; t < T; ++t) {
      slots[t] = t;
      visits[t] = statistics[t].core_.visits();
      for (const auto& kv : statistics[t].core_.get_ucb_statistics()) {
        q[(size_t)t * MA + kv.first] = kv.second.action_value_;
        n[(size_t)t * MA + kv.first] = kv.second.action_count_;
      }
    }
    mamcts_uct_stats st;
    std::memset(&st, 0, sizeof(st));
    st.mem = MAMCTS_MEM_HOST; st.num_slots = T; st.max_actions = MA;
    st.value = value.data(); st.latest_return = latest.data(); st.total_visits = visits.data();
    st.q = q.data(); st.n = n.data();
    double mq[MAMCTS_MAX_ACTIONS], mv = 0.0;
    uint64_t mn[MAMCTS_MAX_ACTIONS], mvis = 0;
    uint32_t occ[MAMCTS_MAX_ACTIONS], best = 0;
    CudaEngine::check(mamcts_root_merge(CudaEngine::get(), &st, T, MAMCTS_MEM_HOST, slots.data(), MA, mq, mn,
                                        occ, &mvis, &mv, &best), "mamcts_root_merge");
    core_.ucb().clear();
    for (uint32_t a = 0; a < MA; ++a)
      if (occ[a]) core_.ucb()[a] = UcbPair((unsigned)mn[a], mq[a], 0.0);
    core_.visits() += (unsigned)mvis;
    core_.value() = mv;
  }

  // Executes everything recorded since the last read as ONE mamcts_backup_uct call. The recorded
  // calls form one path per iteration (several when a batched driver advanced many trees before
  // anything was read): per agent, a leaf update starts a path and each update_statistic(child)
  // continues the path whose last statistic is `child`.
  static void flush() {
    std::vector<Op>& ops = pending();
    if (ops.empty()) return;
    std::vector<Op> work;
    work.swap(ops);
    std::vector<AgentIdx> agents;
    std::unordered_map<AgentIdx, std::vector<const Op*>> per_agent;
    for (const Op& op : work) {
      const AgentIdx id = op.self->agent_idx_;
      if (!per_agent.count(id)) agents.push_back(id);
      per_agent[id].push_back(&op);
    }
    const uint32_t A = (uint32_t)agents.size();
    struct Chain {
      const Op* leaf = nullptr;                    // the update_from_heuristic of the leaf, if any
      const CudaUctStatistic* leaf_ro = nullptr;   // re-visited terminal leaf: read only
      std::vector<const Op*> edges;
      const CudaUctStatistic* tail() const {
        if (!edges.empty()) return edges.back()->self;
        return leaf ? leaf->self : leaf_ro;
      }
    };
    std::vector<std::vector<Chain>> chains(A);
    for (uint32_t k = 0; k < A; ++k) {
      for (const Op* op : per_agent[agents[k]]) {
        if (op->is_leaf) {
          chains[k].emplace_back();
          chains[k].back().leaf = op;
        } else if (!chains[k].empty() && chains[k].back().tail() == op->child) {
          chains[k].back().edges.push_back(op);
        } else {
          chains[k].emplace_back();
          chains[k].back().leaf_ro = op->child;
          chains[k].back().edges.push_back(op);
        }
      }
    }
    const uint32_t P = (uint32_t)chains[0].size();
    for (uint32_t k = 1; k < A; ++k) MCTS_EXPECT_TRUE(chains[k].size() == P);
    // node numbering: path p occupies nodes base[p] (leaf) .. base[p] + len[p]
    std::vector<uint32_t> base(P), len(P), off(P), leaf_node(P);
    std::vector<uint8_t> has(P);
    uint32_t n_nodes = 0, n_edges = 0;
    for (uint32_t pth = 0; pth < P; ++pth) {
      base[pth] = n_nodes;
      len[pth] = (uint32_t)chains[0][pth].edges.size();
      off[pth] = n_edges;
      leaf_node[pth] = n_nodes;
      has[pth] = chains[0][pth].leaf ? 1 : 0;
      n_nodes += len[pth] + 1;
      n_edges += len[pth];
      for (uint32_t k = 1; k < A; ++k) MCTS_EXPECT_TRUE(chains[k][pth].edges.size() == len[pth]);
    }
    auto max_act = [](const CudaUctStatistic* s) {
      uint32_t m = (uint32_t)s->num_actions_;
      for (const auto& kv : s->core_.get_ucb_statistics()) m = std::max<uint32_t>(m, (uint32_t)kv.first + 1);
      return m;
    };
    uint32_t MA = 1;
    for (uint32_t k = 0; k < A; ++k)
      for (const Chain& c : chains[k]) {
        if (c.tail()) MA = std::max(MA, max_act(c.leaf ? c.leaf->self : c.leaf_ro));
        for (const Op* op : c.edges) MA = std::max(MA, max_act(op->self));
      }
    const size_t S = (size_t)n_nodes * A;
    std::vector<double> value(S, 0.0), latest(S, 0.0), q(S * MA, 0.0);
    std::vector<uint32_t> visits(S, 0), n(S * MA, 0);
    auto fill = [&](size_t slot, const CudaUctStatistic* s) {
      value[slot] = s->core_.value(); latest[slot] = s->core_.latest(); visits[slot] = s->core_.visits();
      for (const auto& kv : s->core_.get_ucb_statistics()) {
        q[slot * MA + kv.first] = kv.second.action_value_;
        n[slot * MA + kv.first] = kv.second.action_count_;
      }
    };
    std::vector<double> ego_ret(P, 0.0), other_ret((size_t)(A > 1 ? A - 1 : 1) * P, 0.0);
    std::vector<uint32_t> edge_node(n_edges ? n_edges : 1);
    std::vector<uint8_t> edge_action((size_t)(n_edges ? n_edges : 1) * A, 0);
    std::vector<double> edge_reward((size_t)(n_edges ? n_edges : 1) * A, 0.0);
    for (uint32_t k = 0; k < A; ++k) {
      for (uint32_t pth = 0; pth < P; ++pth) {
        const Chain& c = chains[k][pth];
        fill((size_t)base[pth] * A + k, c.leaf ? c.leaf->self : c.leaf_ro);
        if (c.leaf) (k == 0 ? ego_ret[pth] : other_ret[(size_t)(k - 1) * P + pth]) = c.leaf->heuristic;
        for (uint32_t e = 0; e < len[pth]; ++e) {
          const Op* op = c.edges[e];
          fill((size_t)(base[pth] + e + 1) * A + k, op->self);
          edge_node[off[pth] + e] = base[pth] + e + 1;
          edge_action[(size_t)(off[pth] + e) * A + k] = (uint8_t)op->action;
          edge_reward[(size_t)(off[pth] + e) * A + k] = op->reward;
        }
      }
    }
    mamcts_params mp;
    mamcts_default_params(&mp);
    mp.discount_factor = work.front().self->core_.discount();
    mamcts_uct_stats st;
    std::memset(&st, 0, sizeof(st));
    st.mem = MAMCTS_MEM_HOST; st.num_slots = (uint32_t)S; st.max_actions = MA;
    st.value = value.data(); st.latest_return = latest.data(); st.total_visits = visits.data();
    st.q = q.data(); st.n = n.data();
    CudaEngine::check(mamcts_backup_uct(CudaEngine::get(), &mp, A, P, MAMCTS_MEM_HOST, leaf_node.data(),
                                        has.data(), ego_ret.data(), other_ret.data(), off.data(), len.data(),
                                        edge_node.data(), edge_action.data(), edge_reward.data(), &st),
                      "mamcts_backup_uct");
    // scatter the results back into the mirrors
    auto take = [&](size_t slot, CudaUctStatistic* s) {
      s->core_.value() = value[slot];
      s->core_.latest() = latest[slot];
      s->core_.visits() = visits[slot];
    };
    for (uint32_t k = 0; k < A; ++k) {
      for (uint32_t pth = 0; pth < P; ++pth) {
        const Chain& c = chains[k][pth];
        if (c.leaf) take((size_t)base[pth] * A + k, c.leaf->self);
        for (uint32_t e = 0; e < len[pth]; ++e) {
          const Op* op = c.edges[e];
          const size_t slot = (size_t)(base[pth] + e + 1) * A + k;
          take(slot, op->self);
          UcbPair& pair = op->self->core_.ucb()[op->action];
          pair.action_count_ = n[slot * MA + op->action];
          pair.action_value_ = q[slot * MA + op->action];
        }
      }
    }
  }

 protected:
  struct Op {
    CudaUctStatistic* self;
    const CudaUctStatistic* child;
    double heuristic;
    ActionIdx action;
    double reward;
    bool is_leaf;
  };
  // Per host thread (parallel_search runs one Mcts per std::thread, mcts.h:198-210). A thread that
  // ends with recorded updates (single_search returns right after the last iterate()) executes them
  // on exit, while the trees - owned by the parent Mcts - are still alive.
  struct PendingList {
    std::vector<Op> ops;
    ~PendingList() { if (!ops.empty()) CudaUctStatistic::flush(); }
  };
  static std::vector<Op>& pending() {
    static thread_local PendingList list;
    return list.ops;
  }
  static void forget(const CudaUctStatistic* s) {
    // a statistic with recorded updates only dies when its whole tree is torn down unread: drop them
    std::vector<Op>& ops = pending();
    for (const Op& op : ops) if (op.self == s || op.child == s) { ops.clear(); return; }
  }

  Core core_;
};

}  // namespace mcts
#endif
