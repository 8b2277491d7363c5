// cuda_uct_statistic.h - CudaUctStatistic: drop-in for mcts::UctStatistic
// (reference mcts/statistics/uct_statistic.h) as the template parameters SE / SO of
// Mcts<S, SE, SO, H> (mcts/mcts.h:35-36). It is a host mirror with UctStatistic's members and
// accessors (Statistic concept, mcts/node_statistic.h:21-70); the arithmetic of the back-propagation
//   update_from_heuristic  (uct_statistic.h:100-110)
//   update_statistic       (uct_statistic.h:117-120,134-144)
// is not done here: the calls of one iteration are recorded (they arrive leaf-first, agent by agent,
// from Mcts::iterate, mcts.h:309-329) and executed as ONE path by mamcts_backup_uct the moment any
// statistic is read again; merge_node_statistics (uct_statistic.h:165-184) is mamcts_root_merge.
// Selection (choose_next_action, uct_statistic.h:61-76) stays host code here - the device-resident
// search (mamcts_search_*) is where it runs on the GPU.
//
// Copyable like the reference's statistic (stage_node.h:283-287 copies them freely); owns no device
// memory.
#ifndef MAMCTS_B200_HOST_CUDA_UCT_STATISTIC_H_
#define MAMCTS_B200_HOST_CUDA_UCT_STATISTIC_H_

#include <algorithm>
#include <cmath>
#include <cstring>
#include <iomanip>
#include <limits>
#include <map>
#include <numeric>
#include <random>
#include <sstream>
#include <unordered_map>
#include <vector>

#include "cuda_random_heuristic.h"  // CudaEngine, to_mamcts_params
#include "mamcts_b200.h"
#include "mcts/node_statistic.h"
#include "mcts/random_generator.h"

namespace mcts {

class CudaUctStatistic : public mcts::NodeStatistic<CudaUctStatistic>, mcts::RandomGenerator {
 public:
  typedef struct UcbPair {
    UcbPair() : action_count_(0), action_value_(0.0f), init_value_(0.0) {}
    UcbPair(unsigned count, double value, double init_value)
        : action_count_(count), action_value_(value), init_value_(init_value) {}
    unsigned action_count_;
    double action_value_;
    double init_value_;
  } UcbPair;
  typedef std::map<ActionIdx, UcbPair> UcbStatistics;

  CudaUctStatistic(ActionIdx num_actions, AgentIdx agent_idx, const MctsParameters& mcts_parameters)
      : NodeStatistic<CudaUctStatistic>(num_actions, agent_idx, mcts_parameters),
        RandomGenerator(mcts_parameters.RANDOM_SEED),
        value_(0.0f), latest_return_(0.0), ucb_statistics_(), total_node_visits_(0),
        unexpanded_actions_(num_actions),
        upper_bound(mcts_parameters.uct_statistic.UPPER_BOUND),
        lower_bound(mcts_parameters.uct_statistic.LOWER_BOUND),
        k_discount_factor(mcts_parameters.DISCOUNT_FACTOR),
        exploration_constant(mcts_parameters.uct_statistic.EXPLORATION_CONSTANT),
        progressive_widening_k(mcts_parameters.uct_statistic.PROGRESSIVE_WIDENING_K),
        progressive_widening_alpha(mcts_parameters.uct_statistic.PROGRESSIVE_WIDENING_ALPHA),
        use_bound_estimation_(mcts_parameters.USE_BOUND_ESTIMATION) {
    std::iota(unexpanded_actions_.begin(), unexpanded_actions_.end(), 0);
  }
  // a copy must see every recorded update
  CudaUctStatistic(const CudaUctStatistic& o) : NodeStatistic<CudaUctStatistic>(o), RandomGenerator(o) {
    flush();
    copy_fields(o);
  }
  CudaUctStatistic& operator=(const CudaUctStatistic& o) {
    flush();
    NodeStatistic<CudaUctStatistic>::collected_reward_ = o.collected_reward_;
    NodeStatistic<CudaUctStatistic>::collected_cost_ = o.collected_cost_;
    NodeStatistic<CudaUctStatistic>::collected_action_transition_counts_ = o.collected_action_transition_counts_;
    num_actions_ = o.num_actions_;
    agent_idx_ = o.agent_idx_;
    random_generator_ = o.random_generator_;
    copy_fields(o);
    return *this;
  }
  ~CudaUctStatistic() { forget(this); }

  // uct_statistic.h:61-76 (host side: selection is not part of the rollout-and-backup path)
  template <class S>
  ActionIdx choose_next_action(const S& state) {
    flush();
    if (require_progressive_widening_total()) {
      std::uniform_int_distribution<ActionIdx> random_action_selection(0, unexpanded_actions_.size() - 1);
      ActionIdx array_idx = random_action_selection(random_generator_);
      ActionIdx selected_action = unexpanded_actions_[array_idx];
      unexpanded_actions_.erase(unexpanded_actions_.begin() + array_idx);
      ucb_statistics_[selected_action] = UcbPair();
      return selected_action;
    }
    std::unordered_map<ActionIdx, double> values;
    return calculate_ucb_and_max_action(ucb_statistics_, values);
  }

  ActionIdx get_best_action() const {  // :78-89
    flush();
    double temp = ucb_statistics_.begin()->second.action_value_;
    ActionIdx best = ucb_statistics_.begin()->first;
    for (auto it = ucb_statistics_.begin(); it != ucb_statistics_.end(); ++it) {
      if (it->second.action_value_ > temp) { temp = it->second.action_value_; best = it->first; }
    }
    return best;
  }

  Policy get_policy() const {  // :91-98
    flush();
    Policy policy;
    for (auto it = ucb_statistics_.begin(); it != ucb_statistics_.end(); ++it) policy[it->first] = it->second.action_value_;
    return policy;
  }

  // recorded; executed by mamcts_backup_uct at the next read (see flush())
  void update_from_heuristic(const NodeStatistic<CudaUctStatistic>& heuristic_statistic) {
    const CudaUctStatistic& h = heuristic_statistic.impl();
    pending().push_back(Op{this, nullptr, h.value_, 0, 0.0, true});
  }
  void update_statistic(const NodeStatistic<CudaUctStatistic>& changed_child_statistic) {
    const CudaUctStatistic& child = changed_child_statistic.impl();
    pending().push_back(Op{this, &child, 0.0, collected_reward_.first, collected_reward_.second, false});
  }

  void set_heuristic_estimate(const Reward& accum_rewards, const EgoCosts&) { value_ = accum_rewards; }  // :146-154

  // uct_statistic.h:165-184 through mamcts_root_merge (+ the NCCL all-reduce when the engine has a
  // communicator attached)
  void merge_node_statistics(const std::vector<CudaUctStatistic>& statistics) {
    flush();
    const uint32_t T = (uint32_t)statistics.size();
    uint32_t MA = 1;
    for (const auto& s : statistics) MA = std::max<uint32_t>(MA, (uint32_t)s.num_actions_);
    for (const auto& s : statistics)
      for (const auto& kv : s.ucb_statistics_) MA = std::max<uint32_t>(MA, (uint32_t)kv.first + 1);
    std::vector<double> value(T, 0.0), latest(T, 0.0), q((size_t)T * MA, 0.0);
    std::vector<uint32_t> visits(T, 0), n((size_t)T * MA, 0), slots(T);
    for (uint32_t t = 0; t < T; ++t) {
      slots[t] = t;
      visits[t] = statistics[t].total_node_visits_;
      for (const auto& kv : statistics[t].ucb_statistics_) {
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
    ucb_statistics_.clear();
    for (uint32_t a = 0; a < MA; ++a)
      if (occ[a]) ucb_statistics_[a] = UcbPair((unsigned)mn[a], mq[a], 0.0);
    total_node_visits_ += (unsigned)mvis;
    value_ = mv;
  }

  std::string print_node_information() const {
    flush();
    std::stringstream ss;
    ss << std::setprecision(2) << "V=" << value_ << ", N=" << total_node_visits_;
    return ss.str();
  }
  std::string print_edge_information(const ActionIdx& action) const {
    flush();
    std::stringstream ss;
    auto it = ucb_statistics_.find(action);
    if (it != ucb_statistics_.end())
      ss << std::setprecision(2) << "a=" << int(action) << ", N=" << it->second.action_count_ << ", V=" << it->second.action_value_;
    return ss.str();
  }
  std::string sprintf() const {
    flush();
    std::stringstream ss;
    for (const auto& u : ucb_statistics_) ss << "a=" << u.first << ", q=" << u.second.action_value_ << ", n=" << u.second.action_count_ << "|";
    return ss.str();
  }

  const UcbStatistics& get_ucb_statistics() const { flush(); return ucb_statistics_; }
  double get_value() const { flush(); return value_; }
  double get_latest_return() const { flush(); return latest_return_; }
  unsigned get_total_node_visits() const { flush(); return total_node_visits_; }
  Reward get_reward_lower_bound() const { return lower_bound; }
  Reward get_reward_upper_bound() const { return upper_bound; }
  inline unsigned int num_expanded_actions() const { return ucb_statistics_.size(); }

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
      for (const auto& kv : s->ucb_statistics_) m = std::max<uint32_t>(m, (uint32_t)kv.first + 1);
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
      value[slot] = s->value_; latest[slot] = s->latest_return_; visits[slot] = s->total_node_visits_;
      for (const auto& kv : s->ucb_statistics_) {
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
    mp.discount_factor = work.front().self->k_discount_factor;
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
      s->value_ = value[slot];
      s->latest_return_ = latest[slot];
      s->total_node_visits_ = visits[slot];
    };
    for (uint32_t k = 0; k < A; ++k) {
      for (uint32_t pth = 0; pth < P; ++pth) {
        const Chain& c = chains[k][pth];
        if (c.leaf) take((size_t)base[pth] * A + k, c.leaf->self);
        for (uint32_t e = 0; e < len[pth]; ++e) {
          const Op* op = c.edges[e];
          const size_t slot = (size_t)(base[pth] + e + 1) * A + k;
          take(slot, op->self);
          UcbPair& pair = op->self->ucb_statistics_[op->action];
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

  void copy_fields(const CudaUctStatistic& o) {
    value_ = o.value_; latest_return_ = o.latest_return_; ucb_statistics_ = o.ucb_statistics_;
    total_node_visits_ = o.total_node_visits_; unexpanded_actions_ = o.unexpanded_actions_;
    upper_bound = o.upper_bound; lower_bound = o.lower_bound; k_discount_factor = o.k_discount_factor;
    exploration_constant = o.exploration_constant; progressive_widening_k = o.progressive_widening_k;
    progressive_widening_alpha = o.progressive_widening_alpha; use_bound_estimation_ = o.use_bound_estimation_;
  }

  std::pair<Reward, Reward> calculate_bounds_based_on_stat(const UcbStatistics& ucb_stats) const {  // :122-132
    const auto mm = std::minmax_element(ucb_stats.begin(), ucb_stats.end(), [](const auto& a, const auto& b) {
      return a.second.action_value_ < b.second.action_value_; });
    std::pair<Reward, Reward> bounds{mm.first->second.action_value_, mm.second->second.action_value_};
    if (bounds.first == bounds.second) {
      if (bounds.second != 0.0) return std::make_pair(Reward(0.0), Reward(bounds.second));
      return std::make_pair(Reward(0.0), Reward(1.0));
    }
    return bounds;
  }

  ActionIdx calculate_ucb_and_max_action(const UcbStatistics& ucb_statistics,
                                         std::unordered_map<ActionIdx, double>& values) const {  // :223-244
    ActionIdx maximizing_action = 0;
    double max_value = std::numeric_limits<double>::min();
    std::pair<Reward, Reward> bounds(lower_bound, upper_bound);
    if (use_bound_estimation_) bounds = calculate_bounds_based_on_stat(ucb_statistics);
    for (const auto& ucb_pair : ucb_statistics) {
      double action_value_normalized = (ucb_pair.second.action_value_ - bounds.first) / (bounds.second - bounds.first);
      MCTS_EXPECT_TRUE(action_value_normalized >= 0);
      MCTS_EXPECT_TRUE(action_value_normalized <= 1);
      values[ucb_pair.first] = action_value_normalized +
          2 * exploration_constant * sqrt((2 * log(total_node_visits_)) / (ucb_pair.second.action_count_));
      if (values[ucb_pair.first] > max_value) { max_value = values[ucb_pair.first]; maximizing_action = ucb_pair.first; }
    }
    return maximizing_action;
  }

  inline bool require_progressive_widening_total() const {  // :256-262
    const auto widening_term = progressive_widening_k * std::pow(total_node_visits_, progressive_widening_alpha);
    return num_expanded_actions() <= widening_term && num_expanded_actions() < num_actions_;
  }

  double value_;
  double latest_return_;
  UcbStatistics ucb_statistics_;
  unsigned int total_node_visits_;
  std::vector<ActionIdx> unexpanded_actions_;
  double upper_bound, lower_bound, k_discount_factor, exploration_constant;
  double progressive_widening_k, progressive_widening_alpha;
  bool use_bound_estimation_;
};

}  // namespace mcts
#endif
