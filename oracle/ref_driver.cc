// oracle/ref_driver.cc - drives the UNMODIFIED reference (juloberno/mamcts headers, included from
// where they lie under $MAMCTS_REFERENCE, default /root/reference) behind a small C interface so
// that tests can pin oracle/mamcts_oracle.c and the CUDA engine against the real thing, and so
// that bench.py can time the reference's own CPU path.
//
// TEST INFRASTRUCTURE ONLY. Built by oracle/Makefile into oracle/_ref/libmamcts_ref.so (git-ignored,
// travels to the GPU box). No reference source is copied into this repository: this file only
// instantiates the reference's templates and reads their results.
//
// Wall-clock cut-offs (MAX_SEARCH_TIME, random_heuristic.MAX_SEARCH_TIME) are set huge so the
// reference is deterministic.

#define UNIT_TESTING  // mcts/common.h:16-20: MCTS_TEST -> `friend class UctTest`
#include <chrono>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <numeric>
#include <sstream>
#include <thread>
#include <vector>

#include "mcts/mcts.h"
#include "mcts/heuristics/random_heuristic.h"
#include "mcts/heuristics/action_value_random_heuristic.h"
#include "mcts/statistics/uct_statistic.h"
#include "mcts/hypothesis/hypothesis_statistic.h"
#include "environments/crossing_state.h"
#include "test/uct/simple_state.h"

#include "mamcts_b200.h"
#include "mamcts_oracle.h"

using namespace mcts;

namespace {

MctsParameters to_ref_params(const mamcts_params& p) {
  MctsParameters r = mcts_default_parameters();
  r.DISCOUNT_FACTOR = p.discount_factor;
  r.RANDOM_SEED = p.random_seed;
  r.MAX_NUMBER_OF_ITERATIONS = p.max_number_of_iterations;
  r.MAX_SEARCH_TIME = 1000000000u;  // F2
  r.MAX_SEARCH_DEPTH = p.max_search_depth;
  r.MAX_NUMBER_OF_NODES = p.max_number_of_nodes;
  r.USE_BOUND_ESTIMATION = p.use_bound_estimation != 0;
  r.NUM_PARALLEL_MCTS = p.num_parallel_mcts ? p.num_parallel_mcts : 1;
  r.USE_MULTI_THREADING = false;
  r.random_heuristic.MAX_SEARCH_TIME = 1e9;  // F2
  r.random_heuristic.MAX_NUMBER_OF_ITERATIONS = p.rh_max_number_of_iterations;
  r.uct_statistic.LOWER_BOUND = p.uct_lower_bound;
  r.uct_statistic.UPPER_BOUND = p.uct_upper_bound;
  r.uct_statistic.EXPLORATION_CONSTANT = p.uct_exploration_constant;
  r.uct_statistic.PROGRESSIVE_WIDENING_K = p.uct_pw_k;
  r.uct_statistic.PROGRESSIVE_WIDENING_ALPHA = p.uct_pw_alpha;
  return r;
}

CrossingStateParameters<int> to_ref_crossing(const mamcts_crossing_params& c) {
  CrossingStateParameters<int> r = default_crossing_state_parameters<int>();
  r.NUM_OTHER_AGENTS = c.num_other_agents;
  r.COST_ONLY_COLLISION = c.cost_only_collision != 0;
  r.MAX_VELOCITY_OTHER = c.max_velocity_other;
  r.MIN_VELOCITY_OTHER = c.min_velocity_other;
  r.NUM_OTHER_ACTIONS = c.num_other_actions;
  r.MAX_VELOCITY_EGO = c.max_velocity_ego;
  r.MIN_VELOCITY_EGO = c.min_velocity_ego;
  r.CHAIN_LENGTH = c.chain_length;
  r.EGO_GOAL_POS = c.ego_goal_pos;
  r.REWARD_COLLISION = c.reward_collision;
  r.REWARD_GOAL_REACHED = c.reward_goal_reached;
  r.REWARD_STEP = c.reward_step;
  return r;
}

// ---- recording -----------------------------------------------------------------------------
struct StepRecord {
  std::vector<long long> action;  // JointAction entries reinterpreted as signed
  std::vector<double> rewards;
  std::vector<double> cost;
  bool terminal_after;
};
struct Recorder {
  std::vector<StepRecord> steps;
};
thread_local Recorder* g_recorder = nullptr;

// A State that forwards every concept call to the wrapped reference state and logs execute().
// RandomHeuristic::rollout only touches the state through these methods
// (mcts/heuristics/random_heuristic.h:36-91).
template <class Inner>
class RecState : public StateInterface<RecState<Inner>> {
 public:
  explicit RecState(std::shared_ptr<Inner> inner) : inner_(std::move(inner)) {}
  std::shared_ptr<RecState> clone() const { return std::make_shared<RecState>(inner_->clone()); }
  std::shared_ptr<RecState> execute(const JointAction& ja, std::vector<Reward>& rewards,
                                    EgoCosts& cost) const {
    auto next = inner_->execute(ja, rewards, cost);
    if (g_recorder) {
      StepRecord rec;
      for (auto a : ja) rec.action.push_back(static_cast<long long>(a));
      rec.rewards = rewards;
      rec.cost = cost;
      rec.terminal_after = next->is_terminal();
      g_recorder->steps.push_back(rec);
    }
    return std::make_shared<RecState>(next);
  }
  ActionIdx get_num_actions(AgentIdx a) const { return inner_->get_num_actions(a); }
  bool is_terminal() const { return inner_->is_terminal(); }
  const std::vector<AgentIdx> get_other_agent_idx() const { return inner_->get_other_agent_idx(); }
  const AgentIdx get_ego_agent_idx() const { return inner_->get_ego_agent_idx(); }
  const std::size_t get_num_costs() const { return inner_->get_num_costs(); }
  double get_execution_step_length() const { return inner_->get_execution_step_length(); }
  void choose_random_seed(const unsigned& s) { inner_->choose_random_seed(s); }
  std::string sprintf() const { return inner_->sprintf(); }
  const Inner& inner() const { return *inner_; }

 private:
  std::shared_ptr<Inner> inner_;
};

// The same for a hypothesis state: HypothesisStatistic::choose_next_action reaches the state through
// HypothesisStateInterface (mcts/hypothesis/hypothesis_statistic.h:62-81), so the wrapper forwards
// plan_action_current_hypothesis as well.
template <class Inner>
class RecHypState : public HypothesisStateInterface<RecHypState<Inner>> {
 public:
  RecHypState(std::shared_ptr<Inner> inner, const std::unordered_map<AgentIdx, HypothesisId>& hyp)
      : HypothesisStateInterface<RecHypState<Inner>>(hyp), inner_(std::move(inner)), hyp_(hyp) {}
  std::shared_ptr<RecHypState> clone() const { return std::make_shared<RecHypState>(inner_->clone(), hyp_); }
  std::shared_ptr<RecHypState> execute(const JointAction& ja, std::vector<Reward>& rewards, EgoCosts& cost) const {
    auto next = inner_->execute(ja, rewards, cost);
    if (g_recorder) {
      StepRecord rec;
      for (auto a : ja) rec.action.push_back(static_cast<long long>(a));
      rec.rewards = rewards;
      rec.cost = cost;
      rec.terminal_after = next->is_terminal();
      g_recorder->steps.push_back(rec);
    }
    return std::make_shared<RecHypState>(next, hyp_);
  }
  ActionIdx plan_action_current_hypothesis(const AgentIdx& a) const { return inner_->plan_action_current_hypothesis(a); }
  ActionIdx get_num_actions(AgentIdx a) const { return inner_->get_num_actions(a); }
  bool is_terminal() const { return inner_->is_terminal(); }
  const std::vector<AgentIdx> get_other_agent_idx() const { return inner_->get_other_agent_idx(); }
  const AgentIdx get_ego_agent_idx() const { return inner_->get_ego_agent_idx(); }
  const std::size_t get_num_costs() const { return inner_->get_num_costs(); }
  double get_execution_step_length() const { return inner_->get_execution_step_length(); }
  void choose_random_seed(const unsigned& s) { inner_->choose_random_seed(s); }
  std::string sprintf() const { return inner_->sprintf(); }

 private:
  std::shared_ptr<Inner> inner_;
  const std::unordered_map<AgentIdx, HypothesisId>& hyp_;
};

// Long-lived pieces a CrossingState<int> keeps references to (crossing_state.h:320, hypothesis_state.h:42).
struct CrossingWorld {
  CrossingStateParameters<int> params;
  std::unordered_map<AgentIdx, HypothesisId> hypothesis_map;
  explicit CrossingWorld(const mamcts_crossing_params& c) : params(to_ref_crossing(c)) {
    for (AgentIdx i = 1; i <= c.num_other_agents; ++i) hypothesis_map[i] = 0;
  }
  std::shared_ptr<CrossingState<int>> make(const int32_t* x, const int32_t* last, bool terminal) {
    std::vector<AgentState<int>> others;
    for (unsigned i = 0; i < params.NUM_OTHER_AGENTS; ++i)
      others.emplace_back(x[i + 1], last ? last[i + 1] : 2);
    AgentState<int> ego(x[0], last ? last[0] : 2);
    return std::make_shared<CrossingState<int>>(hypothesis_map, params, others, ego, terminal,
                                                false, false,
                                                std::vector<AgentPolicyCrossingState<int>>());
  }
};

// ---- stochastic twin: the reference's Mcts/StageNode/UctStatistic with a Philox heuristic -----
struct PhiloxCtx {
  orc_env env;
  mamcts_params mp;
  uint64_t seed;
  uint32_t tree_id;
  uint32_t K;
  uint64_t rollout_steps;
};
thread_local PhiloxCtx g_philox;

void pack_state(const CrossingState<int>& s, orc_state* o) {
  std::memset(o, 0, sizeof(*o));
  o->x[0] = s.get_ego_state().x_pos;
  o->last[0] = s.get_ego_state().last_action;
  const auto& oth = s.get_agent_states();
  for (size_t i = 0; i < oth.size(); ++i) {
    o->x[i + 1] = oth[i].x_pos;
    o->last[i + 1] = oth[i].last_action;
  }
  o->terminal = s.is_terminal();
  o->goal = s.ego_goal_reached();
  o->collided = s.ego_collided();
}
void pack_state(const SimpleState& s, orc_state* o) {
  std::memset(o, 0, sizeof(*o));
  o->x[0] = s.get_state_length();
  o->terminal = s.is_terminal();
}

// Heuristic concept (mcts/heuristic.h:24-41) whose rollout body is oracle/mamcts_oracle.c's
// orc_rollout in PHILOX mode, stream (tree_id, node_index*K + k); node_index = number of
// heuristic calls so far (== creation index of the node being evaluated).
class PhiloxHeuristic : public mcts::Heuristic<PhiloxHeuristic> {
 public:
  explicit PhiloxHeuristic(const MctsParameters& p) : mcts::Heuristic<PhiloxHeuristic>(p), calls_(0) {}
  template <class S, class SE, class SO, class H>
  std::pair<SE, std::unordered_map<AgentIdx, SO>> calculate_heuristic_values(
      const std::shared_ptr<StageNode<S, SE, SO, H>>& node) {
    ++calls_;
    PhiloxCtx& c = g_philox;
    orc_state st;
    pack_state(*node->get_state(), &st);
    const uint32_t A = c.env.num_agents;
    std::vector<double> ego(c.K), oth(static_cast<size_t>(c.K) * ORC_MAX_AGENTS);
    for (uint32_t k = 0; k < c.K; ++k) {
      orc_rollout_result rr;
      orc_rollout(&c.env, &c.mp, &st, node->get_depth(), MAMCTS_SRC_PHILOX, nullptr, 0, nullptr,
                  c.seed, c.tree_id, calls_ * c.K + k, nullptr, nullptr, &rr, nullptr, 0);
      c.rollout_steps += rr.steps;
      ego[k] = rr.ego_return;
      for (uint32_t j = 0; j + 1 < A; ++j) oth[static_cast<size_t>(j) * c.K + k] = rr.other_return[j];
    }
    SE ego_h(0, node->get_state()->get_ego_agent_idx(), mcts_parameters_);
    ego_h.set_heuristic_estimate(orc_reduce_mean(ego.data(), c.K), EgoCosts());
    std::unordered_map<AgentIdx, SO> others;
    uint32_t j = 0;
    for (auto ai : node->get_state()->get_other_agent_idx()) {
      SO st_o(0, ai, mcts_parameters_);
      st_o.set_heuristic_estimate(orc_reduce_mean(oth.data() + static_cast<size_t>(j) * c.K, c.K),
                                  EgoCosts());
      others.insert(std::pair<AgentIdx, SO>(ai, st_o));
      ++j;
    }
    return std::pair<SE, std::unordered_map<AgentIdx, SO>>(ego_h, others);
  }
  std::string sprintf() const { return "PhiloxHeuristic"; }

 private:
  uint32_t calls_;
};

}  // namespace

// White-box reader: the reference grants `friend class UctTest` under UNIT_TESTING.
class mcts::UctTest {
 public:
  static void read_hyp(const HypothesisStatistic& s, double* value, double* latest, uint32_t* visits,
                       uint32_t* expanded) {
    *value = s.ego_cost_value_;
    *latest = s.latest_ego_cost_;
    *visits = s.total_node_visits_;
    *expanded = s.num_expanded_actions_;
  }
  static void read_stat(const UctStatistic& s, double* value, double* latest, uint32_t* visits,
                        double* q, int64_t* n, uint32_t max_actions) {
    *value = s.value_;
    *latest = s.latest_return_;
    *visits = s.total_node_visits_;
    for (uint32_t a = 0; a < max_actions; ++a) { q[a] = 0.0; n[a] = -1; }
    for (const auto& kv : s.ucb_statistics_) {
      if (kv.first < max_actions) {
        q[kv.first] = kv.second.action_value_;
        n[kv.first] = kv.second.action_count_;
      }
    }
  }
  static void force_visits(UctStatistic& s, unsigned v) { s.total_node_visits_ = v; }

  // canonical dump of a hypothesis-MCTS tree (record layout: include/mamcts_b200.h, mamcts_hyp_search_create)
  template <class S, class H>
  static uint32_t dump_hyp(const std::shared_ptr<StageNode<S, UctStatistic, HypothesisStatistic, H>>& node, double* rec,
                           uint32_t cap, uint32_t count, uint32_t NE, uint32_t MA, uint32_t NH, int min_v) {
    const uint32_t A = node->state_->get_num_agents();
    const uint32_t J = A - 1;
    const uint32_t M = NE > MA ? NE : MA;
    const uint32_t per_other = 6 + NH * (2 + MA) + 2 * NH * MA;
    const uint32_t stride = 4 + (3 + 2 * NE) + J * per_other;
    auto code_of = [&](const JointAction& ja) {
      double code = static_cast<double>(ja[0]), mul = M;
      for (uint32_t a = 1; a < A; ++a) { code += mul * static_cast<double>(aconv<int>(ja[a]) - min_v); mul *= M; }
      return code;
    };
    if (count < cap) {
      double* r = rec + static_cast<size_t>(count) * stride;
      for (uint32_t i = 0; i < stride; ++i) r[i] = 0.0;
      r[0] = node->depth_;
      r[1] = node->is_root() ? -1.0 : code_of(node->joint_action_);
      r[2] = node->visit_count_;
      r[3] = static_cast<double>(node->children_.size());
      {
        const UctStatistic& st = node->ego_int_node_;
        double* e = r + 4;
        e[0] = st.total_node_visits_; e[1] = st.value_; e[2] = st.latest_return_;
        for (uint32_t k = 0; k < NE; ++k) { e[3 + k] = -1.0; e[3 + NE + k] = 0.0; }
        for (const auto& kv : st.ucb_statistics_)
          if (kv.first < NE) { e[3 + kv.first] = kv.second.action_count_; e[3 + NE + kv.first] = kv.second.action_value_; }
      }
      for (uint32_t j = 0; j < J; ++j) {
        const HypothesisStatistic& st = node->other_int_nodes_[j];
        double* o = r + 4 + (3 + 2 * NE) + j * per_other;
        o[0] = st.total_node_visits_; o[1] = st.ego_cost_value_; o[2] = st.latest_ego_cost_; o[3] = st.num_expanded_actions_;
        auto ns = st.total_node_visits_hypothesis_.find(HYPOTHESIS_ID_NOT_SET);
        o[4] = ns == st.total_node_visits_hypothesis_.end() ? 0.0 : ns->second;
        o[5] = 0.0;
        for (uint32_t h = 0; h < NH; ++h) {
          double* oh = o + 6 + h * (2 + MA);
          auto vit = st.total_node_visits_hypothesis_.find(h);
          oh[0] = vit == st.total_node_visits_hypothesis_.end() ? -1.0 : vit->second;
          auto uit = st.ucb_statistics_.find(h);
          oh[1] = uit == st.ucb_statistics_.end() ? -1.0 : static_cast<double>(uit->second.size());
          for (uint32_t k = 0; k < MA; ++k) oh[2 + k] = -1.0;
          double* oc = o + 6 + NH * (2 + MA) + 2 * h * MA;
          for (uint32_t k = 0; k < MA; ++k) { oc[2 * k] = -1.0; oc[2 * k + 1] = 0.0; }
          if (uit != st.ucb_statistics_.end()) {
            uint32_t pos = 0;
            for (const auto& kv : uit->second) {   // the container's own iteration order
              const int a = aconv<int>(kv.first) - min_v;
              if (pos < MA) oh[2 + pos] = a;
              ++pos;
              if (a >= 0 && static_cast<uint32_t>(a) < MA) { oc[2 * a] = kv.second.action_count_; oc[2 * a + 1] = kv.second.action_ego_cost_; }
            }
          }
        }
      }
    }
    ++count;
    std::map<double, std::shared_ptr<StageNode<S, UctStatistic, HypothesisStatistic, H>>> sorted;
    for (const auto& kv : node->children_) sorted[code_of(kv.first)] = kv.second;
    for (const auto& kv : sorted) count = dump_hyp<S, H>(kv.second, rec, cap, count, NE, MA, NH, min_v);
    return count;
  }
  template <class S, class H>
  static std::shared_ptr<StageNode<S, UctStatistic, HypothesisStatistic, H>> root_of_hyp(
      Mcts<S, UctStatistic, HypothesisStatistic, H>& m) {
    return m.root_;
  }

  template <class S, class H>
  static uint32_t dump(const std::shared_ptr<StageNode<S, UctStatistic, UctStatistic, H>>& node,
                       double* rec, uint32_t cap, uint32_t count, uint32_t max_actions) {
    const uint32_t A = node->state_->get_num_agents();
    const uint32_t stride = 4 + A * (3 + 2 * max_actions);
    auto code_of = [&](const JointAction& ja) {
      double code = 0.0, mul = 1.0;
      for (uint32_t a = 0; a < A; ++a) { code += mul * static_cast<double>(ja[a]); mul *= max_actions; }
      return code;
    };
    if (count < cap) {
      double* r = rec + static_cast<size_t>(count) * stride;
      r[0] = node->depth_;
      r[1] = node->is_root() ? -1.0 : code_of(node->joint_action_);
      r[2] = node->visit_count_;
      r[3] = static_cast<double>(node->children_.size());
      auto emit = [&](const UctStatistic& st, uint32_t a) {
        double* s = r + 4 + a * (3 + 2 * max_actions);
        double v, l; uint32_t nv;
        std::vector<double> q(max_actions);
        std::vector<int64_t> n(max_actions);
        read_stat(st, &v, &l, &nv, q.data(), n.data(), max_actions);
        s[0] = nv; s[1] = v; s[2] = l;
        for (uint32_t k = 0; k < max_actions; ++k) {
          s[3 + k] = static_cast<double>(n[k]);
          s[3 + max_actions + k] = q[k];
        }
      };
      emit(node->ego_int_node_, 0);
      for (uint32_t j = 0; j < node->other_int_nodes_.size(); ++j) emit(node->other_int_nodes_[j], j + 1);
    }
    ++count;
    std::map<double, std::shared_ptr<StageNode<S, UctStatistic, UctStatistic, H>>> sorted;
    for (const auto& kv : node->children_) sorted[code_of(kv.first)] = kv.second;
    for (const auto& kv : sorted) count = dump<S, H>(kv.second, rec, cap, count, max_actions);
    return count;
  }

  template <class S, class H>
  static std::shared_ptr<StageNode<S, UctStatistic, UctStatistic, H>> root_of(
      Mcts<S, UctStatistic, UctStatistic, H>& m) {
    return m.root_;
  }
};

namespace {

struct SearchOut {
  uint32_t num_nodes;
  uint32_t num_iterations;
  uint32_t best_action;
};

template <class S, class H>
void run_search_and_dump(const S& state, const MctsParameters& rp, uint32_t num_agents,
                         uint32_t max_actions, double* root_value, double* root_latest,
                         uint32_t* root_visits, double* root_q, int64_t* root_n,
                         uint32_t* counters /*[3]: nodes, iterations, best*/, double* tree_records,
                         uint32_t tree_capacity, uint32_t* tree_count) {
  Mcts<S, UctStatistic, UctStatistic, H> mcts(rp);
  mcts.search(state);
  const auto& root = mcts.get_root();
  UctTest::read_stat(root.get_ego_int_node(), &root_value[0], &root_latest[0], &root_visits[0],
                     root_q, root_n, max_actions);
  for (uint32_t j = 0; j + 1 < num_agents; ++j) {
    UctTest::read_stat(root.get_other_int_nodes()[j], &root_value[j + 1], &root_latest[j + 1],
                       &root_visits[j + 1], root_q + (j + 1) * max_actions,
                       root_n + (j + 1) * max_actions, max_actions);
  }
  counters[0] = mcts.numNodes();
  counters[1] = mcts.numIterations();
  counters[2] = static_cast<uint32_t>(mcts.returnBestAction());
  if (tree_records && tree_count && rp.NUM_PARALLEL_MCTS <= 1) {
    *tree_count = UctTest::dump<S, H>(UctTest::root_of<S, H>(mcts), tree_records, tree_capacity, 0,
                                      max_actions);
  } else if (tree_count) {
    *tree_count = 0;
  }
}

}  // namespace

extern "C" {

// The widening order of a real UctStatistic with `num_actions` actions
// (mcts/statistics/uct_statistic.h:61-69): total_node_visits_ is forced high so that
// require_progressive_widening_total() stays true until every action is expanded.
int ref_uct_expand_order(const mamcts_params* mp, uint32_t num_actions, uint32_t* order) {
  MctsParameters rp = to_ref_params(*mp);
  UctStatistic stat(num_actions, 0, rp);
  UctTest::force_visits(stat, 1000000000u);
  rp.uct_statistic.PROGRESSIVE_WIDENING_K = 1e9;
  UctStatistic wide(num_actions, 0, rp);
  UctTest::force_visits(wide, 1000000000u);
  SimpleState dummy(0);
  for (uint32_t k = 0; k < num_actions; ++k) order[k] = static_cast<uint32_t>(wide.choose_next_action(dummy));
  return 0;
}

// RandomHeuristic::rollout<RecState<CrossingState<int>>, UctStatistic, UctStatistic, RandomHeuristic>
// from n start states (SoA x[a*n+i]); records every execute().
//   rec_actions: int8 [n][rec_stride][A] (ego index, other bit-cast value), rec_rewards [n][rec_stride],
//   rec_steps[n].
int ref_rollout_crossing(const mamcts_params* mp, const mamcts_crossing_params* cp, uint32_t n,
                         const int32_t* x, const int32_t* last, const uint8_t* flags,
                         const uint32_t* depth, double* ego_ret, double* other_ret, double* cost0,
                         double* step_len, uint32_t* steps, int32_t* final_x, int32_t* final_last,
                         uint8_t* final_flags, int8_t* rec_actions, double* rec_rewards,
                         uint32_t rec_stride) {
  const MctsParameters rp = to_ref_params(*mp);
  CrossingWorld world(*cp);
  const uint32_t A = cp->num_other_agents + 1;
  using RS = RecState<CrossingState<int>>;
  for (uint32_t i = 0; i < n; ++i) {
    std::vector<int32_t> xi(A), li(A);
    for (uint32_t a = 0; a < A; ++a) { xi[a] = x[a * n + i]; li[a] = last ? last[a * n + i] : 2; }
    auto inner = world.make(xi.data(), li.data(), flags ? (flags[i] & 1) : false);
    auto start = std::make_shared<RS>(inner);
    Recorder rec;
    g_recorder = &rec;
    auto result = RandomHeuristic::rollout<RS, UctStatistic, UctStatistic, RandomHeuristic>(
        start, rp, depth ? depth[i] : 0);
    g_recorder = nullptr;
    ego_ret[i] = std::get<0>(result);
    const auto& others = std::get<1>(result);
    for (uint32_t j = 0; j + 1 < A; ++j) other_ret[j * n + i] = others.at(j + 1);
    const auto& costs = std::get<2>(result);
    cost0[i] = costs.empty() ? 0.0 : costs[0];
    step_len[i] = std::get<3>(result);
    steps[i] = static_cast<uint32_t>(rec.steps.size());
    // replay the recorded actions on the plain reference state to obtain the final state
    std::shared_ptr<CrossingState<int>> cur = inner;
    for (size_t s = 0; s < rec.steps.size(); ++s) {
      JointAction ja(A);
      for (uint32_t a = 0; a < A; ++a) ja[a] = static_cast<ActionIdx>(rec.steps[s].action[a]);
      std::vector<Reward> r; EgoCosts c;
      cur = cur->execute(ja, r, c);
      if (s < rec_stride) {
        if (rec_actions) {
          rec_actions[(static_cast<size_t>(i) * rec_stride + s) * A + 0] = static_cast<int8_t>(ja[0]);
          for (uint32_t a = 1; a < A; ++a)
            rec_actions[(static_cast<size_t>(i) * rec_stride + s) * A + a] =
                static_cast<int8_t>(aconv<int>(ja[a]));
        }
        if (rec_rewards) rec_rewards[static_cast<size_t>(i) * rec_stride + s] = rec.steps[s].rewards[0];
      }
    }
    orc_state fs;
    pack_state(*cur, &fs);
    for (uint32_t a = 0; a < A; ++a) {
      if (final_x) final_x[a * n + i] = fs.x[a];
      if (final_last) final_last[a * n + i] = fs.last[a];
    }
    if (final_flags) final_flags[i] = (fs.terminal ? 1 : 0) | (fs.goal ? 2 : 0) | (fs.collided ? 4 : 0);
  }
  return 0;
}

// A minimal hypothesis state for driving HypothesisStatistic objects directly (like the reference's
// test/hypothesis/hypothesis_statistic_test_state.h): the current hypothesis and the action the
// hypothesis "plans" are whatever the driver sets before the call.
class DrivenHypState : public HypothesisStateInterface<DrivenHypState> {
 public:
  explicit DrivenHypState(const std::unordered_map<AgentIdx, HypothesisId>& hyp)
      : HypothesisStateInterface<DrivenHypState>(hyp) {}
  ActionIdx plan_action_current_hypothesis(const AgentIdx&) const { return planned; }
  ActionIdx planned = 0;
};

// HypothesisStatistic back-propagation on the arrays of mamcts_backup_hypothesis, executed by REAL
// HypothesisStatistic objects (one per (node, other agent)) in the order Mcts::iterate would call
// them (mcts.h:309-329): leaf update_from_heuristic, then per ancestor choose_next_action (which fixes
// the hypothesis of the iteration, hypothesis_statistic.h:62-66), collect, update_statistic.
// The statistics must start empty (a fresh tree); results are written in mamcts_hyp_stats layout.
int ref_backup_hypothesis(const mamcts_params* mp, uint32_t num_others, uint32_t n_nodes, uint32_t n_paths,
                          const uint32_t* leaf_node, const uint8_t* leaf_has, const double* leaf_cost,
                          uint32_t C, const uint32_t* path_offset, const uint32_t* path_len,
                          const uint32_t* edge_node, const uint8_t* edge_hyp, const uint8_t* edge_action,
                          const double* edge_cost, int32_t chance, uint32_t rounds_paths_per_round,
                          mamcts_hyp_stats* out) {
  MctsParameters rp = to_ref_params(*mp);
  rp.cost_constrained_statistic.USE_CHANCE_CONSTRAINED_UPDATES = {chance != 0};
  const uint32_t J = num_others, H = out->num_hypotheses, MA = out->max_actions;
  (void)rounds_paths_per_round;
  std::vector<std::vector<HypothesisStatistic>> stats(n_nodes);
  for (uint32_t n = 0; n < n_nodes; ++n)
    for (uint32_t j = 0; j < J; ++j) stats[n].emplace_back(MA, j + 1, rp);
  std::unordered_map<AgentIdx, HypothesisId> cur;
  for (uint32_t j = 0; j < J; ++j) cur[j + 1] = 0;
  DrivenHypState state(cur);
  for (uint32_t p = 0; p < n_paths; ++p) {
    for (uint32_t j = 0; j < J; ++j) {
      HypothesisStatistic* child = &stats[leaf_node[p]][j];
      if (leaf_has[p]) {
        HypothesisStatistic heuristic(0, j + 1, rp);
        heuristic.set_heuristic_estimate(0.0, EgoCosts(leaf_cost + (size_t)p * C, leaf_cost + (size_t)p * C + C));
        child->update_from_heuristic(heuristic);
      }
      for (uint32_t e = 0; e < path_len[p]; ++e) {
        const size_t idx = (size_t)path_offset[p] + e;
        HypothesisStatistic& st = stats[edge_node[idx]][j];
        cur[j + 1] = edge_hyp[idx * J + j];
        state.planned = edge_action[idx * J + j];
        st.choose_next_action(state);
        st.collect(0.0, EgoCosts(edge_cost + idx * C, edge_cost + idx * C + C), edge_action[idx * J + j], {0, 0});
        st.update_statistic(*child);
        child = &st;
      }
    }
  }
  for (uint32_t n = 0; n < n_nodes; ++n) {
    for (uint32_t j = 0; j < J; ++j) {
      const size_t slot = (size_t)n * J + j;
      const HypothesisStatistic& st = stats[n][j];
      UctTest::read_hyp(st, &out->ego_cost_value[slot], &out->latest_ego_cost[slot], &out->total_visits[slot],
                        &out->num_expanded[slot]);
      for (const auto& kv : st.get_total_node_visits())
        out->visits_hypothesis[slot * (H + 1) + (kv.first < H ? kv.first : H)] = kv.second;
      for (const auto& hv : st.get_ucb_statistics())
        for (const auto& av : hv.second) {
          out->action_count[(slot * H + hv.first) * MA + av.first] = av.second.action_count_;
          out->action_cost[(slot * H + hv.first) * MA + av.first] = av.second.action_ego_cost_;
        }
    }
  }
  return 0;
}

// RandomHeuristic::rollout on CrossingState<float> (UCT statistics for all agents: constant index-valued
// actions, F1 / F5), plus optional explicit joint-action sequences executed step by step (act_seq != null:
// per rollout seq_len[i] joint actions: ego index, others' velocities) so that arbitrary float velocities
// are covered. Records the executed joint actions as floats for REPLAY.
int ref_rollout_crossing_f32(const mamcts_params* mp, const mamcts_crossing_params_f32* cp, uint32_t n,
                             const float* x, const float* last, const uint32_t* depth, double* ego_ret,
                             double* cost0, double* step_len, uint32_t* steps, float* final_x, float* final_last,
                             uint8_t* final_flags, float* rec_actions, uint32_t rec_stride) {
  const MctsParameters rp = to_ref_params(*mp);
  CrossingStateParameters<float> params = default_crossing_state_parameters<float>();
  params.NUM_OTHER_AGENTS = cp->num_other_agents;
  params.NUM_OTHER_ACTIONS = cp->num_other_actions;
  params.COST_ONLY_COLLISION = cp->cost_only_collision != 0;
  params.MIN_VELOCITY_EGO = cp->min_velocity_ego; params.MAX_VELOCITY_EGO = cp->max_velocity_ego;
  params.MIN_VELOCITY_OTHER = cp->min_velocity_other; params.MAX_VELOCITY_OTHER = cp->max_velocity_other;
  params.CHAIN_LENGTH = cp->chain_length; params.EGO_GOAL_POS = cp->ego_goal_pos;
  params.REWARD_COLLISION = cp->reward_collision; params.REWARD_GOAL_REACHED = cp->reward_goal_reached;
  params.REWARD_STEP = cp->reward_step;
  const uint32_t A = cp->num_other_agents + 1;
  std::unordered_map<AgentIdx, HypothesisId> hyp_map;
  for (AgentIdx j = 1; j < A; ++j) hyp_map[j] = 0;
  using RS = RecState<CrossingState<float>>;
  for (uint32_t i = 0; i < n; ++i) {
    std::vector<AgentState<float>> others;
    for (uint32_t a = 1; a < A; ++a) others.emplace_back(x[a * n + i], last ? last[a * n + i] : 2.0f);
    AgentState<float> ego(x[i], last ? last[i] : 2.0f);
    auto inner = std::make_shared<CrossingState<float>>(hyp_map, params, others, ego, false, false, false,
                                                        std::vector<AgentPolicyCrossingState<float>>());
    auto start = std::make_shared<RS>(inner);
    Recorder rec;
    g_recorder = &rec;
    auto result = RandomHeuristic::rollout<RS, UctStatistic, UctStatistic, RandomHeuristic>(start, rp, depth ? depth[i] : 0);
    g_recorder = nullptr;
    ego_ret[i] = std::get<0>(result);
    const auto& costs = std::get<2>(result);
    cost0[i] = costs.empty() ? 0.0 : costs[0];
    step_len[i] = std::get<3>(result);
    steps[i] = static_cast<uint32_t>(rec.steps.size());
    std::shared_ptr<CrossingState<float>> cur = inner;
    for (size_t s = 0; s < rec.steps.size(); ++s) {
      JointAction ja(A);
      for (uint32_t a = 0; a < A; ++a) ja[a] = static_cast<ActionIdx>(rec.steps[s].action[a]);
      std::vector<Reward> r; EgoCosts c;
      cur = cur->execute(ja, r, c);
      if (s < rec_stride && rec_actions) {
        rec_actions[(static_cast<size_t>(i) * rec_stride + s) * A + 0] = static_cast<float>(ja[0]);
        for (uint32_t a = 1; a < A; ++a)
          rec_actions[(static_cast<size_t>(i) * rec_stride + s) * A + a] = aconv<float>(ja[a]);
      }
    }
    for (uint32_t a = 0; a < A; ++a) {
      const AgentState<float>& st = a == 0 ? cur->get_ego_state() : cur->get_agent_states()[a - 1];
      if (final_x) final_x[a * n + i] = st.x_pos;
      if (final_last) final_last[a * n + i] = st.last_action;
    }
    if (final_flags) final_flags[i] = (cur->is_terminal() ? 1 : 0) | (cur->ego_goal_reached() ? 2 : 0) | (cur->ego_collided() ? 4 : 0);
  }
  return 0;
}

// AgentPolicyCrossingState<float>::calculate_action (environments/crossing_state_agent_policy.h:35-55).
int ref_policy_calculate_action_f32(const mamcts_crossing_params_f32* cp, uint32_t n, const float* agent_x,
                                    const float* agent_last, const float* ego_x, const float* ego_last,
                                    const float* gap, float* out) {
  CrossingStateParameters<float> params = default_crossing_state_parameters<float>();
  params.MIN_VELOCITY_OTHER = cp->min_velocity_other; params.MAX_VELOCITY_OTHER = cp->max_velocity_other;
  params.CHAIN_LENGTH = cp->chain_length;
  AgentPolicyCrossingState<float> policy({0.0f, 1.0f}, params);
  for (uint32_t i = 0; i < n; ++i)
    out[i] = policy.calculate_action(AgentState<float>(agent_x[i], agent_last[i]),
                                     AgentState<float>(ego_x[i], ego_last[i]), gap[i]);
  return 0;
}

// RandomHeuristic::rollout<RecHypState<CrossingState<float>>, UctStatistic, HypothesisStatistic, RandomHeuristic>:
// the float-domain hypothesis rollout (other agents act by AgentPolicyCrossingState<float>::act, agent_policy.h:
// 78-84, with the reference's own uniform_real_distribution draws); records the joint actions for REPLAY.
int ref_rollout_crossing_f32_hypothesis(const mamcts_params* mp, const mamcts_crossing_params_f32* cp, uint32_t n,
                                        const float* x, const float* last, const uint32_t* depth, uint32_t n_hyp,
                                        const float* hyp_lo, const float* hyp_hi, const uint32_t* cur_hyp,
                                        double* ego_ret, double* cost0, double* step_len, uint32_t* steps,
                                        float* final_x, float* final_last, uint8_t* final_flags, float* rec_actions,
                                        uint32_t rec_stride) {
  const MctsParameters rp = to_ref_params(*mp);
  CrossingStateParameters<float> params = default_crossing_state_parameters<float>();
  params.NUM_OTHER_AGENTS = cp->num_other_agents;
  params.NUM_OTHER_ACTIONS = cp->num_other_actions;
  params.COST_ONLY_COLLISION = cp->cost_only_collision != 0;
  params.MIN_VELOCITY_EGO = cp->min_velocity_ego; params.MAX_VELOCITY_EGO = cp->max_velocity_ego;
  params.MIN_VELOCITY_OTHER = cp->min_velocity_other; params.MAX_VELOCITY_OTHER = cp->max_velocity_other;
  params.CHAIN_LENGTH = cp->chain_length; params.EGO_GOAL_POS = cp->ego_goal_pos;
  params.REWARD_COLLISION = cp->reward_collision; params.REWARD_GOAL_REACHED = cp->reward_goal_reached;
  params.REWARD_STEP = cp->reward_step;
  const uint32_t A = cp->num_other_agents + 1;
  using RS = RecHypState<CrossingState<float>>;
  for (uint32_t i = 0; i < n; ++i) {
    std::unordered_map<AgentIdx, HypothesisId> hyp_map;
    for (AgentIdx j = 1; j < A; ++j) hyp_map[j] = cur_hyp[(j - 1) * n + i];
    std::vector<AgentPolicyCrossingState<float>> policies;
    for (uint32_t h = 0; h < n_hyp; ++h) policies.emplace_back(std::make_pair(hyp_lo[h], hyp_hi[h]), params);
    std::vector<AgentState<float>> others;
    for (uint32_t a = 1; a < A; ++a) others.emplace_back(x[a * n + i], last ? last[a * n + i] : 2.0f);
    AgentState<float> ego(x[i], last ? last[i] : 2.0f);
    auto inner = std::make_shared<CrossingState<float>>(hyp_map, params, others, ego, false, false, false, policies);
    auto start = std::make_shared<RS>(inner, hyp_map);
    Recorder rec;
    g_recorder = &rec;
    auto result = RandomHeuristic::rollout<RS, UctStatistic, HypothesisStatistic, RandomHeuristic>(start, rp, depth ? depth[i] : 0);
    g_recorder = nullptr;
    ego_ret[i] = std::get<0>(result);
    const auto& costs = std::get<2>(result);
    cost0[i] = costs.empty() ? 0.0 : costs[0];
    step_len[i] = std::get<3>(result);
    steps[i] = static_cast<uint32_t>(rec.steps.size());
    std::shared_ptr<CrossingState<float>> cur = std::make_shared<CrossingState<float>>(
        hyp_map, params, others, ego, false, false, false, std::vector<AgentPolicyCrossingState<float>>());
    for (size_t s = 0; s < rec.steps.size(); ++s) {
      JointAction ja(A);
      for (uint32_t a = 0; a < A; ++a) ja[a] = static_cast<ActionIdx>(rec.steps[s].action[a]);
      std::vector<Reward> r; EgoCosts c;
      cur = cur->execute(ja, r, c);
      if (s < rec_stride && rec_actions) {
        rec_actions[(static_cast<size_t>(i) * rec_stride + s) * A + 0] = static_cast<float>(ja[0]);
        for (uint32_t a = 1; a < A; ++a)
          rec_actions[(static_cast<size_t>(i) * rec_stride + s) * A + a] = aconv<float>(ja[a]);
      }
    }
    for (uint32_t a = 0; a < A; ++a) {
      const AgentState<float>& st = a == 0 ? cur->get_ego_state() : cur->get_agent_states()[a - 1];
      if (final_x) final_x[a * n + i] = st.x_pos;
      if (final_last) final_last[a * n + i] = st.last_action;
    }
    if (final_flags) final_flags[i] = (cur->is_terminal() ? 1 : 0) | (cur->ego_goal_reached() ? 2 : 0) | (cur->ego_collided() ? 4 : 0);
  }
  return 0;
}

// CrossingState<float>::execute on n states with explicit joint actions (ego index as float, others'
// velocities as floats, passed through aconv like the reference's own float tests do,
// environments/tests/crossing_state_float_test.cc:155-217).
int ref_crossing_execute_f32(const mamcts_crossing_params_f32* cp, uint32_t n, const float* x, const float* last,
                             const float* action, float* next_x, float* next_last, uint8_t* flags,
                             double* reward, double* cost0) {
  CrossingStateParameters<float> params = default_crossing_state_parameters<float>();
  params.NUM_OTHER_AGENTS = cp->num_other_agents;
  params.NUM_OTHER_ACTIONS = cp->num_other_actions;
  params.COST_ONLY_COLLISION = cp->cost_only_collision != 0;
  params.MIN_VELOCITY_EGO = cp->min_velocity_ego; params.MAX_VELOCITY_EGO = cp->max_velocity_ego;
  params.MIN_VELOCITY_OTHER = cp->min_velocity_other; params.MAX_VELOCITY_OTHER = cp->max_velocity_other;
  params.CHAIN_LENGTH = cp->chain_length; params.EGO_GOAL_POS = cp->ego_goal_pos;
  params.REWARD_COLLISION = cp->reward_collision; params.REWARD_GOAL_REACHED = cp->reward_goal_reached;
  params.REWARD_STEP = cp->reward_step;
  const uint32_t A = cp->num_other_agents + 1;
  std::unordered_map<AgentIdx, HypothesisId> hyp_map;
  for (AgentIdx j = 1; j < A; ++j) hyp_map[j] = 0;
  for (uint32_t i = 0; i < n; ++i) {
    std::vector<AgentState<float>> others;
    for (uint32_t a = 1; a < A; ++a) others.emplace_back(x[a * n + i], last[a * n + i]);
    AgentState<float> ego(x[i], last[i]);
    CrossingState<float> st(hyp_map, params, others, ego, false, false, false, std::vector<AgentPolicyCrossingState<float>>());
    JointAction ja(A);
    ja[0] = static_cast<ActionIdx>(action[i]);
    for (uint32_t a = 1; a < A; ++a) ja[a] = aconv<float>(action[a * n + i]);
    std::vector<Reward> r; EgoCosts c;
    auto nx = st.execute(ja, r, c);
    for (uint32_t a = 0; a < A; ++a) {
      const AgentState<float>& s2 = a == 0 ? nx->get_ego_state() : nx->get_agent_states()[a - 1];
      next_x[a * n + i] = s2.x_pos;
      next_last[a * n + i] = s2.last_action;
    }
    flags[i] = (nx->is_terminal() ? 1 : 0) | (nx->ego_goal_reached() ? 2 : 0) | (nx->ego_collided() ? 4 : 0);
    reward[i] = r[0];
    cost0[i] = c.empty() ? 0.0 : c[0];
  }
  return 0;
}

// AgentPolicyCrossingState<int>::calculate_action (environments/crossing_state_agent_policy.h:35-55)
// on n independent inputs.
int ref_policy_calculate_action(const mamcts_crossing_params* cp, uint32_t n, const int32_t* agent_x,
                                const int32_t* agent_last, const int32_t* ego_x, const int32_t* ego_last,
                                const int32_t* gap, int32_t* out) {
  const CrossingStateParameters<int> params = to_ref_crossing(*cp);
  AgentPolicyCrossingState<int> policy({0, 1}, params);
  for (uint32_t i = 0; i < n; ++i)
    out[i] = policy.calculate_action(AgentState<int>(agent_x[i], agent_last[i]),
                                     AgentState<int>(ego_x[i], ego_last[i]), gap[i]);
  return 0;
}

// RandomHeuristic::rollout with HypothesisStatistic for the other agents (the C4 instantiation:
// Mcts<CrossingState<int>, UctStatistic, HypothesisStatistic, RandomHeuristic>, reference
// environments/crossing_state_episode_runner.h:61): other agent j follows hypothesis
// cur_hyp[j] of the n_hyp hypotheses {[hyp_lo[h], hyp_hi[h]]}. Records the joint actions (ego index,
// other agents' Domain values) for REPLAY and returns the rollout's results and final state.
int ref_rollout_crossing_hypothesis(const mamcts_params* mp, const mamcts_crossing_params* cp, uint32_t n,
                                    const int32_t* x, const int32_t* last, const uint32_t* depth,
                                    uint32_t n_hyp, const int32_t* hyp_lo, const int32_t* hyp_hi,
                                    const uint32_t* cur_hyp /* [(A-1)*n] */, double* ego_ret,
                                    double* other_ret, double* cost0, double* step_len, uint32_t* steps,
                                    int32_t* final_x, int32_t* final_last, uint8_t* final_flags,
                                    int8_t* rec_actions, uint32_t rec_stride) {
  const MctsParameters rp = to_ref_params(*mp);
  const CrossingStateParameters<int> params = to_ref_crossing(*cp);
  const uint32_t A = cp->num_other_agents + 1;
  using RS = RecHypState<CrossingState<int>>;
  for (uint32_t i = 0; i < n; ++i) {
    std::unordered_map<AgentIdx, HypothesisId> hyp_map;
    for (AgentIdx j = 1; j < A; ++j) hyp_map[j] = cur_hyp[(j - 1) * n + i];
    std::vector<AgentPolicyCrossingState<int>> policies;
    for (uint32_t h = 0; h < n_hyp; ++h) policies.emplace_back(std::make_pair(hyp_lo[h], hyp_hi[h]), params);
    std::vector<AgentState<int>> others;
    for (uint32_t a = 1; a < A; ++a) others.emplace_back(x[a * n + i], last ? last[a * n + i] : 2);
    AgentState<int> ego(x[i], last ? last[i] : 2);
    auto inner = std::make_shared<CrossingState<int>>(hyp_map, params, others, ego, false, false, false, policies);
    auto start = std::make_shared<RS>(inner, hyp_map);
    Recorder rec;
    g_recorder = &rec;
    auto result = RandomHeuristic::rollout<RS, UctStatistic, HypothesisStatistic, RandomHeuristic>(
        start, rp, depth ? depth[i] : 0);
    g_recorder = nullptr;
    ego_ret[i] = std::get<0>(result);
    const auto& oth = std::get<1>(result);
    for (uint32_t j = 0; j + 1 < A; ++j) other_ret[j * n + i] = oth.at(j + 1);
    const auto& costs = std::get<2>(result);
    cost0[i] = costs.empty() ? 0.0 : costs[0];
    step_len[i] = std::get<3>(result);
    steps[i] = static_cast<uint32_t>(rec.steps.size());
    // replay on a plain state (no policies needed) for the final state
    std::shared_ptr<CrossingState<int>> cur = std::make_shared<CrossingState<int>>(
        hyp_map, params, others, ego, false, false, false, std::vector<AgentPolicyCrossingState<int>>());
    for (size_t s = 0; s < rec.steps.size(); ++s) {
      JointAction ja(A);
      for (uint32_t a = 0; a < A; ++a) ja[a] = static_cast<ActionIdx>(rec.steps[s].action[a]);
      std::vector<Reward> r; EgoCosts c;
      cur = cur->execute(ja, r, c);
      if (s < rec_stride && rec_actions) {
        rec_actions[(static_cast<size_t>(i) * rec_stride + s) * A + 0] = static_cast<int8_t>(ja[0]);
        for (uint32_t a = 1; a < A; ++a)
          rec_actions[(static_cast<size_t>(i) * rec_stride + s) * A + a] = static_cast<int8_t>(aconv<int>(ja[a]));
      }
    }
    orc_state fs;
    pack_state(*cur, &fs);
    for (uint32_t a = 0; a < A; ++a) {
      if (final_x) final_x[a * n + i] = fs.x[a];
      if (final_last) final_last[a * n + i] = fs.last[a];
    }
    if (final_flags) final_flags[i] = (fs.terminal ? 1 : 0) | (fs.goal ? 2 : 0) | (fs.collided ? 4 : 0);
  }
  return 0;
}

// Same for SimpleState (agents 4 and 5 -> positions 0 and 1).
int ref_rollout_simple(const mamcts_params* mp, uint32_t n, const int32_t* len,
                       const uint32_t* depth, double* ego_ret, double* other_ret,
                       double* step_len, uint32_t* steps, int32_t* final_len,
                       int8_t* rec_actions, double* rec_rewards, uint32_t rec_stride) {
  const MctsParameters rp = to_ref_params(*mp);
  using RS = RecState<SimpleState>;
  for (uint32_t i = 0; i < n; ++i) {
    auto inner = std::make_shared<SimpleState>(len[i]);
    auto start = std::make_shared<RS>(inner);
    Recorder rec;
    g_recorder = &rec;
    auto result = RandomHeuristic::rollout<RS, UctStatistic, UctStatistic, RandomHeuristic>(
        start, rp, depth ? depth[i] : 0);
    g_recorder = nullptr;
    ego_ret[i] = std::get<0>(result);
    other_ret[i] = std::get<1>(result).at(5);
    step_len[i] = std::get<3>(result);
    steps[i] = static_cast<uint32_t>(rec.steps.size());
    std::shared_ptr<SimpleState> cur = inner;
    for (size_t s = 0; s < rec.steps.size(); ++s) {
      JointAction ja(2);
      ja[0] = rec.steps[s].action[0]; ja[1] = rec.steps[s].action[1];
      std::vector<Reward> r; EgoCosts c;
      cur = cur->execute(ja, r, c);
      if (s < rec_stride) {
        if (rec_actions) {
          rec_actions[(static_cast<size_t>(i) * rec_stride + s) * 2 + 0] = static_cast<int8_t>(ja[0]);
          rec_actions[(static_cast<size_t>(i) * rec_stride + s) * 2 + 1] = static_cast<int8_t>(ja[1]);
        }
        if (rec_rewards) rec_rewards[static_cast<size_t>(i) * rec_stride + s] = rec.steps[s].rewards[0];
      }
    }
    if (final_len) final_len[i] = cur->get_state_length();
  }
  return 0;
}

// One CrossingState<int>::execute on arbitrary states/actions (ego index, other bit-cast value).
int ref_crossing_execute(const mamcts_crossing_params* cp, uint32_t n, const int32_t* x,
                         const int32_t* last, const int32_t* action, int32_t* next_x,
                         int32_t* next_last, uint8_t* next_flags, double* ego_reward,
                         double* cost0) {
  CrossingWorld world(*cp);
  const uint32_t A = cp->num_other_agents + 1;
  for (uint32_t i = 0; i < n; ++i) {
    std::vector<int32_t> xi(A), li(A);
    for (uint32_t a = 0; a < A; ++a) { xi[a] = x[a * n + i]; li[a] = last ? last[a * n + i] : 2; }
    auto st = world.make(xi.data(), li.data(), false);
    JointAction ja(A);
    ja[0] = static_cast<ActionIdx>(action[0 * n + i]);
    for (uint32_t a = 1; a < A; ++a) ja[a] = aconv<int>(action[a * n + i]);
    std::vector<Reward> r; EgoCosts c;
    auto nx = st->execute(ja, r, c);
    orc_state fs;
    pack_state(*nx, &fs);
    for (uint32_t a = 0; a < A; ++a) { next_x[a * n + i] = fs.x[a]; next_last[a * n + i] = fs.last[a]; }
    next_flags[i] = (fs.terminal ? 1 : 0) | (fs.goal ? 2 : 0) | (fs.collided ? 4 : 0);
    ego_reward[i] = r[0];
    cost0[i] = c[0];
  }
  return 0;
}

// Mcts<CrossingState<int>, UctStatistic, UctStatistic, RandomHeuristic>::search (mcts/mcts.h:133-141).
// Root statistics per agent position: value/latest/visits [A], q/n [A*max_actions] (n = -1: unexpanded).
int ref_search_crossing(const mamcts_params* mp, const mamcts_crossing_params* cp,
                        const int32_t* root_x, const int32_t* root_last, uint32_t max_actions,
                        double* root_value, double* root_latest, uint32_t* root_visits,
                        double* root_q, int64_t* root_n, uint32_t* counters, double* tree_records,
                        uint32_t tree_capacity, uint32_t* tree_count) {
  const MctsParameters rp = to_ref_params(*mp);
  CrossingWorld world(*cp);
  auto state = world.make(root_x, root_last, false);
  run_search_and_dump<CrossingState<int>, RandomHeuristic>(
      *state, rp, cp->num_other_agents + 1, max_actions, root_value, root_latest, root_visits,
      root_q, root_n, counters, tree_records, tree_capacity, tree_count);
  return 0;
}

int ref_search_simple(const mamcts_params* mp, int32_t state_length, uint32_t max_actions,
                      double* root_value, double* root_latest, uint32_t* root_visits,
                      double* root_q, int64_t* root_n, uint32_t* counters, double* tree_records,
                      uint32_t tree_capacity, uint32_t* tree_count) {
  const MctsParameters rp = to_ref_params(*mp);
  SimpleState state(state_length);
  run_search_and_dump<SimpleState, RandomHeuristic>(state, rp, 2, max_actions, root_value,
                                                    root_latest, root_visits, root_q, root_n,
                                                    counters, tree_records, tree_capacity,
                                                    tree_count);
  return 0;
}

// Stochastic twin: the reference's own tree code with the Philox heuristic.
// env: MAMCTS_ENV_*; tree_id = Philox stream_hi; K rollouts per leaf.
int ref_search_philox(int32_t env, const mamcts_params* mp, const mamcts_crossing_params* cp,
                      const mamcts_simple_params* sp, const int32_t* root_x,
                      const int32_t* root_last, uint64_t seed, uint32_t tree_id, uint32_t K,
                      uint32_t max_actions, double* root_value, double* root_latest,
                      uint32_t* root_visits, double* root_q, int64_t* root_n, uint32_t* counters,
                      uint64_t* rollout_steps, double* tree_records, uint32_t tree_capacity,
                      uint32_t* tree_count) {
  MctsParameters rp = to_ref_params(*mp);
  rp.NUM_PARALLEL_MCTS = 1;
  PhiloxCtx& c = g_philox;
  c.mp = *mp;
  c.seed = seed;
  c.tree_id = tree_id;
  c.K = K;
  c.rollout_steps = 0;
  if (env == MAMCTS_ENV_CROSSING_I32) {
    orc_env_crossing(&c.env, cp);
    CrossingWorld world(*cp);
    auto state = world.make(root_x, root_last, false);
    run_search_and_dump<CrossingState<int>, PhiloxHeuristic>(
        *state, rp, cp->num_other_agents + 1, max_actions, root_value, root_latest, root_visits,
        root_q, root_n, counters, tree_records, tree_capacity, tree_count);
  } else {
    orc_env_simple(&c.env, sp);
    SimpleState state(root_x[0]);
    run_search_and_dump<SimpleState, PhiloxHeuristic>(state, rp, 2, max_actions, root_value,
                                                      root_latest, root_visits, root_q, root_n,
                                                      counters, tree_records, tree_capacity,
                                                      tree_count);
  }
  if (rollout_steps) *rollout_steps = c.rollout_steps;
  return 0;
}

// UctStatistic::merge_node_statistics (uct_statistic.h:165-184) on `n` ego root statistics given as
// arrays (q/cnt [n*max_actions], cnt = -1: action absent; visits[n]).
int ref_merge_uct(const mamcts_params* mp, uint32_t n, uint32_t num_actions, uint32_t max_actions,
                  const double* q, const int64_t* cnt, const uint32_t* visits, double* merged_q,
                  int64_t* merged_n, uint32_t* merged_visits, double* merged_value,
                  uint32_t* best_action) {
  const MctsParameters rp = to_ref_params(*mp);
  std::vector<UctStatistic> stats;
  for (uint32_t t = 0; t < n; ++t) {
    UctStatistic s(num_actions, 0, rp);
    UctStatistic::UcbStatistics ucb;
    for (uint32_t a = 0; a < num_actions; ++a) {
      if (cnt[t * max_actions + a] >= 0)
        ucb[a] = UctStatistic::UcbPair(static_cast<unsigned>(cnt[t * max_actions + a]),
                                       q[t * max_actions + a], 0.0);
    }
    s.SetUcbStatistics(ucb);
    UctTest::force_visits(s, visits[t]);
    stats.push_back(s);
  }
  UctStatistic merged(num_actions, 0, rp);
  merged.merge_node_statistics(stats);
  double latest;
  UctTest::read_stat(merged, merged_value, &latest, merged_visits, merged_q, merged_n, max_actions);
  *best_action = static_cast<uint32_t>(merged.get_best_action());
  return 0;
}

// ---- ActionValueRandomHeuristic (mcts/heuristics/action_value_random_heuristic.h:27-112) --------------------
// Its non-cost-constrained branch calls a map-valued SE::set_heuristic_estimate(action_returns, action_costs)
// (:93) that UctStatistic does not offer, so the reference only ever instantiates it with
// CostConstrainedStatistic (needs OR-tools). These two test-only adapters add exactly that overload (and let
// the mean cost handed to the other agents' statistics be read), so the reference's own loop :37-85 compiles
// and runs unmodified as the checker of mamcts_action_values_crossing_i32.
namespace {
class AvEgoStat : public UctStatistic {
 public:
  using UctStatistic::UctStatistic;
  using UctStatistic::set_heuristic_estimate;
  void set_heuristic_estimate(const std::unordered_map<ActionIdx, Reward>& action_returns,
                              const std::unordered_map<ActionIdx, EgoCosts>& action_costs) {
    returns = action_returns;
    costs = action_costs;
    set_heuristic_estimate_from_backpropagated(action_returns);   // uct_statistic.h:156-163
  }
  std::unordered_map<ActionIdx, Reward> returns;
  std::unordered_map<ActionIdx, EgoCosts> costs;
};
class AvOtherStat : public UctStatistic {
 public:
  using UctStatistic::UctStatistic;
  void set_heuristic_estimate(const Reward& accum_rewards, const EgoCosts& accum_ego_cost) {
    seen_cost = accum_ego_cost;
    reward = accum_rewards;
    UctStatistic::set_heuristic_estimate(accum_rewards, accum_ego_cost);
  }
  EgoCosts seen_cost;
  Reward reward = 0.0;
};
}  // namespace

extern "C" int ref_action_values_crossing(const mamcts_params* mp, const mamcts_crossing_params* cp, uint32_t n,
                                          const int32_t* x, const int32_t* last, const uint8_t* flags,
                                          const uint32_t* depth, uint32_t max_actions, double* action_return,
                                          double* action_cost, double* other_ret, double* mean_cost) {
  const MctsParameters rp = to_ref_params(*mp);
  CrossingWorld world(*cp);
  const uint32_t A = cp->num_other_agents + 1;
  using S = CrossingState<int>;
  using H = ActionValueRandomHeuristic;
  using Node = StageNode<S, AvEgoStat, AvOtherStat, H>;
  for (uint32_t i = 0; i < n; ++i) {
    std::vector<int32_t> xi(A), li(A);
    for (uint32_t a = 0; a < A; ++a) { xi[a] = x[a * n + i]; li[a] = last ? last[a * n + i] : 2; }
    auto state = world.make(xi.data(), li.data(), flags ? (flags[i] & 1) : false);
    const unsigned d = depth ? depth[i] : 0;
    auto node = std::make_shared<Node, std::shared_ptr<Node>, std::shared_ptr<S>, const JointAction&, const unsigned int&>(
        nullptr, state->clone(), JointAction(), d, rp);
    H heuristic(rp);
    auto result = heuristic.calculate_heuristic_values<S, AvEgoStat, AvOtherStat, H>(node);
    for (uint32_t a = 0; a < max_actions; ++a) { action_return[(size_t)i * max_actions + a] = 0.0; action_cost[(size_t)i * max_actions + a] = 0.0; }
    for (const auto& kv : result.first.returns) action_return[(size_t)i * max_actions + kv.first] = kv.second;
    for (const auto& kv : result.first.costs) action_cost[(size_t)i * max_actions + kv.first] = kv.second.empty() ? 0.0 : kv.second[0];
    uint32_t j = 0;
    for (auto ai : state->get_other_agent_idx()) {
      const AvOtherStat& so = result.second.at(ai);
      other_ret[(size_t)j * n + i] = so.reward;
      mean_cost[i] = so.seen_cost.empty() ? 0.0 : so.seen_cost[0];
      ++j;
    }
  }
  return 0;
}

// ---- hypothesis-MCTS (BASELINE configs[3]) -------------------------------------------------------------------
// Mcts<CrossingState<int>, UctStatistic, HypothesisStatistic, RandomHeuristic>::search(state, belief_tracker)
// (mcts/mcts.h:143-164) exactly as environments/tests/crossing_state_int_test.cc:141-229 sets it up: a
// HypothesisBeliefTracker, a CrossingState referring to its sampled hypotheses, add_hypothesis per gap range,
// belief_update(state, state) to initialise the beliefs from the priors. `skip` earlier
// sample_current_hypothesis() calls emulate a tracker that has already served other decisions.
// Outputs: the hypothesis ids the tracker hands out during the search (from a copy of the tracker, sequence
// [iterations][J]), the canonical tree dump (UctTest::dump_hyp) and counters {nodes, iterations, best action}.
namespace {
MctsParameters to_ref_hyp_params(const mamcts_params& mp, const mamcts_hypothesis_params& hp) {
  MctsParameters r = to_ref_params(mp);
  r.hypothesis_statistic.UPPER_COST_BOUND = hp.upper_cost_bound;
  r.hypothesis_statistic.LOWER_COST_BOUND = hp.lower_cost_bound;
  r.hypothesis_statistic.PROGRESSIVE_WIDENING_K = hp.pw_k;
  r.hypothesis_statistic.PROGRESSIVE_WIDENING_ALPHA = hp.pw_alpha;
  r.hypothesis_statistic.EXPLORATION_CONSTANT = hp.exploration_constant;
  r.hypothesis_statistic.COST_BASED_ACTION_SELECTION = hp.cost_based_action_selection != 0;
  r.hypothesis_statistic.PROGRESSIVE_WIDENING_HYPOTHESIS_BASED = hp.progressive_widening_hypothesis_based != 0;
  r.cost_constrained_statistic.USE_CHANCE_CONSTRAINED_UPDATES.clear();
  if (hp.chance_constrained_update) r.cost_constrained_statistic.USE_CHANCE_CONSTRAINED_UPDATES.push_back(true);
  return r;
}
}  // namespace

extern "C" int ref_search_hypothesis(const mamcts_params* mp, const mamcts_hypothesis_params* hp,
                                     const mamcts_crossing_params* cp, uint32_t num_hyp, const int32_t* gap_lo,
                                     const int32_t* gap_hi, const int32_t* root_x, const int32_t* root_last,
                                     uint32_t skip, uint8_t* sequence /* [iterations][J] */, uint32_t* counters /* [3] */,
                                     double* tree_records, uint32_t tree_capacity, uint32_t* tree_count,
                                     double* iterations_per_s) {
  const MctsParameters rp = to_ref_hyp_params(*mp, *hp);
  CrossingStateParameters<int> params = to_ref_crossing(*cp);
  params.OTHER_AGENTS_POLICY_RANDOM_SEED = hp->policy_random_seed;
  const uint32_t J = cp->num_other_agents;
  HypothesisBeliefTracker tracker(rp);
  std::vector<AgentState<int>> others;
  for (uint32_t j = 0; j < J; ++j) others.emplace_back(root_x[j + 1], root_last ? root_last[j + 1] : 2);
  AgentState<int> ego(root_x[0], root_last ? root_last[0] : 2);
  std::vector<AgentPolicyCrossingState<int>> hyps;
  for (uint32_t h = 0; h < num_hyp; ++h) hyps.emplace_back(std::pair<int, int>(gap_lo[h], gap_hi[h]), params);
  // the map the state refers to must hold every agent BEFORE the first sample (crossing_state_int_test.cc:147)
  auto state = std::make_shared<CrossingState<int>>(tracker.sample_current_hypothesis(), params, others, ego, false, false,
                                                    false, hyps);
  tracker.belief_update(*state, *state);
  for (uint32_t k = 0; k < skip; ++k) tracker.sample_current_hypothesis();
  if (sequence) {
    HypothesisBeliefTracker copy = tracker;
    for (uint32_t it = 0; it < rp.MAX_NUMBER_OF_ITERATIONS; ++it) {
      const auto& cur = copy.sample_current_hypothesis();
      for (uint32_t j = 0; j < J; ++j) sequence[static_cast<size_t>(it) * J + j] = static_cast<uint8_t>(cur.at(j + 1));
    }
  }
  using M = Mcts<CrossingState<int>, UctStatistic, HypothesisStatistic, RandomHeuristic>;
  M mcts(rp);
  const auto t0 = std::chrono::steady_clock::now();
  mcts.search(*state, tracker);
  const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (iterations_per_s) *iterations_per_s = mcts.numIterations() / el;
  counters[0] = mcts.numNodes();
  counters[1] = mcts.numIterations();
  counters[2] = static_cast<uint32_t>(mcts.returnBestAction());
  if (tree_records && tree_count) {
    const uint32_t NE = static_cast<uint32_t>(cp->max_velocity_ego - cp->min_velocity_ego + 1);
    const uint32_t MA = static_cast<uint32_t>(cp->max_velocity_other - cp->min_velocity_other + 1);
    *tree_count = UctTest::dump_hyp<CrossingState<int>, RandomHeuristic>(
        UctTest::root_of_hyp<CrossingState<int>, RandomHeuristic>(mcts), tree_records, tree_capacity, 0, NE, MA, num_hyp,
        cp->min_velocity_other);
  }
  return 0;
}

// `threads` independent hypothesis searches of the same configuration, one std::thread each (the reference's own
// root-parallel mode for hypothesis searches does the same, mcts.h:237-251); returns iterations / wall time.
extern "C" int ref_time_search_hypothesis(const mamcts_params* mp, const mamcts_hypothesis_params* hp,
                                          const mamcts_crossing_params* cp, uint32_t num_hyp, const int32_t* gap_lo,
                                          const int32_t* gap_hi, uint32_t threads, double* iterations_per_s,
                                          double* seconds_out) {
  std::vector<uint64_t> iters(threads, 0);
  const uint32_t A = cp->num_other_agents + 1;
  std::vector<int32_t> x(A, 5), last(A, 2);
  const auto t0 = std::chrono::steady_clock::now();
  auto body = [&](uint32_t tid) {
    uint32_t counters[3] = {0, 0, 0};
    ref_search_hypothesis(mp, hp, cp, num_hyp, gap_lo, gap_hi, x.data(), last.data(), tid, nullptr, counters, nullptr, 0,
                          nullptr, nullptr);
    iters[tid] = counters[1];
  };
  std::vector<std::thread> pool;
  for (uint32_t t = 0; t < threads; ++t) pool.emplace_back(body, t);
  for (auto& th : pool) th.join();
  const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  uint64_t it = 0;
  for (uint32_t t = 0; t < threads; ++t) it += iters[t];
  *iterations_per_s = it / el;
  *seconds_out = el;
  return 0;
}

// ---- timing of the reference's own CPU path (bench.py --impl reference / cpu_baseline) ---------

// `threads` independent loops of RandomHeuristic::rollout from the default start state for about
// `seconds` s; returns env-steps/s (execute() calls / wall time) and rollouts/s.
int ref_time_rollouts(const mamcts_params* mp, const mamcts_crossing_params* cp, uint32_t threads,
                      double seconds, double* env_steps_per_s, double* rollouts_per_s) {
  const MctsParameters rp = to_ref_params(*mp);
  std::vector<uint64_t> steps(threads, 0), rollouts(threads, 0);
  const auto t0 = std::chrono::steady_clock::now();
  auto body = [&](uint32_t tid) {
    CrossingWorld world(*cp);
    const uint32_t A = cp->num_other_agents + 1;
    std::vector<int32_t> x(A, 5), last(A, 2);
    auto start = world.make(x.data(), last.data(), false);
    for (;;) {
      for (int rep = 0; rep < 64; ++rep) {
        auto r = RandomHeuristic::rollout<CrossingState<int>, UctStatistic, UctStatistic,
                                          RandomHeuristic>(start, rp, 0);
        steps[tid] += static_cast<uint64_t>(std::get<3>(r));
        rollouts[tid] += 1;
      }
      const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (el >= seconds) break;
    }
  };
  std::vector<std::thread> pool;
  for (uint32_t t = 0; t < threads; ++t) pool.emplace_back(body, t);
  for (auto& th : pool) th.join();
  const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  uint64_t s = 0, r = 0;
  for (uint32_t t = 0; t < threads; ++t) { s += steps[t]; r += rollouts[t]; }
  *env_steps_per_s = s / el;
  *rollouts_per_s = r / el;
  return 0;
}

// `threads` independent Mcts::search calls (single_search, mcts/mcts.h:166-186), `repeats` each;
// returns iterations/s and the env-steps/s of the rollouts inside them (execute() calls counted
// with the recording wrapper would perturb timing, so steps are derived from one untimed recorded
// search and scaled).
int ref_time_search(const mamcts_params* mp, const mamcts_crossing_params* cp, uint32_t threads,
                    uint32_t repeats, double* iterations_per_s, double* seconds_out) {
  const MctsParameters rp = to_ref_params(*mp);
  std::vector<uint64_t> iters(threads, 0);
  const auto t0 = std::chrono::steady_clock::now();
  auto body = [&](uint32_t tid) {
    CrossingWorld world(*cp);
    const uint32_t A = cp->num_other_agents + 1;
    std::vector<int32_t> x(A, 5), last(A, 2);
    auto start = world.make(x.data(), last.data(), false);
    for (uint32_t r = 0; r < repeats; ++r) {
      Mcts<CrossingState<int>, UctStatistic, UctStatistic, RandomHeuristic> mcts(rp);
      mcts.search(*start);
      iters[tid] += mcts.numIterations();
    }
  };
  std::vector<std::thread> pool;
  for (uint32_t t = 0; t < threads; ++t) pool.emplace_back(body, t);
  for (auto& th : pool) th.join();
  const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  uint64_t it = 0;
  for (uint32_t t = 0; t < threads; ++t) it += iters[t];
  *iterations_per_s = it / el;
  *seconds_out = el;
  return 0;
}

// One untimed Mcts::search with the recording state as S: counts every execute() call (expansion
// + rollout steps) and the iterations, so timed runs can be converted to env-steps/s exactly (the
// search is deterministic).
int ref_search_count_steps(const mamcts_params* mp, const mamcts_crossing_params* cp,
                           uint64_t* execute_calls, uint32_t* iterations, uint32_t* nodes) {
  const MctsParameters rp = to_ref_params(*mp);
  CrossingWorld world(*cp);
  const uint32_t A = cp->num_other_agents + 1;
  std::vector<int32_t> x(A, 5), last(A, 2);
  using RS = RecState<CrossingState<int>>;
  RS start(world.make(x.data(), last.data(), false));
  Recorder rec;
  g_recorder = &rec;
  Mcts<RS, UctStatistic, UctStatistic, RandomHeuristic> mcts(rp);
  mcts.search(start);
  g_recorder = nullptr;
  *execute_calls = rec.steps.size();
  *iterations = mcts.numIterations();
  *nodes = mcts.numNodes();
  return 0;
}

// The reference's own root-parallel path: Mcts::parallel_search with USE_MULTI_THREADING
// (mcts/mcts.h:188-221) - the runnable stand-in for test/parallel_mcts (needs OR-tools).
int ref_time_parallel_search(const mamcts_params* mp, const mamcts_crossing_params* cp,
                             uint32_t num_parallel, double* iterations_per_s, double* seconds_out,
                             double* root_q, int64_t* root_n, uint32_t max_actions) {
  MctsParameters rp = to_ref_params(*mp);
  rp.NUM_PARALLEL_MCTS = num_parallel;
  rp.USE_MULTI_THREADING = true;
  CrossingWorld world(*cp);
  const uint32_t A = cp->num_other_agents + 1;
  std::vector<int32_t> x(A, 5), last(A, 2);
  auto start = world.make(x.data(), last.data(), false);
  const auto t0 = std::chrono::steady_clock::now();
  Mcts<CrossingState<int>, UctStatistic, UctStatistic, RandomHeuristic> mcts(rp);
  mcts.search(*start);
  const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  *iterations_per_s = static_cast<double>(num_parallel) * rp.MAX_NUMBER_OF_ITERATIONS / el;
  *seconds_out = el;
  if (root_q && root_n) {
    double v, l; uint32_t nv;
    UctTest::read_stat(mcts.get_root().get_ego_int_node(), &v, &l, &nv, root_q, root_n, max_actions);
  }
  return 0;
}

}  // extern "C"
