// tests/dropin/dropin_test.cc - the drop-in check: the reference's OWN host code
// (Mcts<S,SE,SO,H>::search, StageNode, parallel_search; juloberno/mamcts headers, unmodified) is
// instantiated with this repository's CudaRandomHeuristic / CudaUctStatistic and compared with the
// stock RandomHeuristic / UctStatistic instantiation in the same binary, tree by tree, bit by bit.
//
// Built by oracle/Makefile (target dropin) where the reference checkout exists; the binary travels
// to the GPU box and is run by tests/test_gpu_dropin.py. Exit code = number of failed checks.
#define UNIT_TESTING  // MCTS_TEST -> friend class UctTest (mcts/common.h:16-20)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <iterator>
#include <set>
#include <string>
#include <map>
#include <sstream>
#include <vector>

#include "mcts/mcts.h"
#include "mcts/heuristics/random_heuristic.h"
#include "mcts/statistics/uct_statistic.h"
#include "environments/crossing_state.h"
#include "test/uct/simple_state.h"

#include "batched_mcts.h"
#include "cuda_random_heuristic.h"
#include "cuda_state.h"
#include "cuda_uct_statistic.h"
#include "device_mcts.h"
#include "mamcts_oracle.h"  // CPU checker for the PHILOX cases (test infrastructure)

using namespace mcts;

static int g_failures = 0;
#define CHECK_MSG(cond, name)                                        \
  do {                                                               \
    if (cond) { std::printf("PASS %s\n", name); }                    \
    else { std::printf("FAIL %s\n", name); ++g_failures; }           \
  } while (0)

// statistic readers: UctStatistic keeps its members protected (friend UctTest), CudaUctStatistic has
// public accessors.
class mcts::UctTest {
 public:
  static void read(const UctStatistic& s, double* v, double* g, unsigned* n, std::map<ActionIdx, std::pair<unsigned, double>>* ucb) {
    *v = s.value_; *g = s.latest_return_; *n = s.total_node_visits_;
    for (const auto& kv : s.ucb_statistics_) (*ucb)[kv.first] = {kv.second.action_count_, kv.second.action_value_};
  }
  static void read(const CudaUctStatistic& s, double* v, double* g, unsigned* n, std::map<ActionIdx, std::pair<unsigned, double>>* ucb) {
    *v = s.get_value(); *g = s.get_latest_return(); *n = s.get_total_node_visits();
    for (const auto& kv : s.get_ucb_statistics()) (*ucb)[kv.first] = {kv.second.action_count_, kv.second.action_value_};
  }
  template <class M>
  static auto root_of(M& m) { return m.root_; }
};

// canonical dump (same record as mamcts_search_dump_tree / orc_tree_dump)
template <class NodePtr>
static void dump_tree(const NodePtr& node, uint32_t max_actions, std::vector<double>* out) {
  const uint32_t A = node->get_state()->get_num_agents();
  auto code_of = [&](const JointAction& ja) {
    double c = 0.0, mul = 1.0;
    for (uint32_t a = 0; a < A; ++a) { c += mul * (double)ja[a]; mul *= max_actions; }
    return c;
  };
  auto children = node->get_children();
  std::map<double, typename decltype(children)::mapped_type> sorted;
  for (auto& kv : children) sorted[code_of(kv.first)] = kv.second;
  out->push_back(node->get_depth());
  out->push_back(-2.0);  // joint-action code is written by the parent below
  const size_t code_pos = out->size() - 1;
  (void)code_pos;
  out->push_back(node->get_visit_count());
  out->push_back((double)children.size());
  auto emit = [&](const auto& stat) {
    double v, g; unsigned n;
    std::map<ActionIdx, std::pair<unsigned, double>> ucb;
    UctTest::read(stat, &v, &g, &n, &ucb);
    out->push_back(n); out->push_back(v); out->push_back(g);
    for (uint32_t k = 0; k < max_actions; ++k) out->push_back(ucb.count(k) ? (double)ucb[k].first : -1.0);
    for (uint32_t k = 0; k < max_actions; ++k) out->push_back(ucb.count(k) ? ucb[k].second : 0.0);
  };
  emit(node->get_ego_int_node());
  for (const auto& o : node->get_other_int_nodes()) emit(o);
  for (auto& kv : sorted) {
    const size_t pos = out->size();
    dump_tree(kv.second, max_actions, out);
    (*out)[pos + 1] = kv.first;
  }
}

template <class NodePtr>
static std::vector<double> dump_root(const NodePtr& root, uint32_t max_actions) {
  std::vector<double> out;
  dump_tree(root, max_actions, &out);
  out[1] = -1.0;
  return out;
}

static bool bit_equal(const std::vector<double>& a, const std::vector<double>& b) {
  return a.size() == b.size() && (a.empty() || std::memcmp(a.data(), b.data(), a.size() * sizeof(double)) == 0);
}

static MctsParameters test_params(unsigned iterations, unsigned seed = 1000) {
  MctsParameters p = mcts_default_parameters();
  p.RANDOM_SEED = seed;
  p.MAX_NUMBER_OF_ITERATIONS = iterations;
  p.MAX_NUMBER_OF_NODES = 100000000;
  p.MAX_SEARCH_TIME = 1000000000;          // no wall-clock cut-offs
  p.random_heuristic.MAX_SEARCH_TIME = 1e9;
  p.random_heuristic.MAX_NUMBER_OF_ITERATIONS = 50;
  return p;
}

struct CrossingFixture {
  CrossingStateParameters<int> params;
  std::unordered_map<AgentIdx, HypothesisId> hyp;
  explicit CrossingFixture(unsigned others) : params(default_crossing_state_parameters<int>()) {
    params.NUM_OTHER_AGENTS = others;
    for (AgentIdx i = 1; i <= others; ++i) hyp[i] = 0;
    CudaState<CrossingState<int>>::set_parameters(params);
  }
  CrossingState<int> state() { return CrossingState<int>(hyp, params); }
};

template <class S, class SE, class SO, class H>
static std::vector<double> search_and_dump(const S& state, const MctsParameters& p, uint32_t max_actions,
                                           ActionIdx* best = nullptr) {
  Mcts<S, SE, SO, H> mcts(p);
  mcts.search(state);
  if (best) *best = mcts.returnBestAction();
  return dump_root(UctTest::root_of(mcts), max_actions);
}

static std::vector<double> oracle_dump(const orc_env& env, const mamcts_params& mp, const int32_t* root_x,
                                       uint32_t iterations, uint64_t seed, uint32_t tree_id, uint32_t max_actions) {
  orc_search_config cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.env = env;
  cfg.mcts = mp;
  cfg.action_source = MAMCTS_SRC_PHILOX;
  cfg.rollouts_per_leaf = 1;
  cfg.philox_seed = seed;
  cfg.tree_id = tree_id;
  for (uint32_t a = 0; a < env.num_agents; ++a) mamcts_reference_expand_order(mp.random_seed, env.num_actions[a], cfg.expand_order[a]);
  orc_state root;
  std::memset(&root, 0, sizeof(root));
  for (uint32_t a = 0; a < env.num_agents; ++a) { root.x[a] = root_x[a]; root.last[a] = 2; }
  orc_tree* t = orc_tree_create(&cfg, &root);
  orc_tree_run(t, iterations);
  const uint32_t nn = orc_tree_num_nodes(t);
  std::vector<double> rec((size_t)nn * (4 + env.num_agents * (3 + 2 * max_actions)));
  orc_tree_dump(t, rec.data(), nn, max_actions);
  orc_tree_destroy(t);
  return rec;
}

// Check 6: the back-propagation invariant of the reference's verify_uct (test/uct/uct_test_class.h),
// stated positionally so that it also holds on PHILOX trees (non-diagonal joint actions, terminal
// children that are re-visited: stage_node.h:183-187 returns without a rollout but mcts.h:314-329
// still back-propagates the child's latest return):
//   n_a(parent) = sum over children c with ja_c[p] == a of (1 + visit_count_c)
//   Q_a(parent) = sum over those children of (1 + visit_count_c) (r_c[p] + gamma V_c) / n_a(parent)
//   N(parent)   = sum_a n_a(parent) + (parent is not the root: 1 for its own heuristic update)
template <class NodePtr>
static void check_invariant(const NodePtr& node, double gamma, long* nodes_checked, long* bad) {
  auto children = node->get_children();
  if (children.empty()) return;
  const auto* st = node->get_state();
  const unsigned A = st->get_num_agents();
  auto read_pos = [](const auto& nd, unsigned p, double* v, unsigned* n, std::map<ActionIdx, std::pair<unsigned, double>>* ucb) {
    double g;
    if (p == 0) UctTest::read(nd->get_ego_int_node(), v, &g, n, ucb);
    else UctTest::read(nd->get_other_int_nodes()[p - 1], v, &g, n, ucb);
  };
  for (unsigned p = 0; p < A; ++p) {
    double pv; unsigned pn;
    std::map<ActionIdx, std::pair<unsigned, double>> pucb;
    read_pos(node, p, &pv, &pn, &pucb);
    std::map<ActionIdx, double> exp_n, exp_q;
    for (const auto& kv : children) {
      const JointAction& ja = kv.first;
      std::vector<Reward> rewards;
      EgoCosts cost;
      st->execute(ja, rewards, cost);
      double cv; unsigned cn;
      std::map<ActionIdx, std::pair<unsigned, double>> cucb;
      read_pos(kv.second, p, &cv, &cn, &cucb);
      const double w = 1.0 + kv.second->get_visit_count();
      exp_n[ja[p]] += w;
      exp_q[ja[p]] += w * (rewards[p] + gamma * cv);
    }
    double sum_n = 0.0;
    for (const auto& kv : exp_n) {
      const ActionIdx a = kv.first;
      const bool have = pucb.count(a) == 1;
      const double q = exp_q[a] / kv.second;
      if (!have || (double)pucb[a].first != kv.second ||
          std::fabs(pucb[a].second - q) > 1e-9 * (1.0 + std::fabs(q))) {
        if (*bad < 5) std::printf("  invariant violated: depth %u agent position %u action %zu: n %u vs %.0f, q %.17g vs %.17g\n",
                                  node->get_depth(), p, (size_t)a, have ? pucb[a].first : 0u, kv.second, have ? pucb[a].second : 0.0, q);
        ++*bad;
      }
      sum_n += kv.second;
    }
    if ((double)pn != sum_n + (node->is_root() ? 0.0 : 1.0)) {
      if (*bad < 5) std::printf("  invariant violated: depth %u agent position %u: N %u vs %.0f\n", node->get_depth(), p, pn, sum_n);
      ++*bad;
    }
  }
  *nodes_checked += 1;
  for (const auto& kv : children) check_invariant(kv.second, gamma, nodes_checked, bad);
}

template <class S, class SE, class SO>
static bool philox_tree_satisfies_invariant(const S& state, const MctsParameters& p, long min_nodes) {
  CudaRandomHeuristic::options() = CudaRandomHeuristic::Options();
  CudaRandomHeuristic::options().action_source = MAMCTS_SRC_PHILOX;
  CudaRandomHeuristic::options().philox_seed = 77;
  Mcts<S, SE, SO, CudaRandomHeuristic> mcts(p);
  mcts.search(state);
  long nodes = 0, bad = 0;
  check_invariant(UctTest::root_of(mcts), p.DISCOUNT_FACTOR, &nodes, &bad);
  CudaRandomHeuristic::options() = CudaRandomHeuristic::Options();
  std::printf("  %ld inner nodes checked, %ld violations\n", nodes, bad);
  return bad == 0 && nodes >= min_nodes;
}

int main() {
  // 1. CudaRandomHeuristic (REFERENCE_CONST) under the reference's Mcts reproduces the stock tree.
  for (unsigned others : {1u, 2u}) {
    CrossingFixture fx(others);
    const MctsParameters p = test_params(others == 1 ? 1500 : 700);
    CudaRandomHeuristic::options() = CudaRandomHeuristic::Options();
    auto state = fx.state();
    auto ref = search_and_dump<CrossingState<int>, UctStatistic, UctStatistic, RandomHeuristic>(state, p, 7);
    auto gpu = search_and_dump<CrossingState<int>, UctStatistic, UctStatistic, CudaRandomHeuristic>(state, p, 7);
    CHECK_MSG(bit_equal(ref, gpu), others == 1 ? "crossing2_cuda_heuristic_tree_bit_equal" : "crossing3_cuda_heuristic_tree_bit_equal");
    // 2. ... and with CudaUctStatistic as SE and SO (backup through mamcts_backup_uct).
    ActionIdx best_ref = 0, best_gpu = 0;
    auto ref2 = search_and_dump<CrossingState<int>, UctStatistic, UctStatistic, RandomHeuristic>(state, p, 7, &best_ref);
    auto gpu2 = search_and_dump<CrossingState<int>, CudaUctStatistic, CudaUctStatistic, CudaRandomHeuristic>(state, p, 7, &best_gpu);
    CHECK_MSG(bit_equal(ref2, gpu2) && best_ref == best_gpu,
              others == 1 ? "crossing2_cuda_statistic_tree_bit_equal" : "crossing3_cuda_statistic_tree_bit_equal");
  }
  {  // SimpleState (config 1: test/uct/uct_test.cc parameters, 20 and 300 iterations, 100 nodes)
    for (unsigned iters : {20u, 300u}) {
      MctsParameters p = test_params(iters);
      p.MAX_NUMBER_OF_NODES = 100;
      p.random_heuristic.MAX_NUMBER_OF_ITERATIONS = 40;
      SimpleState state(4);
      auto ref = search_and_dump<SimpleState, UctStatistic, UctStatistic, RandomHeuristic>(state, p, 2);
      auto gpu = search_and_dump<SimpleState, CudaUctStatistic, CudaUctStatistic, CudaRandomHeuristic>(state, p, 2);
      CHECK_MSG(bit_equal(ref, gpu), iters == 20 ? "simple_20_iterations_tree_bit_equal" : "simple_300_iterations_tree_bit_equal");
    }
  }
  {  // 3. static rollout<> keeps RandomHeuristic's tuple (used by ActionValueRandomHeuristic, :61)
    CrossingFixture fx(2);
    const MctsParameters p = test_params(10);
    auto st = std::make_shared<CrossingState<int>>(fx.state());
    auto a = RandomHeuristic::rollout<CrossingState<int>, UctStatistic, UctStatistic, RandomHeuristic>(st, p, 0);
    auto b = CudaRandomHeuristic::rollout<CrossingState<int>, UctStatistic, UctStatistic, CudaRandomHeuristic>(st, p, 0);
    bool same = std::get<0>(a) == std::get<0>(b) && std::get<2>(a) == std::get<2>(b) && std::get<3>(a) == std::get<3>(b) &&
                std::get<1>(a) == std::get<1>(b);
    CHECK_MSG(same && std::get<0>(b) == 0.0013962886019873665 && std::get<3>(b) == 7.0, "static_rollout_tuple_equal");
  }
  {  // 4. parallel_search (threads) + merge_node_statistics through mamcts_root_merge
    CrossingFixture fx(1);
    MctsParameters p = test_params(400);
    p.NUM_PARALLEL_MCTS = 4;
    p.USE_MULTI_THREADING = true;
    auto state = fx.state();
    Mcts<CrossingState<int>, UctStatistic, UctStatistic, RandomHeuristic> ref(p);
    ref.search(state);
    Mcts<CrossingState<int>, CudaUctStatistic, CudaUctStatistic, CudaRandomHeuristic> gpu(p);
    gpu.search(state);
    double v, g, v2, g2; unsigned n, n2;
    std::map<ActionIdx, std::pair<unsigned, double>> a, b;
    UctTest::read(ref.get_root().get_ego_int_node(), &v, &g, &n, &a);
    UctTest::read(gpu.get_root().get_ego_int_node(), &v2, &g2, &n2, &b);
    bool ok = n == n2 && a.size() == b.size() && std::fabs(v - v2) <= 1e-12 * std::fabs(v) &&
              ref.returnBestAction() == gpu.returnBestAction();
    for (auto& kv : a) ok = ok && b.count(kv.first) && b[kv.first].first == kv.second.first &&
                            std::fabs(b[kv.first].second - kv.second.second) <= 1e-12 * std::fabs(kv.second.second) + 1e-300;
    CHECK_MSG(ok, "parallel_search_merge_matches");
  }
  // 5. leaf-parallel batcher over host trees: one rollout launch (and one backup launch) per wave.
  for (int variant = 0; variant < 2; ++variant) {
    CrossingFixture fx(1);
    const MctsParameters p = test_params(300);
    const unsigned T = 24, first = 500;
    const uint64_t seed = 4242;
    auto state = fx.state();
    orc_env env;
    orc_env_crossing(&env, &CudaState<CrossingState<int>>::params());
    const mamcts_params mp = to_mamcts_params(p);
    const int32_t root_x[2] = {5, 5};
    bool ok = true;
    uint64_t steps = 0;
    ActionIdx best = 99;
    std::vector<std::vector<double>> dumps;
    if (variant == 0) {
      BatchedRootParallelMcts<CrossingState<int>, UctStatistic, UctStatistic> b(p, T, seed, first);
      b.search(state);
      for (unsigned t = 0; t < T; ++t) dumps.push_back(dump_root(b.get_root_ptr(t), 7));
      steps = b.rolloutSteps();
      best = b.returnBestAction();
    } else {
      BatchedRootParallelMcts<CrossingState<int>, CudaUctStatistic, CudaUctStatistic> b(p, T, seed, first);
      b.search(state);
      for (unsigned t = 0; t < T; ++t) dumps.push_back(dump_root(b.get_root_ptr(t), 7));
      steps = b.rolloutSteps();
      best = b.returnBestAction();
    }
    for (unsigned t = 0; t < T; ++t) ok = ok && bit_equal(dumps[t], oracle_dump(env, mp, root_x, 300, seed, first + t, 7));
    CHECK_MSG(ok && steps > 0 && best < 4, variant == 0 ? "batched_host_trees_equal_oracle" : "batched_host_trees_cuda_statistic_equal_oracle");
  }
  {  // 6. positional back-propagation invariant on PHILOX trees (stock and CUDA statistics)
    CrossingFixture fx(2);
    MctsParameters p = test_params(1500);
    auto cs = fx.state();
    CHECK_MSG((philox_tree_satisfies_invariant<CrossingState<int>, UctStatistic, UctStatistic>(cs, p, 50)),
              "philox_crossing_tree_backprop_invariant");
    CHECK_MSG((philox_tree_satisfies_invariant<CrossingState<int>, CudaUctStatistic, CudaUctStatistic>(cs, p, 50)),
              "philox_crossing_tree_backprop_invariant_cuda_statistic");
    MctsParameters ps = test_params(1500);
    ps.random_heuristic.MAX_NUMBER_OF_ITERATIONS = 40;
    SimpleState ss(4);
    CHECK_MSG((philox_tree_satisfies_invariant<SimpleState, UctStatistic, UctStatistic>(ss, ps, 50)),
              "philox_simple_tree_backprop_invariant");
    CHECK_MSG((philox_tree_satisfies_invariant<SimpleState, CudaUctStatistic, CudaUctStatistic>(ss, ps, 50)),
              "philox_simple_tree_backprop_invariant_cuda_statistic");
  }
  {  // 7. DeviceRootParallelMcts: the Mcts surface over the device-resident search
    CrossingFixture fx(1);
    auto state = fx.state();
    auto same_stat = [](const auto& host_stat, const DeviceRootParallelMcts<CrossingState<int>>::RootStatistic& d) {
      double v, g; unsigned n;
      std::map<ActionIdx, std::pair<unsigned, double>> ucb;
      UctTest::read(host_stat, &v, &g, &n, &ucb);
      bool ok = n == d.total_node_visits && std::memcmp(&v, &d.value, sizeof(double)) == 0 && ucb.size() == d.ucb_statistics.size();
      for (const auto& kv : ucb) {
        auto it = d.ucb_statistics.find(kv.first);
        ok = ok && it != d.ucb_statistics.end() && it->second.first == kv.second.first &&
             std::memcmp(&it->second.second, &kv.second.second, sizeof(double)) == 0;
      }
      return ok;
    };
    {  // REFERENCE_CONST: every device tree is the stock reference's tree
      const MctsParameters p = test_params(1200);
      Mcts<CrossingState<int>, UctStatistic, UctStatistic, RandomHeuristic> ref(p);
      ref.search(state);
      DeviceRootParallelMcts<CrossingState<int>> dev(p, 5, 0, 0, MAMCTS_SRC_REFERENCE_CONST);
      dev.search(state);
      bool ok = dev.returnBestAction() == ref.returnBestAction() && dev.numIterations() == ref.numIterations();
      for (unsigned t : {0u, 4u}) {
        ok = ok && same_stat(ref.get_root().get_ego_int_node(), dev.get_root_statistic(t, 0));
        ok = ok && same_stat(ref.get_root().get_other_int_nodes()[0], dev.get_root_statistic(t, 1));
      }
      ok = ok && dev.numNodes() == 5ull * ref.numNodes();
      dev.search(state);  // a second decision re-roots the same arenas
      ok = ok && same_stat(ref.get_root().get_ego_int_node(), dev.get_root_statistic(2, 0));
      CHECK_MSG(ok, "device_mcts_reference_const_equals_stock_mcts");
      // Graphviz export: same labels as the reference's printTreeToDotFile (ids and line order differ)
      ref.printTreeToDotFile("/tmp/mamcts_dropin_ref_tree");
      dev.printTreeToDotFile("/tmp/mamcts_dropin_dev_tree", 3);
      auto labels = [](const char* path) {
        std::ifstream f(path);
        std::string text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        std::multiset<std::string> out;
        size_t pos = 0, clusters = 0;
        while ((pos = text.find("[label=\"", pos)) != std::string::npos) {
          const size_t end = text.find("\"]", pos);
          out.insert(text.substr(pos + 8, end - pos - 8));
          pos = end;
        }
        pos = 0;
        while ((pos = text.find("subgraph cluster_node_", pos)) != std::string::npos) { ++clusters; ++pos; }
        out.insert("clusters=" + std::to_string(clusters));
        return out;
      };
      const auto lr = labels("/tmp/mamcts_dropin_ref_tree.gv"), ld = labels("/tmp/mamcts_dropin_dev_tree.gv");
      CHECK_MSG(lr == ld && lr.size() > 100, "device_mcts_dot_export_labels_equal_reference");
    }
    {  // PHILOX: device trees == the reference's StageNode trees fed by the same GPU rollouts
      const MctsParameters p = test_params(400);
      const unsigned T = 16, first = 300;
      BatchedRootParallelMcts<CrossingState<int>, UctStatistic, UctStatistic> host(p, T, 4242, first);
      host.search(state);
      DeviceRootParallelMcts<CrossingState<int>> dev(p, T, 4242, first);
      dev.search(state);
      bool ok = dev.returnBestAction() == host.returnBestAction() && dev.numEnvSteps() > 0;
      for (unsigned t = 0; t < T; ++t) {
        ok = ok && same_stat(host.get_root(t).get_ego_int_node(), dev.get_root_statistic(t, 0));
        ok = ok && same_stat(host.get_root(t).get_other_int_nodes()[0], dev.get_root_statistic(t, 1));
      }
      const auto merged = dev.get_merged_root_statistic();
      ok = ok && merged.total_node_visits == T * 400;
      CHECK_MSG(ok, "device_mcts_philox_equals_host_trees");
    }
    {  // SimpleState (config 1)
      MctsParameters p = test_params(20);
      p.MAX_NUMBER_OF_NODES = 100;
      p.random_heuristic.MAX_NUMBER_OF_ITERATIONS = 40;
      SimpleState ss(4);
      Mcts<SimpleState, UctStatistic, UctStatistic, RandomHeuristic> ref(p);
      ref.search(ss);
      DeviceRootParallelMcts<SimpleState> dev(p, 3, 0, 0, MAMCTS_SRC_REFERENCE_CONST);
      dev.search(ss);
      double v, g; unsigned n;
      std::map<ActionIdx, std::pair<unsigned, double>> ucb;
      UctTest::read(ref.get_root().get_ego_int_node(), &v, &g, &n, &ucb);
      const auto d = dev.get_root_statistic(1, 0);
      bool ok = n == d.total_node_visits && ucb.size() == d.ucb_statistics.size() && dev.returnBestAction() == ref.returnBestAction();
      for (const auto& kv : ucb) ok = ok && d.ucb_statistics.at(kv.first).first == kv.second.first && d.ucb_statistics.at(kv.first).second == kv.second.second;
      CHECK_MSG(ok, "device_mcts_simple_state_20_iterations");
    }
  }
  {  // 8. NON-DEFAULT environment parameters (other rewards, chain length, goal, velocity range): the device
     //    search follows them - through the constructor or through a later set_parameters() - and is never
     //    left simulating the parameters it was first created with; a terminal root is refused.
    CrossingStateParameters<int> cp = default_crossing_state_parameters<int>();
    cp.NUM_OTHER_AGENTS = 1;
    cp.CHAIN_LENGTH = 41;
    cp.EGO_GOAL_POS = 26;
    cp.REWARD_COLLISION = -300;
    cp.REWARD_GOAL_REACHED = 77;
    cp.REWARD_STEP = -1;
    cp.MAX_VELOCITY_EGO = 3;          // 5 ego actions
    std::unordered_map<AgentIdx, HypothesisId> hyp{{1, 0}};
    CrossingState<int> state(hyp, cp);
    const MctsParameters p = test_params(700);
    CudaState<CrossingState<int>>::set_parameters(cp);
    Mcts<CrossingState<int>, UctStatistic, UctStatistic, RandomHeuristic> ref(p);
    ref.search(state);
    auto same_root = [&](DeviceRootParallelMcts<CrossingState<int>>& dev) {
      double v, g; unsigned n;
      std::map<ActionIdx, std::pair<unsigned, double>> ucb;
      UctTest::read(ref.get_root().get_ego_int_node(), &v, &g, &n, &ucb);
      const auto d = dev.get_root_statistic(1, 0);
      bool ok = n == d.total_node_visits && std::memcmp(&v, &d.value, sizeof(double)) == 0 && ucb.size() == d.ucb_statistics.size() &&
                dev.returnBestAction() == ref.returnBestAction();
      for (const auto& kv : ucb) {
        auto it = d.ucb_statistics.find(kv.first);
        ok = ok && it != d.ucb_statistics.end() && it->second.first == kv.second.first &&
             std::memcmp(&it->second.second, &kv.second.second, sizeof(double)) == 0;
      }
      return ok;
    };
    {
      // created under the DEFAULT parameters first, then the parameters change: the search is re-created
      CrossingFixture fx(1);   // registers the defaults
      DeviceRootParallelMcts<CrossingState<int>> dev(p, 2, 0, 0, MAMCTS_SRC_REFERENCE_CONST);
      auto st0 = fx.state();
      dev.search(st0);
      CudaState<CrossingState<int>>::set_parameters(cp);
      dev.search(state);
      const bool ok_global = same_root(dev);
      // parameters handed to the constructor win over whatever is registered globally
      CrossingFixture fx2(1);  // registers the defaults again
      const mamcts_crossing_params env = CudaState<CrossingState<int>>::to_params(cp);
      DeviceRootParallelMcts<CrossingState<int>> dev2(p, 2, 0, 0, MAMCTS_SRC_REFERENCE_CONST, 0, 1, &env);
      dev2.search(state);
      CHECK_MSG(ok_global && same_root(dev2), "device_mcts_non_default_parameters_follow_the_state");
      // a state that does not fit the registered parameters is refused, never simulated under the wrong ones
      bool threw = false;
      try { DeviceRootParallelMcts<CrossingState<int>> dev3(p, 2, 0, 0, MAMCTS_SRC_REFERENCE_CONST); dev3.search(state); }
      catch (const std::logic_error&) { threw = true; }
      CHECK_MSG(threw, "device_mcts_mismatching_parameters_are_refused");
    }
    {  // terminal root (ego already past the goal): the reference does not expand it (stage_node.h:183-187)
      CudaState<CrossingState<int>>::set_parameters(cp);
      std::vector<Reward> rewards; EgoCosts cost;
      std::shared_ptr<CrossingState<int>> s = std::make_shared<CrossingState<int>>(state);
      for (int k = 0; k < 40 && !s->is_terminal(); ++k) s = s->execute(JointAction{4, 0}, rewards, cost);
      bool threw = false;
      try { DeviceRootParallelMcts<CrossingState<int>> dev(p, 2, 0, 0, MAMCTS_SRC_REFERENCE_CONST); dev.search(*s); }
      catch (const std::logic_error&) { threw = true; }
      CHECK_MSG(s->is_terminal() && threw, "device_mcts_terminal_root_is_refused");
    }
  }
  std::printf("%s (%d failed)\n", g_failures ? "DROPIN FAILED" : "DROPIN OK", g_failures);
  return g_failures;
}
