// device_mcts.h - DeviceRootParallelMcts<S>: the reference's Mcts<S, UctStatistic, UctStatistic,
// RandomHeuristic> surface (mcts/mcts.h:43-76: search, returnBestAction, numIterations, numNodes,
// searchTime) over the DEVICE-RESIDENT search of the C ABI (mamcts_search_*): num_trees independent trees
// live in HBM and whole iterations (select/expand, rollout, back-propagation) run on the GPU; the host
// sends the decision state and reads root statistics. This is the root-parallel mode of the reference
// (NUM_PARALLEL_MCTS trees merged by UctStatistic::merge_node_statistics, mcts.h:188-221,270-292) with
// thousands of trees instead of a handful of threads.
//
//   DeviceRootParallelMcts<CrossingState<int>> mcts(params, /*num_trees=*/65536);
//   mcts.search(state);
//   ActionIdx a = mcts.returnBestAction();
//
// Trees are the same trees BatchedRootParallelMcts builds on the host with the reference's own
// StageNode code (same Philox streams: tree t, evaluation c -> (first_tree_id + t, c)), which
// tests/dropin/dropin_test.cc checks statistic by statistic. With MAMCTS_SRC_REFERENCE_CONST every tree
// is the stock reference's tree, bit for bit.
#ifndef MAMCTS_B200_HOST_DEVICE_MCTS_H_
#define MAMCTS_B200_HOST_DEVICE_MCTS_H_

#include <chrono>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <cmath>
#include <vector>

#include "cuda_random_heuristic.h"  // CudaEngine, to_mamcts_params
#include "cuda_state.h"
#include "mamcts_b200.h"
#include "mcts/mcts_parameters.h"
#include "mcts/state.h"

namespace mcts {

template <class S>
class DeviceRootParallelMcts {
 public:
  // One agent's root statistic in UctStatistic's vocabulary (uct_statistic.h:272-289).
  struct RootStatistic {
    double value = 0.0;                                              // value_
    unsigned total_node_visits = 0;                                  // total_node_visits_
    std::map<ActionIdx, std::pair<unsigned, double>> ucb_statistics;  // action -> (action_count_, action_value_)
  };

  // env_parameters: the environment's parameters in the C ABI's form (CudaState<S>::to_params(...)); nullptr =
  // whatever is registered with CudaState<S>::set_parameters at the time of each search(). Either way the
  // device search is re-created when the parameters (or the shape of the state) differ from the ones it was
  // created with - a later set_parameters() is never silently ignored.
  DeviceRootParallelMcts(const MctsParameters& mcts_parameters, unsigned num_trees, uint64_t philox_seed = 1000,
                         unsigned first_tree_id = 0, int action_source = MAMCTS_SRC_PHILOX,
                         unsigned max_inner_nodes_per_tree = 0, unsigned rollouts_per_leaf = 1,
                         const typename CudaState<S>::EnvParams* env_parameters = nullptr)
      : mcts_parameters_(mcts_parameters), num_trees_(num_trees), philox_seed_(philox_seed),
        first_tree_id_(first_tree_id), action_source_(action_source), max_inner_(max_inner_nodes_per_tree),
        rollouts_per_leaf_(rollouts_per_leaf),
        search_(nullptr), num_iterations_(0), search_time_(0), merged_valid_(false),
        own_env_(env_parameters != nullptr) {
    std::memset(&created_env_, 0, sizeof(created_env_));
    if (env_parameters) env_ = *env_parameters;
    CudaEngine::get();
  }
  ~DeviceRootParallelMcts() { if (search_) mamcts_search_destroy(search_); }
  DeviceRootParallelMcts(const DeviceRootParallelMcts&) = delete;
  DeviceRootParallelMcts& operator=(const DeviceRootParallelMcts&) = delete;

  // Mcts::search (mcts.h:133-141) for every tree, bounded by MAX_NUMBER_OF_ITERATIONS only
  // (the wall-clock limit MAX_SEARCH_TIME has no device counterpart).
  void search(const S& current_state) {
    using CS = CudaState<S>;
    const auto start = std::chrono::high_resolution_clock::now();
    int32_t x[MAMCTS_MAX_AGENTS] = {0}, last[MAMCTS_MAX_AGENTS] = {0};
    uint8_t flags = 0;
    if (!own_env_) env_ = CS::params();
    CS::check_state(current_state, env_);
    CS::pack(current_state, x, last, &flags);
    if (flags & 1u) {
      // the reference returns from select_or_expand without expanding a terminal root (stage_node.h:183-187):
      // there is nothing to search and no best action to report
      merged_valid_ = false;
      throw std::logic_error("DeviceRootParallelMcts::search: the state is terminal");
    }
    if (search_ && std::memcmp(&created_env_, &env_, sizeof(env_)) != 0) {   // parameters changed since creation
      mamcts_search_destroy(search_);
      search_ = nullptr;
    }
    num_agents_ = CS::num_agents(current_state);
    num_ego_actions_ = (unsigned)current_state.get_num_actions(current_state.get_ego_agent_idx());
    agent_ids_.assign(1, current_state.get_ego_agent_idx());
    num_actions_.assign(1, num_ego_actions_);
    for (auto ai : current_state.get_other_agent_idx()) {
      agent_ids_.push_back(ai);
      num_actions_.push_back((unsigned)current_state.get_num_actions(ai));
    }
    if (!search_) {
      mamcts_search_config cfg;
      std::memset(&cfg, 0, sizeof(cfg));
      cfg.env = CS::kEnv;
      cfg.action_source = action_source_;
      cfg.num_trees = num_trees_;
      cfg.first_tree_id = first_tree_id_;
      cfg.rollouts_per_leaf = rollouts_per_leaf_;   // 1 = the reference's heuristic; 2..16 = leaf-batched mean
      cfg.layout = MAMCTS_LAYOUT_AUTO;
      cfg.max_inner_nodes_per_tree = max_inner_;
      cfg.philox_seed = philox_seed_;
      cfg.mcts = to_mamcts_params(mcts_parameters_);
      fill_env(cfg);
      CudaEngine::check(mamcts_search_create(CudaEngine::get(), &cfg, x, last, &search_), "mamcts_search_create");
      created_env_ = env_;
    } else {
      CudaEngine::check(mamcts_search_reset(search_, x, last), "mamcts_search_reset");
    }
    CudaEngine::check(mamcts_search_run(search_, mcts_parameters_.MAX_NUMBER_OF_ITERATIONS), "mamcts_search_run");
    merge();
    num_iterations_ = mcts_parameters_.MAX_NUMBER_OF_ITERATIONS;
    search_time_ = (unsigned)std::chrono::duration_cast<std::chrono::milliseconds>(
                       std::chrono::high_resolution_clock::now() - start).count();
  }

  // get_best_action of the merged ego root statistic (uct_statistic.h:78-89 after :165-184)
  ActionIdx returnBestAction() const { require_search(); return merged_best_; }
  unsigned numIterations() const { return num_iterations_; }   // per tree
  unsigned searchTime() const { return search_time_; }         // ms, like Mcts::searchTime
  unsigned numTrees() const { return num_trees_; }
  uint64_t numNodes() const {                                  // over all trees
    require_search();
    uint64_t nodes = 0;
    CudaEngine::check(mamcts_search_counters(search_, nullptr, nullptr, nullptr, &nodes, nullptr), "mamcts_search_counters");
    return nodes;
  }
  uint64_t numEnvSteps() const {                               // execute() calls: rollouts + expansions
    require_search();
    uint64_t r = 0, x = 0;
    CudaEngine::check(mamcts_search_counters(search_, nullptr, &r, &x, nullptr, nullptr), "mamcts_search_counters");
    return r + x;
  }

  // The merged ego root statistic (what Mcts::get_root().get_ego_int_node() holds after parallel_search).
  RootStatistic get_merged_root_statistic() const {
    require_search();
    RootStatistic s;
    s.value = merged_value_;
    s.total_node_visits = (unsigned)merged_visits_;
    for (unsigned a = 0; a < num_ego_actions_; ++a)
      if (merged_occ_[a]) s.ucb_statistics[a] = {(unsigned)merged_n_[a], merged_q_[a]};
    return s;
  }

  // Root statistic of one tree; agent_position 0 = ego, 1.. = others in get_other_agent_idx() order.
  RootStatistic get_root_statistic(unsigned tree, unsigned agent_position) const {
    require_search();
    double value = 0.0, q[MAMCTS_MAX_ACTIONS];
    uint32_t visits = 0, n[MAMCTS_MAX_ACTIONS], mask = 0, nodes = 0;
    CudaEngine::check(mamcts_search_root_stats(search_, tree, agent_position, &value, &visits, q, n, &mask, &nodes),
                      "mamcts_search_root_stats");
    RootStatistic s;
    s.value = value;
    s.total_node_visits = visits;
    for (unsigned a = 0; a < MAMCTS_MAX_ACTIONS; ++a)
      if ((mask >> a) & 1u) s.ucb_statistics[a] = {n[a], q[a]};
    return s;
  }

  // Mcts::printTreeToDotFile (mcts.h:370-372 -> StageNode::printTree / printLayer, stage_node.h:322-391) for
  // one device tree: the same Graphviz text - a dotted cluster per stage node with one "V=.., N=.."
  // node per agent (UctStatistic::print_node_information), one "a=.., N=.., V=.." edge per agent and
  // child (print_edge_information) - down to max_depth. Node ids are depth-first indices (the
  // reference numbers nodes in creation order); an other agent's edge label uses its POSITION in the
  // joint action (the reference indexes the joint action by agent id there, which only coincides for
  // CrossingState).
  void printTreeToDotFile(std::string filename = "tree", unsigned tree = 0, unsigned max_depth = 5) const {
    require_search();
    uint32_t nodes = 0;
    {
      double v; uint32_t vis, mask; double q[MAMCTS_MAX_ACTIONS]; uint32_t n[MAMCTS_MAX_ACTIONS];
      CudaEngine::check(mamcts_search_root_stats(search_, tree, 0, &v, &vis, q, n, &mask, &nodes), "mamcts_search_root_stats");
    }
    unsigned max_actions = 0;
    for (unsigned na : num_actions_) max_actions = std::max(max_actions, na);
    const size_t stride = 4 + (size_t)num_agents_ * (3 + 2 * max_actions);
    std::vector<double> rec((size_t)nodes * stride);
    uint32_t count = 0;
    CudaEngine::check(mamcts_search_dump_tree(search_, tree, rec.data(), nodes, &count), "mamcts_search_dump_tree");
    std::ofstream out(filename + ".gv");
    out << "digraph G {" << std::endl << "labelloc = \"t\";" << std::endl;
    auto stat = [&](uint32_t i, unsigned p) { return rec.data() + (size_t)i * stride + 4 + (size_t)p * (3 + 2 * max_actions); };
    auto node_label = [&](uint32_t i, unsigned p) {
      std::stringstream ss;
      ss << std::setprecision(2) << "V=" << stat(i, p)[1] << ", N=" << (unsigned)stat(i, p)[0];
      return ss.str();
    };
    std::vector<uint32_t> path;  // depth-first: path[d] = record index of the current ancestor at depth d
    for (uint32_t i = 0; i < count; ++i) {
      const double* r = rec.data() + (size_t)i * stride;
      const unsigned depth = (unsigned)r[0];
      path.resize(depth + 1);
      path[depth] = i;
      if (depth > max_depth) continue;
      const uint32_t id = i + 1;
      out << "subgraph cluster_node_" << id << "{" << std::endl;
      for (unsigned p = 0; p < num_agents_; ++p)
        out << "node" << id << "_" << int(agent_ids_[p]) << "[label=\"" << node_label(i, p) << " \n Ag." << int(agent_ids_[p])
            << "\"]" << ";" << std::endl;
      out << "label= \"ID " << id << "\";" << std::endl << "graph[style=dotted]; }" << std::endl;
      if (depth == 0) continue;
      // the edges from the parent (line order is irrelevant to Graphviz; the reference prints them after
      // the child's subtree)
      const uint32_t parent = path[depth - 1];
      double code = r[1];
      for (unsigned p = 0; p < num_agents_; ++p) {
        const unsigned a = (unsigned)std::fmod(code, (double)max_actions);
        code = std::floor(code / max_actions);
        std::stringstream lab;
        if (stat(parent, p)[3 + a] >= 0.0)
          lab << std::setprecision(2) << "a=" << int(a) << ", N=" << (unsigned)stat(parent, p)[3 + a] << ", V="
              << stat(parent, p)[3 + max_actions + a];
        out << "node" << (parent + 1) << "_" << int(agent_ids_[p]) << " -> " << "node" << id << "_" << int(agent_ids_[p])
            << "[label=\"" << lab.str() << "\"]" << ";" << std::endl;
      }
    }
    out << "}" << std::endl;
  }

 private:
  void require_search() const {
    if (!search_ || !merged_valid_) {
      std::cerr << "DeviceRootParallelMcts: search() has not been called" << std::endl;
      std::terminate();
    }
  }
  void merge() {
    CudaEngine::check(mamcts_search_merge_roots(search_, merged_q_, merged_n_, merged_occ_, &merged_visits_,
                                                &merged_value_, &merged_best_),
                      "mamcts_search_merge_roots");
    merged_valid_ = true;
  }
  void fill_env(mamcts_search_config& cfg) {
    if constexpr (CudaState<S>::kEnv == MAMCTS_ENV_CROSSING_I32) cfg.crossing = env_;
    else cfg.simple = env_;
  }

  const MctsParameters mcts_parameters_;
  unsigned num_trees_;
  uint64_t philox_seed_;
  unsigned first_tree_id_;
  int action_source_;
  unsigned rollouts_per_leaf_ = 1;
  unsigned max_inner_;
  mamcts_search* search_;
  unsigned num_iterations_, search_time_;
  unsigned num_agents_ = 0, num_ego_actions_ = 0;
  std::vector<AgentIdx> agent_ids_;      // positional: ego, then get_other_agent_idx() order
  std::vector<unsigned> num_actions_;
  bool merged_valid_;
  bool own_env_;
  typename CudaState<S>::EnvParams env_{}, created_env_{};
  double merged_q_[MAMCTS_MAX_ACTIONS] = {0.0}, merged_value_ = 0.0;
  uint64_t merged_n_[MAMCTS_MAX_ACTIONS] = {0}, merged_visits_ = 0;
  uint32_t merged_occ_[MAMCTS_MAX_ACTIONS] = {0}, merged_best_ = 0;
};

}  // namespace mcts
#endif
