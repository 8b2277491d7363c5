// module.cpp - the pybind11 module `mamcts_cuda`: everything the reference's own `mamcts` module defines
// (python/bindings/module.cpp:12-15 - define_mamcts and define_environments, compiled UNMODIFIED from
// python/bindings/define_mamcts.cpp and define_environments.cpp where they lie in the reference checkout)
// plus, next to the reference's MctsCrossingStateIntUctUct (define_mamcts.cpp:223-229), the CUDA-backed
// searches of this repository:
//
//   MctsCrossingStateIntCudaUctCudaUct      Mcts<CrossingState<int>, CudaUctStatistic, CudaUctStatistic,
//                                                CudaRandomHeuristic>: the reference's own Mcts<> with the
//                                                drop-in heuristic and statistic (rollouts and back-propagation
//                                                through the C ABI)
//   MctsCrossingStateIntUctUctCuda          the same with the reference's UctStatistic, CUDA rollouts only
//   MctsCrossingStateIntUctUctRandom        the stock instantiation (for side-by-side checks from Python)
//   DeviceRootParallelMctsCrossingStateInt  the device-resident root-parallel search (mamcts_search_*)
//
// so a Python user of the reference switches with `import mamcts_cuda as mamcts`. Built by
// `make -C oracle pybind` into oracle/_ref/ (it includes the reference's headers, which do not travel to the
// GPU box; the built module does).
#include <memory>
#include <unordered_map>

#include "python/bindings/define_environments.hpp"
#include "python/bindings/define_mamcts.hpp"

#include "environments/crossing_state.h"
#include "mcts/heuristics/random_heuristic.h"
#include "mcts/mcts.h"
#include "mcts/statistics/uct_statistic.h"

#include "batched_mcts.h"
#include "cuda_random_heuristic.h"
#include "cuda_state.h"
#include "cuda_uct_statistic.h"
#include "device_mcts.h"

namespace py = pybind11;
using namespace mcts;

namespace {

template <class M>
void define_mcts_surface(py::module& m, const char* name) {
  // the public surface of Mcts<> (mcts/mcts.h:43-76)
  py::class_<M, std::shared_ptr<M>>(m, name)
      .def(py::init<const MctsParameters&>())
      .def("search", [](M& self, const CrossingState<int>& state) { self.search(state); })
      .def("returnBestAction", [](M& self) { return self.returnBestAction(); })
      .def("numIterations", &M::numIterations)
      .def("numNodes", &M::numNodes)
      .def("searchTime", &M::searchTime)
      .def("root_ego_policy", [](M& self) { return self.get_root().get_ego_int_node().get_policy(); })
      .def("root_ego_counts", [](M& self) {
        std::unordered_map<ActionIdx, unsigned> counts;
        for (const auto& kv : self.get_root().get_ego_int_node().get_ucb_statistics()) counts[kv.first] = kv.second.action_count_;
        return counts;
      })
      .def("__repr__", [name](const M&) { return std::string("mamcts_cuda.") + name; });
}

void define_mamcts_cuda(py::module m) {
  m.def("MctsDefaultParameters", &mcts_default_parameters);   // mcts/mcts_parameters.h:97-135
  // the environment parameters the CUDA engine simulates (CrossingState keeps its own private)
  m.def("set_crossing_state_parameters", [](const CrossingStateParameters<int>& p) {
    CudaState<CrossingState<int>>::set_parameters(p);
  });
  // action source of CudaRandomHeuristic: "reference" = the stock deterministic rollouts (bit-identical trees),
  // "philox" = uniform Philox4x32-10 joint actions, K rollouts per leaf
  m.def("set_cuda_heuristic_options", [](const std::string& source, unsigned rollouts_per_leaf, uint64_t philox_seed, unsigned tree_id) {
    auto& o = CudaRandomHeuristic::options();
    if (source == "reference") o.action_source = MAMCTS_SRC_REFERENCE_CONST;
    else if (source == "philox") o.action_source = MAMCTS_SRC_PHILOX;
    else throw std::invalid_argument("source must be 'reference' or 'philox'");
    o.rollouts_per_leaf = rollouts_per_leaf;
    o.philox_seed = philox_seed;
    o.tree_id = tree_id;
  }, py::arg("source") = "reference", py::arg("rollouts_per_leaf") = 1, py::arg("philox_seed") = 1000, py::arg("tree_id") = 0);

  define_mcts_surface<Mcts<CrossingState<int>, CudaUctStatistic, CudaUctStatistic, CudaRandomHeuristic>>(
      m, "MctsCrossingStateIntCudaUctCudaUct");
  define_mcts_surface<Mcts<CrossingState<int>, UctStatistic, UctStatistic, CudaRandomHeuristic>>(
      m, "MctsCrossingStateIntUctUctCuda");
  define_mcts_surface<Mcts<CrossingState<int>, UctStatistic, UctStatistic, RandomHeuristic>>(
      m, "MctsCrossingStateIntUctUctRandom");

  using Dev = DeviceRootParallelMcts<CrossingState<int>>;
  py::class_<Dev, std::shared_ptr<Dev>>(m, "DeviceRootParallelMctsCrossingStateInt")
      .def(py::init([](const MctsParameters& p, unsigned num_trees, uint64_t philox_seed, unsigned first_tree_id,
                       const std::string& source, unsigned max_inner, unsigned rollouts_per_leaf,
                       const CrossingStateParameters<int>* env) {
             const int src = source == "reference" ? MAMCTS_SRC_REFERENCE_CONST : MAMCTS_SRC_PHILOX;
             if (source != "reference" && source != "philox") throw std::invalid_argument("source must be 'reference' or 'philox'");
             if (env) {
               const mamcts_crossing_params ep = CudaState<CrossingState<int>>::to_params(*env);
               return std::make_shared<Dev>(p, num_trees, philox_seed, first_tree_id, src, max_inner, rollouts_per_leaf, &ep);
             }
             return std::make_shared<Dev>(p, num_trees, philox_seed, first_tree_id, src, max_inner, rollouts_per_leaf);
           }),
           py::arg("mcts_parameters"), py::arg("num_trees"), py::arg("philox_seed") = 1000, py::arg("first_tree_id") = 0,
           py::arg("source") = "philox", py::arg("max_inner_nodes_per_tree") = 0, py::arg("rollouts_per_leaf") = 1,
           py::arg("crossing_state_parameters") = static_cast<const CrossingStateParameters<int>*>(nullptr))
      .def("search", [](Dev& self, const CrossingState<int>& state) { self.search(state); })
      .def("returnBestAction", &Dev::returnBestAction)
      .def("numIterations", &Dev::numIterations)
      .def("numNodes", &Dev::numNodes)
      .def("numEnvSteps", &Dev::numEnvSteps)
      .def("numTrees", &Dev::numTrees)
      .def("searchTime", &Dev::searchTime)
      .def("merged_root_statistic", [](const Dev& self) {
        const auto s = self.get_merged_root_statistic();
        py::dict d, counts, values;
        for (const auto& kv : s.ucb_statistics) { counts[py::int_(kv.first)] = kv.second.first; values[py::int_(kv.first)] = kv.second.second; }
        d["value"] = s.value; d["total_node_visits"] = s.total_node_visits; d["action_counts"] = counts; d["action_values"] = values;
        return d;
      })
      .def("root_statistic", [](const Dev& self, unsigned tree, unsigned agent_position) {
        const auto s = self.get_root_statistic(tree, agent_position);
        py::dict d, counts, values;
        for (const auto& kv : s.ucb_statistics) { counts[py::int_(kv.first)] = kv.second.first; values[py::int_(kv.first)] = kv.second.second; }
        d["value"] = s.value; d["total_node_visits"] = s.total_node_visits; d["action_counts"] = counts; d["action_values"] = values;
        return d;
      }, py::arg("tree"), py::arg("agent_position") = 0)
      .def("printTreeToDotFile", &Dev::printTreeToDotFile, py::arg("filename") = "tree", py::arg("tree") = 0, py::arg("max_depth") = 5)
      .def("__repr__", [](const Dev&) { return "mamcts_cuda.DeviceRootParallelMctsCrossingStateInt"; });
}

}  // namespace

PYBIND11_MODULE(mamcts_cuda, m) {
  define_mamcts(m);         // the reference's own definitions, unmodified
  define_environments(m);
  define_mamcts_cuda(m);
}
