// tests/dropin/dropin_verify_uct.cc - runs the REFERENCE'S OWN back-propagation invariant checker
// (UctTest::verify_uct, reference test/uct/uct_test_class.h:23-215: for every node, recompute the
// expected visit counts and Q(s,a) = sum_child n_child (r + gamma V_child) / n_a from the children)
// on trees whose rollouts came from the CUDA engine through CudaRandomHeuristic. Mirrors the
// reference's test/uct/uct_test.cc:42-52. Built by oracle/Makefile (target dropin).
#include "gtest/gtest.h"

#define UNIT_TESTING
#include "test/uct/uct_test_class.h"
#include "mcts/heuristics/random_heuristic.h"
#include "mcts/statistics/uct_statistic.h"
#include "environments/crossing_state.h"
#include "test/uct/simple_state.h"

#include "cuda_random_heuristic.h"
#include "cuda_state.h"

using namespace std;
using namespace mcts;

// reference test/uct/uct_test.cc:22-40, wall-clock limits neutralised
MctsParameters default_uct_params() {
  MctsParameters parameters(mcts_default_parameters());
  parameters.DISCOUNT_FACTOR = 0.9;
  parameters.RANDOM_SEED = 1000;
  parameters.MAX_NUMBER_OF_ITERATIONS = 2000;
  parameters.MAX_SEARCH_TIME = 1000000000;
  parameters.MAX_SEARCH_DEPTH = 1000;
  parameters.MAX_NUMBER_OF_NODES = 100;
  parameters.NUM_PARALLEL_MCTS = 1;
  parameters.random_heuristic.MAX_SEARCH_TIME = 1e9;
  parameters.random_heuristic.MAX_NUMBER_OF_ITERATIONS = 40;
  parameters.uct_statistic.LOWER_BOUND = -1000;
  parameters.uct_statistic.UPPER_BOUND = 100;
  parameters.uct_statistic.EXPLORATION_CONSTANT = 0.7;
  return parameters;
}

TEST(dropin_mcts, verify_uct_simple_state_philox_rollouts) {
  CudaRandomHeuristic::options().action_source = MAMCTS_SRC_PHILOX;
  CudaRandomHeuristic::options().rollouts_per_leaf = 4;
  Mcts<SimpleState, UctStatistic, UctStatistic, CudaRandomHeuristic> mcts(default_uct_params());
  SimpleState state(4);
  mcts.search(state);
  UctTest test;
  test.verify_uct(mcts, 1000);
}

TEST(dropin_mcts, verify_uct_crossing_state_philox_rollouts) {
  CudaRandomHeuristic::options().action_source = MAMCTS_SRC_PHILOX;
  CudaRandomHeuristic::options().rollouts_per_leaf = 1;
  auto cparams = default_crossing_state_parameters<int>();
  cparams.NUM_OTHER_AGENTS = 2;
  CudaState<CrossingState<int>>::set_parameters(cparams);
  std::unordered_map<AgentIdx, HypothesisId> hyp{{1, 0}, {2, 0}};
  auto params = default_uct_params();
  params.MAX_NUMBER_OF_NODES = 100000;
  params.MAX_NUMBER_OF_ITERATIONS = 1500;
  Mcts<CrossingState<int>, UctStatistic, UctStatistic, CudaRandomHeuristic> mcts(params);
  CrossingState<int> state(hyp, cparams);
  mcts.search(state);
  UctTest test;
  test.verify_uct(mcts, 1000);
}

int main(int argc, char** argv) {
  ::testing::InitGoogleTest(&argc, argv);
  return RUN_ALL_TESTS();
}
