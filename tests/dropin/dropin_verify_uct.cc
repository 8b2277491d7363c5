// tests/dropin/dropin_verify_uct.cc - the reference's back-propagation invariant on trees whose
// rollouts came from the CUDA engine.
//
// 1. The reference's OWN checker UctTest::verify_uct (reference test/uct/uct_test_class.h:23-215)
//    run exactly as the reference's test does (test/uct/uct_test.cc:22-52: SimpleState(4), 10 000
//    iterations, 100 nodes) with CudaRandomHeuristic plugged into Mcts<>.
//    verify_uct is only sound on the trees that test produces: it looks actions up with
//    std::find over the WHOLE joint action (uct_test_class.h:44-47) and indexes other agents'
//    actions by their position among the others, i.e. one slot too low (:96-101, :160), which is
//    harmless only while every child has a "diagonal" joint action (a, a) - the case under the
//    reference's deterministic rollouts. So it runs with the REFERENCE_CONST source.
//    The same invariant stated positionally, for PHILOX trees, is check 6 of dropin_test.cc.
// Built by oracle/Makefile (target dropin); run on the GPU box by tests/test_gpu_dropin.py.
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

// The parameter set of the reference's UCT tests (test/uct/uct_test.cc:22-40: gamma 0.9, seed 1000,
// 10 000 iterations, 100 nodes, bounds [-1000, 100], c = 0.7), with both wall-clock limits
// neutralised and the rollout depth capped (its rollouts never terminate, F1).
static MctsParameters uct_test_parameters() {
  MctsParameters p(mcts_default_parameters());
  p.DISCOUNT_FACTOR = 0.9;          p.RANDOM_SEED = 1000;
  p.MAX_NUMBER_OF_ITERATIONS = 10000; p.MAX_NUMBER_OF_NODES = 100;
  p.MAX_SEARCH_DEPTH = 1000;        p.NUM_PARALLEL_MCTS = 1;
  p.MAX_SEARCH_TIME = 1000000000;   p.random_heuristic.MAX_SEARCH_TIME = 1e9;
  p.random_heuristic.MAX_NUMBER_OF_ITERATIONS = 40;
  p.uct_statistic.LOWER_BOUND = -1000; p.uct_statistic.UPPER_BOUND = 100;
  p.uct_statistic.EXPLORATION_CONSTANT = 0.7;
  return p;
}

// ---- 1. the reference's own test with the CUDA heuristic -----------------------------------------
TEST(dropin_mcts, reference_verify_uct_simple_state_cuda_heuristic) {
  CudaRandomHeuristic::options() = CudaRandomHeuristic::Options();  // REFERENCE_CONST
  Mcts<SimpleState, UctStatistic, UctStatistic, CudaRandomHeuristic> mcts(uct_test_parameters());
  SimpleState state(4);
  mcts.search(state);
  EXPECT_EQ(mcts.numIterations(), 10000u);
  UctTest test;
  test.verify_uct(mcts, 1000);
}

// small_search_depth (test/uct/uct_test.cc:54-64) with the CUDA heuristic
TEST(dropin_mcts, reference_small_search_depth_cuda_heuristic) {
  CudaRandomHeuristic::options() = CudaRandomHeuristic::Options();
  auto params = uct_test_parameters();
  params.MAX_SEARCH_DEPTH = 2;
  Mcts<SimpleState, UctStatistic, UctStatistic, CudaRandomHeuristic> mcts(params);
  SimpleState state(4);
  mcts.search(state);
  EXPECT_EQ(mcts.numIterations(), 10000u);
  UctTest test;  // as in the reference: the depth-limited search only has to run
}

int main(int argc, char** argv) {
  ::testing::InitGoogleTest(&argc, argv);
  return RUN_ALL_TESTS();
}
