// tests/dropin/dropin_bench.cc - how fast is the reference's OWN host tree code when its rollouts (and
// back-propagation) go through the C ABI? Three drivers over the same C2-shaped search
// (CrossingState<int>, 2 agents, Philox rollouts, depth cap 50):
//   1. Mcts<.., UctStatistic, UctStatistic, RandomHeuristic>        the stock CPU path, one tree
//   2. Mcts<.., UctStatistic, UctStatistic, CudaRandomHeuristic>    one synchronous launch per leaf
//   3. BatchedRootParallelMcts<..>                                  one launch per wave of T leaves
// Prints iterations/s; numbers are quoted in INTEGRATION.md. Built by oracle/Makefile (target dropin).
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "mcts/mcts.h"
#include "mcts/heuristics/random_heuristic.h"
#include "mcts/statistics/uct_statistic.h"
#include "environments/crossing_state.h"

#include "batched_mcts.h"
#include "cuda_random_heuristic.h"
#include "cuda_state.h"
#include "cuda_uct_statistic.h"

using namespace mcts;
using Clock = std::chrono::steady_clock;

static double seconds_since(Clock::time_point t0) {
  return std::chrono::duration<double>(Clock::now() - t0).count();
}

int main(int argc, char** argv) {
  const unsigned iterations = argc > 1 ? std::atoi(argv[1]) : 2000;
  const unsigned trees = argc > 2 ? std::atoi(argv[2]) : 4096;
  const bool only_threads = argc > 3 && std::atoi(argv[3]) == 1;   // only the all-host-threads batched driver
  MctsParameters p = mcts_default_parameters();
  p.MAX_NUMBER_OF_ITERATIONS = iterations;
  p.MAX_NUMBER_OF_NODES = 100000000;
  p.MAX_SEARCH_TIME = 1000000000;
  p.random_heuristic.MAX_SEARCH_TIME = 1e9;
  p.random_heuristic.MAX_NUMBER_OF_ITERATIONS = 50;
  auto cparams = default_crossing_state_parameters<int>();
  cparams.NUM_OTHER_AGENTS = 1;
  CudaState<CrossingState<int>>::set_parameters(cparams);
  std::unordered_map<AgentIdx, HypothesisId> hyp{{1, 0}};
  CrossingState<int> state(hyp, cparams);
  if (!only_threads) {
    Mcts<CrossingState<int>, UctStatistic, UctStatistic, RandomHeuristic> m(p);
    auto t0 = Clock::now();
    m.search(state);
    std::printf("stock CPU, 1 tree:                 %9.0f iterations/s\n", m.numIterations() / seconds_since(t0));
  }
  if (!only_threads) {
    CudaRandomHeuristic::options() = CudaRandomHeuristic::Options();
    CudaRandomHeuristic::options().action_source = MAMCTS_SRC_PHILOX;
    Mcts<CrossingState<int>, UctStatistic, UctStatistic, CudaRandomHeuristic> m(p);
    m.search(state);  // warm-up (context, first launches)
    auto t0 = Clock::now();
    m.search(state);
    std::printf("drop-in heuristic, 1 leaf/launch:  %9.0f iterations/s\n", m.numIterations() / seconds_since(t0));
  }
  if (!only_threads) {
    BatchedRootParallelMcts<CrossingState<int>, UctStatistic, UctStatistic> b(p, trees, 1000);
    auto t0 = Clock::now();
    b.search(state);
    const double s = seconds_since(t0);
    std::printf("batched driver, %5u trees:       %9.0f iterations/s (%.2f s, %.3g rollout env-steps/s)\n", trees,
                (double)trees * iterations / s, s, (double)b.rolloutSteps() / s);
  }
  {
    BatchedRootParallelMcts<CrossingState<int>, UctStatistic, UctStatistic> b(p, trees, 1000, 0, 1, /*host_threads=*/0);
    auto t0 = Clock::now();
    b.search(state);
    const double s = seconds_since(t0);
    std::printf("batched driver, %5u trees, all host threads: %9.0f iterations/s (%.2f s)\n", trees,
                (double)trees * iterations / s, s);
  }
  if (!only_threads) {
    BatchedRootParallelMcts<CrossingState<int>, CudaUctStatistic, CudaUctStatistic> b(p, trees / 8, 1000);
    auto t0 = Clock::now();
    b.search(state);
    const double s = seconds_since(t0);
    std::printf("batched driver + CUDA statistic, %5u trees: %9.0f iterations/s (%.2f s)\n", trees / 8,
                (double)(trees / 8) * iterations / s, s);
  }
  return 0;
}
