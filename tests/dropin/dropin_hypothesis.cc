// tests/dropin/dropin_hypothesis.cc - C4 (BASELINE.json configs[3]) through the reference's own host
// code: Mcts<CrossingState<int>, UctStatistic, HypothesisStatistic, H>::search(state, belief_tracker)
// with H = CudaRandomHeuristic in MAMCTS_SRC_HYPOTHESIS mode, i.e. every rollout runs on the GPU with
// the other agents following the hypotheses the belief tracker sampled for the iteration. Mirrors the
// reference's closed-loop tests environments/tests/crossing_state_int_test.cc:141-229
// (mcts_goal_reached_true_hypothesis, mcts_goal_reached_wrong_hypothesis): plan, act, update the belief,
// until the ego agent reaches its goal without a collision. Wall-clock limits are replaced by an
// iteration budget. Built by oracle/Makefile (target dropin).
#include "gtest/gtest.h"

#include <random>

#include "mcts/mcts.h"
#include "mcts/heuristics/random_heuristic.h"
#include "mcts/statistics/uct_statistic.h"
#include "mcts/hypothesis/hypothesis_statistic.h"
#include "mcts/hypothesis/hypothesis_belief_tracker.h"
#include "environments/crossing_state.h"

#include "cuda_random_heuristic.h"
#include "cuda_state.h"
#include "device_hypothesis_mcts.h"
#include "device_mcts.h"

using namespace mcts;
typedef int Domain;

static MctsParameters hypothesis_params(unsigned iterations) {
  MctsParameters p = mcts_default_parameters();
  p.MAX_NUMBER_OF_ITERATIONS = iterations;
  p.MAX_SEARCH_TIME = 1000000000;
  p.random_heuristic.MAX_SEARCH_TIME = 1e9;
  p.random_heuristic.MAX_NUMBER_OF_ITERATIONS = 50;
  return p;
}

template <class H>
static bool closed_loop(const std::vector<std::pair<int, int>>& hypotheses, std::pair<int, int> true_policy_gap,
                        unsigned iterations, int max_steps, bool* collision, int* steps_taken) {
  const auto params = default_crossing_state_parameters<Domain>();
  CudaState<CrossingState<Domain>>::set_parameters(params);
  CudaState<CrossingState<Domain>>::set_hypotheses(hypotheses);
  auto mcts_params = hypothesis_params(iterations);
  HypothesisBeliefTracker belief_tracker(mcts_params);
  auto state = std::make_shared<CrossingState<Domain>>(belief_tracker.sample_current_hypothesis(), params);
  for (const auto& h : hypotheses) state->add_hypothesis(AgentPolicyCrossingState<Domain>(h, params));
  auto next_state = state;
  belief_tracker.belief_update(*state, *next_state);
  AgentPolicyCrossingState<Domain> true_agents_policy(true_policy_gap, params);
  std::vector<Reward> rewards;
  EgoCosts cost;
  *collision = false;
  int i = 0;
  for (; i < max_steps; ++i) {
    auto jointaction = JointAction(state->get_num_agents());
    Mcts<CrossingState<Domain>, UctStatistic, HypothesisStatistic, H> mcts(mcts_params);
    mcts.search(*state, belief_tracker);
    jointaction[CrossingState<Domain>::ego_agent_idx] = mcts.returnBestAction();
    for (auto agent_idx : state->get_other_agent_idx()) {
      const auto action = true_agents_policy.act(state->get_agent_state(agent_idx), state->get_ego_state());
      jointaction[agent_idx] = aconv<Domain>(action);
    }
    next_state = state->execute(jointaction, rewards, cost);
    belief_tracker.belief_update(*state, *next_state);
    state = next_state;
    if (state->is_terminal() && !state->ego_goal_reached()) { *collision = true; break; }
    if (state->is_terminal()) break;
  }
  *steps_taken = i + 1;
  return state->ego_goal_reached();
}

TEST(dropin_hypothesis, mcts_goal_reached_true_hypothesis_cuda_rollouts) {
  CudaRandomHeuristic::options() = CudaRandomHeuristic::Options();
  CudaRandomHeuristic::options().action_source = MAMCTS_SRC_HYPOTHESIS;
  CudaRandomHeuristic::options().rollouts_per_leaf = 8;
  bool collision = false;
  int steps = 0;
  const bool goal = closed_loop<CudaRandomHeuristic>({{5, 5}}, {5, 5}, 1500, 100, &collision, &steps);
  std::printf("  true hypothesis: goal %d collision %d after %d steps\n", (int)goal, (int)collision, steps);
  EXPECT_TRUE(goal);
  EXPECT_FALSE(collision);
}

TEST(dropin_hypothesis, mcts_goal_reached_wrong_hypothesis_cuda_rollouts) {
  CudaRandomHeuristic::options() = CudaRandomHeuristic::Options();
  CudaRandomHeuristic::options().action_source = MAMCTS_SRC_HYPOTHESIS;
  CudaRandomHeuristic::options().rollouts_per_leaf = 8;
  bool collision = false;
  int steps = 0;
  const bool goal = closed_loop<CudaRandomHeuristic>({{-2, 1}, {2, 3}, {4, 5}}, {-2, -2}, 1500, 30, &collision, &steps);
  std::printf("  wrong hypothesis: goal %d collision %d after %d steps\n", (int)goal, (int)collision, steps);
  EXPECT_TRUE(goal);
  EXPECT_FALSE(collision);
}

// the same loops with the stock heuristic, as a sanity check of the harness itself
TEST(dropin_hypothesis, stock_heuristic_reaches_goal_too) {
  bool collision = false;
  int steps = 0;
  EXPECT_TRUE(closed_loop<RandomHeuristic>({{5, 5}}, {5, 5}, 1500, 100, &collision, &steps));
  EXPECT_FALSE(collision);
}

// Closed loop in the manner of CrossingStateEpisodeRunner::step (reference
// environments/crossing_state_episode_runner.h:51-97) with the DEVICE-RESIDENT search as the planner:
// every step 2 048 root-parallel UCT trees are searched on the GPU from the current state and the ego
// executes the merged best action. The planner's model of the other agents is the UCT one (action index
// i = velocity +i), so the true agents here move by a uniformly drawn index as well - with
// gap-keeping true agents that model is simply wrong (they never pass the crossing the ego waits at),
// which is what the hypothesis statistics above are for.
TEST(dropin_hypothesis, episode_with_device_resident_planner_reaches_goal) {
  const auto params = default_crossing_state_parameters<Domain>();
  CudaState<CrossingState<Domain>>::set_parameters(params);
  std::unordered_map<AgentIdx, HypothesisId> hyp;
  for (AgentIdx i = 1; i <= params.NUM_OTHER_AGENTS; ++i) hyp[i] = 0;
  auto state = std::make_shared<CrossingState<Domain>>(hyp, params);
  std::mt19937 true_agents_rng(12345);
  std::uniform_int_distribution<int> true_agents_action(0, (int)params.NUM_OTHER_ACTIONS - 1);
  auto mcts_params = hypothesis_params(400);
  mcts_params.MAX_NUMBER_OF_NODES = 100000000;
  DeviceRootParallelMcts<CrossingState<Domain>> planner(mcts_params, 2048, 1000);
  std::vector<Reward> rewards;
  EgoCosts cost;
  bool collision = false;
  int i = 0;
  for (; i < 40; ++i) {
    planner.search(*state);
    auto jointaction = JointAction(state->get_num_agents());
    jointaction[CrossingState<Domain>::ego_agent_idx] = planner.returnBestAction();
    for (auto agent_idx : state->get_other_agent_idx())
      jointaction[agent_idx] = (ActionIdx)true_agents_action(true_agents_rng);
    state = state->execute(jointaction, rewards, cost);
    if (state->is_terminal()) { collision = !state->ego_goal_reached(); break; }
  }
  std::printf("  device planner: goal %d collision %d after %d steps, %u iterations x %u trees per step\n",
              (int)state->ego_goal_reached(), (int)collision, i + 1, planner.numIterations(), planner.numTrees());
  EXPECT_TRUE(state->ego_goal_reached());
  EXPECT_FALSE(collision);
}

// The device-resident hypothesis-MCTS as the planner of the reference's own closed-loop tests
// (crossing_state_int_test.cc:141-229): at every step the device search (one tree, reference-identical
// generators) must choose the action, build as many nodes and leave the belief tracker in the state the
// reference's own search does - then the loop goes on with the device planner's action.
static bool closed_loop_device_vs_stock(const std::vector<std::pair<int, int>>& hypotheses, std::pair<int, int> true_gap,
                                        unsigned iterations, int max_steps, bool* collision, int* mismatches) {
  const auto params = default_crossing_state_parameters<Domain>();
  CudaState<CrossingState<Domain>>::set_parameters(params);
  CudaState<CrossingState<Domain>>::set_hypotheses(hypotheses);
  auto mcts_params = hypothesis_params(iterations);
  mcts_params.MAX_NUMBER_OF_NODES = 100000000;
  HypothesisBeliefTracker tracker_dev(mcts_params), tracker_ref(mcts_params);
  auto state = std::make_shared<CrossingState<Domain>>(tracker_dev.sample_current_hypothesis(), params);
  for (const auto& h : hypotheses) state->add_hypothesis(AgentPolicyCrossingState<Domain>(h, params));
  tracker_dev.belief_update(*state, *state);
  AgentPolicyCrossingState<Domain> true_agents_policy(true_gap, params);
  DeviceHypothesisMcts planner(mcts_params, 1);
  std::vector<Reward> rewards;
  EgoCosts cost;
  *collision = false;
  *mismatches = 0;
  for (int i = 0; i < max_steps; ++i) {
    // the reference's own search from the same state with a tracker in the same condition
    tracker_ref = tracker_dev;
    auto ref_state = state->change_belief_reference(tracker_ref.sample_current_hypothesis());   // no beliefs change
    tracker_ref = tracker_dev;   // undo the sample: both trackers start the search alike
    // (change_belief_reference keeps a reference to tracker_ref's map, which assignment updates in place)
    Mcts<CrossingState<Domain>, UctStatistic, HypothesisStatistic, RandomHeuristic> ref(mcts_params);
    ref.search(*ref_state, tracker_ref);
    planner.search(*state, tracker_dev);
    const auto root = planner.get_root_statistic(0);
    bool same = planner.returnBestAction() == ref.returnBestAction() && planner.numNodes() == ref.numNodes() &&
                root.total_node_visits == iterations;
    // both trackers must be in the same state afterwards: their next samples agree
    HypothesisBeliefTracker a = tracker_dev, b = tracker_ref;
    for (int k = 0; k < 8 && same; ++k) same = a.sample_current_hypothesis() == b.sample_current_hypothesis();
    if (!same) ++*mismatches;
    auto jointaction = JointAction(state->get_num_agents());
    jointaction[CrossingState<Domain>::ego_agent_idx] = planner.returnBestAction();
    for (auto agent_idx : state->get_other_agent_idx())
      jointaction[agent_idx] = aconv<Domain>(true_agents_policy.act(state->get_agent_state(agent_idx), state->get_ego_state()));
    auto next_state = state->execute(jointaction, rewards, cost);
    tracker_dev.belief_update(*state, *next_state);
    state = next_state;
    if (state->is_terminal() && !state->ego_goal_reached()) { *collision = true; break; }
    if (state->is_terminal()) break;
  }
  return state->ego_goal_reached();
}

TEST(dropin_hypothesis, device_hypothesis_mcts_true_hypothesis_equals_reference_search_every_step) {
  bool collision = false;
  int mismatches = 0;
  const bool goal = closed_loop_device_vs_stock({{5, 5}}, {5, 5}, 800, 100, &collision, &mismatches);
  std::printf("  device hypothesis planner, true hypothesis: goal %d collision %d mismatching steps %d\n", (int)goal, (int)collision, mismatches);
  EXPECT_TRUE(goal);
  EXPECT_FALSE(collision);
  EXPECT_EQ(mismatches, 0);
}

TEST(dropin_hypothesis, device_hypothesis_mcts_wrong_hypothesis_equals_reference_search_every_step) {
  bool collision = false;
  int mismatches = 0;
  const bool goal = closed_loop_device_vs_stock({{-2, 1}, {2, 3}, {4, 5}}, {-2, -2}, 800, 30, &collision, &mismatches);
  std::printf("  device hypothesis planner, wrong hypothesis: goal %d collision %d mismatching steps %d\n", (int)goal, (int)collision, mismatches);
  EXPECT_TRUE(goal);
  EXPECT_FALSE(collision);
  EXPECT_EQ(mismatches, 0);
}

// 4 096 decorrelated root-parallel hypothesis trees per decision
TEST(dropin_hypothesis, device_hypothesis_mcts_root_parallel_reaches_goal) {
  const auto params = default_crossing_state_parameters<Domain>();
  CudaState<CrossingState<Domain>>::set_parameters(params);
  const std::vector<std::pair<int, int>> hypotheses{{-2, 1}, {2, 3}, {4, 5}};
  CudaState<CrossingState<Domain>>::set_hypotheses(hypotheses);
  auto mcts_params = hypothesis_params(300);
  mcts_params.MAX_NUMBER_OF_NODES = 100000000;
  HypothesisBeliefTracker tracker(mcts_params);
  auto state = std::make_shared<CrossingState<Domain>>(tracker.sample_current_hypothesis(), params);
  for (const auto& h : hypotheses) state->add_hypothesis(AgentPolicyCrossingState<Domain>(h, params));
  tracker.belief_update(*state, *state);
  AgentPolicyCrossingState<Domain> true_agents_policy({-2, -2}, params);
  DeviceHypothesisMcts planner(mcts_params, 4096, true);
  std::vector<Reward> rewards;
  EgoCosts cost;
  bool collision = false;
  for (int i = 0; i < 30; ++i) {
    planner.search(*state, tracker);
    auto jointaction = JointAction(state->get_num_agents());
    jointaction[CrossingState<Domain>::ego_agent_idx] = planner.returnBestAction();
    for (auto agent_idx : state->get_other_agent_idx())
      jointaction[agent_idx] = aconv<Domain>(true_agents_policy.act(state->get_agent_state(agent_idx), state->get_ego_state()));
    auto next_state = state->execute(jointaction, rewards, cost);
    tracker.belief_update(*state, *next_state);
    state = next_state;
    if (state->is_terminal()) { collision = !state->ego_goal_reached(); break; }
  }
  EXPECT_TRUE(planner.get_merged_root_statistic().total_node_visits == 4096u * 300u);
  EXPECT_TRUE(state->ego_goal_reached());
  EXPECT_FALSE(collision);
}

int main(int argc, char** argv) {
  ::testing::InitGoogleTest(&argc, argv);
  return RUN_ALL_TESTS();
}
