"""Python-side constructors of the parameter structs, mirroring the reference defaults.

mcts_default_parameters()                 : reference mcts/mcts_parameters.h:97-135
default_crossing_state_parameters<int>()  : reference environments/crossing_state_parameters.h:36-59
SimpleState constants                     : reference test/uct/simple_state.h:16
"""
from __future__ import annotations

from ._abi import CrossingParams, CrossingParamsF32, HypothesisParams, Params, SimpleParams


def default_params(**over) -> Params:
    p = Params()
    p.discount_factor = 0.9
    p.random_seed = 1000
    p.max_number_of_iterations = 10000
    p.max_number_of_nodes = 10000
    p.max_search_depth = 1000
    p.use_bound_estimation = 1
    p.num_parallel_mcts = 1
    p.rh_max_number_of_iterations = 1000
    p.uct_lower_bound = -1000.0
    p.uct_upper_bound = 100.0
    p.uct_exploration_constant = 0.7
    p.uct_pw_k = 4.0
    p.uct_pw_alpha = 0.25
    for k, v in over.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def default_hypothesis_params(**over) -> HypothesisParams:
    """MctsParameters::hypothesis_statistic defaults (mcts_parameters.h:118-124) and the policy seed
    (default_crossing_state_parameters: OTHER_AGENTS_POLICY_RANDOM_SEED = 1000)."""
    h = HypothesisParams()
    h.cost_based_action_selection = 0
    h.lower_cost_bound = 0.0
    h.upper_cost_bound = 100.0
    h.progressive_widening_hypothesis_based = 1
    h.pw_alpha = 0.5
    h.pw_k = 1.0
    h.exploration_constant = 0.7
    h.chance_constrained_update = 0
    h.policy_random_seed = 1000
    for k, v in over.items():
        if not hasattr(h, k):
            raise AttributeError(k)
        setattr(h, k, v)
    return h


def default_crossing_params(**over) -> CrossingParams:
    c = CrossingParams()
    c.num_other_agents = 2
    c.cost_only_collision = 0
    c.max_velocity_other = 3
    c.min_velocity_other = -3
    c.max_velocity_ego = 2
    c.min_velocity_ego = -1
    c.chain_length = 21
    c.ego_goal_pos = 12
    c.num_other_actions = 7
    c.reward_collision = -1000.0
    c.reward_goal_reached = 100.0
    c.reward_step = 0.0
    for k, v in over.items():
        if not hasattr(c, k):
            raise AttributeError(k)
        setattr(c, k, v)
    return c


def default_crossing_params_f32(**over) -> CrossingParamsF32:
    """default_crossing_state_parameters<float>(): the same values, 30 other actions (:50-51)."""
    c = CrossingParamsF32()
    c.num_other_agents = 2
    c.cost_only_collision = 0
    c.max_velocity_other = 3.0
    c.min_velocity_other = -3.0
    c.max_velocity_ego = 2.0
    c.min_velocity_ego = -1.0
    c.chain_length = 21.0
    c.ego_goal_pos = 12.0
    c.num_other_actions = 30
    c.reward_collision = -1000.0
    c.reward_goal_reached = 100.0
    c.reward_step = 0.0
    for k, v in over.items():
        if not hasattr(c, k):
            raise AttributeError(k)
        setattr(c, k, v)
    return c


def default_simple_params() -> SimpleParams:
    s = SimpleParams()
    s.winning_state_length = 10
    s.loosing_state_length = -1
    return s
