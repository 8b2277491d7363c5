"""CrossingState<float>, CPU side: the oracle's FP32 restatement of execute and of
rollouts over it against golden vectors generated from the unmodified reference
(tests/golden/float_domain.json; float32 values stored as bit patterns)."""
from __future__ import annotations

import numpy as np

import oracle_api as O
from helpers import assert_bit_equal, golden, unhex
from mamcts_b200 import default_crossing_params_f32, default_params
from mamcts_b200._abi import SRC_REFERENCE_CONST, SRC_REPLAY


def f32(bits, shape=None):
    a = np.array(bits, dtype=np.uint32).view(np.float32)
    return a if shape is None else a.reshape(shape)


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a, dtype=np.float32).view(np.uint32),
                          np.ascontiguousarray(b, dtype=np.float32).view(np.uint32))


def test_execute_matches_golden():
    seen = set()
    for case in golden("float_domain.json")["execute"]:
        cp = default_crossing_params_f32(num_other_agents=case["num_others"], chain_length=case["chain_length"],
                                         ego_goal_pos=case["ego_goal_pos"], reward_step=case["reward_step"],
                                         cost_only_collision=case["cost_only_collision"])
        A = case["num_others"] + 1
        x, last, act = (f32(case[k], (A, -1)) for k in ("x", "last", "action"))
        nx, nl = f32(case["next_x"], (A, -1)), f32(case["next_last"], (A, -1))
        r, c0 = unhex(case["reward"]), unhex(case["cost0"])
        for i in range(x.shape[1]):
            ox, ol, of, orr, oc = O.crossing_execute_f32(cp, x[:, i], last[:, i], act[:, i])
            assert same_bits(ox, nx[:, i]) and same_bits(ol, nl[:, i]) and of == case["flags"][i]
            assert_bit_equal(orr, r[i], "reward")
            assert_bit_equal(oc, c0[i], "cost")
        seen |= set(case["flags"])
    assert {0, 1, 3, 5} <= seen   # running, out of map, goal and collision all occur


def test_rollouts_match_golden():
    for case in golden("float_domain.json")["rollouts"]:
        mp = default_params(random_seed=case["seed"], rh_max_number_of_iterations=50)
        cp = default_crossing_params_f32(num_other_agents=case["num_others"])
        A = case["num_others"] + 1
        x = f32(case["x"], (A, -1))
        n = x.shape[1]
        rec = f32(case["rec_actions"], (n, 50, A))
        fx, fl = f32(case["final_x"], (A, n)), f32(case["final_last_action"], (A, n))
        ego, cost = unhex(case["ego_return"]), unhex(case["ego_cost"])
        for i in range(n):
            for kw in (dict(kind=SRC_REPLAY, replay=rec[i, :case["steps"][i]]),
                       dict(kind=SRC_REFERENCE_CONST, const_actions=case["const_actions"])):
                o = O.rollout_f32(cp, mp, x[:, i], depth=case["depth"][i], **kw)
                assert o["steps"] == case["steps"][i]
                assert_bit_equal(o["ego_return"], ego[i], "ego return")
                assert_bit_equal(o["cost0"], cost[i], "cost")
                assert same_bits(o["final_x"], fx[:, i]) and same_bits(o["final_last"], fl[:, i])
                assert o["final_flags"] == case["final_flags"][i]


def test_float_policy_matches_golden():
    """AgentPolicyCrossingState<float>::calculate_action (crossing_state_agent_policy.h:35-55): the oracle's FP32
    restatement against the reference's results on 400 inputs (zero gaps, signed zeros included), bit for bit."""
    g = golden("float_hypothesis.json")["policy"]
    cp = default_crossing_params_f32(num_other_agents=2)
    ax, al, ex, el, gap, act = (f32(g[k]) for k in ("agent_x", "agent_last", "ego_x", "ego_last", "gap", "action"))
    got = np.array([O.crossing_policy_action_f32(cp, ax[i], al[i], ex[i], el[i], gap[i]) for i in range(ax.shape[0])],
                   dtype=np.float32)
    assert same_bits(got, act)


def test_float_hypothesis_rollouts_replay_matches_golden():
    """The reference's float-domain hypothesis rollouts (its own uniform_real_distribution gap draws), replayed by
    the oracle from the recorded velocities: steps, returns, costs and final states bit for bit."""
    for case in golden("float_hypothesis.json")["rollouts"]:
        mp = default_params(random_seed=case["seed"], rh_max_number_of_iterations=50)
        cp = default_crossing_params_f32(num_other_agents=case["num_others"])
        A = case["num_others"] + 1
        x, last = f32(case["x"], (A, -1)), f32(case["last"], (A, -1))
        n = x.shape[1]
        rec = f32(case["rec_actions"], (n, 50, A))
        fx, fl = f32(case["final_x"], (A, n)), f32(case["final_last_action"], (A, n))
        ego, cost = unhex(case["ego_return"]), unhex(case["ego_cost"])
        for i in range(n):
            o = O.rollout_f32(cp, mp, x[:, i], last[:, i], depth=case["depth"][i], kind=SRC_REPLAY,
                              replay=rec[i, :case["steps"][i]])
            assert o["steps"] == case["steps"][i]
            assert_bit_equal(o["ego_return"], ego[i], "ego return")
            assert_bit_equal(o["cost0"], cost[i], "cost")
            assert same_bits(o["final_x"], fx[:, i]) and same_bits(o["final_last"], fl[:, i])
            assert o["final_flags"] == case["final_flags"][i]
