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
