"""CrossingState<float> rollouts on the GPU (mamcts_rollout_crossing_f32): recorded
reference rollouts replay bit-exactly (terminal state, step count, rewards -> returns, cost), the
reference's constant-action rollouts are reproduced from the action indices alone, and Philox rollouts
equal the oracle driven by the same streams."""
from __future__ import annotations

import numpy as np
import pytest

import oracle_api as O
from helpers import assert_bit_equal, golden, unhex
from mamcts_b200 import default_crossing_params_f32, default_params
from mamcts_b200._abi import SRC_HYPOTHESIS, SRC_PHILOX, SRC_REFERENCE_CONST, SRC_REPLAY
from mamcts_b200.engine import MamctsError
from test_float_domain_cpu import f32, same_bits

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["replay", "const"])
def test_float_golden_bit_exact(engine, mode):
    for case in golden("float_domain.json")["rollouts"]:
        mp = default_params(random_seed=case["seed"], rh_max_number_of_iterations=50)
        cp = default_crossing_params_f32(num_other_agents=case["num_others"])
        A = case["num_others"] + 1
        x = f32(case["x"], (A, -1))
        n = x.shape[1]
        if mode == "replay":
            kw = dict(kind=SRC_REPLAY, replay_actions=f32(case["rec_actions"], (n, 50, A)),
                      replay_steps=np.array(case["steps"], dtype=np.uint32))
        else:
            kw = dict(kind=SRC_REFERENCE_CONST, const_actions=case["const_actions"])
        out = engine.rollout_crossing_f32(mp, cp, x, None, None, np.array(case["depth"], dtype=np.uint32), K=1,
                                          parity=True, **kw)
        assert out["steps"].tolist() == case["steps"]
        assert same_bits(out["final_x"], f32(case["final_x"], (A, n)))
        assert same_bits(out["final_last_action"], f32(case["final_last_action"], (A, n)))
        assert out["final_flags"].tolist() == case["final_flags"]
        assert_bit_equal(out["ego_return"], unhex(case["ego_return"]), "ego return")
        assert_bit_equal(out["ego_cost"], unhex(case["ego_cost"]), "cost")
        assert_bit_equal(out["step_length"], unhex(case["step_length"]), "step length")
        assert not out["other_return"].any()


@pytest.mark.parametrize("num_others", [1, 2, 3, 7])
def test_float_philox_and_arbitrary_replay_equal_oracle(engine, num_others):
    rng = np.random.default_rng(60 + num_others)
    A = num_others + 1
    mp = default_params(rh_max_number_of_iterations=50)
    cp = default_crossing_params_f32(num_other_agents=num_others, chain_length=20.0, ego_goal_pos=11.5,
                                     reward_step=-0.25, cost_only_collision=num_others % 2)
    n = 150
    x = (rng.random((A, n)) * 12).astype(np.float32)
    last = (rng.random((A, n)) * 4 - 2).astype(np.float32)
    depth = rng.integers(0, 4, size=n).astype(np.uint32)
    # Philox, index-valued (the others' indices act as denormal velocities)
    out = engine.rollout_crossing_f32(mp, cp, x, last, None, depth, K=1, parity=True, kind=SRC_PHILOX, seed=99,
                                      stream_hi_base=3, stream_lo_base=1000)
    for i in range(n):
        o = O.rollout_f32(cp, mp, x[:, i], last[:, i], depth=int(depth[i]), kind=SRC_PHILOX, seed=99, stream_hi=3,
                          stream_lo=1000 + i)
        assert out["steps"][i] == o["steps"] and out["final_flags"][i] == o["final_flags"]
        assert_bit_equal(out["ego_return"][i], o["ego_return"], "ego")
        assert_bit_equal(out["ego_cost"][i], o["cost0"], "cost")
        assert same_bits(out["final_x"][:, i], o["final_x"])
    # arbitrary float velocities through REPLAY (what a float-domain hypothesis policy would produce)
    steps = rng.integers(0, 30, size=n).astype(np.uint32)
    rec = np.zeros((n, 30, A), dtype=np.float32)
    rec[:, :, 0] = rng.integers(0, 4, size=(n, 30))
    rec[:, :, 1:] = (rng.random((n, 30, A - 1)) * 5 - 1.5).astype(np.float32)
    out = engine.rollout_crossing_f32(mp, cp, x, last, None, depth, K=1, parity=True, kind=SRC_REPLAY,
                                      replay_actions=rec, replay_steps=steps)
    for i in range(n):
        o = O.rollout_f32(cp, mp, x[:, i], last[:, i], depth=int(depth[i]), kind=SRC_REPLAY, replay=rec[i, :steps[i]])
        assert out["steps"][i] == o["steps"] and out["final_flags"][i] == o["final_flags"]
        assert_bit_equal(out["ego_return"][i], o["ego_return"], "ego")
        assert same_bits(out["final_x"][:, i], o["final_x"]) and same_bits(out["final_last_action"][:, i], o["final_last"])


def test_float_leaf_parallel_and_errors(engine):
    mp = default_params(rh_max_number_of_iterations=50)
    cp = default_crossing_params_f32(num_other_agents=2)
    x = np.full((3, 64), 5.0, dtype=np.float32)
    a = engine.rollout_crossing_f32(mp, cp, x, K=32, kind=SRC_PHILOX, seed=5)
    b = engine.rollout_crossing_f32(mp, cp, x, K=32, kind=SRC_PHILOX, seed=5)
    assert_bit_equal(a["ego_return"], b["ego_return"], "determinism")
    assert a["total_steps"] == b["total_steps"] > 0
    with pytest.raises(MamctsError, match="HYPOTHESIS"):   # the gap ranges are required
        engine.rollout_crossing_f32(mp, cp, x, kind=SRC_HYPOTHESIS)
    with pytest.raises(MamctsError, match="compiled for"):
        engine.rollout_crossing_f32(mp, default_crossing_params_f32(num_other_agents=5), np.zeros((6, 4), dtype=np.float32))


def test_float_hypothesis_policy_on_device(engine):
    """AgentPolicyCrossingState<float> on the device (MAMCTS_SRC_HYPOTHESIS of mamcts_rollout_crossing_f32): (1) the
    reference's own float-domain hypothesis rollouts - its uniform_real_distribution gap draws recorded as
    velocities - replay bit-exactly; (2) the device policy (Philox gap, calculate_action in FP32,
    crossing_state_agent_policy.h:35-55,78-84) equals the oracle twin step for step on the same streams."""
    for case in golden("float_hypothesis.json")["rollouts"]:
        mp = default_params(random_seed=case["seed"], rh_max_number_of_iterations=50)
        cp = default_crossing_params_f32(num_other_agents=case["num_others"])
        A = case["num_others"] + 1
        x, last = f32(case["x"], (A, -1)), f32(case["last"], (A, -1))
        n = x.shape[1]
        depth = np.array(case["depth"], dtype=np.uint32)
        out = engine.rollout_crossing_f32(mp, cp, x, last, None, depth, K=1, parity=True, kind=SRC_REPLAY,
                                          replay_actions=f32(case["rec_actions"], (n, 50, A)),
                                          replay_steps=np.array(case["steps"], dtype=np.uint32))
        assert out["steps"].tolist() == case["steps"]
        assert same_bits(out["final_x"], f32(case["final_x"], (A, n)))
        assert same_bits(out["final_last_action"], f32(case["final_last_action"], (A, n)))
        assert out["final_flags"].tolist() == case["final_flags"]
        assert_bit_equal(out["ego_return"], unhex(case["ego_return"]), "ego return")
        assert_bit_equal(out["ego_cost"], unhex(case["ego_cost"]), "cost")
        # the device policy with Philox gaps against the oracle twin
        hyps = case["hyps"]
        cur = np.array(case["cur_hyp"])
        lo = np.array([[hyps[h][0] for h in row] for row in cur], dtype=np.float32)
        hi = np.array([[hyps[h][1] for h in row] for row in cur], dtype=np.float32)
        got = engine.rollout_crossing_f32(mp, cp, x, last, None, depth, K=1, parity=True, kind=SRC_HYPOTHESIS, seed=321,
                                          stream_hi_base=5, stream_lo_base=70, hyp_gap_lo=lo, hyp_gap_hi=hi)
        for i in range(n):
            o = O.rollout_f32_hypothesis(cp, mp, x[:, i], last[:, i], depth=int(depth[i]), seed=321, stream_hi=5,
                                         stream_lo=70 + i, gap_lo=lo[:, i], gap_hi=hi[:, i])
            assert got["steps"][i] == o["steps"] and got["final_flags"][i] == o["final_flags"]
            assert_bit_equal(got["ego_return"][i], o["ego_return"], "ego")
            assert_bit_equal(got["ego_cost"][i], o["cost0"], "cost")
            assert same_bits(got["final_x"][:, i], o["final_x"]) and same_bits(got["final_last_action"][:, i], o["final_last"])
    # K rollouts per leaf, 7 other agents
    cp7 = default_crossing_params_f32(num_other_agents=7)
    x = np.full((8, 256), 5.0, dtype=np.float32)
    lo, hi = np.full((7, 256), 2.0, dtype=np.float32), np.full((7, 256), 4.5, dtype=np.float32)
    a = engine.rollout_crossing_f32(default_params(rh_max_number_of_iterations=50), cp7, x, K=32, kind=SRC_HYPOTHESIS, seed=1,
                                    hyp_gap_lo=lo, hyp_gap_hi=hi)
    assert a["total_steps"] > 256 * 32
