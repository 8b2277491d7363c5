"""GPU parity tests for the batched rollout kernels, called through the C ABI
(mamcts_rollout_crossing_i32 / mamcts_rollout_simple) and checked against
  - the golden vectors generated from the unmodified reference (REPLAY and REFERENCE_CONST), bit-exact
  - the CPU oracle on fresh seeded inputs (all three action sources, K rollouts per leaf), bit-exact
  - the compiled reference itself when oracle/_ref travelled to the box
  - size-independent properties at full batch sizes."""
from __future__ import annotations

import numpy as np
import pytest

import oracle_api as O
from helpers import assert_bit_equal, golden, unhex
from mamcts_b200 import default_crossing_params, default_params, default_simple_params
from mamcts_b200._abi import SRC_HYPOTHESIS, SRC_PHILOX, SRC_REFERENCE_CONST, SRC_REPLAY
from mamcts_b200.engine import MamctsError

pytestmark = pytest.mark.gpu


def _const_actions(seed, nacts):
    order = golden("expand_order.json")[str(seed)]
    return [order[str(n)][0] for n in nacts]


def _check_against(out, exp, A, n, K, trace_steps=None):
    assert out["steps"].tolist() == exp["steps"].tolist()
    assert out["final_x"].tolist() == exp["final_x"].tolist()
    assert out["final_flags"].tolist() == exp["final_flags"].tolist()
    assert out["total_steps"] == int(np.sum(exp["steps"], dtype=np.uint64))
    assert_bit_equal(out["ego_return"], exp["ego_return"], "ego_return")
    assert_bit_equal(out["other_return"], exp["other_return"], "other_return")
    assert_bit_equal(out["ego_cost"], exp["ego_cost"], "ego_cost")
    assert_bit_equal(out["step_length"], exp["step_length"], "step_length")


@pytest.mark.parametrize("mode", ["replay", "const"])
def test_crossing_golden_bit_exact(engine, mode):
    """Replayed rollouts: same terminal state, step count, rewards and returns as the reference."""
    for case in golden("rollout_crossing.json"):
        mp = default_params(random_seed=case["seed"], rh_max_number_of_iterations=50)
        cp = default_crossing_params(num_other_agents=case["num_others"], reward_step=case["reward_step"])
        A = case["num_others"] + 1
        x = np.array(case["x"], dtype=np.int32)
        n = x.shape[1]
        steps = np.array(case["steps"], dtype=np.uint32)
        if mode == "replay":
            kw = dict(kind=SRC_REPLAY, replay_actions=np.array(case["rec_actions"], dtype=np.int8),
                      replay_steps=steps)
        else:
            kw = dict(kind=SRC_REFERENCE_CONST,
                      const_actions=_const_actions(case["seed"], [4] + [7] * (A - 1)))
        out = engine.rollout_crossing(mp, cp, x, None, np.array(case["flags"], dtype=np.uint8),
                                      np.array(case["depth"], dtype=np.uint32), K=1, parity=True,
                                      trace_stride=50, **kw)
        assert out["steps"].tolist() == case["steps"]
        assert out["final_x"].tolist() == case["final_x"]
        assert out["final_last_action"].tolist() == case["final_last_action"]
        assert out["final_flags"].tolist() == case["final_flags"]
        assert out["total_steps"] == sum(case["steps"])
        assert_bit_equal(out["ego_return"], unhex(case["ego_return"]), "ego_return")
        assert_bit_equal(out["other_return"], unhex(case["other_return"], (A - 1, n)), "other_return")
        assert_bit_equal(out["ego_cost"], unhex(case["ego_cost"]), "ego_cost")
        assert_bit_equal(out["step_length"], unhex(case["step_length"]), "step_length")
        rec = unhex(case["rec_rewards"], (n, 50))
        for i in range(n):
            assert_bit_equal(out["reward_trace"][i, :steps[i]], rec[i, :steps[i]], "reward trace")


def test_simple_golden_bit_exact(engine):
    sp = default_simple_params()
    for case in golden("rollout_simple.json"):
        mp = default_params(random_seed=case["seed"], rh_max_number_of_iterations=40)
        length = np.array(case["length"], dtype=np.int32)
        for kw in (dict(kind=SRC_REFERENCE_CONST, const_actions=_const_actions(case["seed"], [2, 2])),
                   dict(kind=SRC_REPLAY, replay_actions=np.array(case["rec_actions"], dtype=np.int8),
                        replay_steps=np.array(case["steps"], dtype=np.uint32))):
            out = engine.rollout_simple(mp, sp, length, K=1, parity=True, **kw)
            assert out["steps"].tolist() == case["steps"]
            assert out["final_x"][0].tolist() == case["final_len"]
            assert_bit_equal(out["ego_return"], unhex(case["ego_return"]), "ego")
            assert_bit_equal(out["other_return"][0], unhex(case["other_return"]), "other")
            assert_bit_equal(out["step_length"], unhex(case["step_length"]), "len")


@pytest.mark.parametrize("num_others", [1, 2, 3, 4, 5, 6, 7, 11, 15])
def test_crossing_philox_equals_oracle(engine, num_others):
    rng = np.random.default_rng(100 + num_others)
    A = num_others + 1
    mp = default_params(rh_max_number_of_iterations=50)
    cp = default_crossing_params(num_other_agents=num_others, reward_step=-0.25,
                                 cost_only_collision=num_others % 2)
    n = 193  # ragged: not a multiple of the warp size
    x = rng.integers(0, 12, size=(A, n)).astype(np.int32)
    depth = rng.integers(0, 5, size=n).astype(np.uint32)
    flags = (rng.random(n) < 0.05).astype(np.uint8)
    stream_hi = rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    kw = dict(kind=SRC_PHILOX, seed=0xDEADBEEF12345678, stream_hi=stream_hi, stream_lo_base=4000000000)
    out = engine.rollout_crossing(mp, cp, x, None, flags, depth, K=1, parity=True, **kw)
    exp = O.rollout_batch(O.make_env(cp=cp), mp, x, None, flags, depth, K=1, **kw)
    _check_against(out, exp, A, n, 1)
    assert out["final_last_action"].tolist() == exp["final_last_action"].tolist()


def test_hypothesis_golden_replay_bit_exact(engine):
    """C4 rollout path: rollouts recorded from the reference's
    RandomHeuristic::rollout<.., UctStatistic, HypothesisStatistic, ..> (other agents follow their
    behaviour hypothesis, incl. negative velocities) replayed on the GPU: same terminal state, step
    count, returns and cost, bit for bit."""
    for case in golden("hypothesis.json")["rollouts"]:
        mp = default_params(random_seed=case["seed"], rh_max_number_of_iterations=50)
        cp = default_crossing_params(num_other_agents=case["num_others"])
        x, last = np.array(case["x"], dtype=np.int32), np.array(case["last"], dtype=np.int32)
        A, n = x.shape
        out = engine.rollout_crossing(mp, cp, x, last, K=1, parity=True, kind=SRC_REPLAY,
                                      replay_actions=np.array(case["rec_actions"], dtype=np.int8),
                                      replay_steps=np.array(case["steps"], dtype=np.uint32))
        assert out["steps"].tolist() == case["steps"]
        assert out["final_x"].tolist() == case["final_x"]
        assert out["final_last_action"].tolist() == case["final_last_action"]
        assert out["final_flags"].tolist() == case["final_flags"]
        assert_bit_equal(out["ego_return"], unhex(case["ego_return"]), "ego return")
        assert_bit_equal(out["ego_cost"], unhex(case["ego_cost"]), "cost")
        assert_bit_equal(out["other_return"], unhex(case["other_return"], (A - 1, n)), "other return")
        assert_bit_equal(out["step_length"], unhex(case["step_length"]), "step length")


@pytest.mark.parametrize("num_others", [1, 2, 3, 7])
def test_hypothesis_source_equals_oracle(engine, num_others):
    """MAMCTS_SRC_HYPOTHESIS: other agents draw a gap from their hypothesis' range and act by the
    gap-keeping policy on the device; equal to the oracle driven by the same Philox streams."""
    rng = np.random.default_rng(500 + num_others)
    A = num_others + 1
    mp = default_params(rh_max_number_of_iterations=50)
    cp = default_crossing_params(num_other_agents=num_others, reward_step=-0.5)
    n = 171
    x = rng.integers(0, 12, size=(A, n)).astype(np.int32)
    last = rng.integers(-1, 3, size=(A, n)).astype(np.int32)
    depth = rng.integers(0, 5, size=n).astype(np.uint32)
    hyps = np.array([(4, 5), (5, 6), (-2, 1), (2, 3), (0, 0)])
    pick = rng.integers(0, len(hyps), size=(num_others, n))
    lo, hi = hyps[pick, 0].astype(np.int32), hyps[pick, 1].astype(np.int32)
    stream_hi = rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    for K in (1, 4):
        kw = dict(kind=SRC_HYPOTHESIS, seed=0xABCDEF0123, stream_hi=stream_hi, stream_lo_base=17,
                  hyp_gap_lo=lo, hyp_gap_hi=hi)
        out = engine.rollout_crossing(mp, cp, x, last, None, depth, K=K, parity=True, **kw)
        exp = O.rollout_batch(O.make_env(cp=cp), mp, x, last, None, depth, K=K, **kw)
        _check_against(out, exp, A, n, K)
        assert out["final_last_action"].tolist() == exp["final_last_action"].tolist()
    assert (out["final_last_action"][1:] < 0).any()      # braking happened


def test_hypothesis_source_rejects_bad_arguments(engine):
    mp = default_params(rh_max_number_of_iterations=10)
    with pytest.raises(MamctsError, match="HYPOTHESIS"):
        engine.rollout_simple(mp, default_simple_params(), np.zeros(4, dtype=np.int32), kind=SRC_HYPOTHESIS,
                              hyp_gap_lo=np.zeros((1, 4), dtype=np.int32), hyp_gap_hi=np.zeros((1, 4), dtype=np.int32))


@pytest.mark.parametrize("K", [2, 4, 8, 16, 32, 64, 96, 256, 512])
def test_leaf_parallel_mean_equals_oracle(engine, K):
    """K rollouts per leaf: the fixed-order warp-shuffle segmented reduction is deterministic."""
    rng = np.random.default_rng(K)
    mp = default_params(rh_max_number_of_iterations=50)
    cp = default_crossing_params(num_other_agents=2, reward_step=-1.0)
    n = 7 if K >= 256 else 21
    x = rng.integers(2, 11, size=(3, n)).astype(np.int32)
    kw = dict(kind=SRC_PHILOX, seed=77, stream_hi_base=5, stream_lo_base=123)
    out = engine.rollout_crossing(mp, cp, x, K=K, parity=True, **kw)
    exp = O.rollout_batch(O.make_env(cp=cp), mp, x, K=K, **kw)
    _check_against(out, exp, 3, n, K)
    out2 = engine.rollout_crossing(mp, cp, x, K=K, parity=False, **kw)
    assert_bit_equal(out2["ego_return"], out["ego_return"], "run-to-run determinism")


def test_simple_philox_equals_oracle(engine):
    mp = default_params(rh_max_number_of_iterations=60)
    sp = default_simple_params()
    length = np.arange(-1, 11, dtype=np.int32).repeat(5)
    kw = dict(kind=SRC_PHILOX, seed=5, stream_hi_base=9)
    out = engine.rollout_simple(mp, sp, length, K=4, parity=True, **kw)
    exp = O.rollout_batch(O.make_env(sp=sp), mp, length.reshape(1, -1), K=4, **kw)
    assert out["steps"].tolist() == exp["steps"].tolist()
    assert out["final_x"].tolist() == exp["final_x"][:1].tolist()
    assert_bit_equal(out["ego_return"], exp["ego_return"], "ego")
    assert_bit_equal(out["other_return"], exp["other_return"], "other")


@pytest.mark.skipif(not O.reference_available(), reason="compiled reference not on this box")
def test_replay_of_fresh_reference_rollouts(engine):
    """Record new reference rollouts here, replay them on the GPU: bit-exact."""
    rng = np.random.default_rng(31337)
    for num_others, seed in ((1, 1000), (2, 7), (7, 42), (3, 9)):
        mp = default_params(random_seed=seed, rh_max_number_of_iterations=50)
        cp = default_crossing_params(num_other_agents=num_others)
        A = num_others + 1
        n = 300
        x = rng.integers(0, 13, size=(A, n)).astype(np.int32)
        ref = O.ref_rollout_crossing(mp, cp, x, rec_stride=50)
        out = engine.rollout_crossing(mp, cp, x, K=1, kind=SRC_REPLAY, parity=True, trace_stride=50,
                                      replay_actions=ref["rec_actions"], replay_steps=ref["steps"])
        assert out["steps"].tolist() == ref["steps"].tolist()
        assert out["final_x"].tolist() == ref["final_x"].tolist()
        assert out["final_last_action"].tolist() == ref["final_last_action"].tolist()
        assert out["final_flags"].tolist() == ref["final_flags"].tolist()
        assert_bit_equal(out["ego_return"], ref["ego_return"], "ego")
        assert_bit_equal(out["ego_cost"], ref["ego_cost"], "cost")
        for i in range(n):
            s = ref["steps"][i]
            assert_bit_equal(out["reward_trace"][i, :s], ref["rec_rewards"][i, :s], "rewards")


def test_edge_cases(engine):
    mp = default_params(rh_max_number_of_iterations=50)
    cp = default_crossing_params(num_other_agents=1)
    # empty batch
    out = engine.rollout_crossing(mp, cp, np.zeros((2, 0), dtype=np.int32), kind=SRC_PHILOX, seed=1)
    assert out["ego_return"].shape == (0,) and out["total_steps"] == 0
    # every leaf terminal / at the depth cap / zero rollout budget: no step is executed
    x = np.full((2, 65), 5, dtype=np.int32)
    out = engine.rollout_crossing(mp, cp, x, flags=np.ones(65, dtype=np.uint8), kind=SRC_PHILOX, seed=1,
                                  parity=True)
    assert out["total_steps"] == 0 and not out["ego_return"].any() and not out["steps"].any()
    out = engine.rollout_crossing(mp, cp, x, depth=np.full(65, 1000, dtype=np.uint32), kind=SRC_PHILOX,
                                  seed=1, parity=True)
    assert out["total_steps"] == 0
    mp0 = default_params(rh_max_number_of_iterations=0)
    out = engine.rollout_crossing(mp0, cp, x, kind=SRC_PHILOX, seed=1, parity=True)
    assert out["total_steps"] == 0
    # invalid arguments are refused with a message, never executed
    with pytest.raises(MamctsError, match="rollouts_per_leaf"):
        engine.rollout_crossing(mp, cp, x, K=3, kind=SRC_PHILOX, seed=1)
    with pytest.raises(MamctsError, match="const action"):
        engine.rollout_crossing(mp, cp, x, kind=SRC_REFERENCE_CONST, const_actions=[9, 9])


def test_full_size_properties(engine):
    """BASELINE config 3 shape (A = 8, 1M rollouts, depth cap 50): properties that need no oracle."""
    mp = default_params(rh_max_number_of_iterations=50)
    cp = default_crossing_params(num_other_agents=7)
    n, K = 4096, 256  # 1 048 576 rollouts
    x = np.full((8, n), 5, dtype=np.int32)
    kw = dict(kind=SRC_PHILOX, seed=1000, stream_hi_base=0, stream_lo_base=0)
    out = engine.rollout_crossing(mp, cp, x, K=K, parity=True, **kw)
    steps = out["steps"].reshape(n, K)
    assert steps.min() >= 1 and steps.max() <= 50
    assert out["total_steps"] == int(steps.sum(dtype=np.uint64))
    np.testing.assert_allclose(out["step_length"], steps.mean(axis=1), rtol=1e-12)
    fl = out["final_flags"]
    fx = out["final_x"]
    term = (fl & 1).astype(bool)
    assert np.all(term | (out["steps"] == 50))            # only the cap ends a non-terminal rollout
    goal, coll = (fl & 2).astype(bool), (fl & 4).astype(bool)
    assert not np.any(goal & coll)
    assert np.all(fx[0][goal] >= cp.ego_goal_pos)
    assert np.all(fx[1:].min(axis=0) >= 0)                 # others are clamped at 0
    oob = term & ~goal & ~coll
    assert np.all(fx[0][oob] < 0)
    assert np.all(out["other_return"] == 0.0)              # CrossingState gives others no reward
    # sample of leaves against the oracle (same streams), bit-exact
    exp = O.rollout_batch(O.make_env(cp=cp), mp, x[:, :2], K=K, **kw)
    assert_bit_equal(out["ego_return"][:2], exp["ego_return"], "ego sample")
    assert out["steps"][:2 * K].tolist() == exp["steps"].tolist()
    # idempotence: same seeds -> same bits
    out2 = engine.rollout_crossing(mp, cp, x, K=K, **kw)
    assert_bit_equal(out2["ego_return"], out["ego_return"], "idempotence")
    assert out2["total_steps"] == out["total_steps"]


def test_moments_inside_the_rollout_call(engine):
    """out->moments (one decision, many rollouts: C3): the same {sum, sum^2, count} whether the rollouts are
    launched as n leaves x K or as n*K leaves x 1 with the same global stream ids, and equal to the separate
    mamcts_rollout_moments call over the per-rollout returns - bit for bit (same values, same fixed order)."""
    from helpers import assert_bit_equal
    mp = default_params(rh_max_number_of_iterations=50)
    cp = default_crossing_params(num_other_agents=3)
    n, K = 40, 32
    a = engine.rollout_crossing(mp, cp, np.full((4, n), 5, dtype=np.int32), K=K, kind=SRC_PHILOX, seed=9,
                                stream_hi_base=3, stream_lo_base=1000, moments=True)
    b = engine.rollout_crossing(mp, cp, np.full((4, n * K), 5, dtype=np.int32), K=1, kind=SRC_PHILOX, seed=9,
                                stream_hi_base=3, stream_lo_base=1000, moments=True)
    assert a["total_steps"] == b["total_steps"] and a["moments"][-1] == n * K
    assert_bit_equal(a["moments"], b["moments"], "moments K=32 vs K=1")
    assert_bit_equal(b["moments"], engine.rollout_moments(b["ego_return"], b["other_return"]), "in-call vs separate call")
    assert np.allclose(a["moments"][0] / (n * K), b["ego_return"].mean(), rtol=1e-12, atol=1e-15)


def test_action_values_match_reference_golden_and_oracle(engine):
    """mamcts_action_values_crossing_i32 = the compute part of ActionValueRandomHeuristic
    (action_value_random_heuristic.h:37-85): REFERENCE_CONST against the committed vectors of the reference's own
    loop (bit for bit), PHILOX against the oracle on the same streams (bit for bit), 1 to 7 other agents."""
    from helpers import assert_bit_equal
    for case in golden("action_values.json"):
        cp = default_crossing_params(num_other_agents=case["num_others"], **case["params"])
        mp = default_params(random_seed=case["seed"], rh_max_number_of_iterations=50, max_search_depth=60)
        x = np.array(case["x"], dtype=np.int32)
        n, nA, J = x.shape[1], cp.num_ego_actions, case["num_others"]
        last, flags, depth = np.array(case["last"]), np.array(case["flags"], dtype=np.uint8), np.array(case["depth"])
        got = engine.action_values_crossing(mp, cp, x, last, flags, depth, kind=SRC_REFERENCE_CONST,
                                            const_actions=case["const_actions"])
        assert_bit_equal(got["action_return"], unhex(case["action_return"], (n, nA)), "action returns")
        assert_bit_equal(got["action_cost"], unhex(case["action_cost"], (n, nA)), "action costs")
        assert_bit_equal(got["other_return"], unhex(case["other_return"], (J, n)), "other returns")
        assert_bit_equal(got["mean_cost"], unhex(case["mean_cost"]), "mean cost")
        # uniform Philox actions: the oracle twin on the same streams, per-leaf stream ids
        rng = np.random.default_rng(case["seed"])
        hi = rng.integers(0, 2**31, size=n).astype(np.uint32)
        lo = rng.integers(0, 2**20, size=n).astype(np.uint32)
        a = engine.action_values_crossing(mp, cp, x, last, flags, depth, kind=SRC_PHILOX, seed=77, stream_hi=hi, stream_lo=lo)
        b = O.action_values_batch(O.make_env(cp=cp), mp, x, last, flags, depth, kind=SRC_PHILOX, seed=77, stream_hi=hi, stream_lo=lo)
        for k in ("action_return", "action_cost", "action_step_length", "other_return", "mean_cost"):
            assert_bit_equal(a[k], b[k], f"philox {k}")
        assert a["total_steps"] == b["total_steps"] > 0
        assert np.all(a["action_return"][flags == 1] == 0.0) and np.all(a["action_step_length"][flags == 1] == 0.0)
    # a larger batch: 20 000 leaves x 4 actions in one call, stream ids from the bases
    cp = default_crossing_params(num_other_agents=2)
    mp = default_params(rh_max_number_of_iterations=50)
    n = 20000
    x = np.random.default_rng(1).integers(0, 12, size=(3, n)).astype(np.int32)
    a = engine.action_values_crossing(mp, cp, x, kind=SRC_PHILOX, seed=5, stream_hi_base=9, stream_lo_base=100)
    sel = np.arange(0, n, 997)
    b = O.action_values_batch(O.make_env(cp=cp), mp, x[:, sel], kind=SRC_PHILOX, seed=5,
                              stream_hi=np.full(sel.size, 9), stream_lo=100 + sel)
    assert_bit_equal(a["action_return"][sel], b["action_return"], "large batch action returns")
    assert_bit_equal(a["mean_cost"][sel], b["mean_cost"], "large batch mean cost")
