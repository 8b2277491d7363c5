"""C4 (hypothesis-MCTS on CrossingState), rollout path, CPU side: the oracle's restatement of the
behaviour-hypothesis policy and of rollouts whose other agents follow it, against the golden vectors
generated from the unmodified reference (tests/golden/hypothesis.json) and, where oracle/_ref exists,
against the reference itself on fresh random inputs."""
from __future__ import annotations

import numpy as np
import pytest

import oracle_api as O
from helpers import assert_bit_equal, golden, unhex
from mamcts_b200 import default_crossing_params, default_params
from mamcts_b200._abi import SRC_HYPOTHESIS, SRC_REPLAY


def test_policy_matches_golden():
    """AgentPolicyCrossingState<int>::calculate_action (crossing_state_agent_policy.h:35-55)."""
    for case in golden("hypothesis.json")["policy"]:
        cp = default_crossing_params(chain_length=case["chain_length"],
                                     min_velocity_other=case["min_velocity_other"],
                                     max_velocity_other=case["max_velocity_other"])
        got = [O.crossing_policy_action(cp, *v) for v in zip(case["agent_x"], case["agent_last"], case["ego_x"],
                                                             case["ego_last"], case["gap"])]
        assert got == case["action"]


def test_reference_hypothesis_friendly_known_answer():
    """reference environments/tests/crossing_state_int_test.cc:60-95 (hypothesis_friendly): the ego moves
    +1 per step (action index 2), the other agents follow the deterministic hypothesis [5, 5]; the
    reference asserts min_distance_to_ego() == 5 and no collision when the ego reaches its goal."""
    cp = default_crossing_params()
    x, last = [5, 5, 5], [2, 2, 2]
    flags = 0
    for _ in range(100):
        act = [2] + [O.crossing_policy_action(cp, x[a], last[a], x[0], last[0], 5) for a in (1, 2)]
        x, last, flags, _, _ = O.crossing_execute(cp, x, last, act)
        if flags & 1:
            break
    assert flags & 2 and not flags & 4          # goal reached, no collision
    assert min(x[0] - x[1], x[0] - x[2]) == 5   # the desired gap


def test_replay_of_reference_hypothesis_rollouts_matches_golden():
    for case in golden("hypothesis.json")["rollouts"]:
        mp = default_params(random_seed=case["seed"], rh_max_number_of_iterations=50)
        cp = default_crossing_params(num_other_agents=case["num_others"])
        env = O.make_env(cp=cp)
        x, last = np.array(case["x"]), np.array(case["last"])
        A, n = x.shape
        rec = np.array(case["rec_actions"], dtype=np.int8)
        ego, cost = unhex(case["ego_return"]), unhex(case["ego_cost"])
        oth = unhex(case["other_return"], (A - 1, n))
        negative = False
        for i in range(n):
            steps = case["steps"][i]
            r = O.rollout(env, mp, x[:, i], last[:, i], kind=SRC_REPLAY, replay=rec[i, :steps])
            assert r["steps"] == steps
            assert_bit_equal(r["ego_return"], ego[i], "ego return")
            assert_bit_equal(r["cost0"], cost[i], "cost")
            assert_bit_equal(r["other_return"], oth[:, i], "other return")
            assert r["final_x"] == np.array(case["final_x"])[:, i].tolist()
            assert r["final_last"] == np.array(case["final_last_action"])[:, i].tolist()
            assert r["final_flags"] == case["final_flags"][i]
            # every recorded other-agent action is what the policy yields for SOME gap of its hypothesis
            st_x, st_last = x[:, i].copy(), last[:, i].copy()
            for s in range(steps):
                for a in range(1, A):
                    h = case["cur_hyp"][a - 1][i]
                    allowed = {O.crossing_policy_action(cp, int(st_x[a]), int(st_last[a]), int(st_x[0]),
                                                        int(st_last[0]), g)
                               for g in range(case["hyp_lo"][h], case["hyp_hi"][h] + 1)}
                    assert int(rec[i, s, a]) in allowed
                    negative |= rec[i, s, a] < 0
                nx, nl, _, _, _ = O.crossing_execute(cp, st_x, st_last, rec[i, s])
                st_x, st_last = np.array(nx), np.array(nl)
        assert negative or case["num_others"] == 1   # braking (negative Domain values) is exercised


def test_hypothesis_source_gaps_are_uniform_and_reproducible():
    """The oracle's HYPOTHESIS source: deterministic per stream; over many streams the first-step action
    of an agent is distributed as the policy applied to a uniform gap."""
    mp = default_params(rh_max_number_of_iterations=50)
    cp = default_crossing_params(num_other_agents=1)
    env = O.make_env(cp=cp)
    lo, hi = [-2], [3]
    a = O.rollout(env, mp, [5, 4], [2, 1], kind=SRC_HYPOTHESIS, seed=9, stream_hi=1, stream_lo=2, gap_lo=lo, gap_hi=hi)
    b = O.rollout(env, mp, [5, 4], [2, 1], kind=SRC_HYPOTHESIS, seed=9, stream_hi=1, stream_lo=2, gap_lo=lo, gap_hi=hi)
    assert a["steps"] == b["steps"] and a["ego_return"] == b["ego_return"] and a["final_x"] == b["final_x"]
    counts = {}
    mp1 = default_params(rh_max_number_of_iterations=1)
    for s in range(3000):
        r = O.rollout(env, mp1, [5, 4], [2, 1], kind=SRC_HYPOTHESIS, seed=9, stream_hi=0, stream_lo=s, gap_lo=lo, gap_hi=hi)
        counts[r["final_last"][1]] = counts.get(r["final_last"][1], 0) + 1
    expect = {}
    for g in range(lo[0], hi[0] + 1):
        v = O.crossing_policy_action(cp, 4, 1, 5, 2, g)
        expect[v] = expect.get(v, 0) + 3000 / 6
    assert set(counts) == set(expect)
    for v in expect:
        assert abs(counts[v] - expect[v]) < 5 * np.sqrt(expect[v])


@pytest.mark.skipif(not O.reference_available(), reason="compiled reference not on this box")
def test_policy_and_rollouts_equal_reference_on_fresh_inputs():
    rng = np.random.default_rng(31337)
    cp = default_crossing_params(num_other_agents=3, chain_length=41, ego_goal_pos=26)
    n = 3000
    ax = rng.integers(-2, 41, n); al = rng.integers(-3, 4, n)
    ex = rng.integers(-2, 41, n); el = rng.integers(-1, 3, n); gap = rng.integers(-8, 10, n)
    ref = O.ref_policy_calculate_action(cp, ax, al, ex, el, gap)
    mine = [O.crossing_policy_action(cp, int(ax[i]), int(al[i]), int(ex[i]), int(el[i]), int(gap[i])) for i in range(n)]
    assert ref.tolist() == mine
    mp = default_params(rh_max_number_of_iterations=50)
    A, m = 4, 40
    x = rng.integers(0, 20, size=(A, m)).astype(np.int32)
    last = rng.integers(-1, 3, size=(A, m)).astype(np.int32)
    cur = rng.integers(0, 2, size=(A - 1, m)).astype(np.uint32)
    r = O.ref_rollout_crossing_hypothesis(mp, cp, x, last, [4, -1], [6, 2], cur, rec_stride=50)
    env = O.make_env(cp=cp)
    for i in range(m):
        o = O.rollout(env, mp, x[:, i], last[:, i], kind=SRC_REPLAY, replay=r["rec_actions"][i, :r["steps"][i]])
        assert o["steps"] == r["steps"][i] and o["final_x"] == r["final_x"][:, i].tolist()
        assert_bit_equal(o["ego_return"], r["ego_return"][i], "ego")
        assert_bit_equal(o["cost0"], r["ego_cost"][i], "cost")


# ---- HypothesisStatistic back-propagation ------------------------------------------------------------

def _f32(v):
    return float(np.float32(v))


def _run_batches(run, case, stats):
    for b in case["batches"]:
        if not b["leaf_node"]:
            continue
        run(stats, b["leaf_node"], b["leaf_has"], np.array(b["leaf_cost"]).reshape(-1, case["C"]),
            b["path_offset"], b["path_len"], b["edge_node"], np.array(b["edge_hyp"]).reshape(-1, case["J"]),
            np.array(b["edge_action"]).reshape(-1, case["J"]), np.array(b["edge_cost"]).reshape(-1, case["C"]))


def check_hyp_stats_against_golden(stats, exp):
    S, Hp1 = stats["visits_hypothesis"].shape
    assert stats["total_visits"].tolist() == exp["total_visits"]
    assert stats["num_expanded"].tolist() == exp["num_expanded"]
    assert stats["visits_hypothesis"].tolist() == exp["visits_hypothesis"]
    assert stats["action_count"].tolist() == exp["action_count"]
    assert_bit_equal(stats["ego_cost_value"], unhex(exp["ego_cost_value"]), "ego_cost_value_")
    assert_bit_equal(stats["latest_ego_cost"], unhex(exp["latest_ego_cost"]), "latest_ego_cost_")
    assert_bit_equal(stats["action_cost"], unhex(exp["action_cost"], stats["action_cost"].shape), "action_ego_cost_")


def test_hypothesis_backup_reference_known_answers():
    """reference test/hypothesis/hypothesis_statistic_test.cc:23-118 (backprop_hypothesis_action_selection):
    the four updates of agent 1's statistic with their stated expected values (EXPECT_NEAR 1e-3)."""
    from mamcts_b200.engine import Engine
    mp = default_params()
    st = Engine.new_hyp_stats(5, 2, 6)   # node 0 = the parent statistic, nodes 1..4 = the four children

    def upd(child, cost, leaf_cost, hyp, act):
        O.backup_hypothesis(mp, 1, st, [child], [1], [[_f32(leaf_cost[0]), _f32(leaf_cost[1])]], [0], [1], [0],
                            [[hyp]], [[act]], [[_f32(cost[0]), _f32(cost[1])]])

    upd(1, (2.3, 5.0), (20.0, 1.2323), 0, 5)
    assert abs(st["action_cost"][0, 0, 5] - ((5.0 + 2.3) / 2.0 + 0.9 * (1.2323 + 20.0) / 2.0)) < 1e-3
    assert st["action_count"][0, 0, 5] == 1 and st["visits_hypothesis"][0, 0] == 1
    upd(2, (4.3, 4.0), (24.5, 23.0), 0, 5)
    assert abs(st["action_cost"][0, 0, 5] - ((5.0 + 4.0 + 4.3 + 2.3) / 2.0 + 0.9 * (1.2323 + 20.0) / 2.0
                                            + 0.9 * (23.0 + 24.5) / 2.0) / 2.0) < 1e-3
    assert st["action_count"][0, 0, 5] == 2 and st["visits_hypothesis"][0, 0] == 2
    upd(3, (1000.3, 1000.0), (450.5, 450.0), 0, 3)
    assert abs(st["action_cost"][0, 0, 3] - ((1000.0 + 1000.3) / 2.0 + 0.9 * (450.0 + 450.5) / 2.0)) < 1e-3
    assert st["action_count"][0, 0, 3] == 1 and st["visits_hypothesis"][0, 0] == 3
    upd(4, (10.3, 10.0), (45.5, 45.0), 1, 4)
    assert abs(st["action_cost"][0, 1, 4] - ((10.0 + 10.3) / 2.0 + 0.9 * (45.0 + 45.5) / 2.0)) < 1e-3
    assert st["action_count"][0, 1, 4] == 1 and st["visits_hypothesis"][0, 1] == 1
    assert st["total_visits"][0] == 4 and st["num_expanded"][0] == 4
    # the children counted their heuristic update under HYPOTHESIS_ID_NOT_SET
    assert st["visits_hypothesis"][1:, 2].tolist() == [1, 1, 1, 1]


def test_hypothesis_backup_matches_golden():
    from mamcts_b200.engine import Engine
    mp = default_params()
    for case in golden("hypothesis_backup.json"):
        st = Engine.new_hyp_stats(case["n_nodes"] * case["J"], case["H"], case["MA"])
        _run_batches(lambda s, *a: O.backup_hypothesis(mp, case["J"], s, *a, chance=case["chance"]), case, st)
        check_hyp_stats_against_golden(st, case["expected"])
