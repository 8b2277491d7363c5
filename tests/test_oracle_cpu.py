"""CPU tests: pin the oracle (oracle/mamcts_oracle.c) against
  (1) the committed golden vectors generated from the unmodified reference (tests/golden/), and
  (2) the compiled reference itself (oracle/_ref), when it is available,
plus the published Philox4x32-10 known-answer vectors. No GPU needed."""
from __future__ import annotations

import numpy as np
import pytest

import oracle_api as O
from helpers import assert_bit_equal, golden, tree_digest, unhex
from mamcts_b200 import default_crossing_params, default_params, default_simple_params
from mamcts_b200._abi import (ENV_CROSSING_I32, ENV_SIMPLE, SRC_PHILOX, SRC_REFERENCE_CONST,
                              SRC_REPLAY)

needs_ref = pytest.mark.skipif(not O.reference_available(), reason="compiled reference not available")


# ---- Philox -----------------------------------------------------------------------------------

def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10 (Salmon et al., SC'11).
    assert O.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert O.philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6,
                                                                    0x6D5451FD]
    assert O.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344],
                           [0xA4093822, 0x299F31D0]) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420,
                                                          0x24126EA1]


def test_philox_draw_uniform_and_in_range():
    for n in (2, 4, 7, 30):
        counts = np.zeros(n, dtype=np.int64)
        for lo in range(4000):
            for step in range(4):
                counts[O.philox_draw(1000, 3, lo, 2, step, 1, n)] += 1
        total = counts.sum()
        chi2 = float(((counts - total / n) ** 2 / (total / n)).sum())
        # 99.99% quantile of chi2 with n-1 dof is < 4 * (n - 1) + 25 for these n
        assert chi2 < 4 * (n - 1) + 25, (n, counts.tolist())


def test_draw_positions_do_not_collide():
    for A in (2, 3, 4, 5, 8, 12, 16):
        seen = set()
        for step in range(40):
            for agent in range(A):
                import ctypes as C
                b, h = C.c_uint32(), C.c_uint32()
                O.oracle().orc_draw_position(A, step, agent, C.byref(b), C.byref(h))
                assert h.value < 8
                assert (b.value, h.value) not in seen
                seen.add((b.value, h.value))


# ---- golden: execute ----------------------------------------------------------------------------

def test_crossing_execute_matches_golden():
    for case in golden("crossing_execute.json"):
        cp = default_crossing_params(num_other_agents=case["num_others"],
                                     chain_length=case["chain_length"],
                                     ego_goal_pos=case["ego_goal_pos"],
                                     cost_only_collision=case["cost_only_collision"])
        x = np.array(case["x"])
        last = np.array(case["last"])
        act = np.array(case["action"])
        rew = unhex(case["reward"])
        c0 = unhex(case["cost0"])
        for i in range(x.shape[1]):
            nx, nl, fl, r, c = O.crossing_execute(cp, x[:, i], last[:, i], act[:, i])
            assert nx == [row[i] for row in case["next_x"]]
            assert nl == [row[i] for row in case["next_last"]]
            assert fl == case["flags"][i]
            assert_bit_equal(r, rew[i], "reward")
            assert_bit_equal(c, c0[i], "cost")


def test_reference_collision_test_case():
    # environments/tests/crossing_state_int_test.cc:29-58: ego idx 2 (+1), others move +1 from x=5
    # => terminal without reaching the goal (collision at the crossing point).
    cp = default_crossing_params()
    x, last = [5, 5, 5], [2, 2, 2]
    collided = False
    for _ in range(100):
        x, last, fl, r, _c = O.crossing_execute(cp, x, last, [2, 1, 1])
        if fl & 1:
            collided = bool(fl & 4) and not (fl & 2)
            break
    assert collided and r == -1000.0


# ---- golden: rollouts ---------------------------------------------------------------------------

def _const_actions(seed, nacts):
    order = golden("expand_order.json")[str(seed)]
    return [order[str(n)][0] for n in nacts]


@pytest.mark.parametrize("mode", ["const", "replay"])
def test_crossing_rollouts_match_golden(mode):
    for case in golden("rollout_crossing.json"):
        mp = default_params(random_seed=case["seed"], rh_max_number_of_iterations=50)
        cp = default_crossing_params(num_other_agents=case["num_others"], reward_step=case["reward_step"])
        env = O.make_env(cp=cp)
        A = case["num_others"] + 1
        x = np.array(case["x"], dtype=np.int32)
        n = x.shape[1]
        steps = np.array(case["steps"], dtype=np.uint32)
        kw = {}
        if mode == "const":
            kw = dict(kind=SRC_REFERENCE_CONST, const_actions=_const_actions(case["seed"], [4] + [7] * (A - 1)))
        else:
            kw = dict(kind=SRC_REPLAY, replay=np.array(case["rec_actions"], dtype=np.int8),
                      replay_steps=steps)
        out = O.rollout_batch(env, mp, x, None, np.array(case["flags"], dtype=np.uint8),
                              np.array(case["depth"], dtype=np.uint32), K=1, trace_stride=50, **kw)
        assert out["steps"].tolist() == case["steps"]
        assert out["final_x"].tolist() == case["final_x"]
        assert out["final_last_action"].tolist() == case["final_last_action"]
        assert out["final_flags"].tolist() == case["final_flags"]
        assert_bit_equal(out["ego_return"], unhex(case["ego_return"]), "ego_return")
        assert_bit_equal(out["other_return"], unhex(case["other_return"], (A - 1, n)), "other_return")
        assert_bit_equal(out["ego_cost"], unhex(case["ego_cost"]), "ego_cost")
        assert_bit_equal(out["step_length"], unhex(case["step_length"]), "step_length")
        rec = unhex(case["rec_rewards"], (n, 50))
        for i in range(n):
            assert_bit_equal(out["reward_trace"][i, :steps[i]], rec[i, :steps[i]], "reward trace")


def test_simple_rollouts_match_golden():
    sp = default_simple_params()
    env = O.make_env(sp=sp)
    for case in golden("rollout_simple.json"):
        mp = default_params(random_seed=case["seed"], rh_max_number_of_iterations=40)
        length = np.array(case["length"], dtype=np.int32).reshape(1, -1)
        out = O.rollout_batch(env, mp, length, K=1, kind=SRC_REFERENCE_CONST,
                              const_actions=_const_actions(case["seed"], [2, 2]))
        assert out["steps"].tolist() == case["steps"]
        assert out["final_x"][0].tolist() == case["final_len"]
        assert_bit_equal(out["ego_return"], unhex(case["ego_return"]), "ego")
        assert_bit_equal(out["other_return"][0], unhex(case["other_return"]), "other")


# ---- golden: whole searches ---------------------------------------------------------------------

def _orders(seed, nacts):
    table = golden("expand_order.json")[str(seed)]
    return [table[str(n)] for n in nacts]


def _run_golden_search(case):
    if case["env"] == "simple":
        mp = default_params(random_seed=case["seed"], max_number_of_iterations=case["iterations"],
                            max_number_of_nodes=case["max_nodes"])
        env = O.make_env(sp=default_simple_params())
        root, nacts, ma = [case["state_length"]], [2, 2], 2
    else:
        mp = default_params(random_seed=case["seed"], max_number_of_iterations=case["iterations"],
                            max_number_of_nodes=100000000, rh_max_number_of_iterations=50)
        cp = default_crossing_params(num_other_agents=case["num_others"],
                                     chain_length=case["chain_length"], ego_goal_pos=case["ego_goal_pos"])
        env = O.make_env(cp=cp)
        A = case["num_others"] + 1
        root, nacts, ma = [5] * A, [4] + [7] * (A - 1), 7
    t = O.OracleTree(env, mp, root, action_source=SRC_REFERENCE_CONST,
                     expand_orders=_orders(case["seed"], nacts))
    t.run(case["iterations"])
    return t, nacts, ma


def test_search_matches_golden():
    for case in golden("search.json"):
        t, nacts, ma = _run_golden_search(case)
        A = len(nacts)
        assert t.num_nodes == case["num_nodes"]
        gq = unhex(case["q"], (A, ma))
        for a in range(A):
            st = t.root_stat(a)
            assert st["visits"] == case["visits"][a]
            assert st["n"] == case["n"][a][:nacts[a]]
            assert_bit_equal(st["q"], gq[a, :nacts[a]], f"root q agent {a}")
            assert_bit_equal(st["value"], unhex(case["value"])[a], "root value")
        assert tree_digest(t.dump(ma, case["tree_count"])) == case["tree_sha256"]
        t.close()


def test_survey_known_answers():
    # 10 000 iterations, 1 other, seed 1000 -> root ego statistics of the reference.
    case = [c for c in golden("search.json") if c["env"] == "crossing" and c["iterations"] == 10000][0]
    assert case["n"][0][:4] == [7666, 2017, 279, 38]
    assert unhex(case["q"])[0] == 0.4601886772625664
    assert case["num_nodes"] == 10001 and case["best"] == 0
    # rollout from the default start, seed 1000: [2 4 4] x 7, ego return 100 * 0.9^14 * (1/4)^7
    mp = default_params(rh_max_number_of_iterations=50)
    env = O.make_env(cp=default_crossing_params())
    r = O.rollout(env, mp, [5, 5, 5], kind=SRC_REFERENCE_CONST, const_actions=[2, 4, 4])
    assert r["steps"] == 7 and r["ego_return"] == 0.0013962886019873665
    assert r["cost0"] == -0.006103515625 and r["final_x"] == [12, 33, 33]


# ---- golden: merge ------------------------------------------------------------------------------

def test_merge_matches_golden():
    import ctypes as C
    for case in golden("merge.json"):
        nA = case["num_actions"]
        cnt = np.array(case["cnt"])
        q = unhex(case["q"], cnt.shape)
        stats = []
        for t in range(cnt.shape[0]):
            s = O.OrcStat()
            s.num_actions = nA
            s.total_visits = case["visits"][t]
            for a in range(nA):
                if cnt[t, a] >= 0:
                    s.expanded[a] = 1
                    s.n[a] = int(cnt[t, a])
                    s.q[a] = q[t, a]
                    s.num_expanded += 1
            stats.append(s)
        arr = (C.POINTER(O.OrcStat) * len(stats))(*[C.pointer(s) for s in stats])
        merged = O.OrcStat()
        merged.num_actions = nA
        O.oracle().orc_stat_merge(arr, len(stats), C.byref(merged))
        mq = unhex(case["merged_q"])
        for a in range(nA):
            if case["merged_n"][a] >= 0:
                assert merged.expanded[a]
                assert merged.n[a] == case["merged_n"][a]
                assert_bit_equal(merged.q[a], mq[a], "merged q")
            else:
                assert not merged.expanded[a]
        assert merged.total_visits == case["merged_visits"]
        assert_bit_equal(merged.value, float.fromhex(case["merged_value"]), "merged value")
        assert O.oracle().orc_stat_best_action(C.byref(merged)) == case["best"]


# ---- reduction order ----------------------------------------------------------------------------

def test_reduce_mean_close_to_numpy():
    rng = np.random.default_rng(3)
    for K in (1, 2, 4, 8, 16, 32, 64, 96, 256, 512, 1024):
        v = rng.normal(size=K) * 100
        assert abs(O.reduce_mean(v) - v.mean()) <= 1e-12 * max(1.0, abs(v).max())


# ---- oracle against the compiled reference (fresh inputs, not only the fixtures) ----------------

@needs_ref
def test_oracle_rollouts_equal_reference_random_states():
    rng = np.random.default_rng(2024)
    for num_others, seed in ((1, 3), (2, 1000), (3, 11), (7, 5)):
        mp = default_params(random_seed=seed, rh_max_number_of_iterations=30)
        cp = default_crossing_params(num_other_agents=num_others, cost_only_collision=int(seed % 2))
        A = num_others + 1
        n = 40
        x = rng.integers(0, 14, size=(A, n)).astype(np.int32)
        ref = O.ref_rollout_crossing(mp, cp, x, rec_stride=30)
        env = O.make_env(cp=cp)
        out = O.rollout_batch(env, mp, x, K=1, kind=SRC_REPLAY, replay=ref["rec_actions"],
                              replay_steps=ref["steps"])
        assert out["steps"].tolist() == ref["steps"].tolist()
        assert out["final_x"].tolist() == ref["final_x"].tolist()
        assert out["final_flags"].tolist() == ref["final_flags"].tolist()
        assert_bit_equal(out["ego_return"], ref["ego_return"], "ego")
        assert_bit_equal(out["ego_cost"], ref["ego_cost"], "cost")


@needs_ref
def test_oracle_search_equals_reference_other_params():
    # non-default exploration / widening / discount, fixed bounds
    mp = default_params(random_seed=5, max_number_of_iterations=1500, max_number_of_nodes=100000000,
                        rh_max_number_of_iterations=50, discount_factor=0.95,
                        uct_exploration_constant=1.3, uct_pw_k=2.0, uct_pw_alpha=0.4)
    cp = default_crossing_params(num_other_agents=2)
    ref = O.ref_search_crossing(mp, cp, max_actions=7, tree_capacity=1501)
    t = O.OracleTree(O.make_env(cp=cp), mp, [5, 5, 5], action_source=SRC_REFERENCE_CONST)
    t.run(1500)
    assert t.num_nodes == ref["num_nodes"]
    assert np.array_equal(t.dump(7), ref["tree"])
    # node budget: the reference stops expanding once num_nodes_ > MAX_NUMBER_OF_NODES
    mp2 = default_params(max_number_of_iterations=400, max_number_of_nodes=50, rh_max_number_of_iterations=50)
    ref2 = O.ref_search_crossing(mp2, cp, max_actions=7, tree_capacity=401)
    t2 = O.OracleTree(O.make_env(cp=cp), mp2, [5, 5, 5], action_source=SRC_REFERENCE_CONST)
    t2.run(400)
    assert t2.num_nodes == ref2["num_nodes"] == 51
    assert np.array_equal(t2.dump(7), ref2["tree"])


@needs_ref
def test_philox_twin_equals_oracle_tree():
    # The reference's own Mcts/StageNode/UctStatistic driven by the Philox heuristic (the stochastic
    # twin of SURVEY.md §8c) must build the same tree as the oracle's arena search.
    for env_kind, num_others in ((ENV_CROSSING_I32, 1), (ENV_CROSSING_I32, 2), (ENV_SIMPLE, 1)):
        mp = default_params(max_number_of_iterations=800, max_number_of_nodes=100000000,
                            rh_max_number_of_iterations=50)
        if env_kind == ENV_CROSSING_I32:
            cp, sp = default_crossing_params(num_other_agents=num_others), None
            env, root, ma = O.make_env(cp=cp), [5] * (num_others + 1), 7
        else:
            cp, sp = None, default_simple_params()
            env, root, ma = O.make_env(sp=sp), [4], 2
        for tree_id in (0, 3):
            ref = O.ref_search_philox(env_kind, mp, cp, sp, root_x=root, seed=99, tree_id=tree_id, K=1,
                                      max_actions=ma, tree_capacity=801)
            t = O.OracleTree(env, mp, root, action_source=SRC_PHILOX, seed=99, tree_id=tree_id)
            t.run(800)
            assert t.num_nodes == ref["num_nodes"]
            assert t.rollout_steps == ref["rollout_steps"]
            assert np.array_equal(t.dump(ma), ref["tree"])
            t.close()


@needs_ref
def test_expand_order_golden_equals_reference():
    table = golden("expand_order.json")
    for seed, per_n in table.items():
        mp = default_params(random_seed=int(seed))
        for n, order in per_n.items():
            assert O.ref_expand_order(mp, int(n)) == order


@pytest.mark.skipif(not O.reference_available(), reason="compiled reference not on this box")
def test_oracle_action_values_equal_reference():
    """ActionValueRandomHeuristic's per-ego-action first step + rollout (action_value_random_heuristic.h:37-85),
    executed by the REFERENCE'S OWN loop (compiled with the two test-only statistic adapters of ref_driver.cc),
    against the restatement orc_action_values: action returns, action costs, the other agents' returns and
    the mean cost handed to their statistics, bit for bit - random states, depths near the depth limit,
    terminal states, other parameter sets and seeds (the first-draw constants change with the seed)."""
    rng = np.random.default_rng(11)
    for others, seed, kw in ((1, 1000, {}), (2, 1000, {}), (3, 7, {}), (2, 42, dict(chain_length=41, ego_goal_pos=26)),
                             (1, 5, dict(reward_step=-1.0, cost_only_collision=1))):
        cp = default_crossing_params(num_other_agents=others, **kw)
        mp = default_params(random_seed=seed, rh_max_number_of_iterations=50, max_search_depth=60)
        A, n = others + 1, 160
        x = rng.integers(-1, cp.chain_length + 3, size=(A, n)).astype(np.int32)
        last = rng.integers(-2, 4, size=(A, n)).astype(np.int32)
        flags = (rng.random(n) < 0.1).astype(np.uint8)
        depth = rng.choice([0, 3, 20, 58, 59, 60], size=n).astype(np.uint32)
        ref = O.ref_action_values_crossing(mp, cp, x, last, flags, depth)
        env = O.make_env(cp=cp)
        const = [O.ref_expand_order(mp, env.num_actions[a])[0] for a in range(A)]
        orc = O.action_values_batch(env, mp, x, last, flags, depth, kind=SRC_REFERENCE_CONST, const_actions=const)
        assert_bit_equal(orc["action_return"], ref["action_return"], "action returns")
        assert_bit_equal(orc["action_cost"], ref["action_cost"], "action costs")
        assert_bit_equal(orc["other_return"], ref["other_return"], "other returns")
        assert_bit_equal(orc["mean_cost"], ref["mean_cost"], "mean cost")
        assert np.all(orc["action_return"][flags == 1] == 0.0)


def test_oracle_action_values_equal_golden():
    """The committed vectors of the reference's ActionValueRandomHeuristic loop (tests/golden/action_values.json,
    generated by make_golden.py from oracle/_ref) against the restatement - runs wherever the oracle builds."""
    for case in golden("action_values.json"):
        cp = default_crossing_params(num_other_agents=case["num_others"], **case["params"])
        mp = default_params(random_seed=case["seed"], rh_max_number_of_iterations=50, max_search_depth=60)
        x = np.array(case["x"], dtype=np.int32)
        n, nA = x.shape[1], cp.num_ego_actions
        orc = O.action_values_batch(O.make_env(cp=cp), mp, x, np.array(case["last"]), np.array(case["flags"], dtype=np.uint8),
                                    np.array(case["depth"]), kind=SRC_REFERENCE_CONST, const_actions=case["const_actions"])
        assert_bit_equal(orc["action_return"], unhex(case["action_return"], (n, nA)), "action returns")
        assert_bit_equal(orc["action_cost"], unhex(case["action_cost"], (n, nA)), "action costs")
        assert_bit_equal(orc["other_return"], unhex(case["other_return"], (case["num_others"], n)), "other returns")
        assert_bit_equal(orc["mean_cost"], unhex(case["mean_cost"]), "mean cost")
