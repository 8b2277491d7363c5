"""GPU parity tests for the backup kernel, the root merge and the device-resident search, through
the C ABI. REFERENCE_CONST searches must reproduce the reference's tree bit for bit (golden digests
from the unmodified reference); PHILOX searches must equal the CPU oracle's tree driven by the same
streams, and the reference's own Mcts/StageNode/UctStatistic when oracle/_ref is on the box."""
from __future__ import annotations

import numpy as np
import pytest

import oracle_api as O
from helpers import assert_bit_equal, golden, rel_close, tree_digest, unhex
from mamcts_b200 import default_crossing_params, default_params, default_simple_params
from mamcts_b200._abi import (ENV_CROSSING_I32, ENV_SIMPLE, LAYOUT_AUTO, LAYOUT_CLASSIC, LAYOUT_COMPACT,
                              SRC_PHILOX, SRC_REFERENCE_CONST)
from mamcts_b200.engine import MamctsError, Search, reference_expand_order

pytestmark = pytest.mark.gpu


# ---- backup kernel ------------------------------------------------------------------------------

def _oracle_backup(mp, A, MA, n_nodes, paths):
    """CPU statement of the same batch of path backups with orc_stat_* (oracle)."""
    import ctypes as C
    lib = O.oracle()
    lib.orc_stat_backprop.argtypes = [C.c_void_p, C.c_double, C.c_double]
    lib.orc_stat_update_from_heuristic.argtypes = [C.c_void_p, C.c_double]
    stats = [[O.OrcStat() for _ in range(A)] for _ in range(n_nodes)]
    for nd in stats:
        for s in nd:
            s.num_actions = MA
    return stats, lib


def test_backup_matches_oracle(engine):
    import ctypes as C
    rng = np.random.default_rng(5)
    A, MA = 3, 7
    mp = default_params(discount_factor=0.9)
    n_paths, L = 500, 9
    n_nodes = n_paths * (L + 1)
    stats = dict(value=np.zeros(n_nodes * A), latest_return=np.zeros(n_nodes * A),
                 total_visits=np.zeros(n_nodes * A, dtype=np.uint32), q=np.zeros((n_nodes * A, MA)),
                 n=np.zeros((n_nodes * A, MA), dtype=np.uint32))
    ostats, lib = _oracle_backup(mp, A, MA, n_nodes, None)
    # three rounds over the same trees so that running means are exercised beyond the first visit
    for rnd in range(3):
        leaf = np.arange(n_paths, dtype=np.uint32) * (L + 1) + L
        has = (rng.random(n_paths) < 0.8).astype(np.uint8)
        ego = rng.normal(size=n_paths)
        oth = rng.normal(size=(A - 1, n_paths))
        plen = rng.integers(0, L + 1, size=n_paths).astype(np.uint32)
        off = np.concatenate([[0], np.cumsum(plen)[:-1]]).astype(np.uint32)
        ne = int(plen.sum())
        enode = np.zeros(ne, dtype=np.uint32)
        eact = rng.integers(0, MA, size=(ne, A)).astype(np.uint8)
        erew = rng.choice([0.0, -1000.0, 100.0, 1.0], size=(ne, A))
        for p in range(n_paths):
            for e in range(plen[p]):
                enode[off[p] + e] = p * (L + 1) + (L - 1 - e)
        engine.backup_uct(mp, A, leaf, has, ego, oth, off, plen, enode, eact, erew, stats)
        for p in range(n_paths):
            for a in range(A):
                ls = ostats[leaf[p]][a]
                if has[p]:
                    lib.orc_stat_update_from_heuristic(C.byref(ls), ego[p] if a == 0 else oth[a - 1, p])
                G = ls.latest_return
                for e in range(plen[p]):
                    s = ostats[enode[off[p] + e]][a]
                    s.collected_action = int(eact[off[p] + e, a])
                    s.collected_reward = float(erew[off[p] + e, a])
                    lib.orc_stat_backprop(C.byref(s), mp.discount_factor, G)
                    G = s.latest_return
    exp_v = np.array([ostats[i][a].value for i in range(n_nodes) for a in range(A)])
    exp_g = np.array([ostats[i][a].latest_return for i in range(n_nodes) for a in range(A)])
    exp_N = np.array([ostats[i][a].total_visits for i in range(n_nodes) for a in range(A)])
    exp_q = np.array([[ostats[i][a].q[k] for k in range(MA)] for i in range(n_nodes) for a in range(A)])
    exp_n = np.array([[ostats[i][a].n[k] for k in range(MA)] for i in range(n_nodes) for a in range(A)])
    assert_bit_equal(stats["value"], exp_v, "value")
    assert_bit_equal(stats["latest_return"], exp_g, "latest")
    assert stats["total_visits"].tolist() == exp_N.tolist()
    assert_bit_equal(stats["q"], exp_q, "q")
    assert stats["n"].tolist() == exp_n.tolist()


def test_hypothesis_backup_matches_reference_golden_and_oracle(engine):
    """mamcts_backup_hypothesis (C4: the other agents' HypothesisStatistic, hypothesis_statistic.h:100-134):
    golden results produced by REAL HypothesisStatistic objects of the reference, bit for bit; then a
    larger random batch against the oracle."""
    from test_hypothesis_cpu import _run_batches, check_hyp_stats_against_golden
    from mamcts_b200.engine import Engine
    mp = default_params()
    for case in golden("hypothesis_backup.json"):
        st = Engine.new_hyp_stats(case["n_nodes"] * case["J"], case["H"], case["MA"])
        _run_batches(lambda s, *a: engine.backup_hypothesis(mp, case["J"], *a, s, chance_constrained_update=case["chance"]),
                     case, st)
        check_hyp_stats_against_golden(st, case["expected"])
    # 3 000 paths x 3 other agents, three rounds over growing chains
    rng = np.random.default_rng(8)
    J, H, MA, Cn, T, L = 3, 4, 9, 2, 3000, 4
    a = Engine.new_hyp_stats(T * (L + 1) * J, H, MA)
    b = Engine.new_hyp_stats(T * (L + 1) * J, H, MA)
    depth = np.zeros(T, dtype=np.int64)
    for rnd in range(4):
        grow = (rng.random(T) < 0.7) & (depth < L)
        depth += grow
        leaf = np.arange(T) * (L + 1) + depth
        off = np.concatenate([[0], np.cumsum(depth)[:-1]])
        ne = int(depth.sum())
        en = np.zeros(ne, dtype=np.uint32)
        for t in range(T):
            en[off[t]:off[t] + depth[t]] = t * (L + 1) + np.arange(depth[t] - 1, -1, -1)
        eh, ea = rng.integers(0, H, (ne, J)), rng.integers(0, MA, (ne, J))
        ec, lc = rng.choice([0.0, 1000.0, -1.5], (ne, Cn)), rng.normal(size=(T, Cn))
        engine.backup_hypothesis(mp, J, leaf, grow, lc, off, depth, en, eh, ea, ec, a)
        O.backup_hypothesis(mp, J, b, leaf, grow, lc, off, depth, en, eh, ea, ec)
    for k in a:
        if a[k].dtype == np.float64:
            assert_bit_equal(a[k], b[k], k)
        else:
            assert np.array_equal(a[k], b[k]), k
    with pytest.raises(MamctsError, match="out of range"):
        engine.backup_hypothesis(mp, J, [0], [1], [[0.0, 0.0]], [0], [1], [0], [[H, 0, 0]], [[0, 0, 0]], [[0.0, 0.0]], a)


def test_root_merge_matches_golden(engine):
    for case in golden("merge.json"):
        nA = case["num_actions"]
        cnt = np.array(case["cnt"])
        T = cnt.shape[0]
        q = unhex(case["q"], cnt.shape)
        stats = dict(value=np.zeros(T), latest_return=np.zeros(T),
                     total_visits=np.array(case["visits"], dtype=np.uint32),
                     q=np.ascontiguousarray(q), n=np.maximum(cnt, 0).astype(np.uint32))
        m = engine.root_merge(stats, np.arange(T, dtype=np.uint32), nA)
        mq = unhex(case["merged_q"])
        present = np.array(case["merged_n"]) >= 0
        assert rel_close(m["q"][present], mq[present], 1e-12)
        assert m["n"][present].tolist() == np.array(case["merged_n"])[present].tolist()
        assert m["visits"] == case["merged_visits"]
        assert rel_close(m["value"], float.fromhex(case["merged_value"]), 1e-12)
        assert m["best"] == case["best"]


def test_root_merge_many_trees_multi_block(engine):
    """The multi-block warp-shuffle reduction (reduce.cu): 70 001 roots (274 blocks + a ragged tail), actions that
    are missing in some trees, against the host statement of merge_node_statistics (sharding.local_partial /
    finish); counts exact, sums of Q to rounding; two calls give bit-identical results (fixed order)."""
    from mamcts_b200 import sharding
    rng = np.random.default_rng(77)
    T, nA, MA = 70001, 5, 7
    n = rng.integers(0, 4000, size=(T, MA)).astype(np.uint32)
    n[rng.random((T, MA)) < 0.2] = 0
    q = rng.normal(size=(T, MA)) * 50.0
    vis = rng.integers(0, 10000, size=T).astype(np.uint32)
    stats = dict(value=np.zeros(T), latest_return=np.zeros(T), total_visits=vis, q=np.ascontiguousarray(q),
                 n=np.ascontiguousarray(n))
    slots = rng.permutation(T).astype(np.uint32)
    m1 = engine.root_merge(stats, slots, nA)
    m2 = engine.root_merge(stats, slots, nA)
    exp = sharding.finish(sharding.local_partial(q[:, :nA], n[:, :nA], n[:, :nA] > 0, vis), nA)
    assert m1["n"].tolist() == exp["n"].tolist() and m1["occurrences"].tolist() == exp["occurrences"].tolist()
    assert m1["visits"] == exp["visits"] and m1["best"] == exp["best"]
    assert rel_close(m1["q"], exp["q"], 1e-9) and rel_close(m1["value"], exp["value"], 1e-9)
    assert_bit_equal(m1["q"], m2["q"], "root merge is deterministic")
    with pytest.raises(MamctsError, match="out of range"):
        engine.root_merge(stats, np.array([T], dtype=np.uint32), nA)


def test_rollout_moments_device_reduction(engine):
    """mamcts_rollout_moments (C3's exchange step): {sum, sum^2} per agent + count, reduced on the device in a
    fixed order; against numpy (to rounding) and against itself (bit-identical on repetition)."""
    from mamcts_b200 import sharding
    rng = np.random.default_rng(3)
    for n, others in ((1, 0), (37, 1), (4096, 3), (300007, 7)):
        ego = rng.normal(size=n) * 3.0
        oth = rng.normal(size=(others, n)) if others else None
        a = engine.rollout_moments(ego, oth)
        b = engine.rollout_moments(ego, oth)
        assert_bit_equal(a, b, "moments are deterministic")
        rows = np.vstack([ego[None, :]] + ([oth] if others else []))
        exp = sharding.moments_partial(rows)
        assert a[-1] == n and a.shape == exp.shape
        assert np.allclose(a, exp, rtol=1e-11, atol=1e-9)
        fin = sharding.moments_finish(a)
        assert np.allclose(fin["mean"], rows.mean(axis=1), rtol=1e-9, atol=1e-12)


# ---- device-resident search ---------------------------------------------------------------------

def _golden_search(engine, case, num_trees=3, layout=LAYOUT_AUTO):
    if case["env"] == "simple":
        mp = default_params(random_seed=case["seed"], max_number_of_iterations=case["iterations"],
                            max_number_of_nodes=case["max_nodes"])
        s = Search(engine, ENV_SIMPLE, mp, num_trees, sp=default_simple_params(),
                   root_x=[case["state_length"]], action_source=SRC_REFERENCE_CONST, layout=layout)
        nacts, ma = [2, 2], 2
    else:
        mp = default_params(random_seed=case["seed"], max_number_of_iterations=case["iterations"],
                            max_number_of_nodes=100000000, rh_max_number_of_iterations=50)
        cp = default_crossing_params(num_other_agents=case["num_others"],
                                     chain_length=case["chain_length"], ego_goal_pos=case["ego_goal_pos"])
        s = Search(engine, ENV_CROSSING_I32, mp, num_trees, cp=cp, action_source=SRC_REFERENCE_CONST, layout=layout)
        nacts, ma = [4] + [7] * case["num_others"], 7
    return s, nacts, ma


def test_expand_order_matches_golden():
    for seed, per_n in golden("expand_order.json").items():
        for n, order in per_n.items():
            assert reference_expand_order(int(seed), int(n)) == order


@pytest.mark.parametrize("layout", [LAYOUT_CLASSIC, LAYOUT_AUTO])
def test_reference_const_search_reproduces_reference_tree(engine, layout):
    """Whole Mcts::search: every statistic of every node equals the reference's, bit for bit
    (classic layout, and the compact one wherever AUTO selects it: 2 agents, SimpleState)."""
    for case in golden("search.json"):
        s, nacts, ma = _golden_search(engine, case, layout=layout)
        # in two launches: the trees must carry on exactly where they stopped
        first = case["iterations"] // 3
        s.run(first)
        s.run(case["iterations"] - first)
        A = len(nacts)
        gq = unhex(case["q"], (A, ma))
        for tree in (0, 2):  # the reference's parallel trees are identical
            for a in range(A):
                st = s.root_stats(tree, a)
                assert st["num_nodes"] == case["num_nodes"]
                assert st["visits"] == case["visits"][a]
                assert st["n"] == case["n"][a][:nacts[a]]
                assert_bit_equal(st["q"], gq[a, :nacts[a]], f"root q agent {a}")
                assert_bit_equal(st["value"], unhex(case["value"])[a], "root value")
            assert tree_digest(s.dump_tree(tree, case["tree_count"])) == case["tree_sha256"]
        c = s.counters()
        assert c["iterations"] == 3 * case["iterations"]
        assert c["nodes"] == 3 * case["num_nodes"]
        m = s.merge_roots()
        assert m["best"] == case["best"]
        present = np.array(case["n"][0][:nacts[0]]) >= 0
        assert rel_close(m["q"][present], gq[0, :nacts[0]][present], 1e-12)
        assert m["n"][present].tolist() == (3 * np.array(case["n"][0][:nacts[0]])[present]).tolist()
        s.close()


@pytest.mark.parametrize("num_others,layout", [(1, LAYOUT_CLASSIC), (1, LAYOUT_COMPACT), (2, LAYOUT_AUTO),
                                               (3, LAYOUT_AUTO), (7, LAYOUT_AUTO), (2, LAYOUT_CLASSIC),
                                               (3, LAYOUT_COMPACT), (7, LAYOUT_CLASSIC), (7, LAYOUT_COMPACT),
                                               (4, LAYOUT_COMPACT), (4, LAYOUT_CLASSIC), (5, LAYOUT_AUTO),
                                               (5, LAYOUT_CLASSIC), (6, LAYOUT_COMPACT), (6, LAYOUT_CLASSIC)])
def test_philox_search_equals_oracle_tree(engine, num_others, layout):
    iters = 600
    mp = default_params(max_number_of_iterations=iters, max_number_of_nodes=100000000,
                        rh_max_number_of_iterations=50)
    cp = default_crossing_params(num_other_agents=num_others)
    T, first = 70, 1000
    s = Search(engine, ENV_CROSSING_I32, mp, T, cp=cp, action_source=SRC_PHILOX, philox_seed=4242,
               first_tree_id=first, layout=layout)
    s.run(iters // 2)       # two launches: the trees carry on exactly where they stopped
    s.run(iters - iters // 2)
    env = O.make_env(cp=cp)
    total_steps = 0
    for tree in (0, 1, 33, 69):
        t = O.OracleTree(env, mp, [5] * (num_others + 1), action_source=SRC_PHILOX, seed=4242,
                         tree_id=first + tree)
        t.run(iters)
        assert s.root_stats(tree, 0)["num_nodes"] == t.num_nodes
        got = s.dump_tree(tree, t.num_nodes)
        assert np.array_equal(got, t.dump(7)), f"tree {tree} differs from the oracle"
        t.close()
    # trees with different ids are different trees (unlike the reference's identical ones)
    assert not np.array_equal(s.dump_tree(0, iters + 1), s.dump_tree(1, iters + 1))
    c = s.counters()
    assert c["iterations"] == T * iters
    s.close()


@pytest.mark.parametrize("layout", [LAYOUT_CLASSIC, LAYOUT_COMPACT])
def test_philox_search_simple_equals_oracle_tree(engine, layout):
    iters = 300
    mp = default_params(max_number_of_iterations=iters, max_number_of_nodes=100,
                        rh_max_number_of_iterations=30)
    sp = default_simple_params()
    s = Search(engine, ENV_SIMPLE, mp, 40, sp=sp, root_x=[4], action_source=SRC_PHILOX, philox_seed=1,
               layout=layout)
    s.run(iters)
    for tree in (0, 39):
        t = O.OracleTree(O.make_env(sp=sp), mp, [4], action_source=SRC_PHILOX, seed=1, tree_id=tree)
        t.run(iters)
        assert np.array_equal(s.dump_tree(tree, t.num_nodes), t.dump(2))
        assert s.root_stats(tree, 0)["num_nodes"] == t.num_nodes == 101  # MAX_NUMBER_OF_NODES + 1
        t.close()
    s.close()


@pytest.mark.skipif(not O.reference_available(), reason="compiled reference not on this box")
def test_philox_search_equals_reference_twin(engine):
    """The reference's own Mcts<>/StageNode/UctStatistic code with the Philox heuristic."""
    iters = 500
    mp = default_params(max_number_of_iterations=iters, max_number_of_nodes=100000000,
                        rh_max_number_of_iterations=50)
    cp = default_crossing_params(num_other_agents=2)
    s = Search(engine, ENV_CROSSING_I32, mp, 8, cp=cp, action_source=SRC_PHILOX, philox_seed=99)
    s.run(iters)
    steps = 0
    for tree in range(8):
        ref = O.ref_search_philox(ENV_CROSSING_I32, mp, cp, None, root_x=[5, 5, 5], seed=99,
                                  tree_id=tree, K=1, max_actions=7, tree_capacity=iters + 1)
        assert np.array_equal(s.dump_tree(tree, iters + 1), ref["tree"])
        steps += ref["rollout_steps"]
    assert s.counters()["rollout_steps"] == steps
    # root-parallel merge == UctStatistic::merge_node_statistics over the 8 reference roots
    q = np.zeros((8, 7)); cnt = np.full((8, 7), -1, dtype=np.int64); vis = np.zeros(8, dtype=np.uint32)
    for tree in range(8):
        st = s.root_stats(tree, 0)
        q[tree, :4], cnt[tree, :4], vis[tree] = st["q"], st["n"], st["visits"]
    exp = O.ref_merge_uct(mp, q, cnt, vis, 4)
    m = s.merge_roots()
    assert rel_close(m["q"], exp["q"][:4], 1e-6)
    assert m["n"].tolist() == exp["n"][:4].tolist()
    assert m["visits"] == exp["visits"] and m["best"] == exp["best"]
    assert rel_close(m["value"], exp["value"], 1e-6)
    s.close()


def test_multi_agent_compact_layout_agrees_with_classic_on_large_trees(engine):
    """3, 4 and 8 agents (the reference's own 4-agent setting included: chain 41, goal 26,
    environments/tests/crossing_state_int_test.cc:231-252) at the benchmark's tree size: the counts-only compact
    layout (search_compact_multi.cuh) and the classic one hold the same trees, statistic for statistic, and one of
    them is checked against the CPU oracle; 70 000-iteration configurations use the 32-bit-count records."""
    for others, kw, iters, max_iter in ((2, {}, 10000, 10000), (3, dict(chain_length=41, ego_goal_pos=26), 6000, 6000),
                                        (7, {}, 10000, 10000), (7, {}, 1500, 70000)):
        mp = default_params(max_number_of_iterations=max_iter, max_number_of_nodes=20000 if max_iter > 65000 else 100000000,
                            rh_max_number_of_iterations=50)
        cp = default_crossing_params(num_other_agents=others, **kw)
        dumps, merged, counters = [], [], []
        for layout in (LAYOUT_CLASSIC, LAYOUT_COMPACT):
            s = Search(engine, ENV_CROSSING_I32, mp, 40, cp=cp, action_source=SRC_PHILOX, philox_seed=1000,
                       first_tree_id=3, layout=layout)
            s.run(iters // 2)
            s.run(iters - iters // 2)
            dumps.append([s.dump_tree(t, iters + 1) for t in (0, 39)])
            merged.append(s.merge_roots())
            counters.append(s.counters())
            for agent in (0, others):
                assert s.root_stats(5, agent)["visits"] == iters
            s.close()
        for a, b in zip(*dumps):
            assert a.shape == b.shape and np.array_equal(a, b)
        assert counters[0] == counters[1]
        assert merged[0]["n"].tolist() == merged[1]["n"].tolist() and merged[0]["best"] == merged[1]["best"]
        assert_bit_equal(merged[0]["q"], merged[1]["q"], "merged q")
        t = O.OracleTree(O.make_env(cp=cp), mp, [5] * (others + 1), action_source=SRC_PHILOX, seed=1000, tree_id=3 + 39)
        t.run(iters)
        assert np.array_equal(dumps[1][1], t.dump(7))
        t.close()


def test_compact_and_classic_layouts_agree_on_large_trees(engine):
    """The two tree layouts are two storages of the same search: every node statistic of every tree is
    bit-identical, at the benchmark's tree size (10 000 iterations), and the merged roots agree."""
    iters = 10000
    mp = default_params(max_number_of_iterations=iters, max_number_of_nodes=100000000,
                        rh_max_number_of_iterations=50)
    cp = default_crossing_params(num_other_agents=1)
    dumps, merged, counters = [], [], []
    for layout, inner in ((LAYOUT_CLASSIC, 0), (LAYOUT_COMPACT, 3200)):
        s = Search(engine, ENV_CROSSING_I32, mp, 96, cp=cp, action_source=SRC_PHILOX, philox_seed=1000,
                   first_tree_id=7, layout=layout, max_inner_nodes_per_tree=inner)
        s.run(iters)
        dumps.append([s.dump_tree(t, iters + 1) for t in (0, 31, 95)])
        merged.append(s.merge_roots())
        counters.append(s.counters())
        s.close()
    for a, b in zip(*dumps):
        assert a.shape == b.shape and np.array_equal(a, b)
    assert counters[0] == counters[1]
    assert merged[0]["n"].tolist() == merged[1]["n"].tolist() and merged[0]["best"] == merged[1]["best"]
    assert_bit_equal(merged[0]["q"], merged[1]["q"], "merged q")
    # ... and against the CPU oracle for one of them
    t = O.OracleTree(O.make_env(cp=cp), mp, [5, 5], action_source=SRC_PHILOX, seed=1000, tree_id=7 + 31)
    t.run(iters)
    assert np.array_equal(dumps[1][1], t.dump(7))
    t.close()


@pytest.mark.parametrize("max_iter,max_nodes", [(70000, 5000), (70000, 100000000)])
def test_compact_layout_wide_counters(engine, max_iter, max_nodes):
    """The compact record has three instantiations: 16-bit counts and child entries (searches below
    65 535 iterations: every other test), 32-bit counts with 16-bit entries, and 32-bit both; the
    latter two are selected by MAX_NUMBER_OF_ITERATIONS / the node capacity of the configuration."""
    iters = 700
    mp = default_params(max_number_of_iterations=max_iter, max_number_of_nodes=max_nodes,
                        rh_max_number_of_iterations=50)
    cp = default_crossing_params(num_other_agents=1)
    s = Search(engine, ENV_CROSSING_I32, mp, 40, cp=cp, action_source=SRC_PHILOX, philox_seed=31,
               layout=LAYOUT_COMPACT, max_inner_nodes_per_tree=1000)
    s.run(iters)
    for tree in (0, 39):
        t = O.OracleTree(O.make_env(cp=cp), mp, [5, 5], action_source=SRC_PHILOX, seed=31, tree_id=tree)
        t.run(iters)
        assert np.array_equal(s.dump_tree(tree, iters + 1), t.dump(7))
        t.close()
    m = s.merge_roots()
    assert m["visits"] == 40 * iters
    s.close()


@pytest.mark.parametrize("num_others,layout,ego_v,n_oth", [(1, LAYOUT_COMPACT, (-1, 3), 5), (1, LAYOUT_CLASSIC, (-2, 1), 8),
                                                          (2, LAYOUT_AUTO, (0, 2), 6), (3, LAYOUT_AUTO, (-1, 1), 4)])
def test_other_action_counts_equal_oracle(engine, num_others, layout, ego_v, n_oth):
    """CrossingStateParameters lets the velocity ranges be chosen freely: searches with other action
    counts than the default 4 / 7 run on the generic (up to 8 actions) instantiation."""
    iters = 500
    mp = default_params(max_number_of_iterations=iters, max_number_of_nodes=100000000,
                        rh_max_number_of_iterations=50, random_seed=77)
    cp = default_crossing_params(num_other_agents=num_others, min_velocity_ego=ego_v[0], max_velocity_ego=ego_v[1],
                                 num_other_actions=n_oth)
    s = Search(engine, ENV_CROSSING_I32, mp, 20, cp=cp, action_source=SRC_PHILOX, philox_seed=5, layout=layout)
    s.run(iters)
    ma = max(ego_v[1] - ego_v[0] + 1, n_oth)
    for tree in (0, 19):
        t = O.OracleTree(O.make_env(cp=cp), mp, [5] * (num_others + 1), action_source=SRC_PHILOX, seed=5, tree_id=tree)
        t.run(iters)
        assert np.array_equal(s.dump_tree(tree, t.num_nodes), t.dump(ma))
        t.close()
    m = s.merge_roots()
    assert m["visits"] == 20 * iters and len(m["q"]) == ego_v[1] - ego_v[0] + 1
    s.close()


def test_compact_layout_reports_exhausted_inner_records(engine):
    mp = default_params(max_number_of_iterations=2000, max_number_of_nodes=100000000,
                        rh_max_number_of_iterations=50)
    cp = default_crossing_params(num_other_agents=1)
    s = Search(engine, ENV_CROSSING_I32, mp, 8, cp=cp, action_source=SRC_PHILOX, philox_seed=5,
               layout=LAYOUT_COMPACT, max_inner_nodes_per_tree=50)
    s.run(2000)
    with pytest.raises(MamctsError, match="inner node records"):
        s.merge_roots()
    with pytest.raises(MamctsError, match="inner node records"):
        s.counters()
    s.close()
    # the same for the multi-agent compact layout
    s = Search(engine, ENV_CROSSING_I32, mp, 8, cp=default_crossing_params(num_other_agents=3), action_source=SRC_PHILOX,
               philox_seed=5, layout=LAYOUT_COMPACT, max_inner_nodes_per_tree=50)
    s.run(2000)
    with pytest.raises(MamctsError, match="inner node records"):
        s.merge_roots()
    s.close()
    # a compact layout that does not exist for the configuration is refused, not silently replaced:
    # a negative discount (the other agents' values are no longer known to be +0.0), trees of >= 32 767 nodes
    with pytest.raises(MamctsError, match="compact layouts"):
        Search(engine, ENV_CROSSING_I32, default_params(max_number_of_iterations=100, discount_factor=-0.5), 1,
               cp=default_crossing_params(num_other_agents=2), layout=LAYOUT_COMPACT)
    with pytest.raises(MamctsError, match="compact layouts"):
        Search(engine, ENV_CROSSING_I32, default_params(max_number_of_iterations=40000, max_number_of_nodes=100000000), 1,
               cp=default_crossing_params(num_other_agents=2), layout=LAYOUT_COMPACT)


@pytest.mark.parametrize("layout", [LAYOUT_CLASSIC, LAYOUT_COMPACT])
def test_search_reset_and_limits(engine, layout):
    mp = default_params(max_number_of_iterations=100, rh_max_number_of_iterations=50)
    cp = default_crossing_params(num_other_agents=1)
    s = Search(engine, ENV_CROSSING_I32, mp, 33, cp=cp, action_source=SRC_PHILOX, philox_seed=3, layout=layout)
    s.run(100)
    a = s.dump_tree(5, 101)
    with pytest.raises(MamctsError, match="MAX_NUMBER_OF_ITERATIONS"):
        s.run(1)
    s.reset([5, 5])
    assert s.counters()["iterations"] == 0
    s.run(100)
    assert np.array_equal(s.dump_tree(5, 101), a)   # idempotent after reset
    s.reset([11, 9])                                # a different decision state
    s.run(50)
    assert s.root_stats(0, 0)["visits"] == 50
    s.close()
    with pytest.raises(MamctsError, match="compiled for"):
        Search(engine, ENV_CROSSING_I32, mp, 1, cp=default_crossing_params(num_other_agents=9))
    with pytest.raises(MamctsError, match="compiled for"):
        Search(engine, ENV_CROSSING_I32, mp, 1, cp=default_crossing_params(num_other_agents=1, num_other_actions=9))


def test_root_policy_statistically_matches_cpu_twin(engine):
    """Root action-selection frequencies over many independent trees: GPU trees [0, T) against CPU
    oracle trees [T, 2T) (disjoint Philox streams, same distribution). Stated tolerance: total
    variation distance of the pooled root visit distribution <= 0.02, best action equal."""
    iters, T = 400, 256
    mp = default_params(max_number_of_iterations=iters, max_number_of_nodes=100000000,
                        rh_max_number_of_iterations=50)
    cp = default_crossing_params(num_other_agents=1)
    s = Search(engine, ENV_CROSSING_I32, mp, T, cp=cp, action_source=SRC_PHILOX, philox_seed=1000)
    s.run(iters)
    m = s.merge_roots()
    gpu_freq = m["n"].astype(np.float64) / m["n"].sum()
    cpu_n = np.zeros(4)
    cpu_q = np.zeros(4)
    env = O.make_env(cp=cp)
    for tree in range(T, 2 * T):
        t = O.OracleTree(env, mp, [5, 5], action_source=SRC_PHILOX, seed=1000, tree_id=tree)
        t.run(iters)
        st = t.root_stat(0)
        cpu_n += np.maximum(st["n"], 0)
        cpu_q += st["q"]
        t.close()
    cpu_freq = cpu_n / cpu_n.sum()
    tv = 0.5 * np.abs(gpu_freq - cpu_freq).sum()
    assert tv <= 0.02, (tv, gpu_freq, cpu_freq)
    assert int(np.argmax(cpu_q)) == m["best"]
    s.close()


@pytest.mark.parametrize("lpw,helpers", [(2, "1"), (2, "0"), (4, "1"), (1, "0"), (8, "1"), (32, "1")])
@pytest.mark.parametrize("layout", [LAYOUT_CLASSIC, LAYOUT_COMPACT])
def test_trees_per_warp_do_not_change_results(engine, lpw, helpers, layout, monkeypatch):
    """mamcts_search_create packs 1..32 trees into a warp depending on the tree count (few trees: one per
    warp); the small tests above all run one tree per warp, the bench 32. Forced here through the
    profiling knob: every packing must build the same trees - with at most 4 trees per warp the compact layout
    gives a tree a group of 8 lanes (Philox blocks of the rollout side by side, one lane per level of the
    backup; MAMCTS_SEARCH_HELPERS=0 switches that off), which must not change a bit either."""
    monkeypatch.setenv("MAMCTS_SEARCH_LPW", str(lpw))
    monkeypatch.setenv("MAMCTS_SEARCH_HELPERS", helpers)
    iters, T = 500, 100
    mp = default_params(max_number_of_iterations=iters, max_number_of_nodes=100000000,
                        rh_max_number_of_iterations=50)
    cp = default_crossing_params(num_other_agents=1)
    s = Search(engine, ENV_CROSSING_I32, mp, T, cp=cp, action_source=SRC_PHILOX, philox_seed=77, layout=layout)
    s.run(iters)
    env = O.make_env(cp=cp)
    for tree in (0, 1, lpw - 1, lpw % T, 63, 99):
        t = O.OracleTree(env, mp, [5, 5], action_source=SRC_PHILOX, seed=77, tree_id=tree)
        t.run(iters)
        assert np.array_equal(s.dump_tree(tree, t.num_nodes), t.dump(7)), f"lpw {lpw}: tree {tree} differs"
        t.close()
    assert s.counters()["iterations"] == T * iters
    s.close()


@pytest.mark.parametrize("spl,spw", [(2, 0), (2, 40), (4, 0)])
@pytest.mark.parametrize("source", ["philox", "reference_const"])
def test_async_tree_pool_scheduler_builds_the_same_trees(engine, spl, spw, source, monkeypatch):
    """search_async.cuh (opt-in, MAMCTS_SEARCH_ASYNC = tree slots per lane): trees are not bound to lanes, a warp
    owns a pool of 32 * spl trees and runs one kind of work item (one descent level / one Philox block of the
    rollout / one backup) per pass for up to 32 of them. Per tree the order of operations is unchanged, so
    every tree must equal the oracle's (and the lane-per-tree kernel's) bit for bit - also across several
    launches and with a partly filled last pool."""
    monkeypatch.setenv("MAMCTS_SEARCH_ASYNC", str(spl))
    if spw:
        monkeypatch.setenv("MAMCTS_SEARCH_ASYNC_SPW", str(spw))
    iters, T = 600, 333
    src = SRC_PHILOX if source == "philox" else SRC_REFERENCE_CONST
    mp = default_params(max_number_of_iterations=iters, max_number_of_nodes=100000000,
                        rh_max_number_of_iterations=50)
    cp = default_crossing_params(num_other_agents=1)
    s = Search(engine, ENV_CROSSING_I32, mp, T, cp=cp, action_source=src, philox_seed=31, layout=LAYOUT_COMPACT)
    s.run(250); s.run(1); s.run(iters - 251)
    env = O.make_env(cp=cp)
    for tree in (0, 1, 31, 32, 127, 128, 300, T - 1):
        t = O.OracleTree(env, mp, [5, 5], action_source=src, seed=31, tree_id=tree)
        t.run(iters)
        assert np.array_equal(s.dump_tree(tree, t.num_nodes), t.dump(7)), f"spl {spl}: tree {tree} differs"
        t.close()
    c = s.counters()
    assert c["iterations"] == T * iters
    m = s.merge_roots()
    assert int(m["n"].sum()) == T * iters
    s.close()
    # the counters of the two schedulers agree (same trees, same rollouts)
    monkeypatch.delenv("MAMCTS_SEARCH_ASYNC")
    s2 = Search(engine, ENV_CROSSING_I32, mp, T, cp=cp, action_source=src, philox_seed=31, layout=LAYOUT_COMPACT)
    s2.run(iters)
    c2 = s2.counters()
    assert c2 == c, (c, c2)
    s2.close()


def test_full_wave_sampled_trees_equal_oracle(engine):
    """As many trees as the bench runs per wave class (automatic packing: 32 trees per warp): sampled
    trees against the oracle, and the sum of the root visit counts against iterations x trees."""
    iters, T = 150, 40000
    mp = default_params(max_number_of_iterations=iters, max_number_of_nodes=100000000,
                        rh_max_number_of_iterations=50)
    cp = default_crossing_params(num_other_agents=1)
    s = Search(engine, ENV_CROSSING_I32, mp, T, cp=cp, action_source=SRC_PHILOX, philox_seed=9, layout=LAYOUT_COMPACT)
    s.run(iters)
    env = O.make_env(cp=cp)
    for tree in (0, 31, 32, 20011, T - 1):
        t = O.OracleTree(env, mp, [5, 5], action_source=SRC_PHILOX, seed=9, tree_id=tree)
        t.run(iters)
        assert np.array_equal(s.dump_tree(tree, t.num_nodes), t.dump(7)), f"tree {tree} differs"
        t.close()
    m = s.merge_roots()
    assert int(m["n"].sum()) == T * iters   # every iteration of every tree backed up through its root
    assert s.counters()["iterations"] == T * iters
    s.close()


@pytest.mark.parametrize("K", [2, 4, 16])
@pytest.mark.parametrize("layout", [LAYOUT_CLASSIC, LAYOUT_COMPACT])
def test_k_rollouts_per_leaf_equal_oracle_tree(engine, K, layout):
    """Leaf-batched rollouts: the value of a new leaf is the mean of K Philox rollouts (stream = creation
    index * K + k) summed in the fixed pairwise order of DESIGN.md section 6; the oracle builds the same tree."""
    iters = 300
    mp = default_params(max_number_of_iterations=iters, max_number_of_nodes=100000000,
                        rh_max_number_of_iterations=50)
    cp = default_crossing_params(num_other_agents=1)
    s = Search(engine, ENV_CROSSING_I32, mp, 40, cp=cp, action_source=SRC_PHILOX, philox_seed=21, layout=layout,
               rollouts_per_leaf=K)
    s.run(iters)
    env = O.make_env(cp=cp)
    steps = 0
    for tree in (0, 7, 39):
        t = O.OracleTree(env, mp, [5, 5], action_source=SRC_PHILOX, K=K, seed=21, tree_id=tree)
        t.run(iters)
        assert np.array_equal(s.dump_tree(tree, t.num_nodes), t.dump(7)), f"K {K}: tree {tree} differs"
        t.close()
    s.close()


def test_k_rollouts_per_leaf_validation(engine):
    mp = default_params(max_number_of_iterations=10)
    cp = default_crossing_params(num_other_agents=1)
    for bad in (0, 3, 32):
        with pytest.raises(MamctsError):
            Search(engine, ENV_CROSSING_I32, mp, 1, cp=cp, action_source=SRC_PHILOX, rollouts_per_leaf=bad)
    with pytest.raises(MamctsError):   # the reference's own rollouts are all identical (F1)
        Search(engine, ENV_CROSSING_I32, mp, 1, cp=cp, action_source=SRC_REFERENCE_CONST, rollouts_per_leaf=4)
