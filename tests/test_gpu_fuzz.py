"""Differential fuzzing of the device-resident search against the CPU oracle: random parameter sets
(discount, exploration constant, progressive widening, bound estimation on/off with explicit bounds,
step reward, cost mode, chain geometry, depth / node / rollout caps, both tree layouts, both action
sources) - every statistic of every node of the sampled trees must be bit-identical."""
from __future__ import annotations

import numpy as np
import pytest

import oracle_api as O
from mamcts_b200 import default_crossing_params, default_params, default_simple_params
from mamcts_b200._abi import (ENV_CROSSING_I32, ENV_SIMPLE, LAYOUT_CLASSIC, LAYOUT_COMPACT, SRC_PHILOX,
                              SRC_REFERENCE_CONST)
from mamcts_b200.engine import Search

pytestmark = pytest.mark.gpu


def _random_case(rng):
    num_others = int(rng.choice([1, 1, 1, 2, 3]))
    chain, goal = [(21, 12), (41, 26), (15, 9)][int(rng.integers(0, 3))]
    iters = int(rng.integers(50, 900))
    mp = default_params(
        discount_factor=float(rng.choice([0.9, 0.95, 1.0, 0.5, -0.9])),
        random_seed=int(rng.integers(1, 5000)),
        max_number_of_iterations=iters,
        max_number_of_nodes=int(rng.choice([100000000, 100000000, 60, 300])),
        max_search_depth=int(rng.choice([1000, 1000, 3, 6])),
        use_bound_estimation=int(rng.integers(0, 2)),
        rh_max_number_of_iterations=int(rng.choice([50, 10, 1, 0])),
        uct_lower_bound=-1000.0, uct_upper_bound=100.0,
        uct_exploration_constant=float(rng.choice([0.7, 0.1, 2.0])),
        uct_pw_k=float(rng.choice([4.0, 1.0, 2.0])), uct_pw_alpha=float(rng.choice([0.25, 0.5, 0.1])))
    cp = default_crossing_params(num_other_agents=num_others, chain_length=chain, ego_goal_pos=goal,
                                 reward_step=float(rng.choice([0.0, -1.0, 0.125])),
                                 cost_only_collision=int(rng.integers(0, 2)),
                                 reward_goal_reached=float(rng.choice([100.0, 50.0])))
    root = [int(v) for v in rng.integers(0, goal, size=num_others + 1)]
    return mp, cp, root, iters


@pytest.mark.parametrize("seed", range(32))
def test_crossing_search_fuzz(engine, seed):
    rng = np.random.default_rng(9000 + seed)
    for _ in range(3):
        mp, cp, root, iters = _random_case(rng)
        A = cp.num_other_agents + 1
        source = SRC_PHILOX if rng.random() < 0.8 else SRC_REFERENCE_CONST
        layouts = [LAYOUT_CLASSIC, LAYOUT_COMPACT] if A == 2 else [LAYOUT_CLASSIC]
        pseed, first = int(rng.integers(0, 2 ** 40)), int(rng.integers(0, 10 ** 6))
        T = 40
        env = O.make_env(cp=cp)
        oracle_trees = {}
        for tree in (0, 17, 39):
            t = O.OracleTree(env, mp, root, action_source=source, seed=pseed, tree_id=first + tree)
            t.run(iters)
            oracle_trees[tree] = (t.dump(7), t.num_nodes, t.rollout_steps)
            t.close()
        for layout in layouts:
            s = Search(engine, ENV_CROSSING_I32, mp, T, cp=cp, root_x=root, action_source=source, philox_seed=pseed,
                       first_tree_id=first, layout=layout)
            cut = int(rng.integers(0, iters + 1))
            if cut:
                s.run(cut)
            if iters - cut:
                s.run(iters - cut)
            for tree, (dump, nodes, _steps) in oracle_trees.items():
                assert s.root_stats(tree, 0)["num_nodes"] == nodes, (seed, layout, "node count")
                got = s.dump_tree(tree, nodes)
                assert got.shape == dump.shape and np.array_equal(got, dump), (seed, layout, tree)
            s.close()


@pytest.mark.parametrize("seed", range(3))
def test_simple_search_fuzz(engine, seed):
    rng = np.random.default_rng(7000 + seed)
    for _ in range(3):
        iters = int(rng.integers(20, 500))
        mp = default_params(discount_factor=float(rng.choice([0.9, 1.0])), random_seed=int(rng.integers(1, 999)),
                            max_number_of_iterations=iters, max_number_of_nodes=int(rng.choice([100, 100000])),
                            max_search_depth=int(rng.choice([1000, 4])),
                            rh_max_number_of_iterations=int(rng.choice([40, 5])),
                            use_bound_estimation=int(rng.integers(0, 2)))
        sp = default_simple_params()
        start = int(rng.integers(0, 9))
        source = SRC_PHILOX if rng.random() < 0.7 else SRC_REFERENCE_CONST
        pseed = int(rng.integers(0, 2 ** 40))
        for layout in (LAYOUT_CLASSIC, LAYOUT_COMPACT):
            s = Search(engine, ENV_SIMPLE, mp, 9, sp=sp, root_x=[start], action_source=source, philox_seed=pseed,
                       layout=layout)
            s.run(iters)
            for tree in (0, 8):
                t = O.OracleTree(O.make_env(sp=sp), mp, [start], action_source=source, seed=pseed, tree_id=tree)
                t.run(iters)
                got = s.dump_tree(tree, t.num_nodes)
                assert np.array_equal(got, t.dump(2)), (seed, layout, tree)
                t.close()
            s.close()
