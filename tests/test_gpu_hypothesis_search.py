"""GPU parity of the device-resident hypothesis-MCTS (BASELINE configs[3], mamcts_hyp_search_*): every node statistic
of a device tree - the ego's UctStatistic, every other agent's HypothesisStatistic per hypothesis and action, even the
iteration order of the reference's action maps - equals the tree the REFERENCE'S OWN
Mcts<CrossingState<int>, UctStatistic, HypothesisStatistic, RandomHeuristic>::search(state, belief_tracker) builds
(oracle/ref_driver.cc: ref_search_hypothesis), bit for bit: live against oracle/_ref where it is on the box, and
against the committed digests of tests/golden/hypothesis_search.json everywhere."""
from __future__ import annotations

import numpy as np
import pytest

import oracle_api as O
from helpers import golden, tree_digest
from mamcts_b200 import default_crossing_params, default_hypothesis_params, default_params
from mamcts_b200.engine import HypSearch, MamctsError, reference_hypothesis_sequence

pytestmark = pytest.mark.gpu

# the hypothesis sets of the reference's own tests: crossing_state_int_test.cc:105-106, :150, :190-192
CASES = [
    dict(name="int_test_belief", others=2, gaps=[(4, 5), (5, 6)], iters=1500, kw={}, hkw={}, skip=0),
    dict(name="int_test_true_hypothesis", others=2, gaps=[(5, 5)], iters=1200, kw={}, hkw={}, skip=3),
    dict(name="int_test_wrong_hypothesis", others=2, gaps=[(-2, 1), (2, 3), (4, 5)], iters=2000, kw={}, hkw={}, skip=0),
    dict(name="cost_based_selection", others=2, gaps=[(-2, 1), (2, 3), (4, 5)], iters=1500, kw={},
         hkw=dict(cost_based_action_selection=1), skip=7),
    dict(name="one_other_cost_only_collision", others=1, gaps=[(1, 3), (4, 6)], iters=1500, kw=dict(cost_only_collision=1),
         hkw=dict(pw_k=2.0, pw_alpha=0.3), skip=0),
    dict(name="three_others_chance_update", others=3, gaps=[(0, 2), (3, 5)], iters=800, kw={},
         hkw=dict(chance_constrained_update=1, cost_based_action_selection=1), skip=1),
    dict(name="node_budget", others=2, gaps=[(4, 5), (5, 6)], iters=900, kw={}, hkw={}, skip=0, max_nodes=300),
]


def _params(case):
    mp = default_params(max_number_of_iterations=case["iters"], max_number_of_nodes=case.get("max_nodes", 100000000),
                        rh_max_number_of_iterations=50)
    hp = default_hypothesis_params(**case["hkw"])
    cp = default_crossing_params(num_other_agents=case["others"], **case["kw"])
    return mp, hp, cp


def _device_tree(engine, case, seq, trees=3):
    mp, hp, cp = _params(case)
    s = HypSearch(engine, mp, hp, cp, trees, case["gaps"], seq)
    first = case["iters"] // 3
    s.run(first)                       # two launches: the trees carry on exactly where they stopped
    s.run(case["iters"] - first)
    return s


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_device_tree_equals_reference_golden(engine, case):
    g = {c["name"]: c for c in golden("hypothesis_search.json")}[case["name"]]
    seq = np.array(g["sequence"], dtype=np.uint8).reshape(case["iters"], case["others"])
    s = _device_tree(engine, case, seq)
    for tree in (0, 2):                # with decorrelate_trees = 0 every tree is the reference's tree
        d = s.dump_tree(tree, g["num_nodes"] + 1)
        assert d.shape[0] == g["num_nodes"]
        assert tree_digest(d) == g["tree_sha256"], f"{case['name']}: tree {tree} differs from the reference"
    m = s.merge_roots()
    root_visits = int(float.fromhex(g["root_record"][4]))   # the ego root's total_node_visits_ in the reference
    assert m["best"] == g["best"] and m["visits"] == 3 * root_visits
    c = s.counters()
    assert c["iterations"] == 3 * case["iters"] and c["nodes"] == 3 * g["num_nodes"]
    s.close()


@pytest.mark.skipif(not O.reference_available(), reason="compiled reference not on this box")
@pytest.mark.parametrize("case", CASES[:4], ids=[c["name"] for c in CASES[:4]])
def test_device_tree_equals_live_reference(engine, case):
    mp, hp, cp = _params(case)
    ref = O.ref_search_hypothesis(mp, hp, cp, case["gaps"], skip=case["skip"], tree_capacity=case["iters"] + 1)
    s = _device_tree(engine, case, ref["sequence"], trees=2)
    d = s.dump_tree(1, ref["num_nodes"] + 1)
    assert d.shape == ref["tree"].shape
    bad = np.nonzero(np.any(d.view(np.uint64) != ref["tree"].view(np.uint64), axis=1))[0]
    assert bad.size == 0, f"first differing node record {bad[:3]}:\n{d[bad[0]]}\n{ref['tree'][bad[0]]}"
    s.close()


def test_decorrelated_trees_and_second_decision(engine):
    """decorrelate_trees: differently seeded generators per tree (independent root-parallel trees, results
    independent of the sharding: keyed by the global tree id); a reset re-roots the same arenas."""
    case = CASES[2]
    mp, hp, cp = _params(case)
    seq = reference_hypothesis_sequence(1000, np.full((2, 3), 1.0 / 3), 0, case["iters"])
    a = HypSearch(engine, mp, hp, cp, 40, case["gaps"], seq, decorrelate_trees=True, first_tree_id=100)
    a.run(case["iters"])
    d0, d1 = a.dump_tree(0, case["iters"] + 1), a.dump_tree(1, case["iters"] + 1)
    assert d0.shape != d1.shape or not np.array_equal(d0, d1)
    b = HypSearch(engine, mp, hp, cp, 8, case["gaps"], seq, decorrelate_trees=True, first_tree_id=130)
    b.run(case["iters"])
    assert np.array_equal(b.dump_tree(3, case["iters"] + 1), a.dump_tree(33, case["iters"] + 1))
    m = a.merge_roots()
    assert m["visits"] == 40 * case["iters"]
    # second decision from another state with another sequence, then back: the first result again
    seq2 = reference_hypothesis_sequence(1000, np.full((2, 3), 1.0 / 3), case["iters"], case["iters"])
    a.reset([7, 4, 6], [1, 2, 1], seq2)
    a.run(case["iters"])
    assert not np.array_equal(a.dump_tree(0, case["iters"] + 1)[:5], d0[:5])
    a.reset([5, 5, 5], [2, 2, 2], seq)
    a.run(case["iters"])
    assert np.array_equal(a.dump_tree(0, case["iters"] + 1), d0)
    a.close(); b.close()


def test_unsupported_configurations_fail_loudly(engine):
    mp, hp, cp = _params(CASES[0])
    seq = np.zeros((CASES[0]["iters"], 2), dtype=np.uint8)
    with pytest.raises(MamctsError, match="PROGRESSIVE_WIDENING_HYPOTHESIS_BASED"):
        HypSearch(engine, mp, default_hypothesis_params(progressive_widening_hypothesis_based=0), cp, 2, [(4, 5)], seq)
    with pytest.raises(MamctsError, match="velocities"):
        HypSearch(engine, mp, hp, default_crossing_params(num_other_agents=2, min_velocity_other=-6, max_velocity_other=3), 2,
                  [(4, 5)], seq)
