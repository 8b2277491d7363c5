"""__graft_entry__.smoke(): one small invocation of the hot path on cuda:0, checked against the
CPU oracle (rollout kernel, backup via the device-resident search, root merge)."""
from __future__ import annotations

import numpy as np

import oracle_api as O
from helpers import assert_bit_equal
from mamcts_b200 import default_crossing_params, default_params
from mamcts_b200._abi import ENV_CROSSING_I32, SRC_PHILOX
from mamcts_b200.engine import Engine, Search


def run():
    e = Engine(0)
    mp = default_params(rh_max_number_of_iterations=50, max_number_of_iterations=200,
                        max_number_of_nodes=100000000)
    cp = default_crossing_params(num_other_agents=1)
    # rollouts: 2048 leaves x 1 rollout, Philox
    rng = np.random.default_rng(0)
    x = rng.integers(0, 12, size=(2, 2048)).astype(np.int32)
    kw = dict(kind=SRC_PHILOX, seed=1000, stream_hi_base=1)
    out = e.rollout_crossing(mp, cp, x, K=1, parity=True, **kw)
    exp = O.rollout_batch(O.make_env(cp=cp), mp, x[:, :256], K=1, **kw)
    assert out["steps"][:256].tolist() == exp["steps"].tolist()
    assert_bit_equal(out["ego_return"][:256], exp["ego_return"], "smoke rollout ego return")
    # search: 64 trees x 200 iterations (select/expand + rollout + backup on the device)
    s = Search(e, ENV_CROSSING_I32, mp, 64, cp=cp, action_source=SRC_PHILOX, philox_seed=1000)
    s.run(200)
    t = O.OracleTree(O.make_env(cp=cp), mp, [5, 5], action_source=SRC_PHILOX, seed=1000, tree_id=7)
    t.run(200)
    assert np.array_equal(s.dump_tree(7, 201), t.dump(7)), "smoke search tree differs from the oracle"
    m = s.merge_roots()
    assert m["visits"] == 64 * 200
    s.close()
    # 8 agents: the counts-only compact layout (search_compact_multi.cuh)
    cp8 = default_crossing_params(num_other_agents=7)
    s = Search(e, ENV_CROSSING_I32, mp, 32, cp=cp8, action_source=SRC_PHILOX, philox_seed=1000)
    s.run(200)
    t = O.OracleTree(O.make_env(cp=cp8), mp, [5] * 8, action_source=SRC_PHILOX, seed=1000, tree_id=3)
    t.run(200)
    assert np.array_equal(s.dump_tree(3, 201), t.dump(7)), "smoke 8-agent search tree differs from the oracle"
    s.close()
    # per-ego-action heuristic (ActionValueRandomHeuristic's compute part)
    cp3 = default_crossing_params(num_other_agents=2)
    av = e.action_values_crossing(mp, cp3, x[:, :64].repeat(2, axis=0)[:3], kind=SRC_PHILOX, seed=5, stream_hi_base=2)
    ex = O.action_values_batch(O.make_env(cp=cp3), mp, x[:, :64].repeat(2, axis=0)[:3], kind=SRC_PHILOX, seed=5, stream_hi_base=2)
    assert_bit_equal(av["action_return"], ex["action_return"], "smoke action values")
    # hypothesis-MCTS on the device against the digest of the reference's own search (tests/golden)
    from helpers import golden, tree_digest
    from mamcts_b200 import default_hypothesis_params
    from mamcts_b200.engine import HypSearch
    g = [c for c in golden("hypothesis_search.json") if c["name"] == "int_test_wrong_hypothesis"][0]
    mph = default_params(max_number_of_iterations=g["iters"], max_number_of_nodes=100000000, rh_max_number_of_iterations=50)
    hs = HypSearch(e, mph, default_hypothesis_params(), default_crossing_params(num_other_agents=g["others"]), 2,
                   [tuple(r) for r in g["gaps"]], np.array(g["sequence"], dtype=np.uint8).reshape(g["iters"], g["others"]))
    hs.run(g["iters"])
    assert tree_digest(hs.dump_tree(1, g["num_nodes"] + 1)) == g["tree_sha256"], "smoke hypothesis tree differs from the reference"
    hs.close()
    e.close()
