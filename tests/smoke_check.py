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
    e.close()
