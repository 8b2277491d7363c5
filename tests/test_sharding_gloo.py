"""The N > 1 path on CPU: two processes over gloo (world size 2) shard root-parallel trees by global
id, reduce their trees' ego root statistics to the partial vector the GPUs all-reduce, all-reduce it,
and finish - the result must equal the single-process merge of all trees and the reference's
merge_node_statistics golden vectors. Trees are produced by the CPU oracle (test infrastructure)
with the same (seed, global tree id) Philox streams the device search uses, so the union of the
shards is independent of the world size."""
from __future__ import annotations

import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from helpers import golden, unhex  # noqa: E402
from mamcts_b200 import sharding  # noqa: E402

TOTAL_TREES = 7      # odd on purpose: ranks get 4 and 3
ITERATIONS = 300
SEED = 4242


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _root_stats(first: int, count: int):
    """Ego root statistics of trees [first, first+count) from the CPU oracle."""
    import oracle_api as O
    from mamcts_b200 import default_crossing_params, default_params
    from mamcts_b200._abi import SRC_PHILOX
    mp_ = default_params(max_number_of_iterations=ITERATIONS, max_number_of_nodes=100000000,
                         rh_max_number_of_iterations=50)
    cp = default_crossing_params(num_other_agents=1)
    nA = 4
    q = np.zeros((count, nA)); n = np.zeros((count, nA)); ex = np.zeros((count, nA), dtype=bool)
    vis = np.zeros(count, dtype=np.int64)
    for i in range(count):
        t = O.OracleTree(O.make_env(cp=cp), mp_, [5, 5], action_source=SRC_PHILOX, seed=SEED, tree_id=first + i)
        t.run(ITERATIONS)
        st = t.root_stat(0)
        for a in range(nA):
            ex[i, a] = st["n"][a] >= 0; q[i, a] = st["q"][a]; n[i, a] = max(st["n"][a], 0)
        vis[i] = st["visits"]
        t.close()
    return q, n, ex, vis


def _worker(rank: int, world: int, port: int, out_dir: str):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = sharding.strong_shard(TOTAL_TREES, world, rank)
    q, n, ex, vis = _root_stats(sh.first_tree_id, sh.num_trees)
    acc = torch.from_numpy(sharding.local_partial(q, n, ex, vis))
    dist.all_reduce(acc, op=dist.ReduceOp.SUM)          # the one collective of the path
    # bench.py's timing reduction: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res = sharding.finish(acc.numpy(), 4)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), q=res["q"], n=res["n"], occ=res["occurrences"],
             visits=res["visits"], value=res["value"], best=res["best"], tmax=t.item(),
             first=sh.first_tree_id, count=sh.num_trees)
    dist.barrier()
    dist.destroy_process_group()


def test_shards_cover_all_trees_disjointly():
    for total in (1, 7, 64, 65536):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                sh = sharding.strong_shard(total, world, r)
                seen += list(range(sh.first_tree_id, sh.first_tree_id + sh.num_trees))
            assert seen == list(range(total))
    assert sharding.weak_shard(32768, 8, 3) == sharding.Shard(3 * 32768, 32768)
    with pytest.raises(ValueError):
        sharding.strong_shard(8, 2, 2)


def test_finish_matches_reference_merge_golden():
    """sharding.local_partial + finish == UctStatistic::merge_node_statistics (golden from the reference)."""
    for case in golden("merge.json"):
        nA = case["num_actions"]
        cnt = np.array(case["cnt"])
        q = unhex(case["q"], cnt.shape)
        res = sharding.finish(sharding.local_partial(q, np.maximum(cnt, 0), cnt >= 0, case["visits"]), nA)
        mq = unhex(case["merged_q"])
        for a in range(nA):
            if case["merged_n"][a] >= 0:
                assert res["n"][a] == case["merged_n"][a]
                assert abs(res["q"][a] - mq[a]) <= 1e-12 * max(1.0, abs(mq[a]))
            else:
                assert res["occurrences"][a] == 0
        assert res["visits"] == case["merged_visits"]
        assert abs(res["value"] - float.fromhex(case["merged_value"])) <= 1e-12
        assert res["best"] == case["best"]


def test_world_size_2_gloo_equals_single_process(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    assert (int(r0["first"]), int(r0["count"]), int(r1["first"]), int(r1["count"])) == (0, 4, 4, 3)
    assert r0["tmax"] == 2.0 and r1["tmax"] == 2.0
    # every rank holds the same merged result
    for k in ("q", "n", "occ", "visits", "value", "best"):
        assert np.array_equal(r0[k], r1[k]), k
    # ... equal to one process merging all trees
    q, n, ex, vis = _root_stats(0, TOTAL_TREES)
    ref = sharding.finish(sharding.local_partial(q, n, ex, vis), 4)
    assert np.array_equal(r0["n"], ref["n"]) and np.array_equal(r0["occ"], ref["occurrences"])
    assert int(r0["visits"]) == ref["visits"] == TOTAL_TREES * ITERATIONS
    np.testing.assert_allclose(r0["q"], ref["q"], rtol=1e-12, atol=0)
    assert abs(float(r0["value"]) - ref["value"]) <= 1e-12 * abs(ref["value"])
    assert int(r0["best"]) == ref["best"]
    # ... and to the oracle's restatement of merge_node_statistics over all trees
    import oracle_api as O
    stats = []
    for t in range(TOTAL_TREES):
        s = O.OrcStat()
        s.num_actions = 4
        s.total_visits = int(vis[t])
        for a in range(4):
            if ex[t, a]:
                s.expanded[a] = 1; s.n[a] = int(n[t, a]); s.q[a] = q[t, a]; s.num_expanded += 1
        stats.append(s)
    arr = (C.POINTER(O.OrcStat) * len(stats))(*[C.pointer(s) for s in stats])
    merged = O.OrcStat()
    merged.num_actions = 4
    O.oracle().orc_stat_merge(arr, len(stats), C.byref(merged))
    for a in range(4):
        if merged.expanded[a]:
            assert merged.n[a] == int(r0["n"][a])
            assert abs(merged.q[a] - r0["q"][a]) <= 1e-12 * max(1.0, abs(merged.q[a]))
    assert O.oracle().orc_stat_best_action(C.byref(merged)) == int(r0["best"])


# ---------------------------------------------------------------------------------------------
# C3: the rollout index range sharded over ranks, one all-reduce of {sum, sum of squares, count}
# ---------------------------------------------------------------------------------------------
C3_TOTAL = 2003          # ragged on purpose
C3_OTHERS = 3


def _c3_returns(first, count):
    """Returns of rollouts [first, first + count) of one decision from the oracle (global stream ids)."""
    import oracle_api as O
    from mamcts_b200 import default_crossing_params, default_params
    from mamcts_b200._abi import SRC_PHILOX
    mpar = default_params(rh_max_number_of_iterations=50)
    cp = default_crossing_params(num_other_agents=C3_OTHERS)
    A = C3_OTHERS + 1
    x = np.full((A, count), 5, dtype=np.int32)
    out = O.rollout_batch(O.make_env(cp=cp), mpar, x, None, None, None, K=1, kind=SRC_PHILOX, seed=1000,
                          stream_hi_base=0, stream_lo_base=first)
    return np.vstack([out["ego_return"][None, :], out["other_return"].reshape(A - 1, count)])


def _c3_worker(rank, world, port, outdir):
    import torch.distributed as dist
    import torch
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", world_size=world, rank=rank)
    sh = sharding.rollout_shard(C3_TOTAL, world, rank)
    acc = torch.from_numpy(sharding.moments_partial(_c3_returns(sh.first_rollout, sh.num_rollouts)))
    dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    res = sharding.moments_finish(acc.numpy())
    np.savez(os.path.join(outdir, f"c3_rank{rank}.npz"), first=sh.first_rollout, count=sh.num_rollouts,
             total=res["count"], mean=res["mean"], variance=res["variance"])
    dist.destroy_process_group()


def test_c3_rollout_range_world_size_2_gloo_equals_single_process(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_c3_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "c3_rank0.npz")
    r1 = np.load(tmp_path / "c3_rank1.npz")
    assert (int(r0["first"]), int(r0["count"]), int(r1["first"]), int(r1["count"])) == (0, 1002, 1002, 1001)
    for k in ("total", "mean", "variance"):
        assert np.array_equal(r0[k], r1[k]), k
    one = sharding.moments_finish(sharding.moments_partial(_c3_returns(0, C3_TOTAL)))
    assert int(r0["total"]) == one["count"] == C3_TOTAL
    np.testing.assert_allclose(r0["mean"], one["mean"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(r0["variance"], one["variance"], rtol=1e-9, atol=1e-15)
    assert np.all(one["mean"][1:] == 0.0)      # reward-free other agents (crossing_state.h:145)


def test_rollout_shards_cover_the_range():
    for world in (1, 2, 3, 8):
        got = []
        for r in range(world):
            sh = sharding.rollout_shard(C3_TOTAL, world, r)
            got += list(range(sh.first_rollout, sh.first_rollout + sh.num_rollouts))
        assert got == list(range(C3_TOTAL))
