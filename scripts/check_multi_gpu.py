"""Multi-GPU check of the one collective of the path (run under torchrun, one rank per GPU):
every rank searches its shard of root-parallel trees, mamcts_search_merge_roots all-reduces the root
statistics over the engine's NCCL communicator, and the result must equal a single-GPU search over
ALL trees (same global tree ids => same trees)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from mamcts_b200 import default_crossing_params, default_params, sharding
from mamcts_b200._abi import ENV_CROSSING_I32, SRC_PHILOX
from mamcts_b200.engine import Engine, Search

def main():
    world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    e = Engine(local)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.tensor(list(e.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    e.comm_init_rank(bytes(uid.cpu().tolist()), world, rank)
    total, iters = 1000 * world + 37, 400
    mp = default_params(max_number_of_iterations=iters, max_number_of_nodes=100000000, rh_max_number_of_iterations=50)
    cp = default_crossing_params(num_other_agents=1)
    sh = sharding.strong_shard(total, world, rank)
    s = Search(e, ENV_CROSSING_I32, mp, sh.num_trees, cp=cp, action_source=SRC_PHILOX, philox_seed=77,
               first_tree_id=sh.first_tree_id)
    s.run(iters)
    m = s.merge_roots()           # local reduce + ncclAllReduce + finish, identical on every rank
    s.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, (m["n"].tolist(), m["q"].tolist(), m["visits"], m["best"]))
    ok = all(g == gathered[0] for g in gathered)
    # C3: one decision, the rollout index range sharded over the ranks (global Philox stream ids), one
    # all-reduce of {sum, sum of squares, count} per agent
    c3_total, c3_others = 100003, 7
    mp3 = default_params(rh_max_number_of_iterations=50)
    cp3 = default_crossing_params(num_other_agents=c3_others)

    def c3_returns(engine, first, count):
        x = np.full((c3_others + 1, count), 5, dtype=np.int32)
        out = engine.rollout_crossing(mp3, cp3, x, None, None, None, K=1, kind=SRC_PHILOX, seed=1000,
                                      stream_hi_base=0, stream_lo_base=first)
        return np.vstack([out["ego_return"][None, :], out["other_return"].reshape(c3_others, count)])

    rs = sharding.rollout_shard(c3_total, world, rank)
    c3 = sharding.moments_finish(e.allreduce_f64(sharding.moments_partial(c3_returns(e, rs.first_rollout, rs.num_rollouts))))
    e.comm_destroy()
    if rank == 0:
        e1 = Engine(local)
        s1 = Search(e1, ENV_CROSSING_I32, mp, total, cp=cp, action_source=SRC_PHILOX, philox_seed=77)
        s1.run(iters)
        r = s1.merge_roots()
        s1.close()
        ok = ok and r["n"].tolist() == m["n"].tolist() and r["visits"] == m["visits"] == total * iters
        ok = ok and r["best"] == m["best"] and np.allclose(r["q"], m["q"], rtol=1e-12, atol=0)
        one = sharding.moments_finish(sharding.moments_partial(c3_returns(e1, 0, c3_total)))
        c3_ok = (c3["count"] == one["count"] == c3_total and np.allclose(c3["mean"], one["mean"], rtol=1e-12, atol=1e-15)
                 and np.allclose(c3["variance"], one["variance"], rtol=1e-9, atol=1e-15))
        ok = ok and c3_ok
        print(f"MULTI_GPU_CHECK_C3 world={world} rollouts={c3_total} {'OK' if c3_ok else 'FAILED'} "
              f"mean ego return {c3['mean'][0]:.9g} (single GPU {one['mean'][0]:.9g})")
        e1.close()
        print(f"MULTI_GPU_CHECK world={world} trees={total} {'OK' if ok else 'FAILED'} n={m['n'].tolist()} q={m['q'].tolist()}")
    dist.barrier()
    dist.destroy_process_group()
    e.close()
    sys.exit(0 if ok else 1)
main()
