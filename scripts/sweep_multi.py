"""Profiling helper: the device-resident search with more than two agents, classic vs compact layout.
python scripts/sweep_multi.py <others> <trees> <iterations> <layout 1|2> [max_inner]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from mamcts_b200 import default_crossing_params, default_params
from mamcts_b200._abi import ENV_CROSSING_I32, SRC_PHILOX
from mamcts_b200.engine import Engine, Search

others, trees, iters, layout = (int(v) for v in sys.argv[1:5])
max_inner = int(sys.argv[5]) if len(sys.argv) > 5 else 0
e = Engine(0)
mp = default_params(max_number_of_iterations=iters, max_number_of_nodes=100000000, rh_max_number_of_iterations=50)
cp = default_crossing_params(num_other_agents=others)
s = Search(e, ENV_CROSSING_I32, mp, trees, cp=cp, action_source=SRC_PHILOX, philox_seed=1000, layout=layout,
           max_inner_nodes_per_tree=max_inner)
s.run(iters); e.sync()
best = 1e9
for _ in range(2):
    s.reset(None); e.sync()
    t0 = time.perf_counter(); s.run(iters); e.sync(); best = min(best, time.perf_counter() - t0)
c = s.counters()
d = s.dump_tree(0, iters + 1)
inner = [int((s.dump_tree(t, iters + 1)[:, 2] > 0).sum()) for t in range(0, min(trees, 64), 7)]
print(f"others={others} trees={trees} iters={iters} layout={layout} max_inner={max_inner} ms={best*1e3:.1f} "
      f"it/s={c['iterations']/best:.4g} env_steps/s={(c['rollout_steps']+c['expand_steps'])/best:.4g} "
      f"mean_depth={c['backup_levels']/c['iterations']:.2f} nodes/tree={c['nodes']/trees:.0f} inner/tree max={max(inner)} mean={np.mean(inner):.0f} "
      f"free_GB={torch.cuda.mem_get_info(0)[0]/2**30:.0f}")
