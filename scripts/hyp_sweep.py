"""Hypothesis-MCTS search (C4 shape) over (trees, iterations): prints one line per point. Profiling helper.
   python scripts/hyp_sweep.py "trees,iterations;..." """
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mamcts_b200 import default_crossing_params, default_hypothesis_params, default_params
from mamcts_b200.engine import Engine, HypSearch, reference_hypothesis_sequence

def main():
    pts = [tuple(int(x) for x in p.split(",")) for p in sys.argv[1].split(";") if p]
    e = Engine(0)
    gaps = [(-2, 1), (2, 3), (4, 5)]
    for trees, iters in pts:
        mp = default_params(max_number_of_iterations=iters, max_number_of_nodes=100000000, rh_max_number_of_iterations=50)
        hp = default_hypothesis_params()
        cp = default_crossing_params(num_other_agents=2)
        seq = reference_hypothesis_sequence(1000, np.full((2, len(gaps)), 1.0 / len(gaps)), 0, iters)
        s = HypSearch(e, mp, hp, cp, trees, gaps, seq, decorrelate_trees=True)
        s.run(iters); e.sync()
        best = 1e9
        for _ in range(2):
            s.reset([5, 5, 5], [2, 2, 2], seq); e.sync()
            t0 = time.perf_counter(); s.run(iters); e.sync(); best = min(best, time.perf_counter() - t0)
        c = s.counters()
        m = s.merge_roots()
        print(f"hyp trees={trees} iters={iters} ms={best*1e3:.1f} it/s={c['iterations']/best:.4g} "
              f"env_steps/s={(c['rollout_steps']+c['expand_steps'])/best:.4g} nodes/tree={c['nodes']/trees:.0f} "
              f"levels/it={c['backup_levels']/max(c['iterations'],1):.2f} best={int(m['best'])} visits={int(m['visits'])}", flush=True)
        s.close()
    e.close()
main()
