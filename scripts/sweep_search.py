"""Sweep of the device-resident search over (lanes per warp, block threads, trees); prints one line per
point. Profiling helper, not part of the product or the tests.
   python scripts/sweep_search.py "lpw,threads,trees,iterations[,max_nodes,layout,inner,K];..." """
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mamcts_b200 import default_crossing_params, default_params
from mamcts_b200._abi import ENV_CROSSING_I32, SRC_PHILOX
from mamcts_b200.engine import Engine, Search

def main():
    pts = [tuple(int(x) for x in p.split(",")) for p in sys.argv[1].split(";") if p]
    others = int(os.environ.get("SWEEP_OTHERS", "1"))
    e = Engine(0)
    if os.environ.get("SWEEP_L2_GRAN"):   # cudaLimitMaxL2FetchGranularity = 5
        import ctypes
        rt = ctypes.CDLL("libcudart.so.12")
        rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(os.environ["SWEEP_L2_GRAN"])))
        v = ctypes.c_size_t(0); rt.cudaDeviceGetLimit(ctypes.byref(v), 5)
        print("L2 fetch granularity set rc", rc, "now", v.value, flush=True)
    for pt in pts:
        lpw, threads, trees, iters = pt[:4]
        maxn = pt[4] if len(pt) > 4 else 0
        layout = pt[5] if len(pt) > 5 else 0
        inner = pt[6] if len(pt) > 6 else 0
        K = pt[7] if len(pt) > 7 else 1
        os.environ["MAMCTS_SEARCH_LPW"] = str(lpw)
        os.environ["MAMCTS_SEARCH_THREADS"] = str(threads)
        mp = default_params(max_number_of_iterations=iters, max_number_of_nodes=100000000, rh_max_number_of_iterations=50)
        cp = default_crossing_params(num_other_agents=others)
        s = Search(e, ENV_CROSSING_I32, mp, trees, cp=cp, action_source=SRC_PHILOX, philox_seed=1000,
                   max_nodes_per_tree=maxn, layout=layout, max_inner_nodes_per_tree=inner, rollouts_per_leaf=K)
        s.run(iters); e.sync()
        best = 1e9
        for _ in range(2):
            s.reset(None); e.sync()
            t0 = time.perf_counter(); s.run(iters); e.sync(); dt = time.perf_counter() - t0
            best = min(best, dt)
        c = s.counters()
        steps = c["rollout_steps"] + c["expand_steps"]
        print(f"lpw={lpw} threads={threads} trees={trees} iters={iters} maxn={maxn} layout={layout} inner={inner} K={K} ms={best*1e3:.1f} "
              f"env_steps/s={steps/best:.4g} it/s={c['iterations']/best:.4g} "
              f"exact_ucb/it={s.exact_ucb_evaluations()/max(c['iterations'],1):.4f} levels/it={c['backup_levels']/max(c['iterations'],1):.3f}", flush=True)
        s.close()
    e.close()
main()
