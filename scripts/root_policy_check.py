"""Root policy of the C2 search at full depth (10 000 iterations per tree): GPU trees against CPU twin
trees with disjoint Philox streams (same distribution), and against the reference's own deterministic
tree. Evidence for BASELINE.json's "root policies statistically indistinguishable"; the asserting version
with fewer iterations is tests/test_gpu_search.py::test_root_policy_statistically_matches_cpu_twin.
Not part of the product: it uses the oracle as the checker, like the tests do.
   python scripts/root_policy_check.py [gpu_trees] [cpu_trees]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_api as O
from mamcts_b200 import default_crossing_params, default_params
from mamcts_b200._abi import ENV_CROSSING_I32, SRC_PHILOX
from mamcts_b200.engine import Engine, Search

def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    TC = int(sys.argv[2]) if len(sys.argv) > 2 else 192
    iters = 10000
    mp = default_params(max_number_of_iterations=iters, max_number_of_nodes=100000000, rh_max_number_of_iterations=50)
    cp = default_crossing_params(num_other_agents=1)
    e = Engine(0)
    s = Search(e, ENV_CROSSING_I32, mp, T, cp=cp, action_source=SRC_PHILOX, philox_seed=1000)
    s.run(iters)
    n = np.zeros((T, 4)); q = np.zeros((T, 4))
    for t in range(T):
        st = s.root_stats(t, 0)
        n[t] = np.maximum(st["n"][:4], 0); q[t] = st["q"][:4]
    m = s.merge_roots()
    s.close(); e.close()
    env = O.make_env(cp=cp)
    cn = np.zeros((TC, 4)); cq = np.zeros((TC, 4))
    t0 = time.time()
    for i in range(TC):
        tr = O.OracleTree(env, mp, [5, 5], action_source=SRC_PHILOX, seed=1000, tree_id=1000000 + i)
        tr.run(iters)
        st = tr.root_stat(0)
        cn[i] = np.maximum(st["n"], 0); cq[i] = st["q"]
        tr.close()
    print(f"C2 root statistics of the ego agent, {iters} iterations per tree")
    print(f"GPU: {T} trees (Philox tree ids 0..{T-1}); CPU twin (oracle): {TC} trees (ids 1000000..), {time.time()-t0:.0f} s")
    gf, cf = n.sum(0) / n.sum(), cn.sum(0) / cn.sum()
    print("visit frequency  GPU", np.round(gf, 5), " CPU", np.round(cf, 5), f" total variation {0.5*np.abs(gf-cf).sum():.5f}")
    for a in range(4):
        se = np.sqrt(q[:, a].var(ddof=1) / T + cq[:, a].var(ddof=1) / TC)
        print(f"action {a}: mean Q GPU {q[:, a].mean():+.6f}  CPU {cq[:, a].mean():+.6f}  difference / standard error {(q[:, a].mean()-cq[:, a].mean())/se:+.2f}"
              f"   mean n GPU {n[:, a].mean():8.1f}  CPU {cn[:, a].mean():8.1f}  diff / s.e. "
              f"{(n[:, a].mean()-cn[:, a].mean())/np.sqrt(n[:, a].var(ddof=1)/T + cn[:, a].var(ddof=1)/TC):+.2f}")
    print("best action: GPU merged", m["best"], " CPU mean-Q argmax", int(np.argmax(cq.mean(0))),
          " per-tree best-action agreement GPU", np.round(np.bincount(q.argmax(1), minlength=4) / T, 4),
          " CPU", np.round(np.bincount(cq.argmax(1), minlength=4) / TC, 4))
main()
