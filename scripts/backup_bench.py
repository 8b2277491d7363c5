"""The batched backup / root-merge entry points at a size where the kernels' own time is visible (profiling helper;
ncu reads the kernel durations and DRAM bytes, this script only drives the calls through HOST buffers and prints the
algorithmic bytes of SURVEY.md 8(d) next to the wall time of a call).
   python scripts/backup_bench.py [paths] [levels]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mamcts_b200 import default_params
from mamcts_b200.engine import Engine

def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 75776
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    rng = np.random.default_rng(1)
    e = Engine(0)
    mp = default_params(discount_factor=0.9)
    # ---- mamcts_backup_uct: P paths of L edges, 2 agents, one chain of L + 1 nodes per path -----------------
    A, MA = 2, 7
    n_nodes = P * (L + 1)
    stats = dict(value=np.zeros(n_nodes * A), latest_return=np.zeros(n_nodes * A),
                 total_visits=np.zeros(n_nodes * A, dtype=np.uint32), q=np.zeros((n_nodes * A, MA)),
                 n=np.zeros((n_nodes * A, MA), dtype=np.uint32))
    leaf = np.arange(P, dtype=np.uint32) * (L + 1) + L
    has = np.ones(P, dtype=np.uint8)
    ego, oth = rng.normal(size=P), rng.normal(size=(A - 1, P))
    plen = np.full(P, L, dtype=np.uint32)
    off = (np.arange(P) * L).astype(np.uint32)
    enode = (np.arange(P)[:, None] * (L + 1) + (L - 1 - np.arange(L))[None, :]).astype(np.uint32).ravel()
    eact = rng.integers(0, MA, size=(P * L, A)).astype(np.uint8)
    erew = rng.choice([0.0, -1000.0, 100.0], size=(P * L, A))
    for rep in range(3):
        t0 = time.perf_counter()
        e.backup_uct(mp, A, leaf, has, ego, oth, off, plen, enode, eact, erew, stats)
        dt = time.perf_counter() - t0
    print(f"backup_uct: paths={P} levels={L} agents={A} algorithmic bytes 80*A*L per path = {80*A*L*P/1e6:.1f} MB, "
          f"call through host buffers {dt*1e3:.2f} ms", flush=True)
    # ---- mamcts_backup_hypothesis: P paths of L edges, 2 other agents, 3 hypotheses ---------------------------
    J, H, MAo, Cn = 2, 3, 8, 2
    hs = Engine.new_hyp_stats(n_nodes * J, H, MAo)
    eh, ea = rng.integers(0, H, (P * L, J)), rng.integers(0, MAo, (P * L, J))
    ec, lc = rng.choice([0.0, 1000.0, -1.5], (P * L, Cn)), rng.normal(size=(P, Cn))
    for rep in range(3):
        t0 = time.perf_counter()
        e.backup_hypothesis(mp, J, leaf, has, lc, off, plen, enode, eh, ea, ec, hs)
        dt = time.perf_counter() - t0
    print(f"backup_hypothesis: paths={P} levels={L} others={J} algorithmic bytes ((8C+2)+56) per level and agent = "
          f"{((8*Cn+2)+56)*J*L*P/1e6:.1f} MB, call through host buffers {dt*1e3:.2f} ms", flush=True)
    # ---- mamcts_root_merge: P roots, 4 ego actions ------------------------------------------------------------
    nA = 4
    rs = dict(value=np.zeros(P), latest_return=np.zeros(P), total_visits=rng.integers(0, 10000, P).astype(np.uint32),
              q=np.ascontiguousarray(rng.normal(size=(P, MA))), n=rng.integers(1, 4000, (P, MA)).astype(np.uint32))
    slots = np.arange(P, dtype=np.uint32)
    for rep in range(3):
        t0 = time.perf_counter()
        m = e.root_merge(rs, slots, nA)
        dt = time.perf_counter() - t0
    print(f"root_merge: roots={P} actions={nA} algorithmic bytes 12*nA + 4 per root = {(12*nA+4)*P/1e6:.2f} MB, "
          f"call through host buffers {dt*1e3:.2f} ms, best {m['best']}", flush=True)
    e.close()
main()
