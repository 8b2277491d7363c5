#!/usr/bin/env python
"""bench.py - headline benchmark of the MA-MCTS rollout-and-backup hot path on B200.

Workload (BASELINE.json configs[1], "C2"): CrossingState<int>, 2 agents (ego 4 actions, other 7),
UctStatistic for ego and other, 10 000 MCTS iterations per tree, rollout depth cap 50, on T
root-parallel trees per GPU that live in HBM. One bench "step" = one decision: every tree is
re-rooted at the start state and searched for 10 000 iterations (select/expand -> rollout -> backup,
all on the device). Metric: env-steps/s = CrossingState::execute() calls (rollout + expansion) per
second over all GPUs; MCTS iterations/s is reported beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W]           # this repository's CUDA engine
  python bench.py --impl reference [...]                        # the reference's own CPU search

Prints ONE JSON line (rank 0). See DESIGN.md §8 for how every field is derived.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))  # oracle_api: ONLY for the cpu_baseline / reference legs

METRIC = "crossingstate_env_steps_per_s"
UNIT = "env-steps/s"
ITERATIONS = 10000        # C2: 10k iterations per tree
DEPTH_CAP = 50            # random_heuristic.MAX_NUMBER_OF_ITERATIONS
NUM_OTHERS = 1            # C2: 2 agents


def _params():
    from mamcts_b200 import default_crossing_params, default_params
    mp = default_params(max_number_of_iterations=ITERATIONS, max_number_of_nodes=100000000,
                        rh_max_number_of_iterations=DEPTH_CAP)
    cp = default_crossing_params(num_other_agents=NUM_OTHERS)
    return mp, cp


def _config(trees_per_gpu, n_gpus, action_source, layout="compact (16-B leaf + 256-B inner records)"):
    return {
        "workload": "C2: CrossingState<int> 2 agents (ego 4 / other 7 actions), UctStatistic ego+other, "
                    f"{ITERATIONS} iterations per tree, rollout depth cap {DEPTH_CAP}, "
                    f"{trees_per_gpu} root-parallel trees per GPU (one decision per step)",
        "trees_per_gpu": trees_per_gpu, "trees_total": trees_per_gpu * n_gpus,
        "iterations_per_tree": ITERATIONS, "rollout_depth_cap": DEPTH_CAP, "agents": NUM_OTHERS + 1,
        "action_source": action_source,
        "tree_layout": layout,
        "cache": "inputs larger than L2: tree arenas are tens of GB, no flush needed",
        "parallelism": f"root-parallel x{n_gpus} (trees sharded by global id, one NCCL all-reduce of "
                       "root statistics per decision)",
    }


def _ncu_entry(**match):
    """The committed ncu capture (profiles/r02_traffic.json) of exactly this configuration, else None."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        for e in json.load(f)["entries"]:
            if all(e.get(k) == v for k, v in match.items()):
                return e
    return None


def _measured_traffic(layout, trees, iterations):
    e = _ncu_entry(layout=layout if layout else 2, trees=trees, iterations=iterations)
    return e["dram_bytes_per_launch"] if e else None


def _issue_roofline(entry, kernel_ms, sm_mhz):
    """Issue-limited roofline: executed warp instructions per launch (ncu) / kernel time,
    against 148 SMs x 4 schedulers x 1 warp instruction per clock at the clock observed in this run."""
    if not entry or not sm_mhz:
        return None
    peak = 148 * 4 * sm_mhz * 1e6
    achieved = entry["warp_inst_per_launch"] / (kernel_ms * 1e-3)
    return {"bound": "issue", "achieved": achieved, "peak": peak, "unit": "warp-inst/s", "frac": achieved / peak,
            "warp_inst_per_launch": entry["warp_inst_per_launch"], "sm_mhz": sm_mhz, "source": entry["source"]}


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------

class ClockSampler:
    """nvidia-smi clock / throttle sampling DURING the timed region (B200_PROFILING.md)."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.samples = []
        self.proc = None
        self.index = index
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6:
                self.samples.append(parts)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [names[i] for i in range(4) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# the reference's own CPU path (cpu_baseline and --impl reference)
# ---------------------------------------------------------------------------------------------

def cpu_reference_search(threads: int, repeats: int):
    """`threads` independent Mcts<CrossingState<int>,UctStatistic,UctStatistic,RandomHeuristic>::search
    calls of the C2 configuration on the host cores (oracle/_ref = the unmodified reference)."""
    import oracle_api as O
    mp, cp = _params()
    if O.reference_available():
        calls, iters, _nodes = O.ref_search_count_steps(mp, cp)   # deterministic: exact step count
        it_per_s, seconds = O.ref_time_search(mp, cp, threads, repeats)
        steps_per_s = it_per_s * calls / iters
        return dict(kind="reference", env_steps_per_s=steps_per_s, iterations_per_s=it_per_s,
                    seconds=seconds, searches=threads * repeats,
                    env_steps_per_search=calls)
    # oracle port (scalar C restatement) when the compiled reference did not travel
    from mamcts_b200._abi import SRC_REFERENCE_CONST
    t0 = time.time()
    steps = 0
    for _ in range(repeats):
        t = O.OracleTree(O.make_env(cp=cp), mp, [5] * (NUM_OTHERS + 1), action_source=SRC_REFERENCE_CONST)
        t.run(ITERATIONS)
        steps += t.rollout_steps + t.num_nodes - 1
        t.close()
    el = time.time() - t0
    return dict(kind="port", env_steps_per_s=steps / el, iterations_per_s=repeats * ITERATIONS / el,
                seconds=el, searches=repeats, env_steps_per_search=steps / repeats)


def cpu_baselines(repeats: int):
    """Every CPU leg BASELINE.md promises, timed on this box's host cores BEFORE any GPU rank starts work
    (the other ranks of a torchrun launch wait for rank 0's flag file, see _cpu_gate), each a bounded sample:
    (i) `nproc` independent single_search of C2 (the embarrassingly parallel ceiling, the headline baseline),
    (ii) the reference's own root-parallel path Mcts::parallel_search with USE_MULTI_THREADING
    (mcts/mcts.h:188-221; the runnable stand-in for test/parallel_mcts, which needs OR-tools),
    (iii) RandomHeuristic::rollout loops at A = 2 and A = 8 (random_heuristic.h:27-104)."""
    import oracle_api as O
    from mamcts_b200 import default_crossing_params, default_params
    threads = os.cpu_count() or 1
    cb = cpu_reference_search(threads, max(1, repeats))
    out = {"value": cb["env_steps_per_s"], "unit": UNIT, "cores": threads, "kind": cb["kind"],
           "iterations_per_s": cb["iterations_per_s"],
           "sample": f"{cb['searches']} whole C2 searches ({ITERATIONS} iterations each, the reference's stock "
                     f"deterministic rollouts, {cb['env_steps_per_search']:.0f} execute() calls per search) on "
                     f"{threads} host threads, {cb['seconds']:.1f} s; timed before the GPU ranks start"}
    if O.reference_available():
        mp, cp = _params()
        it_s, sec, _q, n = O.ref_time_parallel_search(mp, cp, threads)
        out["parallel_search"] = {
            "iterations_per_s": it_s, "seconds": sec, "num_parallel_mcts": threads,
            "env_steps_per_s": it_s * cb["env_steps_per_search"] / ITERATIONS,
            "root_visits": int(sum(int(v) for v in n if v > 0)),
            "what": "Mcts::parallel_search, USE_MULTI_THREADING = true, NUM_PARALLEL_MCTS = nproc "
                    "(reference mcts/mcts.h:188-221), one call"}
        mpr = default_params(rh_max_number_of_iterations=DEPTH_CAP)
        out["c4_hypothesis"] = cpu_c4_hypothesis(threads)
        # C1 (BASELINE configs[0], the reference's own CPU-runnable case): SimpleState(4), 20 iterations, one thread
        mp1 = default_params(max_number_of_iterations=20, rh_max_number_of_iterations=1000)
        n1, t1 = 0, time.perf_counter()
        while time.perf_counter() - t1 < 1.0:
            O.ref_search_simple(mp1, 4)
            n1 += 1
        el1 = time.perf_counter() - t1
        out["c1_simple"] = {"iterations_per_s": 20 * n1 / el1, "searches": n1, "cores": 1,
                            "what": "Mcts<SimpleState, UctStatistic, UctStatistic, RandomHeuristic>::search, 20 iterations "
                                    "(test/uct/uct_test.cc:22-40), one host thread; every rollout runs into its 1 000-step cap "
                                    "(the stock policy repeats one joint action, which never ends a SimpleState "
                                    "episode: SURVEY.md F1), so an iteration is ~1 000 execute() calls"}
        # C3-shaped search: 8 agents, the C2 search parameters
        mp8, _ = _params()
        it8, sec8 = O.ref_time_search(mp8, default_crossing_params(num_other_agents=C3_OTHERS), threads, 1)
        out["search_a8"] = {"iterations_per_s": it8, "seconds": sec8, "cores": threads,
                            "what": f"{threads} independent searches of the C2 parameters with {C3_OTHERS + 1} agents "
                                    f"({ITERATIONS} iterations each), one per host thread"}
        for others in (1, C3_OTHERS):
            st, ro = O.ref_time_rollouts(mpr, default_crossing_params(num_other_agents=others), threads, 2.0)
            out[f"rollout_loop_a{others + 1}"] = {
                "env_steps_per_s": st, "rollouts_per_s": ro, "cores": threads,
                "what": f"RandomHeuristic::rollout from the start state, {others + 1} agents, depth cap {DEPTH_CAP}, "
                        "2 s on every host thread"}
    return out


def _cpu_gate_path():
    # all ranks of one torchrun launch share the launcher as parent process
    return os.path.join("/tmp", f"mamcts_bench_cpu_{os.getppid()}_{os.environ.get('MASTER_PORT', '0')}.json")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # each bench step = every host thread runs `rep` whole searches; sized to ~1 s per step
    rep = 1
    cpu_reference_search(threads, 1) if args.warmup else None
    results = [cpu_reference_search(threads, rep) for _ in range(args.steps)]
    steps_per_s = sum(r["env_steps_per_s"] * r["seconds"] for r in results) / sum(r["seconds"] for r in results)
    its_per_s = sum(r["iterations_per_s"] * r["seconds"] for r in results) / sum(r["seconds"] for r in results)
    ms = 1e3 * sum(r["seconds"] for r in results) / len(results)
    kind = results[0]["kind"]
    sample = (f"{threads} host threads x {rep} independent single_search of the C2 configuration per step "
              f"({ITERATIONS} iterations each, deterministic reference rollouts, "
              f"{results[0]['env_steps_per_search']:.0f} execute() calls per search)")
    cfg = _config(threads * rep, 1, "the reference's stock RandomHeuristic: every rollout step applies the one joint "
                  "action a fresh statistic draws first (mt19937 first draw, SURVEY.md F1) - the reference has no "
                  "other rollout policy; 5.7 execute() calls per iteration against 11.9 with uniform Philox actions, "
                  "so compare the arms in iterations_per_s",
                  "n/a for this arm (the reference keeps each tree in shared_ptr / std::map nodes on the host)")
    cfg["workload"] = ("C2: CrossingState<int> 2 agents (ego 4 / other 7 actions), UctStatistic ego+other, "
                       f"{ITERATIONS} iterations per tree, rollout depth cap {DEPTH_CAP}, {threads * rep} independent "
                       "trees per step on the host cores (one per thread)")
    cfg.pop("trees_per_gpu"); cfg.pop("trees_total")
    cfg["cache"] = "n/a (host arm: every tree is built in host memory)"
    cfg["trees_per_step"] = threads * rep
    cfg["parallelism"] = f"{threads} host threads, one single_search each (mcts/mcts.h:166-186)"
    line = {
        "impl": "reference", "metric": METRIC, "value": steps_per_s, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "iterations_per_s": its_per_s,
        "cpu_baseline": {"value": steps_per_s, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
                         "iterations_per_s": its_per_s},
        "e2e": {"value": steps_per_s, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# C3: the batched rollout kernel alone (BASELINE.json configs[2])
# ---------------------------------------------------------------------------------------------

C3_OTHERS = 7            # 8 agents
C3_ROLLOUTS = 1 << 20    # 1M rollouts per decision
C5_TREES = 65536         # BASELINE.json configs[4]: 65 536 trees in total over the GPUs
C5_ITERATIONS = 1000


def measure_rollout_c3(engine, stream, world=1, rank=0, steps=5, warmup=3, dist=None, e2e=True):
    """CrossingState<int>, 8 agents, 2^20 Philox rollouts per decision from the decision state, depth cap 50.
    The rollout index range is sharded over the ranks (sharding.rollout_shard; rollout r keeps its GLOBAL
    Philox stream id, so the set of rollouts does not depend on the number of GPUs). Each decision is ONE
    call, mamcts_rollout_crossing_i32 with out->moments: the rollout launch, the on-device moments reduction
    ({sum, sum^2} per agent + count over the per-rollout returns, reduce.cu), the K-rollout means, and one
    ncclAllReduce of the 2A+1 doubles. Device-resident inputs; env-steps/s from the kernel's own step counter,
    CUDA events on the engine stream, max over ranks, L2 flushed between decisions; plus the HBM roofline of
    rollout_kernel (algorithmic bytes 16A+8C+8 per rollout)."""
    import numpy as np
    import torch
    from mamcts_b200 import default_crossing_params, default_params, sharding
    from mamcts_b200._abi import SRC_PHILOX
    mp = default_params(rh_max_number_of_iterations=DEPTH_CAP)
    cp = default_crossing_params(num_other_agents=C3_OTHERS)
    A = C3_OTHERS + 1
    K = 256
    sh = sharding.rollout_shard(C3_ROLLOUTS, world, rank)
    if sh.num_rollouts % K or sh.first_rollout % K:
        raise SystemExit("bench.py: the C3 shard is not a whole number of leaves")
    n = sh.num_rollouts // K            # start states ("leaves") of this rank, K rollouts each
    dev = torch.device("cuda", engine.device)
    with torch.cuda.stream(stream):
        # the decision state (all agents at x = 5, reference crossing_state_common.h:31) in every slot
        x = torch.full((A, n), 5, dtype=torch.int32, device=dev)
        ego = torch.zeros(n, dtype=torch.float64, device=dev)
        oth = torch.zeros((A - 1, n), dtype=torch.float64, device=dev)
        tot = torch.zeros(1, dtype=torch.int64, device=dev)
        mom = torch.zeros(2 * A + 1, dtype=torch.float64, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > L2 (126 MB)

    def one(i):
        # rollout r = leaf * K + k of this rank draws from stream (i, first_rollout + r): global ids
        engine.rollout_crossing_device(mp, cp, n, K, x.data_ptr(), ego.data_ptr(), oth.data_ptr(),
                                       0, 0, tot.data_ptr(), kind=SRC_PHILOX, d_moments=mom.data_ptr(),
                                       seed=1000, stream_hi_base=i, stream_lo_base=sh.first_rollout)

    def barrier():
        if dist is not None and world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        one(i)
    barrier()
    times, total_steps = [], 0
    launches0 = engine.launch_count
    for i in range(steps):
        with torch.cuda.stream(stream):
            flush.fill_(i & 0xFF)                     # flush L2 between timed iterations
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        one(warmup + i)
        b.record(stream)
        stream.synchronize()
        times.append(a.elapsed_time(b))
        total_steps += int(tot.item())
    launches = engine.launch_count - launches0
    moments = mom.cpu().numpy()
    agg = torch.tensor([sum(times), float(total_steps)], dtype=torch.float64, device=dev)
    mx = agg.clone()
    if dist is not None and world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    total_ms, all_steps = float(mx[0]), float(agg[1])
    ms = total_ms / steps
    steps_per_s = all_steps / (total_ms * 1e-3)
    fin = sharding.moments_finish(moments)
    out = {
        "workload": f"C3: CrossingState<int> {A} agents, {C3_ROLLOUTS} Philox rollouts per decision from the decision "
                    f"state ({C3_ROLLOUTS // K} start-state slots x {K}) sharded over {world} GPU(s) by global rollout "
                    f"index, depth cap {DEPTH_CAP}; one call per decision: rollouts + device-side moments reduction + one "
                    "ncclAllReduce of 2A+1 doubles; L2 flushed between decisions",
        "env_steps_per_s": steps_per_s, "rollouts_per_s": C3_ROLLOUTS / (ms * 1e-3), "ms_per_decision": ms,
        "n_gpus": world, "scaling": "strong",
        "mean_steps_per_rollout": all_steps / (steps * C3_ROLLOUTS), "gpu_launches": int(launches),
        "rollouts_counted": int(fin["count"]), "mean_ego_return": float(fin["mean"][0]),
        "kernel_ms_estimate": ms,
    }
    if fin["count"] != C3_ROLLOUTS:
        raise SystemExit(f"bench.py: C3 all-reduced rollout count {fin['count']} != {C3_ROLLOUTS}")
    if e2e:
        # e2e: host buffers through the public call: start states in (numpy), per-slot means + moments out
        xh = np.full((A, n), 5, dtype=np.int32)
        engine.rollout_crossing(mp, cp, xh, K=K, kind=SRC_PHILOX, seed=1000, stream_hi_base=99, moments=True,
                                stream_lo_base=sh.first_rollout)
        barrier()
        t0 = time.perf_counter()
        e2e_steps = 0
        for i in range(steps):
            r = engine.rollout_crossing(mp, cp, xh, K=K, kind=SRC_PHILOX, seed=1000, stream_hi_base=100 + i,
                                        stream_lo_base=sh.first_rollout, moments=True)
            e2e_steps += r["total_steps"]
        barrier()
        e2e_s = time.perf_counter() - t0
        ag = torch.tensor([e2e_s, float(e2e_steps)], dtype=torch.float64, device=dev)
        mx2 = ag.clone()
        if dist is not None and world > 1:
            dist.all_reduce(mx2, op=dist.ReduceOp.MAX)
            dist.all_reduce(ag, op=dist.ReduceOp.SUM)
        out.update({"e2e_env_steps_per_s": float(ag[1]) / float(mx2[0]), "e2e_h2d_bytes": int(xh.nbytes),
                    "e2e_d2h_bytes": int(8 * n * (A + 2) + 8 + 8 * (2 * A + 1))})
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6650.0
    bytes_per_rollout = 16 * A + 8 * 1 + 8
    achieved = bytes_per_rollout * sh.num_rollouts / (ms * 1e-3) / 1e9
    out["roofline"] = {"bound": "hbm", "kernel": "rollout_kernel<CrossingEnv<7>,PHILOX>", "achieved": achieved,
                       "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                       "algorithmic_bytes_per_rollout": bytes_per_rollout,
                       "note": "per GPU, over the whole decision (rollouts + reductions); state lives in registers for "
                               "the whole episode: issue-bound, see profiles/ for smsp__issue_active and instructions "
                               "per env-step"}
    return out


# ---------------------------------------------------------------------------------------------
# this repository's engine
# ---------------------------------------------------------------------------------------------

def measure_c5_strong(engine, stream, world, rank, dist, steps=5, warmup=2):
    """BASELINE.json configs[4] (C5): 65 536 root-parallel C2 trees IN TOTAL, sharded over the ranks by global
    tree id (sharding.strong_shard - the union of the ranks' trees is the same for every N), 1 000 iterations
    per tree, one NCCL all-reduce of the ego root statistics per decision. Strong scaling: total work is fixed.
    A step = one decision through the public call (host root state in, merged statistics and best action out:
    Search.reset(host) + run + merge_roots), timed with CUDA events on the engine stream, max over ranks."""
    import numpy as np
    import torch
    from mamcts_b200 import default_crossing_params, default_params, sharding
    from mamcts_b200._abi import ENV_CROSSING_I32, SRC_PHILOX
    from mamcts_b200.engine import Search
    mp = default_params(max_number_of_iterations=C5_ITERATIONS, max_number_of_nodes=100000000,
                        rh_max_number_of_iterations=DEPTH_CAP)
    cp = default_crossing_params(num_other_agents=NUM_OTHERS)
    sh = sharding.strong_shard(C5_TREES, world, rank)
    s = Search(engine, ENV_CROSSING_I32, mp, sh.num_trees, cp=cp, action_source=SRC_PHILOX, philox_seed=1000,
               first_tree_id=sh.first_tree_id)
    root = np.full(NUM_OTHERS + 1, 5, dtype=np.int32)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    merged = None
    for _ in range(warmup):
        s.reset(root); s.run(C5_ITERATIONS); merged = s.merge_roots()
    barrier()
    launches0 = engine.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kern = []
    t0 = time.perf_counter()
    ev[0].record(stream)
    for _ in range(steps):
        ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.reset(root)
        ka.record(stream)
        s.run(C5_ITERATIONS)
        kb.record(stream)
        merged = s.merge_roots()          # local reduce + ncclAllReduce + device -> host read
        kern.append((ka, kb))
    ev[1].record(stream)
    barrier()
    wall_s = time.perf_counter() - t0
    launches = engine.launch_count - launches0
    c = s.counters()
    dev = torch.device("cuda", engine.device)
    mx = torch.tensor([ev[0].elapsed_time(ev[1]), sum(a.elapsed_time(b) for a, b in kern) / len(kern), wall_s * 1e3],
                      dtype=torch.float64, device=dev)
    sm = torch.tensor([c["rollout_steps"] + c["expand_steps"], c["iterations"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    total_ms, kern_ms, wall_ms = (float(v) for v in mx.tolist())
    env_steps, iters = (float(v) for v in sm.tolist())
    s.close()
    # every tree's root was visited once per iteration, on every rank, and the collective saw all of them
    if int(merged["visits"]) != C5_TREES * C5_ITERATIONS or int(sum(int(v) for v in merged["n"])) != C5_TREES * C5_ITERATIONS:
        raise SystemExit(f"bench.py: C5 merged root visits {merged['visits']} != {C5_TREES * C5_ITERATIONS}")
    return {
        "workload": f"C5: {C5_TREES} root-parallel C2 trees in total over {world} GPU(s) ({sh.num_trees} on rank 0), "
                    f"{C5_ITERATIONS} iterations per tree, one NCCL all-reduce of 13 doubles per decision; a step = one "
                    "decision through Search.reset(host root) + run + merge_roots",
        "scaling": "strong", "n_gpus": world, "trees_total": C5_TREES, "trees_per_gpu": sh.num_trees,
        "iterations_per_tree": C5_ITERATIONS, "steps": steps,
        "env_steps_per_s": env_steps * steps / (total_ms * 1e-3), "iterations_per_s": iters * steps / (total_ms * 1e-3),
        "ms_per_decision": total_ms / steps, "search_kernel_ms": kern_ms, "wall_ms_per_decision": wall_ms / steps,
        "merged_root_visits": int(merged["visits"]), "merged_root_n": [int(v) for v in merged["n"]],
        "best_action": int(merged["best"]), "gpu_launches": int(launches),
    }


def _c3_variants(engine, stream, steps=5):
    """The two other C3 inputs SURVEY.md 8(d) names, on ONE GPU (rank 0, called after the engine's communicator is
    gone: a call that asks for the moments all-reduces them when a communicator is attached, which a single rank must
    never do - so these calls do not ask for them either), device-resident, same kernel as the main section: (1) longer, branchier episodes - CHAIN_LENGTH = 41, EGO_GOAL_POS = 26, the reference's own 4-agent
    setting (environments/tests/crossing_state_int_test.cc:232-235) with C3's 8 agents; (2) a ragged batch as a
    leaf-parallel search hands it over: start states at depths 1..6 instead of the decision state - every slot's
    state is what 1..6 random joint actions lead to (produced by the same entry point: a rollout capped at that
    many steps returns its final state), terminal ones included (a leaf can be terminal: zero steps)."""
    import numpy as np
    import torch
    from mamcts_b200 import default_crossing_params, default_params
    from mamcts_b200._abi import SRC_PHILOX
    dev = torch.device("cuda", engine.device)
    mp = default_params(rh_max_number_of_iterations=DEPTH_CAP)
    A, K = C3_OTHERS + 1, 256
    n = C3_ROLLOUTS // K
    res = {}

    def timed(cp, x_t, flags_t, depth_t, tag):
        with torch.cuda.stream(stream):
            ego = torch.zeros(n, dtype=torch.float64, device=dev)
            oth = torch.zeros((A - 1, n), dtype=torch.float64, device=dev)
            tot = torch.zeros(1, dtype=torch.int64, device=dev)
        def one(i):
            engine.rollout_crossing_device(mp, cp, n, K, x_t.data_ptr(), ego.data_ptr(), oth.data_ptr(), 0, 0,
                                           tot.data_ptr(), d_flags=flags_t.data_ptr() if flags_t is not None else 0,
                                           d_depth=depth_t.data_ptr() if depth_t is not None else 0,
                                           kind=SRC_PHILOX, seed=1000, stream_hi_base=1000 + i, stream_lo_base=0)
        for i in range(2):
            one(i)
        torch.cuda.synchronize()
        total_ms, total_steps = 0.0, 0
        for i in range(steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            one(2 + i)
            b.record(stream)
            stream.synchronize()
            total_ms += a.elapsed_time(b)
            total_steps += int(tot.item())
        return {"what": tag, "env_steps_per_s": total_steps / (total_ms * 1e-3), "ms_per_decision": total_ms / steps,
                "mean_steps_per_rollout": total_steps / (steps * C3_ROLLOUTS)}

    # (1) longer episodes
    cp_long = default_crossing_params(num_other_agents=A - 1, chain_length=41, ego_goal_pos=26)
    with torch.cuda.stream(stream):
        x0 = torch.full((A, n), 5, dtype=torch.int32, device=dev)
    res["long_episodes"] = timed(cp_long, x0, None, None,
                                 f"CHAIN_LENGTH 41, EGO_GOAL_POS 26, {A} agents, {C3_ROLLOUTS} rollouts from the decision state")
    # (2) ragged leaf batch: states and depths as a search's leaves have them
    cp = default_crossing_params(num_other_agents=A - 1)
    rng = np.random.default_rng(1000)
    depth = rng.integers(1, 7, size=n).astype(np.uint32)
    xs = np.full((A, n), 5, dtype=np.int32)
    fl = np.zeros(n, dtype=np.uint8)
    for d in range(1, 7):
        sel = np.nonzero(depth == d)[0]
        if sel.size == 0:
            continue
        mpd = default_params(rh_max_number_of_iterations=d)
        r = engine.rollout_crossing(mpd, cp, np.full((A, sel.size), 5, dtype=np.int32), K=1, kind=SRC_PHILOX,
                                    parity=True, seed=4711, stream_hi_base=d)
        xs[:, sel] = r["final_x"]
        fl[sel] = r["final_flags"]
        depth[sel] = r["steps"]            # a walk that ended early sits at the depth it reached
    with torch.cuda.stream(stream):
        x_t = torch.from_numpy(xs).to(dev)
        f_t = torch.from_numpy(fl).to(dev)
        d_t = torch.from_numpy(depth.astype(np.int32)).to(dev)
    r2 = timed(cp, x_t, f_t, d_t, f"{n} leaf states at depths 1..6 (terminal ones included) x {K} rollouts each")
    r2["terminal_leaves"] = int((fl & 1).sum())
    res["leaf_states"] = r2
    return res


def measure_c4_hypothesis(engine, stream, trees=8192, iterations=10000, steps=3):
    """BASELINE.json configs[3] (C4): hypothesis-MCTS on CrossingState<int> - 2 other agents whose behaviour
    hypotheses are the three desired-gap ranges of the reference's own closed-loop test
    (environments/tests/crossing_state_int_test.cc:190-192), a hypothesis per agent and iteration drawn from uniform
    beliefs as HypothesisBeliefTracker draws them, UctStatistic for the ego, HypothesisStatistic for the others,
    `iterations` per tree, `trees` decorrelated root-parallel trees on one GPU (mamcts_hyp_search_*). With
    decorrelate_trees = 0 every tree is the reference's tree bit for bit (tests/test_gpu_hypothesis_search.py)."""
    import numpy as np
    import torch
    from mamcts_b200 import default_crossing_params, default_hypothesis_params, default_params
    from mamcts_b200.engine import HypSearch, reference_hypothesis_sequence
    gaps = [(-2, 1), (2, 3), (4, 5)]
    mp = default_params(max_number_of_iterations=iterations, max_number_of_nodes=100000000,
                        rh_max_number_of_iterations=DEPTH_CAP)
    hp = default_hypothesis_params()
    cp = default_crossing_params(num_other_agents=2)
    seq = reference_hypothesis_sequence(1000, np.full((2, len(gaps)), 1.0 / len(gaps)), 0, iterations)
    s = HypSearch(engine, mp, hp, cp, trees, gaps, seq, decorrelate_trees=True)
    s.run(iterations)                       # warm-up decision
    torch.cuda.synchronize()
    launches0 = engine.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(steps):
        s.reset([5, 5, 5], [2, 2, 2], seq)  # host -> device: root state and the iteration's hypothesis ids
        s.run(iterations)
        merged = s.merge_roots()
    ev1.record(stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = ev0.elapsed_time(ev1) / steps
    c = s.counters()
    launches = engine.launch_count - launches0
    s.close()
    if int(merged["visits"]) != trees * iterations:
        raise SystemExit(f"bench.py: C4 merged root visits {merged['visits']} != {trees * iterations}")
    return {
        "workload": f"C4: hypothesis-MCTS, CrossingState<int>, 2 other agents, hypotheses {gaps} (desired-gap ranges of "
                    f"crossing_state_int_test.cc:190-192), uniform beliefs, {iterations} iterations per tree, {trees} "
                    "decorrelated root-parallel trees on one GPU; a step = one decision through HypSearch.reset(host root, "
                    "host hypothesis sequence) + run + merge_roots",
        "trees": trees, "iterations_per_tree": iterations, "steps": steps,
        "iterations_per_s": c["iterations"] / (ms * 1e-3),
        "env_steps_per_s": (c["rollout_steps"] + c["expand_steps"]) / (ms * 1e-3),
        "ms_per_decision": ms, "wall_ms_per_decision": 1e3 * wall / steps, "nodes_per_tree": c["nodes"] / trees,
        "best_action": int(merged["best"]), "gpu_launches": int(launches),
    }


def cpu_c4_hypothesis(threads: int, iterations=10000):
    """The reference's own hypothesis-MCTS on the host cores: `threads` independent
    Mcts<CrossingState<int>, UctStatistic, HypothesisStatistic, RandomHeuristic>::search(state, belief_tracker)
    calls of the C4 configuration (oracle/_ref: ref_time_search_hypothesis), one std::thread each."""
    import oracle_api as O
    if not O.reference_available():
        return None
    from mamcts_b200 import default_crossing_params, default_hypothesis_params, default_params
    mp = default_params(max_number_of_iterations=iterations, max_number_of_nodes=100000000,
                        rh_max_number_of_iterations=DEPTH_CAP)
    it_s, sec = O.ref_time_search_hypothesis(mp, default_hypothesis_params(), default_crossing_params(num_other_agents=2),
                                             [(-2, 1), (2, 3), (4, 5)], threads)
    return {"iterations_per_s": it_s, "cores": threads, "seconds": sec,
            "what": f"{threads} independent hypothesis searches of the C4 configuration ({iterations} iterations each), "
                    "one per host thread (mcts/mcts.h:143-164)"}


def _check_sampled_tree(search, mp, cp, first_tree_id, tree):
    """Outside the timed region: one of the trees the bench just timed, node for node, against the CPU oracle
    driven by the same Philox stream (tests/oracle_api.py; the checker, never the thing measured)."""
    import numpy as np
    import oracle_api as O
    from mamcts_b200._abi import SRC_PHILOX
    t = O.OracleTree(O.make_env(cp=cp), mp, [5] * (NUM_OTHERS + 1), action_source=SRC_PHILOX, seed=1000,
                     tree_id=first_tree_id + tree)
    t.run(ITERATIONS)
    got = search.dump_tree(tree, t.num_nodes + 1)
    ok = got.shape[0] == t.num_nodes and np.array_equal(got, t.dump(7))
    nodes = t.num_nodes
    t.close()
    return bool(ok), int(nodes)


def run_native(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # ---- cpu baseline first: the reference's own search on the host cores while no GPU rank is working ----
    cpu_baseline = None
    gate = _cpu_gate_path()
    if rank == 0:
        cpu_baseline = cpu_baselines(args.cpu_repeats)
        if world > 1:
            with open(gate + ".tmp", "w") as f:
                json.dump({"done": True}, f)
            os.replace(gate + ".tmp", gate)
    else:
        t_wait = time.time()
        while not os.path.exists(gate) and time.time() - t_wait < 600:
            time.sleep(0.2)

    import numpy as np
    import torch
    import torch.distributed as dist

    from mamcts_b200._abi import ENV_CROSSING_I32, SRC_PHILOX, lib_path
    from mamcts_b200.engine import Engine, Search

    if not os.path.exists(lib_path()):
        import __graft_entry__
        __graft_entry__.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the mamcts_b200 engine has no CPU fallback")

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        if rank == 0:
            try:
                os.remove(gate)
            except OSError:
                pass
    engine = Engine(local_rank)
    stream = torch.cuda.Stream(device=local_rank)
    engine.set_stream(stream.cuda_stream)
    if world > 1:
        # the engine's own NCCL communicator; the unique id travels over torch.distributed
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.tensor(list(engine.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        engine.comm_init_rank(bytes(uid.cpu().tolist()), world, rank)

    from mamcts_b200 import sharding
    if args.workload == "c3":
        c3 = measure_rollout_c3(engine, stream, world, rank, args.steps, args.warmup, dist if world > 1 else None)
        if rank == 0:
            print(json.dumps({"metric": METRIC, "unit": UNIT, **c3}))
        engine.close()
        return
    if args.workload == "c5":
        c5 = measure_c5_strong(engine, stream, world, rank, dist, args.steps, args.warmup)
        if rank == 0:
            print(json.dumps({"metric": METRIC, "unit": UNIT, **c5}))
        engine.close()
        return
    mp, cp = _params()
    T = args.trees
    # the tree arenas must fit into what is free on this GPU (compact layout: inner + leaf records; classic:
    # 288-byte nodes); fewer trees are named in `config`, never silently
    free_b, _total_b = torch.cuda.mem_get_info(local_rank)
    if args.layout == 1:
        per_tree = (ITERATIONS + 1) * 288
    else:
        per_tree = (min(args.max_inner, ITERATIONS + 1) if args.max_inner else ITERATIONS + 1) * 256 + (ITERATIONS + 1) * 16
    fit = int((free_b - (2 << 30)) // (per_tree + 64))
    if fit < T:
        wave = 148 * 64
        T_fit = max(wave, (fit // wave) * wave) if fit >= wave else max(1, fit)
        if world > 1:   # every rank runs the same number of trees
            t = torch.tensor([T_fit], device=f"cuda:{local_rank}", dtype=torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            T_fit = int(t.item())
        sys.stderr.write(f"bench.py: {T} trees need {T * per_tree / 2**30:.0f} GiB, {free_b / 2**30:.0f} GiB are free: "
                         f"running {T_fit} trees per GPU\n")
        T = T_fit
    shard = sharding.weak_shard(T, world, rank)   # fixed trees per GPU, global ids rank*T ...
    search = Search(engine, ENV_CROSSING_I32, mp, shard.num_trees, cp=cp, action_source=SRC_PHILOX,
                    philox_seed=1000, first_tree_id=shard.first_tree_id, layout=args.layout,
                    max_inner_nodes_per_tree=args.max_inner)
    root = np.full(NUM_OTHERS + 1, 5, dtype=np.int32)
    root_pinned = torch.from_numpy(root.copy()).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        search.reset(None)          # root state already resident in HBM
        search.run(ITERATIONS)

    def e2e_step():
        # the call a user makes: host root state in, merged root statistics / best action out
        search.reset(root_pinned.numpy())     # host -> device copy of the decision state
        search.run(ITERATIONS)
        return search.merge_roots()           # local reduce (+ NCCL all-reduce) + device -> host read

    # ---- value: device-resident, CUDA events on the engine stream --------------------------------
    for _ in range(args.warmup):
        device_step()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = engine.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms = []
    ev0.record(stream)
    for _ in range(args.steps):
        ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        search.reset(None)
        ka.record(stream)
        search.run(ITERATIONS)
        kb.record(stream)
        kern_ms.append((ka, kb))
    ev1.record(stream)
    barrier()
    launches = engine.launch_count - launches0
    clock_info = clocks.stop() if rank == 0 else None
    elapsed_ms = ev0.elapsed_time(ev1)
    search_kernel_ms = sum(a.elapsed_time(b) for a, b in kern_ms) / len(kern_ms)
    c = search.counters()  # counters of the LAST step (reset zeroes them)
    t_ms = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([c["rollout_steps"] + c["expand_steps"], c["iterations"], c["backup_levels"],
                        c["rollout_steps"], c["nodes"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    elapsed_ms = float(t_ms.item())
    env_steps_per_step, iters_per_step, levels_per_step, rollout_steps, nodes = (float(v) for v in cnt.tolist())
    value = env_steps_per_step * args.steps / (elapsed_ms * 1e-3)
    its = iters_per_step * args.steps / (elapsed_ms * 1e-3)

    # ---- e2e: through the public API with host buffers, wall clock around the whole call -------
    for _ in range(min(args.warmup, 1)):  # e2e warm-up
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        merged = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    t_e2e = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = env_steps_per_step * args.steps / float(t_e2e.item())
    e2e_its = iters_per_step * args.steps / float(t_e2e.item())
    nA = 4

    # ---- correctness gate on the trees just timed (outside every timed region) -------------------
    # every iteration of every tree on every rank went through its root, and the all-reduce saw all of them;
    # one tree of rank 0 is compared node for node with the CPU oracle on the same Philox stream
    total_trees = T * world
    if int(merged["visits"]) != total_trees * ITERATIONS or sum(int(v) for v in merged["n"]) != total_trees * ITERATIONS:
        raise SystemExit(f"bench.py: merged root visits {merged['visits']} / counts {merged['n']} != trees x iterations "
                         f"= {total_trees * ITERATIONS}")
    gate_info = None
    if rank == 0:
        sample = (12345 * 7919) % T
        ok, nodes_sample = _check_sampled_tree(search, mp, cp, shard.first_tree_id, sample)
        if not ok:
            raise SystemExit(f"bench.py: tree {sample} of the timed search differs from the CPU oracle")
        gate_info = {"merged_root_visits": int(merged["visits"]), "expected": total_trees * ITERATIONS,
                     "sampled_tree": int(sample), "sampled_tree_nodes": nodes_sample,
                     "sampled_tree_equals_oracle": True,
                     "what": "sum of root visits over all ranks = trees x iterations; one timed tree dumped and "
                             "compared node for node (bit-exact) with the CPU oracle on the same Philox stream"}

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel (the search kernel), rank 0's launch -----------------
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            with open(peaks_path) as f:
                peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        A = NUM_OTHERS + 1
        it_local = c["iterations"]
        mean_depth = c["backup_levels"] / max(it_local, 1)
        # ALGORITHMIC bytes per iteration, SURVEY.md 8(d): rollout 16A + 8C + 8 and backup 80 * A * L
        bytes_per_it = (16 * A + 8 + 8) + 80 * A * mean_depth
        # the same plus the selection reads of the device-resident tree:
        # (12 nA + 12) per agent per level = 60 (ego) + 96 (other); most of these hit the L2
        bytes_per_it_sel = bytes_per_it + (60 + 96 * (A - 1)) * (mean_depth + 1)
        kern_s = search_kernel_ms * 1e-3
        achieved = bytes_per_it * it_local / kern_s / 1e9
        traffic = _measured_traffic(args.layout, T, ITERATIONS)
        roofline = {"bound": "hbm", "kernel": ("search_kernel" if args.layout == 1 else "search_compact_kernel") +
                                              "<CrossingEnv<1>,4,7,PHILOX>",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "peak_source": peak_src, "traffic": traffic,
                    "algorithmic_bytes_per_iteration": bytes_per_it, "mean_leaf_depth": mean_depth,
                    "dram_frac": (traffic / kern_s / 1e9 / peak) if traffic else None,
                    "frac_with_selection_reads": bytes_per_it_sel * it_local / kern_s / 1e9 / peak,
                    "algorithmic_bytes_per_iteration_with_selection_reads": bytes_per_it_sel,
                    "kernel_ms": search_kernel_ms,
                    "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum of the "
                                    "committed capture of this configuration, profiles/*_traffic.json; null when "
                                    "this configuration was not captured)",
                    "note": "frac = SURVEY.md 8(d) bytes (rollout in/out + backup read-modify-write) / kernel time / "
                            "measured HBM copy bandwidth; dram_frac = measured DRAM bytes of the capture / this run's "
                            "kernel time / the same peak. One lane per tree, one randomly placed 256-B record per "
                            "level: random-record access tops out near 2.8 TB/s of DRAM traffic "
                            "(profiles/r01_randrec_microbench.txt), not at the streaming peak; DESIGN.md section 7"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": _config(T, world, "philox4x32-10 uniform joint actions",
                              "classic (288-B node records)" if args.layout == 1 else
                              f"compact (16-B leaf + 256-B inner records, {args.max_inner} inner per tree)"),
            "iterations_per_s": its,
            "env_steps_per_decision": env_steps_per_step, "rollout_steps_per_decision": rollout_steps,
            "nodes_per_decision": nodes,
            "clocks": clock_info,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 64,  # the int32[16] root-state vector mamcts_search_reset copies
                    "d2h_bytes_per_step": 8 * (3 * nA + 3), "iterations_per_s": e2e_its,
                    "api": "Search.reset(host root) + Search.run + Search.merge_roots -> best action",
                    "best_action": int(merged["best"])},
            "gpu_launches": int(launches),
            "correctness_gate": gate_info,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
        }
        sm_mhz = (clock_info or {}).get("sm_mhz")
        line["roofline_issue"] = _issue_roofline(_ncu_entry(layout=args.layout if args.layout else 2, trees=T,
                                                            iterations=ITERATIONS), search_kernel_ms, sm_mhz)
    search.close()
    leaf_batched = None
    if rank == 0 and args.leaf_batch > 1 and world == 1:
        # the same search with K rollouts per new leaf (leaf-batched heuristic, BASELINE.json configs[1]):
        # a different algorithm setting than the reference's K = 1, so it is reported beside the headline,
        # never as the headline
        K = args.leaf_batch
        sk = Search(engine, ENV_CROSSING_I32, mp, T, cp=cp, action_source=SRC_PHILOX, philox_seed=1000,
                    layout=args.layout, max_inner_nodes_per_tree=args.max_inner, rollouts_per_leaf=K)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            sk.run(ITERATIONS)           # warm-up decision
            torch.cuda.synchronize()
            reps = 2
            ev0.record(stream)
            for _ in range(reps):
                sk.reset(None)
                sk.run(ITERATIONS)
            ev1.record(stream)
            torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / reps
        ck = sk.counters()
        leaf_batched = {"rollouts_per_leaf": K, "ms_per_decision": ms,
                        "env_steps_per_s": (ck["rollout_steps"] + ck["expand_steps"]) / (ms * 1e-3),
                        "iterations_per_s": ck["iterations"] / (ms * 1e-3),
                        "note": "leaf value = mean of K Philox rollouts in a fixed order (DESIGN.md section 6); "
                                "bit-equal to the oracle's K-rollout trees (tests/test_gpu_search.py)"}
        sk.close()
    # ---- the sections every N carries: C5 (strong scaling) and C3 (rollouts sharded), after the trees are freed
    c5 = measure_c5_strong(engine, stream, world, rank, dist) if not args.no_sections else None
    c3 = measure_rollout_c3(engine, stream, world, rank, dist=dist if world > 1 else None) if not args.no_sections else None
    if world > 1:
        engine.comm_destroy()   # every collective of the bench is done; what follows runs on rank 0 alone
    if rank == 0 and c3 is not None:
        # rank 0 alone: only ever after comm_destroy() above (tests/test_bench_structure_cpu.py)
        try:
            c3["variants"] = _c3_variants(engine, stream)
        except Exception as ex:   # a side section must never cost the line
            c3["variants"] = {"error": f"{type(ex).__name__}: {ex}"}
    c4 = None
    if rank == 0 and not args.no_sections:
        try:
            c4 = measure_c4_hypothesis(engine, stream)
        except SystemExit:
            raise                  # its correctness check failed: that must fail the run
        except Exception as ex:    # e.g. not enough free HBM for 8 192 hypothesis trees: a side section, not the line
            c4 = {"error": f"{type(ex).__name__}: {ex}"}
    if rank == 0:
        line["leaf_batched"] = leaf_batched
        line["c4_hypothesis"] = c4
        if c4 is not None and "error" not in c4 and cpu_baseline and cpu_baseline.get("c4_hypothesis"):
            c4["cpu_iterations_per_s"] = cpu_baseline["c4_hypothesis"]["iterations_per_s"]
            c4["cpu_cores"] = cpu_baseline["c4_hypothesis"]["cores"]
        line["c5_strong"] = c5
        if c3 is not None:
            c3["roofline_issue"] = _issue_roofline(_ncu_entry(kernel="rollout_kernel", workload="c3"),
                                                   c3["kernel_ms_estimate"], sm_mhz)
            if cpu_baseline and "rollout_loop_a8" in cpu_baseline:
                c3["cpu_rollout_loop_a8_env_steps_per_s"] = cpu_baseline["rollout_loop_a8"]["env_steps_per_s"]
        line["rollout_c3"] = c3
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    engine.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--trees", type=int, default=75776,
                    help="root-parallel trees per GPU: 75776 = 148 SMs x 8 resident blocks x 64 lanes, one full "
                         "wave of the search grid (compact layout: 75776 trees x (3200 x 256 B inner + "
                         "10001 x 16 B leaf records) = 74 GB of HBM)")
    ap.add_argument("--cpu-repeats", type=int, default=24, help="whole searches per host thread for cpu_baseline")
    ap.add_argument("--layout", type=int, default=2, help="tree layout: 0 auto, 1 classic, 2 compact")
    ap.add_argument("--max-inner", type=int, default=3200,
                    help="compact layout: inner records per tree (0 = one per node); a 10 000-iteration C2 "
                         "tree needs 2 500 +- 76 (max 2 734 over 120 trees); a tree that runs out fails the run")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c5"],
                    help="c2 = the headline search benchmark (default; carries the c5_strong and rollout_c3 sections); "
                         "c3 / c5 = only that section (profiling)")
    ap.add_argument("--no-sections", action="store_true", help="c2 only: skip the c5_strong and rollout_c3 sections")
    ap.add_argument("--leaf-batch", type=int, default=4,
                    help="also time the search with this many rollouts per new leaf (N = 1 only; 1 = skip)")
    ap.add_argument("--iterations", type=int, default=ITERATIONS,
                    help="iterations per tree (C2 = 10000; smaller values only for profiling runs)")
    args = ap.parse_args()
    globals()["ITERATIONS"] = args.iterations
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
