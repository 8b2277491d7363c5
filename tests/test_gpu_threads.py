"""SURVEY.md 8(b) "Threading": parallel_search calls single_search from NUM_PARALLEL_MCTS std::threads, each on
its own Mcts (mcts.h:198-210), so engine handles must be usable from several host threads at once - on one
shared engine (internal lock) and on one engine per thread. ctypes releases the GIL during a call, so these
threads really are concurrent inside the library."""
from __future__ import annotations

import threading

import numpy as np
import pytest

from mamcts_b200 import default_crossing_params, default_params
from mamcts_b200._abi import ENV_CROSSING_I32, SRC_PHILOX
from mamcts_b200.engine import Engine, Search

pytestmark = pytest.mark.gpu


def _inputs(seed, n=4096):
    rng = np.random.default_rng(seed)
    x = rng.integers(0, 12, size=(2, n)).astype(np.int32)
    depth = rng.integers(0, 5, size=n).astype(np.uint32)
    return x, depth


def _rollouts(engine, seed):
    mp = default_params(rh_max_number_of_iterations=50)
    cp = default_crossing_params(num_other_agents=1)
    x, depth = _inputs(seed)
    out = engine.rollout_crossing(mp, cp, x, None, None, depth, K=4, kind=SRC_PHILOX, seed=seed, stream_lo_base=7)
    return out["ego_return"].copy(), int(out["total_steps"])


def _search(engine, seed):
    mp = default_params(max_number_of_iterations=300, max_number_of_nodes=100000000, rh_max_number_of_iterations=50)
    cp = default_crossing_params(num_other_agents=1)
    s = Search(engine, ENV_CROSSING_I32, mp, 64, cp=cp, action_source=SRC_PHILOX, philox_seed=seed)
    s.run(300)
    m = s.merge_roots()
    s.close()
    return m["n"].copy(), m["q"].copy()


def _run_threads(jobs):
    results, errors = [None] * len(jobs), []

    def work(i, fn, args):
        try:
            for _ in range(3):
                results[i] = fn(*args)
        except Exception as exc:  # noqa: BLE001 - reported below
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(i, fn, args)) for i, (fn, args) in enumerate(jobs)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    return results


def test_one_engine_shared_by_eight_threads(engine):
    seeds = list(range(11, 19))
    expect_r = [_rollouts(engine, sd) for sd in seeds]
    expect_s = [_search(engine, sd) for sd in seeds[:4]]
    got = _run_threads([(_rollouts, (engine, sd)) for sd in seeds] + [(_search, (engine, sd)) for sd in seeds[:4]])
    for (e_ret, e_steps), (g_ret, g_steps) in zip(expect_r, got[:8]):
        assert g_steps == e_steps and np.array_equal(g_ret, e_ret)
    for (e_n, e_q), (g_n, g_q) in zip(expect_s, got[8:]):
        assert np.array_equal(g_n, e_n) and np.array_equal(g_q, e_q)


def test_one_engine_per_thread(engine):
    seeds = [31, 32, 33, 34]
    expect = [_rollouts(engine, sd) for sd in seeds]
    engines = [Engine(0) for _ in seeds]
    try:
        got = _run_threads([(_rollouts, (e, sd)) for e, sd in zip(engines, seeds)])
    finally:
        for e in engines:
            e.close()
    for (e_ret, e_steps), (g_ret, g_steps) in zip(expect, got):
        assert g_steps == e_steps and np.array_equal(g_ret, e_ret)
