"""The pybind11 module `mamcts_cuda` (mamcts_b200/python_bindings/module.cpp, built by `make -C oracle pybind` where
the reference checkout exists): it must carry everything the reference's own `mamcts` module defines
(python/bindings/module.cpp:12-15) - checked with the reference's own wrapper tests, pickling included
(python/bindings/test/wrapper_test.py:64-102, restated here because /root/reference does not travel) - plus the
CUDA-backed classes. No GPU is touched here: the CUDA classes are only looked up, the stock instantiation is run."""
from __future__ import annotations

import glob
import os
import pickle
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def load_module():
    if not glob.glob(os.path.join(REF_DIR, "mamcts_cuda*.so")):
        pytest.skip("oracle/_ref/mamcts_cuda*.so is not built (needs the reference checkout: python __graft_entry__.py)")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import mamcts_cuda
    return mamcts_cuda


def test_module_is_a_superset_of_the_reference_module():
    m = load_module()
    # what python/bindings/define_mamcts.cpp and define_crossing_state.hpp define
    for name in ("MctsParameters", "MctsCrossingStateIntUctUct", "HypothesisBeliefTracker", "Viewer",
                 "CrossingStateParametersInt", "CrossingStateParametersFloat", "CrossingStateInt", "CrossingStateFloat",
                 "AgentPolicyCrossingStateInt", "CrossingStateEpisodeRunnerInt", "CrossingStateDefaultParametersInt",
                 "CrossingStateDefaultParametersFloat"):
        assert hasattr(m, name), name
    # what this repository adds next to MctsCrossingStateIntUctUct
    for name in ("MctsCrossingStateIntCudaUctCudaUct", "MctsCrossingStateIntUctUctCuda", "MctsCrossingStateIntUctUctRandom",
                 "DeviceRootParallelMctsCrossingStateInt", "set_crossing_state_parameters", "set_cuda_heuristic_options",
                 "MctsDefaultParameters"):
        assert hasattr(m, name), name


def test_reference_wrapper_tests_pickling():
    """python/bindings/test/wrapper_test.py: parameter structs survive a pickle round trip."""
    m = load_module()
    p = m.MctsDefaultParameters()
    p.uct_statistic.EXPLORATION_CONSTANT = 1.25
    p.random_heuristic.MAX_NUMBER_OF_ITERATIONS = 77
    q = pickle.loads(pickle.dumps(p))
    assert (q.DISCOUNT_FACTOR, q.RANDOM_SEED, q.MAX_NUMBER_OF_ITERATIONS) == (p.DISCOUNT_FACTOR, p.RANDOM_SEED, p.MAX_NUMBER_OF_ITERATIONS)
    assert q.uct_statistic.EXPLORATION_CONSTANT == 1.25 and q.random_heuristic.MAX_NUMBER_OF_ITERATIONS == 77
    cp = m.CrossingStateDefaultParametersInt()
    cp.NUM_OTHER_AGENTS, cp.CHAIN_LENGTH = 3, 41
    cq = pickle.loads(pickle.dumps(cp))
    assert (cq.NUM_OTHER_AGENTS, cq.CHAIN_LENGTH, cq.EGO_GOAL_POS, cq.REWARD_COLLISION) == (3, 41, cp.EGO_GOAL_POS, cp.REWARD_COLLISION)


def test_stock_search_through_the_module():
    m = load_module()
    p = m.MctsDefaultParameters()
    p.MAX_NUMBER_OF_ITERATIONS, p.MAX_SEARCH_TIME = 400, 1000000000
    p.random_heuristic.MAX_SEARCH_TIME = 1000000000
    cp = m.CrossingStateDefaultParametersInt()
    cp.NUM_OTHER_AGENTS = 1
    s = m.CrossingStateInt({1: 0}, cp)
    a, b = m.MctsCrossingStateIntUctUctRandom(p), m.MctsCrossingStateIntUctUctRandom(p)
    a.search(s); b.search(s)
    assert a.numIterations() == 400 and sum(a.root_ego_counts().values()) == 400
    assert a.root_ego_counts() == b.root_ego_counts() and a.root_ego_policy() == b.root_ego_policy()
