"""GPU tests of the pybind11 module `mamcts_cuda`: the CUDA-backed classes bound next to the reference's
MctsCrossingStateIntUctUct give the stock reference's results from Python."""
from __future__ import annotations

import pytest

from test_pybind_cpu import load_module

pytestmark = pytest.mark.gpu


def _setup(m, iterations):
    p = m.MctsDefaultParameters()
    p.MAX_NUMBER_OF_ITERATIONS, p.MAX_SEARCH_TIME = iterations, 1000000000
    p.MAX_NUMBER_OF_NODES = 100000000
    p.random_heuristic.MAX_SEARCH_TIME = 1000000000
    p.random_heuristic.MAX_NUMBER_OF_ITERATIONS = 50
    cp = m.CrossingStateDefaultParametersInt()
    cp.NUM_OTHER_AGENTS = 1
    m.set_crossing_state_parameters(cp)
    return p, cp, m.CrossingStateInt({1: 0}, cp)


def test_cuda_mcts_classes_reproduce_the_stock_search_from_python():
    m = load_module()
    p, cp, state = _setup(m, 500)
    stock = m.MctsCrossingStateIntUctUctRandom(p)
    stock.search(state)
    m.set_cuda_heuristic_options("reference")
    for cls in (m.MctsCrossingStateIntUctUctCuda, m.MctsCrossingStateIntCudaUctCudaUct):
        cuda = cls(p)
        cuda.search(state)
        assert cuda.numIterations() == stock.numIterations() and cuda.numNodes() == stock.numNodes()
        assert cuda.root_ego_counts() == stock.root_ego_counts()
        assert cuda.root_ego_policy() == stock.root_ego_policy()          # bit-identical action values
        assert cuda.returnBestAction() == stock.returnBestAction()


def test_device_resident_search_from_python():
    m = load_module()
    p, cp, state = _setup(m, 800)
    stock = m.MctsCrossingStateIntUctUctRandom(p)
    stock.search(state)
    dev = m.DeviceRootParallelMctsCrossingStateInt(p, 6, source="reference", crossing_state_parameters=cp)
    dev.search(state)
    assert dev.numIterations() == 800 and dev.numNodes() == 6 * stock.numNodes()
    r = dev.root_statistic(3, 0)
    assert r["action_counts"] == stock.root_ego_counts() and r["action_values"] == stock.root_ego_policy()
    assert dev.returnBestAction() == stock.returnBestAction()
    merged = dev.merged_root_statistic()
    assert merged["total_node_visits"] == 6 * 800
    # uniform Philox rollouts: 2 000 trees, a decision from Python
    big = m.DeviceRootParallelMctsCrossingStateInt(p, 2000, philox_seed=5, source="philox")
    big.search(state)
    assert big.merged_root_statistic()["total_node_visits"] == 2000 * 800 and big.numEnvSteps() > 2000 * 800
