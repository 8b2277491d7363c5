"""GPU drop-in tests: the reference's OWN host code (Mcts<S,SE,SO,H>::search, StageNode,
parallel_search, UctTest::verify_uct - compiled unmodified from the reference checkout into
oracle/_ref/ by `make -C oracle dropin`) instantiated with this repository's CudaRandomHeuristic /
CudaUctStatistic / BatchedRootParallelMcts (mamcts_b200/host/*.h), which reach the kernels through
the C ABI. The binaries are prebuilt where the reference exists and travel to the GPU box; nothing
here reads /root/reference at run time.

  dropin_test        - tests/dropin/dropin_test.cc: stock vs CUDA instantiation, tree by tree, bit by bit
  dropin_verify_uct  - tests/dropin/dropin_verify_uct.cc: the reference's own test/uct/uct_test.cc:42-64
                       (verify_uct, test/uct/uct_test_class.h:23-215) with the CUDA heuristic plugged in
"""
from __future__ import annotations

import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def _run(name: str, timeout: int = 600) -> str:
    exe = os.path.join(REF_DIR, name)
    if not os.path.exists(exe):
        pytest.fail(f"{exe} is missing: run `python __graft_entry__.py` where the reference checkout exists "
                    f"(the binary travels to the GPU box with the snapshot)")
    pr = subprocess.run([exe], capture_output=True, text=True, timeout=timeout)
    out = pr.stdout + pr.stderr
    assert pr.returncode == 0, f"{name} exited {pr.returncode}:\n{out[-4000:]}"
    return out


def test_reference_mcts_with_cuda_heuristic_and_statistic_is_bit_equal():
    out = _run("dropin_test")
    assert "DROPIN OK" in out
    assert "FAIL" not in out
    # every named check ran
    for name in ("crossing2_cuda_heuristic_tree_bit_equal", "crossing3_cuda_heuristic_tree_bit_equal",
                 "crossing2_cuda_statistic_tree_bit_equal", "crossing3_cuda_statistic_tree_bit_equal",
                 "simple_20_iterations_tree_bit_equal", "simple_300_iterations_tree_bit_equal",
                 "static_rollout_tuple_equal", "parallel_search_merge_matches",
                 "batched_host_trees_equal_oracle", "batched_host_trees_cuda_statistic_equal_oracle",
                 "philox_crossing_tree_backprop_invariant", "philox_crossing_tree_backprop_invariant_cuda_statistic",
                 "philox_simple_tree_backprop_invariant", "philox_simple_tree_backprop_invariant_cuda_statistic",
                 "device_mcts_reference_const_equals_stock_mcts", "device_mcts_dot_export_labels_equal_reference",
                 "device_mcts_philox_equals_host_trees",
                 "device_mcts_simple_state_20_iterations",
                 "device_mcts_non_default_parameters_follow_the_state",
                 "device_mcts_mismatching_parameters_are_refused", "device_mcts_terminal_root_is_refused"):
        assert f"PASS {name}" in out, f"{name} did not pass:\n{out[-4000:]}"


def test_reference_verify_uct_on_gpu_rollouts():
    out = _run("dropin_verify_uct")
    assert "FAIL" not in out and " 0 failures" in out, out[-4000:]
    for name in ("dropin_mcts.reference_verify_uct_simple_state_cuda_heuristic",
                 "dropin_mcts.reference_small_search_depth_cuda_heuristic"):
        assert f"PASS {name}" in out, out[-4000:]


def test_reference_hypothesis_mcts_closed_loop_with_gpu_rollouts():
    """C4: the reference's closed-loop hypothesis tests (crossing_state_int_test.cc:141-229) with every
    rollout on the GPU (MAMCTS_SRC_HYPOTHESIS through CudaRandomHeuristic)."""
    out = _run("dropin_hypothesis", timeout=900)
    assert "FAIL" not in out and " 0 failures" in out, out[-4000:]
    for name in ("dropin_hypothesis.mcts_goal_reached_true_hypothesis_cuda_rollouts",
                 "dropin_hypothesis.mcts_goal_reached_wrong_hypothesis_cuda_rollouts",
                 "dropin_hypothesis.stock_heuristic_reaches_goal_too",
                 "dropin_hypothesis.episode_with_device_resident_planner_reaches_goal",
                 "dropin_hypothesis.device_hypothesis_mcts_true_hypothesis_equals_reference_search_every_step",
                 "dropin_hypothesis.device_hypothesis_mcts_wrong_hypothesis_equals_reference_search_every_step",
                 "dropin_hypothesis.device_hypothesis_mcts_root_parallel_reaches_goal"):
        assert f"PASS {name}" in out, out[-4000:]
