"""The reference's OWN tests for the path, compiled unmodified from the reference checkout against the
gtest / glog / boost stand-ins under oracle/shim (oracle/Makefile target `reftests`): test/uct/uct_test.cc
(verify_uct and friends: 5 tests, 663 checks), environments/tests/crossing_state_int_test.cc (execute,
hypothesis policies, belief tracker, closed-loop episodes: 7 tests), crossing_state_float_test.cc (the float
domain: 10 tests), test/hypothesis/hypothesis_statistic_test.cc (3 tests: the known answers the hypothesis
backup kernel is also checked against) and belief_tracker_test.cc (2 tests). They pin the reference build that
oracle/_ref (the checker the parity tests compare with) is made from: same sources, same flags, same shims.
The binaries are built where the reference checkout exists and travel with the repository snapshot."""
from __future__ import annotations

import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REFERENCE = os.environ.get("MAMCTS_REFERENCE", "/root/reference")


def _binary(name: str) -> str:
    path = os.path.join(REF_DIR, name)
    if not os.path.exists(path):
        if not os.path.isdir(os.path.join(REFERENCE, "mcts")):
            pytest.skip("no reference checkout and no prebuilt oracle/_ref/" + name)
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "reftests", f"REF={REFERENCE}"])
    return path


def _run(name: str, tmp_path, timeout: int):
    out = subprocess.run([_binary(name)], cwd=tmp_path, capture_output=True, text=True, timeout=timeout)
    passed = re.findall(r"^PASS (\S+)$", out.stdout, flags=re.M)
    failed = re.findall(r"^FAIL (\S+)$", out.stdout, flags=re.M)
    m = re.search(r"^(\d+) checks, (\d+) failures$", out.stdout, flags=re.M)
    assert m, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.returncode == 0 and not failed and int(m.group(2)) == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return passed, int(m.group(1))


def test_reference_uct_test_passes(tmp_path):
    passed, checks = _run("ref_uct_test", tmp_path, 120)
    assert passed == ["test_mcts.verify_uct", "test_mcts.small_search_depth", "test_mcts.generate_dot_file",
                      "test_mcts.visit_edges", "test_mcts.visit_nodes"]
    assert checks == 663


def test_reference_crossing_state_int_test_passes(tmp_path):
    passed, _checks = _run("ref_crossing_state_int_test", tmp_path, 300)
    assert passed == ["hypothesis_crossing_state.collision", "hypothesis_crossing_state.hypothesis_friendly",
                      "hypothesis_crossing_state.hypothesis_belief_correct",
                      "crossing_state.mcts_goal_reached_true_hypothesis",
                      "crossing_state.mcts_goal_reached_wrong_hypothesis",
                      "episode_runner.four_agents_reached_goal", "episode_runner.run_some_steps"]


def test_reference_crossing_state_float_test_passes(tmp_path):
    passed, _checks = _run("ref_crossing_state_float_test", tmp_path, 400)
    assert len(passed) == 10 and "hypothesis_crossing_state_float.policy_probability" in passed
    assert "hypothesis_crossing_state_float.collision" in passed and "episode_runner.run_some_steps" in passed


def test_reference_hypothesis_statistic_test_passes(tmp_path):
    passed, checks = _run("ref_hypothesis_statistic_test", tmp_path, 120)
    assert passed == ["hypothesis_statistic.backprop_hypothesis_action_selection",
                      "hypothesis_statistic.backprop_heuristic_hyp1",
                      "hypothesis_statistic.worst_case_action_selection"]
    assert checks == 20


def test_reference_belief_tracker_test_passes(tmp_path):
    passed, checks = _run("ref_belief_tracker_test", tmp_path, 120)
    assert passed == ["belief_tracker.simple_tracking_state", "belief_tracker.fixed_hypothesis_set"]
    assert checks == 13
