"""Groundwork for pinning HypothesisStatistic's selection on the device:
get_worst_case_action (reference mcts/hypothesis/hypothesis_statistic.h:178-202) breaks ties by the iteration
order of std::unordered_map<ActionIdx, UcbPair>. With this toolchain's libstdc++ that order is, for the up to
7 integer keys of an agent's actions, exactly the REVERSE of the insertion order (13 buckets after the first
insertion, identity hash, every key in its own bucket, new nodes linked at the front). tests/aux/
unordered_map_order.cc checks that on 2 000 random insertion sequences."""
from __future__ import annotations

import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_unordered_map_iterates_small_integer_keys_in_reverse_insertion_order(tmp_path):
    if not shutil.which("g++"):
        pytest.skip("no g++")
    exe = tmp_path / "um"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", str(exe),
                           os.path.join(ROOT, "tests", "aux", "unordered_map_order.cc")])
    out = subprocess.check_output([str(exe)], text=True)
    assert "mismatches 0 of 2000; bucket_count after 1 insert: 13" in out, out
