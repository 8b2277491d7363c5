"""Host-side pieces of the device-resident hypothesis-MCTS that need no GPU."""
from __future__ import annotations

import numpy as np

from helpers import golden
from mamcts_b200.engine import reference_hypothesis_sequence


def test_hypothesis_sequence_helper_equals_the_tracker():
    """mamcts_reference_hypothesis_sequence = the draws of a real HypothesisBeliefTracker with uniform beliefs
    (reference hypothesis_belief_tracker.h:136-149), recorded from the reference's own tracker by
    tests/golden/make_golden.py (sequence, agents in the tracker's container order, earlier samples skipped)."""
    for case in golden("hypothesis_search.json"):
        J, H = case["others"], len(case["gaps"])
        seq = reference_hypothesis_sequence(1000, np.full((J, H), 1.0 / H), case["skip"], case["iters"])
        assert np.array_equal(seq.ravel(), np.array(case["sequence"], dtype=np.uint8)), case["name"]
