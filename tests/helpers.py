"""Shared helpers for the test-suite (golden fixtures, comparisons)."""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name: str):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def unhex(vals, shape=None) -> np.ndarray:
    a = np.array([float.fromhex(v) for v in vals], dtype=np.float64)
    return a if shape is None else a.reshape(shape)


def bits(a) -> np.ndarray:
    """Bit pattern view for exact float comparison (distinguishes -0.0 from 0.0)."""
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_bit_equal(a, b, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    bad = np.nonzero(bits(a).ravel() != bits(b).ravel())[0]
    assert bad.size == 0, f"{what}: {bad.size} of {a.size} values differ, first at {bad[:5]}: " \
                          f"{a.ravel()[bad[:5]]} vs {b.ravel()[bad[:5]]}"


def tree_digest(tree: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(tree, dtype=np.float64).tobytes()).hexdigest()


def rel_close(a, b, tol=1e-6):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= tol * np.maximum(np.abs(a), np.abs(b)) + 1e-300)
