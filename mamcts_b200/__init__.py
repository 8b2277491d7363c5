"""mamcts_b200 - B200-native rollout-and-backup engine for multi-agent MCTS (juloberno/mamcts).

The product is mamcts_b200/csrc/libmamcts_b200.so (hand-written CUDA for sm_100a behind the C ABI
in include/mamcts_b200.h). This package is the thin Python binding over that ABI; it has no CPU
fallback.
"""
from . import _abi  # noqa: F401
from .params import (default_crossing_params, default_crossing_params_f32, default_hypothesis_params,  # noqa: F401
                     default_params, default_simple_params)

__all__ = ["default_params", "default_crossing_params", "default_crossing_params_f32", "default_simple_params",
           "default_hypothesis_params"]
