"""ctypes mirror of include/mamcts_b200.h (the C ABI) and the loader of libmamcts_b200.so.

The shared library is the product; this module is only the Python binding a maintainer would
write against the header. There is NO fallback: if the library (or a CUDA device, for compute
calls) is missing the import / call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

MAX_AGENTS = 16
MAX_ACTIONS = 32
NCCL_UNIQUE_ID_BYTES = 128

OK = 0
MEM_HOST, MEM_DEVICE = 0, 1
SRC_REPLAY, SRC_REFERENCE_CONST, SRC_PHILOX, SRC_HYPOTHESIS = 0, 1, 2, 3
ENV_SIMPLE, ENV_CROSSING_I32 = 0, 1

c_double_p = C.POINTER(C.c_double)
c_u8_p = C.POINTER(C.c_uint8)
c_i8_p = C.POINTER(C.c_int8)
c_u32_p = C.POINTER(C.c_uint32)
c_i32_p = C.POINTER(C.c_int32)
c_u64_p = C.POINTER(C.c_uint64)
c_i64_p = C.POINTER(C.c_int64)


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("flags", C.c_uint32)]


class Params(C.Structure):
    """mamcts_params <- MctsParameters (reference mcts/mcts_parameters.h:18-41)."""

    _fields_ = [
        ("discount_factor", C.c_double),
        ("uct_lower_bound", C.c_double),
        ("uct_upper_bound", C.c_double),
        ("uct_exploration_constant", C.c_double),
        ("uct_pw_k", C.c_double),
        ("uct_pw_alpha", C.c_double),
        ("random_seed", C.c_uint32),
        ("max_number_of_iterations", C.c_uint32),
        ("max_search_depth", C.c_uint32),
        ("max_number_of_nodes", C.c_uint32),
        ("use_bound_estimation", C.c_uint32),
        ("rh_max_number_of_iterations", C.c_uint32),
        ("num_parallel_mcts", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class CrossingParams(C.Structure):
    """mamcts_crossing_params <- CrossingStateParameters<int> (crossing_state_parameters.h:15-33)."""

    _fields_ = [
        ("reward_collision", C.c_double),
        ("reward_goal_reached", C.c_double),
        ("reward_step", C.c_double),
        ("num_other_agents", C.c_uint32),
        ("num_other_actions", C.c_uint32),
        ("cost_only_collision", C.c_int32),
        ("min_velocity_ego", C.c_int32),
        ("max_velocity_ego", C.c_int32),
        ("min_velocity_other", C.c_int32),
        ("max_velocity_other", C.c_int32),
        ("chain_length", C.c_int32),
        ("ego_goal_pos", C.c_int32),
        ("reserved", C.c_int32),
    ]

    @property
    def num_agents(self) -> int:
        return self.num_other_agents + 1

    @property
    def num_ego_actions(self) -> int:
        return self.max_velocity_ego - self.min_velocity_ego + 1

    @property
    def crossing_point(self) -> int:
        return (self.chain_length - 1) // 2 + 1


class SimpleParams(C.Structure):
    _fields_ = [("winning_state_length", C.c_int32), ("loosing_state_length", C.c_int32)]


class CrossingParamsF32(C.Structure):
    """mamcts_crossing_params_f32 <- CrossingStateParameters<float>."""

    _fields_ = [
        ("reward_collision", C.c_double),
        ("reward_goal_reached", C.c_double),
        ("reward_step", C.c_double),
        ("num_other_agents", C.c_uint32),
        ("num_other_actions", C.c_uint32),
        ("cost_only_collision", C.c_int32),
        ("min_velocity_ego", C.c_float),
        ("max_velocity_ego", C.c_float),
        ("min_velocity_other", C.c_float),
        ("max_velocity_other", C.c_float),
        ("chain_length", C.c_float),
        ("ego_goal_pos", C.c_float),
        ("reserved", C.c_int32),
    ]


class ActionSource(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("mem", C.c_int32),
        ("replay_actions", C.c_void_p),
        ("replay_actions_f32", C.c_void_p),
        ("replay_steps", C.c_void_p),
        ("replay_stride", C.c_uint32),
        ("reserved", C.c_uint32),
        ("const_actions", C.c_uint32 * MAX_AGENTS),
        ("philox_seed", C.c_uint64),
        ("stream_hi", C.c_void_p),
        ("stream_lo", C.c_void_p),
        ("stream_hi_base", C.c_uint32),
        ("stream_lo_base", C.c_uint32),
        ("hyp_gap_lo", C.c_void_p),
        ("hyp_gap_hi", C.c_void_p),
        ("hyp_gap_lo_f32", C.c_void_p),
        ("hyp_gap_hi_f32", C.c_void_p),
    ]


class RolloutOut(C.Structure):
    _fields_ = [
        ("mem", C.c_int32),
        ("reserved", C.c_int32),
        ("ego_return", C.c_void_p),
        ("other_return", C.c_void_p),
        ("ego_cost", C.c_void_p),
        ("step_length", C.c_void_p),
        ("total_steps", C.c_void_p),
        ("steps", C.c_void_p),
        ("final_x", C.c_void_p),
        ("final_last_action", C.c_void_p),
        ("final_flags", C.c_void_p),
        ("final_x_f32", C.c_void_p),
        ("final_last_action_f32", C.c_void_p),
        ("reward_trace", C.c_void_p),
        ("trace_stride", C.c_uint32),
        ("reserved2", C.c_uint32),
        ("moments", C.c_void_p),
    ]


class HypothesisParams(C.Structure):
    """mamcts_hypothesis_params: MctsParameters::hypothesis_statistic (reference mcts/mcts_parameters.h:71-79)."""
    _fields_ = [
        ("upper_cost_bound", C.c_double),
        ("lower_cost_bound", C.c_double),
        ("pw_k", C.c_double),
        ("pw_alpha", C.c_double),
        ("exploration_constant", C.c_double),
        ("cost_based_action_selection", C.c_uint32),
        ("progressive_widening_hypothesis_based", C.c_uint32),
        ("chance_constrained_update", C.c_uint32),
        ("policy_random_seed", C.c_uint32),
    ]


MAX_HYPOTHESES = 4
HYP_MAX_OTHER_ACTIONS = 8


class HypSearchConfig(C.Structure):
    _fields_ = [
        ("num_trees", C.c_uint32),
        ("first_tree_id", C.c_uint32),
        ("max_nodes_per_tree", C.c_uint32),
        ("num_hypotheses", C.c_uint32),
        ("gap_lo", C.c_int32 * MAX_HYPOTHESES),
        ("gap_hi", C.c_int32 * MAX_HYPOTHESES),
        ("decorrelate_trees", C.c_uint32),
        ("reserved", C.c_uint32),
        ("mcts", Params),
        ("hypothesis", HypothesisParams),
        ("crossing", CrossingParams),
    ]


class ActionValuesOut(C.Structure):
    _fields_ = [
        ("mem", C.c_int32),
        ("reserved", C.c_int32),
        ("action_return", C.c_void_p),
        ("action_cost", C.c_void_p),
        ("action_step_length", C.c_void_p),
        ("other_return", C.c_void_p),
        ("mean_cost", C.c_void_p),
        ("total_steps", C.c_void_p),
    ]


class UctStats(C.Structure):
    _fields_ = [
        ("mem", C.c_int32),
        ("num_slots", C.c_uint32),
        ("max_actions", C.c_uint32),
        ("reserved", C.c_uint32),
        ("value", C.c_void_p),
        ("latest_return", C.c_void_p),
        ("total_visits", C.c_void_p),
        ("q", C.c_void_p),
        ("n", C.c_void_p),
    ]


class HypStats(C.Structure):
    _fields_ = [
        ("mem", C.c_int32),
        ("num_slots", C.c_uint32),
        ("num_hypotheses", C.c_uint32),
        ("max_actions", C.c_uint32),
        ("ego_cost_value", C.c_void_p),
        ("latest_ego_cost", C.c_void_p),
        ("total_visits", C.c_void_p),
        ("num_expanded", C.c_void_p),
        ("visits_hypothesis", C.c_void_p),
        ("action_cost", C.c_void_p),
        ("action_count", C.c_void_p),
    ]


LAYOUT_AUTO, LAYOUT_CLASSIC, LAYOUT_COMPACT = 0, 1, 2


class SearchConfig(C.Structure):
    _fields_ = [
        ("env", C.c_int32),
        ("action_source", C.c_int32),
        ("num_trees", C.c_uint32),
        ("first_tree_id", C.c_uint32),
        ("rollouts_per_leaf", C.c_uint32),
        ("max_nodes_per_tree", C.c_uint32),
        ("layout", C.c_uint32),
        ("max_inner_nodes_per_tree", C.c_uint32),
        ("philox_seed", C.c_uint64),
        ("mcts", Params),
        ("crossing", CrossingParams),
        ("simple", SimpleParams),
    ]


_LIB = None
_LIB_PATH = os.environ.get("MAMCTS_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc",
                                                            "libmamcts_b200.so")


def lib_path() -> str:
    return _LIB_PATH


def load() -> C.CDLL:
    """Loads libmamcts_b200.so (built in-tree by __graft_entry__.build()). Raises if absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(
            f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the mamcts_b200 engine)"
        )
    lib = C.CDLL(_LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, u32, u64, sz = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64, C.c_size_t
    P = C.POINTER
    sigs = {
        "mamcts_create": (i32, [P(Config), P(vp)]),
        "mamcts_destroy": (None, [vp]),
        "mamcts_last_error": (C.c_char_p, [vp]),
        "mamcts_abi_version": (i32, []),
        "mamcts_set_stream": (i32, [vp, vp]),
        "mamcts_get_stream": (vp, [vp]),
        "mamcts_sync": (i32, [vp]),
        "mamcts_launch_count": (u64, [vp]),
        "mamcts_device_alloc": (i32, [vp, sz, P(vp)]),
        "mamcts_device_free": (i32, [vp, vp]),
        "mamcts_memcpy_h2d": (i32, [vp, vp, vp, sz]),
        "mamcts_memcpy_d2h": (i32, [vp, vp, vp, sz]),
        "mamcts_default_params": (None, [P(Params)]),
        "mamcts_default_crossing_params": (None, [P(CrossingParams)]),
        "mamcts_default_simple_params": (None, [P(SimpleParams)]),
        "mamcts_reference_expand_order": (i32, [u32, u32, c_u32_p]),
        "mamcts_rollout_crossing_i32": (
            i32,
            [vp, P(Params), P(CrossingParams), u32, u32, i32, vp, vp, vp, vp, P(ActionSource),
             P(RolloutOut)],
        ),
        "mamcts_rollout_crossing_f32": (
            i32,
            [vp, P(Params), P(CrossingParamsF32), u32, u32, i32, vp, vp, vp, vp, P(ActionSource),
             P(RolloutOut)],
        ),
        "mamcts_default_crossing_params_f32": (None, [P(CrossingParamsF32)]),
        "mamcts_action_values_crossing_i32": (
            i32,
            [vp, P(Params), P(CrossingParams), u32, i32, vp, vp, vp, vp, P(ActionSource), P(ActionValuesOut)],
        ),
        "mamcts_rollout_simple": (
            i32,
            [vp, P(Params), P(SimpleParams), u32, u32, i32, vp, vp, P(ActionSource), P(RolloutOut)],
        ),
        "mamcts_backup_uct": (
            i32,
            [vp, P(Params), u32, u32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, P(UctStats)],
        ),
        "mamcts_backup_hypothesis": (
            i32,
            [vp, P(Params), u32, u32, i32, vp, vp, vp, u32, vp, vp, vp, vp, vp, vp, i32, P(HypStats)],
        ),
        "mamcts_root_merge": (
            i32,
            [vp, P(UctStats), u32, i32, vp, u32, c_double_p, c_u64_p, c_u32_p, c_u64_p, c_double_p,
             c_u32_p],
        ),
        "mamcts_rollout_moments": (i32, [vp, i32, vp, vp, u32, u32, vp]),
        "mamcts_comm_get_unique_id": (i32, [vp, c_u8_p]),
        "mamcts_comm_init_rank": (i32, [vp, c_u8_p, i32, i32]),
        "mamcts_comm_destroy": (i32, [vp]),
        "mamcts_allreduce_f64": (i32, [vp, vp, sz]),
        "mamcts_default_hypothesis_params": (None, [P(HypothesisParams)]),
        "mamcts_reference_hypothesis_sequence": (i32, [u32, vp, u32, u32, u32, u32, c_u8_p]),
        "mamcts_hyp_search_create": (i32, [vp, P(HypSearchConfig), c_i32_p, c_i32_p, c_u8_p, P(vp)]),
        "mamcts_hyp_search_reset": (i32, [vp, c_i32_p, c_i32_p, c_u8_p]),
        "mamcts_search_create": (i32, [vp, P(SearchConfig), c_i32_p, c_i32_p, P(vp)]),
        "mamcts_search_destroy": (i32, [vp]),
        "mamcts_search_reset": (i32, [vp, c_i32_p, c_i32_p]),
        "mamcts_search_run": (i32, [vp, u32]),
        "mamcts_search_counters": (i32, [vp, c_u64_p, c_u64_p, c_u64_p, c_u64_p, c_u64_p]),
        "mamcts_search_exact_ucb_evaluations": (i32, [vp, c_u64_p]),
        "mamcts_search_root_stats": (
            i32, [vp, u32, u32, c_double_p, c_u32_p, c_double_p, c_u32_p, c_u32_p, c_u32_p]),
        "mamcts_search_merge_roots": (
            i32, [vp, c_double_p, c_u64_p, c_u32_p, c_u64_p, c_double_p, c_u32_p]),
        "mamcts_search_dump_tree": (i32, [vp, u32, c_double_p, u32, c_u32_p]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export the symbol
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


EXPORTED_SYMBOLS = [
    "mamcts_create", "mamcts_destroy", "mamcts_last_error", "mamcts_abi_version",
    "mamcts_set_stream", "mamcts_get_stream", "mamcts_sync", "mamcts_launch_count",
    "mamcts_device_alloc", "mamcts_device_free", "mamcts_memcpy_h2d", "mamcts_memcpy_d2h",
    "mamcts_default_params", "mamcts_default_crossing_params", "mamcts_default_simple_params",
    "mamcts_reference_expand_order", "mamcts_rollout_crossing_i32", "mamcts_rollout_simple",
    "mamcts_rollout_crossing_f32", "mamcts_default_crossing_params_f32", "mamcts_action_values_crossing_i32",
    "mamcts_backup_uct", "mamcts_backup_hypothesis", "mamcts_root_merge", "mamcts_rollout_moments", "mamcts_comm_get_unique_id",
    "mamcts_comm_init_rank", "mamcts_comm_destroy", "mamcts_allreduce_f64",
    "mamcts_default_hypothesis_params", "mamcts_reference_hypothesis_sequence", "mamcts_hyp_search_create",
    "mamcts_hyp_search_reset",
    "mamcts_search_create", "mamcts_search_destroy", "mamcts_search_reset", "mamcts_search_run",
    "mamcts_search_counters", "mamcts_search_exact_ucb_evaluations", "mamcts_search_root_stats", "mamcts_search_merge_roots",
    "mamcts_search_dump_tree",
]
