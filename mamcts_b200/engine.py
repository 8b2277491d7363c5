"""Thin Python host over the C ABI (include/mamcts_b200.h): numpy in / numpy out for HOST buffers,
raw device pointers (e.g. torch tensors' data_ptr()) for DEVICE buffers.

Mirrors the reference surfaces that sit on the hot path:
  Engine.rollout_crossing / rollout_simple  <- RandomHeuristic::rollout
                                               (reference mcts/heuristics/random_heuristic.h:27-104)
  Engine.backup_uct                         <- the backprop loop, mcts/mcts.h:309-329
  Engine.root_merge                         <- UctStatistic::merge_node_statistics, uct_statistic.h:165-184
  Search                                    <- Mcts<S,UctStatistic,UctStatistic,RandomHeuristic>::search
                                               for T root-parallel trees, mcts/mcts.h:133-221
There is no CPU path here: every compute call goes to libmamcts_b200.so and raises on error.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _abi
from ._abi import (ENV_CROSSING_I32, ENV_SIMPLE, MAX_ACTIONS, MEM_DEVICE, MEM_HOST, SRC_HYPOTHESIS, SRC_PHILOX,
                   SRC_REFERENCE_CONST, SRC_REPLAY, ActionSource, Config, CrossingParams, Params,
                   RolloutOut, SearchConfig, SimpleParams, UctStats)


class MamctsError(RuntimeError):
    pass


def _vp(a):
    """void* of a numpy array (host) / int device pointer / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))


def reference_expand_order(seed: int, num_actions: int):
    """UctStatistic's deterministic widening order (reference uct_statistic.h:61-69)."""
    lib = _abi.load()
    out = (C.c_uint32 * num_actions)()
    rc = lib.mamcts_reference_expand_order(seed, num_actions, out)
    if rc != 0:
        raise MamctsError(lib.mamcts_last_error(None).decode())
    return list(out)


class Engine:
    """One CUDA device + stream (mamcts_engine)."""

    def __init__(self, device: int = 0):
        self.lib = _abi.load()
        self.h = C.c_void_p()
        cfg = Config(device=device, flags=0)
        rc = self.lib.mamcts_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            raise MamctsError(f"mamcts_create failed ({rc}): {self.lib.mamcts_last_error(None).decode()}")
        self.device = device
        self._searches = weakref.WeakSet()   # closed before the engine they live on

    # -- plumbing ---------------------------------------------------------------------------
    def _check(self, rc: int, what: str):
        if rc != 0:
            raise MamctsError(f"{what} failed ({rc}): {self.lib.mamcts_last_error(self.h).decode()}")

    def close(self):
        if self.h:
            for s in list(self._searches):
                s.close()
            self.lib.mamcts_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._check(self.lib.mamcts_sync(self.h), "mamcts_sync")

    def set_stream(self, cuda_stream: int):
        self._check(self.lib.mamcts_set_stream(self.h, C.c_void_p(cuda_stream)), "mamcts_set_stream")

    @property
    def stream(self) -> int:
        return int(self.lib.mamcts_get_stream(self.h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.lib.mamcts_launch_count(self.h))

    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(self.lib.mamcts_device_alloc(self.h, nbytes, C.byref(p)), "mamcts_device_alloc")
        return int(p.value)

    def device_free(self, ptr: int):
        self._check(self.lib.mamcts_device_free(self.h, C.c_void_p(ptr)), "mamcts_device_free")

    def h2d(self, dptr: int, arr: np.ndarray):
        self._check(self.lib.mamcts_memcpy_h2d(self.h, C.c_void_p(dptr), _vp(arr), arr.nbytes), "h2d")

    def d2h(self, arr: np.ndarray, dptr: int):
        self._check(self.lib.mamcts_memcpy_d2h(self.h, _vp(arr), C.c_void_p(dptr), arr.nbytes), "d2h")

    # -- NCCL ---------------------------------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = (C.c_uint8 * _abi.NCCL_UNIQUE_ID_BYTES)()
        self._check(self.lib.mamcts_comm_get_unique_id(self.h, buf), "mamcts_comm_get_unique_id")
        return bytes(buf)

    def comm_init_rank(self, unique_id: bytes, world_size: int, rank: int):
        buf = (C.c_uint8 * _abi.NCCL_UNIQUE_ID_BYTES)(*unique_id)
        self._check(self.lib.mamcts_comm_init_rank(self.h, buf, world_size, rank), "mamcts_comm_init_rank")

    def comm_destroy(self):
        self._check(self.lib.mamcts_comm_destroy(self.h), "mamcts_comm_destroy")

    def allreduce_f64(self, values: np.ndarray) -> np.ndarray:
        """Sum all-reduce of a small vector of doubles over the engine's NCCL communicator (host in, host
        out: the payload goes through a device buffer; the collective itself is mamcts_allreduce_f64)."""
        v = np.ascontiguousarray(values, dtype=np.float64)
        d = self.device_alloc(v.nbytes)
        try:
            self.h2d(d, v)
            self._check(self.lib.mamcts_allreduce_f64(self.h, C.c_void_p(d), v.size), "mamcts_allreduce_f64")
            out = np.empty_like(v)
            self.d2h(out, d)
            self.sync()
        finally:
            self.device_free(d)
        return out

    def rollout_moments(self, ego_return, other_return, num_others: int = None, n: int = None, d_out: int = 0):
        """{sum, sum of squares} per agent + count of per-rollout returns (C3's exchange step), reduced on the
        device and all-reduced over the engine's communicator (mamcts_rollout_moments). numpy arrays = HOST
        buffers ([n] and [A-1, n]), result returned as numpy; ints = device pointers (num_others and n
        required): with d_out (device, 2A+1 doubles) the call is asynchronous on the engine stream and returns
        None, without it the result is copied back."""
        if isinstance(ego_return, np.ndarray):
            ego = np.ascontiguousarray(ego_return, dtype=np.float64)
            oth = None if other_return is None else np.ascontiguousarray(other_return, dtype=np.float64)
            n = ego.shape[0]
            num_others = 0 if oth is None else oth.reshape(-1, n).shape[0]
            out = np.zeros(2 * (num_others + 1) + 1)
            self._check(self.lib.mamcts_rollout_moments(self.h, MEM_HOST, _vp(ego), _vp(oth), num_others, n, _vp(out)),
                        "mamcts_rollout_moments")
            return out
        nacc = 2 * (num_others + 1) + 1
        d = d_out or self.device_alloc(8 * nacc)
        try:
            self._check(self.lib.mamcts_rollout_moments(self.h, MEM_DEVICE, _vp(ego_return), _vp(other_return or None),
                                                        num_others, n, C.c_void_p(d)), "mamcts_rollout_moments")
            if d_out:
                return None
            out = np.zeros(nacc)
            self.d2h(out, d)
            self.sync()
        finally:
            if not d_out:
                self.device_free(d)
        return out

    # -- rollouts -----------------------------------------------------------------------------
    @staticmethod
    def _source(kind, A, replay_actions=None, replay_steps=None, const_actions=None, seed=0,
                stream_hi=None, stream_hi_base=0, stream_lo_base=0, stream_lo=None, mem=MEM_HOST,
                hyp_gap_lo=None, hyp_gap_hi=None, f32=False):
        src = ActionSource()
        src.kind = kind
        src.mem = mem
        keep = []
        if kind == SRC_REPLAY:
            if mem == MEM_HOST:
                ra = np.ascontiguousarray(replay_actions, dtype=np.int8)
                rs = np.ascontiguousarray(replay_steps, dtype=np.uint32)
                keep += [ra, rs]
                src.replay_stride = ra.shape[1]
                src.replay_actions = ra.ctypes.data
                src.replay_steps = rs.ctypes.data
            else:  # device pointers; replay_actions = (ptr, stride)
                src.replay_actions, src.replay_stride = int(replay_actions[0]), int(replay_actions[1])
                src.replay_steps = int(replay_steps)
        elif kind == SRC_REFERENCE_CONST:
            for a in range(A):
                src.const_actions[a] = int(const_actions[a])
        else:
            src.philox_seed = seed
            src.stream_hi_base = stream_hi_base
            src.stream_lo_base = stream_lo_base
            if stream_hi is not None:
                if mem == MEM_HOST:
                    sh = np.ascontiguousarray(stream_hi, dtype=np.uint32)
                    keep.append(sh)
                    src.stream_hi = sh.ctypes.data
                else:
                    src.stream_hi = int(stream_hi)
            if stream_lo is not None:
                if mem == MEM_HOST:
                    sl = np.ascontiguousarray(stream_lo, dtype=np.uint32)
                    keep.append(sl)
                    src.stream_lo = sl.ctypes.data
                else:
                    src.stream_lo = int(stream_lo)
            if kind == SRC_HYPOTHESIS and hyp_gap_lo is not None and hyp_gap_hi is not None:
                # [A-1, n] desired-gap ranges of the other agents' current hypotheses (missing: the call fails)
                if mem == MEM_HOST:
                    lo = np.ascontiguousarray(hyp_gap_lo, dtype=np.float32 if f32 else np.int32)
                    hi = np.ascontiguousarray(hyp_gap_hi, dtype=np.float32 if f32 else np.int32)
                    keep += [lo, hi]
                    plo, phi = lo.ctypes.data, hi.ctypes.data
                else:
                    plo, phi = int(hyp_gap_lo), int(hyp_gap_hi)
                if f32:   # CrossingState<float>: float ranges
                    src.hyp_gap_lo_f32, src.hyp_gap_hi_f32 = plo, phi
                else:
                    src.hyp_gap_lo, src.hyp_gap_hi = plo, phi
        return src, keep

    def _rollout_host(self, env, mp, envp, A, x, last, flags, depth, K, src_kw, want_parity,
                      trace_stride, want_moments=False):
        x = np.ascontiguousarray(x, dtype=np.int32)
        if env == ENV_SIMPLE:
            x = x.reshape(1, -1)
        n = x.shape[1]
        total = n * K
        last_a = None if last is None else np.ascontiguousarray(last, dtype=np.int32)
        flags_a = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
        depth_a = None if depth is None else np.ascontiguousarray(depth, dtype=np.uint32)
        src, keep = self._source(A=A, **src_kw)
        res = dict(ego_return=np.zeros(n), other_return=np.zeros((A - 1, n)), ego_cost=np.zeros(n),
                   step_length=np.zeros(n), total_steps=np.zeros(1, dtype=np.uint64))
        out = RolloutOut()
        out.mem = MEM_HOST
        out.ego_return = res["ego_return"].ctypes.data
        out.other_return = res["other_return"].ctypes.data
        out.ego_cost = res["ego_cost"].ctypes.data
        out.step_length = res["step_length"].ctypes.data
        out.total_steps = res["total_steps"].ctypes.data
        if want_moments:   # {sum, sum^2} per agent + count over all n*K per-rollout returns, reduced on the device
            res["moments"] = np.zeros(2 * A + 1)
            out.moments = res["moments"].ctypes.data
        if want_parity:
            xr = 1 if env == ENV_SIMPLE else A
            res.update(steps=np.zeros(total, dtype=np.uint32),
                       final_x=np.zeros((xr, total), dtype=np.int32),
                       final_last_action=np.zeros((A, total), dtype=np.int32),
                       final_flags=np.zeros(total, dtype=np.uint8))
            out.steps = res["steps"].ctypes.data
            out.final_x = res["final_x"].ctypes.data
            out.final_last_action = res["final_last_action"].ctypes.data
            out.final_flags = res["final_flags"].ctypes.data
            if trace_stride:
                res["reward_trace"] = np.zeros((total, trace_stride))
                out.reward_trace = res["reward_trace"].ctypes.data
                out.trace_stride = trace_stride
        if env == ENV_CROSSING_I32:
            rc = self.lib.mamcts_rollout_crossing_i32(
                self.h, C.byref(mp), C.byref(envp), n, K, MEM_HOST, _vp(x), _vp(last_a), _vp(flags_a),
                _vp(depth_a), C.byref(src), C.byref(out))
            self._check(rc, "mamcts_rollout_crossing_i32")
        else:
            rc = self.lib.mamcts_rollout_simple(self.h, C.byref(mp), C.byref(envp), n, K, MEM_HOST,
                                                _vp(x), _vp(depth_a), C.byref(src), C.byref(out))
            self._check(rc, "mamcts_rollout_simple")
        res["total_steps"] = int(res["total_steps"][0])
        del keep
        return res

    def rollout_crossing(self, mp: Params, cp: CrossingParams, x, last=None, flags=None, depth=None,
                         K: int = 1, kind: int = SRC_PHILOX, parity: bool = False,
                         trace_stride: int = 0, moments: bool = False, **src_kw):
        """HOST-buffer batched rollout through CrossingState<int>::execute. x: [A, n] int32.
        moments=True adds res["moments"]: {sum, sum^2} per agent + count over all n*K rollouts."""
        return self._rollout_host(ENV_CROSSING_I32, mp, cp, cp.num_other_agents + 1, x, last, flags,
                                  depth, K, dict(kind=kind, **src_kw), parity, trace_stride, moments)

    def action_values_crossing(self, mp: Params, cp: CrossingParams, x, last=None, flags=None, depth=None,
                               kind: int = SRC_PHILOX, **src_kw):
        """ActionValueRandomHeuristic's compute part (reference action_value_random_heuristic.h:37-85) for n leaves:
        per ego action one first step + one rollout (mamcts_action_values_crossing_i32). x: [A, n] int32, HOST.
        Returns action_return / action_cost / action_step_length [n, nA], other_return [A-1, n], mean_cost [n]."""
        from ._abi import ActionValuesOut
        A, nA = cp.num_other_agents + 1, cp.num_ego_actions
        x = np.ascontiguousarray(x, dtype=np.int32)
        n = x.shape[1]
        last_a = None if last is None else np.ascontiguousarray(last, dtype=np.int32)
        flags_a = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
        depth_a = None if depth is None else np.ascontiguousarray(depth, dtype=np.uint32)
        src, keep = self._source(A=A, kind=kind, **src_kw)
        res = dict(action_return=np.zeros((n, nA)), action_cost=np.zeros((n, nA)), action_step_length=np.zeros((n, nA)),
                   other_return=np.zeros((A - 1, n)), mean_cost=np.zeros(n), total_steps=np.zeros(1, dtype=np.uint64))
        out = ActionValuesOut()
        out.mem = MEM_HOST
        for k in res:
            setattr(out, k, res[k].ctypes.data)
        rc = self.lib.mamcts_action_values_crossing_i32(self.h, C.byref(mp), C.byref(cp), n, MEM_HOST, _vp(x), _vp(last_a),
                                                        _vp(flags_a), _vp(depth_a), C.byref(src), C.byref(out))
        self._check(rc, "mamcts_action_values_crossing_i32")
        res["total_steps"] = int(res["total_steps"][0])
        del keep
        return res

    def rollout_crossing_f32(self, mp: Params, cp, x, last=None, flags=None, depth=None, K: int = 1,
                             kind: int = SRC_PHILOX, parity: bool = False, replay_actions=None,
                             replay_steps=None, **src_kw):
        """HOST-buffer batched rollout through CrossingState<float>::execute. x / last: [A, n] float32;
        replay_actions: [n, stride, A] float32 (ego index, others' velocities)."""
        A = cp.num_other_agents + 1
        x = np.ascontiguousarray(x, dtype=np.float32)
        n = x.shape[1]
        total = n * K
        last_a = None if last is None else np.ascontiguousarray(last, dtype=np.float32)
        flags_a = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
        depth_a = None if depth is None else np.ascontiguousarray(depth, dtype=np.uint32)
        if kind == SRC_REPLAY:
            src = ActionSource()
            src.kind, src.mem = SRC_REPLAY, MEM_HOST
            ra = np.ascontiguousarray(replay_actions, dtype=np.float32)
            rs = np.ascontiguousarray(replay_steps, dtype=np.uint32)
            src.replay_actions_f32, src.replay_steps, src.replay_stride = ra.ctypes.data, rs.ctypes.data, ra.shape[1]
            keep = [ra, rs]
        else:
            src, keep = self._source(A=A, kind=kind, f32=True, **src_kw)
        res = dict(ego_return=np.zeros(n), other_return=np.zeros((A - 1, n)), ego_cost=np.zeros(n),
                   step_length=np.zeros(n), total_steps=np.zeros(1, dtype=np.uint64))
        out = RolloutOut()
        out.mem = MEM_HOST
        for k in ("ego_return", "other_return", "ego_cost", "step_length", "total_steps"):
            setattr(out, k, res[k].ctypes.data)
        if parity:
            res.update(steps=np.zeros(total, dtype=np.uint32), final_x=np.zeros((A, total), dtype=np.float32),
                       final_last_action=np.zeros((A, total), dtype=np.float32),
                       final_flags=np.zeros(total, dtype=np.uint8))
            out.steps = res["steps"].ctypes.data
            out.final_x_f32 = res["final_x"].ctypes.data
            out.final_last_action_f32 = res["final_last_action"].ctypes.data
            out.final_flags = res["final_flags"].ctypes.data
        rc = self.lib.mamcts_rollout_crossing_f32(self.h, C.byref(mp), C.byref(cp), n, K, MEM_HOST, _vp(x), _vp(last_a),
                                                  _vp(flags_a), _vp(depth_a), C.byref(src), C.byref(out))
        self._check(rc, "mamcts_rollout_crossing_f32")
        res["total_steps"] = int(res["total_steps"][0])
        del keep
        return res

    def rollout_simple(self, mp: Params, sp: SimpleParams, length, depth=None, K: int = 1,
                       kind: int = SRC_PHILOX, parity: bool = False, trace_stride: int = 0, **src_kw):
        """HOST-buffer batched rollout through SimpleState::execute. length: [n] int32."""
        return self._rollout_host(ENV_SIMPLE, mp, sp, 2, length, None, None, depth, K,
                                  dict(kind=kind, **src_kw), parity, trace_stride)

    def rollout_crossing_device(self, mp: Params, cp: CrossingParams, n: int, K: int, d_x: int,
                                d_ego: int, d_other: int = 0, d_cost: int = 0, d_len: int = 0,
                                d_total_steps: int = 0, d_last: int = 0, d_flags: int = 0,
                                d_depth: int = 0, kind: int = SRC_PHILOX, d_moments: int = 0, **src_kw):
        """DEVICE-buffer variant: raw device pointers, fully asynchronous on the engine stream."""
        src, keep = self._source(A=cp.num_other_agents + 1, kind=kind, mem=MEM_DEVICE, **src_kw)
        out = RolloutOut()
        out.mem = MEM_DEVICE
        out.ego_return = d_ego or None
        out.other_return = d_other or None
        out.ego_cost = d_cost or None
        out.step_length = d_len or None
        out.total_steps = d_total_steps or None
        out.moments = d_moments or None
        rc = self.lib.mamcts_rollout_crossing_i32(
            self.h, C.byref(mp), C.byref(cp), n, K, MEM_DEVICE, C.c_void_p(d_x),
            C.c_void_p(d_last) if d_last else None, C.c_void_p(d_flags) if d_flags else None,
            C.c_void_p(d_depth) if d_depth else None, C.byref(src), C.byref(out))
        self._check(rc, "mamcts_rollout_crossing_i32")
        del keep

    # -- backup -------------------------------------------------------------------------------
    def backup_uct(self, mp: Params, num_agents: int, leaf_node, leaf_has_heuristic, ego_return,
                   other_return, path_offset, path_len, edge_node, edge_action, edge_reward,
                   stats: dict):
        """HOST-buffer batched backup; `stats` holds numpy arrays value/latest_return/total_visits
        [slots], q/n [slots, max_actions] and is updated in place."""
        leaf_node = np.ascontiguousarray(leaf_node, dtype=np.uint32)
        n_paths = leaf_node.shape[0]
        has = np.ascontiguousarray(leaf_has_heuristic, dtype=np.uint8)
        ego = np.ascontiguousarray(ego_return, dtype=np.float64)
        oth = np.ascontiguousarray(other_return, dtype=np.float64)
        off = np.ascontiguousarray(path_offset, dtype=np.uint32)
        ln = np.ascontiguousarray(path_len, dtype=np.uint32)
        en = np.ascontiguousarray(edge_node, dtype=np.uint32)
        ea = np.ascontiguousarray(edge_action, dtype=np.uint8)
        er = np.ascontiguousarray(edge_reward, dtype=np.float64)
        st = self._stats_struct(stats)
        rc = self.lib.mamcts_backup_uct(self.h, C.byref(mp), num_agents, n_paths, MEM_HOST,
                                        _vp(leaf_node), _vp(has), _vp(ego), _vp(oth), _vp(off),
                                        _vp(ln), _vp(en), _vp(ea), _vp(er), C.byref(st))
        self._check(rc, "mamcts_backup_uct")

    @staticmethod
    def new_hyp_stats(num_slots: int, num_hypotheses: int, max_actions: int) -> dict:
        """Empty HypothesisStatistic storage (numpy, HOST) in the layout of mamcts_hyp_stats."""
        S, H, MA = num_slots, num_hypotheses, max_actions
        return dict(ego_cost_value=np.zeros(S), latest_ego_cost=np.zeros(S),
                    total_visits=np.zeros(S, dtype=np.uint32), num_expanded=np.zeros(S, dtype=np.uint32),
                    visits_hypothesis=np.zeros((S, H + 1), dtype=np.uint32),
                    action_cost=np.zeros((S, H, MA)), action_count=np.zeros((S, H, MA), dtype=np.uint32))

    @staticmethod
    def hyp_stats_struct(stats: dict):
        from ._abi import HypStats
        st = HypStats()
        st.mem = MEM_HOST
        st.num_slots = stats["ego_cost_value"].shape[0]
        st.num_hypotheses = stats["action_cost"].shape[1]
        st.max_actions = stats["action_cost"].shape[2]
        for k in ("ego_cost_value", "latest_ego_cost", "total_visits", "num_expanded", "visits_hypothesis",
                  "action_cost", "action_count"):
            assert stats[k].flags["C_CONTIGUOUS"], k
            setattr(st, k, stats[k].ctypes.data)
        return st

    def backup_hypothesis(self, mp: Params, num_others: int, leaf_node, leaf_has_heuristic, leaf_cost,
                          path_offset, path_len, edge_node, edge_hypothesis, edge_action, edge_cost,
                          stats: dict, chance_constrained_update: bool = False):
        """HOST-buffer batched HypothesisStatistic backup (C4); `stats` from new_hyp_stats, updated in
        place. leaf_cost: [n_paths, C]; edge_cost: [n_edges, C]; edge_hypothesis / edge_action: [n_edges, J]."""
        leaf_node = np.ascontiguousarray(leaf_node, dtype=np.uint32)
        has = np.ascontiguousarray(leaf_has_heuristic, dtype=np.uint8)
        lc = np.ascontiguousarray(leaf_cost, dtype=np.float64)
        off = np.ascontiguousarray(path_offset, dtype=np.uint32)
        ln = np.ascontiguousarray(path_len, dtype=np.uint32)
        en = np.ascontiguousarray(edge_node, dtype=np.uint32)
        eh = np.ascontiguousarray(edge_hypothesis, dtype=np.uint8)
        ea = np.ascontiguousarray(edge_action, dtype=np.uint8)
        ec = np.ascontiguousarray(edge_cost, dtype=np.float64)
        st = self.hyp_stats_struct(stats)
        rc = self.lib.mamcts_backup_hypothesis(self.h, C.byref(mp), num_others, leaf_node.shape[0], MEM_HOST,
                                               _vp(leaf_node), _vp(has), _vp(lc), lc.shape[1], _vp(off), _vp(ln),
                                               _vp(en), _vp(eh), _vp(ea), _vp(ec),
                                               1 if chance_constrained_update else 0, C.byref(st))
        self._check(rc, "mamcts_backup_hypothesis")

    @staticmethod
    def _stats_struct(stats: dict) -> UctStats:
        st = UctStats()
        st.mem = MEM_HOST
        st.num_slots = stats["value"].shape[0]
        st.max_actions = stats["q"].shape[1]
        for k, dt in (("value", np.float64), ("latest_return", np.float64),
                      ("total_visits", np.uint32), ("q", np.float64), ("n", np.uint32)):
            a = stats[k]
            assert a.dtype == dt and a.flags["C_CONTIGUOUS"], k
        st.value = stats["value"].ctypes.data
        st.latest_return = stats["latest_return"].ctypes.data
        st.total_visits = stats["total_visits"].ctypes.data
        st.q = stats["q"].ctypes.data
        st.n = stats["n"].ctypes.data
        return st

    def root_merge(self, stats: dict, root_slot, num_actions: int):
        st = self._stats_struct(stats)
        slots = np.ascontiguousarray(root_slot, dtype=np.uint32)
        return self._merge_call(
            lambda *o: self.lib.mamcts_root_merge(self.h, C.byref(st), slots.shape[0], MEM_HOST,
                                                  _vp(slots), num_actions, *o),
            num_actions, "mamcts_root_merge")

    def _merge_call(self, fn, num_actions, what):
        q = (C.c_double * MAX_ACTIONS)()
        n = (C.c_uint64 * MAX_ACTIONS)()
        occ = (C.c_uint32 * MAX_ACTIONS)()
        vis = C.c_uint64(0)
        val = C.c_double(0)
        best = C.c_uint32(0)
        self._check(fn(q, n, occ, C.byref(vis), C.byref(val), C.byref(best)), what)
        return dict(q=np.array(q[:num_actions]), n=np.array(n[:num_actions], dtype=np.uint64),
                    occurrences=np.array(occ[:num_actions], dtype=np.uint32), visits=vis.value,
                    value=val.value, best=best.value)


class Search:
    """T device-resident root-parallel trees (mamcts_search)."""

    def __init__(self, engine: Engine, env: int, mp: Params, num_trees: int, cp: CrossingParams = None,
                 sp: SimpleParams = None, root_x=None, action_source: int = SRC_PHILOX,
                 philox_seed: int = 0, first_tree_id: int = 0, rollouts_per_leaf: int = 1,
                 max_nodes_per_tree: int = 0, layout: int = 0, max_inner_nodes_per_tree: int = 0):
        self.engine = engine
        self.lib = engine.lib
        cfg = SearchConfig()
        cfg.env = env
        cfg.action_source = action_source
        cfg.num_trees = num_trees
        cfg.first_tree_id = first_tree_id
        cfg.rollouts_per_leaf = rollouts_per_leaf
        cfg.max_nodes_per_tree = max_nodes_per_tree
        cfg.layout = layout
        cfg.max_inner_nodes_per_tree = max_inner_nodes_per_tree
        cfg.philox_seed = philox_seed
        cfg.mcts = mp
        if cp is not None:
            cfg.crossing = cp
        if sp is not None:
            cfg.simple = sp
        self.cfg = cfg
        if env == ENV_CROSSING_I32:
            self.A = cp.num_other_agents + 1
            self.num_actions = [cp.num_ego_actions] + [cp.num_other_actions] * (self.A - 1)
            default_root = [5] * self.A
        else:
            self.A = 2
            self.num_actions = [2, 2]
            default_root = [4]
        rx = np.zeros(_abi.MAX_AGENTS, dtype=np.int32)
        root = default_root if root_x is None else list(root_x)
        rx[:len(root)] = root
        self.h = C.c_void_p()
        rc = self.lib.mamcts_search_create(engine.h, C.byref(cfg), rx.ctypes.data_as(_abi.c_i32_p), None,
                                           C.byref(self.h))
        engine._check(rc, "mamcts_search_create")
        engine._searches.add(self)

    def close(self):
        if self.h and self.engine.h:
            self.lib.mamcts_search_destroy(self.h)
        self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, root_x=None):
        """New decision from root_x (host); None = keep the root already on the device."""
        ptr = None
        if root_x is not None:
            rx = np.zeros(_abi.MAX_AGENTS, dtype=np.int32)
            rx[:len(root_x)] = root_x
            ptr = rx.ctypes.data_as(_abi.c_i32_p)
        self.engine._check(self.lib.mamcts_search_reset(self.h, ptr, None), "mamcts_search_reset")

    def run(self, iterations: int):
        self.engine._check(self.lib.mamcts_search_run(self.h, iterations), "mamcts_search_run")

    def counters(self):
        v = [C.c_uint64(0) for _ in range(5)]
        self.engine._check(self.lib.mamcts_search_counters(self.h, *[C.byref(x) for x in v]),
                           "mamcts_search_counters")
        return dict(iterations=v[0].value, rollout_steps=v[1].value, expand_steps=v[2].value,
                    nodes=v[3].value, backup_levels=v[4].value)

    def exact_ucb_evaluations(self) -> int:
        v = C.c_uint64(0)
        self.engine._check(self.lib.mamcts_search_exact_ucb_evaluations(self.h, C.byref(v)),
                           "mamcts_search_exact_ucb_evaluations")
        return v.value

    def root_stats(self, tree: int, agent: int):
        nA = self.num_actions[agent]
        q = (C.c_double * MAX_ACTIONS)()
        n = (C.c_uint32 * MAX_ACTIONS)()
        val = C.c_double(0)
        vis = C.c_uint32(0)
        mask = C.c_uint32(0)
        nn = C.c_uint32(0)
        self.engine._check(self.lib.mamcts_search_root_stats(self.h, tree, agent, C.byref(val),
                                                             C.byref(vis), q, n, C.byref(mask),
                                                             C.byref(nn)), "mamcts_search_root_stats")
        return dict(value=val.value, visits=vis.value, num_nodes=nn.value,
                    q=[q[a] if (mask.value >> a) & 1 else 0.0 for a in range(nA)],
                    n=[int(n[a]) if (mask.value >> a) & 1 else -1 for a in range(nA)])

    def merge_roots(self):
        return self.engine._merge_call(
            lambda *o: self.lib.mamcts_search_merge_roots(self.h, *o), self.num_actions[0],
            "mamcts_search_merge_roots")

    def dump_tree(self, tree: int, capacity: int):
        max_actions = max(self.num_actions)
        stride = 4 + self.A * (3 + 2 * max_actions)
        rec = np.zeros((capacity, stride))
        cnt = C.c_uint32(0)
        self.engine._check(self.lib.mamcts_search_dump_tree(self.h, tree, rec.ctypes.data_as(_abi.c_double_p),
                                                            capacity, C.byref(cnt)),
                           "mamcts_search_dump_tree")
        return rec[:min(cnt.value, capacity)]


def reference_hypothesis_sequence(seed: int, beliefs, skip_iterations: int, n_iterations: int) -> np.ndarray:
    """The hypothesis of every other agent in every iteration as HypothesisBeliefTracker::sample_current_hypothesis
    draws it (reference hypothesis_belief_tracker.h:136-149; mamcts_reference_hypothesis_sequence).
    beliefs: [num_others, num_hypotheses]. Returns uint8 [n_iterations, num_others]."""
    lib = _abi.load()
    b = np.ascontiguousarray(beliefs, dtype=np.float32)
    out = np.zeros((n_iterations, b.shape[0]), dtype=np.uint8)
    rc = lib.mamcts_reference_hypothesis_sequence(seed, b.ctypes.data_as(C.c_void_p), b.shape[0], b.shape[1], skip_iterations,
                                                  n_iterations, out.ctypes.data_as(_abi.c_u8_p))
    if rc != 0:
        raise MamctsError(lib.mamcts_last_error(None).decode())
    return out


class HypSearch(Search):
    """T device-resident hypothesis-MCTS trees (mamcts_hyp_search_create): Mcts<CrossingState<int>, UctStatistic,
    HypothesisStatistic, RandomHeuristic>::search(state, belief_tracker), reference mcts/mcts.h:143-164.
    gaps: [(lo, hi)] desired-gap range per hypothesis; hypothesis_sequence: uint8 [iterations, num_others]."""

    def __init__(self, engine: Engine, mp: Params, hp, cp: CrossingParams, num_trees: int, gaps, hypothesis_sequence,
                 root_x=None, root_last=None, first_tree_id: int = 0, decorrelate_trees: bool = False,
                 max_nodes_per_tree: int = 0):
        from ._abi import HypSearchConfig
        self.engine = engine
        self.lib = engine.lib
        cfg = HypSearchConfig()
        cfg.num_trees = num_trees
        cfg.first_tree_id = first_tree_id
        cfg.max_nodes_per_tree = max_nodes_per_tree
        cfg.num_hypotheses = len(gaps)
        for h, (lo, hi) in enumerate(gaps):
            cfg.gap_lo[h], cfg.gap_hi[h] = lo, hi
        cfg.decorrelate_trees = 1 if decorrelate_trees else 0
        cfg.mcts, cfg.hypothesis, cfg.crossing = mp, hp, cp
        self.cfg = cfg
        self.A = cp.num_other_agents + 1
        self.num_actions = [cp.num_ego_actions] + [cp.num_other_actions] * (self.A - 1)
        self.MA = cp.max_velocity_other - cp.min_velocity_other + 1
        self.H = len(gaps)
        rx, rl = self._root(root_x, root_last)
        seq = np.ascontiguousarray(hypothesis_sequence, dtype=np.uint8)
        if seq.shape != (mp.max_number_of_iterations, self.A - 1):
            raise ValueError("hypothesis_sequence must be [MAX_NUMBER_OF_ITERATIONS, num_others]")
        self.h = C.c_void_p()
        rc = self.lib.mamcts_hyp_search_create(engine.h, C.byref(cfg), rx.ctypes.data_as(_abi.c_i32_p),
                                               rl.ctypes.data_as(_abi.c_i32_p), seq.ctypes.data_as(_abi.c_u8_p),
                                               C.byref(self.h))
        engine._check(rc, "mamcts_hyp_search_create")
        engine._searches.add(self)

    def _root(self, root_x, root_last):
        rx = np.zeros(_abi.MAX_AGENTS, dtype=np.int32)
        rl = np.zeros(_abi.MAX_AGENTS, dtype=np.int32)
        rx[:self.A] = [5] * self.A if root_x is None else list(root_x)
        rl[:self.A] = [2] * self.A if root_last is None else list(root_last)
        return rx, rl

    def reset(self, root_x=None, root_last=None, hypothesis_sequence=None):
        """New decision: root state (None = the resident one) and hypothesis sequence (None = the resident one)."""
        px = pl = ps = None
        if root_x is not None:
            rx, rl = self._root(root_x, root_last)
            px, pl = rx.ctypes.data_as(_abi.c_i32_p), rl.ctypes.data_as(_abi.c_i32_p)
        if hypothesis_sequence is not None:
            seq = np.ascontiguousarray(hypothesis_sequence, dtype=np.uint8)
            ps = seq.ctypes.data_as(_abi.c_u8_p)
        self.engine._check(self.lib.mamcts_hyp_search_reset(self.h, px, pl, ps), "mamcts_hyp_search_reset")

    @property
    def dump_stride(self) -> int:
        NE = self.num_actions[0]
        return 4 + (3 + 2 * NE) + (self.A - 1) * (6 + self.H * (2 + self.MA) + 2 * self.H * self.MA)

    def dump_tree(self, tree: int, capacity: int):
        rec = np.zeros((capacity, self.dump_stride))
        cnt = C.c_uint32(0)
        self.engine._check(self.lib.mamcts_search_dump_tree(self.h, tree, rec.ctypes.data_as(_abi.c_double_p),
                                                            capacity, C.byref(cnt)), "mamcts_search_dump_tree")
        return rec[:min(cnt.value, capacity)]
