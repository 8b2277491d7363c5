"""ctypes access to the CPU checker (oracle/) for tests, smoke() and bench.py's CPU legs.

  Oracle   : oracle/_build/libmamcts_oracle.so - the plain-C restatement (always available;
             built on demand with gcc).
  Reference: oracle/_ref/libmamcts_ref.so      - the UNMODIFIED reference compiled from
             /root/reference (built where the reference checkout exists; the prebuilt .so
             travels to the GPU box). `reference_available()` says whether it can be used.

TEST INFRASTRUCTURE ONLY: nothing under mamcts_b200/ imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from mamcts_b200._abi import (MAX_ACTIONS, MAX_AGENTS, SRC_PHILOX, SRC_REFERENCE_CONST,
                              SRC_REPLAY, ENV_CROSSING_I32, ENV_SIMPLE, CrossingParams, Params,
                              SimpleParams)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libmamcts_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libmamcts_ref.so")
REFERENCE_SRC = os.environ.get("MAMCTS_REFERENCE", "/root/reference")


def _ptr(a, ty=None):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class OrcState(C.Structure):
    _fields_ = [("x", C.c_int32 * MAX_AGENTS), ("last", C.c_int32 * MAX_AGENTS),
                ("terminal", C.c_uint8), ("goal", C.c_uint8), ("collided", C.c_uint8),
                ("pad", C.c_uint8)]


class OrcEnv(C.Structure):
    _fields_ = [("env", C.c_int32), ("num_agents", C.c_uint32),
                ("num_actions", C.c_uint32 * MAX_AGENTS), ("crossing", CrossingParams),
                ("simple", SimpleParams)]


class OrcRolloutResult(C.Structure):
    _fields_ = [("ego_return", C.c_double), ("other_return", C.c_double * MAX_AGENTS),
                ("cost0", C.c_double), ("step_length", C.c_double), ("steps", C.c_uint32),
                ("final_state", OrcState)]


class OrcStat(C.Structure):
    _fields_ = [("value", C.c_double), ("latest_return", C.c_double),
                ("total_visits", C.c_uint32), ("num_actions", C.c_uint32),
                ("num_expanded", C.c_uint32), ("expanded", C.c_uint8 * MAX_ACTIONS),
                ("n", C.c_uint32 * MAX_ACTIONS), ("q", C.c_double * MAX_ACTIONS),
                ("collected_action", C.c_uint32), ("collected_reward", C.c_double)]


class OrcSearchConfig(C.Structure):
    _fields_ = [("env", OrcEnv), ("mcts", Params), ("action_source", C.c_int32),
                ("rollouts_per_leaf", C.c_uint32), ("philox_seed", C.c_uint64),
                ("tree_id", C.c_uint32),
                ("expand_order", (C.c_uint32 * MAX_ACTIONS) * MAX_AGENTS)]


def build_oracle(force: bool = False) -> str:
    src = os.path.join(ORACLE_DIR, "mamcts_oracle.c")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(ORACLE_DIR, "mamcts_oracle.h"))):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])
    return ORACLE_SO


def build_reference(force: bool = False) -> str | None:
    """(Re)builds oracle/_ref when the reference checkout is present; otherwise uses a prebuilt one."""
    if os.path.isdir(os.path.join(REFERENCE_SRC, "mcts")):
        srcs = [os.path.join(ORACLE_DIR, f) for f in ("ref_driver.cc", "mamcts_oracle.c",
                                                       "mamcts_oracle.h")]
        if force or not os.path.exists(REF_SO) or os.path.getmtime(REF_SO) < max(
                os.path.getmtime(s) for s in srcs):
            subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref", f"REF={REFERENCE_SRC}"])
    return REF_SO if os.path.exists(REF_SO) else None


def reference_available() -> bool:
    return build_reference() is not None


_oracle = None
_ref = None


def oracle() -> C.CDLL:
    global _oracle
    if _oracle is None:
        lib = C.CDLL(build_oracle())
        lib.orc_philox_draw.restype = C.c_uint32
        lib.orc_philox_draw.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                        C.c_uint32, C.c_uint32]
        lib.orc_reduce_mean.restype = C.c_double
        lib.orc_reduce_mean.argtypes = [C.c_void_p, C.c_uint32]
        lib.orc_tree_create.restype = C.c_void_p
        lib.orc_tree_create.argtypes = [C.POINTER(OrcSearchConfig), C.POINTER(OrcState)]
        lib.orc_tree_destroy.argtypes = [C.c_void_p]
        lib.orc_tree_run.argtypes = [C.c_void_p, C.c_uint32]
        lib.orc_tree_num_nodes.restype = C.c_uint32
        lib.orc_tree_num_nodes.argtypes = [C.c_void_p]
        lib.orc_tree_rollout_steps.restype = C.c_uint64
        lib.orc_tree_rollout_steps.argtypes = [C.c_void_p]
        lib.orc_tree_root_stat.restype = C.POINTER(OrcStat)
        lib.orc_tree_root_stat.argtypes = [C.c_void_p, C.c_uint32]
        lib.orc_tree_dump.restype = C.c_uint32
        lib.orc_tree_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        lib.orc_stat_choose_next_action.restype = C.c_uint32
        lib.orc_stat_best_action.restype = C.c_uint32
        _oracle = lib
    return _oracle


def reference() -> C.CDLL:
    global _ref
    if _ref is None:
        path = build_reference()
        if path is None:
            raise RuntimeError("oracle/_ref/libmamcts_ref.so not built and no reference checkout")
        _ref = C.CDLL(path)
    return _ref


# ------------------------------------------------------------------------------------------------
# environments
# ------------------------------------------------------------------------------------------------

def make_env(cp: CrossingParams | None = None, sp: SimpleParams | None = None) -> OrcEnv:
    env = OrcEnv()
    if cp is not None:
        oracle().orc_env_crossing(C.byref(env), C.byref(cp))
    else:
        oracle().orc_env_simple(C.byref(env), C.byref(sp))
    return env


def make_state(x, last=None, terminal=False) -> OrcState:
    s = OrcState()
    for i, v in enumerate(x):
        s.x[i] = int(v)
        s.last[i] = int(last[i]) if last is not None else 2
    s.terminal = 1 if terminal else 0
    return s


def philox4x32_10(ctr, key):
    out = (C.c_uint32 * 4)()
    oracle().orc_philox4x32_10((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
    return list(out)


def philox_draw(seed, hi, lo, A, step, agent, n) -> int:
    return int(oracle().orc_philox_draw(seed, hi, lo, A, step, agent, n))


def reduce_mean(vals) -> float:
    v = np.ascontiguousarray(vals, dtype=np.float64)
    return float(oracle().orc_reduce_mean(_ptr(v), len(v)))


def crossing_execute(cp: CrossingParams, x, last, action):
    """Oracle CrossingState<int>::execute on one state. action[0]: ego index; [1:]: other values."""
    s = make_state(x, last)
    nxt = OrcState()
    r = C.c_double()
    c0 = C.c_double()
    act = (C.c_int32 * MAX_AGENTS)(*[int(a) for a in action])
    oracle().orc_crossing_execute(C.byref(cp), C.byref(s), act, C.byref(nxt), C.byref(r),
                                  C.byref(c0))
    A = cp.num_other_agents + 1
    flags = (1 if nxt.terminal else 0) | (2 if nxt.goal else 0) | (4 if nxt.collided else 0)
    return list(nxt.x[:A]), list(nxt.last[:A]), flags, r.value, c0.value


def crossing_policy_action(cp: CrossingParams, agent_x, agent_last, ego_x, ego_last, gap) -> int:
    """Oracle AgentPolicyCrossingState<int>::calculate_action."""
    f = oracle().orc_crossing_policy_action
    f.restype = C.c_int32
    return int(f(C.byref(cp), C.c_int32(agent_x), C.c_int32(agent_last), C.c_int32(ego_x), C.c_int32(ego_last),
                 C.c_int32(gap)))


def rollout(env: OrcEnv, mp: Params, x, last=None, terminal=False, depth=0, kind=SRC_PHILOX,
            replay=None, const_actions=None, seed=0, stream_hi=0, stream_lo=0, trace_cap=0,
            gap_lo=None, gap_hi=None):
    """One oracle rollout; returns a dict."""
    st = make_state(x, last, terminal)
    res = OrcRolloutResult()
    rp = None
    nrep = 0
    if replay is not None:
        rep = np.ascontiguousarray(replay, dtype=np.int8)
        nrep = rep.shape[0]
        rp = _ptr(rep)
    ca = (C.c_uint32 * MAX_AGENTS)(*(const_actions or []))
    trace = np.zeros(max(trace_cap, 1), dtype=np.float64)
    glo = (C.c_int32 * MAX_AGENTS)(*[int(v) for v in gap_lo]) if gap_lo is not None else None
    ghi = (C.c_int32 * MAX_AGENTS)(*[int(v) for v in gap_hi]) if gap_hi is not None else None
    oracle().orc_rollout(C.byref(env), C.byref(mp), C.byref(st), C.c_uint32(depth),
                         C.c_int32(kind), rp, C.c_uint32(nrep), ca, C.c_uint64(seed),
                         C.c_uint32(stream_hi), C.c_uint32(stream_lo), glo, ghi, C.byref(res),
                         _ptr(trace) if trace_cap else None, C.c_uint32(trace_cap))
    A = env.num_agents
    fs = res.final_state
    return dict(ego_return=res.ego_return, other_return=list(res.other_return[:A - 1]),
                cost0=res.cost0, step_length=res.step_length, steps=res.steps,
                final_x=list(fs.x[:A]), final_last=list(fs.last[:A]),
                final_flags=(1 if fs.terminal else 0) | (2 if fs.goal else 0) |
                (4 if fs.collided else 0),
                reward_trace=trace[:min(trace_cap, res.steps)].copy())


def rollout_batch(env: OrcEnv, mp: Params, x, last=None, flags=None, depth=None, K=1,
                  kind=SRC_PHILOX, replay=None, replay_steps=None, const_actions=None, seed=0,
                  stream_hi=None, stream_hi_base=0, stream_lo_base=0, stream_lo=None, trace_stride=0,
                  hyp_gap_lo=None, hyp_gap_hi=None):
    """Oracle counterpart of mamcts_rollout_* over n leaves x K rollouts (same output layout)."""
    x = np.asarray(x, dtype=np.int32)
    A = env.num_agents
    n = x.shape[1]
    out = dict(ego_return=np.zeros(n), other_return=np.zeros((max(A - 1, 0), n)),
               ego_cost=np.zeros(n), step_length=np.zeros(n), total_steps=0,
               steps=np.zeros(n * K, dtype=np.uint32), final_x=np.zeros((A, n * K), dtype=np.int32),
               final_last_action=np.zeros((A, n * K), dtype=np.int32),
               final_flags=np.zeros(n * K, dtype=np.uint8),
               rollout_ego_return=np.zeros(n * K),
               reward_trace=np.zeros((n * K, max(trace_stride, 1))))
    for i in range(n):
        ego, cost, sl = np.zeros(K), np.zeros(K), np.zeros(K)
        oth = np.zeros((max(A - 1, 1), K))
        for k in range(K):
            hi = int(stream_hi[i]) if stream_hi is not None else stream_hi_base
            lo = ((int(stream_lo[i]) + k) if stream_lo is not None
                  else (stream_lo_base + i * K + k)) & 0xFFFFFFFF
            rep = None
            if kind == SRC_REPLAY:
                rep = np.asarray(replay[i][:int(replay_steps[i])], dtype=np.int8).reshape(-1, A)
            r = rollout(env, mp, x[:, i], None if last is None else np.asarray(last)[:, i],
                        bool(flags[i] & 1) if flags is not None else False,
                        int(depth[i]) if depth is not None else 0, kind, rep, const_actions, seed,
                        hi, lo, trace_stride,
                        None if hyp_gap_lo is None else np.asarray(hyp_gap_lo)[:, i],
                        None if hyp_gap_hi is None else np.asarray(hyp_gap_hi)[:, i])
            idx = i * K + k
            ego[k], cost[k], sl[k] = r["ego_return"], r["cost0"], r["step_length"]
            for j in range(A - 1):
                oth[j, k] = r["other_return"][j]
            out["steps"][idx] = r["steps"]
            out["final_x"][:, idx] = r["final_x"]
            out["final_last_action"][:, idx] = r["final_last"]
            out["final_flags"][idx] = r["final_flags"]
            out["rollout_ego_return"][idx] = r["ego_return"]
            out["total_steps"] += r["steps"]
            if trace_stride:
                out["reward_trace"][idx, :len(r["reward_trace"])] = r["reward_trace"]
        out["ego_return"][i] = reduce_mean(ego)
        out["ego_cost"][i] = reduce_mean(cost)
        out["step_length"][i] = reduce_mean(sl)
        for j in range(A - 1):
            out["other_return"][j, i] = reduce_mean(oth[j])
    return out


class OrcActionValuesResult(C.Structure):
    _fields_ = [("action_return", C.c_double * MAX_ACTIONS), ("action_cost", C.c_double * MAX_ACTIONS),
                ("action_len", C.c_double * MAX_ACTIONS), ("other_return", C.c_double * MAX_AGENTS),
                ("mean_cost", C.c_double), ("steps", C.c_uint32)]


def action_values_batch(env: OrcEnv, mp: Params, x, last=None, flags=None, depth=None, kind=SRC_PHILOX,
                        const_actions=None, seed=0, stream_hi=None, stream_hi_base=0, stream_lo_base=0,
                        stream_lo=None):
    """Oracle counterpart of mamcts_action_values_crossing_i32 (ActionValueRandomHeuristic's per-ego-action
    first step + rollout, action_value_random_heuristic.h:37-85) over n leaves; same output layout."""
    x = np.asarray(x, dtype=np.int32)
    A, nA = env.num_agents, env.num_actions[0]
    n = x.shape[1]
    out = dict(action_return=np.zeros((n, nA)), action_cost=np.zeros((n, nA)), action_step_length=np.zeros((n, nA)),
               other_return=np.zeros((max(A - 1, 0), n)), mean_cost=np.zeros(n), total_steps=0)
    ca = (C.c_uint32 * MAX_AGENTS)(*(const_actions or []))
    for i in range(n):
        st = make_state(x[:, i], None if last is None else np.asarray(last)[:, i],
                        bool(flags[i] & 1) if flags is not None else False)
        hi = int(stream_hi[i]) if stream_hi is not None else stream_hi_base
        lo = (int(stream_lo[i]) if stream_lo is not None else (stream_lo_base + i)) & 0xFFFFFFFF
        res = OrcActionValuesResult()
        oracle().orc_action_values(C.byref(env), C.byref(mp), C.byref(st), C.c_uint32(int(depth[i]) if depth is not None else 0),
                                   C.c_int32(kind), ca, C.c_uint64(seed), C.c_uint32(hi), C.c_uint32(lo), C.byref(res))
        out["action_return"][i] = res.action_return[:nA]
        out["action_cost"][i] = res.action_cost[:nA]
        out["action_step_length"][i] = res.action_len[:nA]
        out["other_return"][:, i] = res.other_return[:A - 1]
        out["mean_cost"][i] = res.mean_cost
        out["total_steps"] += res.steps
    return out


def ref_action_values_crossing(mp: Params, cp: CrossingParams, x, last=None, flags=None, depth=None):
    """The reference's own ActionValueRandomHeuristic::calculate_heuristic_values (with the two test-only
    statistic adapters of oracle/ref_driver.cc) over n leaves."""
    x = np.ascontiguousarray(x, dtype=np.int32)
    A = cp.num_other_agents + 1
    n = x.shape[1]
    nA = cp.num_ego_actions
    la = None if last is None else np.ascontiguousarray(last, dtype=np.int32)
    fl = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
    de = None if depth is None else np.ascontiguousarray(depth, dtype=np.uint32)
    ret, cost = np.zeros((n, nA)), np.zeros((n, nA))
    oth, mean = np.zeros((max(A - 1, 1), n)), np.zeros(n)
    reference().ref_action_values_crossing(C.byref(mp), C.byref(cp), C.c_uint32(n), _ptr(x), _ptr(la) if la is not None else None,
                                           _ptr(fl) if fl is not None else None, _ptr(de) if de is not None else None,
                                           C.c_uint32(nA), _ptr(ret), _ptr(cost), _ptr(oth), _ptr(mean))
    return dict(action_return=ret, action_cost=cost, other_return=oth[:A - 1], mean_cost=mean)


# ------------------------------------------------------------------------------------------------
# search (oracle restatement)
# ------------------------------------------------------------------------------------------------

def expand_order_libstdcxx(seed: int, n: int):
    """The widening order the product computes on the host (same libstdc++ algorithm); used by the
    oracle so that it does not depend on the product: asks the reference when available, else the
    golden table in tests/golden/expand_order.json."""
    if reference_available():
        from mamcts_b200.params import default_params
        mp = default_params(random_seed=seed)
        order = (C.c_uint32 * n)()
        reference().ref_uct_expand_order(C.byref(mp), C.c_uint32(n), order)
        return list(order)
    import json
    with open(os.path.join(ROOT, "tests", "golden", "expand_order.json")) as f:
        table = json.load(f)
    return table[str(seed)][str(n)]


class OracleTree:
    def __init__(self, env: OrcEnv, mp: Params, root_x, root_last=None, action_source=SRC_PHILOX,
                 K=1, seed=0, tree_id=0, expand_orders=None):
        cfg = OrcSearchConfig()
        cfg.env = env
        cfg.mcts = mp
        cfg.action_source = action_source
        cfg.rollouts_per_leaf = K
        cfg.philox_seed = seed
        cfg.tree_id = tree_id
        self.env = env
        for a in range(env.num_agents):
            nA = env.num_actions[a]
            order = (expand_orders[a] if expand_orders is not None
                     else expand_order_libstdcxx(mp.random_seed, nA))
            for k in range(nA):
                cfg.expand_order[a][k] = order[k]
        st = make_state(root_x, root_last)
        self.cfg = cfg
        self.h = oracle().orc_tree_create(C.byref(cfg), C.byref(st))

    def run(self, iterations: int):
        oracle().orc_tree_run(self.h, iterations)

    @property
    def num_nodes(self) -> int:
        return int(oracle().orc_tree_num_nodes(self.h))

    @property
    def rollout_steps(self) -> int:
        return int(oracle().orc_tree_rollout_steps(self.h))

    def root_stat(self, agent: int):
        s = oracle().orc_tree_root_stat(self.h, agent).contents
        nA = s.num_actions
        return dict(value=s.value, latest=s.latest_return, visits=s.total_visits,
                    n=[int(s.n[a]) if s.expanded[a] else -1 for a in range(nA)],
                    q=[s.q[a] if s.expanded[a] else 0.0 for a in range(nA)])

    def dump(self, max_actions: int, capacity: int | None = None):
        cap = capacity or self.num_nodes
        stride = 4 + self.env.num_agents * (3 + 2 * max_actions)
        rec = np.zeros((cap, stride))
        cnt = oracle().orc_tree_dump(self.h, _ptr(rec), cap, max_actions)
        return rec[:min(cnt, cap)]

    def close(self):
        if self.h:
            oracle().orc_tree_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# the compiled reference
# ------------------------------------------------------------------------------------------------

def ref_expand_order(mp: Params, n: int):
    order = (C.c_uint32 * n)()
    reference().ref_uct_expand_order(C.byref(mp), C.c_uint32(n), order)
    return list(order)


def ref_rollout_crossing(mp: Params, cp: CrossingParams, x, last=None, flags=None, depth=None,
                         rec_stride=64):
    x = np.ascontiguousarray(x, dtype=np.int32)
    A, n = x.shape
    last_a = None if last is None else np.ascontiguousarray(last, dtype=np.int32)
    flags_a = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
    depth_a = None if depth is None else np.ascontiguousarray(depth, dtype=np.uint32)
    out = dict(ego_return=np.zeros(n), other_return=np.zeros((A - 1, n)), ego_cost=np.zeros(n),
               step_length=np.zeros(n), steps=np.zeros(n, dtype=np.uint32),
               final_x=np.zeros((A, n), dtype=np.int32),
               final_last_action=np.zeros((A, n), dtype=np.int32),
               final_flags=np.zeros(n, dtype=np.uint8),
               rec_actions=np.zeros((n, rec_stride, A), dtype=np.int8),
               rec_rewards=np.zeros((n, rec_stride)))
    reference().ref_rollout_crossing(
        C.byref(mp), C.byref(cp), C.c_uint32(n), _ptr(x), _ptr(last_a), _ptr(flags_a),
        _ptr(depth_a), _ptr(out["ego_return"]), _ptr(out["other_return"]), _ptr(out["ego_cost"]),
        _ptr(out["step_length"]), _ptr(out["steps"]), _ptr(out["final_x"]),
        _ptr(out["final_last_action"]), _ptr(out["final_flags"]), _ptr(out["rec_actions"]),
        _ptr(out["rec_rewards"]), C.c_uint32(rec_stride))
    return out


# ---- CrossingState<float> -----------------------------------------------------------------------

class OrcStateF32(C.Structure):
    _fields_ = [("x", C.c_float * MAX_AGENTS), ("last", C.c_float * MAX_AGENTS), ("terminal", C.c_uint8),
                ("goal", C.c_uint8), ("collided", C.c_uint8), ("pad", C.c_uint8)]


class OrcRolloutResultF32(C.Structure):
    _fields_ = [("ego_return", C.c_double), ("other_return", C.c_double * MAX_AGENTS), ("cost0", C.c_double),
                ("step_length", C.c_double), ("steps", C.c_uint32), ("final_state", OrcStateF32)]


def _state_f32(x, last=None, terminal=False):
    s = OrcStateF32()
    for a, v in enumerate(x):
        s.x[a] = float(v)
        s.last[a] = 2.0 if last is None else float(last[a])
    s.terminal = 1 if terminal else 0
    return s


def crossing_execute_f32(cp, x, last, action):
    """Oracle CrossingState<float>::execute. action[0]: ego index; [1:]: other velocities (float)."""
    s = _state_f32(x, last)
    nxt = OrcStateF32()
    r, c0 = C.c_double(), C.c_double()
    act = (C.c_float * MAX_AGENTS)(*[float(a) for a in action])
    oracle().orc_crossing_f32_execute(C.byref(cp), C.byref(s), act, C.byref(nxt), C.byref(r), C.byref(c0))
    A = cp.num_other_agents + 1
    flags = (1 if nxt.terminal else 0) | (2 if nxt.goal else 0) | (4 if nxt.collided else 0)
    return (np.array(nxt.x[:A], dtype=np.float32), np.array(nxt.last[:A], dtype=np.float32), flags, r.value, c0.value)


def rollout_f32(cp, mp: Params, x, last=None, terminal=False, depth=0, kind=SRC_PHILOX, replay=None,
                const_actions=None, seed=0, stream_hi=0, stream_lo=0):
    """One oracle rollout on CrossingState<float>."""
    st = _state_f32(x, last, terminal)
    res = OrcRolloutResultF32()
    rp, nrep = None, 0
    if replay is not None:
        rep_a = np.ascontiguousarray(replay, dtype=np.float32)
        nrep, rp = rep_a.shape[0], _ptr(rep_a)
    ca = (C.c_uint32 * MAX_AGENTS)(*(const_actions or []))
    oracle().orc_rollout_f32(C.byref(cp), C.byref(mp), C.byref(st), C.c_uint32(depth), C.c_int32(kind), rp,
                             C.c_uint32(nrep), ca, C.c_uint64(seed), C.c_uint32(stream_hi), C.c_uint32(stream_lo),
                             C.byref(res))
    A = cp.num_other_agents + 1
    fs = res.final_state
    return dict(ego_return=res.ego_return, other_return=list(res.other_return[:A - 1]), cost0=res.cost0,
                step_length=res.step_length, steps=res.steps,
                final_x=np.array(fs.x[:A], dtype=np.float32), final_last=np.array(fs.last[:A], dtype=np.float32),
                final_flags=(1 if fs.terminal else 0) | (2 if fs.goal else 0) | (4 if fs.collided else 0))


def crossing_policy_action_f32(cp, agent_x, agent_last, ego_x, ego_last, gap) -> float:
    f = oracle().orc_crossing_policy_action_f32
    f.restype = C.c_float
    f.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]
    return float(f(C.byref(cp), float(agent_x), float(agent_last), float(ego_x), float(ego_last), float(gap)))


def rollout_f32_hypothesis(cp, mp: Params, x, last=None, depth=0, seed=0, stream_hi=0, stream_lo=0, gap_lo=None, gap_hi=None):
    """One oracle hypothesis rollout on CrossingState<float> (Philox gap draws)."""
    st = _state_f32(x, last, False)
    res = OrcRolloutResultF32()
    glo = (C.c_float * MAX_AGENTS)(*[float(v) for v in gap_lo])
    ghi = (C.c_float * MAX_AGENTS)(*[float(v) for v in gap_hi])
    oracle().orc_rollout_f32_hypothesis(C.byref(cp), C.byref(mp), C.byref(st), C.c_uint32(depth), C.c_uint64(seed),
                                        C.c_uint32(stream_hi), C.c_uint32(stream_lo), glo, ghi, C.byref(res))
    A = cp.num_other_agents + 1
    fs = res.final_state
    return dict(ego_return=res.ego_return, cost0=res.cost0, step_length=res.step_length, steps=res.steps,
                final_x=np.array(fs.x[:A], dtype=np.float32), final_last=np.array(fs.last[:A], dtype=np.float32),
                final_flags=(1 if fs.terminal else 0) | (2 if fs.goal else 0) | (4 if fs.collided else 0))


def ref_policy_calculate_action_f32(cp, agent_x, agent_last, ego_x, ego_last, gap):
    arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (agent_x, agent_last, ego_x, ego_last, gap)]
    out = np.zeros(arrs[0].shape[0], dtype=np.float32)
    reference().ref_policy_calculate_action_f32(C.byref(cp), C.c_uint32(out.shape[0]), *[_ptr(a) for a in arrs], _ptr(out))
    return out


def ref_rollout_crossing_f32_hypothesis(mp: Params, cp, x, last, hyp_lo, hyp_hi, cur_hyp, depth=None, rec_stride=64):
    """The reference's float-domain hypothesis rollout (RandomHeuristic::rollout<.., UctStatistic, HypothesisStatistic, ..>
    over CrossingState<float>, other agents by AgentPolicyCrossingState<float>::act), recorded for REPLAY."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    A, n = x.shape
    last_a = None if last is None else np.ascontiguousarray(last, dtype=np.float32)
    depth_a = None if depth is None else np.ascontiguousarray(depth, dtype=np.uint32)
    lo, hi = np.ascontiguousarray(hyp_lo, dtype=np.float32), np.ascontiguousarray(hyp_hi, dtype=np.float32)
    ch = np.ascontiguousarray(cur_hyp, dtype=np.uint32)
    out = dict(ego_return=np.zeros(n), ego_cost=np.zeros(n), step_length=np.zeros(n),
               steps=np.zeros(n, dtype=np.uint32), final_x=np.zeros((A, n), dtype=np.float32),
               final_last_action=np.zeros((A, n), dtype=np.float32), final_flags=np.zeros(n, dtype=np.uint8),
               rec_actions=np.zeros((n, rec_stride, A), dtype=np.float32))
    reference().ref_rollout_crossing_f32_hypothesis(
        C.byref(mp), C.byref(cp), C.c_uint32(n), _ptr(x), _ptr(last_a) if last_a is not None else None,
        _ptr(depth_a) if depth_a is not None else None, C.c_uint32(lo.shape[0]), _ptr(lo), _ptr(hi), _ptr(ch),
        _ptr(out["ego_return"]), _ptr(out["ego_cost"]), _ptr(out["step_length"]), _ptr(out["steps"]), _ptr(out["final_x"]),
        _ptr(out["final_last_action"]), _ptr(out["final_flags"]), _ptr(out["rec_actions"]), C.c_uint32(rec_stride))
    return out


def ref_rollout_crossing_f32(mp: Params, cp, x, last=None, depth=None, rec_stride=64):
    """The reference's RandomHeuristic::rollout on CrossingState<float> (UCT statistics), recorded."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    A, n = x.shape
    last_a = None if last is None else np.ascontiguousarray(last, dtype=np.float32)
    depth_a = None if depth is None else np.ascontiguousarray(depth, dtype=np.uint32)
    out = dict(ego_return=np.zeros(n), ego_cost=np.zeros(n), step_length=np.zeros(n),
               steps=np.zeros(n, dtype=np.uint32), final_x=np.zeros((A, n), dtype=np.float32),
               final_last_action=np.zeros((A, n), dtype=np.float32), final_flags=np.zeros(n, dtype=np.uint8),
               rec_actions=np.zeros((n, rec_stride, A), dtype=np.float32))
    reference().ref_rollout_crossing_f32(
        C.byref(mp), C.byref(cp), C.c_uint32(n), _ptr(x), _ptr(last_a), _ptr(depth_a), _ptr(out["ego_return"]),
        _ptr(out["ego_cost"]), _ptr(out["step_length"]), _ptr(out["steps"]), _ptr(out["final_x"]),
        _ptr(out["final_last_action"]), _ptr(out["final_flags"]), _ptr(out["rec_actions"]), C.c_uint32(rec_stride))
    return out


def ref_crossing_execute_f32(cp, x, last, action):
    x = np.ascontiguousarray(x, dtype=np.float32)
    A, n = x.shape
    last = np.ascontiguousarray(last, dtype=np.float32)
    action = np.ascontiguousarray(action, dtype=np.float32)
    nx, nl = np.zeros((A, n), dtype=np.float32), np.zeros((A, n), dtype=np.float32)
    fl, r, c0 = np.zeros(n, dtype=np.uint8), np.zeros(n), np.zeros(n)
    reference().ref_crossing_execute_f32(C.byref(cp), C.c_uint32(n), _ptr(x), _ptr(last), _ptr(action), _ptr(nx),
                                         _ptr(nl), _ptr(fl), _ptr(r), _ptr(c0))
    return nx, nl, fl, r, c0


def ref_policy_calculate_action(cp: CrossingParams, agent_x, agent_last, ego_x, ego_last, gap):
    """The reference's AgentPolicyCrossingState<int>::calculate_action on n inputs."""
    arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in (agent_x, agent_last, ego_x, ego_last, gap)]
    out = np.zeros(arrs[0].shape[0], dtype=np.int32)
    reference().ref_policy_calculate_action(C.byref(cp), C.c_uint32(out.shape[0]), *[_ptr(a) for a in arrs],
                                            _ptr(out))
    return out


def ref_rollout_crossing_hypothesis(mp: Params, cp: CrossingParams, x, last, hyp_lo, hyp_hi, cur_hyp,
                                    depth=None, rec_stride=64):
    """RandomHeuristic::rollout<.., UctStatistic, HypothesisStatistic, ..> on n start states; other agent
    j of state i follows hypothesis cur_hyp[j, i] of {[hyp_lo[h], hyp_hi[h]]}. Records the joint actions."""
    x = np.ascontiguousarray(x, dtype=np.int32)
    A, n = x.shape
    last_a = None if last is None else np.ascontiguousarray(last, dtype=np.int32)
    depth_a = None if depth is None else np.ascontiguousarray(depth, dtype=np.uint32)
    lo = np.ascontiguousarray(hyp_lo, dtype=np.int32)
    hi = np.ascontiguousarray(hyp_hi, dtype=np.int32)
    cur = np.ascontiguousarray(cur_hyp, dtype=np.uint32)
    out = dict(ego_return=np.zeros(n), other_return=np.zeros((A - 1, n)), ego_cost=np.zeros(n),
               step_length=np.zeros(n), steps=np.zeros(n, dtype=np.uint32),
               final_x=np.zeros((A, n), dtype=np.int32),
               final_last_action=np.zeros((A, n), dtype=np.int32),
               final_flags=np.zeros(n, dtype=np.uint8),
               rec_actions=np.zeros((n, rec_stride, A), dtype=np.int8))
    reference().ref_rollout_crossing_hypothesis(
        C.byref(mp), C.byref(cp), C.c_uint32(n), _ptr(x), _ptr(last_a), _ptr(depth_a),
        C.c_uint32(lo.shape[0]), _ptr(lo), _ptr(hi), _ptr(cur), _ptr(out["ego_return"]),
        _ptr(out["other_return"]), _ptr(out["ego_cost"]), _ptr(out["step_length"]), _ptr(out["steps"]),
        _ptr(out["final_x"]), _ptr(out["final_last_action"]), _ptr(out["final_flags"]),
        _ptr(out["rec_actions"]), C.c_uint32(rec_stride))
    return out


def _hyp_arrays(leaf_node, leaf_has, leaf_cost, path_offset, path_len, edge_node, edge_hyp, edge_action, edge_cost):
    return (np.ascontiguousarray(leaf_node, dtype=np.uint32), np.ascontiguousarray(leaf_has, dtype=np.uint8),
            np.ascontiguousarray(leaf_cost, dtype=np.float64), np.ascontiguousarray(path_offset, dtype=np.uint32),
            np.ascontiguousarray(path_len, dtype=np.uint32), np.ascontiguousarray(edge_node, dtype=np.uint32),
            np.ascontiguousarray(edge_hyp, dtype=np.uint8), np.ascontiguousarray(edge_action, dtype=np.uint8),
            np.ascontiguousarray(edge_cost, dtype=np.float64))


def backup_hypothesis(mp: Params, num_others, stats: dict, leaf_node, leaf_has, leaf_cost, path_offset,
                      path_len, edge_node, edge_hyp, edge_action, edge_cost, chance=False):
    """Oracle HypothesisStatistic back-propagation on mamcts_hyp_stats-shaped numpy arrays (in place)."""
    from mamcts_b200.engine import Engine
    ln, lh, lc, po, pl, en, eh, ea, ec = _hyp_arrays(leaf_node, leaf_has, leaf_cost, path_offset, path_len,
                                                     edge_node, edge_hyp, edge_action, edge_cost)
    st = Engine.hyp_stats_struct(stats)
    oracle().orc_backup_hypothesis(C.c_double(mp.discount_factor), C.c_uint32(num_others), C.c_uint32(ln.shape[0]),
                                   _ptr(ln), _ptr(lh), _ptr(lc), C.c_uint32(lc.shape[1]), _ptr(po), _ptr(pl),
                                   _ptr(en), _ptr(eh), _ptr(ea), _ptr(ec), C.c_int32(1 if chance else 0),
                                   C.byref(st))


def ref_backup_hypothesis(mp: Params, num_others, n_nodes, stats: dict, leaf_node, leaf_has, leaf_cost,
                          path_offset, path_len, edge_node, edge_hyp, edge_action, edge_cost, chance=False):
    """The same arrays executed by REAL HypothesisStatistic objects of the reference (fresh statistics);
    results written into `stats`."""
    from mamcts_b200.engine import Engine
    ln, lh, lc, po, pl, en, eh, ea, ec = _hyp_arrays(leaf_node, leaf_has, leaf_cost, path_offset, path_len,
                                                     edge_node, edge_hyp, edge_action, edge_cost)
    st = Engine.hyp_stats_struct(stats)
    reference().ref_backup_hypothesis(C.byref(mp), C.c_uint32(num_others), C.c_uint32(n_nodes),
                                      C.c_uint32(ln.shape[0]), _ptr(ln), _ptr(lh), _ptr(lc),
                                      C.c_uint32(lc.shape[1]), _ptr(po), _ptr(pl), _ptr(en), _ptr(eh), _ptr(ea),
                                      _ptr(ec), C.c_int32(1 if chance else 0), C.c_uint32(0), C.byref(st))


def ref_rollout_simple(mp: Params, length, depth=None, rec_stride=64):
    length = np.ascontiguousarray(length, dtype=np.int32)
    n = length.shape[0]
    depth_a = None if depth is None else np.ascontiguousarray(depth, dtype=np.uint32)
    out = dict(ego_return=np.zeros(n), other_return=np.zeros((1, n)), step_length=np.zeros(n),
               steps=np.zeros(n, dtype=np.uint32), final_len=np.zeros(n, dtype=np.int32),
               rec_actions=np.zeros((n, rec_stride, 2), dtype=np.int8),
               rec_rewards=np.zeros((n, rec_stride)))
    reference().ref_rollout_simple(
        C.byref(mp), C.c_uint32(n), _ptr(length), _ptr(depth_a), _ptr(out["ego_return"]),
        _ptr(out["other_return"]), _ptr(out["step_length"]), _ptr(out["steps"]),
        _ptr(out["final_len"]), _ptr(out["rec_actions"]), _ptr(out["rec_rewards"]),
        C.c_uint32(rec_stride))
    return out


def ref_crossing_execute(cp: CrossingParams, x, last, action):
    x = np.ascontiguousarray(x, dtype=np.int32)
    A, n = x.shape
    last = np.ascontiguousarray(last, dtype=np.int32)
    action = np.ascontiguousarray(action, dtype=np.int32)
    nx = np.zeros((A, n), dtype=np.int32)
    nl = np.zeros((A, n), dtype=np.int32)
    fl = np.zeros(n, dtype=np.uint8)
    r = np.zeros(n)
    c0 = np.zeros(n)
    reference().ref_crossing_execute(C.byref(cp), C.c_uint32(n), _ptr(x), _ptr(last), _ptr(action),
                                     _ptr(nx), _ptr(nl), _ptr(fl), _ptr(r), _ptr(c0))
    return nx, nl, fl, r, c0


def _search_outputs(A, max_actions, tree_capacity):
    return dict(value=np.zeros(A), latest=np.zeros(A), visits=np.zeros(A, dtype=np.uint32),
                q=np.zeros((A, max_actions)), n=np.zeros((A, max_actions), dtype=np.int64),
                counters=np.zeros(3, dtype=np.uint32),
                tree=np.zeros((max(tree_capacity, 1), 4 + A * (3 + 2 * max_actions))),
                tree_count=C.c_uint32(0))


def _finish_search(o, tree_capacity):
    cnt = o.pop("tree_count").value
    o["tree"] = o["tree"][:min(cnt, tree_capacity)]
    o["tree_count"] = cnt
    o["num_nodes"], o["num_iterations"], o["best_action"] = (int(v) for v in o.pop("counters"))
    return o


def ref_search_crossing(mp: Params, cp: CrossingParams, root_x=None, root_last=None,
                        max_actions=7, tree_capacity=0):
    A = cp.num_other_agents + 1
    rx = np.ascontiguousarray(root_x if root_x is not None else [5] * A, dtype=np.int32)
    rl = np.ascontiguousarray(root_last if root_last is not None else [2] * A, dtype=np.int32)
    o = _search_outputs(A, max_actions, tree_capacity)
    reference().ref_search_crossing(
        C.byref(mp), C.byref(cp), _ptr(rx), _ptr(rl), C.c_uint32(max_actions), _ptr(o["value"]),
        _ptr(o["latest"]), _ptr(o["visits"]), _ptr(o["q"]), _ptr(o["n"]), _ptr(o["counters"]),
        _ptr(o["tree"]) if tree_capacity else None, C.c_uint32(tree_capacity),
        C.byref(o["tree_count"]))
    return _finish_search(o, tree_capacity)


def ref_search_simple(mp: Params, state_length=4, max_actions=2, tree_capacity=0):
    o = _search_outputs(2, max_actions, tree_capacity)
    reference().ref_search_simple(
        C.byref(mp), C.c_int32(state_length), C.c_uint32(max_actions), _ptr(o["value"]),
        _ptr(o["latest"]), _ptr(o["visits"]), _ptr(o["q"]), _ptr(o["n"]), _ptr(o["counters"]),
        _ptr(o["tree"]) if tree_capacity else None, C.c_uint32(tree_capacity),
        C.byref(o["tree_count"]))
    return _finish_search(o, tree_capacity)


def ref_search_philox(env_kind, mp: Params, cp=None, sp=None, root_x=None, root_last=None, seed=0,
                      tree_id=0, K=1, max_actions=7, tree_capacity=0):
    A = (cp.num_other_agents + 1) if env_kind == ENV_CROSSING_I32 else 2
    rx = np.ascontiguousarray(root_x if root_x is not None else [5] * A, dtype=np.int32)
    rl = np.ascontiguousarray(root_last if root_last is not None else [2] * A, dtype=np.int32)
    o = _search_outputs(A, max_actions, tree_capacity)
    steps = C.c_uint64(0)
    reference().ref_search_philox(
        C.c_int32(env_kind), C.byref(mp), C.byref(cp) if cp is not None else None,
        C.byref(sp) if sp is not None else None, _ptr(rx), _ptr(rl), C.c_uint64(seed),
        C.c_uint32(tree_id), C.c_uint32(K), C.c_uint32(max_actions), _ptr(o["value"]),
        _ptr(o["latest"]), _ptr(o["visits"]), _ptr(o["q"]), _ptr(o["n"]), _ptr(o["counters"]),
        C.byref(steps), _ptr(o["tree"]) if tree_capacity else None, C.c_uint32(tree_capacity),
        C.byref(o["tree_count"]))
    o = _finish_search(o, tree_capacity)
    o["rollout_steps"] = steps.value
    return o


def ref_merge_uct(mp: Params, q, cnt, visits, num_actions):
    q = np.ascontiguousarray(q, dtype=np.float64)
    cnt = np.ascontiguousarray(cnt, dtype=np.int64)
    visits = np.ascontiguousarray(visits, dtype=np.uint32)
    n, max_actions = q.shape
    mq = np.zeros(max_actions)
    mn = np.zeros(max_actions, dtype=np.int64)
    mv = C.c_uint32(0)
    val = C.c_double(0)
    best = C.c_uint32(0)
    reference().ref_merge_uct(C.byref(mp), C.c_uint32(n), C.c_uint32(num_actions),
                              C.c_uint32(max_actions), _ptr(q), _ptr(cnt), _ptr(visits), _ptr(mq),
                              _ptr(mn), C.byref(mv), C.byref(val), C.byref(best))
    return dict(q=mq, n=mn, visits=mv.value, value=val.value, best=best.value)


def hyp_dump_stride(cp: CrossingParams, num_hyp: int) -> int:
    NE = cp.num_ego_actions
    MA = cp.max_velocity_other - cp.min_velocity_other + 1
    return 4 + (3 + 2 * NE) + cp.num_other_agents * (6 + num_hyp * (2 + MA) + 2 * num_hyp * MA)


def ref_search_hypothesis(mp: Params, hp, cp: CrossingParams, gaps, root_x=None, root_last=None, skip=0,
                          tree_capacity=0):
    """The reference's own hypothesis-MCTS: Mcts<CrossingState<int>, UctStatistic, HypothesisStatistic,
    RandomHeuristic>::search(state, belief_tracker) (oracle/ref_driver.cc: ref_search_hypothesis).
    gaps: [(lo, hi)] per hypothesis. Returns the hypothesis sequence the tracker handed out, the canonical
    tree dump and the counters."""
    A = cp.num_other_agents + 1
    H = len(gaps)
    lo = np.ascontiguousarray([g[0] for g in gaps], dtype=np.int32)
    hi = np.ascontiguousarray([g[1] for g in gaps], dtype=np.int32)
    rx = np.ascontiguousarray(root_x if root_x is not None else [5] * A, dtype=np.int32)
    rl = np.ascontiguousarray(root_last if root_last is not None else [2] * A, dtype=np.int32)
    seq = np.zeros((mp.max_number_of_iterations, A - 1), dtype=np.uint8)
    counters = np.zeros(3, dtype=np.uint32)
    stride = hyp_dump_stride(cp, H)
    tree = np.zeros((max(tree_capacity, 1), stride))
    count = C.c_uint32(0)
    its = C.c_double(0)
    reference().ref_search_hypothesis(C.byref(mp), C.byref(hp), C.byref(cp), C.c_uint32(H), _ptr(lo), _ptr(hi), _ptr(rx),
                                      _ptr(rl), C.c_uint32(skip), _ptr(seq), _ptr(counters),
                                      _ptr(tree) if tree_capacity else None, C.c_uint32(tree_capacity), C.byref(count),
                                      C.byref(its))
    return dict(sequence=seq, num_nodes=int(counters[0]), iterations=int(counters[1]), best=int(counters[2]),
                tree=tree[:min(count.value, tree_capacity)], tree_count=count.value, iterations_per_s=its.value)


def ref_time_search_hypothesis(mp: Params, hp, cp: CrossingParams, gaps, threads: int):
    lo = np.ascontiguousarray([g[0] for g in gaps], dtype=np.int32)
    hi = np.ascontiguousarray([g[1] for g in gaps], dtype=np.int32)
    it, sec = C.c_double(0), C.c_double(0)
    reference().ref_time_search_hypothesis(C.byref(mp), C.byref(hp), C.byref(cp), C.c_uint32(len(gaps)), _ptr(lo), _ptr(hi),
                                           C.c_uint32(threads), C.byref(it), C.byref(sec))
    return it.value, sec.value


def ref_time_rollouts(mp: Params, cp: CrossingParams, threads: int, seconds: float):
    s, r = C.c_double(0), C.c_double(0)
    reference().ref_time_rollouts(C.byref(mp), C.byref(cp), C.c_uint32(threads),
                                  C.c_double(seconds), C.byref(s), C.byref(r))
    return s.value, r.value


def ref_time_search(mp: Params, cp: CrossingParams, threads: int, repeats: int):
    it, sec = C.c_double(0), C.c_double(0)
    reference().ref_time_search(C.byref(mp), C.byref(cp), C.c_uint32(threads), C.c_uint32(repeats),
                                C.byref(it), C.byref(sec))
    return it.value, sec.value


def ref_search_count_steps(mp: Params, cp: CrossingParams):
    """(execute() calls, iterations, nodes) of one deterministic reference search."""
    calls, it, nodes = C.c_uint64(0), C.c_uint32(0), C.c_uint32(0)
    reference().ref_search_count_steps(C.byref(mp), C.byref(cp), C.byref(calls), C.byref(it),
                                       C.byref(nodes))
    return calls.value, it.value, nodes.value


def ref_time_parallel_search(mp: Params, cp: CrossingParams, num_parallel: int, max_actions=7):
    it, sec = C.c_double(0), C.c_double(0)
    q = np.zeros(max_actions)
    n = np.zeros(max_actions, dtype=np.int64)
    reference().ref_time_parallel_search(C.byref(mp), C.byref(cp), C.c_uint32(num_parallel),
                                         C.byref(it), C.byref(sec), _ptr(q), _ptr(n),
                                         C.c_uint32(max_actions))
    return it.value, sec.value, q, n
