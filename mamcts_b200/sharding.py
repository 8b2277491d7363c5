"""Host-side logic of the root-parallel sharding over GPUs.

Trees are independent until the final merge (reference mcts/mcts.h:188-221), so the only things the
host decides are (1) which global tree ids a rank owns and (2) how the per-rank partial root
statistics combine. The partial vector is exactly what the device all-reduces
(`root_merge_kernel` -> `ncclAllReduce(sum, f64)` -> `root_finish_kernel`):

    acc = [sum_q[0..nA), sum_n[0..nA), occurrences[0..nA), sum_total_visits]        (3*nA + 1 doubles)

and the finish step is UctStatistic::merge_node_statistics + get_best_action
(reference mcts/statistics/uct_statistic.h:165-184, :78-89). Nothing here computes rollouts or
backups; this module has no device code and no CPU fallback for it - it is the glue that the
multi-process (torchrun) callers share, kept separate so it can be tested with gloo on CPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import numpy as np


@dataclass(frozen=True)
class Shard:
    first_tree_id: int   # global id of this rank's local tree 0 (the Philox stream_hi base)
    num_trees: int


def weak_shard(trees_per_gpu: int, world_size: int, rank: int) -> Shard:
    """Fixed trees per GPU (bench.py, `scaling: weak`): rank r owns [r*T, (r+1)*T)."""
    _check(world_size, rank)
    return Shard(rank * trees_per_gpu, trees_per_gpu)


def strong_shard(total_trees: int, world_size: int, rank: int) -> Shard:
    """Fixed total (C5: 65 536 trees over 8 GPUs): contiguous ranges, the first `total % world`
    ranks take one extra tree. The union over ranks is [0, total) for any world size."""
    _check(world_size, rank)
    base, extra = divmod(total_trees, world_size)
    first = rank * base + min(rank, extra)
    return Shard(first, base + (1 if rank < extra else 0))


def _check(world_size: int, rank: int):
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} of world size {world_size}")


def partial_size(num_actions: int) -> int:
    return 3 * num_actions + 1


def local_partial(q: np.ndarray, n: np.ndarray, expanded: np.ndarray, visits: Sequence[int]) -> np.ndarray:
    """The all-reduce payload of one rank from its trees' ego root statistics.
    q, n, expanded: [trees, nA]; visits: [trees]. Trees are added in index order (the device kernel's
    fixed order differs; sums of counts are exact and sums of Q agree to rounding)."""
    q = np.asarray(q, dtype=np.float64)
    n = np.asarray(n, dtype=np.float64)
    ex = np.asarray(expanded, dtype=bool)
    nA = q.shape[1]
    acc = np.zeros(partial_size(nA))
    for t in range(q.shape[0]):
        for a in range(nA):
            if ex[t, a]:
                acc[a] += q[t, a]
                acc[nA + a] += n[t, a]
                acc[2 * nA + a] += 1.0
        acc[3 * nA] += float(visits[t])
    return acc


def finish(acc: np.ndarray, num_actions: int) -> dict:
    """merge_node_statistics' division and mean (uct_statistic.h:173-183) + get_best_action (:78-89,
    first maximum in action order) on a fully reduced partial vector; mirrors root_finish_kernel."""
    nA = num_actions
    acc = np.asarray(acc, dtype=np.float64)
    if acc.shape != (partial_size(nA),):
        raise ValueError("partial vector has the wrong length")
    q = np.zeros(nA)
    value, present, best, best_q = 0.0, 0, 0, 0.0
    for a in range(nA):
        occ = acc[2 * nA + a]
        if occ > 0.0:
            q[a] = acc[a] / occ
            value += q[a]
            if present == 0 or q[a] > best_q:
                best, best_q = a, q[a]
            present += 1
    return dict(q=q, n=acc[nA:2 * nA].astype(np.uint64), occurrences=acc[2 * nA:3 * nA].astype(np.uint32),
                visits=int(acc[3 * nA]), value=(value / present) if present else 0.0, best=best)


# ---------------------------------------------------------------------------------------------
# C3: one decision, a large number of rollouts from the same leaf set
# ---------------------------------------------------------------------------------------------
# The rollout index range is partitioned like the trees; rollout r draws from Philox stream
# (stream_hi, stream_lo_base + r) with the GLOBAL index r, so the set of rollouts is the same for any
# number of GPUs. The exchange step is one all-reduce of {sum, sum of squares, count} per agent.

@dataclass(frozen=True)
class RolloutShard:
    first_rollout: int   # global index of this rank's rollout 0 (added to stream_lo_base)
    num_rollouts: int


def rollout_shard(total_rollouts: int, world_size: int, rank: int) -> RolloutShard:
    s = strong_shard(total_rollouts, world_size, rank)
    return RolloutShard(s.first_tree_id, s.num_trees)


def moments_partial(returns: np.ndarray) -> np.ndarray:
    """The all-reduce payload of one rank: per agent [sum, sum of squares], then the rollout count.
    returns: [agents, local rollouts] (ego first, then the other agents)."""
    r = np.atleast_2d(np.asarray(returns, dtype=np.float64))
    acc = np.zeros(2 * r.shape[0] + 1)
    acc[0:2 * r.shape[0]:2] = r.sum(axis=1)
    acc[1:2 * r.shape[0]:2] = (r * r).sum(axis=1)
    acc[-1] = r.shape[1]
    return acc


def moments_finish(acc: np.ndarray) -> dict:
    """Mean and (population) variance per agent from a fully reduced payload."""
    acc = np.asarray(acc, dtype=np.float64)
    count = acc[-1]
    if count <= 0:
        raise ValueError("no rollouts")
    s, ss = acc[0:-1:2], acc[1:-1:2]
    mean = s / count
    return dict(count=int(count), mean=mean, variance=np.maximum(ss / count - mean * mean, 0.0))
