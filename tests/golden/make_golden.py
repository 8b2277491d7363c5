"""Generates tests/golden/*.json from the UNMODIFIED reference (oracle/_ref/libmamcts_ref.so, built
from /root/reference by oracle/Makefile). Run in the container that has the reference checkout:

    python tests/golden/make_golden.py

The vectors pin (a) CrossingState<int>::execute, (b) RandomHeuristic::rollout incl. the recorded
action / reward sequences used for REPLAY, (c) whole Mcts::search root statistics and a digest of
every node statistic of the tree, (d) UctStatistic::merge_node_statistics, (e) the libstdc++
widening orders. Floats are stored as hex so nothing is lost in JSON.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_api as O  # noqa: E402
from mamcts_b200 import default_crossing_params, default_params  # noqa: E402


def fhex(a):
    return [float(v).hex() for v in np.asarray(a, dtype=np.float64).ravel()]


def tree_digest(tree: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(tree, dtype=np.float64).tobytes()).hexdigest()


def dump(name, obj):
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(obj, f, separators=(",", ":"))
    print("wrote", name)


def gen_expand_order():
    table = {}
    for seed in (1000, 7, 42, 1, 2024):
        mp = default_params(random_seed=seed)
        table[str(seed)] = {str(n): O.ref_expand_order(mp, n) for n in (2, 3, 4, 5, 7, 8, 30)}
    dump("expand_order.json", table)


def gen_execute():
    rng = np.random.default_rng(12345)
    cases = []
    for num_others, chain, goal in ((1, 21, 12), (2, 21, 12), (3, 41, 26), (7, 21, 12)):
        for cost_only in (0, 1):
            cp = default_crossing_params(num_other_agents=num_others, chain_length=chain,
                                         ego_goal_pos=goal, cost_only_collision=cost_only)
            A = num_others + 1
            n = 64
            cpt = cp.crossing_point
            x = rng.integers(-1, cpt + 8, size=(A, n)).astype(np.int32)
            x[1:] = np.maximum(x[1:], 0)
            last = rng.integers(-3, 7, size=(A, n)).astype(np.int32)
            act = np.zeros((A, n), dtype=np.int32)
            act[0] = rng.integers(0, cp.num_ego_actions, size=n)
            act[1:] = rng.integers(-3, 7, size=(A - 1, n))
            nx, nl, fl, r, c0 = O.ref_crossing_execute(cp, x, last, act)
            cases.append(dict(num_others=num_others, chain_length=chain, ego_goal_pos=goal,
                              cost_only_collision=cost_only, x=x.tolist(), last=last.tolist(),
                              action=act.tolist(), next_x=nx.tolist(), next_last=nl.tolist(),
                              flags=fl.tolist(), reward=fhex(r), cost0=fhex(c0)))
    dump("crossing_execute.json", cases)


def gen_rollouts():
    rng = np.random.default_rng(777)
    cases = []
    for seed in (1000, 7, 42):
        for num_others in (1, 2, 7):
            for reward_step in (0.0, -1.5):
                mp = default_params(random_seed=seed, rh_max_number_of_iterations=50)
                cp = default_crossing_params(num_other_agents=num_others, reward_step=reward_step)
                A = num_others + 1
                n = 24
                x = rng.integers(0, 12, size=(A, n)).astype(np.int32)
                x[:, 0] = 5  # the reference's default start state
                depth = rng.integers(0, 4, size=n).astype(np.uint32)
                depth[1] = 999   # one step before MAX_SEARCH_DEPTH (1000)
                depth[2] = 1000  # at the cap: loop never runs
                flags = np.zeros(n, dtype=np.uint8)
                flags[3] = 1     # terminal start state
                r = O.ref_rollout_crossing(mp, cp, x, None, flags, depth, rec_stride=50)
                cases.append(dict(seed=seed, num_others=num_others, reward_step=reward_step,
                                  x=x.tolist(), depth=depth.tolist(), flags=flags.tolist(),
                                  ego_return=fhex(r["ego_return"]), other_return=fhex(r["other_return"]),
                                  ego_cost=fhex(r["ego_cost"]), step_length=fhex(r["step_length"]),
                                  steps=r["steps"].tolist(), final_x=r["final_x"].tolist(),
                                  final_last_action=r["final_last_action"].tolist(),
                                  final_flags=r["final_flags"].tolist(),
                                  rec_actions=r["rec_actions"].tolist(),
                                  rec_rewards=fhex(r["rec_rewards"])))
    dump("rollout_crossing.json", cases)
    # SimpleState: rollouts from every length, default cap 1000 would never end (F1): cap at 40.
    scases = []
    for seed in (1000, 7):
        mp = default_params(random_seed=seed, rh_max_number_of_iterations=40)
        length = np.arange(-1, 11, dtype=np.int32)
        r = O.ref_rollout_simple(mp, length, None, rec_stride=40)
        scases.append(dict(seed=seed, length=length.tolist(), ego_return=fhex(r["ego_return"]),
                           other_return=fhex(r["other_return"]), step_length=fhex(r["step_length"]),
                           steps=r["steps"].tolist(), final_len=r["final_len"].tolist(),
                           rec_actions=r["rec_actions"].tolist(), rec_rewards=fhex(r["rec_rewards"])))
    dump("rollout_simple.json", scases)


def gen_hypothesis():
    """C4 (BASELINE.json configs[3]): the rollout path of hypothesis-MCTS. (a) the gap-keeping policy
    AgentPolicyCrossingState<int>::calculate_action, (b) rollouts of
    RandomHeuristic::rollout<.., UctStatistic, HypothesisStatistic, ..> with the hypothesis sets of the
    reference's own tests (environments/tests/crossing_state_int_test.cc:105-106 and :190-192), recorded
    for REPLAY."""
    rng = np.random.default_rng(4242)
    pol = []
    for chain, vmin, vmax in ((21, -3, 3), (41, -2, 4)):
        cp = default_crossing_params(chain_length=chain, min_velocity_other=vmin, max_velocity_other=vmax)
        n = 600
        ax = rng.integers(-2, chain, n); al = rng.integers(vmin, vmax + 1, n)
        ex = rng.integers(-2, chain, n); el = rng.integers(-1, 3, n); gap = rng.integers(-6, 9, n)
        out = O.ref_policy_calculate_action(cp, ax, al, ex, el, gap)
        pol.append(dict(chain_length=chain, min_velocity_other=vmin, max_velocity_other=vmax,
                        agent_x=ax.tolist(), agent_last=al.tolist(), ego_x=ex.tolist(), ego_last=el.tolist(),
                        gap=gap.tolist(), action=out.tolist()))
    cases = []
    for seed, num_others, hyps in ((1000, 2, [(4, 5), (5, 6)]), (1000, 2, [(-2, 1), (2, 3), (4, 5)]),
                                   (7, 3, [(4, 5), (5, 6)]), (42, 1, [(1, 4)])):
        mp = default_params(random_seed=seed, rh_max_number_of_iterations=50)
        cp = default_crossing_params(num_other_agents=num_others)
        A = num_others + 1
        n = 48
        x = rng.integers(0, 12, size=(A, n)).astype(np.int32)
        x[:, 0] = 5
        last = rng.integers(-1, 3, size=(A, n)).astype(np.int32)
        last[:, 0] = 2
        cur = rng.integers(0, len(hyps), size=(num_others, n)).astype(np.uint32)
        lo = [h[0] for h in hyps]; hi = [h[1] for h in hyps]
        r = O.ref_rollout_crossing_hypothesis(mp, cp, x, last, lo, hi, cur, rec_stride=50)
        cases.append(dict(seed=seed, num_others=num_others, hyp_lo=lo, hyp_hi=hi, cur_hyp=cur.tolist(),
                          x=x.tolist(), last=last.tolist(), ego_return=fhex(r["ego_return"]),
                          other_return=fhex(r["other_return"]), ego_cost=fhex(r["ego_cost"]),
                          step_length=fhex(r["step_length"]), steps=r["steps"].tolist(),
                          final_x=r["final_x"].tolist(), final_last_action=r["final_last_action"].tolist(),
                          final_flags=r["final_flags"].tolist(), rec_actions=r["rec_actions"].tolist()))
    dump("hypothesis.json", dict(policy=pol, rollouts=cases))


def hypothesis_backup_case(rng, J, H, MA, C, n_trees, L, rounds, chance):
    """`rounds` batches of one path per tree. Every tree is a chain of up to L+1 nodes (node ids
    t*(L+1) + depth) that grows like a search does: a batch either expands the chain by one NEW leaf,
    which receives the heuristic update (the reference asserts it is the node's first visit,
    hypothesis_statistic.h:105), or re-visits the current deepest node without one (a terminal node,
    stage_node.h:183-187); all ancestors are updated in both cases, so running means see many visits."""
    batches = []
    depth = np.zeros(n_trees, dtype=np.int64)       # current deepest node of every chain
    for _ in range(rounds):
        grow = (rng.random(n_trees) < 0.75) & (depth < L)
        depth = depth + grow
        plen = depth.copy()                          # edges between the leaf and the root
        leaf = np.arange(n_trees) * (L + 1) + depth
        has = grow.astype(np.uint8)
        lc = np.round(rng.normal(size=(n_trees, C)) * 10, 3)
        po = np.concatenate([[0], np.cumsum(plen)[:-1]])
        ne = int(plen.sum())
        en = np.zeros(ne, dtype=np.uint32)
        for t in range(n_trees):
            for e in range(plen[t]):
                en[po[t] + e] = t * (L + 1) + (plen[t] - 1 - e)
        eh = rng.integers(0, H, (ne, J))
        ea = rng.integers(0, MA, (ne, J))
        ec = rng.choice([0.0, 1000.0, -0.5, 2.25], (ne, C))
        batches.append(dict(leaf_node=leaf.tolist(), leaf_has=has.tolist(), leaf_cost=lc.tolist(),
                            path_offset=po.tolist(), path_len=plen.tolist(), edge_node=en.tolist(),
                            edge_hyp=eh.tolist(), edge_action=ea.tolist(), edge_cost=ec.tolist()))
    return dict(J=J, H=H, MA=MA, C=C, n_nodes=n_trees * (L + 1), chance=chance, batches=batches)


def concat_batches(batches):
    """All batches as one sequential path list (what the reference driver executes)."""
    out = {k: [] for k in ("leaf_node", "leaf_has", "leaf_cost", "path_offset", "path_len", "edge_node",
                           "edge_hyp", "edge_action", "edge_cost")}
    base = 0
    for b in batches:
        for k in ("leaf_node", "leaf_has", "leaf_cost", "path_len", "edge_node", "edge_hyp", "edge_action", "edge_cost"):
            out[k] += b[k]
        out["path_offset"] += [o + base for o in b["path_offset"]]
        base += len(b["edge_node"])
    return out


def gen_hypothesis_backup():
    """HypothesisStatistic back-propagation (hypothesis_statistic.h:100-134) executed by real
    HypothesisStatistic objects of the reference on random path batches."""
    from mamcts_b200.engine import Engine
    rng = np.random.default_rng(2718)
    cases = []
    for J, H, MA, C, chance in ((1, 2, 6, 2, False), (2, 3, 7, 2, False), (3, 2, 7, 1, True), (2, 3, 7, 2, True)):
        case = hypothesis_backup_case(rng, J, H, MA, C, n_trees=12, L=5, rounds=9, chance=chance)
        mp = default_params()
        st = Engine.new_hyp_stats(case["n_nodes"] * J, H, MA)
        cat = concat_batches(case["batches"])
        if not cat["edge_node"]:
            continue
        O.ref_backup_hypothesis(mp, J, case["n_nodes"], st, cat["leaf_node"], cat["leaf_has"], cat["leaf_cost"],
                                cat["path_offset"], cat["path_len"], cat["edge_node"], cat["edge_hyp"],
                                cat["edge_action"], cat["edge_cost"], chance)
        case["expected"] = dict(ego_cost_value=fhex(st["ego_cost_value"]), latest_ego_cost=fhex(st["latest_ego_cost"]),
                                total_visits=st["total_visits"].tolist(), num_expanded=st["num_expanded"].tolist(),
                                visits_hypothesis=st["visits_hypothesis"].tolist(),
                                action_cost=fhex(st["action_cost"]), action_count=st["action_count"].tolist())
        cases.append(case)
    dump("hypothesis_backup.json", cases)


def f32hex(a):
    """float32 arrays as their uint32 bit patterns (exact, incl. denormals and -0.0)."""
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).ravel().tolist()


def gen_float_domain():
    """CrossingState<float>: execute on random float states / velocities and whole
    RandomHeuristic rollouts with UCT statistics (index-valued actions: an other agent's index i acts as
    the float with bit pattern i, F5), recorded for REPLAY."""
    from mamcts_b200 import default_crossing_params_f32
    rng = np.random.default_rng(31415)
    ex = []
    for num_others, chain, goal, cost_only in ((1, 21.0, 12.0, 0), (2, 20.0, 11.5, 1), (3, 41.0, 26.0, 0)):
        cp = default_crossing_params_f32(num_other_agents=num_others, chain_length=chain, ego_goal_pos=goal,
                                         cost_only_collision=cost_only, reward_step=-0.25)
        A, n = num_others + 1, 96
        x = (rng.random((A, n)) * (chain / 2 + 6) - 1).astype(np.float32)
        x[1:] = np.maximum(x[1:], 0)
        last = (rng.random((A, n)) * 6 - 3).astype(np.float32)
        act = np.zeros((A, n), dtype=np.float32)
        act[0] = rng.integers(0, 4, n)
        act[1:] = (rng.random((A - 1, n)) * 8 - 4).astype(np.float32)
        # a third of the cases sit just before the crossing point and move forward: collisions
        cpt = (chain - 1) / 2 + 1
        k = n // 3
        x[:, :k] = (cpt - rng.random((A, k)) * 2.5).astype(np.float32)
        act[0, :k] = rng.integers(2, 4, k)
        act[1:, :k] = (rng.random((A - 1, k)) * 4).astype(np.float32)
        nx, nl, fl, r, c0 = O.ref_crossing_execute_f32(cp, x, last, act)
        ex.append(dict(num_others=num_others, chain_length=chain, ego_goal_pos=goal, cost_only_collision=cost_only,
                       reward_step=-0.25, x=f32hex(x), last=f32hex(last), action=f32hex(act), next_x=f32hex(nx),
                       next_last=f32hex(nl), flags=fl.tolist(), reward=fhex(r), cost0=fhex(c0)))
    ro = []
    for seed, num_others in ((1000, 2), (7, 1), (42, 3)):
        mp = default_params(random_seed=seed, rh_max_number_of_iterations=50)
        cp = default_crossing_params_f32(num_other_agents=num_others)
        A, n = num_others + 1, 40
        x = (rng.random((A, n)) * 12).astype(np.float32)
        x[:, 0] = 5.0
        depth = rng.integers(0, 4, size=n).astype(np.uint32)
        r = O.ref_rollout_crossing_f32(mp, cp, x, None, depth, rec_stride=50)
        ro.append(dict(seed=seed, num_others=num_others, x=f32hex(x), depth=depth.tolist(),
                       const_actions=[O.ref_expand_order(mp, 4)[0]] + [O.ref_expand_order(mp, 30)[0]] * num_others,
                       ego_return=fhex(r["ego_return"]), ego_cost=fhex(r["ego_cost"]),
                       step_length=fhex(r["step_length"]), steps=r["steps"].tolist(), final_x=f32hex(r["final_x"]),
                       final_last_action=f32hex(r["final_last_action"]), final_flags=r["final_flags"].tolist(),
                       rec_actions=f32hex(r["rec_actions"])))
    dump("float_domain.json", dict(execute=ex, rollouts=ro))


def gen_float_hypothesis():
    """CrossingState<float> with behaviour hypotheses (AgentPolicyCrossingState<float>, crossing_state_agent_policy.h:
    35-55,78-84): calculate_action on random inputs, and whole hypothesis rollouts of the reference (its own
    uniform_real_distribution gap draws) recorded for REPLAY."""
    from mamcts_b200 import default_crossing_params_f32
    rng = np.random.default_rng(27182)
    cp = default_crossing_params_f32(num_other_agents=2)
    n = 400
    ax = (rng.random(n) * 16 - 1).astype(np.float32)
    al = (rng.random(n) * 6 - 3).astype(np.float32)
    ex = (rng.random(n) * 16 - 1).astype(np.float32)
    el = (rng.random(n) * 3 - 1).astype(np.float32)
    gap = (rng.random(n) * 10 - 4).astype(np.float32)
    gap[:40] = 0.0
    al[40:80] = np.float32(-0.0)
    pol = dict(agent_x=f32hex(ax), agent_last=f32hex(al), ego_x=f32hex(ex), ego_last=f32hex(el), gap=f32hex(gap),
               action=f32hex(O.ref_policy_calculate_action_f32(cp, ax, al, ex, el, gap)))
    ro = []
    for seed, num_others, hyps in ((1000, 2, [(4.0, 5.0), (5.0, 6.0)]), (7, 1, [(-2.0, 1.0), (2.0, 3.0), (4.0, 5.5)]),
                                   (42, 3, [(0.5, 2.5)])):
        mp = default_params(random_seed=seed, rh_max_number_of_iterations=50)
        cpx = default_crossing_params_f32(num_other_agents=num_others)
        A, n = num_others + 1, 40
        x = (rng.random((A, n)) * 10).astype(np.float32)
        last = (rng.random((A, n)) * 3 - 1).astype(np.float32)
        depth = rng.integers(0, 4, size=n).astype(np.uint32)
        cur = rng.integers(0, len(hyps), size=(num_others, n)).astype(np.uint32)
        r = O.ref_rollout_crossing_f32_hypothesis(mp, cpx, x, last, [h[0] for h in hyps], [h[1] for h in hyps], cur, depth,
                                                  rec_stride=50)
        ro.append(dict(seed=seed, num_others=num_others, hyps=hyps, x=f32hex(x), last=f32hex(last), depth=depth.tolist(),
                       cur_hyp=cur.tolist(), ego_return=fhex(r["ego_return"]), ego_cost=fhex(r["ego_cost"]),
                       steps=r["steps"].tolist(), final_x=f32hex(r["final_x"]), final_last_action=f32hex(r["final_last_action"]),
                       final_flags=r["final_flags"].tolist(), rec_actions=f32hex(r["rec_actions"])))
    dump("float_hypothesis.json", dict(policy=pol, rollouts=ro))


def gen_search():
    cases = []
    # C1: SimpleState(4), 20 iterations (reference test/uct/uct_test.cc:22-40 parameters)
    for iters, nodes in ((20, 100), (200, 100), (2000, 100000)):
        mp = default_params(max_number_of_iterations=iters, max_number_of_nodes=nodes)
        s = O.ref_search_simple(mp, 4, max_actions=2, tree_capacity=iters + 1)
        cases.append(dict(env="simple", iterations=iters, max_nodes=nodes, seed=1000, state_length=4,
                          value=fhex(s["value"]), visits=s["visits"].tolist(), q=fhex(s["q"]),
                          n=s["n"].tolist(), num_nodes=s["num_nodes"], best=s["best_action"],
                          tree_count=s["tree_count"], tree_sha256=tree_digest(s["tree"])))
    # C2: CrossingState, 1 other, 10 000 iterations; plus A = 3, 4, 8 and other seeds
    for seed, num_others, iters, chain, goal in ((1000, 1, 10000, 21, 12), (1000, 2, 3000, 21, 12),
                                                 (7, 1, 3000, 21, 12), (42, 2, 2000, 21, 12),
                                                 (1000, 3, 2000, 41, 26), (1000, 7, 1500, 21, 12)):
        mp = default_params(random_seed=seed, max_number_of_iterations=iters,
                            max_number_of_nodes=100000000, rh_max_number_of_iterations=50)
        cp = default_crossing_params(num_other_agents=num_others, chain_length=chain,
                                     ego_goal_pos=goal)
        s = O.ref_search_crossing(mp, cp, max_actions=7, tree_capacity=iters + 1)
        cases.append(dict(env="crossing", iterations=iters, seed=seed, num_others=num_others,
                          chain_length=chain, ego_goal_pos=goal, value=fhex(s["value"]),
                          visits=s["visits"].tolist(), q=fhex(s["q"]), n=s["n"].tolist(),
                          num_nodes=s["num_nodes"], best=s["best_action"],
                          tree_count=s["tree_count"], tree_sha256=tree_digest(s["tree"])))
    dump("search.json", cases)


def gen_merge():
    rng = np.random.default_rng(99)
    cases = []
    mp = default_params()
    for n_trees, nA in ((2, 4), (8, 4), (5, 7), (64, 4)):
        q = rng.normal(size=(n_trees, nA)) * 10
        cnt = rng.integers(1, 500, size=(n_trees, nA)).astype(np.int64)
        absent = rng.random(size=(n_trees, nA)) < 0.25
        absent[0, 0] = False
        cnt[absent] = -1
        q[absent] = 0.0
        visits = np.array([max(1, int(c[c > 0].sum())) for c in cnt], dtype=np.uint32)
        m = O.ref_merge_uct(mp, q, cnt, visits, nA)
        cases.append(dict(num_actions=nA, q=fhex(q), cnt=cnt.tolist(), visits=visits.tolist(),
                          merged_q=fhex(m["q"]), merged_n=m["n"].tolist(), merged_visits=m["visits"],
                          merged_value=float(m["value"]).hex(), best=m["best"]))
    dump("merge.json", cases)


def gen_action_values():
    """ActionValueRandomHeuristic::calculate_heuristic_values (action_value_random_heuristic.h:27-112) run by the
    reference's own loop (oracle/ref_driver.cc: ref_action_values_crossing) on random CrossingState<int> leaves."""
    rng = np.random.default_rng(2024)
    cases = []
    for others, seed, kw in ((1, 1000, {}), (2, 1000, {}), (3, 7, dict(chain_length=41, ego_goal_pos=26)),
                             (7, 42, {}), (1, 5, dict(reward_step=-1.0, cost_only_collision=1))):
        cp = default_crossing_params(num_other_agents=others, **kw)
        mp = default_params(random_seed=seed, rh_max_number_of_iterations=50, max_search_depth=60)
        A, n = others + 1, 48
        x = rng.integers(-1, cp.chain_length + 3, size=(A, n)).astype(np.int32)
        last = rng.integers(-2, 4, size=(A, n)).astype(np.int32)
        flags = (rng.random(n) < 0.1).astype(np.uint8)
        depth = rng.choice([0, 3, 20, 58, 59, 60], size=n).astype(np.uint32)
        r = O.ref_action_values_crossing(mp, cp, x, last, flags, depth)
        cases.append(dict(num_others=others, seed=seed, params=kw, x=x.tolist(), last=last.tolist(), flags=flags.tolist(),
                          depth=depth.tolist(), const_actions=[O.ref_expand_order(mp, 4 if a == 0 else 7)[0] for a in range(A)],
                          action_return=fhex(r["action_return"]), action_cost=fhex(r["action_cost"]),
                          other_return=fhex(r["other_return"]), mean_cost=fhex(r["mean_cost"])))
    dump("action_values.json", cases)


def gen_hypothesis_search():
    """Whole hypothesis-MCTS searches of the reference (Mcts<CrossingState<int>, UctStatistic, HypothesisStatistic,
    RandomHeuristic>::search(state, belief_tracker)): the hypothesis sequence its belief tracker handed out, the
    number of nodes, the best action and a digest of the canonical dump of every node statistic."""
    from mamcts_b200 import default_hypothesis_params
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_hypothesis_search import CASES
    out = []
    for case in CASES:
        mp = default_params(max_number_of_iterations=case["iters"], max_number_of_nodes=case.get("max_nodes", 100000000),
                            rh_max_number_of_iterations=50)
        hp = default_hypothesis_params(**case["hkw"])
        cp = default_crossing_params(num_other_agents=case["others"], **case["kw"])
        r = O.ref_search_hypothesis(mp, hp, cp, case["gaps"], skip=case["skip"], tree_capacity=case["iters"] + 1)
        assert r["tree_count"] == r["num_nodes"]
        out.append(dict(name=case["name"], others=case["others"], gaps=case["gaps"], iters=case["iters"], skip=case["skip"],
                        sequence=r["sequence"].ravel().tolist(), num_nodes=r["num_nodes"], best=r["best"],
                        tree_sha256=tree_digest(r["tree"]), root_record=fhex(r["tree"][0])))
    dump("hypothesis_search.json", out)


if __name__ == "__main__":
    assert O.reference_available(), "needs /root/reference (or a prebuilt oracle/_ref)"
    gen_expand_order()
    gen_execute()
    gen_rollouts()
    gen_hypothesis()
    gen_hypothesis_backup()
    gen_float_domain()
    gen_float_hypothesis()
    gen_search()
    gen_merge()
    gen_action_values()
    gen_hypothesis_search()
