"""CPU checks of the mctx restatement (oracle/mcts_oracle.c).  mctx itself is absent, so these pin the restatement
against mctx's *published* sequential-halving routine (restated verbatim in Python below) and against invariants of
the algorithm; MCTS parity versus a live mctx stays "unpinned" (DESIGN.md)."""
import math

import numpy as np
import pytest

import oracle as O


def mctx_sequence(m, S):
    """mctx/_src/seq_halving.py get_sequence_of_considered_visits, verbatim logic"""
    if m <= 1:
        return tuple(range(S))
    log2max = int(math.ceil(math.log2(m)))
    sequence, visits, k = [], [0] * m, m
    while len(sequence) < S:
        extra = max(1, int(S / (log2max * k)))
        for _ in range(extra):
            sequence.extend(visits[:k])
            for i in range(k):
                visits[i] += 1
        k = max(2, k // 2)
    return tuple(sequence[:S])


def test_considered_visit_closed_form_equals_mctx_table():
    for S in (1, 2, 5, 16, 50, 64, 75, 100, 200):
        for m in range(0, 17):
            s = mctx_sequence(m, S)
            assert [O.considered_visit(m, S, i) for i in range(S)] == list(s)


def _cfg(policy, qt, S, depth, A, Cn=0, E=4, **kw):
    d = dict(policy=policy, qtransform=qt, num_simulations=S, max_depth=depth, num_actions=A, num_chance=Cn, embed_dim=E,
             max_num_considered_actions=16, q_min=-1.0, q_max=1.0, value_scale=0.5, maxvisit_init=50.0, epsilon=1e-8,
             pb_c_init=1.25, pb_c_base=19652.0, dirichlet_fraction=0.25, temperature=1.0, gumbel_scale=1.0)
    d.update(kw)
    return O.MctsCfg(**d)


def _net(rng, A, E, Cn=0):
    W = rng.standard_normal((A + Cn, E, E)).astype(np.float32) * 0.7
    P = rng.standard_normal((E, A)).astype(np.float32)
    Pc = rng.standard_normal((E, max(Cn, 1))).astype(np.float32)

    def f(action, emb):
        nxt = np.tanh(np.einsum("nij,nj->ni", W[action], emb)).astype(np.float32)
        return dict(prior=(nxt @ P).astype(np.float32), value=np.tanh(nxt.sum(1)).astype(np.float32),
                    reward=(0.1 * nxt[:, 0]).astype(np.float32), discount=np.where(nxt[:, 1] > 0, 1.0, -1.0).astype(np.float32),
                    emb=nxt, chance=(nxt @ Pc).astype(np.float32))
    return f


def _search(cfg, n, seed, invalid_frac=0.3):
    rng = np.random.default_rng(seed)
    A, Cn, E = cfg.num_actions, cfg.num_chance, cfg.embed_dim
    tree = O.MctsTree(cfg, n)
    keys = rng.integers(0, 2**32, (n, 2), dtype=np.uint64).astype(np.uint32)
    invalid = (rng.random((n, A)) < invalid_frac).astype(np.uint8)
    invalid[np.arange(n), rng.integers(0, A, n)] = 0
    noise = rng.dirichlet(np.full(A, 0.3), n).astype(np.float32)
    O.mcts_init(tree, keys, rng.standard_normal((n, A)).astype(np.float32), rng.uniform(-1, 1, n).astype(np.float32),
                rng.standard_normal((n, E)).astype(np.float32), invalid, noise)
    net = _net(rng, A, E, Cn)
    for sim in range(cfg.num_simulations):
        parent, action, emb, isdec = O.mcts_select(tree, sim)
        o = net(np.clip(action, 0, A + Cn - 1), emb)
        if cfg.policy == 2:
            O.mcts_expand(tree, sim, parent, action, o["prior"], o["value"], o["reward"], o["discount"], o["emb"], o["chance"],
                          o["value"], o["emb"])
        else:
            O.mcts_expand(tree, sim, parent, action, o["prior"], o["value"], o["reward"], o["discount"], o["emb"])
    return tree, invalid, O.mcts_policy_output(tree)


@pytest.mark.parametrize("policy,qt,A,Cn", [(0, 0, 9, 0), (0, 1, 24, 0), (1, 2, 24, 0), (1, 2, 806, 0), (2, 1, 4, 6)])
def test_search_invariants(policy, qt, A, Cn):
    S = 40
    cfg = _cfg(policy, qt, S, 12, A, Cn)
    tree, invalid, (action, weights, value) = _search(cfg, 16, 5 + policy)
    root_visits = tree.children_visits[:, 0, :]
    assert (root_visits.sum(1) == S).all() and (tree.node_visits[:, 0] == S + 1).all()
    assert (root_visits[:, :A][invalid.astype(bool)] == 0).all()          # invalid root actions are never visited
    assert (root_visits[:, A:] == 0).all()                                 # decision root never picks a chance slot
    assert (invalid[np.arange(16), action] == 0).all()
    assert np.allclose(weights.sum(1), 1.0, atol=1e-5) and (np.abs(value) <= 1.5).all()
    # every expanded node hangs below its parent through the recorded action
    for g in range(16):
        for node in range(1, S + 1):
            p, a = tree.parents[g, node], tree.action_from_parent[g, node]
            if p >= 0:
                assert tree.children_index[g, p, a] == node
    if policy == 2:  # levels alternate
        for g in range(16):
            for node in range(1, S + 1):
                p = tree.parents[g, node]
                if p >= 0:
                    assert tree.is_decision[g, node] != tree.is_decision[g, p]


def test_gumbel_visits_follow_sequential_halving():
    """with m considered actions the root visit counts are exactly those of the halving schedule"""
    S, A = 32, 24
    cfg = _cfg(1, 2, S, 8, A)
    tree, invalid, _ = _search(cfg, 32, 11, invalid_frac=0.0)
    for g in range(32):
        counts = sorted(tree.children_visits[g, 0][tree.children_visits[g, 0] > 0].tolist(), reverse=True)
        # reconstruct the schedule: every phase visits each of the k considered actions `extra` more times
        k, log2max, pos = 16, 4, 0
        visits = [0] * 16
        while pos < S:
            extra = max(1, int(S / (log2max * k)))
            for _ in range(extra):
                for i in range(k):
                    if pos < S:
                        visits[i] += 1
                        pos += 1
            k = max(2, k // 2)
        assert counts == sorted([v for v in visits if v > 0], reverse=True)
