"""MCTS parity proper: the CUDA tree kernels (through the C-ABI) versus the CPU oracle, simulation by simulation, fed
IDENTICAL network outputs (computed once on the GPU by a small torch stand-in for the Flax networks).  Visit counts and
every tree array must match exactly (floats bit-for-bit, which is tighter than the 1e-5 relative bar of BASELINE)."""
import numpy as np
import pytest
import torch

import oracle as O

pytestmark = pytest.mark.gpu

TREE_FIELDS = ("node_visits", "raw_values", "node_values", "parents", "action_from_parent", "children_index",
               "children_prior_logits", "children_visits", "children_rewards", "children_discounts", "children_values",
               "embeddings", "root_invalid_actions", "search_key", "policy_key")


def _mk_cfgs(policy, qt, S, depth, A, Cn, E, **kw):
    from exploring_muzero_on_dog_b200 import _lib
    d = dict(policy=policy, qtransform=qt, num_simulations=S, max_depth=depth, num_actions=A, num_chance=Cn, embed_dim=E,
             max_num_considered_actions=16, q_min=-1.0, q_max=1.0, value_scale=0.5, maxvisit_init=50.0, epsilon=1e-8,
             pb_c_init=1.25, pb_c_base=19652.0, dirichlet_fraction=0.25, temperature=1.0, gumbel_scale=1.0)
    d.update(kw)
    return _lib.MctsCfg(**d), O.MctsCfg(**d)


class Net:
    """deterministic stand-in for dynamics+prediction: emb' = tanh(W[a] emb), heads are linear maps of emb'"""

    def __init__(self, A, Cn, E, seed, quantize=False):
        self.quantize = quantize  # priors / values on a coarse grid: many exactly equal scores, the first index must win
        g = torch.Generator(device="cuda").manual_seed(seed)
        self.W = torch.randn(A + Cn, E, E, device="cuda", generator=g) * 0.7
        self.P = torch.randn(E, A, device="cuda", generator=g)
        self.Pc = torch.randn(E, max(Cn, 1), device="cuda", generator=g)

    def __call__(self, action, emb):
        nxt = torch.tanh(torch.einsum("nij,nj->ni", self.W[action], emb))
        if self.quantize:
            return dict(prior=torch.round(nxt @ self.P), value=torch.round(torch.tanh(nxt.sum(1)) * 2) / 2, reward=torch.zeros_like(nxt[:, 0]),
                        discount=torch.where(nxt[:, 1] > 0, 1.0, -1.0), emb=nxt, chance=torch.round(nxt @ self.Pc))
        return dict(prior=nxt @ self.P, value=torch.tanh(nxt.sum(1)), reward=0.1 * nxt[:, 0],
                    discount=torch.where(nxt[:, 1] > 0, 1.0, -1.0), emb=nxt, chance=nxt @ self.Pc)


def _compare_trees(search, otree, what):
    t = search.tree
    for k in TREE_FIELDS + (("is_decision",) if t.is_decision is not None else ()) + (("root_gumbel",) if t.root_gumbel is not None else ()):
        got = getattr(t, k).cpu().numpy()
        exp = getattr(otree, k)
        if got.dtype == np.float32:
            same = (got.view(np.int32) == exp.view(np.int32)) | (np.isnan(got) & np.isnan(exp))
        else:
            same = got == exp
        if not same.all():
            idx = np.argwhere(~same)[0]
            raise AssertionError(f"{what}: tree.{k}{tuple(idx)}: cuda {got[tuple(idx)]!r} oracle {exp[tuple(idx)]!r}")


CASES = [
    # (policy, qtransform, A, C, sims, depth)      the reference's four search configurations + DOG-sized trees
    (0, 0, 9, 0, 50, 9),       # TicTacToe run_mcts: muzero_policy, by_min_max(-1,1), max_depth 9       (TicTacToe/mcts.py:13-22)
    (1, 0, 9, 0, 50, 9),       # TicTacToe run_gumbel                                                   (TicTacToe/mcts.py:29-37)
    (1, 2, 24, 0, 100, 50),    # det MADN run_muzero_mcts: gumbel, completed_by_mix_value(0.5)          (muzero_deterministic_madn.py:673-684)
    (2, 1, 4, 6, 64, 50),      # classic MADN run_stochastic_muzero_mcts: by_parent_and_siblings        (muzero_classic_madn.py:488-501)
    (0, 1, 24, 0, 40, 5),      # muzero_policy, depth cut revisits
    (1, 2, 806, 0, 100, 50),   # DOG-wide action space (config 5)
    (0, 2, 806, 0, 30, 8),
    (1, 2, 806, 0, 100, 50, "ties"),   # quantised priors and values: exact ties everywhere (argmax keeps the FIRST maximum)
    (1, 2, 100, 0, 60, 50, "ties"),
    (1, 2, 806, 0, 100, 50, "peaked"),  # network-like rows: a few actions hold nearly all of the prior mass (the visited children
                                        # then carry > 90 % of it: the cancellation guard of the row-free level)
    (1, 2, 806, 0, 100, 50, "large"),   # logits beyond +-100: the magnitude guards send every level to the exact evaluation
    (1, 2, 806, 0, 100, 50, "flat"),    # nearly equal logits (1e-4 apart): two near-maximal children at every level
    (1, 2, 33, 0, 40, 50),             # the narrowest wide tree
    (1, 2, 832, 0, 20, 50),            # the widest register-path tree
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(str(x) for x in c))
def test_cuda_search_equals_oracle(case):
    from exploring_muzero_on_dog_b200 import mcts
    policy, qt, A, Cn, S, depth = case[:6]
    ties = len(case) > 6 and case[6] == "ties"
    shape = case[6] if len(case) > 6 else ""
    n, E = 96, 16
    ccfg, ocfg = _mk_cfgs(policy, qt, S, depth, A, Cn, E)
    rng = np.random.default_rng(1000 * policy + A)
    keys = rng.integers(0, 2**32, (n, 2), dtype=np.uint64).astype(np.uint32)
    prior = rng.standard_normal((n, A)).astype(np.float32) * 2
    if ties:
        prior = np.round(prior)
    value = rng.uniform(-1, 1, n).astype(np.float32)
    emb = rng.standard_normal((n, E)).astype(np.float32)
    invalid = (rng.random((n, A)) < 0.4).astype(np.uint8)
    invalid[np.arange(n), rng.integers(0, A, n)] = 0
    invalid[:4] = 0
    noise = rng.dirichlet(np.full(A, 0.3), n).astype(np.float32)
    net = Net(A, Cn, E, 7, quantize=ties)
    if shape in ("peaked", "large", "flat"):  # reshape every prior row the search sees, the root's and the network's
        def reshape_np(p, r):
            if shape == "peaked":
                p = p * 3.0
                p[np.arange(p.shape[0])[:, None], r.integers(0, A, (p.shape[0], 3))] += 22.0
                return p.astype(np.float32)
            if shape == "large":
                return (p * 70.0).astype(np.float32)
            return (np.round(p * 4) * 1e-4 + 3.0).astype(np.float32)
        prior = reshape_np(prior, rng)
        base_net = net

        def net(action, pemb, _r=np.random.default_rng(77)):
            o = base_net(action, pemb)
            o["prior"] = torch.as_tensor(reshape_np(o["prior"].float().cpu().numpy(), _r), device="cuda")
            return o

    s = mcts.Search(ccfg, n)
    dev = lambda x: torch.as_tensor(x, device="cuda")
    keys_t = torch.from_numpy(keys).cuda()
    s.init(keys_t, mcts.RootFnOutput(dev(prior), dev(value), dev(emb)), dev(invalid), dev(noise))
    otree = O.MctsTree(ocfg, n)
    O.mcts_init(otree, keys, prior, value, emb, invalid, noise)
    _compare_trees(s, otree, "init")
    for sim in range(S):
        parent, action, pemb, isdec = s.select(sim)
        op, oa, oemb, oisdec = O.mcts_select(otree, sim)
        assert np.array_equal(parent.cpu().numpy(), op), f"parent differs at sim {sim}"
        assert np.array_equal(action.cpu().numpy(), oa), f"action differs at sim {sim}: {np.argwhere(action.cpu().numpy() != oa)[:3].tolist()}"
        assert np.array_equal(pemb.cpu().numpy(), oemb)
        o = net(action.long().clamp(0, A + Cn - 1), pemb)
        h = {k: v.float().contiguous().cpu().numpy() for k, v in o.items()}
        if policy == 2:
            s.expand(sim, o["prior"], o["value"], o["reward"], o["discount"], o["emb"], o["chance"], o["value"], o["emb"])
            O.mcts_expand(otree, sim, op, oa, h["prior"], h["value"], h["reward"], h["discount"], h["emb"], h["chance"], h["value"], h["emb"])
        else:
            s.expand(sim, o["prior"], o["value"], o["reward"], o["discount"], o["emb"])
            O.mcts_expand(otree, sim, op, oa, h["prior"], h["value"], h["reward"], h["discount"], h["emb"])
    _compare_trees(s, otree, "after search")
    out, root_value = s.policy_output()
    oact, ow, ov = O.mcts_policy_output(otree)
    assert np.array_equal(out.action.cpu().numpy(), oact)
    assert np.array_equal(out.action_weights.cpu().numpy().view(np.int32), ow.view(np.int32))
    assert np.array_equal(root_value.cpu().numpy().view(np.int32), ov.view(np.int32))
    vc = s.tree.children_visits[:, 0].cpu().numpy()
    assert (vc.sum(1) == S).all() and (vc[:, :A][invalid.astype(bool)] == 0).all()


class WideNet:
    """stand-in network for the real-shape cases (no per-action [E, E] matrices: A x E x E would not fit at E = 258)"""

    def __init__(self, A, Cn, E, seed):
        g = torch.Generator(device="cuda").manual_seed(seed)
        self.Wd = torch.randn(E, E, device="cuda", generator=g) * (1.6 / E ** 0.5)
        self.Wa = torch.randn(A + Cn, E, device="cuda", generator=g)
        self.P = torch.randn(E, A, device="cuda", generator=g) * (2.0 / E ** 0.5)
        self.Pc = torch.randn(E, max(Cn, 1), device="cuda", generator=g) * (2.0 / E ** 0.5)

    def __call__(self, action, emb):
        nxt = torch.tanh(emb @ self.Wd + self.Wa[action])
        return dict(prior=nxt @ self.P, value=torch.tanh(nxt[:, :8].sum(1)), reward=0.1 * nxt[:, 0],
                    discount=torch.where(nxt[:, 1] > 0, 1.0, -1.0), emb=nxt, chance=nxt @ self.Pc)


REAL_SHAPES = [
    # (policy, qtransform, A, C, E, sims, depth, games)
    (2, 1, 4, 6, 258, 64, 50, 4096),    # BASELINE config 3: stochastic MuZero, 4,096 games x 64 sims, embedding 258 (two warp_copy passes)
    (1, 2, 806, 0, 256, 100, 50, 1024), # BASELINE config 5: Gumbel over DOG's 806 actions, 100 sims, latent 256 (select cache path)
    (1, 2, 24, 0, 256, 100, 50, 2048),  # det MADN run_muzero_mcts at training shape
]


@pytest.mark.parametrize("case", REAL_SHAPES, ids=lambda c: "-".join(str(x) for x in c))
def test_cuda_search_equals_oracle_at_real_shapes(case):
    """the BASELINE shapes, through the PRODUCTION launch sequence (select(0), then expand_select fused launches, a final expand):
    parent / action of every simulation, every tree array and the policy output bit-equal to the oracle"""
    from exploring_muzero_on_dog_b200 import mcts
    policy, qt, A, Cn, E, S, depth, n = case
    ccfg, ocfg = _mk_cfgs(policy, qt, S, depth, A, Cn, E, dirichlet_fraction=0.25)
    rng = np.random.default_rng(77 + A)
    keys = rng.integers(0, 2**32, (n, 2), dtype=np.uint64).astype(np.uint32)
    prior = rng.standard_normal((n, A)).astype(np.float32) * 2
    value = rng.uniform(-1, 1, n).astype(np.float32)
    emb = rng.standard_normal((n, E)).astype(np.float32)
    invalid = (rng.random((n, A)) < 0.5).astype(np.uint8)
    invalid[np.arange(n), rng.integers(0, A, n)] = 0
    noise = rng.dirichlet(np.full(A, 0.3), n).astype(np.float32)
    net = WideNet(A, Cn, E, 11)
    s = mcts.Search(ccfg, n)
    dev = lambda x: torch.as_tensor(x, device="cuda")
    s.init(torch.from_numpy(keys).cuda(), mcts.RootFnOutput(dev(prior), dev(value), dev(emb)), dev(invalid), dev(noise))
    otree = O.MctsTree(ocfg, n)
    O.mcts_init(otree, keys, prior, value, emb, invalid, noise)
    parent, action, pemb, _ = s.select(0)
    for sim in range(S):
        op, oa, oemb, _ = O.mcts_select(otree, sim)
        assert np.array_equal(parent.cpu().numpy(), op), f"parent differs at sim {sim}"
        assert np.array_equal(action.cpu().numpy(), oa), f"action differs at sim {sim}: games {np.flatnonzero(action.cpu().numpy() != oa)[:5].tolist()}"
        assert np.array_equal(pemb.cpu().numpy().view(np.int32), oemb.view(np.int32)), f"embedding differs at sim {sim}"
        o = net(action.long().clamp(0, A + Cn - 1), pemb)
        h = {k: v.float().contiguous().cpu().numpy() for k, v in o.items()}
        step = s.expand_select if sim + 1 < S else s.expand
        if policy == 2:
            step(sim, o["prior"], o["value"], o["reward"], o["discount"], o["emb"], o["chance"], o["value"], o["emb"])
            O.mcts_expand(otree, sim, op, oa, h["prior"], h["value"], h["reward"], h["discount"], h["emb"], h["chance"], h["value"], h["emb"])
        else:
            step(sim, o["prior"], o["value"], o["reward"], o["discount"], o["emb"])
            O.mcts_expand(otree, sim, op, oa, h["prior"], h["value"], h["reward"], h["discount"], h["emb"])
    _compare_trees(s, otree, "after search")
    out, root_value = s.policy_output()
    oact, ow, ov = O.mcts_policy_output(otree)
    assert np.array_equal(out.action.cpu().numpy(), oact)
    assert np.array_equal(out.action_weights.cpu().numpy().view(np.int32), ow.view(np.int32))
    assert np.array_equal(root_value.cpu().numpy().view(np.int32), ov.view(np.int32))
    assert (s.tree.children_visits[:, 0].sum(1) == S).all()


def test_mctx_shaped_policies_run():
    """the drop-in wrappers (same names / kwargs as mctx) on a torch recurrent_fn"""
    import functools
    from exploring_muzero_on_dog_b200 import mcts
    n, A, E = 64, 24, 32
    net = Net(A, 6, E, 3)
    g = torch.Generator(device="cuda").manual_seed(0)
    root = mcts.RootFnOutput(torch.randn(n, A, device="cuda", generator=g), torch.rand(n, device="cuda", generator=g) * 2 - 1,
                             torch.randn(n, E, device="cuda", generator=g))
    keys = torch.randint(0, 2**31, (n, 2), device="cuda", generator=g).to(torch.uint32)
    invalid = torch.rand(n, A, device="cuda", generator=g) < 0.5
    invalid[:, 0] = False

    def recurrent_fn(params, rng_key, action, embedding):
        o = net(action, embedding)
        return mcts.RecurrentFnOutput(o["reward"], o["discount"], o["prior"], o["value"]), o["emb"]

    out = mcts.gumbel_muzero_policy(None, keys, root, recurrent_fn, 50, invalid_actions=invalid, max_depth=25,
                                    qtransform=functools.partial(mcts.qtransform_completed_by_mix_value, value_scale=0.5), gumbel_scale=1.0)
    assert out.action.shape == (n,) and not invalid[torch.arange(n), out.action.long()].any()
    assert torch.allclose(out.action_weights.sum(1), torch.ones(n, device="cuda"), atol=1e-5)
    assert (out.search_tree.summary().visit_counts.sum(1) == 50).all()
    out = mcts.muzero_policy(None, keys, root, recurrent_fn, 30, invalid_actions=invalid, max_depth=9,
                             qtransform=functools.partial(mcts.qtransform_by_min_max, min_value=-1, max_value=1), dirichlet_fraction=0.0)
    assert (out.search_tree.summary().visit_counts.sum(1) == 30).all()

    root4 = mcts.RootFnOutput(root.prior_logits[:, :4].contiguous(), root.value, root.embedding)
    net4 = Net(4, 6, E, 5)

    def dec(params, rng_key, action, embedding):
        o = net4(action, embedding)
        return mcts.DecisionRecurrentFnOutput(o["chance"], o["value"]), torch.cat([o["emb"], o["reward"][:, None], o["discount"][:, None]], 1)

    def ch(params, rng_key, outcome, afterstate):
        o = net4(outcome + 4, afterstate[:, :-2])
        return mcts.ChanceRecurrentFnOutput(o["prior"], o["value"], afterstate[:, -2], afterstate[:, -1]), o["emb"]

    out = mcts.stochastic_muzero_policy(None, keys, root4, dec, ch, 64, invalid_actions=invalid[:, :4].contiguous(), max_depth=50,
                                        qtransform=mcts.qtransform_by_parent_and_siblings, temperature=1.0)
    assert (out.search_tree.summary().visit_counts.sum(1) == 64).all()
    assert out.search_tree.children_visits.shape[-1] == 10


def test_stochastic_policy_with_unpadded_embeddings_equals_the_padded_search():
    """stochastic_muzero_policy hands the state / afterstate embeddings to expand at their own widths (cfg.state_embed_dim /
    afterstate_embed_dim: the kernel zero-fills) and takes the callbacks' action indices from select: same tree and policy
    output as the explicit loop that clamps the actions and pads both embeddings on the caller's side (what mctx does)"""
    from exploring_muzero_on_dog_b200 import mcts
    n, A, Cn, Es, S = 160, 4, 6, 24, 40
    Ea = Es + 2
    net4 = Net(A, Cn, Es, 5)
    g = torch.Generator(device="cuda").manual_seed(4)
    root = mcts.RootFnOutput(torch.randn(n, A, device="cuda", generator=g), torch.rand(n, device="cuda", generator=g) * 2 - 1,
                             torch.randn(n, Es, device="cuda", generator=g))
    keys = torch.randint(0, 2**31, (n, 2), device="cuda", generator=g).to(torch.uint32)
    invalid = torch.rand(n, A, device="cuda", generator=g) < 0.4
    invalid[:, 1] = False
    noise = torch.distributions.Dirichlet(torch.full((A,), 0.3, device="cuda")).sample((n,))

    def dec(params, rng_key, action, embedding):
        o = net4(action.long(), embedding)
        return mcts.DecisionRecurrentFnOutput(o["chance"], o["value"]), torch.cat([o["emb"], o["reward"][:, None], o["discount"][:, None]], 1)

    def ch(params, rng_key, outcome, afterstate):
        o = net4(outcome.long() + A, afterstate[:, :-2])
        return mcts.ChanceRecurrentFnOutput(o["prior"], o["value"], afterstate[:, -2], afterstate[:, -1]), o["emb"]

    out = mcts.stochastic_muzero_policy(None, keys, root, dec, ch, S, invalid_actions=invalid, max_depth=50, dirichlet_noise=noise)
    # the explicit loop: padded rows of width E = max(Es, Ea), clamps on the caller's side
    E = Ea
    cfg = mcts._cfg(mcts.STOCHASTIC, mcts.qtransform_by_parent_and_siblings, S, 50, A, Cn, E)
    s = mcts.Search(cfg, n)
    pad = lambda x: torch.nn.functional.pad(x, (0, E - x.shape[1]))
    s.init(keys, mcts.RootFnOutput(root.prior_logits, root.value, pad(root.embedding)), invalid, noise)
    _, action, emb, _ = s.select(0)
    for sim in range(S):
        a = action.long()
        d, after = dec(None, None, a.clamp(max=A - 1), emb[:, :Es])
        c, nxt = ch(None, None, (a - A).clamp(min=0, max=Cn - 1), emb[:, :Ea])
        step = s.expand_select if sim + 1 < S else s.expand
        step(sim, c.action_logits, c.value, c.reward, c.discount, pad(nxt), d.chance_logits, d.afterstate_value, pad(after))
    ref, _ = s.policy_output()
    assert torch.equal(out.action, ref.action) and torch.equal(out.action_weights, ref.action_weights)
    for k in ("children_visits", "children_values", "node_values", "embeddings", "children_index"):
        assert torch.equal(getattr(out.search_tree, k), getattr(s.tree, k)), k


def test_cuda_graph_replay_equals_eager_search():
    """mcts.GraphCache: the captured search replayed with new inputs gives bit-identical outputs to the eager loop, for the
    three policies (launch-bound loops are replayed as one graph)"""
    import functools
    from exploring_muzero_on_dog_b200 import mcts
    n, A, E = 96, 24, 32

    class ElemNet:  # elementwise only: bit-identical in eager and captured execution (cuBLAS may pick other kernels)
        def __init__(self, A, Cn, E, seed):
            g = torch.Generator(device="cuda").manual_seed(seed)
            self.w = torch.randn(A + Cn, E, device="cuda", generator=g)
            self.A, self.Cn = A, Cn

        def __call__(self, action, emb):
            nxt = torch.tanh(emb * self.w[action] + 0.1 * torch.roll(emb, 1, 1))
            return dict(prior=3.0 * nxt[:, :self.A], value=torch.tanh(nxt[:, 0] + nxt[:, 1] - nxt[:, 2]), reward=0.1 * nxt[:, 3],
                        discount=torch.where(nxt[:, 4] > 0, 1.0, -1.0), emb=nxt, chance=2.0 * nxt[:, 5:5 + max(self.Cn, 1)])

    net = ElemNet(A, 6, E, 3)
    net4 = ElemNet(4, 6, E, 5)

    def recurrent_fn(params, rng_key, action, embedding):
        o = net(action, embedding)
        return mcts.RecurrentFnOutput(o["reward"], o["discount"], o["prior"], o["value"]), o["emb"]

    def dec(params, rng_key, action, embedding):
        o = net4(action, embedding)
        return mcts.DecisionRecurrentFnOutput(o["chance"], o["value"]), torch.cat([o["emb"], o["reward"][:, None], o["discount"][:, None]], 1)

    def ch(params, rng_key, outcome, afterstate):
        o = net4(outcome + 4, afterstate[:, :-2])
        return mcts.ChanceRecurrentFnOutput(o["prior"], o["value"], afterstate[:, -2], afterstate[:, -1]), o["emb"]

    cache = mcts.GraphCache()
    qt = functools.partial(mcts.qtransform_completed_by_mix_value, value_scale=0.5)
    for rep in range(3):  # rep 0 captures, reps 1-2 replay with fresh inputs
        g = torch.Generator(device="cuda").manual_seed(10 + rep)
        root = mcts.RootFnOutput(torch.randn(n, A, device="cuda", generator=g), torch.rand(n, device="cuda", generator=g) * 2 - 1,
                                 torch.randn(n, E, device="cuda", generator=g))
        keys = torch.randint(0, 2**31, (n, 2), device="cuda", generator=g).to(torch.uint32)
        invalid = torch.rand(n, A, device="cuda", generator=g) < 0.5
        invalid[:, 0] = False
        for run in (lambda gc: mcts.gumbel_muzero_policy(None, keys, root, recurrent_fn, 40, invalid_actions=invalid, max_depth=20, qtransform=qt, graph_cache=gc),
                    lambda gc: mcts.muzero_policy(None, keys, root, recurrent_fn, 30, invalid_actions=invalid, max_depth=9, dirichlet_fraction=0.0, graph_cache=gc),
                    lambda gc: mcts.stochastic_muzero_policy(None, keys, mcts.RootFnOutput(root.prior_logits[:, :4].contiguous(), root.value, root.embedding),
                                                             dec, ch, 32, invalid_actions=invalid[:, :4].contiguous(), max_depth=50,
                                                             dirichlet_fraction=0.0, graph_cache=gc)):
            eager, graphed = run(None), run(cache)
            assert torch.equal(eager.action, graphed.action)
            assert torch.equal(eager.action_weights, graphed.action_weights)
            assert torch.equal(eager.search_tree.children_visits, graphed.search_tree.children_visits)
            assert torch.equal(eager.search_tree.node_values, graphed.search_tree.node_values)
    assert len(cache.entries) == 3


def test_device_exp_matches_libm():
    """the exp of the float contract, (float)exp((double)x), evaluated on the device against libm (math.exp) on softmax-like,
    uniform and boundary arguments"""
    import ctypes as C
    import math
    from exploring_muzero_on_dog_b200 import _lib
    rng = np.random.default_rng(0)
    x = np.concatenate([-np.abs(rng.standard_normal(1 << 19)).astype(np.float32) * 8, rng.uniform(-150, 5, 1 << 18).astype(np.float32),
                        np.array([0.0, -0.0, -np.inf, -87.4, -100.0, -103.9, -104.0, -150.0, -3.4028235e38, 80.0, 88.7], np.float32)])
    want = np.array([math.exp(float(v)) for v in x], np.float64).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    out = torch.empty_like(xd)
    _lib.check(_lib.lib().dogstep_exp_f32(_lib.ptr(xd), C.c_int64(x.size), _lib.ptr(out), _lib.stream()), "exp_f32")
    got = out.cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_wide_select_cache_is_consistent_at_scale():
    """config-5-shaped search (806 actions, 100 simulations) on 1,024 games with random network outputs: the per-node select
    cache the wide Gumbel kernels maintain (select_aux) must describe the dense tree arrays exactly — bitmap of the children
    with visits, visit sum / maximum, prior softmax statistics — and the usual search invariants hold."""
    from exploring_muzero_on_dog_b200 import jaxrand, mcts
    n, S, A, E = 1024, 100, 806, 32
    g = torch.Generator(device="cuda").manual_seed(5)
    rnd = lambda *shape: torch.randn(*shape, device="cuda", generator=g)
    cfg = mcts._cfg(mcts.GUMBEL, mcts.qtransform_completed_by_mix_value(value_scale=0.5), S, 50, A, 0, E)
    srch = mcts.Search(cfg, n)
    keys = jaxrand.split(jaxrand.PRNGKey(9), n)
    invalid = torch.rand(n, A, device="cuda", generator=g) < 0.3
    invalid[:, 0] = False
    srch.init(keys, mcts.RootFnOutput(rnd(n, A) * 2, torch.zeros(n, device="cuda"), rnd(n, E)), invalid, None)
    srch.select(0)
    for sim in range(S):
        step = srch.expand_select if sim + 1 < S else srch.expand
        step(sim, rnd(n, A) * 2, torch.tanh(rnd(n)), 0.1 * rnd(n), torch.where(rnd(n) > 0, 1.0, -1.0), rnd(n, E))
    t = srch.tree
    vc = t.children_visits                                            # [n, N, A]
    aux = t.select_aux.view(torch.int32)                              # [n, N + 1, 52]
    N = S + 1
    assert (vc[:, 0].sum(1) == S).all() and (vc[:, 0][invalid] == 0).all()
    assert (t.node_visits[:, 1:] >= 1).all() and (t.node_visits[:, 0] == S + 1).all()
    # visit sum / maximum per node
    assert torch.equal(aux[:, :N, 34], vc.sum(2).to(torch.int32)) and torch.equal(aux[:, :N, 35], vc.max(2).values.to(torch.int32))
    # bitmap: word l, bit j <-> child l + 32 j
    a = torch.arange(A, device="cuda")
    words = aux[:, :N, :32].to(torch.int64) & 0xFFFFFFFF
    bit = (words[:, :, a % 32] >> (a // 32)) & 1
    assert torch.equal(bit.bool(), vc > 0)
    # prior statistics of every expanded node: max and softmax denominator (float32 sums in a different order: tolerance)
    lg = t.children_prior_logits
    m1 = aux[:, :N, 32].view(torch.float32)
    s1 = aux[:, :N, 33].view(torch.float32)
    assert torch.equal(m1, lg.max(2).values)
    assert torch.allclose(s1, torch.exp(lg.double() - m1.double().unsqueeze(2)).sum(2).float(), rtol=1e-5)
    # the eight largest prior logits of every expanded node, descending, ties by index; a list may end early (index -1), never lie
    created = t.node_visits > 0
    top_v = aux[:, :N, 36:44].view(torch.float32)
    top_i = aux[:, :N, 44:52].to(torch.int64)
    srt = torch.sort(lg, dim=2, descending=True, stable=True)
    known = top_i >= 0
    assert known[created][:, :3].all()                                    # every lane ranks three: the first three are always known
    assert torch.equal(torch.where(known, top_i, srt.indices[:, :, :8])[created], srt.indices[:, :, :8][created])
    assert torch.equal(torch.where(known, top_v, srt.values[:, :, :8])[created], srt.values[:, :, :8][created])
    assert (known[:, :, :-1] | ~known[:, :, 1:])[created].all()           # unknown entries only at the tail
    assert known[created].float().mean() > 0.95
    # root_invalid bitmap and valid count in the extra slot
    rwords = aux[:, N, :32].to(torch.int64) & 0xFFFFFFFF
    rbit = (rwords[:, a % 32] >> (a // 32)) & 1
    assert torch.equal(rbit.bool(), invalid) and torch.equal(aux[:, N, 32], (~invalid).sum(1).to(torch.int32))
