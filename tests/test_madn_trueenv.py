"""The true-env mctx callbacks of deterministic MADN (MADN/deterministic_madn.py:480-590: policy_function, rollout, root_fn,
recurrent_fn): the C oracle against outputs of the reference's own functions (tests/golden/madn_det_reference_trueenv.npz, made by
tests/golden/gen_madn_trueenv_goldens.py on the jaxshim), and the CUDA path against the oracle."""
import os

import numpy as np
import pytest

import oracle as O
from helpers import RULE_BITS

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "madn_det_reference_trueenv.npz"))
LEAVES = ("board", "current_player", "pins", "reward", "done", "action_set")


def _cfgs():
    keys = [str(k) for k in G["rule_keys"]]
    return [O.MadnCfg(4, 0xF, 10, sum(RULE_BITS[k] for k, v in zip(keys, row) if v)) for row in G["rule_values"]]


def _state(cfg, idx, prefix="s_"):
    s = O.MadnState(cfg, len(idx))
    for k in LEAVES:
        v = G[prefix + k][idx]
        setattr(s, k, np.ascontiguousarray(v.astype(np.uint8 if k == "done" else np.int8)))
    return s


def _groups():
    for ri, cfg in enumerate(_cfgs()):
        idx = np.flatnonzero(G["rules"] == ri)
        yield ri, cfg, idx


def test_golden_file_covers_what_it_should():
    has = G["has_rollout"]
    assert int(G["n"]) >= 150 and has.sum() >= 120
    v = G["rollout"][has]
    assert (v == v[:, :1]).all() and set(np.unique(v)) == {-1.0, 1.0}      # four equal entries, +-1 (never 0: see the oracle)
    assert (G["s_done"] != 0).sum() >= 10                                  # finished states as well
    assert (G["policy"] >= 300).any() and (G["policy"] == 100).any()       # winning moves occur


def test_oracle_policy_function_matches_reference():
    for ri, cfg, idx in _groups():
        s = _state(cfg, idx)
        assert np.array_equal(O.madn_det_policy_function(s), G["policy"][idx]), ri


def test_oracle_root_fn_matches_reference():
    for ri, cfg, idx in _groups():
        idx = idx[G["has_rollout"][idx]]
        s = _state(cfg, idx)
        prior, value, emb = O.madn_det_root_fn(s, G["key"][idx])
        assert np.array_equal(prior, G["policy"][idx])
        assert np.array_equal(value, G["rollout"][idx][:, 0]), ri
        assert np.array_equal(emb, O.madn_det_embedding(s))


def test_oracle_recurrent_fn_matches_reference():
    for ri, cfg, idx in _groups():
        idx = idx[G["has_rollout"][idx]]
        s = _state(cfg, idx)
        prior, value, reward, discount, emb = O.madn_det_recurrent_fn(cfg, G["key"][idx], G["action"][idx], O.madn_det_embedding(s))
        assert np.array_equal(reward, G["rec_reward"][idx]) and np.array_equal(discount, G["rec_discount"][idx]), ri
        assert np.array_equal(prior, G["rec_prior"][idx]), ri
        assert np.array_equal(value, G["rec_value"][idx][:, 0]), ri
        assert np.array_equal(emb, O.madn_det_embedding(_state(cfg, idx, "n_"))), ri


# ------------------------------------------------------------------ CUDA
def _upload(dm, s, cfg_rules):
    env = dm.env_reset(0, num_players=4, seed=np.zeros(s.n, np.int32), **cfg_rules)
    return env.replace(board=s.board, pins=s.pins, current_player=s.current_player, reward=s.reward, done=s.done.astype(bool),
                       action_set=s.action_set, key=s.key)


def _rules(ri):
    keys = [str(k) for k in G["rule_keys"]]
    return {k: bool(v) for k, v in zip(keys, G["rule_values"][ri])}


@pytest.mark.gpu
def test_cuda_callbacks_match_reference_goldens():
    """policy_function / root_fn / recurrent_fn through the C-ABI against the reference's own outputs"""
    import torch
    from exploring_muzero_on_dog_b200.MADN import deterministic_madn as dm
    for ri, cfg, idx in _groups():
        s = _state(cfg, idx)
        env = _upload(dm, s, _rules(ri))
        assert np.array_equal(dm.policy_function(env).cpu().numpy(), G["policy"][idx]), ri
        has = G["has_rollout"][idx]
        keys = torch.from_numpy(G["key"][idx].astype(np.uint32)).cuda()
        root = dm.root_fn(env, keys)
        assert np.array_equal(root.prior_logits.cpu().numpy(), G["policy"][idx])
        assert np.array_equal(root.value.cpu().numpy()[has], G["rollout"][idx][has][:, 0]), ri
        assert np.array_equal(root.embedding.cpu().numpy(), O.madn_det_embedding(s))
        out, nxt = dm.make_recurrent_fn(env)(None, keys, torch.from_numpy(G["action"][idx]).cuda(), root.embedding)
        assert np.array_equal(out.reward.cpu().numpy()[has], G["rec_reward"][idx][has]), ri
        assert np.array_equal(out.discount.cpu().numpy()[has], G["rec_discount"][idx][has])
        assert np.array_equal(out.prior_logits.cpu().numpy()[has], G["rec_prior"][idx][has])
        assert np.array_equal(out.value.cpu().numpy()[has], G["rec_value"][idx][has][:, 0]), ri
        assert np.array_equal(nxt.cpu().numpy()[has], O.madn_det_embedding(_state(cfg, idx, "n_"))[has]), ri


@pytest.mark.gpu
@pytest.mark.parametrize("num_players", [4, 2])
def test_cuda_callbacks_match_oracle_on_reachable_states(num_players):
    """a wider sweep than the goldens: reachable states of random rule sets, arbitrary (also illegal / out-of-range) actions"""
    import torch
    from helpers import all_rule_sets, mask_of
    from exploring_muzero_on_dog_b200.MADN import deterministic_madn as dm
    rng = np.random.default_rng(7 + num_players)
    for rules in all_rule_sets(rng, 3):
        n = 192
        cfg = O.MadnCfg(num_players, 0xF if num_players == 4 else 0x5, 10, mask_of(rules))
        s = O.madn_reset(cfg, rng.integers(0, 1_000_000, n).astype(np.int32), -1)
        key = rng.integers(0, 2**32, 2, dtype=np.uint64).astype(np.uint32)
        for b in range(4):  # four groups at different depths of a random game
            sub = O.MadnState(cfg, n // 4)
            for f, v in s.fields().items():
                getattr(sub, f)[...] = v[b * (n // 4):(b + 1) * (n // 4)]
            O.madn_det_play_random(sub, key, int(rng.integers(0, 450)), game_offset=b * (n // 4))
            for f, v in sub.fields().items():
                getattr(s, f)[b * (n // 4):(b + 1) * (n // 4)] = v
        env = dm.env_reset(0, num_players=num_players, layout=None if num_players == 4 else [True, False, True, False],
                           seed=np.zeros(n, np.int32), **rules).replace(
            board=s.board, pins=s.pins, current_player=s.current_player, reward=s.reward, done=s.done.astype(bool),
            action_set=s.action_set, key=s.key)
        keys = rng.integers(0, 2**32, (n, 2), dtype=np.uint64).astype(np.uint32)
        action = np.where(rng.random(n) < 0.15, rng.integers(-40, 200, n), rng.integers(0, 24, n)).astype(np.int32)
        root = dm.root_fn(env, torch.from_numpy(keys).cuda())
        op, ov, oe = O.madn_det_root_fn(s, keys)
        assert np.array_equal(op, root.prior_logits.cpu().numpy()), rules
        assert np.array_equal(ov, root.value.cpu().numpy()), rules
        assert np.array_equal(oe, root.embedding.cpu().numpy())
        out, nxt = dm.make_recurrent_fn(env)(None, torch.from_numpy(keys).cuda(), torch.from_numpy(action).cuda(), root.embedding)
        rp, rv, rr, rd, re = O.madn_det_recurrent_fn(cfg, keys, action, oe)
        for a_, b_ in ((rp, out.prior_logits), (rv, out.value), (rr, out.reward), (rd, out.discount), (re, nxt)):
            assert np.array_equal(a_, b_.cpu().numpy()), rules


@pytest.mark.gpu
def test_cuda_true_env_gumbel_search_equals_oracle():
    """run_gumbel (MADN/simulate_deterministicMADN.py:12-35) on the true env: CUDA search + CUDA callbacks against the oracle search
    + oracle callbacks, simulation by simulation; then the mirror's run_gumbel gives the same action weights"""
    import torch
    from helpers import TRAIN_RULES, mask_of
    from exploring_muzero_on_dog_b200 import _lib, mcts
    from exploring_muzero_on_dog_b200.MADN import deterministic_madn as dm
    n, S = 64, 24
    rng = np.random.default_rng(3)
    cfg = O.MadnCfg(4, 0xF, 10, mask_of(TRAIN_RULES))
    s = O.madn_reset(cfg, rng.integers(0, 1_000_000, n).astype(np.int32), 0)
    O.madn_det_play_random(s, np.array([1, 2], np.uint32), 360)      # late positions: short rollouts, some games already over
    s.done[...] = 0
    env = _upload(dm, s, TRAIN_RULES)
    E = O.madn_det_embed_dim(cfg)
    keys = rng.integers(0, 2**32, (n, 2), dtype=np.uint64).astype(np.uint32)
    rkeys = rng.integers(0, 2**32, (n, 2), dtype=np.uint64).astype(np.uint32)
    invalid = ~O.madn_det_valid_action(s).reshape(n, 24)
    d = dict(policy=1, qtransform=0, num_simulations=S, max_depth=350, num_actions=24, num_chance=0, embed_dim=E,
             max_num_considered_actions=16, q_min=-1.0, q_max=1.0, value_scale=0.1, maxvisit_init=50.0, epsilon=1e-8,
             pb_c_init=1.25, pb_c_base=19652.0, dirichlet_fraction=0.0, temperature=1.0, gumbel_scale=1.0)
    search = mcts.Search(_lib.MctsCfg(**d), n)
    root = dm.root_fn(env, torch.from_numpy(rkeys).cuda())
    search.init(torch.from_numpy(keys).cuda(), root, invalid_actions=torch.from_numpy(invalid).cuda())
    op, ov, oe = O.madn_det_root_fn(s, rkeys)
    otree = O.MctsTree(O.MctsCfg(**d), n)
    O.mcts_init(otree, keys, op, ov, oe, invalid=invalid)
    rec = dm.make_recurrent_fn(env)
    for sim in range(S):
        parent, action, emb, _ = search.select(sim)
        p2, a2, e2, _ = O.mcts_select(otree, sim)
        assert np.array_equal(action.cpu().numpy(), a2) and np.array_equal(parent.cpu().numpy(), p2), sim
        out, nxt = rec(None, search.expand_key, action, emb)
        search.expand(sim, out.prior_logits, out.value, out.reward, out.discount, nxt)
        rp, rv, rr, rd, re = O.madn_det_recurrent_fn(cfg, otree.expand_key, a2, e2)
        assert np.array_equal(rv, out.value.cpu().numpy()) and np.array_equal(re, nxt.cpu().numpy()), sim
        O.mcts_expand(otree, sim, p2, a2, rp, rv, rr, rd, re)
    po, _ = search.policy_output()
    oa, ow, _ = O.mcts_policy_output(otree)
    assert np.array_equal(po.action.cpu().numpy(), oa) and np.array_equal(po.action_weights.cpu().numpy(), ow)
    assert not invalid[np.arange(n), oa][invalid.sum(1) < 24].any()          # the chosen action is legal wherever one exists
