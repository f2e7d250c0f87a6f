"""The true-env mctx callbacks of the dice game (MADN/classic_madn.py:541-714, SURVEY 8 row b4: policy_function, rollout, root_fn,
recurrent_fn = decision node, recurrent_chance_fn): the C oracle against what the reference's own function bodies compute
(tests/golden/madn_cls_reference_trueenv.npz, made by tests/golden/gen_madn_cls_trueenv_goldens.py on the jaxshim — see its
header for the one constructor call that has to be made to go through), and the CUDA path against both."""
import os

import numpy as np
import pytest

import oracle as O
from helpers import RULE_BITS

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "madn_cls_reference_trueenv.npz"))
LEAVES = ("board", "current_player", "pins", "reward", "done", "die", "key")


def _rules(ri):
    keys = [str(k) for k in G["rule_keys"]]
    return {k: bool(v) for k, v in zip(keys, G["rule_values"][ri])}


def _cfg(ri):
    return O.MadnCfg(4, 0xF, 10, sum(RULE_BITS[k] for k, v in _rules(ri).items() if v))


def _state(cfg, idx, prefix="s_"):
    s = O.MadnState(cfg, len(idx), det=False)
    for k in LEAVES:
        v = G[prefix + k][idx]
        setattr(s, k, np.ascontiguousarray(v.astype({"done": np.uint8, "key": np.uint32}.get(k, np.int8))))
    return s


def _groups():
    for ri in range(len(G["rule_values"])):
        yield ri, _cfg(ri), np.flatnonzero(G["rules"] == ri)


def test_golden_file_covers_what_it_should():
    assert int(G["n"]) >= 150
    for k in ("root_value", "after_value", "ch_value"):
        v = G[k]
        assert (v == v[:, :1]).all() and set(np.unique(v)) == {-1.0, 1.0}, k      # four equal entries, +-1
    assert (G["s_done"] != 0).sum() >= 10 and (G["a_done"] != 0).sum() >= 10        # finished states, finishing moves
    assert (G["policy"] >= 300).any()                                               # winning moves occur
    assert (G["a_current_player"] != G["s_current_player"]).any() and (G["ch_reward"] != 0).any()
    assert np.unique(G["chance_logits"]).size == 1


def test_oracle_policy_and_root_fn_match_reference():
    for ri, cfg, idx in _groups():
        s = _state(cfg, idx)
        assert np.array_equal(O.madn_cls_policy_function(s), G["policy"][idx]), ri
        prior, value, emb = O.madn_cls_root_fn(s, G["keys"][idx][:, 0])
        assert np.array_equal(prior, G["policy"][idx])
        assert np.array_equal(value, G["root_value"][idx][:, 0]), ri
        assert np.array_equal(emb, O.madn_cls_embedding(s))


def test_oracle_decision_and_chance_fn_match_reference():
    for ri, cfg, idx in _groups():
        s = _state(cfg, idx)
        cl, av, after = O.madn_cls_decision_recurrent_fn(cfg, G["keys"][idx][:, 1], G["action"][idx], O.madn_cls_embedding(s))
        assert np.array_equal(cl, G["chance_logits"][idx]), ri
        assert np.array_equal(av, G["after_value"][idx][:, 0]), ri
        assert np.array_equal(after, O.madn_cls_embedding(_state(cfg, idx, "a_"))), ri
        al, v, r, d, nxt = O.madn_cls_chance_recurrent_fn(cfg, G["keys"][idx][:, 2], G["outcome"][idx], after)
        assert np.array_equal(al, G["ch_logits"][idx]) and np.array_equal(v, G["ch_value"][idx][:, 0]), ri
        assert np.array_equal(r, G["ch_reward"][idx]) and np.array_equal(d, G["ch_discount"][idx]), ri
        assert np.array_equal(nxt, O.madn_cls_embedding(_state(cfg, idx, "n_"))), ri


# ------------------------------------------------------------------ CUDA
def _upload(cm, s, rules):
    env = cm.env_reset(0, num_players=4, seed=np.zeros(s.n, np.int32), **rules)
    return env.replace(board=s.board, pins=s.pins, current_player=s.current_player, reward=s.reward, done=s.done.astype(bool),
                       die=s.die, key=s.key)


@pytest.mark.gpu
def test_cuda_callbacks_match_reference_goldens():
    import torch
    from exploring_muzero_on_dog_b200.MADN import classic_madn as cm
    for ri, cfg, idx in _groups():
        s = _state(cfg, idx)
        env = _upload(cm, s, _rules(ri))
        k = [torch.from_numpy(np.ascontiguousarray(G["keys"][idx][:, j]).astype(np.uint32)).cuda() for j in range(3)]
        assert np.array_equal(cm.policy_function(env).cpu().numpy(), G["policy"][idx]), ri
        root = cm.root_fn(env, k[0])
        assert np.array_equal(root.prior_logits.cpu().numpy(), G["policy"][idx])
        assert np.array_equal(root.value.cpu().numpy(), G["root_value"][idx][:, 0]), ri
        assert np.array_equal(root.embedding.cpu().numpy(), O.madn_cls_embedding(s))
        dec, after = cm.make_recurrent_fn(env)(None, k[1], torch.from_numpy(G["action"][idx]).cuda(), root.embedding)
        assert np.array_equal(dec.chance_logits.cpu().numpy(), G["chance_logits"][idx])
        assert np.array_equal(dec.afterstate_value.cpu().numpy(), G["after_value"][idx][:, 0]), ri
        assert np.array_equal(after.cpu().numpy(), O.madn_cls_embedding(_state(cfg, idx, "a_"))), ri
        ch, nxt = cm.make_recurrent_chance_fn(env)(None, k[2], torch.from_numpy(G["outcome"][idx]).cuda(), after)
        assert np.array_equal(ch.action_logits.cpu().numpy(), G["ch_logits"][idx]) and np.array_equal(ch.value.cpu().numpy(), G["ch_value"][idx][:, 0])
        assert np.array_equal(ch.reward.cpu().numpy(), G["ch_reward"][idx]) and np.array_equal(ch.discount.cpu().numpy(), G["ch_discount"][idx])
        assert np.array_equal(nxt.cpu().numpy(), O.madn_cls_embedding(_state(cfg, idx, "n_"))), ri


@pytest.mark.gpu
@pytest.mark.parametrize("num_players", [4, 3])
def test_cuda_callbacks_match_oracle_on_reachable_states(num_players):
    """a wider sweep than the goldens: reachable states of random rule sets (dice rules included), arbitrary actions / outcomes"""
    import torch
    from helpers import all_rule_sets, mask_of
    from exploring_muzero_on_dog_b200.MADN import classic_madn as cm
    rng = np.random.default_rng(17 + num_players)
    for rules in all_rule_sets(rng, 3, dice=True):
        n = 160
        cfg = O.MadnCfg(num_players, 0xF if num_players == 4 else 0x7, 10, mask_of(rules))
        s = O.madn_reset(cfg, rng.integers(0, 1_000_000, n).astype(np.int32), -1, det=False)
        for t in range(int(rng.integers(20, 420))):  # lockstep random play on the oracle
            O.madn_cls_throw_die(s)
            m = O.madn_cls_valid_action(s)
            a = np.where(m.any(1), np.argmax(m * rng.random(m.shape), 1), 0).astype(np.int8)
            stepped, skipped = s.copy(), s.copy()
            O.madn_cls_step(stepped, a)
            O.madn_cls_no_step(skipped)
            live, has = s.done == 0, m.any(1)
            for f, v in s.fields().items():
                sel = (live & has).reshape((-1,) + (1,) * (v.ndim - 1)); skp = (live & ~has).reshape((-1,) + (1,) * (v.ndim - 1))
                v[...] = np.where(sel, getattr(stepped, f), np.where(skp, getattr(skipped, f), v))
        O.madn_cls_throw_die(s)
        env = cm.env_reset(0, num_players=num_players, seed=np.zeros(n, np.int32), **rules).replace(
            board=s.board, pins=s.pins, current_player=s.current_player, reward=s.reward, done=s.done.astype(bool), die=s.die, key=s.key)
        keys = [rng.integers(0, 2**32, (n, 2), dtype=np.uint64).astype(np.uint32) for _ in range(3)]
        tk = [torch.from_numpy(k).cuda() for k in keys]
        action = np.where(rng.random(n) < 0.15, rng.integers(-9, 40, n), rng.integers(0, 4, n)).astype(np.int32)
        outcome = rng.integers(0, 6, n).astype(np.int32)
        root = cm.root_fn(env, tk[0])
        op, ov, oe = O.madn_cls_root_fn(s, keys[0])
        for a_, b_ in ((op, root.prior_logits), (ov, root.value), (oe, root.embedding)):
            assert np.array_equal(a_, b_.cpu().numpy()), rules
        dec, after = cm.make_recurrent_fn(env)(None, tk[1], torch.from_numpy(action).cuda(), root.embedding)
        cl, av, oa = O.madn_cls_decision_recurrent_fn(cfg, keys[1], action, oe)
        for a_, b_ in ((cl, dec.chance_logits), (av, dec.afterstate_value), (oa, after)):
            assert np.array_equal(a_, b_.cpu().numpy()), rules
        ch, nxt = cm.make_recurrent_chance_fn(env)(None, tk[2], torch.from_numpy(outcome).cuda(), after)
        al, v, r, d, on = O.madn_cls_chance_recurrent_fn(cfg, keys[2], outcome, oa)
        for a_, b_ in ((al, ch.action_logits), (v, ch.value), (r, ch.reward), (d, ch.discount), (on, nxt)):
            assert np.array_equal(a_, b_.cpu().numpy()), rules


@pytest.mark.gpu
def test_run_mcts_search_on_the_true_env():
    """run_mcts_search (MADN/simulate_classicMADN.py:51-76): the stochastic search runs end to end on the true-env callbacks and
    puts its visits on legal pins only"""
    import torch
    from helpers import TRAIN_RULES
    from exploring_muzero_on_dog_b200 import jaxrand
    from exploring_muzero_on_dog_b200.MADN import classic_madn as cm
    n = 48
    env = cm.env_reset(0, num_players=4, seed=np.arange(n, dtype=np.int32), enable_dice_rethrow=True, **TRAIN_RULES)
    env = cm.throw_die(env)
    out = cm.run_mcts_search(env, jaxrand.split(jaxrand.PRNGKey(4), n), num_simulations=20,
                             dirichlet_noise=torch.full((n, 4), 0.25, device="cuda"))
    valid = cm.valid_action(env).reshape(n, 4)
    w = out.action_weights
    assert torch.all(w[~valid] == 0) and torch.allclose(w.sum(1)[valid.any(1)], torch.ones(int(valid.any(1).sum()), device="cuda"))
    assert torch.all(valid[torch.arange(n), out.action.long()][valid.any(1)])


@pytest.mark.gpu
def test_throw_die_matches_the_die_values_the_reference_printed():
    """MADN/jupyter_code/test_functions.ipynb cell 1 (output committed with the reference): choice(split(PRNGKey(s))[1],
    [1..6], p = uniform) is 4, 1, 6 for s = 1, 2, 3 — the draw throw_die (classic_madn.py:230-242) makes from an env whose key
    is PRNGKey(s) (see tests/test_oracle_threefry.py for the CPU side of this known answer)."""
    from exploring_muzero_on_dog_b200 import jaxrand
    from exploring_muzero_on_dog_b200.MADN import classic_madn as cm
    keys = np.stack([np.asarray(jaxrand.PRNGKey(s), np.uint32) for s in (1, 2, 3)])
    for players in (2, 3, 4):
        env = cm.env_reset(0, num_players=players, distance=10, seed=np.array([1, 2, 3], np.int32), enable_dice_rethrow=True,
                           enable_initial_free_pin=True)
        env = env.replace(key=keys)
        env = cm.throw_die(env)
        assert env.die.cpu().tolist() == [4, 1, 6]
        assert np.array_equal(env.numpy()["key"], np.stack([jaxrand.split_host(k)[0] for k in keys]))


@pytest.mark.gpu
def test_dice_probabilities_match_the_reference_notebook():
    """the distributions the reference printed for soft-locked / free players (tests/test_oracle_madn.py), on the device"""
    from test_oracle_madn import _NOTEBOOK_BASE, notebook_dice_states
    from exploring_muzero_on_dog_b200.MADN import classic_madn as cm
    for flags, s, exp in notebook_dice_states():
        env = cm.env_reset(0, num_players=2, distance=10, layout=[True, False, True, False], seed=np.array([0], np.int32),
                           **dict(_NOTEBOOK_BASE, **flags))
        env = env.replace(pins=s.pins, board=s.board)
        got = cm.dice_probabilities(env).cpu().numpy().reshape(6)
        assert np.array_equal(got, O.madn_cls_dice_probabilities(s)[0])
        assert np.array_equal(np.round(got.astype(np.float64), 8), np.round(exp.astype(np.float64), 8))
