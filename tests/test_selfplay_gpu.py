"""Self-play loop (play_batch_of_games): the CUDA mirror against the NumPy restatement with the SAME supplied search
results, then the full loop with the tree search and a torch network checked by trajectory-replay consistency."""
import numpy as np
import pytest
import torch

import oracle as O
from helpers import TRAIN_RULES, assert_state_equal, mask_of
from oracle import selfplay_oracle

pytestmark = pytest.mark.gpu


def _fake_search(A):
    """deterministic stand-in for the search: picks the first legal action by a hash of the step key"""
    def on_host(step_keys, obs, invalid):
        n = invalid.shape[0]
        h = (step_keys[:, 0].astype(np.uint64) * 2654435761 + step_keys[:, 1]) % (2**31)
        score = ((h[:, None] + np.arange(A)[None, :] * 40503) % 1009).astype(np.float32)
        score[invalid] = -1
        action = score.argmax(1).astype(np.int32)
        w = np.where(invalid, 0.0, score + 1).astype(np.float32)
        w = (w / np.maximum(w.sum(1, keepdims=True), 1)).astype(np.float32)
        value = ((h % 2001).astype(np.float32) / 1000.0 - 1.0).astype(np.float32)
        return action, w, value

    def on_device(params, step_keys, obs, invalid):
        a, w, v = on_host(step_keys.cpu().numpy(), None, invalid.cpu().numpy())
        return torch.as_tensor(a, device="cuda"), torch.as_tensor(w, device="cuda"), torch.as_tensor(v, device="cuda")
    return on_host, on_device


@pytest.mark.parametrize("det", [True, False])
def test_loop_bookkeeping_matches_oracle(det):
    from exploring_muzero_on_dog_b200 import game_agent, jaxrand
    from exploring_muzero_on_dog_b200.MADN import classic_madn as cm, deterministic_madn as dm
    n, max_steps = 256, 90
    rules = dict(TRAIN_RULES)
    if not det:
        rules["enable_dice_rethrow"] = True
    key = jaxrand.split_host(jaxrand.PRNGKey(3))[1]
    seeds = O.randint(key, n, 0, 1_000_000)
    host_fn, dev_fn = _fake_search(24 if det else 4)
    s = O.madn_reset(O.MadnCfg(4, 0xF, 10, mask_of(rules)), seeds, 0, det=det)
    exp = selfplay_oracle.play_batch_of_games(s, max_steps, key, host_fn, teams=True)
    mod = dm if det else cm
    envs = mod.env_reset(0, seed=seeds, **rules)
    shape = (34, 56) if det else (11, 56)
    got = game_agent.play_batch_of_games(envs, n, shape, None, key, 0, 0, max_steps, 1.0, search_fn=dev_fn)
    assert_state_equal(s, envs.numpy())
    for k, v in exp.items():
        assert np.array_equal(got[k].cpu().numpy(), v), k


def test_full_selfplay_with_tree_search_is_replay_consistent():
    """config-5-shaped slice on det MADN: Gumbel search (16 sims) with a torch network; the recorded actions replayed through
    the CPU oracle env must reproduce players / reward classes / discount classes / masks and the final states."""
    from exploring_muzero_on_dog_b200 import game_agent, jaxrand, mcts
    n, max_steps, E, A = 128, 60, 32, 24
    g = torch.Generator(device="cuda").manual_seed(1)
    Wr = torch.randn(34 * 56, E, device="cuda", generator=g) * 0.05
    Wp, Wv = torch.randn(E, A, device="cuda", generator=g), torch.randn(E, device="cuda", generator=g)
    Wd = torch.randn(A, E, E, device="cuda", generator=g) * 0.4

    def root_fn(params, obs):
        e = torch.tanh(obs.reshape(obs.shape[0], -1) @ Wr)
        return mcts.RootFnOutput(e @ Wp, torch.tanh(e @ Wv), e)

    def recurrent_fn(params, rng, action, emb):
        e = torch.tanh(torch.einsum("nij,nj->ni", Wd[action], emb))
        return mcts.RecurrentFnOutput(0.1 * e[:, 0], torch.where(e[:, 1] > 0, 1.0, -1.0), e @ Wp, torch.tanh(e @ Wv)), e

    key = jaxrand.PRNGKey(11)
    buf = game_agent.play_n_games_v3(None, key, (34, 56), n, 16, 8, max_steps, 1.0, root_fn=root_fn, recurrent_fn=recurrent_fn,
                                     obs_dtype=torch.int8)
    b = {k: v.cpu().numpy() for k, v in buf.items()}
    assert (b["idx"] == max_steps).all() or (b["idx"] <= max_steps).all()
    sub = jaxrand.split_host(key)[1]
    seeds = O.randint(sub, n, 0, 1_000_000)
    s = O.madn_reset(O.MadnCfg(4, 0xF, 10, mask_of(TRAIN_RULES)), seeds, 0)
    for t in range(int(b["idx"].max())):
        live = (s.done == 0) & (t < b["idx"])
        valid = O.madn_det_valid_action(s).reshape(n, -1)
        obs = O.madn_det_encode_board(s)
        has = valid.any(1)
        act = b["act"][:, t]
        assert (b["mask"][live, t] == has[live]).all() and (act[live & ~has] == -1).all()
        assert valid[live & has, act[live & has]].all()                       # only legal actions are played
        assert (b["player"][live, t] == s.current_player[live]).all() and (b["team"][live, t] == s.current_player[live] % 2).all()
        assert np.array_equal(b["obs"][live & has, t], obs[live & has])
        assert np.allclose(b["pol"][live & has, t].sum(1), 1.0, atol=1e-5) and (b["pol"][live & has, t][~valid[live & has]] < 1e-30).all()
        stepped, skipped = s.copy(), s.copy()
        a = np.where(has, act, 0)
        r, d = O.madn_det_step(stepped, np.stack([a // 6, a % 6 + 1], 1).astype(np.int8))
        O.madn_det_no_step(skipped)
        rew_t = np.where(d & (r > 0), 2, np.where(d & (r < 0), 0, 1))
        same = (s.current_player % 2) == (stepped.current_player % 2)
        disc_t = np.where(d, 1, np.where(same, 2, 0))
        sel = live & has
        assert (b["rew"][sel, t] == rew_t[sel]).all() and (b["discount"][sel, t] == disc_t[sel]).all()
        for f in s.fields():
            cur = getattr(s, f)
            cur[live & has] = getattr(stepped, f)[live & has]
            cur[live & ~has] = getattr(skipped, f)[live & ~has]


def test_dog_loop_bookkeeping_matches_oracle():
    """config 5's loop on the DOG env (806 actions, swap phase, re-deals, no_step discards) with a supplied search"""
    from helpers import DOG_RULES
    from exploring_muzero_on_dog_b200 import game_agent, jaxrand
    from exploring_muzero_on_dog_b200.DOG import dog as dg
    n, max_steps = 64, 1300
    key = jaxrand.split_host(jaxrand.PRNGKey(5))[1]
    seeds = O.randint(key, n, 0, 1_000_000)
    cfg = O.DogCfg(4, 0xF, 10, mask_of(DOG_RULES))
    host_fn, dev_fn = _fake_search(cfg.num_actions)
    s = O.dog_reset(cfg, seeds, 0)
    exp = selfplay_oracle.play_batch_of_games_dog(s, max_steps, key, host_fn, teams=True)
    envs = dg.env_reset(0, seed=seeds, **DOG_RULES)
    got = game_agent.play_batch_of_games(envs, n, (dg.RAW_OBS_SIZE,), None, key, 0, 0, max_steps, 1.0, search_fn=dev_fn)
    assert_state_equal(s, envs.numpy())
    assert (exp["mask"] == 0).any() and (exp["rew"] != 1).any() and s.done.any()   # no_step and game ends are exercised
    for k, v in exp.items():
        assert np.array_equal(got[k].cpu().numpy(), v), k


def test_dog_selfplay_with_gumbel_search_plays_legal_moves():
    """config-5-shaped slice: Gumbel search over DOG's 806 actions with a torch stand-in network; the recorded actions replayed
    through the CPU oracle must be legal and reproduce players / masks / reward classes and the final state"""
    from helpers import DOG_RULES
    from exploring_muzero_on_dog_b200 import game_agent, jaxrand, mcts
    n, max_steps, E, A = 64, 40, 32, 806
    g = torch.Generator(device="cuda").manual_seed(2)
    Wr = torch.randn(74, E, device="cuda", generator=g) * 0.2
    Wp, Wv = torch.randn(E, A, device="cuda", generator=g), torch.randn(E, device="cuda", generator=g)
    Wa, Wd = torch.randn(A, E, device="cuda", generator=g), torch.randn(E, E, device="cuda", generator=g) * 0.4

    def root_fn(params, obs):
        e = torch.tanh(obs.reshape(obs.shape[0], -1) @ Wr)
        return mcts.RootFnOutput(e @ Wp, torch.tanh(e @ Wv), e)

    def recurrent_fn(params, rng, action, emb):
        e = torch.tanh(emb @ Wd + Wa[action])
        return mcts.RecurrentFnOutput(0.1 * e[:, 0], torch.where(e[:, 1] > 0, 1.0, -1.0), e @ Wp, torch.tanh(e @ Wv)), e

    key = jaxrand.PRNGKey(13)
    envs, buf = game_agent.play_n_dog_games(None, key, n, 12, 6, max_steps, 1.0, root_fn=root_fn, recurrent_fn=recurrent_fn)
    b = {k: v.cpu().numpy() for k, v in buf.items()}
    sub = jaxrand.split_host(key)[1]
    seeds = O.randint(sub, n, 0, 1_000_000)
    s = O.dog_reset(O.DogCfg(4, 0xF, 10, mask_of(DOG_RULES)), seeds, 0)
    for t in range(int(b["idx"].max())):
        live = (s.done == 0) & (t < b["idx"])
        valid = O.dog_valid_actions(s)
        has = valid.any(1)
        act = b["act"][:, t]
        assert (b["mask"][live, t] == has[live]).all() and (act[live & ~has] == -1).all()
        assert valid[live & has, act[live & has]].all()
        assert (b["player"][live, t] == s.current_player[live]).all()
        assert np.array_equal(b["obs"][live & has, t], selfplay_oracle.dog_raw_observation(s)[live & has])
        assert np.allclose(b["pol"][live & has, t].sum(1), 1.0, atol=1e-4) and (b["pol"][live & has, t][~valid[live & has]] < 1e-30).all()
        stepped, skipped = s.copy(), s.copy()
        r, d = O.dog_step(stepped, np.where(has, act, 0))
        O.dog_no_step(skipped)
        rew_t = np.where(d & (r > 0), 2, np.where(d & (r < 0), 0, 1))
        sel = live & has
        assert (b["rew"][sel, t] == rew_t[sel]).all()
        for f in s.fields():
            cur = getattr(s, f)
            cur[live & has] = getattr(stepped, f)[live & has]
            cur[live & ~has] = getattr(skipped, f)[live & ~has]
    assert_state_equal(s, envs.numpy())


def test_split_chain_equals_host_key_chain():
    """dogstep_random_split_chain: rng_key, *step_keys = split(rng_key, n + 1) with the loop key on the device"""
    import ctypes as C
    from exploring_muzero_on_dog_b200 import _lib, jaxrand
    n = 1000
    key = jaxrand.split_host(jaxrand.PRNGKey(42))[1]
    dkey = torch.from_numpy(key.copy()).cuda()
    out = torch.empty((n, 2), dtype=torch.uint32, device="cuda")
    for it in range(5):
        _lib.check(_lib.lib().dogstep_random_split_chain(_lib.ptr(dkey), C.c_int64(n), _lib.ptr(out), _lib.stream()), "split_chain")
        exp = O.split(key, n + 1)
        assert np.array_equal(out.cpu().numpy(), exp[1:]) and np.array_equal(dkey.cpu().numpy(), exp[0])
        key = exp[0]


@pytest.mark.parametrize("env_kind", ["det", "cls", "dog"])
def test_graph_replayed_loop_equals_eager_loop_and_oracle(env_kind):
    """SelfPlayLoop with cuda_graph=True (one captured iteration replayed, lagging termination poll, run twice on the same
    buffers) == the eager loop == the NumPy restatement, with a capturable device-side stand-in for the search"""
    from helpers import DOG_RULES
    from exploring_muzero_on_dog_b200 import game_agent, jaxrand
    from exploring_muzero_on_dog_b200.DOG import dog as dg
    from exploring_muzero_on_dog_b200.MADN import classic_madn as cm, deterministic_madn as dm
    n = 96
    max_steps = {"det": 120, "cls": 120, "dog": 260}[env_kind]
    key = jaxrand.split_host(jaxrand.PRNGKey(17))[1]
    seeds = O.randint(key, n, 0, 1_000_000)
    if env_kind == "dog":
        rules, mod, A, shape = DOG_RULES, dg, 806, (dg.RAW_OBS_SIZE,)
        ocfg = O.DogCfg(4, 0xF, 10, mask_of(rules))
    else:
        rules = dict(TRAIN_RULES, **({} if env_kind == "det" else {"enable_dice_rethrow": True}))
        mod, A, shape = (dm, 24, (34, 56)) if env_kind == "det" else (cm, 4, (11, 56))
        ocfg = O.MadnCfg(4, 0xF, 10, mask_of(rules))
    ar = np.arange(A, dtype=np.int64)

    def host_fn(step_keys, obs, invalid):  # integer hash of the step key: exact on both sides
        h = (step_keys[:, 0].astype(np.int64) * 7 + step_keys[:, 1].astype(np.int64)) % 1000003
        score = (h[:, None] + ar[None, :] * 40503) % 1009
        score[invalid] = -1
        action = score.argmax(1).astype(np.int32)
        w = np.where(invalid, 0.0, 1.0).astype(np.float32)
        w = w / np.maximum(w.sum(1, keepdims=True), 1)
        value = ((h % 2001) - 1000).astype(np.float32) / np.float32(1024.0)  # power-of-two scale: exact however it is evaluated
        return action, w.astype(np.float32), value

    ar_d = torch.arange(A, device="cuda", dtype=torch.int64)

    def dev_fn(params, step_keys, obs, invalid):
        k = step_keys.to(torch.int64)
        h = (k[:, 0] * 7 + k[:, 1]) % 1000003
        score = (h[:, None] + ar_d[None, :] * 40503) % 1009
        score = torch.where(invalid, torch.full_like(score, -1), score)
        action = score.argmax(1).to(torch.int32)
        w = (~invalid).to(torch.float32)
        w = w / w.sum(1, keepdim=True).clamp(min=1)
        value = ((h % 2001) - 1000).to(torch.float32) / 1024.0
        return action, w, value

    def reset(out=None):
        return mod.env_reset(0, seed=seeds, out=out, **rules)

    if env_kind == "dog":
        s = O.dog_reset(ocfg, seeds, 0)
        exp = selfplay_oracle.play_batch_of_games_dog(s, max_steps, key, host_fn, teams=True)
    else:
        s = O.madn_reset(ocfg, seeds, 0, det=env_kind == "det")
        exp = selfplay_oracle.play_batch_of_games(s, max_steps, key, host_fn, teams=True)
    eager_env = reset()
    eager = game_agent.play_batch_of_games(eager_env, n, shape, None, key, 0, 0, max_steps, 1.0, search_fn=dev_fn)
    assert_state_equal(s, eager_env.numpy())
    for k, v in exp.items():
        assert np.array_equal(eager[k].cpu().numpy(), v), ("eager", k)
    genv = reset()
    loop = game_agent.SelfPlayLoop(genv, n, shape, None, max_steps, search_fn=dev_fn, cuda_graph=True, lookahead=3)
    for rep in range(2):  # the second run reuses the buffers and the captured graph on the re-seeded env
        if rep:
            reset(out=genv)
        got = loop.run(key)
        assert_state_equal(s, genv.numpy())
        for k, v in exp.items():
            assert np.array_equal(got[k].cpu().numpy(), v), ("graph", rep, k)
        assert loop.iterations == int(exp["idx"].max()) or not s.done.all()
        assert loop.enqueued <= min(max_steps, loop.iterations + 2 * 3 + 1)
