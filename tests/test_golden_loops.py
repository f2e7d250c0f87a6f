"""The loops around the env, pinned to the REFERENCE'S OWN LOOP CODE (tests/golden/gen_loop_goldens.py ran
play_eval_loop_jitted and play_batch_of_games_jitted of the reference on the jaxshim, with the search replaced by a deterministic
stand-in): oracle/eval_oracle.py and oracle/selfplay_oracle.py must reproduce every recorded buffer / winner / final leaf
(CPU), and so must the CUDA loops (GPU)."""
import json
import os

import numpy as np
import pytest

import oracle as O
from helpers import mask_of
from oracle import eval_oracle, selfplay_oracle

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loops_reference.npz"))
BUF_KEYS = ("obs", "act", "rew", "val", "pol", "mask", "player", "team", "discount", "idx")


def _rules(tag):
    return json.loads(bytes(Z[f"{tag}_rules"]).decode())


def _host_search(A, split_first):
    """the stand-in search of the generator (an integer hash of the key the loop hands to the search)"""
    ar = np.arange(A, dtype=np.int64)

    def fn(step_keys, obs, invalid):
        keys = np.stack([O.split(k)[1] for k in step_keys]) if split_first else step_keys   # stochastic loop: key1, key2 = split(key)
        h = (keys[:, 0].astype(np.int64) * 7 + keys[:, 1].astype(np.int64)) % 1000003
        score = (h[:, None] + ar[None, :] * 40503) % 1009
        score[invalid] = -1
        action = score.argmax(1).astype(np.int32)
        w = np.where(invalid, 0.0, 1.0).astype(np.float32)
        w = (w / np.maximum(w.sum(1, keepdims=True), 1).astype(np.float32)).astype(np.float32)
        value = ((h % 2001) - 1000).astype(np.float32) / np.float32(1024.0)
        return action, w, value
    return fn


def _check_buffers(tag, ci, got, to_np=np.asarray):
    keys = BUF_KEYS + (("dice", "dice_dist") if "cls" in tag else ())
    for k in keys:
        exp = Z[f"{tag}_{ci}_buf_{k}"]
        g = to_np(got[k])
        assert g.shape == exp.shape, (tag, ci, k, g.shape, exp.shape)
        assert np.array_equal(g.astype(np.float64), exp.astype(np.float64)), (tag, ci, k, np.argwhere(g.astype(np.float64) != exp.astype(np.float64))[:3].tolist())


@pytest.mark.parametrize("tag", ["selfplay_det", "selfplay_cls"])
@pytest.mark.parametrize("ci", [0, 1])
def test_selfplay_oracle_reproduces_the_reference_loop(tag, ci):
    det = tag.endswith("det")
    rules = _rules(tag)
    s = O.madn_reset(O.MadnCfg(4, 0xF, 10, mask_of(rules)), Z[f"{tag}_{ci}_seeds"], 0, det=det)
    buf = selfplay_oracle.play_batch_of_games(s, int(Z[f"{tag}_{ci}_max_steps"]), Z[f"{tag}_{ci}_key"], _host_search(24 if det else 4, not det),
                                              teams=rules["enable_teams"])
    _check_buffers(tag, ci, buf)
    assert (buf["mask"] == 0).any() and buf["idx"].max() > 100          # skipped turns are exercised
    if ci == 1:
        assert s.done.all() and (buf["rew"] != 1).any()                 # games end inside the recording: terminal reward classes


@pytest.mark.parametrize("tag", ["eval_det", "eval_cls"])
@pytest.mark.parametrize("ci", [0, 1])
def test_eval_oracle_reproduces_the_reference_loop(tag, ci):
    det = tag.endswith("det")
    rules = _rules(tag)
    seeds = Z[f"{tag}_{ci}_seeds"]
    s = O.madn_reset(O.MadnCfg(4, 0xF, 10, mask_of(rules)), seeds, 0, det=det)
    s.current_player[:] = np.repeat(np.arange(4), len(seeds) // 4)       # batch_reset(seeds, repeat(arange(4), num_envs))
    loop = eval_oracle.play_eval_loop if det else eval_oracle.play_eval_loop_classic
    winners = loop(s, [int(t) for t in Z[f"{tag}_{ci}_types"]], Z[f"{tag}_{ci}_key"])
    assert np.array_equal(winners, Z[f"{tag}_{ci}_winners"])
    for k, v in s.fields().items():
        assert np.array_equal(np.asarray(v).astype(np.int64), Z[f"{tag}_{ci}_final_{k}"].astype(np.int64)), (tag, ci, k)
    assert s.done.all()


# ----------------------------------------------------------------------------------------------------------------- GPU
def _dev_search(A, split_first):
    import torch
    from exploring_muzero_on_dog_b200 import game_agent
    ar = torch.arange(A, device="cuda", dtype=torch.int64)

    def fn(params, step_keys, obs, invalid):
        k = (game_agent._split_each(step_keys, 1) if split_first else step_keys).to(torch.int64)
        h = (k[:, 0] * 7 + k[:, 1]) % 1000003
        score = (h[:, None] + ar[None, :] * 40503) % 1009
        score = torch.where(invalid, torch.full_like(score, -1), score)
        action = score.argmax(1).to(torch.int32)
        w = (~invalid).to(torch.float32)
        w = w / w.sum(1, keepdim=True).clamp(min=1)
        value = ((h % 2001) - 1000).to(torch.float32) / 1024.0
        return action, w, value
    return fn


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["selfplay_det", "selfplay_cls"])
@pytest.mark.parametrize("ci", [0, 1])
@pytest.mark.parametrize("graph", [False, True])
def test_cuda_selfplay_loop_reproduces_the_reference_loop(tag, ci, graph):
    from exploring_muzero_on_dog_b200 import game_agent
    from exploring_muzero_on_dog_b200.MADN import classic_madn as cm, deterministic_madn as dm
    det = tag.endswith("det")
    rules = _rules(tag)
    seeds = Z[f"{tag}_{ci}_seeds"]
    envs = (dm if det else cm).env_reset(0, seed=seeds, **rules)
    got = game_agent.play_batch_of_games(envs, len(seeds), (34, 56) if det else (11, 56), None, Z[f"{tag}_{ci}_key"], 16, 8,
                                         int(Z[f"{tag}_{ci}_max_steps"]), 1.0, search_fn=_dev_search(24 if det else 4, not det),
                                         cuda_graph=graph)
    _check_buffers(tag, ci, got, to_np=lambda t: t.cpu().numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["eval_det", "eval_cls"])
@pytest.mark.parametrize("ci", [0, 1])
def test_cuda_eval_loop_reproduces_the_reference_loop(tag, ci):
    import torch
    from exploring_muzero_on_dog_b200 import evaluate_agent as ea
    from exploring_muzero_on_dog_b200.MADN import classic_madn as cm, deterministic_madn as dm
    det = tag.endswith("det")
    rules = _rules(tag)
    seeds = Z[f"{tag}_{ci}_seeds"]
    n = len(seeds)
    envs = (dm if det else cm).env_reset(0, seed=seeds, **rules)
    envs.raw("current_player").copy_(torch.as_tensor(np.repeat(np.arange(4), n // 4), dtype=torch.int8, device="cuda"))
    params = tuple({"type": int(t)} for t in Z[f"{tag}_{ci}_types"])
    _, winners = ea.play_eval_loop(envs, params, Z[f"{tag}_{ci}_key"], n)
    assert np.array_equal(winners.cpu().numpy(), Z[f"{tag}_{ci}_winners"])
    got = envs.numpy()
    for k in got:
        assert np.array_equal(got[k].astype(np.int64), Z[f"{tag}_{ci}_final_{k}"].astype(np.int64)), (tag, ci, k)
