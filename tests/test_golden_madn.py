"""Goldens produced by the REFERENCE's own code (tests/golden/gen_madn_goldens.py ran
/root/reference/MADN/*.py on the jaxshim) replayed through the C oracle (CPU) and the CUDA path (GPU)."""
import json
import os

import numpy as np
import pytest

import oracle as O
from helpers import mask_of

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LEAVES_DET = ("board", "current_player", "pins", "reward", "done", "key", "action_set")
LEAVES_CLS = ("board", "current_player", "pins", "reward", "done", "key", "die")


def _load(name):
    z = np.load(os.path.join(G, f"madn_{name}_reference_trajectories.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


def _check_leaves(got, z, gi, t, leaves, where):
    for k in leaves:
        exp = z[f"g{gi}_state_{k}"][t]
        a = np.asarray(got[k][0]).astype(np.int64)
        assert np.array_equal(a, np.asarray(exp).astype(np.int64)), f"{where}: game {gi} ply {t} leaf {k}: got {a.tolist()} want {exp.tolist()}"


def _replay_oracle(det):
    z, meta = _load("det" if det else "cls")
    leaves = LEAVES_DET if det else LEAVES_CLS
    plies = 0
    for gi, m in enumerate(meta):
        cfg = O.MadnCfg(m["num_players"], 0xF, 10, mask_of(m["rules"]))
        s = O.madn_reset(cfg, [m["seed"]], m["starting_player"], det=det)
        _check_leaves(s.fields(), z, gi, 0, leaves, "reset")
        for t in range(m["plies"]):
            if not det:
                assert np.array_equal(O.madn_cls_dice_probabilities(s)[0], z[f"g{gi}_dice_probs"][t])
                O.madn_cls_throw_die(s)
            mask = (O.madn_det_valid_action(s) if det else O.madn_cls_valid_action(s))[0]
            assert np.array_equal(mask, z[f"g{gi}_mask"][t]), f"mask: game {gi} ply {t}"
            obs = (O.madn_det_encode_board(s) if det else O.madn_cls_encode_board(s))[0]
            assert np.array_equal(obs, z[f"g{gi}_obs"][t]), f"obs: game {gi} ply {t}"
            act = z[f"g{gi}_action"][t]
            if z[f"g{gi}_kind"][t] == 1:
                r, d = O.madn_det_step(s, [act]) if det else O.madn_cls_step(s, [act[0]])
            else:
                r, d = O.madn_det_no_step(s) if det else O.madn_cls_no_step(s)
            assert int(r[0]) == int(z[f"g{gi}_reward"][t]) and bool(d[0]) == bool(z[f"g{gi}_done"][t])
            _check_leaves(s.fields(), z, gi, t + 1, leaves, "step")
            plies += 1
    return plies


def test_oracle_reproduces_reference_trajectories_deterministic():
    assert _replay_oracle(True) > 5000


def test_oracle_reproduces_reference_trajectories_classic():
    assert _replay_oracle(False) > 5000


def test_oracle_reproduces_reference_random_driver():
    """The reference's do_random lockstep loop (jax.random.categorical on the shim, float gumbel argmax)
    versus the oracle's integer categorical: same actions, same final states, same carried key."""
    z = np.load(os.path.join(G, "madn_det_reference_random_driver.npz"))
    from helpers import TRAIN_RULES
    cfg = O.MadnCfg(4, 0xF, 10, mask_of(TRAIN_RULES))
    key = O.split(O.prng_key(0))[1]
    seeds = O.randint(key, 6, 0, 1_000_000)
    assert np.array_equal(seeds, z["seeds"])
    T = z["actions"].shape[0]
    for fg in (False, True):
        s = O.madn_reset(cfg, seeds, 0)
        glen, total, kout = O.madn_det_play_random(s, key, T, float_gumbel=fg)
        assert np.array_equal(s.pins, z["final_pins"]) and np.array_equal(s.board, z["final_board"])
        assert np.array_equal(s.action_set, z["final_action_set"]) and np.array_equal(s.current_player, z["final_current_player"])
        assert np.array_equal(kout, z["final_key"])


# ------------------------------------------------------------------------------------------- GPU
def _replay_cuda(det):
    import torch
    if det:
        from exploring_muzero_on_dog_b200.MADN import deterministic_madn as mod
    else:
        from exploring_muzero_on_dog_b200.MADN import classic_madn as mod
    z, meta = _load("det" if det else "cls")
    leaves = LEAVES_DET if det else LEAVES_CLS
    for gi, m in enumerate(meta):
        env = mod.env_reset(0, num_players=m["num_players"], distance=10, starting_player=m["starting_player"],
                            seed=np.array([m["seed"]], np.int32), **m["rules"])
        _check_leaves(env.numpy(), z, gi, 0, leaves, "reset")
        stride = 1 if m["plies"] <= 300 else 1
        for t in range(m["plies"]):
            if not det:
                if t % 16 == 0:
                    assert np.array_equal(mod.dice_probabilities(env).cpu().numpy()[0], z[f"g{gi}_dice_probs"][t])
                mod.throw_die(env, inplace=True)
            mask = mod.valid_action(env).cpu().numpy()[0]
            assert np.array_equal(mask, z[f"g{gi}_mask"][t]), f"mask: game {gi} ply {t}"
            if t % 16 == 0:
                assert np.array_equal(mod.encode_board(env).cpu().numpy()[0], z[f"g{gi}_obs"][t]), f"obs: game {gi} ply {t}"
            act = z[f"g{gi}_action"][t]
            if z[f"g{gi}_kind"][t] == 1:
                _, r, d = mod.env_step(env, act[None] if det else act[:1], inplace=True)
            else:
                _, r, d = mod.no_step(env, inplace=True)
            assert int(r[0]) == int(z[f"g{gi}_reward"][t]) and bool(d[0]) == bool(z[f"g{gi}_done"][t])
            _check_leaves(env.numpy(), z, gi, t + 1, leaves, "step")


@pytest.mark.gpu
def test_cuda_reproduces_reference_trajectories_deterministic():
    _replay_cuda(True)


@pytest.mark.gpu
def test_cuda_reproduces_reference_trajectories_classic():
    _replay_cuda(False)


@pytest.mark.gpu
def test_cuda_reproduces_reference_random_driver():
    import torch
    from exploring_muzero_on_dog_b200 import jaxrand
    from exploring_muzero_on_dog_b200.MADN import deterministic_madn as dm
    from helpers import TRAIN_RULES
    z = np.load(os.path.join(G, "madn_det_reference_random_driver.npz"))
    key = jaxrand.split_host(jaxrand.PRNGKey(0))[1]
    seeds = jaxrand.randint(key, 6, 0, 1_000_000)
    assert np.array_equal(seeds.cpu().numpy(), z["seeds"])
    env = dm.env_reset(0, seed=seeds, **TRAIN_RULES)
    dm.play_random(env, key, max_steps=z["actions"].shape[0])
    got = env.numpy()
    assert np.array_equal(got["pins"], z["final_pins"]) and np.array_equal(got["board"], z["final_board"])
    assert np.array_equal(got["action_set"], z["final_action_set"])
    assert np.array_equal(got["current_player"], z["final_current_player"])
