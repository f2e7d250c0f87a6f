"""DOG parity tests proper: CUDA path (through the C-ABI) vs the CPU oracle / the reference's goldens, bit-exact."""
import json
import os

import numpy as np
import pytest
import torch

import oracle as O
from helpers import DOG_CASE_SETS, DOG_CODE_WINS, DOG_RULES, assert_state_equal, dog_case_rules, dog_rule_sets, mask_of

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [(name, kind, fn, i) for name, kind, fn in DOG_CASE_SETS for i in range({"test_normal_move": 52, "test_neg_move": 17,
                                                                                "test_swap_move": 14, "test_7_move": 29}[name])]


def _dog():
    from exploring_muzero_on_dog_b200.DOG import dog
    return dog


@pytest.mark.parametrize("name,kind,fn,i", CASES, ids=[f"{c[0]}-{c[3]}" for c in CASES])
def test_reference_cases(ref_cases, name, kind, fn, i):
    dog = _dog()
    c = ref_cases[f"DOG/test.py::{name}"][i]
    pins = torch.tensor(c["pins"])
    env = dog.env_reset(0, num_players=len(pins), distance=10, **dog_case_rules(c["rules"]))
    env = env.replace(pins=pins, board=dog.set_pins_on_board(env.board, pins), current_player=c["player"])
    step = {0: lambda: dog.step_normal_move(env, c["pin"], c["move"]), 1: lambda: dog.step_neg_move(env, c["pin"], c["move"]),
            2: lambda: dog.step_swap(env, c["pin"], c["pos"]), 3: lambda: dog.step_hot_7(env, torch.tensor(c["dist"]))}[kind]
    board, p, reward, done = step()
    assert p.cpu().tolist() == DOG_CODE_WINS.get((name, i), c["expected_valid"])


def _upload(dog, s, rules, num_players=4):
    env = dog.env_reset(0, num_players=num_players, seed=np.zeros(s.n, np.int32), **rules)
    f = s.fields()
    return env.replace(**{k: (v.astype(bool) if k == "done" else v) for k, v in f.items()})


def _reachable(cfg, n, seed, max_plies=400):
    rng = np.random.default_rng(seed)
    s = O.dog_reset(cfg, rng.integers(0, 1_000_000, n), -1)
    key = rng.integers(0, 2**32, 2, dtype=np.uint64).astype(np.uint32)
    chunk = n // 8
    parts = []
    for b in range(8):
        sub = O.DogState(cfg, chunk)
        for k, v in s.fields().items():
            getattr(sub, k)[...] = v[b * chunk:(b + 1) * chunk]
        O.dog_play_random(sub, key, int(rng.integers(0, max_plies)), game_offset=b * chunk, nthreads=4)
        parts.append(sub)
    out = O.DogState(cfg, chunk * 8)
    for k in out.fields():
        getattr(out, k)[...] = np.concatenate([getattr(p, k) for p in parts])
    return out


@pytest.mark.parametrize("num_players", [4, 2, 3])
def test_valid_actions_step_no_step_match_oracle(num_players):
    dog = _dog()
    rng = np.random.default_rng(40 + num_players)
    for ri, rules in enumerate(dog_rule_sets(rng, 4)):
        cfg = O.DogCfg(num_players, 0xF, 10, mask_of(rules))
        s = _reachable(cfg, 1024, 3 * ri + num_players)
        env = _upload(dog, s, rules, num_players)
        m = O.dog_valid_actions(s)
        got = dog.valid_actions(env).cpu().numpy()
        if not np.array_equal(m, got):
            g, a = np.argwhere(m != got)[0]
            raise AssertionError(f"{rules}: mask differs at game {g} action {a}: oracle {m[g, a]}")
        # half legal actions, half arbitrary indices (illegal cards / moves, play indices in the swap phase, ...)
        legal = np.array([rng.choice(np.flatnonzero(r)) if r.any() else 0 for r in m])
        act = np.where(rng.random(s.n) < 0.5, legal, rng.integers(0, 806, s.n)).astype(np.int32)
        env2, reward, done = dog.env_step(env, act)
        r, d = O.dog_step(s, act)
        assert_state_equal(s, env2.numpy())
        assert np.array_equal(r, reward.cpu().numpy()) and np.array_equal(d, done.cpu().numpy())
        env3, r0, d0 = dog.no_step(env2)
        O.dog_no_step(s)
        assert_state_equal(s, env3.numpy())
        env4 = dog.distribute_cards(env3)
        O.dog_distribute_cards(s)
        assert_state_equal(s, env4.numpy())


def test_reset_matches_oracle():
    dog = _dog()
    for rules in (DOG_RULES, dict(DOG_RULES, enable_teams=False, enable_initial_free_pin=True)):
        for sp in (0, 3, -1):
            seeds = np.arange(512, dtype=np.int32) * 1237
            env = dog.env_reset(0, seed=seeds, starting_player=sp, **rules)
            assert_state_equal(O.dog_reset(O.DogCfg(4, 0xF, 10, mask_of(rules)), seeds, sp), env.numpy())


def test_cuda_reproduces_reference_trajectories():
    dog = _dog()
    z = np.load(os.path.join(G, "dog_reference_trajectories.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    for gi, m in enumerate(meta):
        env = dog.env_reset(0, num_players=m["num_players"], distance=10, starting_player=m["starting_player"],
                            seed=np.array([m["seed"]], np.int32), **m["rules"])

        def check(t):
            got = env.numpy()
            for k in ("board", "current_player", "pins", "reward", "done", "deck", "hands", "swap_choices", "round_starter",
                      "phase", "key", "hand_size"):
                exp = z[f"g{gi}_state_{k}"][t]
                assert np.array_equal(got[k][0].astype(np.int64), np.asarray(exp).astype(np.int64)), f"game {gi} ply {t} leaf {k}"

        check(0)
        for t in range(m["plies"]):
            mask = dog.valid_actions(env).cpu().numpy()[0]
            assert np.array_equal(mask, np.unpackbits(z[f"g{gi}_mask"][t])[:mask.size].astype(bool)), f"mask game {gi} ply {t}"
            if z[f"g{gi}_kind"][t] == 1:
                _, r, d = dog.env_step(env, z[f"g{gi}_action"][t:t + 1], inplace=True)
            else:
                _, r, d = dog.no_step(env, inplace=True)
            assert int(r[0]) == int(z[f"g{gi}_reward"][t]) and bool(d[0]) == bool(z[f"g{gi}_done"][t])
            check(t + 1)


def test_play_random_matches_oracle_and_lockstep_kernel():
    dog = _dog()
    from exploring_muzero_on_dog_b200 import jaxrand
    n = 512
    key = jaxrand.split_host(jaxrand.PRNGKey(0))[1]
    seeds = jaxrand.randint(key, n, 0, 1_000_000)
    env = dog.env_reset(0, seed=seeds, **DOG_RULES)
    s = O.dog_reset(O.DogCfg(4, 0xF, 10, mask_of(DOG_RULES)), seeds.cpu().numpy(), 0)
    a = env.clone()
    total = torch.zeros(1, dtype=torch.int64, device="cuda")
    _, glen = dog.play_random(a, key, max_steps=2000, total_steps=total)
    olen, ototal, okey = O.dog_play_random(s, key, 2000, nthreads=16)
    assert_state_equal(s, a.numpy())
    assert np.array_equal(olen, glen.cpu().numpy()) and int(total.item()) == ototal
    b = env.clone()
    k = key
    T = 300
    s2 = O.dog_reset(O.DogCfg(4, 0xF, 10, mask_of(DOG_RULES)), seeds.cpu().numpy(), 0)
    O.dog_play_random(s2, key, T, nthreads=16)
    for t in range(T):
        dog.random_step(b, k)
        k = jaxrand.split_host(k, 1)[0]
    assert_state_equal(s2, b.numpy())


def test_full_size_properties():
    """BASELINE config 4 at full size (16,384 games): all games terminate; card conservation; winners own full lanes;
    board == set_pins_on_board(pins); sharding by game_offset does not change a game."""
    dog = _dog()
    from exploring_muzero_on_dog_b200 import jaxrand
    n = 16384
    key = jaxrand.split_host(jaxrand.PRNGKey(0))[1]
    seeds = jaxrand.randint(key, n, 0, 1_000_000)
    env = dog.env_reset(0, seed=seeds, **DOG_RULES)
    _, glen = dog.play_random(env, key, max_steps=2000)
    st = env.numpy()
    assert st["done"].mean() > 0.99
    fin = st["done"].astype(bool)
    lanes = (st["board"][:, 40:56] >= 0).reshape(n, 4, 4).all(-1)
    assert ((lanes[:, 0] & lanes[:, 2]) ^ (lanes[:, 1] & lanes[:, 3]))[fin].all()
    assert (st["hands"] >= 0).all() and (st["deck"] >= 0).all()
    assert (st["hands"].sum((1, 2)) + st["deck"].sum(1) <= 112).all()
    rebuilt = dog.set_pins_on_board(env.board, env.pins).cpu().numpy()
    assert np.array_equal(rebuilt, st["board"])
    half = dog.env_reset(0, seed=seeds[n // 2:].contiguous(), **DOG_RULES)
    _, glen2 = dog.play_random(half, key, max_steps=2000, game_offset=n // 2)
    assert torch.equal(glen2, glen[n // 2:]) and np.array_equal(half.numpy()["pins"], st["pins"][n // 2:])


def test_action_maps_roundtrip():
    dog = _dog()
    env = dog.env_reset(0, seed=1, **DOG_RULES)
    acts = np.arange(792)
    mv = dog.map_action_to_move(env, acts)
    assert np.array_equal(mv.numpy(), O.dog_map_action_to_move(O.DogCfg(4, 0xF, 10, mask_of(DOG_RULES)), acts))
    back = dog.map_move_to_action(env, mv)
    assert np.array_equal(back.numpy(), acts)
    assert int(dog.map_move_to_action(env, np.zeros(6, np.int64))) == 739  # DOG/speedtest.ipynb cell 16
    assert dog.get_play_action_size(env) == 792


@pytest.mark.parametrize("n", [1, 33, 200])
def test_play_random_ragged_sizes_caps_and_resume(n):
    """the persistent phase-synchronous kernel with fewer games than warps / CTAs, a step cap and resumption"""
    dog = _dog()
    from exploring_muzero_on_dog_b200 import jaxrand
    key = jaxrand.split_host(jaxrand.PRNGKey(3))[0]
    seeds = np.arange(n, dtype=np.int32) * 17 + 1
    cfg = O.DogCfg(4, 0xF, 10, mask_of(DOG_RULES))
    env = dog.env_reset(0, seed=seeds, **DOG_RULES)
    s = O.dog_reset(cfg, seeds, 0)
    for cap in (0, 25, 700, 2000):
        _, glen = dog.play_random(env, key, max_steps=cap, game_offset=3)
        olen, _, _ = O.dog_play_random(s, key, cap, game_offset=3, nthreads=8)
        assert_state_equal(s, env.numpy())
        assert np.array_equal(olen, glen.cpu().numpy())


def test_encode_board_matches_the_plane_list():
    """dog.encode_board (this repo's design: the reference has no DOG encoder) vs its NumPy twin, on states from the middle of
    random games (swap phase, play phase, finished players) for 4 / 3 / 2 players, plus the information boundary: the planes
    do not change when the OTHER seats' hands are permuted among card types"""
    dog = _dog()
    from oracle import selfplay_oracle
    from exploring_muzero_on_dog_b200 import jaxrand
    key = jaxrand.split_host(jaxrand.PRNGKey(2))[1]
    for num_players, rules in ((4, DOG_RULES), (3, dict(DOG_RULES, enable_teams=False)), (2, dict(DOG_RULES, enable_teams=False))):
        n = 256
        seeds = O.randint(key, n, 0, 1_000_000)
        cfg = O.DogCfg(num_players, 0xF, 10, mask_of(dict(rules, enable_teams=rules["enable_teams"] and num_players == 4)))
        s = O.dog_reset(cfg, seeds, -1)
        env = dog.env_reset(0, num_players=num_players, seed=seeds, starting_player=-1, **rules)
        for cap in (0, 3, 40, 400):
            O.dog_play_random(s, key, cap, nthreads=8)
            dog.play_random(env, key, max_steps=cap)
            got = dog.encode_board(env).cpu().numpy()
            exp = selfplay_oracle.dog_encode_board(s)
            assert got.shape == (n, dog.obs_planes(env), 56) and np.array_equal(got, exp), (num_players, cap, np.argwhere(got != exp)[:3].tolist())
    # information boundary (4 players): shuffling the card TYPES inside another seat's hand leaves the observation unchanged
    st = env4 = dog.env_reset(0, seed=O.randint(key, 64, 0, 1_000_000), **DOG_RULES)
    dog.play_random(env4, key, max_steps=30)
    hands = env4.raw("hands").clone()
    cur = env4.raw("current_player").long()
    other = (cur + 1) % 4
    idx = torch.arange(64, device="cuda")
    hands[idx, other] = hands[idx, other].flip(-1)
    assert torch.equal(dog.encode_board(env4), dog.encode_board(env4.replace(hands=hands)))
