"""Parity tests proper: the CUDA path (through the C-ABI) against the CPU oracle, bit-exact."""
import numpy as np
import pytest
import torch

import oracle as O
from helpers import TRAIN_RULES, all_rule_sets, assert_state_equal, madn_case_rules, mask_of

pytestmark = pytest.mark.gpu


def _dm():
    from exploring_muzero_on_dog_b200.MADN import deterministic_madn as dm
    return dm


def _cm():
    from exploring_muzero_on_dog_b200.MADN import classic_madn as cm
    return cm


# ------------------------------------------------------------------ the reference's own cases, through CUDA
@pytest.mark.parametrize("i", range(64))
def test_reference_cases_deterministic(ref_cases, i):
    dm = _dm()
    c = ref_cases["MADN/test.py::test_normal_move_deterministic_MADN"][i]
    pins = torch.tensor(c["pins"])
    env = dm.env_reset(0, num_players=len(pins), distance=10, **madn_case_rules(c["rules"]))
    env = env.replace(pins=pins, board=dm.set_pins_on_board(env.board, pins), current_player=c["player"])
    env, reward, done = dm.env_step(env, [c["pin"], c["move"]])
    assert env.pins.cpu().tolist() == c["expected_valid"]


@pytest.mark.parametrize("i", range(64))
def test_reference_cases_classic(ref_cases, i):
    cm = _cm()
    c = ref_cases["MADN/test.py::test_normal_move_classic_MADN"][i]
    pins = torch.tensor(c["pins"])
    env = cm.env_reset(0, num_players=len(pins), distance=10, **madn_case_rules(c["rules"]))
    env = env.replace(pins=pins, board=cm.set_pins_on_board(env.board, pins), current_player=c["player"])
    env = cm.set_die(env, c["move"])
    valid_moves = cm.valid_action(env)
    env, reward, done = cm.env_step(env, c["pin"])
    assert bool(valid_moves[c["pin"]]) or int(reward) == -1
    assert env.pins.cpu().tolist() == c["expected_valid"]


# ------------------------------------------------------------------ differential tests on reachable states
def _upload_det(dm, s, rules, num_players=4):
    env = dm.env_reset(0, num_players=num_players, seed=np.zeros(s.n, np.int32), **rules)
    return env.replace(board=s.board, pins=s.pins, current_player=s.current_player, reward=s.reward,
                       done=s.done.astype(bool), action_set=s.action_set, key=s.key)


def _reachable_det(cfg, n, seed, max_plies=500):
    """oracle states after a random number of random-policy plies (mix of early / mid / late / finished games)"""
    rng = np.random.default_rng(seed)
    s = O.madn_reset(cfg, rng.integers(0, 1_000_000, n), -1)
    key = rng.integers(0, 2**32, 2, dtype=np.uint64).astype(np.uint32)
    parts = []
    chunk = n // 8
    for b in range(8):
        sub = O.MadnState(cfg, chunk)
        for f, v in s.fields().items():
            getattr(sub, f)[...] = v[b * chunk:(b + 1) * chunk]
        O.madn_det_play_random(sub, key, int(rng.integers(0, max_plies)), game_offset=b * chunk)
        parts.append(sub)
    out = O.MadnState(cfg, chunk * 8)
    for f in out.fields():
        getattr(out, f)[...] = np.concatenate([getattr(p, f) for p in parts])
    return out


@pytest.mark.parametrize("num_players", [4, 3, 2])
def test_det_valid_action_and_step_match_oracle(num_players):
    dm = _dm()
    rng = np.random.default_rng(100 + num_players)
    for ri, rules in enumerate(all_rule_sets(rng, 6)):
        cfg = O.MadnCfg(num_players, 0xF, 10, mask_of(rules))
        s = _reachable_det(cfg, 4096, 7 * ri + num_players)
        env = _upload_det(dm, s, rules, num_players)
        assert np.array_equal(O.madn_det_valid_action(s), dm.valid_action(env).cpu().numpy()), rules
        # random actions: mostly arbitrary (many invalid), some taken from the legal mask
        act = np.stack([rng.integers(0, 4, s.n), rng.integers(1, 7, s.n)], 1).astype(np.int8)
        env2, reward, done = dm.env_step(env, act)
        r, d = O.madn_det_step(s, act)
        assert_state_equal(s, env2.numpy())
        assert np.array_equal(r, reward.cpu().numpy()) and np.array_equal(d, done.cpu().numpy())
        # a second step from the new states with no_step mixed in
        env3, r0, d0 = dm.no_step(env2)
        O.madn_det_no_step(s)
        assert_state_equal(s, env3.numpy())
        assert int(r0.abs().sum()) == 0


def test_det_step_garbage_actions_follow_jax_index_rules():
    """pin / move outside their ranges: gathers clamp (negative wraps once), scatters drop (Appendix A.0)."""
    dm = _dm()
    cfg = O.MadnCfg(4, 0xF, 10, mask_of(TRAIN_RULES))
    s = _reachable_det(cfg, 2048, 3)
    env = _upload_det(dm, s, TRAIN_RULES)
    rng = np.random.default_rng(5)
    act = np.stack([rng.integers(-3, 8, s.n), rng.integers(-4, 12, s.n)], 1).astype(np.int8)
    env2, reward, done = dm.env_step(env, act)
    r, d = O.madn_det_step(s, act)
    assert_state_equal(s, env2.numpy())
    assert np.array_equal(r, reward.cpu().numpy())


def test_det_reset_matches_oracle():
    dm = _dm()
    for rules in (TRAIN_RULES, dict(TRAIN_RULES, enable_initial_free_pin=False)):
        for sp in (0, 2, -1):
            seeds = np.arange(1000, dtype=np.int32) * 977
            env = dm.env_reset(0, seed=seeds, starting_player=sp, **rules)
            s = O.madn_reset(O.MadnCfg(4, 0xF, 10, mask_of(rules)), seeds, sp)
            assert_state_equal(s, env.numpy())
    env = dm.env_reset(0, seed=42, **TRAIN_RULES)
    assert env.key.cpu().numpy().tolist() == [1832780943, 270669613]
    assert env.pins.shape == (4, 4) and env.start.cpu().tolist() == [0, 10, 20, 30]


def test_det_encode_board_matches_oracle():
    dm = _dm()
    rng = np.random.default_rng(9)
    for num_players in (4, 2):
        for rules in all_rule_sets(rng, 2):
            cfg = O.MadnCfg(num_players, 0xF, 10, mask_of(rules))
            s = _reachable_det(cfg, 1024, 11)
            env = _upload_det(dm, s, rules, num_players)
            assert np.array_equal(O.madn_det_encode_board(s), dm.encode_board(env).cpu().numpy())


def test_det_play_random_matches_oracle_and_lockstep_kernel():
    """config-2 workload at a size the oracle finishes in seconds: persistent kernel == oracle ==
    one fused launch per lockstep iteration, state by state."""
    dm = _dm()
    from exploring_muzero_on_dog_b200 import jaxrand
    n = 8192
    key = jaxrand.split_host(jaxrand.PRNGKey(0))[1]
    seeds = jaxrand.randint(key, n, 0, 1_000_000)
    assert np.array_equal(seeds.cpu().numpy(), O.randint(key, n, 0, 1_000_000))
    env = dm.env_reset(0, seed=seeds, **TRAIN_RULES)
    s = O.madn_reset(O.MadnCfg(4, 0xF, 10, mask_of(TRAIN_RULES)), seeds.cpu().numpy(), 0)
    # (a) persistent kernel, sharded as two "ranks" by game_offset
    a = env.clone()
    total = torch.zeros(1, dtype=torch.int64, device="cuda")
    _, glen = dm.play_random(a, key, max_steps=2000, total_steps=total)
    olen, ototal, okey = O.madn_det_play_random(s, key, 2000, nthreads=8)
    assert_state_equal(s, a.numpy())
    assert np.array_equal(olen, glen.cpu().numpy()) and int(total.item()) == ototal
    # (b) lockstep: one launch per iteration with the host-chained key
    b = env.clone()
    active = torch.zeros(1, dtype=torch.int64, device="cuda")
    k = key
    for t in range(int(olen.max())):
        dm.random_step(b, k, active_count=active)
        k = jaxrand.split_host(k, 1)[0]
    assert_state_equal(s, b.numpy())
    assert int(active.item()) == ototal and k.tolist() == okey.tolist()


def test_det_full_size_properties():
    """BASELINE config 2 at full size (65,536 games): every game terminates, the winners own a full goal lane,
    board == set_pins_on_board(pins), and results are independent of how games are sharded."""
    dm = _dm()
    from exploring_muzero_on_dog_b200 import jaxrand
    n = 65536
    key = jaxrand.split_host(jaxrand.PRNGKey(0))[1]
    seeds = jaxrand.randint(key, n, 0, 1_000_000)
    env = dm.env_reset(0, seed=seeds, **TRAIN_RULES)
    _, glen = dm.play_random(env, key, max_steps=2000)
    st = env.numpy()
    assert st["done"].all() and 300 < glen.float().mean().item() < 500
    board = st["board"]
    lanes = (board[:, 40:56] >= 0).reshape(n, 4, 4).all(-1)
    assert ((lanes[:, 0] & lanes[:, 2]) ^ (lanes[:, 1] & lanes[:, 3])).all()
    rebuilt = dm.set_pins_on_board(env.board, env.pins).cpu().numpy()
    assert np.array_equal(rebuilt, board)
    # shard [n/2, n) played alone with its global game offset gives the same games
    half = dm.env_reset(0, seed=seeds[n // 2:].contiguous(), **TRAIN_RULES)
    _, glen2 = dm.play_random(half, key, max_steps=2000, game_offset=n // 2)
    assert torch.equal(glen2, glen[n // 2:]) and np.array_equal(half.numpy()["pins"], st["pins"][n // 2:])


# ------------------------------------------------------------------ classic (dice) MADN
def _upload_cls(cm, s, rules, num_players=4):
    env = cm.env_reset(0, num_players=num_players, seed=np.zeros(s.n, np.int32), **rules)
    return env.replace(board=s.board, pins=s.pins, current_player=s.current_player, reward=s.reward,
                       done=s.done.astype(bool), die=s.die, key=s.key)


@pytest.mark.parametrize("num_players", [4, 2])
def test_cls_lockstep_trajectories_match_oracle(num_players):
    """throw_die -> valid_action -> env_step / no_step for 300 lockstep plies, compared leaf by leaf each ply
    (dice via threefry-matched choice, policy = first legal pin or an arbitrary pin every 7th ply)."""
    cm = _cm()
    rng = np.random.default_rng(num_players)
    for rules in all_rule_sets(rng, 3, dice=True):
        n = 2048
        cfg = O.MadnCfg(num_players, 0xF, 10, mask_of(rules))
        seeds = rng.integers(0, 1_000_000, n).astype(np.int32)
        s = O.madn_reset(cfg, seeds, -1, det=False)
        env = cm.env_reset(0, num_players=num_players, seed=seeds, starting_player=-1, **rules)
        assert_state_equal(s, env.numpy())
        for t in range(300):
            cm.throw_die(env, inplace=True)
            O.madn_cls_throw_die(s)
            if t % 50 == 0:
                assert np.array_equal(O.madn_cls_dice_probabilities(s), cm.dice_probabilities(env).cpu().numpy())
                assert np.array_equal(O.madn_cls_encode_board(s), cm.encode_board(env).cpu().numpy())
            m = O.madn_cls_valid_action(s)
            assert np.array_equal(m, cm.valid_action(env).cpu().numpy()), (rules, t)
            pin = np.where(m.any(1), m.argmax(1), 0).astype(np.int8)
            if t % 7 == 0:
                pin = rng.integers(0, 4, n).astype(np.int8)
            active = m.any(1) | (t % 7 == 0)
            # oracle: step the active games, no_step the others (done on copies, merged by mask)
            s_step, s_skip = s.copy(), s.copy()
            O.madn_cls_step(s_step, pin)
            O.madn_cls_no_step(s_skip)
            for f in s.fields():
                sel = active.reshape((-1,) + (1,) * (getattr(s, f).ndim - 1))
                getattr(s, f)[...] = np.where(sel, getattr(s_step, f), getattr(s_skip, f))
            e_step, _, _ = cm.env_step(env, pin)
            e_skip, _, _ = cm.no_step(env)
            sel_t = torch.as_tensor(active, device="cuda")
            merged = {}
            for f in ("board", "pins", "current_player", "reward", "done", "die"):
                a, b = e_step.raw(f), e_skip.raw(f)
                merged[f] = torch.where(sel_t.reshape((-1,) + (1,) * (a.ndim - 1)), a, b)
            env = env.replace(**merged)
            assert_state_equal(s, env.numpy())


def test_random_helpers_match_oracle():
    from exploring_muzero_on_dog_b200 import jaxrand
    key = jaxrand.split_host(jaxrand.PRNGKey(123))[0]
    assert np.array_equal(jaxrand.split_host(key, 5), O.split(key, 5))
    assert np.array_equal(jaxrand.split(key, 1000).cpu().numpy(), O.split(key, 1000))
    assert np.array_equal(jaxrand.bits(key, 1000).cpu().numpy(), O.random_bits(key, 1000))
    assert np.array_equal(jaxrand.uniform(key, 1000).cpu().numpy(), O.uniform(key, 1000))
    assert np.array_equal(jaxrand.uniform(key, 1000, 1.17549435e-38, 1.0).cpu().numpy(), O.uniform(key, 1000, 1.17549435e-38, 1.0))
    assert np.array_equal(jaxrand.randint(key, 1000, 0, 1_000_000).cpu().numpy(), O.randint(key, 1000, 0, 1_000_000))
    assert np.array_equal(jaxrand.randint(key, 1000, -5, 4).cpu().numpy(), O.randint(key, 1000, -5, 4))


@pytest.mark.parametrize("n", [1, 31, 33, 449, 1000])
def test_det_play_random_ragged_sizes_caps_and_resume(n):
    """the persistent kernel on sizes that leave warps / CTAs partly empty, with a step cap (games stopped mid-way keep
    their exact intermediate state), resumed from that state (some games already done at load), with a game offset"""
    dm = _dm()
    from exploring_muzero_on_dog_b200 import jaxrand
    key = jaxrand.split_host(jaxrand.PRNGKey(7))[0]
    seeds = np.arange(n, dtype=np.int32) * 31 + 5
    cfg = O.MadnCfg(4, 0xF, 10, mask_of(TRAIN_RULES))
    env = dm.env_reset(0, seed=seeds, **TRAIN_RULES)
    s = O.madn_reset(cfg, seeds, 0)
    for cap in (0, 37, 320, 2000):  # resumed again and again: 320 leaves a mix of finished and running games
        _, glen = dm.play_random(env, key, max_steps=cap, game_offset=11)
        olen, _, _ = O.madn_det_play_random(s, key, cap, game_offset=11)
        assert_state_equal(s, env.numpy())
        assert np.array_equal(olen, glen.cpu().numpy())
    assert env.numpy()["done"].all()


def test_det_play_random_non_canonical_games_take_the_generic_rules():
    """a CTA that loads a game outside the canonical form (here: two players on one cell, an out-of-range player id)
    must give the generic kernel's / the oracle's answer for every game it owns"""
    dm = _dm()
    from exploring_muzero_on_dog_b200 import jaxrand
    n = 600
    key = jaxrand.split_host(jaxrand.PRNGKey(9))[1]
    cfg = O.MadnCfg(4, 0xF, 10, mask_of(TRAIN_RULES))
    s = O.madn_reset(cfg, np.arange(n, dtype=np.int32), 0)
    O.madn_det_play_random(s, key, 60)           # some plies in, so that pins are spread out
    s.pins[5, 1, 0] = s.pins[5, 0, 0]            # player 1 stacked onto player 0's first pin
    s.board[5] = O.madn_set_pins_on_board(cfg, s.pins[5:6])[0]
    s.board[450, 17] = 2                         # board that contradicts the pins
    env = _upload_det(dm, s, TRAIN_RULES)
    _, glen = dm.play_random(env, key, max_steps=2000)
    olen, _, _ = O.madn_det_play_random(s, key, 2000)
    assert_state_equal(s, env.numpy())
    assert np.array_equal(olen, glen.cpu().numpy())


def test_det_play_random_states_outside_the_track_rules_take_the_generic_rules():
    """the training-rule program keeps a game as track positions + nibble counts (csrc/madn_track.cuh); a CTA that loads a state
    that representation cannot hold — a card count above 7, a negative count, a complete team whose done flag is not set — must
    give the oracle's answer for every game it owns, and CTAs next to it are unaffected"""
    dm = _dm()
    from exploring_muzero_on_dog_b200 import jaxrand
    n = 1500                                         # 11 games per CTA: games 5, 700, 1200 sit in different CTAs
    key = jaxrand.split_host(jaxrand.PRNGKey(13))[1]
    cfg = O.MadnCfg(4, 0xF, 10, mask_of(TRAIN_RULES))
    s = O.madn_reset(cfg, np.arange(n, dtype=np.int32) + 50, 0)
    O.madn_det_play_random(s, key, 40)
    s.action_set[5, 2, 3] = 9
    s.action_set[700, 0, 0] = -3
    s.pins[1200, 0] = [40, 41, 42, 43]
    s.pins[1200, 2] = [48, 49, 50, 51]
    s.board[1200] = O.madn_set_pins_on_board(cfg, s.pins[1200:1201])[0]
    s.done[1200] = 0
    env = _upload_det(dm, s, TRAIN_RULES)
    _, glen = dm.play_random(env, key, max_steps=2000)
    olen, _, _ = O.madn_det_play_random(s, key, 2000)
    assert_state_equal(s, env.numpy())
    assert np.array_equal(olen, glen.cpu().numpy())


@pytest.mark.parametrize("seed", range(3))
def test_det_play_random_other_rule_sets_and_draw_ahead_sizes(seed):
    """the persistent kernel under random rule dicts (the run-time-rules program: bitboard state) and at batch sizes whose CTAs
    start below the draw-ahead threshold (<= 64 live games per CTA from the first iteration on), in chunks of 1 / 7 / 48 / 49
    iterations per launch (round boundaries, a launch shorter than a draw chunk)"""
    dm = _dm()
    from exploring_muzero_on_dog_b200 import jaxrand
    rng = np.random.default_rng(40 + seed)
    for rules in all_rule_sets(rng, 2):
        n = int(rng.integers(200, 9000))             # 2 .. 61 games per CTA
        key = jaxrand.split_host(jaxrand.PRNGKey(seed))[1]
        cfg = O.MadnCfg(4, 0xF, 10, mask_of(rules))
        seeds = rng.integers(0, 1_000_000, n).astype(np.int32)
        env = dm.env_reset(0, seed=seeds, **rules)
        s = O.madn_reset(cfg, seeds, 0)
        k = key
        for chunk in (1, 7, 48, 49, 3, 2000):
            k_next = dm.random_steps(env, k, chunk) if chunk < 2000 else None
            if chunk == 2000:
                dm.play_random(env, k, max_steps=chunk)
            _, _, okey = O.madn_det_play_random(s, k, chunk)
            assert_state_equal(s, env.numpy())
            if k_next is not None:
                assert np.asarray(k_next).tolist() == okey.tolist()
                k = np.asarray(k_next, dtype=np.uint32)


@pytest.mark.parametrize("n", [65536, 5000])
def test_det_play_random_repeated_runs_are_bit_identical(n):
    """the persistent kernel hands games between warps through shared memory (compaction rounds, the draw-ahead tables behind
    named barriers, the key ring): a missing barrier or fence would show as run-to-run differences.  30 runs of the same
    batch must give the same bytes in every leaf and the same game lengths (the first run is checked against the oracle by
    the tests above / by bench.py at full size)"""
    dm = _dm()
    from exploring_muzero_on_dog_b200 import jaxrand
    key = jaxrand.split_host(jaxrand.PRNGKey(21))[1]
    seeds = jaxrand.randint(key, n, 0, 1_000_000)
    env0 = dm.env_reset(0, seed=seeds, **TRAIN_RULES)
    ref = None
    for rep in range(30):
        env = env0.clone()
        _, glen = dm.play_random(env, key, max_steps=2000)
        got = {k: v.clone() for k, v in env.leaves().items()} if hasattr(env, "leaves") else {k: torch.as_tensor(v) for k, v in env.numpy().items()}
        got["game_len"] = glen.clone()
        if ref is None:
            ref = got
            continue
        for k, v in ref.items():
            assert torch.equal(torch.as_tensor(v).cpu(), torch.as_tensor(got[k]).cpu()), (rep, k)


def test_entry_points_follow_the_device_of_their_tensors():
    """an env created with device='cuda:1' while cuda:0 is current must run on GPU 1, on GPU 1's current stream
    (needs two GPUs; the one-GPU test box skips it, `gpurun --gpus 2` runs it)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from exploring_muzero_on_dog_b200 import jaxrand
    from exploring_muzero_on_dog_b200.MADN import deterministic_madn as dm
    torch.cuda.set_device(0)
    key = jaxrand.split_host(jaxrand.PRNGKey(0))[1]
    seeds = O.randint(key, 512, 0, 1_000_000)
    env = dm.env_reset(0, seed=torch.as_tensor(seeds, device="cuda:1"), device="cuda:1", **TRAIN_RULES)
    assert env.raw("board").device.index == 1 and torch.cuda.current_device() == 0
    mask = dm.valid_action(env)
    _, glen = dm.play_random(env, key, max_steps=2000)
    torch.cuda.synchronize(1)
    s = O.madn_reset(O.MadnCfg(4, 0xF, 10, mask_of(TRAIN_RULES)), seeds, 0)
    assert np.array_equal(O.madn_det_valid_action(s), mask.cpu().numpy())
    olen, _, _ = O.madn_det_play_random(s, key, 2000)
    assert_state_equal(s, env.numpy())
    assert np.array_equal(olen, glen.cpu().numpy()) and glen.device.index == 1
