"""The DOG oracle against (a) the reference's own 112 parametrised cases (DOG/test.py) and (b) trajectories the
reference itself produced on the jaxshim (tests/golden/gen_dog_goldens.py)."""
import json
import os

import numpy as np
import pytest

import oracle as O
from helpers import DOG_CASE_SETS, DOG_CODE_WINS, DOG_RULES, dog_case_rules, mask_of

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [(name, kind, fn, i) for name, kind, fn in DOG_CASE_SETS for i in range({"test_normal_move": 52, "test_neg_move": 17,
                                                                                "test_swap_move": 14, "test_7_move": 29}[name])]


@pytest.mark.parametrize("name,kind,fn,i", CASES, ids=[f"{c[0]}-{c[3]}" for c in CASES])
def test_reference_cases(ref_cases, name, kind, fn, i):
    c = ref_cases[f"DOG/test.py::{name}"][i]
    pins = np.array(c["pins"], np.int32)
    cfg = O.DogCfg(len(pins), 0xF, 10, mask_of(dog_case_rules(c["rules"])))
    s = O.dog_reset(cfg, [42], 0)
    s.pins[0] = pins
    s.board = O.madn_set_pins_on_board(cfg, pins[None].astype(np.int8))
    s.current_player[0] = c["player"]
    board, p, r, d = O.dog_substep(s, [kind], [fn(c)])
    assert p[0].tolist() == DOG_CODE_WINS.get((name, i), c["expected_valid"])


def test_speedtest_notebook_kats():
    """SURVEY Appendix E.1-E.2 (DOG/speedtest.ipynb cells 2, 15, 16): joker-swap of an own pin is invalid -> env
    unchanged except the turn passes with reward -1; map_action_to_move(739) is the all-zero move."""
    rules = dict(enable_teams=False, enable_initial_free_pin=False, enable_circular_board=True, enable_jump_in_goal_area=True,
                 enable_start_blocking=False, enable_friendly_fire=False, must_traverse_start=True)
    cfg = O.DogCfg(2, 0xF, 10, mask_of(rules))
    s = O.dog_reset(cfg, [42], 0)
    pins = np.array([[38, 25, 10, 0], [6, 45, 44, -1]], np.int32)
    s.pins[0] = pins
    s.board = O.madn_set_pins_on_board(cfg, pins[None].astype(np.int8))
    s.hands[...] = 1
    s.current_player[0] = 0
    before = s.copy()
    r, d = O.dog_step(s, [0])
    assert r[0] == -1 and not d[0] and s.current_player[0] == 1
    assert np.array_equal(s.pins, before.pins) and np.array_equal(s.board, before.board) and np.array_equal(s.hands, before.hands)


def test_reset_deal_invariants():
    cfg = O.DogCfg(4, 0xF, 10, mask_of(DOG_RULES))
    s = O.dog_reset(cfg, np.arange(64) * 31, 0)
    assert (s.hands.sum((1, 2)) == 24).all() and (s.deck.sum(1) == 110 - 24).all()
    assert (s.hand_size == 5).all() and (s.phase == 1).all() and (s.round_starter == 0).all()
    assert (s.hands >= 0).all() and (s.deck >= 0).all() and (s.deck[:, 0] <= 6).all()
    m = O.dog_valid_actions(s)
    assert m.shape == (64, 806) and not m[:, :792].any() and (m[:, 792:] == (s.hands[:, 0] > 0)).all()


def _replay(step_fn):
    z = np.load(os.path.join(G, "dog_reference_trajectories.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    plies = 0
    for gi, m in enumerate(meta):
        plies += step_fn(z, gi, m)
    return plies


def _oracle_game(z, gi, m):
    cfg = O.DogCfg(m["num_players"], 0xF, 10, mask_of(m["rules"]))
    s = O.dog_reset(cfg, [m["seed"]], m["starting_player"])

    def check(t, where):
        for k, v in s.fields().items():
            exp = z[f"g{gi}_state_{k}"][t]
            assert np.array_equal(np.asarray(v[0]).astype(np.int64), np.asarray(exp).astype(np.int64)), \
                f"{where}: game {gi} ply {t} leaf {k}: got {np.asarray(v[0]).tolist()} want {exp.tolist()}"

    check(0, "reset")
    for t in range(m["plies"]):
        mask = O.dog_valid_actions(s)[0]
        exp_mask = np.unpackbits(z[f"g{gi}_mask"][t])[:mask.size].astype(bool)
        if not np.array_equal(mask, exp_mask):
            diff = np.flatnonzero(mask != exp_mask)
            raise AssertionError(f"mask: game {gi} ply {t} differs at actions {diff[:10].tolist()} (oracle {mask[diff[:10]].tolist()})")
        if z[f"g{gi}_kind"][t] == 1:
            r, d = O.dog_step(s, [z[f"g{gi}_action"][t]])
        else:
            r, d = O.dog_no_step(s)
        assert int(r[0]) == int(z[f"g{gi}_reward"][t]) and bool(d[0]) == bool(z[f"g{gi}_done"][t]), f"reward/done game {gi} ply {t}"
        check(t + 1, "step")
    return m["plies"]


def test_oracle_reproduces_reference_trajectories():
    assert _replay(_oracle_game) > 1500


def test_random_play_terminates():
    cfg = O.DogCfg(4, 0xF, 10, mask_of(DOG_RULES))
    key = O.split(O.prng_key(0))[1]
    s = O.dog_reset(cfg, O.randint(key, 64, 0, 1_000_000), 0)
    glen, total, _ = O.dog_play_random(s, key, 2000, nthreads=8)
    assert s.done.all() and 100 < glen.mean() < 1500
    s2 = O.dog_reset(cfg, O.randint(key, 64, 0, 1_000_000), 0)
    glen2, _, _ = O.dog_play_random(s2, key, 2000, float_gumbel=True, nthreads=8)
    assert np.array_equal(glen, glen2) and np.array_equal(s.pins, s2.pins)
