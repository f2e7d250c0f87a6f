"""The CPU oracle replays the reference's own 64+64 parametrised MADN cases
(/root/reference/MADN/test.py:7-945, transcribed to tests/golden/reference_cases.json)."""
import numpy as np
import pytest

import oracle as O
from helpers import madn_case_rules, mask_of, TRAIN_RULES


def _run(case, det):
    pins = np.array(case["pins"], np.int8)
    cfg = O.MadnCfg(len(pins), 0xF, 10, mask_of(madn_case_rules(case["rules"])))
    s = O.madn_reset(cfg, [42], 0, det=det)
    s.pins[0] = pins
    s.board = O.madn_set_pins_on_board(cfg, s.pins)
    s.current_player[0] = case["player"]
    if det:
        valid = O.madn_det_valid_action(s)[0, case["pin"], case["move"] - 1]
        r, d = O.madn_det_step(s, [[case["pin"], case["move"]]])
    else:
        s.die[0] = case["move"]
        valid = O.madn_cls_valid_action(s)[0, case["pin"]]
        r, d = O.madn_cls_step(s, [case["pin"]])
    assert valid or r[0] == -1
    assert s.pins[0].tolist() == case["expected_valid"]


@pytest.mark.parametrize("i", range(64))
def test_reference_cases_deterministic(ref_cases, i):
    _run(ref_cases["MADN/test.py::test_normal_move_deterministic_MADN"][i], True)


@pytest.mark.parametrize("i", range(64))
def test_reference_cases_classic(ref_cases, i):
    _run(ref_cases["MADN/test.py::test_normal_move_classic_MADN"][i], False)


def test_reset_key_and_layout():
    cfg = O.MadnCfg(4, 0xF, 10, mask_of(TRAIN_RULES))
    s = O.madn_reset(cfg, [42, 7], 0)
    assert s.key[0].tolist() == [1832780943, 270669613]  # split(PRNGKey(42))[0]
    assert s.pins[0].tolist() == [[0, -1, -1, -1], [10, -1, -1, -1], [20, -1, -1, -1], [30, -1, -1, -1]]
    assert (s.board[0][[0, 10, 20, 30]] == [0, 1, 2, 3]).all() and (s.board[0] >= 0).sum() == 4
    start, target, goal = cfg.geometry()
    assert start.tolist() == [0, 10, 20, 30] and target.tolist() == [39, 9, 19, 29]
    assert goal.tolist() == [[40, 41, 42, 43], [44, 45, 46, 47], [48, 49, 50, 51], [52, 53, 54, 55]]


CLASSIFICATION_FIXTURES = [  # MuZero_det_MADN/classification_test.py:92-133 (name, pins, winning action index)
    ("pre_win", [[35, 41, 42, 43], [5, 15, 7, 12], [48, 49, 50, 51], [25, 28, 33, 30]], 4),
    ("pre_win_6", [[34, 41, 42, 43], [5, 15, 7, 12], [48, 49, 50, 51], [25, 28, 33, 30]], 5),
    ("pre_lose", [[-1, -1, -1, 2], [5, 44, 45, 46], [1, 3, 20, 21], [52, 53, 54, 55]], None),
    ("normal", [[10, 20, 30, -1], [15, 25, -1, -1], [5, 35, -1, -1], [8, 18, -1, -1]], None),
]


@pytest.mark.parametrize("name,pins,win_idx", CLASSIFICATION_FIXTURES)
def test_classification_fixture_winning_action(name, pins, win_idx):
    """with the training rules and P0 to move, the hand-built team positions of
    MuZero_det_MADN/classification_test.py have the winning action the script names (or none)."""
    cfg = O.MadnCfg(4, 0xF, 10, mask_of(TRAIN_RULES))
    wins = []
    for a in range(24):
        s = O.madn_reset(cfg, [0], 0)
        s.pins[0] = np.array(pins, np.int8)
        s.board = O.madn_set_pins_on_board(cfg, s.pins)
        r, d = O.madn_det_step(s, [[a // 6, a % 6 + 1]])
        if r[0] == 1:
            wins.append(a)
            assert d[0]
    assert wins == ([] if win_idx is None else [win_idx])


def test_random_play_terminates_with_plausible_length():
    """random-policy det-MADN games (teams, non-circular) average ~395 plies."""
    cfg = O.MadnCfg(4, 0xF, 10, mask_of(TRAIN_RULES))
    key = O.split(O.prng_key(0))[1]
    seeds = O.randint(key, 256, 0, 1_000_000)
    s = O.madn_reset(cfg, seeds, 0)
    glen, total, _ = O.madn_det_play_random(s, key, 2000, nthreads=4)
    assert s.done.all()
    assert 300 < glen.mean() < 500 and glen.max() < 1500
    # float-gumbel evaluation gives the very same games
    s2 = O.madn_reset(cfg, seeds, 0)
    glen2, _, _ = O.madn_det_play_random(s2, key, 2000, float_gumbel=True, nthreads=4)
    assert np.array_equal(glen, glen2) and np.array_equal(s.pins, s2.pins)


# Printed by the reference itself (MADN/jupyter_code/test_functions.ipynb cells 3-4, "Dice probabilities: ..."): a player whose
# pins are all at home or parked at the end of the goal lane is soft-locked and throws up to three times when
# enable_dice_rethrow is set — 1 - (5/6)^3 for the six, or (1 - (4/6)^3) / 2 each for one and six with enable_start_on_1.
# The notebook's 2-player board (20 ring cells, goal 20..23) is the current layout's 40-cell ring with goal 40..43.
NOTEBOOK_DICE = {(False, False): [0.16666667] * 6, (False, True): [0.16666667] * 6,
                 (True, False): [0.11574074] * 5 + [0.4212963], (True, True): [0.35185185] + [0.07407407] * 4 + [0.35185185]}
_NOTEBOOK_BASE = dict(enable_teams=False, enable_initial_free_pin=False, enable_circular_board=True, enable_start_blocking=False,
                      enable_jump_in_goal_area=True, enable_friendly_fire=False, enable_bonus_turn_on_6=True, must_traverse_start=False)


def notebook_dice_states():
    from exploring_muzero_on_dog_b200 import rules as R
    for (rethrow, on1), expect in NOTEBOOK_DICE.items():
        cfg = O.MadnCfg(2, 0b0101, 10, R.to_mask(dict(_NOTEBOOK_BASE, enable_dice_rethrow=rethrow, enable_start_on_1=on1)))
        for pins, exp in (([[-1, -1, 43, 42], [5, 6, 7, 8]], expect), ([[1, 2, 3, 4], [5, 6, 7, 8]], [0.16666667] * 6)):
            s = O.madn_reset(cfg, np.array([0], np.int32), 0, det=False)
            s.pins[0] = np.array(pins, np.int8)
            s.board[...] = O.madn_set_pins_on_board(cfg, s.pins)
            yield dict(enable_dice_rethrow=rethrow, enable_start_on_1=on1), s, np.array(exp, np.float32)


def test_dice_probabilities_match_the_reference_notebook():
    for _, s, exp in notebook_dice_states():
        assert np.array_equal(np.round(O.madn_cls_dice_probabilities(s)[0].astype(np.float64), 8), np.round(exp.astype(np.float64), 8))
