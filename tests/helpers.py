"""Shared helpers for the parity tests."""
import numpy as np

RULE_BITS = {
    "enable_teams": 1 << 0, "enable_initial_free_pin": 1 << 1, "enable_circular_board": 1 << 2,
    "enable_start_blocking": 1 << 3, "enable_jump_in_goal_area": 1 << 4, "enable_friendly_fire": 1 << 5,
    "enable_start_on_1": 1 << 6, "enable_bonus_turn_on_6": 1 << 7, "must_traverse_start": 1 << 8,
    "enable_dice_rethrow": 1 << 9,
}

# MuZero_det_MADN/game_agent.py:12-22 — the rule set of every training / benchmark config
TRAIN_RULES = dict(enable_teams=True, enable_initial_free_pin=True, enable_circular_board=False,
                   enable_friendly_fire=False, enable_start_blocking=False, enable_jump_in_goal_area=True,
                   enable_start_on_1=True, enable_bonus_turn_on_6=True, must_traverse_start=False)


def mask_of(rules):
    return sum(RULE_BITS[k] for k, v in rules.items() if v)


def madn_case_rules(case_rules):
    """How MADN/test.py:455-470,927-937 builds the env for a case: four rules are passed through,
    start_on_1 / must_traverse_start default to False in the test, the rest are env_reset defaults."""
    r = dict(enable_teams=False, enable_initial_free_pin=False, enable_bonus_turn_on_6=True,
             enable_circular_board=case_rules["enable_circular_board"],
             enable_jump_in_goal_area=case_rules["enable_jump_in_goal_area"],
             enable_start_blocking=case_rules["enable_start_blocking"],
             enable_friendly_fire=case_rules["enable_friendly_fire"],
             enable_start_on_1=case_rules.get("enable_start_on_1", False),
             must_traverse_start=case_rules.get("must_traverse_start", False))
    return r


def all_rule_sets(rng, k, dice=False):
    """k random rule dicts plus the training rule set"""
    keys = list(TRAIN_RULES) + (["enable_dice_rethrow"] if dice else [])
    out = [dict(TRAIN_RULES)]
    for _ in range(k):
        out.append({key: bool(rng.integers(2)) for key in keys})
    return out


def assert_state_equal(oracle_state, got, keys=None):
    for k, v in oracle_state.fields().items():
        if keys and k not in keys:
            continue
        a, b = np.asarray(v).astype(np.int64), np.asarray(got[k]).astype(np.int64)
        if not np.array_equal(a, b):
            bad = np.argwhere(a.reshape(a.shape[0], -1) != b.reshape(b.shape[0], -1))
            raise AssertionError(f"leaf {k} differs at game {bad[0][0]}: oracle {a[bad[0][0]].tolist()} cuda {b[bad[0][0]].tolist()}")


# MuZero_DOG/game_agent.py:12-23 — the DOG rule set of benchmark configs 4 and 5
DOG_RULES = dict(enable_teams=True, enable_initial_free_pin=False, enable_circular_board=True, enable_friendly_fire=True,
                 enable_start_blocking=True, enable_jump_in_goal_area=False, must_traverse_start=True)
DOG_RULE_KEYS = list(DOG_RULES)

# (reference test function, sub-step kind, argument builder, xfail index) — DOG/test.py:376-832
DOG_CASE_SETS = [
    ("test_normal_move", 0, lambda c: [c["pin"], c["move"], 0, 0]),
    ("test_neg_move", 1, lambda c: [c["pin"], c["move"], 0, 0]),
    ("test_swap_move", 2, lambda c: [c["pin"], c["pos"], 0, 0]),
    ("test_7_move", 3, lambda c: list(c["dist"])),
]
# DOG/test.py:367-373 ("Testfall 51") contradicts dog.py:515-519: a pin standing on its own start is exempt from
# start blocking, so the code moves it to 13.  The code is the spec; the reference's own code run on
# the jaxshim gives exactly this result.
DOG_CODE_WINS = {("test_normal_move", 51): [[13, 35, 3, 1], [6, 14, 44, 10]]}


def dog_case_rules(case_rules):
    """how DOG/test.py builds its envs: four rules passed through, must_traverse_start defaults to True"""
    return dict(enable_teams=False, enable_initial_free_pin=False, enable_circular_board=case_rules["enable_circular_board"],
                enable_jump_in_goal_area=case_rules["enable_jump_in_goal_area"],
                enable_start_blocking=case_rules["enable_start_blocking"], enable_friendly_fire=case_rules["enable_friendly_fire"],
                must_traverse_start=case_rules.get("must_traverse_start", True))


def dog_rule_sets(rng, k):
    return [dict(DOG_RULES)] + [{key: bool(rng.integers(2)) for key in DOG_RULE_KEYS} for _ in range(k)]
