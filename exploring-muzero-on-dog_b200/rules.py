"""Rule dict <-> bitmask (include/dogstep_rules.h).  The reference keeps these as a static dict
on every env (MADN/deterministic_madn.py:109-119, MADN/classic_madn.py:119-130)."""

BITS = {
    "enable_teams": 1 << 0,
    "enable_initial_free_pin": 1 << 1,
    "enable_circular_board": 1 << 2,
    "enable_start_blocking": 1 << 3,
    "enable_jump_in_goal_area": 1 << 4,
    "enable_friendly_fire": 1 << 5,
    "enable_start_on_1": 1 << 6,
    "enable_bonus_turn_on_6": 1 << 7,
    "must_traverse_start": 1 << 8,
    "enable_dice_rethrow": 1 << 9,
    "disable_swapping": 1 << 10,
    "disable_hot_seven": 1 << 11,
    "disable_joker": 1 << 12,
}


def to_mask(rules):
    m = 0
    for k, v in rules.items():
        if k not in BITS:
            raise KeyError(f"unknown rule {k!r}")
        if bool(v):
            m |= BITS[k]
    return m


def from_mask(mask, keys):
    return {k: bool(mask & BITS[k]) for k in keys}
