"""Goldens for the true-env mctx callbacks of classic (dice) MADN — MADN/classic_madn.py:541-714: winning_action,
policy_function, rollout, root_fn, recurrent_fn (decision -> afterstate), recurrent_chance_fn (afterstate + die -> state) —
produced by the REFERENCE'S OWN function bodies on the NumPy-backed jaxshim.

    python tests/golden/gen_madn_cls_trueenv_goldens.py [workers]   # build container only; writes madn_cls_reference_trueenv.npz

ONE thing is patched, and it is not in the functions: `winning_action` (:551-565) builds its scratch copy with
`classic_MADN(board=..., ..., rules=...)` and leaves out the dataclass's `key` field, so the reference as it stands raises
TypeError before any of these callbacks can return.  `env_step` never reads that field; here the module's
name `classic_MADN` is rebound to a constructor that fills `key=None` when it is missing, and everything else runs unmodified.
What is recorded is therefore "what the reference's code computes once that constructor call goes through".

The reference's rollout value is a float32[4] of four equal entries (its `winner == -1` test compares a bool array): all four
are stored."""
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "jaxshim"))
sys.path.insert(0, "/root/reference")

import jax  # noqa: E402  (the shim)
import jax.numpy as jnp  # noqa: E402
from MADN import classic_madn as cm  # noqa: E402

_strict = cm.classic_MADN


def _lenient(**kw):
    kw.setdefault("key", None)
    return _strict(**kw)


cm.classic_MADN = _lenient

TRAIN = dict(enable_teams=True, enable_initial_free_pin=True, enable_circular_board=False, enable_friendly_fire=False,
             enable_start_blocking=False, enable_jump_in_goal_area=True, enable_start_on_1=True,
             enable_bonus_turn_on_6=True, must_traverse_start=False, enable_dice_rethrow=True)
OTHER = dict(enable_teams=False, enable_initial_free_pin=True, enable_circular_board=True, enable_friendly_fire=True,
             enable_start_blocking=True, enable_jump_in_goal_area=False, enable_start_on_1=False,
             enable_bonus_turn_on_6=True, must_traverse_start=True, enable_dice_rethrow=False)
RULES = [TRAIN, OTHER]
LEAVES = ("board", "current_player", "pins", "reward", "done", "die", "key")


def leaves(env):
    return {k: np.asarray(getattr(env, k)).copy() for k in LEAVES}


def near_win(env, rng, teams):
    """constructed position (env.replace(pins=..., board=set_pins_on_board(...))): the mover's team needs one more pin"""
    cur = int(env.current_player)
    pins = np.full((4, 4), -1, np.int8)
    used = set()
    for p in range(4):
        goal0, start = 40 + 4 * p, 10 * p
        if p == cur or (teams and p == (cur + 2) % 4):
            for k in range(4):
                pins[p, k] = goal0 + k
            if p == cur:
                free = int(rng.integers(4))
                pos = ((start - 1) % 40 - int(rng.integers(0, 3))) % 40
                pins[p, free] = pos
                used.add(pos)
        else:
            for k in range(4):
                if rng.random() < 0.6:
                    pos = int(rng.integers(40))
                    while pos in used:
                        pos = int(rng.integers(40))
                    used.add(pos)
                    pins[p, k] = pos
    jp = jnp.asarray(pins)
    return env.replace(pins=jp, board=cm.set_pins_on_board(env.board, jp))


def job(args):
    ri, seed, plies = args
    rules = RULES[ri]
    rng = np.random.default_rng(seed)
    env = cm.env_reset(0, num_players=4, distance=10, starting_player=int(rng.integers(4)), seed=int(seed), **rules)
    if plies < 0:
        env = near_win(env, rng, rules["enable_teams"])
        plies = -plies - 1
    for t in range(plies):  # random legal play through the reference's own functions
        if env.done:
            break
        env = cm.throw_die(env)
        m = np.asarray(cm.valid_action(env)).astype(bool).reshape(-1)
        if m.any():
            env, _, _ = cm.env_step(env, jnp.int8(int(rng.choice(np.flatnonzero(m)))))
        else:
            env, _, _ = cm.no_step(env)
    env = cm.throw_die(env)
    out = {"rules": ri, "state": leaves(env), "policy": np.asarray(cm.policy_function(env)).astype(np.float32)}
    keys = [np.asarray(jax.random.split(jax.random.PRNGKey(int(rng.integers(1 << 30))))[0]) for _ in range(3)]
    out["keys"] = np.stack(keys).astype(np.uint32)
    m = np.asarray(cm.valid_action(env)).astype(bool).reshape(-1)
    a = int(rng.integers(4)) if (rng.random() < 0.2 or not m.any()) else int(rng.choice(np.flatnonzero(m)))
    c = int(rng.integers(6))
    out["action"], out["outcome"] = a, c
    root = cm.root_fn(env, jnp.asarray(keys[0]))
    out["root_value"] = np.asarray(root.value).astype(np.float32).reshape(-1)
    dec, after = cm.recurrent_fn(None, jnp.asarray(keys[1]), jnp.int8(a), env)
    out["chance_logits"] = np.asarray(dec.chance_logits).astype(np.float32)
    out["after_value"] = np.asarray(dec.afterstate_value).astype(np.float32).reshape(-1)
    out["after"] = leaves(after)
    ch, nxt = cm.recurrent_chance_fn(None, jnp.asarray(keys[2]), jnp.int32(c), after)
    out["ch_logits"] = np.asarray(ch.action_logits).astype(np.float32).reshape(-1)
    out["ch_value"] = np.asarray(ch.value).astype(np.float32).reshape(-1)
    out["ch_reward"] = np.float32(np.asarray(ch.reward))
    out["ch_discount"] = np.float32(np.asarray(ch.discount))
    out["next"] = leaves(nxt)
    return out


def main():
    workers = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    jobs, seed = [], 5000
    for ri in range(len(RULES)):
        for plies, count in ((2, 4), (60, 6), (200, 8), (320, 10), (400, 12), (460, 12), (520, 10), (-1, 8), (-2, 4), (-4, 4)):
            for _ in range(count):
                jobs.append((ri, seed, plies))
                seed += 1
    with mp.Pool(workers) as pool:
        res = pool.map(job, jobs, chunksize=1)
    out = {"n": np.int64(len(res)), "rules": np.array([r["rules"] for r in res], np.int32)}
    for pre, name in (("s_", "state"), ("a_", "after"), ("n_", "next")):
        for k in LEAVES:
            out[pre + k] = np.stack([r[name][k] for r in res])
    for k in ("policy", "keys", "root_value", "chance_logits", "after_value", "ch_logits", "ch_value"):
        out[k] = np.stack([r[k] for r in res])
    out["action"] = np.array([r["action"] for r in res], np.int32)
    out["outcome"] = np.array([r["outcome"] for r in res], np.int32)
    out["ch_reward"] = np.array([r["ch_reward"] for r in res], np.float32)
    out["ch_discount"] = np.array([r["ch_discount"] for r in res], np.float32)
    rule_keys = sorted(TRAIN)
    out["rule_keys"] = np.array(rule_keys)
    out["rule_values"] = np.array([[int(r[k]) for k in rule_keys] for r in RULES], np.int8)
    np.savez_compressed(os.environ.get("DOGSTEP_GOLDEN_OUT", os.path.join(HERE, "madn_cls_reference_trueenv.npz")), **out)
    print("states", len(res), "done states", int((out["s_done"] != 0).sum()), "winning moves", int((out["policy"] >= 300).any(1).sum()),
          "root values won", int((out["root_value"][:, 0] > 0).sum()), "after done", int((out["a_done"] != 0).sum()))


if __name__ == "__main__":
    main()
