"""Goldens for the true-env mctx callbacks of deterministic MADN, produced by the REFERENCE ITSELF
(/root/reference/MADN/deterministic_madn.py:480-590 — winning_action, policy_function, rollout, root_fn, recurrent_fn —
executed unmodified on the NumPy-backed jaxshim).

    python tests/golden/gen_madn_trueenv_goldens.py [workers]     # build container only; writes madn_det_reference_trueenv.npz

States: positions 3 .. 420 plies into games played with the reference's own functions (random legal policy), so that rollouts
from a handful of plies up to the 300-step cap occur; for each state the policy logits,
for a subset rollout values for given keys (the reference returns a float32[4] with four equal entries — all are stored),
root_fn and recurrent_fn outputs for a random (sometimes illegal) action together with every leaf of the successor state."""
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "jaxshim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import jax  # noqa: E402  (the shim)
import jax.numpy as jnp  # noqa: E402
from MADN import deterministic_madn as dm  # noqa: E402

TRAIN = dict(enable_teams=True, enable_initial_free_pin=True, enable_circular_board=False, enable_friendly_fire=False,
             enable_start_blocking=False, enable_jump_in_goal_area=True, enable_start_on_1=True,
             enable_bonus_turn_on_6=True, must_traverse_start=False)
OTHER = dict(enable_teams=False, enable_initial_free_pin=True, enable_circular_board=True, enable_friendly_fire=True,
             enable_start_blocking=True, enable_jump_in_goal_area=False, enable_start_on_1=False,
             enable_bonus_turn_on_6=True, must_traverse_start=True)
RULES = [TRAIN, OTHER]
LEAVES = ("board", "current_player", "pins", "reward", "done", "action_set")


def leaves(env):
    return {k: np.asarray(getattr(env, k)).copy() for k in LEAVES}


def near_win(env, rng, teams):
    """a constructed position (env.replace(pins=..., board=set_pins_on_board(...)), the reference's own fixture style) in which
    the player to move can finish the game: its partner / nobody else matters is complete, three own pins sit in the goal and
    the fourth a few cells before a free goal cell; the others are scattered"""
    cur = int(env.current_player)
    pins = np.full((4, 4), -1, np.int8)
    used = set()
    for p in range(4):
        goal0, start = 40 + 4 * p, 10 * p
        if p == cur or (teams and p == (cur + 2) % 4):
            free = int(rng.integers(4)) if p == cur else -1
            for k in range(4):
                pins[p, k] = goal0 + k
            if p == cur:
                target = (start - 1) % 40
                back = int(rng.integers(1, 5))                        # cells before the target cell
                pos = (target - back) % 40
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
    return env.replace(pins=jp, board=dm.set_pins_on_board(env.board, jp))


def job(args):
    ri, seed, plies, do_rollout = args
    rules = RULES[ri]
    rng = np.random.default_rng(seed)
    env = dm.env_reset(0, num_players=4, distance=10, starting_player=int(rng.integers(4)), seed=int(seed), **rules)
    if plies < 0:
        env = near_win(env, rng, rules["enable_teams"])
        plies = -plies - 1
    for t in range(plies):  # random legal play through the reference's own functions
        m = np.asarray(dm.valid_action(env)).astype(bool).reshape(-1)
        if env.done:
            break
        if m.any():
            a = int(rng.choice(np.flatnonzero(m)))
            env, _, _ = dm.env_step(env, jnp.array([a // 6, a % 6 + 1], dtype=jnp.int8))
        else:
            env, _, _ = dm.no_step(env)
    out = {"rules": ri, "state": leaves(env), "policy": np.asarray(dm.policy_function(env)).astype(np.float32)}
    key = np.asarray(jax.random.split(jax.random.PRNGKey(int(rng.integers(1 << 30))))[0])
    out["key"] = key
    m = np.asarray(dm.valid_action(env)).astype(bool).reshape(-1)
    a = int(rng.integers(24)) if (rng.random() < 0.2 or not m.any()) else int(rng.choice(np.flatnonzero(m)))
    out["action"] = a
    if do_rollout and not bool(env.done):
        v = np.asarray(dm.rollout(env, jnp.asarray(key))).astype(np.float32).reshape(-1)
        out["rollout"] = v
        ro, env2 = dm.recurrent_fn(None, jnp.asarray(key), jnp.int8(a), env)
        out["rec"] = dict(reward=np.float32(np.asarray(ro.reward)), discount=np.float32(np.asarray(ro.discount)),
                          prior=np.asarray(ro.prior_logits).astype(np.float32), value=np.asarray(ro.value).astype(np.float32).reshape(-1),
                          state=leaves(env2))
    return out


def main():
    workers = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    jobs = []
    seed = 1000
    for ri in range(len(RULES)):
        for plies, count, roll in ((3, 6, True), (40, 8, True), (150, 10, True), (260, 12, True), (330, 14, True), (380, 14, True), (420, 14, True), (480, 10, True)):
            for _ in range(count):
                jobs.append((ri, seed, plies, roll))
                seed += 1
    for ri in range(len(RULES)):  # constructed near-win positions, 0 .. 2 plies on (negative = constructed, then -plies - 1 plies)
        for plies in (-1, -1, -1, -1, -1, -1, -2, -2, -3, -3):
            jobs.append((ri, seed, plies, True))
            seed += 1
    reuse = os.path.join(HERE, "madn_det_reference_trueenv.npz") if os.environ.get("DOGSTEP_GOLDEN_REUSE") else None
    if reuse:  # development shortcut: keep the random-play entries of the existing file, run only the constructed jobs
        old = np.load(reuse)
        n_old = int((np.arange(int(old["n"])) < len([j for j in jobs if j[2] >= 0])).sum())
        jobs_run = [j for j in jobs if j[2] < 0]
    else:
        old, n_old, jobs_run = None, 0, jobs
    with mp.Pool(workers) as pool:
        res = pool.map(job, jobs_run, chunksize=1)
    out = {"n": np.int64(len(res)), "rules": np.array([r["rules"] for r in res], np.int32)}
    for k in LEAVES:
        out["s_" + k] = np.stack([r["state"][k] for r in res])
    out["policy"] = np.stack([r["policy"] for r in res])
    out["key"] = np.stack([r["key"] for r in res]).astype(np.uint32)
    out["action"] = np.array([r["action"] for r in res], np.int32)
    has = np.array(["rollout" in r for r in res])
    out["has_rollout"] = has
    z4, z24 = np.full(4, np.nan, np.float32), np.zeros(24, np.float32)
    out["rollout"] = np.stack([r.get("rollout", z4) for r in res])
    out["rec_reward"] = np.array([r["rec"]["reward"] if "rec" in r else np.nan for r in res], np.float32)
    out["rec_discount"] = np.array([r["rec"]["discount"] if "rec" in r else np.nan for r in res], np.float32)
    out["rec_prior"] = np.stack([r["rec"]["prior"] if "rec" in r else z24 for r in res])
    out["rec_value"] = np.stack([r["rec"]["value"] if "rec" in r else z4 for r in res])
    for k in LEAVES:
        out["n_" + k] = np.stack([r["rec"]["state"][k] if "rec" in r else r["state"][k] for r in res])
    if old is not None:
        for k in list(out):
            if k != "n":
                out[k] = np.concatenate([old[k][:n_old], out[k]])
        out["n"] = np.int64(n_old + len(res))
        has = out["has_rollout"]
    rule_keys = sorted(TRAIN)
    out["rule_keys"] = np.array(rule_keys)
    out["rule_values"] = np.array([[int(r[k]) for k in rule_keys] for r in RULES], np.int8)
    np.savez_compressed(os.environ.get("DOGSTEP_GOLDEN_OUT", os.path.join(HERE, "madn_det_reference_trueenv.npz")), **out)
    print("states", int(out["n"]), "with rollouts", int(has.sum()), "done states", int((out["s_done"] != 0).sum()),
          "states with a winning move", int((out["policy"] >= 300).any(1).sum()), "rollouts won", int((out["rollout"][has][:, 0] > 0).sum()))


if __name__ == "__main__":
    main()
