"""Golden trajectories produced by the REFERENCE ITSELF (its unmodified Python sources under /root/reference,
executed on the NumPy-backed jaxshim because no JAX wheel exists in this image).

    python tests/golden/gen_madn_goldens.py        # build container only; writes tests/golden/madn_*.npz

For every rule set: a handful of 4-player (and 2-/3-player) games are played for a few hundred plies through
the reference's env_reset / valid_action / env_step / no_step / encode_board (and throw_die /
dice_probabilities for the dice variant).  Actions come from a NumPy RNG: mostly a random legal action,
sometimes an arbitrary (often illegal) one, so reward -1 / turn passing / refill quirks are covered.  Every
leaf of every intermediate state is stored; tests replay the same actions through the C oracle (CPU) and
through the CUDA path (GPU) and demand bit-equality.
"""
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
from MADN import deterministic_madn as dm  # noqa: E402

TRAIN = dict(enable_teams=True, enable_initial_free_pin=True, enable_circular_board=False, enable_friendly_fire=False,
             enable_start_blocking=False, enable_jump_in_goal_area=True, enable_start_on_1=True,
             enable_bonus_turn_on_6=True, must_traverse_start=False)
KEYS = list(TRAIN)


def rule_sets(rng, k, dice):
    out = [dict(TRAIN)]
    for _ in range(k):
        r = {key: bool(rng.integers(2)) for key in KEYS}
        out.append(r)
    if dice:
        for i, r in enumerate(out):
            r["enable_dice_rethrow"] = bool(i % 2 == 0)
    return out


def leaves(env, det):
    d = dict(board=np.asarray(env.board), current_player=np.asarray(env.current_player), pins=np.asarray(env.pins),
             reward=np.asarray(env.reward), done=np.asarray(env.done), key=np.asarray(env.key))
    if det:
        d["action_set"] = np.asarray(env.action_set)
    else:
        d["die"] = np.asarray(env.die)
    return d


def play(mod, det, num_players, rules, seed, starting_player, plies, rng):
    env = mod.env_reset(0, num_players=num_players, distance=10, starting_player=starting_player, seed=seed, **rules)
    rec = {k: [v] for k, v in leaves(env, det).items()}
    masks, actions, rewards, dones, kinds, obs, probs = [], [], [], [], [], [], []
    for t in range(plies):
        if not det:
            probs.append(np.asarray(cm.dice_probabilities(env)))
            env = cm.throw_die(env)
        m = np.asarray(mod.valid_action(env)).astype(bool)
        obs.append(np.asarray(mod.encode_board(env)).astype(np.int8))
        flat = m.reshape(-1)
        arbitrary = rng.random() < 0.12
        if flat.any() or arbitrary:
            a = int(rng.integers(flat.size)) if (arbitrary or not flat.any()) else int(rng.choice(np.flatnonzero(flat)))
            if det:
                act = np.array([a // 6, a % 6 + 1], np.int8)
                env, r, d = dm.env_step(env, jnp.array(act))
            else:
                act = np.array([a, 0], np.int8)
                env, r, d = cm.env_step(env, jnp.array(a, dtype=jnp.int8))
            kinds.append(1)
        else:
            act = np.array([-1, -1], np.int8)
            env, r, d = mod.no_step(env)
            kinds.append(0)
        masks.append(m)
        actions.append(act)
        rewards.append(np.asarray(r))
        dones.append(np.asarray(d))
        for k, v in leaves(env, det).items():
            rec[k].append(v)
    out = {f"state_{k}": np.stack(v) for k, v in rec.items()}
    out.update(mask=np.stack(masks), action=np.stack(actions), reward=np.stack(rewards).astype(np.int8),
               done=np.stack(dones).astype(bool), kind=np.array(kinds, np.int8), obs=np.stack(obs))
    if not det:
        out["dice_probs"] = np.stack(probs).astype(np.float32)
    return out


def main():
    rng = np.random.default_rng(20260101)
    for det, mod, name in ((True, dm, "det"), (False, cm, "cls")):
        games, meta = [], []
        for ri, rules in enumerate(rule_sets(rng, 5, dice=not det)):
            plan = ((4, 260, 2), (2, 120, 1), (3, 120, 1)) + (((4, 800, 2),) if ri == 0 else ())  # long games reach termination
            for num_players, plies, reps in plan:
                for rep in range(reps):
                    seed = int(rng.integers(0, 1_000_000))
                    sp = int(rng.integers(-1, num_players))
                    g = play(mod, det, num_players, rules, seed, sp, plies, rng)
                    games.append(g)
                    meta.append(dict(rules=rules, num_players=num_players, seed=seed, starting_player=sp, plies=plies))
                    print(name, ri, num_players, rep, "done at ply", int(np.argmax(g["done"])) if g["done"].any() else None)
        flat = {}
        for i, g in enumerate(games):
            for k, v in g.items():
                flat[f"g{i}_{k}"] = v
        import json
        flat["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(HERE, f"madn_{name}_reference_trajectories.npz"), **flat)

    # the reference's random lockstep driver (MuZero_det_MADN/evaluate_agent.py:733-930, do_random) on 6 games
    N, T = 6, 140
    rng_key = jax.random.split(jax.random.PRNGKey(0))[1]
    seeds = np.asarray(jax.random.randint(rng_key, (N,), 0, 1000000))
    envs = [dm.env_reset(0, num_players=4, distance=10, starting_player=0, seed=int(s), **TRAIN) for s in seeds]
    acts = np.full((T, N), -1, np.int32)
    for t in range(T):
        rng_key, *step_keys = jax.random.split(rng_key, N + 1)
        for j in range(N):
            env = envs[j]
            if bool(env.done):
                continue
            vm = dm.valid_action(env).flatten()
            if bool(jnp.any(vm)):
                logits = jnp.where(vm, 0.0, -1e9)
                a = jax.random.categorical(step_keys[j], logits)
                acts[t, j] = int(a)
                envs[j], _, _ = dm.env_step(env, dm.map_action(a))
            else:
                envs[j], _, _ = dm.no_step(env)
    np.savez_compressed(os.path.join(HERE, "madn_det_reference_random_driver.npz"), seeds=seeds, actions=acts,
                        final_pins=np.stack([np.asarray(e.pins) for e in envs]).astype(np.int8),
                        final_board=np.stack([np.asarray(e.board) for e in envs]).astype(np.int8),
                        final_action_set=np.stack([np.asarray(e.action_set) for e in envs]).astype(np.int8),
                        final_current_player=np.array([int(e.current_player) for e in envs], np.int8),
                        final_key=np.asarray(rng_key).astype(np.uint32))
    print("random driver done")


if __name__ == "__main__":
    main()
