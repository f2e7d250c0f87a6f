"""DOG golden trajectories produced by the REFERENCE ITSELF (/root/reference/DOG/dog.py executed unmodified on
the NumPy-backed jaxshim; build container only).

    python tests/golden/gen_dog_goldens.py      # writes tests/golden/dog_reference_trajectories.npz

Games are driven through env_reset / valid_actions / env_step / no_step (swap phase, play phase, re-deals, deck
resets, hot-7 captures, joker copies).  Actions come from a NumPy RNG: mostly a random legal action, sometimes an
arbitrary index (illegal card, illegal move, play index during the swap phase).  Every leaf of every state, the
full 806-wide legal mask (bit-packed), reward and done are stored.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "jaxshim"))
sys.path.insert(0, "/root/reference")

import jax.numpy as jnp  # noqa: E402  (the shim)
from DOG import dog  # noqa: E402

BENCH = dict(enable_teams=True, enable_initial_free_pin=False, enable_circular_board=True, enable_friendly_fire=True,
             enable_start_blocking=True, enable_jump_in_goal_area=False, must_traverse_start=True)  # MuZero_DOG/game_agent.py:12-23
KEYS = list(BENCH)
LEAVES = ("board", "current_player", "pins", "reward", "done", "deck", "hands", "swap_choices", "round_starter", "phase",
          "key", "hand_size")


def leaves(env):
    return {k: np.asarray(getattr(env, k)) for k in LEAVES}


def play(num_players, rules, seed, sp, plies, rng):
    env = dog.env_reset(0, num_players=num_players, distance=10, starting_player=sp, seed=seed, **rules)
    rec = {k: [v] for k, v in leaves(env).items()}
    masks, actions, rewards, dones, kinds = [], [], [], [], []
    for t in range(plies):
        m = np.asarray(dog.valid_actions(env)).astype(bool)
        arbitrary = rng.random() < 0.1
        if m.any() or arbitrary:
            a = int(rng.integers(m.size)) if (arbitrary or not m.any()) else int(rng.choice(np.flatnonzero(m)))
            env, r, d = dog.env_step(env, jnp.array(a, dtype=jnp.int32))
            kinds.append(1)
        else:
            a = -1
            env, r, d = dog.no_step(env)
            kinds.append(0)
        masks.append(np.packbits(m))
        actions.append(a)
        rewards.append(int(np.asarray(r)))
        dones.append(bool(np.asarray(d)))
        for k, v in leaves(env).items():
            rec[k].append(v)
    out = {f"state_{k}": np.stack(v) for k, v in rec.items()}
    out.update(mask=np.stack(masks), action=np.array(actions, np.int32), reward=np.array(rewards, np.int8),
               done=np.array(dones, bool), kind=np.array(kinds, np.int8))
    return out


def main():
    rng = np.random.default_rng(7)
    sets = [dict(BENCH)] + [{k: bool(rng.integers(2)) for k in KEYS} for _ in range(4)]
    games, meta = [], []
    for ri, rules in enumerate(sets):
        plan = ((4, 330, 2), (2, 120, 1)) if ri == 0 else ((4, 200, 1), (2, 80, 1), (3, 80, 1))
        for num_players, plies, reps in plan:
            for rep in range(reps):
                seed = int(rng.integers(0, 1_000_000))
                sp = int(rng.integers(-1, num_players))
                g = play(num_players, rules, seed, sp, plies, rng)
                games.append(g)
                meta.append(dict(rules=rules, num_players=num_players, seed=seed, starting_player=sp, plies=plies))
                print(ri, num_players, rep, "done at", int(np.argmax(g["done"])) if g["done"].any() else None,
                      "invalid", int((g["reward"] == -1).sum()), "no_step", int((g["kind"] == 0).sum()), flush=True)
    flat = {}
    for i, g in enumerate(games):
        for k, v in g.items():
            flat[f"g{i}_{k}"] = v
    flat["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "dog_reference_trajectories.npz"), **flat)


if __name__ == "__main__":
    main()
