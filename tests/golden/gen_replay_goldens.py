"""Replay goldens from the REFERENCE CLASSES THEMSELVES (pure NumPy; only their final jnp.array() needs the jaxshim).

    python tests/golden/gen_replay_goldens.py      # build container only -> tests/golden/replay_reference.npz

np.random is seeded before sample_batch(); the reference's draw sequence (vec_replay_buffer.py:79-93) is replayed with
the same seed to recover the (episode, t_start) plan, which is stored with the reference's outputs.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "jaxshim"))


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def make_traj(rng, n, T, obs_shape, A, stochastic):
    idx = rng.integers(0, T + 1, n).astype(np.int32)
    idx[0] = 0                      # a zero-length game must be skipped
    idx[1] = T                      # a full-length game
    idx[2] = 1
    tr = dict(obs=rng.integers(-2, 5, (n, T) + obs_shape).astype(np.float32), act=rng.integers(-1, A, (n, T)).astype(np.int32),
              rew=np.ones((n, T), np.int32), val=rng.uniform(-1, 1, (n, T)).astype(np.float32),
              pol=rng.dirichlet(np.ones(A), (n, T)).astype(np.float32), mask=(rng.random((n, T)) < 0.9).astype(np.float32),
              player=rng.integers(0, 4, (n, T)).astype(np.int32), discount=rng.integers(0, 3, (n, T)).astype(np.int32), idx=idx)
    teams = rng.random(n) < 0.7
    tr["team"] = np.where(teams[:, None], tr["player"] % 2, -1).astype(np.int32)
    for i in range(n):              # terminal reward class on the last ply: 2 = win, 0 = loss, 1 = cut off
        if idx[i] > 0:
            tr["rew"][i, idx[i] - 1] = rng.choice([0, 1, 2], p=[0.15, 0.25, 0.6])
    if stochastic:
        tr["dice"] = rng.integers(1, 7, (n, T)).astype(np.int32)
        tr["dice_dist"] = rng.dirichlet(np.ones(6), (n, T)).astype(np.float32)
    return tr


def replay_plan(seed, size, lengths, batch_size, unroll_steps):
    """the reference's own draw sequence (:73-97)"""
    np.random.seed(seed)
    n_term = int(batch_size * 0.25)
    n_norm = batch_size - n_term
    e1 = np.random.randint(0, size, size=n_norm)
    t1 = np.random.randint(0, (lengths[e1] - 1) + 1)
    e2 = np.random.randint(0, size, size=n_term)
    l2 = lengths[e2]
    mk = np.minimum(unroll_steps - 1, l2 - 1)
    tk = np.array([np.random.randint(0, int(m) + 1) for m in mk])
    t2 = np.maximum(l2 - 1 - tk, 0)
    return np.concatenate([e1, e2]).astype(np.int32), np.concatenate([t1, t2]).astype(np.int32)


def main():
    det = load("/root/reference/MuZero_det_MADN/vec_replay_buffer.py", "ref_replay_det")
    sto = load("/root/reference/MuZero_Classic_MADN/vec_replay_buffer_stochastic.py", "ref_replay_sto")
    out = {}
    for name, cls, stochastic, obs_shape, A in (("det", det.VectorizedReplayBuffer, False, (6, 8), 24),
                                                ("sto", sto.VectorizedReplayBufferStochastic, True, (5, 8), 4)):
        rng = np.random.default_rng(3 if stochastic else 2)
        cap, B, U, TD, T = 12, 32, 5, 7, 40
        for boot in (True, False):
            buf = cls(cap, B, U, TD, obs_shape=obs_shape, action_dim=A, max_episode_length=T, bootstrap_value_target=boot)
            tag = f"{name}_{int(boot)}"
            for rnd in range(3):   # 3 saves of 7 games: wraps around the ring of 12
                tr = make_traj(rng, 7, T, obs_shape, A, stochastic)
                buf.save_games_from_buffers(tr)
                for k, v in tr.items():
                    out[f"{tag}_traj{rnd}_{k}"] = v
                for smp in range(2):
                    seed = 100 * rnd + smp
                    ep, ts = replay_plan(seed, buf.size, buf.episode_lengths, B, U)
                    np.random.seed(seed)
                    batch = buf.sample_batch()
                    out[f"{tag}_r{rnd}_s{smp}_ep"] = ep
                    out[f"{tag}_r{rnd}_s{smp}_ts"] = ts
                    for k, v in batch.items():
                        out[f"{tag}_r{rnd}_s{smp}_out_{k}"] = np.asarray(v)
                out[f"{tag}_r{rnd}_position"] = np.int32(buf.position)
                out[f"{tag}_r{rnd}_size"] = np.int32(buf.size)
                out[f"{tag}_r{rnd}_lengths"] = buf.episode_lengths.copy()
                out[f"{tag}_r{rnd}_root_values"] = buf.root_values.copy()
                out[f"{tag}_r{rnd}_observations"] = buf.observations.copy()
    np.savez_compressed(os.path.join(HERE, "replay_reference.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
