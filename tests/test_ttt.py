"""TicTacToe (config 1): the C oracle (CPU) and the CUDA kernels (GPU) against outputs of the reference's own
TicTacToe.py / TicTacToeV2.py (tests/golden/ttt_reference.npz, produced on the jaxshim)."""
import os

import numpy as np
import pytest

import oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ttt_reference.npz")
PLIES = 15  # 14 steps + final state per game


def _state_from(z, tag, variant, sel):
    s = O.TttState(len(sel), variant)
    s.board[...] = z[f"{tag}_board"][sel]
    s.current_player[...] = z[f"{tag}_cur"][sel]
    s.reward[...] = z[f"{tag}_reward"][sel]
    s.done[...] = z[f"{tag}_done"][sel]
    s.memory[...] = z[f"{tag}_memory"][sel]
    return s


@pytest.mark.parametrize("variant", [0, 1])
def test_oracle_env_step_and_policy(variant):
    z = np.load(G)
    tag = f"v{variant}"
    n = z[f"{tag}_board"].shape[0]
    idx = np.array([i for i in range(n) if i % PLIES != PLIES - 1])
    s = _state_from(z, tag, variant, idx)
    assert np.array_equal(O.ttt_policy(s), z[f"{tag}_policy"][idx])
    O.ttt_step(s, z[f"{tag}_action"][idx])
    nxt = idx + 1
    assert np.array_equal(s.board, z[f"{tag}_board"][nxt]) and np.array_equal(s.current_player, z[f"{tag}_cur"][nxt])
    assert np.array_equal(s.reward, z[f"{tag}_reward"][nxt]) and np.array_equal(s.done.astype(bool), z[f"{tag}_done"][nxt])
    if variant == 1:
        assert np.array_equal(s.memory, z[f"{tag}_memory"][nxt])


@pytest.mark.parametrize("variant", [0, 1])
def test_oracle_root_and_recurrent_fn(variant):
    z = np.load(G)
    tag = f"v{variant}"
    idx = np.flatnonzero(z[f"{tag}_rec_action"] >= 0)
    assert idx.size > 30
    s = _state_from(z, tag, variant, idx)
    keys = z[f"{tag}_key"][idx]
    prior, value, emb = O.ttt_root_fn(s, keys)
    assert np.array_equal(prior, z[f"{tag}_policy"][idx]) and np.array_equal(value, z[f"{tag}_root_value"][idx])
    p2, v2, r2, d2, e2 = O.ttt_recurrent_fn(variant, keys, z[f"{tag}_rec_action"][idx], emb)
    assert np.array_equal(p2, z[f"{tag}_rec_prior"][idx]) and np.array_equal(v2, z[f"{tag}_rec_value"][idx])
    assert np.array_equal(r2, z[f"{tag}_rec_reward"][idx]) and np.array_equal(d2, z[f"{tag}_rec_discount"][idx])
    assert np.array_equal(e2[:, :9].reshape(-1, 3, 3), z[f"{tag}_rec_board"][idx]) and np.array_equal(e2[:, 9], z[f"{tag}_rec_cur"][idx])
    assert np.array_equal(e2[:, 11].astype(bool), z[f"{tag}_rec_done"][idx])
