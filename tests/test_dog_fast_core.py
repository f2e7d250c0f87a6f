"""The specialised DOG rules of csrc/dog_fast.cuh (4 players, distance 10, canonical state, bitboards) are plain
__host__ __device__ code: run them here on the CPU (tests/host_core harness) against the oracle — the 806-wide
legal mask and the play-phase state transition, ply by ply over random playouts and random rule sets."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle as O
from helpers import DOG_RULES, dog_rule_sets, mask_of

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_core", "dog_fast_host.cu")
OUT = os.path.join(HERE, "host_core", "_build", "libhostcore_dog.so")
CSRC = os.path.join(os.path.dirname(HERE), "exploring-muzero-on-dog_b200", "csrc")


@pytest.fixture(scope="module")
def hc():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("dog_fast.cuh", "dog_core.cuh", "hostdev.cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.run(["/usr/local/cuda/bin/nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-diag-suppress", "550",
                        "-Xcompiler", "-fPIC", "-shared", "-o", OUT, SRC], check=True)
    return C.CDLL(OUT)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _fast_mask(hc, s):
    mask = np.zeros((s.n, 806), np.uint8)
    canon = np.zeros(s.n, np.uint8)
    soa = s.soa()
    assert hc.hostcore_dog_mask4(C.c_int64(s.n), C.c_uint32(s.cfg.rules), C.byref(soa), _p(mask), _p(canon)) == 0
    return mask.astype(bool), canon.astype(bool)


def _fast_play_step(hc, s, action):
    stepped = np.zeros(s.n, np.uint8)
    deal = np.zeros(s.n, np.uint8)
    action = np.ascontiguousarray(action, np.int32)
    soa = s.soa()
    assert hc.hostcore_dog_play_phase4(C.c_int64(s.n), C.c_uint32(s.cfg.rules), C.byref(soa), _p(action), _p(stepped), _p(deal)) == 0
    return stepped.astype(bool), deal.astype(bool)


def _subset(s, idx):
    o = O.DogState(s.cfg, len(idx))
    for k in s.FIELDS:
        setattr(o, k, np.ascontiguousarray(getattr(s, k)[idx]))
    return o


@pytest.mark.parametrize("seed", range(3))
def test_dog_fast_core_playouts_match_oracle(hc, seed):
    rng = np.random.default_rng(200 + seed)
    kinds = np.zeros(4, np.int64)
    for rules in dog_rule_sets(rng, 3):
        n = 256
        cfg = O.DogCfg(4, 0xF, 10, mask_of(rules))
        s = O.dog_reset(cfg, rng.integers(0, 1_000_000, n).astype(np.int32), int(rng.integers(0, 4)))
        for t in range(400):
            live = s.done == 0
            if not live.any():
                break
            want = O.dog_valid_actions(s)
            got, canon = _fast_mask(hc, s)
            assert canon[live].all(), (rules, t)  # canonical form is closed under the reference's transitions
            bad = np.argwhere(got[live] != want[live])
            assert bad.size == 0, (rules, t, bad[:5].tolist())
            # random legal action; every 9th ply an arbitrary play action (mostly illegal: reward -1 path)
            score = np.where(want, rng.random((n, 806)), -1.0)
            a = score.argmax(1).astype(np.int32)
            if t % 9 == 4:
                a = np.where(s.phase == 0, rng.integers(0, 792, n), a).astype(np.int32)
            has = want.any(1) | ((t % 9 == 4) & (s.phase == 0))
            play = live & has & (s.phase == 0)
            a_cat = (a % 396)
            for lo, hi, k in ((0, 224, 0), (224, 344, 1), (344, 392, 2), (392, 396, 3)):
                kinds[k] += int(((a_cat >= lo) & (a_cat < hi) & play & want[np.arange(n), a]).sum())
            # fast path on a copy (play-phase games only), the deal done by the oracle's distribute_cards
            f = s.copy()
            stepped, deal = _fast_play_step(hc, f, a)
            assert np.array_equal(stepped, s.phase == 0)
            idx = np.flatnonzero(deal)
            if idx.size:
                sub = _subset(f, idx)
                O.dog_distribute_cards(sub)
                for k in f.FIELDS:
                    getattr(f, k)[idx] = getattr(sub, k)
            # oracle transition: env_step where an action is taken, no_step otherwise; finished games untouched
            o_step, o_skip = s.copy(), s.copy()
            O.dog_step(o_step, a)
            O.dog_no_step(o_skip)
            for k in s.FIELDS:
                v = getattr(s, k)
                sel = has.reshape((-1,) + (1,) * (v.ndim - 1))
                lv = live.reshape((-1,) + (1,) * (v.ndim - 1))
                v[...] = np.where(lv, np.where(sel, getattr(o_step, k), getattr(o_skip, k)), v)
            for k in s.FIELDS:
                x, y = getattr(s, k)[play], getattr(f, k)[play]
                if not np.array_equal(x, y):
                    g = np.flatnonzero((x.reshape(x.shape[0], -1) != y.reshape(y.shape[0], -1)).any(1))[0]
                    raise AssertionError((rules, t, k, int(a[play][g]), x[g].tolist(), y[g].tolist()))
        assert t > 30
    assert (kinds > 0).all(), kinds  # swaps, hot sevens, normal moves and -4 moves were all exercised


def test_dog_non_canonical_states_are_detected(hc):
    cfg = O.DogCfg(4, 0xF, 10, mask_of(DOG_RULES))
    s = O.dog_reset(cfg, np.arange(5, dtype=np.int32), 0)
    s.pins[0, 0, 0] = 3
    s.board[0, 3] = 0
    s.current_player[1] = 4
    s.pins[2, 1, 0] = 56
    s.pins[3, 2, 1] = 41
    s.board[3, 41] = 2
    s.board[4, 9] = 1
    _, canon = _fast_mask(hc, s)
    assert canon.tolist() == [True, False, False, False, False]
