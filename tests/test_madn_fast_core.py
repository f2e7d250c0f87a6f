"""The specialised (4 players, distance 10, canonical state) deterministic-MADN rules of csrc/madn_fast.cuh are plain
__host__ __device__ integer code: run them here on the CPU (tests/host_core harness, compiled with nvcc as host code)
against the oracle — mask by mask and step by step over whole random playouts and random rule sets."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle as O
from helpers import TRAIN_RULES, all_rule_sets, mask_of

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_core", "madn_fast_host.cu")
OUT = os.path.join(HERE, "host_core", "_build", "libhostcore.so")
CSRC = os.path.join(os.path.dirname(HERE), "exploring-muzero-on-dog_b200", "csrc")


@pytest.fixture(scope="module")
def hc():
    deps = [SRC, os.path.join(CSRC, "madn_fast.cuh"), os.path.join(CSRC, "madn_core.cuh"), os.path.join(CSRC, "madn_track.cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.run(["/usr/local/cuda/bin/nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                        "-shared", "-o", OUT, SRC], check=True)
    return C.CDLL(OUT)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _fast_mask(hc, s):
    m = np.zeros(s.n, np.uint32)
    gm = np.zeros(s.n, np.uint32)
    canon = np.zeros(s.n, np.uint8)
    rc = hc.hostcore_det_valid_mask4(C.c_int64(s.n), C.c_uint32(s.cfg.rules), 0, _p(s.board), _p(s.current_player), _p(s.pins),
                                     _p(s.reward), _p(s.done), _p(s.action_set), _p(m), _p(canon), _p(gm))
    assert rc == 0
    return m, canon.astype(bool), gm


def _fast_step(hc, s, action):
    stepped = np.zeros(s.n, np.uint8)
    still = np.zeros(s.n, np.uint8)
    action = np.ascontiguousarray(action, np.int32)
    rc = hc.hostcore_det_step4(C.c_int64(s.n), C.c_uint32(s.cfg.rules), _p(s.board), _p(s.current_player), _p(s.pins), _p(s.reward),
                               _p(s.done), _p(s.action_set), _p(action), _p(stepped), _p(still))
    assert rc == 0
    return stepped.astype(bool), still.astype(bool)


def _bits(mask24):
    return ((mask24[:, None] >> np.arange(24, dtype=np.uint32)[None, :]) & 1).astype(bool)


@pytest.mark.parametrize("seed", range(4))
def test_fast_core_playouts_match_oracle(hc, seed):
    rng = np.random.default_rng(100 + seed)
    stacked = 0
    for rules in all_rule_sets(rng, 5):
        n = 512
        cfg = O.MadnCfg(4, 0xF, 10, mask_of(rules))
        s = O.madn_reset(cfg, rng.integers(0, 1_000_000, n).astype(np.int32), int(rng.integers(0, 4)))
        for t in range(700):
            live = s.done == 0
            if not live.any():
                break
            want = O.madn_det_valid_action(s).reshape(n, 24)
            got, canon, generic = _fast_mask(hc, s)
            assert np.array_equal(_bits(generic), want), (rules, t)        # the generic core, on the CPU
            assert np.array_equal(_bits(got)[canon], want[canon]), (rules, t)  # the specialised core
            # random legal action per game (uniform over the legal set)
            score = np.where(want, rng.random((n, 24)), -1.0)
            a = score.argmax(1).astype(np.int32)
            has = want.any(1)
            f = s.copy()
            stepped, still = _fast_step(hc, f, a)
            assert np.array_equal(stepped, canon & live)
            # oracle: step where legal, no_step otherwise (copies merged by mask); finished games untouched
            o_step, o_skip = s.copy(), s.copy()
            O.madn_det_step(o_step, np.stack([a // 6, a % 6 + 1], 1).astype(np.int8))
            O.madn_det_no_step(o_skip)
            for k in s.fields():
                v = getattr(s, k)
                sel = has.reshape((-1,) + (1,) * (v.ndim - 1))
                lv = live.reshape((-1,) + (1,) * (v.ndim - 1))
                v[...] = np.where(lv, np.where(sel, getattr(o_step, k), getattr(o_skip, k)), v)
            for k in s.fields():
                a_, b_ = getattr(s, k)[stepped], getattr(f, k)[stepped]
                assert np.array_equal(a_, b_), (rules, t, k)
            # canonical form is closed under the reference's own transitions
            _, canon_after, _ = _fast_mask(hc, s)
            assert canon_after[stepped].all(), (rules, t)
            pins = s.pins.astype(np.int64)
            srt = np.sort(pins, axis=2)
            stacked += int(((srt[:, :, 1:] == srt[:, :, :-1]) & (srt[:, :, 1:] >= 0)).any((1, 2))[live].sum())
        assert t > 50
    assert stacked > 0  # the stacked-own-pins quirk (proxied home exit) was exercised


@pytest.mark.parametrize("seed", range(3))
def test_track_rules_playouts_match_oracle(hc, seed):
    """csrc/madn_track.cuh (track coordinates, training rule set — what the config-2 play kernel keeps in registers): mask by
    mask and leaf by leaf against the oracle over whole games, through the absolute <-> track conversions, incl. the
    team-proxy phase every game ends with and the stacked-own-pins quirk"""
    rng = np.random.default_rng(300 + seed)
    n = 768
    cfg = O.MadnCfg(4, 0xF, 10, mask_of(TRAIN_RULES))
    s = O.madn_reset(cfg, rng.integers(0, 1_000_000, n).astype(np.int32), int(rng.integers(0, 4)))
    proxied = captured = finished = 0
    for t in range(1200):
        live = s.done == 0
        if not live.any():
            break
        want = O.madn_det_valid_action(s).reshape(n, 24)
        m = np.zeros(n, np.uint32)
        cov = np.zeros(n, np.uint8)
        assert hc.hostcore_track_valid_mask(C.c_int64(n), _p(s.board), _p(s.current_player), _p(s.pins), _p(s.reward), _p(s.done),
                                            _p(s.action_set), _p(m), _p(cov)) == 0
        cov = cov.astype(bool)
        assert cov[live].all(), t                       # every live state the reference produces is covered
        assert not cov[~live].any(), t                  # a finished game (a complete team) is not
        assert np.array_equal(_bits(m)[live], want[live]), t
        score = np.where(want, rng.random((n, 24)), -1.0)
        a = score.argmax(1).astype(np.int32)
        has = want.any(1)
        f = s.copy()
        stepped = np.zeros(n, np.uint8)
        assert hc.hostcore_track_step(C.c_int64(n), _p(f.board), _p(f.current_player), _p(f.pins), _p(f.reward), _p(f.done),
                                      _p(f.action_set), _p(a), _p(stepped)) == 0
        assert np.array_equal(stepped.astype(bool), live)
        o_step, o_skip = s.copy(), s.copy()
        O.madn_det_step(o_step, np.stack([a // 6, a % 6 + 1], 1).astype(np.int8))
        O.madn_det_no_step(o_skip)
        cur = s.current_player.astype(np.int64)
        own_goal = (np.take_along_axis(s.pins.astype(np.int64), cur[:, None, None].repeat(4, 2), 1)[:, 0] >= 40).all(1)
        proxied += int((own_goal & live & has).sum())
        before_home = (s.pins < 0).sum((1, 2))
        for k in s.fields():
            v = getattr(s, k)
            sel = has.reshape((-1,) + (1,) * (v.ndim - 1))
            lv = live.reshape((-1,) + (1,) * (v.ndim - 1))
            v[...] = np.where(lv, np.where(sel, getattr(o_step, k), getattr(o_skip, k)), v)
        captured += int(((s.pins < 0).sum((1, 2)) > before_home).sum())
        finished += int(((s.done != 0) & live).sum())
        for k in s.fields():
            assert np.array_equal(getattr(s, k), getattr(f, k)), (t, k)
    assert finished == n and proxied > 1000 and captured > 1000


def test_track_rules_gate(hc):
    """states outside the track rules' domain are refused (the kernel then runs the generic rules)"""
    cfg = O.MadnCfg(4, 0xF, 10, mask_of(TRAIN_RULES))
    s = O.madn_reset(cfg, np.arange(5, dtype=np.int32), 0)
    s.action_set[1, 2, 3] = 9                       # a count that does not fit the packed row
    s.action_set[2, 0, 0] = -1
    s.pins[3, 0] = [40, 41, 42, 43]                 # team 0 / 2 already complete, done flag not set
    s.pins[3, 2] = [48, 49, 50, 51]
    s.pins[4, 1, 1] = 41                            # not canonical
    for g in (3, 4):
        s.board[g] = O.madn_set_pins_on_board(cfg, s.pins[g:g + 1])[0]
    m = np.zeros(5, np.uint32)
    cov = np.zeros(5, np.uint8)
    assert hc.hostcore_track_valid_mask(C.c_int64(5), _p(s.board), _p(s.current_player), _p(s.pins), _p(s.reward), _p(s.done),
                                        _p(s.action_set), _p(m), _p(cov)) == 0
    assert cov.tolist() == [1, 0, 0, 0, 0]


def test_non_canonical_states_are_detected(hc):
    cfg = O.MadnCfg(4, 0xF, 10, mask_of(TRAIN_RULES))
    s = O.madn_reset(cfg, np.arange(6, dtype=np.int32), 0)
    s.current_player[1] = 5                        # out-of-range player
    s.pins[2, 0, 1] = 60                           # off the board
    s.pins[3, 1, 1] = 41                           # player 1 in player 0's lane
    s.pins[4, 1, 1] = 0                            # two players on one cell
    s.board[5, 7] = 2                              # board does not match the pins
    for g in (2, 3, 4):
        s.board[g] = O.madn_set_pins_on_board(cfg, s.pins[g:g + 1])[0]
    _, canon, _ = _fast_mask(hc, s)
    assert canon.tolist() == [True, False, False, False, False, False]
