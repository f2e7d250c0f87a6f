"""CPU oracle — TEST INFRASTRUCTURE ONLY.

A plain-C restatement of the reference's hot path (see the headers of the *_oracle.c files for
the reference lines each function follows), loaded through ctypes with NumPy arrays.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may
import this package; the product package never does (tests/test_no_oracle_in_product.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    """Compile oracle/*_oracle.c into oracle/_build/liboracle.so with gcc."""
    srcs = sorted(f for f in os.listdir(_HERE) if f.endswith("_oracle.c"))
    newest = max(os.path.getmtime(os.path.join(_HERE, f)) for f in os.listdir(_HERE) if f.endswith((".c", ".h")))
    if not force and os.path.exists(_LIB_PATH) and os.path.getmtime(_LIB_PATH) >= newest:
        return _LIB_PATH
    os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
    cmd = ["gcc", "-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-shared", "-o", _LIB_PATH]
    cmd += [os.path.join(_HERE, s) for s in srcs] + ["-lm", "-lpthread"]
    subprocess.run(cmd, check=True, cwd=_HERE)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    a = np.ascontiguousarray(a, dtype=dt)
    return a


# --------------------------------------------------------------------------- jax.random restatement
def threefry2x32(k0, k1, c0, c1):
    c0 = _c(c0, np.uint32).ravel()
    c1 = _c(c1, np.uint32).ravel()
    o0 = np.empty_like(c0)
    o1 = np.empty_like(c1)
    lib().orc_threefry2x32_v(C.c_uint32(k0), C.c_uint32(k1), C.c_int64(c0.size), _p(c0), _p(c1), _p(o0), _p(o1))
    return o0, o1


def prng_key(seed):
    return np.array([0, np.uint32(np.int64(seed) & 0xFFFFFFFF)], dtype=np.uint32)


def split(key, n=2):
    key = _c(key, np.uint32)
    out = np.empty((n, 2), np.uint32)
    lib().orc_split(_p(key), C.c_int64(n), _p(out))
    return out


def random_bits(key, n):
    key = _c(key, np.uint32)
    out = np.empty(n, np.uint32)
    lib().orc_random_bits(_p(key), C.c_int64(n), _p(out))
    return out


def uniform(key, n, minval=0.0, maxval=1.0):
    key = _c(key, np.uint32)
    out = np.empty(n, np.float32)
    lib().orc_uniform(_p(key), C.c_int64(n), C.c_float(minval), C.c_float(maxval), _p(out))
    return out


def randint(key, n, lo, hi):
    key = _c(key, np.uint32)
    out = np.empty(n, np.int32)
    lib().orc_randint(_p(key), C.c_int64(n), C.c_int32(lo), C.c_int32(hi), _p(out))
    return out


def gumbel(key, n):
    key = _c(key, np.uint32)
    out = np.empty(n, np.float32)
    lib().orc_gumbel(_p(key), C.c_int64(n), _p(out))
    return out


def choice6(key, p):
    key = _c(key, np.uint32)
    p = _c(p, np.float32)
    return int(lib().orc_choice6_v(_p(key), _p(p)))


def categorical_masked(key, mask, float_gumbel=False):
    key = _c(key, np.uint32)
    mask = _c(mask, np.uint8)
    return int(lib().orc_categorical(_p(key), _p(mask), C.c_int(mask.size), C.c_int(int(float_gumbel))))


# --------------------------------------------------------------------------- MADN
class MadnCfg:
    """(num_players, layout, distance, rules bitmask) exactly as handed to the C-ABI."""

    def __init__(self, num_players=4, layout_mask=0xF, distance=10, rules=0):
        self.num_players, self.layout_mask, self.distance, self.rules = num_players, layout_mask, distance, rules
        self.total = 4 * distance + 16

    @property
    def args(self):
        return (C.c_int(self.num_players), C.c_int(self.layout_mask), C.c_int(self.distance), C.c_uint32(self.rules))

    def geometry(self):
        n = self.num_players
        s = np.zeros(n, np.int32)
        t = np.zeros(n, np.int32)
        g = np.zeros((n, 4), np.int32)
        assert lib().orc_madn_geometry(*self.args, _p(s), _p(t), _p(g)) == 0
        return s, t, g


class MadnState:
    """SoA leaves of a batch of deterministic/classic MADN games (NumPy, host)."""

    def __init__(self, cfg, n, det=True):
        P = cfg.num_players
        self.cfg, self.n, self.det = cfg, n, det
        self.board = np.full((n, cfg.total), -1, np.int8)
        self.current_player = np.zeros(n, np.int8)
        self.pins = np.full((n, P, 4), -1, np.int8)
        self.reward = np.zeros(n, np.int8)
        self.done = np.zeros(n, np.uint8)
        self.action_set = np.full((n, P, 6), 4, np.int8) if det else None
        self.die = None if det else np.zeros(n, np.int8)
        self.key = np.zeros((n, 2), np.uint32)

    def copy(self):
        o = MadnState.__new__(MadnState)
        o.cfg, o.n, o.det = self.cfg, self.n, self.det
        for f in ("board", "current_player", "pins", "reward", "done", "action_set", "die", "key"):
            v = getattr(self, f)
            setattr(o, f, None if v is None else v.copy())
        return o

    def fields(self):
        names = ["board", "current_player", "pins", "reward", "done", "key"] + (["action_set"] if self.det else ["die"])
        return {k: getattr(self, k) for k in names}


def madn_reset(cfg, seeds, starting_player=0, det=True):
    seeds = _c(seeds, np.int32)
    s = MadnState(cfg, seeds.size, det)
    rc = lib().orc_madn_reset(*cfg.args, C.c_int64(s.n), _p(seeds), C.c_int(starting_player), _p(s.board),
                              _p(s.current_player), _p(s.pins), _p(s.reward), _p(s.done), _p(s.action_set),
                              _p(s.die), _p(s.key))
    assert rc == 0
    return s


def madn_set_pins_on_board(cfg, pins):
    pins = _c(pins, np.int8)
    n = pins.shape[0]
    board = np.empty((n, cfg.total), np.int8)
    assert lib().orc_madn_set_pins_on_board(*cfg.args, C.c_int64(n), _p(pins), _p(board)) == 0
    return board


def madn_det_valid_action(s):
    mask = np.empty((s.n, 4, 6), np.uint8)
    assert lib().orc_madn_det_valid_action(*s.cfg.args, C.c_int64(s.n), _p(s.board), _p(s.current_player), _p(s.pins),
                                           _p(s.action_set), _p(mask)) == 0
    return mask.astype(bool)


def _det_args(s):
    return (C.c_int64(s.n), _p(s.board), _p(s.current_player), _p(s.pins), _p(s.reward), _p(s.done), _p(s.action_set))


def madn_det_embed_dim(cfg):
    return int(lib().orc_madn_det_embed_dim(*cfg.args))


def madn_det_policy_function(s):
    """policy_function of the true-env search (deterministic_madn.py:495-507) -> f32 [n, 24]"""
    out = np.empty((s.n, 24), np.float32)
    assert lib().orc_madn_det_policy_function(*s.cfg.args, *_det_args(s), _p(out)) == 0
    return out


def madn_det_root_fn(s, keys):
    """root_fn (:551-566) -> (prior f32 [n, 24], value f32 [n], embedding f32 [n, E])"""
    keys = _c(keys, np.uint32)
    prior, value = np.empty((s.n, 24), np.float32), np.empty(s.n, np.float32)
    emb = np.empty((s.n, madn_det_embed_dim(s.cfg)), np.float32)
    assert lib().orc_madn_det_root_fn(*s.cfg.args, *_det_args(s), _p(keys), _p(prior), _p(value), _p(emb)) == 0
    return prior, value, emb


def madn_det_recurrent_fn(cfg, keys, action, emb):
    """recurrent_fn (:568-590) on the env embedding -> (prior, value, reward, discount, embedding_out)"""
    keys, action, emb = _c(keys, np.uint32), _c(action, np.int32), _c(emb, np.float32)
    n = action.size
    prior, value, reward, discount = np.empty((n, 24), np.float32), np.empty(n, np.float32), np.empty(n, np.float32), np.empty(n, np.float32)
    emb_out = np.empty_like(emb)
    assert lib().orc_madn_det_recurrent_fn(*cfg.args, C.c_int64(n), _p(keys), _p(action), _p(emb), _p(prior), _p(value), _p(reward),
                                           _p(discount), _p(emb_out)) == 0
    return prior, value, reward, discount, emb_out


def madn_det_embedding(s):
    """the env as floats: board, current_player, pins, reward, done, action_set"""
    n = s.n
    return np.concatenate([s.board.reshape(n, -1), s.current_player.reshape(n, 1), s.pins.reshape(n, -1), s.reward.reshape(n, 1),
                           s.done.reshape(n, 1), s.action_set.reshape(n, -1)], axis=1).astype(np.float32)


def _cls_args(s):
    return (C.c_int64(s.n), _p(s.board), _p(s.current_player), _p(s.pins), _p(s.reward), _p(s.done), _p(s.die), _p(s.key))


def madn_cls_embed_dim(cfg):
    return int(lib().orc_madn_cls_embed_dim(*cfg.args))


def madn_cls_embedding(s):
    """the dice env as floats: board, current_player, pins, reward, done, die, key as four 16-bit halves"""
    n = s.n
    k = s.key.astype(np.uint32)
    halves = np.stack([k[:, 0] & 0xFFFF, k[:, 0] >> 16, k[:, 1] & 0xFFFF, k[:, 1] >> 16], 1)
    return np.concatenate([s.board.reshape(n, -1), s.current_player.reshape(n, 1), s.pins.reshape(n, -1), s.reward.reshape(n, 1),
                           s.done.reshape(n, 1), s.die.reshape(n, 1), halves], axis=1).astype(np.float32)


def madn_cls_policy_function(s):
    """policy_function of the dice game's true-env search (classic_madn.py:571-583) -> f32 [n, 4]"""
    out = np.empty((s.n, 4), np.float32)
    assert lib().orc_madn_cls_policy_function(*s.cfg.args, *_cls_args(s), _p(out)) == 0
    return out


def madn_cls_root_fn(s, keys):
    keys = _c(keys, np.uint32)
    prior, value = np.empty((s.n, 4), np.float32), np.empty(s.n, np.float32)
    emb = np.empty((s.n, madn_cls_embed_dim(s.cfg)), np.float32)
    assert lib().orc_madn_cls_root_fn(*s.cfg.args, *_cls_args(s), _p(keys), _p(prior), _p(value), _p(emb)) == 0
    return prior, value, emb


def madn_cls_decision_recurrent_fn(cfg, keys, action, emb):
    """recurrent_fn (:657-688) -> (chance_logits f32 [n, 6], afterstate_value, afterstate embedding)"""
    keys, action, emb = _c(keys, np.uint32), _c(action, np.int32), _c(emb, np.float32)
    n = action.size
    cl, av, out = np.empty((n, 6), np.float32), np.empty(n, np.float32), np.empty_like(emb)
    assert lib().orc_madn_cls_decision_recurrent_fn(*cfg.args, C.c_int64(n), _p(keys), _p(action), _p(emb), _p(cl), _p(av), _p(out)) == 0
    return cl, av, out


def madn_cls_chance_recurrent_fn(cfg, keys, outcome, emb):
    """recurrent_chance_fn (:624-655) -> (action_logits f32 [n, 4], value, reward, discount, embedding)"""
    keys, outcome, emb = _c(keys, np.uint32), _c(outcome, np.int32), _c(emb, np.float32)
    n = outcome.size
    al, v, r, d, out = np.empty((n, 4), np.float32), np.empty(n, np.float32), np.empty(n, np.float32), np.empty(n, np.float32), np.empty_like(emb)
    assert lib().orc_madn_cls_chance_recurrent_fn(*cfg.args, C.c_int64(n), _p(keys), _p(outcome), _p(emb), _p(al), _p(v), _p(r), _p(d),
                                                  _p(out)) == 0
    return al, v, r, d, out


def madn_det_step(s, action):
    """in place; returns (reward, done)"""
    action = _c(action, np.int8).reshape(s.n, 2)
    assert lib().orc_madn_det_step(*s.cfg.args, C.c_int64(s.n), _p(s.board), _p(s.current_player), _p(s.pins),
                                   _p(s.reward), _p(s.done), _p(s.action_set), _p(action)) == 0
    return s.reward.copy(), s.done.astype(bool)


def madn_det_no_step(s):
    assert lib().orc_madn_det_no_step(*s.cfg.args, C.c_int64(s.n), _p(s.current_player), _p(s.reward), _p(s.done),
                                      _p(s.action_set)) == 0
    return np.zeros(s.n, np.int8), s.done.astype(bool)


def madn_det_encode_board(s):
    obs = np.empty((s.n, 8 * s.cfg.num_players + 2, s.cfg.total), np.int8)
    assert lib().orc_madn_det_encode_board(*s.cfg.args, C.c_int64(s.n), _p(s.board), _p(s.current_player), _p(s.pins),
                                           _p(s.action_set), _p(obs)) == 0
    return obs


def madn_cls_valid_action(s):
    mask = np.empty((s.n, 4), np.uint8)
    assert lib().orc_madn_cls_valid_action(*s.cfg.args, C.c_int64(s.n), _p(s.board), _p(s.current_player), _p(s.pins),
                                           _p(s.die), _p(mask)) == 0
    return mask.astype(bool)


def madn_cls_step(s, pin):
    pin = _c(pin, np.int8).reshape(s.n)
    assert lib().orc_madn_cls_step(*s.cfg.args, C.c_int64(s.n), _p(s.board), _p(s.current_player), _p(s.pins),
                                   _p(s.reward), _p(s.done), _p(s.die), _p(pin)) == 0
    return s.reward.copy(), s.done.astype(bool)


def madn_cls_no_step(s):
    assert lib().orc_madn_cls_no_step(*s.cfg.args, C.c_int64(s.n), _p(s.current_player)) == 0
    return np.zeros(s.n, np.int8), s.done.astype(bool)


def madn_cls_throw_die(s):
    assert lib().orc_madn_cls_throw_die(*s.cfg.args, C.c_int64(s.n), _p(s.board), _p(s.current_player), _p(s.pins),
                                        _p(s.key), _p(s.die)) == 0
    return s.die.copy()


def madn_cls_dice_probabilities(s):
    p = np.empty((s.n, 6), np.float32)
    assert lib().orc_madn_cls_dice_probabilities(*s.cfg.args, C.c_int64(s.n), _p(s.board), _p(s.current_player),
                                                 _p(s.pins), _p(p)) == 0
    return p


def madn_cls_encode_board(s):
    obs = np.empty((s.n, 2 * s.cfg.num_players + 3, s.cfg.total), np.int8)
    assert lib().orc_madn_cls_encode_board(*s.cfg.args, C.c_int64(s.n), _p(s.board), _p(s.current_player), _p(s.pins),
                                           _p(s.die), _p(obs)) == 0
    return obs


def madn_det_play_random(s, rng_key, max_steps=2000, game_offset=0, float_gumbel=False, nthreads=1):
    """Lockstep random-legal-policy play to termination, in place.  Returns (game_len, total_steps, rng_key_out)."""
    key = _c(rng_key, np.uint32).copy()
    game_len = np.zeros(s.n, np.int32)
    total = C.c_int64(0)
    rc = lib().orc_madn_det_play_random(*s.cfg.args, C.c_int64(s.n), C.c_int64(game_offset), _p(s.board),
                                        _p(s.current_player), _p(s.pins), _p(s.reward), _p(s.done), _p(s.action_set),
                                        _p(key), C.c_int(max_steps), C.c_int(int(float_gumbel)), C.c_int(nthreads),
                                        _p(game_len), C.byref(total))
    assert rc == 0
    return game_len, int(total.value), key


# --------------------------------------------------------------------------- DOG
class _DogSoa(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("board", "cur", "pins", "reward", "done", "deck", "hands", "swap_choices",
                                          "round_starter", "phase", "key", "hand_size")]


class DogCfg(MadnCfg):
    @property
    def num_actions(self):
        return 2 * (4 * (13 + self.total) + 120) + 14


class DogState:
    """SoA leaves of a batch of DOG games (NumPy, host); dtypes as in the reference dataclass (DOG/dog.py:31-56)."""
    FIELDS = ("board", "current_player", "pins", "reward", "done", "deck", "hands", "swap_choices", "round_starter",
              "phase", "key", "hand_size")

    def __init__(self, cfg, n):
        P = cfg.num_players
        self.cfg, self.n = cfg, n
        self.board = np.full((n, cfg.total), -1, np.int8)
        self.current_player = np.zeros(n, np.int8)
        self.pins = np.full((n, P, 4), -1, np.int32)
        self.reward = np.zeros(n, np.int8)
        self.done = np.zeros(n, np.uint8)
        self.deck = np.zeros((n, 14), np.int8)
        self.hands = np.zeros((n, P, 14), np.int8)
        self.swap_choices = np.full((n, 4), -1, np.int8)
        self.round_starter = np.full(n, -1, np.int8)
        self.phase = np.zeros(n, np.int8)
        self.key = np.zeros((n, 2), np.uint32)
        self.hand_size = np.full(n, 6, np.int8)

    def soa(self):
        return _DogSoa(*[_p(getattr(self, k)) for k in self.FIELDS])

    def fields(self):
        return {k: getattr(self, k) for k in self.FIELDS}

    def copy(self):
        o = DogState.__new__(DogState)
        o.cfg, o.n = self.cfg, self.n
        for k in self.FIELDS:
            setattr(o, k, getattr(self, k).copy())
        return o


def dog_reset(cfg, seeds, starting_player=0):
    seeds = _c(seeds, np.int32)
    s = DogState(cfg, seeds.size)
    soa = s.soa()
    assert lib().orc_dog_reset(*cfg.args, C.c_int64(s.n), _p(seeds), C.c_int(starting_player), C.byref(soa)) == 0
    return s


def dog_valid_actions(s):
    mask = np.empty((s.n, s.cfg.num_actions), np.uint8)
    soa = s.soa()
    assert lib().orc_dog_valid_actions(*s.cfg.args, C.c_int64(s.n), C.byref(soa), _p(mask)) == 0
    return mask.astype(bool)


def dog_step(s, action):
    action = _c(action, np.int32).reshape(s.n)
    reward = np.empty(s.n, np.int8)
    done = np.empty(s.n, np.uint8)
    soa = s.soa()
    assert lib().orc_dog_step(*s.cfg.args, C.c_int64(s.n), C.byref(soa), _p(action), _p(reward), _p(done)) == 0
    return reward, done.astype(bool)


def dog_no_step(s):
    reward = np.empty(s.n, np.int8)
    done = np.empty(s.n, np.uint8)
    soa = s.soa()
    assert lib().orc_dog_no_step(*s.cfg.args, C.c_int64(s.n), C.byref(soa), _p(reward), _p(done)) == 0
    return reward, done.astype(bool)


def dog_distribute_cards(s):
    soa = s.soa()
    assert lib().orc_dog_distribute_cards(*s.cfg.args, C.c_int64(s.n), C.byref(soa)) == 0


def dog_map_action_to_move(cfg, action):
    action = _c(action, np.int32).ravel()
    out = np.empty((action.size, 6), np.int32)
    assert lib().orc_dog_map_action_to_move(*cfg.args, C.c_int64(action.size), _p(action), _p(out)) == 0
    return out


def dog_substep(s, kind, args):
    """reference sub-steps (0 normal, 1 neg, 2 swap, 3 hot-7) -> (board, pins, reward, done); env untouched"""
    kind = _c(kind, np.int32).reshape(s.n)
    args = _c(args, np.int32).reshape(s.n, 4)
    board = np.empty_like(s.board)
    pins = np.empty_like(s.pins)
    reward = np.empty(s.n, np.int8)
    done = np.empty(s.n, np.uint8)
    soa = s.soa()
    assert lib().orc_dog_substep(*s.cfg.args, C.c_int64(s.n), C.byref(soa), _p(kind), _p(args), _p(board), _p(pins),
                                 _p(reward), _p(done)) == 0
    return board, pins, reward, done.astype(bool)


def dog_play_random(s, rng_key, max_steps=2000, game_offset=0, float_gumbel=False, nthreads=1):
    key = _c(rng_key, np.uint32).copy()
    game_len = np.zeros(s.n, np.int32)
    total = C.c_int64(0)
    soa = s.soa()
    assert lib().orc_dog_play_random(*s.cfg.args, C.c_int64(s.n), C.c_int64(game_offset), C.byref(soa), _p(key),
                                     C.c_int(max_steps), C.c_int(int(float_gumbel)), C.c_int(nthreads), _p(game_len),
                                     C.byref(total)) == 0
    return game_len, int(total.value), key


# --------------------------------------------------------------------------- MCTS (mctx restatement)
class MctsCfg(C.Structure):
    _fields_ = [("policy", C.c_int32), ("qtransform", C.c_int32), ("num_simulations", C.c_int32), ("max_depth", C.c_int32),
                ("num_actions", C.c_int32), ("num_chance", C.c_int32), ("embed_dim", C.c_int32),
                ("max_num_considered_actions", C.c_int32), ("q_min", C.c_float), ("q_max", C.c_float),
                ("value_scale", C.c_float), ("maxvisit_init", C.c_float), ("epsilon", C.c_float), ("pb_c_init", C.c_float),
                ("pb_c_base", C.c_float), ("dirichlet_fraction", C.c_float), ("temperature", C.c_float),
                ("gumbel_scale", C.c_float)]


_MCTS_FIELDS = (("node_visits", np.int32, "N"), ("raw_values", np.float32, "N"), ("node_values", np.float32, "N"),
                ("parents", np.int32, "N"), ("action_from_parent", np.int32, "N"), ("children_index", np.int32, "NA"),
                ("children_prior_logits", np.float32, "NA"), ("children_visits", np.int32, "NA"),
                ("children_rewards", np.float32, "NA"), ("children_discounts", np.float32, "NA"),
                ("children_values", np.float32, "NA"), ("embeddings", np.float32, "NE"), ("is_decision", np.uint8, "N"),
                ("root_invalid_actions", np.uint8, "A"), ("root_gumbel", np.float32, "A"), ("search_key", np.uint32, "2"),
                ("policy_key", np.uint32, "2"))


class _MctsTreeC(C.Structure):
    _fields_ = [(k, C.c_void_p) for k, _, _ in _MCTS_FIELDS]


class MctsTree:
    def __init__(self, cfg, n):
        N, A, E = cfg.num_simulations + 1, cfg.num_actions + cfg.num_chance, cfg.embed_dim
        shp = {"N": (n, N), "NA": (n, N, A), "NE": (n, N, E), "A": (n, A), "2": (n, 2)}
        for k, dt, s in _MCTS_FIELDS:
            setattr(self, k, np.zeros(shp[s], dt))
        self.cfg, self.n = cfg, n

    def c(self):
        return _MctsTreeC(*[_p(getattr(self, k)) for k, _, _ in _MCTS_FIELDS])


def mcts_init(tree, keys, root_prior, root_value, root_emb, invalid=None, noise=None):
    t = tree.c()
    keys, root_prior = _c(keys, np.uint32), _c(root_prior, np.float32)
    root_value, root_emb = _c(root_value, np.float32), _c(root_emb, np.float32)
    invalid = None if invalid is None else _c(invalid, np.uint8)
    noise = None if noise is None else _c(noise, np.float32)
    assert lib().orc_mcts_init(C.byref(t), C.c_int64(tree.n), C.byref(tree.cfg), _p(keys), _p(root_prior), _p(root_value),
                               _p(root_emb), _p(invalid), _p(noise)) == 0


def mcts_select(tree, sim):
    t = tree.c()
    parent = np.empty(tree.n, np.int32)
    action = np.empty(tree.n, np.int32)
    emb = np.empty((tree.n, tree.cfg.embed_dim), np.float32)
    isdec = np.empty(tree.n, np.uint8)
    tree.expand_key = np.empty((tree.n, 2), np.uint32)
    assert lib().orc_mcts_select(C.byref(t), C.c_int64(tree.n), C.byref(tree.cfg), C.c_int32(sim), _p(parent), _p(action),
                                 _p(emb), _p(isdec), _p(tree.expand_key)) == 0
    return parent, action, emb, isdec


def mcts_expand(tree, sim, parent, action, prior, value, reward, discount, emb, chance_logits=None, after_value=None, after_emb=None):
    t = tree.c()
    f = lambda x: None if x is None else _c(x, np.float32)
    args = [f(prior), f(value), f(reward), f(discount), f(emb), f(chance_logits), f(after_value), f(after_emb)]
    parent, action = _c(parent, np.int32), _c(action, np.int32)
    assert lib().orc_mcts_expand(C.byref(t), C.c_int64(tree.n), C.byref(tree.cfg), C.c_int32(sim), _p(parent), _p(action),
                                 *[_p(a) for a in args]) == 0


def mcts_policy_output(tree):
    t = tree.c()
    action = np.empty(tree.n, np.int32)
    weights = np.empty((tree.n, tree.cfg.num_actions), np.float32)
    value = np.empty(tree.n, np.float32)
    assert lib().orc_mcts_policy_output(C.byref(t), C.c_int64(tree.n), C.byref(tree.cfg), _p(action), _p(weights), _p(value)) == 0
    return action, weights, value


def considered_visit(m, S, i):
    return int(lib().orc_considered_visit(C.c_int(m), C.c_int(S), C.c_int(i)))


# --------------------------------------------------------------------------- TicTacToe
class TttState:
    FIELDS = ("board", "current_player", "reward", "done", "memory")

    def __init__(self, n, variant=1):
        self.n, self.variant = n, variant
        self.board = np.zeros((n, 3, 3), np.int8)
        self.current_player = np.ones(n, np.int8)
        self.reward = np.zeros(n, np.int8)
        self.done = np.zeros(n, np.uint8)
        self.memory = np.full((n, 2, 3), -1, np.int8)

    def args(self):
        return (_p(self.board), _p(self.current_player), _p(self.reward), _p(self.done), _p(self.memory))

    def fields(self):
        return {k: getattr(self, k) for k in self.FIELDS}

    def copy(self):
        o = TttState(self.n, self.variant)
        for k in self.FIELDS:
            setattr(o, k, getattr(self, k).copy())
        return o


def ttt_step(s, action):
    action = _c(action, np.int8).reshape(s.n)
    lib().orc_ttt_step(C.c_int(s.variant), C.c_int64(s.n), *s.args(), _p(action))
    return s.reward.copy(), s.done.astype(bool)


def ttt_policy(s):
    out = np.empty((s.n, 9), np.float32)
    lib().orc_ttt_policy(C.c_int(s.variant), C.c_int64(s.n), *s.args(), _p(out))
    return out


def ttt_root_fn(s, keys):
    keys = _c(keys, np.uint32)
    prior, value, emb = np.empty((s.n, 9), np.float32), np.empty(s.n, np.float32), np.empty((s.n, 18), np.float32)
    lib().orc_ttt_root_fn(C.c_int(s.variant), C.c_int64(s.n), *s.args(), _p(keys), _p(prior), _p(value), _p(emb))
    return prior, value, emb


def ttt_recurrent_fn(variant, keys, action, emb):
    keys, action, emb = _c(keys, np.uint32), _c(action, np.int32), _c(emb, np.float32)
    n = action.size
    prior, value, reward, discount = np.empty((n, 9), np.float32), np.empty(n, np.float32), np.empty(n, np.float32), np.empty(n, np.float32)
    emb_out = np.empty((n, 18), np.float32)
    lib().orc_ttt_recurrent_fn(C.c_int(variant), C.c_int64(n), _p(keys), _p(action), _p(emb), _p(prior), _p(value), _p(reward),
                               _p(discount), _p(emb_out))
    return prior, value, reward, discount, emb_out
