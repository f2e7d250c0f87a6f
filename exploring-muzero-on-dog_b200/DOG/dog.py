"""Host mirror of the reference's DOG/dog.py for the self-play hot path.

Same names / arguments / return tuples as /root/reference/DOG/dog.py — get_play_action_size :58, env_reset :83,
distribute_cards :201, set_pins_on_board :346, valid_actions :693, no_step :713, step_swap :755,
step_normal_move :790, step_neg_move :861, step_hot_7 :913, env_step :1117, map_action_to_move :1134,
map_move_to_action :1198, map_action_to_card :1241 — each already vmapped over the leading game axis and run by one
libdogstep.so kernel (one warp per game).  Pure by default; `inplace=True` updates the leaves in place.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib, rules as _rules
from ..MADN._state import BatchedEnv, reuse_or_alloc, seeds_to_dev, to_dev
from ..MADN.deterministic_madn import _geometry, _layout_mask, _out


def all_pin_distributions(total=7):
    """utils/utility_funcs.py:4-21 — the 120 splits of 7 over four pins, lexicographic in (a, b, c)."""
    return np.array([[a, b, c, total - a - b - c] for a in range(total + 1) for b in range(total + 1)
                     for c in range(total + 1) if total - a - b - c >= 0], dtype=np.int32)


DISTS_7_4 = all_pin_distributions(7)


class DOG(BatchedEnv):
    """Batched leaves of the reference dataclass (:31-56)."""

    LEAVES = {
        "board": (torch.int8, lambda s: (s["total_board_size"],)),
        "current_player": (torch.int8, lambda s: ()),
        "pins": (torch.int32, lambda s: (s["num_players"], 4)),
        "reward": (torch.int8, lambda s: ()),
        "done": (torch.bool, lambda s: ()),
        "deck": (torch.int8, lambda s: (14,)),
        "hands": (torch.int8, lambda s: (s["num_players"], 14)),
        "swap_choices": (torch.int8, lambda s: (4,)),
        "round_starter": (torch.int8, lambda s: ()),
        "phase": (torch.int8, lambda s: ()),
        "key": (torch.uint32, lambda s: (2,)),
        "hand_size": (torch.int8, lambda s: ()),
    }

    def _const(self, name):
        a = torch.as_tensor(self.static["_" + name].astype(np.int32), device=self.device)
        return a.expand((self.n,) + tuple(a.shape)) if self.batched else a

    start = property(lambda self: self._const("start"))
    target = property(lambda self: self._const("target"))
    goal = property(lambda self: self._const("goal"))

    def cfg(self):
        return self.memo("cfg", self._make_cfg)

    def _make_cfg(self):
        s = self.static
        return _lib.MadnCfg(s["num_players"], s["layout_mask"], s["board_size"] // 4, _rules.to_mask(s["rules"]))

    def cstate(self):
        return self.memo("cstate", self._make_cstate)

    def _make_cstate(self):
        t = self._t
        return _lib.tag(_lib.DogState(*[C.c_void_p(t[k].data_ptr()) for k in
                                        ("board", "current_player", "pins", "reward", "done", "deck", "hands", "swap_choices",
                                         "round_starter", "phase", "key", "hand_size")]), t["board"].device)


RAW_OBS_SIZE = 56 + 14 + 4


def raw_observation(env):
    """int8 [n, 74]: board | the mover's own hand | phase, hand_size, current_player, round_starter.  NOT a reference function:
    the reference has no DOG encoder (DOG/dog.py:1264-1272 is a TODO); config 5's self-play slice needs some observation row to
    record, and this one exposes only what the mover may see.  Plain tensor indexing (plumbing, not a kernel)."""
    t = env._t
    n = env.n
    cur = t["current_player"].reshape(n).long() % env.static["num_players"]
    own = t["hands"].reshape(n, env.static["num_players"], 14)[torch.arange(n, device=env.device), cur]
    misc = torch.stack([t["phase"].reshape(n), t["hand_size"].reshape(n), t["current_player"].reshape(n),
                        t["round_starter"].reshape(n)], 1)
    return torch.cat([t["board"].reshape(n, -1), own, misc], 1).to(torch.int8).contiguous()


def obs_planes(env):
    """number of planes of encode_board: 8 + 3 * num_players + 14 (34 for four players)"""
    return 8 + 3 * env.static["num_players"] + 14


def encode_board(env):
    """Observation of the seat to move, int8 [n, obs_planes, total_board_size].  NOT a reference function: DOG/dog.py:1264-1272
    is a TODO and the reference has no DOG networks; the layout continues the MADN encoders (mover's frame, scalar facts as
    planes, only what the mover may know — its own hand, the others' hand sizes).  Plane list: include/dogstep.h."""
    obs = torch.empty((env.n, obs_planes(env), env.static["total_board_size"]), dtype=torch.int8, device=env.device)
    cfg, st = env.cfg(), env.cstate()
    _lib.check(_lib.lib().dogstep_dog_encode_board(C.byref(st), C.c_int64(env.n), C.byref(cfg), _lib.ptr(obs), _lib.stream()),
               "dog_encode_board")
    return _out(env, obs)


def get_play_action_size(env):
    return int(2 * (4 * (12 + 1 + env.static["total_board_size"]) + 120))


def env_reset(_, num_players=4, layout=None, distance=10, starting_player=0, seed=42, enable_teams=False,
              enable_initial_free_pin=False, enable_circular_board=True, enable_start_blocking=False,
              enable_jump_in_goal_area=True, enable_friendly_fire=False, must_traverse_start=True, disable_swapping=False,
              disable_hot_seven=False, disable_joker=False, device="cuda", out=None):
    """env_reset (:83-181), seed scalar or int array [n]."""
    num_players, distance = int(num_players), int(distance)
    batched, seeds = seeds_to_dev(seed, device)
    lm = _layout_mask(layout)
    start, target, goal = _geometry(num_players, lm, distance)
    rules = dict(enable_teams=bool(enable_teams) and num_players == 4, enable_initial_free_pin=bool(enable_initial_free_pin),
                 enable_circular_board=bool(enable_circular_board), enable_start_blocking=bool(enable_start_blocking),
                 enable_jump_in_goal_area=bool(enable_jump_in_goal_area), enable_friendly_fire=bool(enable_friendly_fire),
                 must_traverse_start=bool(must_traverse_start), disable_swapping=bool(disable_swapping),
                 disable_hot_seven=bool(disable_hot_seven), disable_joker=bool(disable_joker))
    static = dict(num_players=num_players, num_cards=14, board_size=4 * distance, total_board_size=4 * distance + 16,
                  rules=rules, layout_mask=lm, _start=start, _target=target, _goal=goal)
    env = reuse_or_alloc(DOG, out, int(seeds.numel()), static, device, batched)  # out=: overwrite that env's leaves
    cfg, st = env.cfg(), env.cstate()
    _lib.check(_lib.lib().dogstep_dog_reset(C.byref(st), C.c_int64(env.n), C.byref(cfg), _lib.ptr(seeds),
                                           C.c_int32(int(starting_player)), _lib.stream()), "dog_reset")
    return env


def valid_actions(env):
    """valid_actions (:693-711) -> bool [n, 806]"""
    na = get_play_action_size(env) + 14
    mask = torch.empty((env.n, na), dtype=torch.uint8, device=env.device)
    cfg, st = env.cfg(), env.cstate()
    _lib.check(_lib.lib().dogstep_dog_valid_actions(C.byref(st), C.c_int64(env.n), C.byref(cfg), _lib.ptr(mask), _lib.stream()),
               "dog_valid_actions")
    return _out(env, mask.view(torch.bool))


def env_step(env, action, inplace=False):
    """env_step (:1117-1131) -> (env, reward, done)"""
    if not inplace:
        env = env.clone()
    act = to_dev(action, torch.int32, env.device).reshape(env.n)
    reward = torch.empty(env.n, dtype=torch.int8, device=env.device)
    done = torch.empty(env.n, dtype=torch.bool, device=env.device)
    cfg, st = env.cfg(), env.cstate()
    _lib.check(_lib.lib().dogstep_dog_step(C.byref(st), C.c_int64(env.n), C.byref(cfg), _lib.ptr(act), _lib.ptr(reward),
                                          _lib.ptr(done), _lib.stream()), "dog_step")
    return env, _out(env, reward), _out(env, done)


def no_step(env, inplace=False):
    """no_step (:713-752)"""
    if not inplace:
        env = env.clone()
    reward = torch.empty(env.n, dtype=torch.int8, device=env.device)
    done = torch.empty(env.n, dtype=torch.bool, device=env.device)
    cfg, st = env.cfg(), env.cstate()
    _lib.check(_lib.lib().dogstep_dog_no_step(C.byref(st), C.c_int64(env.n), C.byref(cfg), _lib.ptr(reward), _lib.ptr(done),
                                             _lib.stream()), "dog_no_step")
    return env, _out(env, reward), _out(env, done)


def distribute_cards(env, inplace=False):
    """distribute_cards (:201-298)"""
    if not inplace:
        env = env.clone()
    cfg, st = env.cfg(), env.cstate()
    _lib.check(_lib.lib().dogstep_dog_distribute_cards(C.byref(st), C.c_int64(env.n), C.byref(cfg), _lib.stream()),
               "dog_distribute_cards")
    return env


def set_pins_on_board(board, pins):
    """set_pins_on_board (:346-358); pins may be int32 (DOG) — positions fit int8."""
    from ..MADN.deterministic_madn import set_pins_on_board as _spb
    dev = board.device if isinstance(board, torch.Tensor) and board.is_cuda else torch.device("cuda")
    return _spb(to_dev(board, torch.int8, dev), to_dev(pins, torch.int32, dev).to(torch.int8))


def _substep(env, kind, args):
    n = env.n
    k = torch.full((n,), kind, dtype=torch.int32, device=env.device)
    a = to_dev(args, torch.int32, env.device).reshape(n, 4)
    board = torch.empty_like(env.raw("board"))
    pins = torch.empty_like(env.raw("pins"))
    reward = torch.empty(n, dtype=torch.int8, device=env.device)
    done = torch.empty(n, dtype=torch.bool, device=env.device)
    cfg, st = env.cfg(), env.cstate()
    _lib.check(_lib.lib().dogstep_dog_substep(C.byref(st), C.c_int64(n), C.byref(cfg), _lib.ptr(k), _lib.ptr(a), _lib.ptr(board),
                                             _lib.ptr(pins), _lib.ptr(reward), _lib.ptr(done), _lib.stream()), "dog_substep")
    return _out(env, board), _out(env, pins), _out(env, reward), _out(env, done)


def _pad4(*cols):
    t = [torch.as_tensor(np.asarray(c.cpu() if isinstance(c, torch.Tensor) else c)).reshape(-1).to(torch.int32) for c in cols]
    z = torch.zeros_like(t[0])
    return torch.stack(t + [z] * (4 - len(t)), dim=-1)


def step_normal_move(env, pin, move):
    """step_normal_move (:790-858) -> (board, pins, reward, done)"""
    return _substep(env, 0, _pad4(pin, move))


def step_neg_move(env, pin, move):
    """step_neg_move (:861-910)"""
    return _substep(env, 1, _pad4(pin, move))


def step_swap(env, pin_idx, swap_pos):
    """step_swap (:755-787)"""
    return _substep(env, 2, _pad4(pin_idx, swap_pos))


def step_hot_7(env, seven_dist):
    """step_hot_7 (:913-984)"""
    return _substep(env, 3, seven_dist)


# ---- pure index maps (host; a handful of integer ops) ----------------------------------------------------
def map_action_to_move(env, action):
    """map_action_to_move (:1134-1196) -> int32 [..., 6] = [is_joker, is_swap, d0, d1, d2, d3]"""
    total = env.static["total_board_size"]
    space = get_play_action_size(env)
    half, pxb = space // 2, 4 * total
    a = np.asarray(action.cpu() if isinstance(action, torch.Tensor) else action).astype(np.int64)
    shape = a.shape
    a = a.reshape(-1)
    out = np.zeros((a.size, 6), np.int32)
    for j, act_full in enumerate(a):
        is_joker = (act_full - half) < 0
        act = int(act_full % half)
        is_swap = act < pxb
        d = np.zeros(4, np.int32)
        if is_swap:
            d[:] = -1
            d[act // total] = act % total
        elif act < pxb + 120:
            d[:] = DISTS_7_4[act - pxb]
        elif act < half - 4:
            na = act - (pxb + 120)
            move = na % 12 + 1
            move += int(move >= 7)
            d[na // 12] = move
        else:
            d[act - (half - 4)] = -4
        out[j] = [int(is_joker), int(is_swap), *d]
    return torch.as_tensor(out.reshape(shape + (6,)))


def map_move_to_action(env, move):
    """map_move_to_action (:1198-1239) -> int32"""
    total = env.static["total_board_size"]
    space = get_play_action_size(env)
    half, pxb = space // 2, 4 * total
    m = np.asarray(move.cpu() if isinstance(move, torch.Tensor) else move).astype(np.int64).reshape(-1, 6)
    out = np.zeros(m.shape[0], np.int32)
    for j, mv in enumerate(m):
        d = mv[2:]
        if mv[1] == 1:
            pi = int(np.argmax(d >= 0))
            idx = pi * total + d[pi]
        elif d.sum() == 7:
            idx = pxb + int(np.argmax((DISTS_7_4 == d[None, :]).all(1)))
        elif (d == -4).any():
            idx = (half - 4) + int(np.argmax(d == -4))
        else:
            pi = int(np.argmax(d != 0))
            mvv = d[pi]
            idx = pxb + 120 + pi * 12 + (mvv - 1 - int(mvv > 7))
        out[j] = idx if mv[0] == 1 else idx + half
    return torch.as_tensor(out if out.size > 1 else out[0])


def map_action_to_card(action):
    """map_action_to_card (:1241-1262): the 6-vector of map_action_to_move -> card id"""
    mv = np.asarray(action.cpu() if isinstance(action, torch.Tensor) else action).astype(np.int64).reshape(-1, 6)
    s = mv[:, 2:].sum(1)
    card = np.where(mv[:, 0] == 1, 0, np.where(mv[:, 1] == 1, 1, np.where(s == -4, 4, np.where(s == 1, 11, s))))
    return torch.as_tensor(card.astype(np.int32) if card.size > 1 else np.int32(card[0]))


def random_step(env, rng_key, game_offset=0, active_count=None):
    """one fused lockstep iteration of the random-legal-policy driver over the 806 actions, in place"""
    ac = None if active_count is None else active_count.data_ptr()
    prep = env.memo(("random_step", game_offset, ac), lambda: _lib.Prepared(
        _lib.lib().dogstep_dog_random_step,
        [C.byref(env.cstate()), C.c_int64(env.n), C.byref(env.cfg()), None, C.c_int64(game_offset), _lib.ptr(active_count), _lib.stream()]))
    prep.args[3] = _lib.host_key(rng_key)
    rc = prep()
    if rc:
        _lib.check(rc, "dog_random_step")
    return env


def play_random(env, rng_key, max_steps=2000, game_offset=0, game_len=None, total_steps=None):
    """the whole random-policy loop as one persistent kernel, in place -> (env, game_len)"""
    if game_len is None:
        game_len = torch.empty(env.n, dtype=torch.int32, device=env.device)
    cfg, st = env.cfg(), env.cstate()
    _lib.check(_lib.lib().dogstep_dog_play_random(C.byref(st), C.c_int64(env.n), C.byref(cfg), _lib.host_key(rng_key),
                                                 C.c_int64(game_offset), C.c_int32(max_steps), _lib.ptr(game_len),
                                                 _lib.ptr(total_steps), _lib.stream()), "dog_play_random")
    return env, game_len
