"""Host mirror of the reference's MADN/classic_madn.py (dice MADN) for the self-play hot path.

Function names / arguments / return tuples follow /root/reference/MADN/classic_madn.py:
env_reset :51, dice_probabilities :208, throw_die :230, set_die :244, env_step :257,
set_pins_on_board :339, no_step :353, valid_action :367, encode_board :463 — each already
vmapped over the leading game axis and executed by one libdogstep.so kernel.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib, rules as _rules
from ._state import BatchedEnv, reuse_or_alloc, seeds_to_dev, to_dev
from .deterministic_madn import _geometry, _layout_mask, _out, set_pins_on_board  # noqa: F401  (same function in both files)


class classic_MADN(BatchedEnv):
    """Batched leaves of the reference dataclass (:33-49)."""

    LEAVES = {
        "board": (torch.int8, lambda s: (s["total_board_size"],)),
        "current_player": (torch.int8, lambda s: ()),
        "pins": (torch.int8, lambda s: (s["num_players"], 4)),
        "reward": (torch.int8, lambda s: ()),
        "done": (torch.bool, lambda s: ()),
        "die": (torch.int8, lambda s: ()),
        "key": (torch.uint32, lambda s: (2,)),
    }

    def _const(self, name):
        a = torch.as_tensor(self.static["_" + name], device=self.device)
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
        return _lib.tag(_lib.MadnClsState(*[C.c_void_p(t[k].data_ptr()) for k in
                                            ("board", "current_player", "pins", "reward", "done", "die", "key")]), t["board"].device)


def env_reset(_, num_players=4, layout=None, distance=10, starting_player=0, seed=42, enable_teams=False,
              enable_initial_free_pin=False, enable_circular_board=True, enable_start_blocking=False,
              enable_jump_in_goal_area=True, enable_friendly_fire=False, enable_start_on_1=True,
              enable_bonus_turn_on_6=True, enable_dice_rethrow=False, must_traverse_start=False, device="cuda", out=None):
    """env_reset (:51-131); `seed` scalar (single env) or int array [n] (vmapped)."""
    num_players, distance = int(num_players), int(distance)
    batched, seeds = seeds_to_dev(seed, device)
    lm = _layout_mask(layout)
    start, target, goal = _geometry(num_players, lm, distance)
    rules = dict(enable_teams=bool(enable_teams) and num_players == 4,
                 enable_initial_free_pin=bool(enable_initial_free_pin),
                 enable_circular_board=bool(enable_circular_board), enable_start_blocking=bool(enable_start_blocking),
                 enable_jump_in_goal_area=bool(enable_jump_in_goal_area), enable_friendly_fire=bool(enable_friendly_fire),
                 enable_start_on_1=bool(enable_start_on_1), enable_bonus_turn_on_6=bool(enable_bonus_turn_on_6),
                 enable_dice_rethrow=bool(enable_dice_rethrow), must_traverse_start=bool(must_traverse_start))
    static = dict(num_players=num_players, board_size=4 * distance, total_board_size=4 * distance + 16, rules=rules,
                  layout_mask=lm, _start=start, _target=target, _goal=goal)
    env = reuse_or_alloc(classic_MADN, out, int(seeds.numel()), static, device, batched)  # out=: overwrite that env's leaves
    cfg, st = env.cfg(), env.cstate()
    _lib.check(_lib.lib().dogstep_madn_cls_reset(C.byref(st), C.c_int64(env.n), C.byref(cfg), _lib.ptr(seeds),
                                                C.c_int32(int(starting_player)), _lib.stream()), "madn_cls_reset")
    return env


def dice_probabilities(env):
    """dice_probabilities (:208-228) -> float32 [n, 6]"""
    p = torch.empty((env.n, 6), dtype=torch.float32, device=env.device)
    cfg, st = env.cfg(), env.cstate()
    _lib.check(_lib.lib().dogstep_madn_cls_dice_probabilities(C.byref(st), C.c_int64(env.n), C.byref(cfg), _lib.ptr(p),
                                                             _lib.stream()), "madn_cls_dice_probabilities")
    return _out(env, p)


def throw_die(env, inplace=False):
    """throw_die (:230-242): key, sub = split(env.key); die = choice(sub, 1..6, p=dice_probabilities)."""
    if not inplace:
        env = env.clone()
    cfg, st = env.cfg(), env.cstate()
    _lib.check(_lib.lib().dogstep_madn_cls_throw_die(C.byref(st), C.c_int64(env.n), C.byref(cfg), _lib.stream()),
               "madn_cls_throw_die")
    return env


def set_die(env, die_value):
    """set_die (:244-255)"""
    v = to_dev(die_value, torch.int8, env.device)
    return env.replace(die=v.expand(env.n) if v.numel() == 1 else v)


def valid_action(env):
    """valid_action (:367-461) -> bool [n, 4]"""
    mask = torch.empty((env.n, 4), dtype=torch.uint8, device=env.device)
    cfg, st = env.cfg(), env.cstate()
    _lib.check(_lib.lib().dogstep_madn_cls_valid_action(C.byref(st), C.c_int64(env.n), C.byref(cfg), _lib.ptr(mask),
                                                       _lib.stream()), "madn_cls_valid_action")
    return _out(env, mask.view(torch.bool))


def env_step(env, pin, inplace=False):
    """env_step (:257-337): move pin `pin` by env.die -> (env, reward, done)"""
    if not inplace:
        env = env.clone()
    act = to_dev(pin, torch.int8, env.device).reshape(env.n)
    reward = torch.empty(env.n, dtype=torch.int8, device=env.device)
    done = torch.empty(env.n, dtype=torch.bool, device=env.device)
    cfg, st = env.cfg(), env.cstate()
    _lib.check(_lib.lib().dogstep_madn_cls_step(C.byref(st), C.c_int64(env.n), C.byref(cfg), _lib.ptr(act), _lib.ptr(reward),
                                               _lib.ptr(done), _lib.stream()), "madn_cls_step")
    return env, _out(env, reward), _out(env, done)


def no_step(env, inplace=False):
    """no_step (:353-365)"""
    if not inplace:
        env = env.clone()
    reward = torch.empty(env.n, dtype=torch.int8, device=env.device)
    done = torch.empty(env.n, dtype=torch.bool, device=env.device)
    cfg, st = env.cfg(), env.cstate()
    _lib.check(_lib.lib().dogstep_madn_cls_no_step(C.byref(st), C.c_int64(env.n), C.byref(cfg), _lib.ptr(reward),
                                                  _lib.ptr(done), _lib.stream()), "madn_cls_no_step")
    return env, _out(env, reward), _out(env, done)


def encode_board(env, dtype=torch.int8):
    """encode_board (:463-497) -> [n, 2*P+3, total]"""
    P, T = env.static["num_players"], env.static["total_board_size"]
    obs = torch.empty((env.n, 2 * P + 3, T), dtype=torch.int8, device=env.device)
    cfg, st = env.cfg(), env.cstate()
    _lib.check(_lib.lib().dogstep_madn_cls_encode_board(C.byref(st), C.c_int64(env.n), C.byref(cfg), _lib.ptr(obs),
                                                       _lib.stream()), "madn_cls_encode_board")
    return _out(env, obs if dtype == torch.int8 else obs.to(dtype))
