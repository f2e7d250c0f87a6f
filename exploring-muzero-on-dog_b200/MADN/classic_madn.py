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


# ---- true-env mctx callbacks (:541-714) and the search that uses them (MADN/simulate_classicMADN.py:51-76) ------
# The reference as it stands raises before any of these returns (winning_action builds its scratch copy without the `key`
# field, :551-565); see include/dogstep.h and tests/golden/gen_madn_cls_trueenv_goldens.py for what is computed and how it is pinned.
def embed_dim(env):
    """floats of the env embedding: board, current_player, pins, reward, done, die, key as four 16-bit halves"""
    return int(_lib.lib().dogstep_madn_cls_embed_dim(C.byref(env.cfg())))


def policy_function(env):
    """policy_function (:571-583) -> f32 [n, 4]: 100 * valid_action + 200 * winning_action (:543-569)"""
    lg = torch.empty((env.n, 4), dtype=torch.float32, device=env.device)
    cfg, st = env.cfg(), env.cstate()
    _lib.check(_lib.lib().dogstep_madn_cls_policy_function(C.byref(st), C.c_int64(env.n), C.byref(cfg), _lib.ptr(lg), _lib.stream()),
               "madn_cls_policy_function")
    return _out(env, lg)


def root_fn(env, rng_key):
    """root_fn (:690-714); rng_key uint32 [n, 2]; value = rollout(env, key) (:585-616) as a scalar per game"""
    from .. import mcts
    n, dev = env.n, env.device
    prior = torch.empty((n, 4), dtype=torch.float32, device=dev)
    value = torch.empty(n, dtype=torch.float32, device=dev)
    emb = torch.empty((n, embed_dim(env)), dtype=torch.float32, device=dev)
    cfg, st = env.cfg(), env.cstate()
    _lib.check(_lib.lib().dogstep_madn_cls_root_fn(C.byref(st), C.c_int64(n), C.byref(cfg), _lib.ptr(rng_key.contiguous()), _lib.ptr(prior),
                                                  _lib.ptr(value), _lib.ptr(emb), _lib.stream()), "madn_cls_root_fn")
    return mcts.RootFnOutput(prior, value, emb)


def value_function(env, rng_key):
    """value_function (:618-622) = rollout(env, rng_key)"""
    return root_fn(env, rng_key).value


rollout = value_function


def _keys_or_zeros(rng_key, n, dev):  # mcts.stochastic_muzero_policy probes the callbacks' shapes with rng_key = None
    return torch.zeros((n, 2), dtype=torch.uint32, device=dev) if rng_key is None else rng_key.contiguous()


def make_recurrent_fn(env):
    """recurrent_fn (:657-688), the decision node: (chance_logits = log(1/6) x 6, afterstate_value), afterstate embedding"""
    from .. import mcts
    cfg = env.cfg()

    def recurrent_fn(params, rng_key, action, embedding):
        n, dev = embedding.shape[0], embedding.device
        chance = torch.empty((n, 6), dtype=torch.float32, device=dev)
        value = torch.empty(n, dtype=torch.float32, device=dev)
        nxt = torch.empty_like(embedding)
        _lib.check(_lib.lib().dogstep_madn_cls_decision_recurrent_fn(
            C.c_int64(n), C.byref(cfg), _lib.ptr(_keys_or_zeros(rng_key, n, dev)), _lib.ptr(action.to(torch.int32).contiguous()),
            _lib.ptr(embedding.contiguous()), _lib.ptr(chance), _lib.ptr(value), _lib.ptr(nxt), _lib.stream()), "madn_cls_decision_recurrent_fn")
        return mcts.DecisionRecurrentFnOutput(chance, value), nxt
    return recurrent_fn


def make_recurrent_chance_fn(env):
    """recurrent_chance_fn (:624-655): (action_logits = valid_action as 0 / 1, value, reward, discount), state embedding"""
    from .. import mcts
    cfg = env.cfg()

    def recurrent_chance_fn(params, rng_key, chance_outcome, afterstate):
        n, dev = afterstate.shape[0], afterstate.device
        logits = torch.empty((n, 4), dtype=torch.float32, device=dev)
        value, reward, discount = (torch.empty(n, dtype=torch.float32, device=dev) for _ in range(3))
        nxt = torch.empty_like(afterstate)
        _lib.check(_lib.lib().dogstep_madn_cls_chance_recurrent_fn(
            C.c_int64(n), C.byref(cfg), _lib.ptr(_keys_or_zeros(rng_key, n, dev)), _lib.ptr(chance_outcome.to(torch.int32).contiguous()),
            _lib.ptr(afterstate.contiguous()), _lib.ptr(logits), _lib.ptr(value), _lib.ptr(reward), _lib.ptr(discount), _lib.ptr(nxt),
            _lib.stream()), "madn_cls_chance_recurrent_fn")
        return mcts.ChanceRecurrentFnOutput(logits, value, reward, discount), nxt
    return recurrent_chance_fn


def run_mcts_search(env, rng_key, num_simulations=100, dirichlet_noise=None, graph_cache=None):
    """run_mcts_search (MADN/simulate_classicMADN.py:51-76): mctx.stochastic_muzero_policy on the true env with rollout values,
    invalid_actions = ~valid_action, max_depth 500, qtransform_by_min_max(-1, 1) — batched over games; rng_key uint32 [n, 2].
    dirichlet_noise: the root noise sample (mcts.stochastic_muzero_policy; jax's gamma sampler is not reproduced)"""
    import functools
    from .. import mcts
    from ..TicTacToe.mcts import _split_each
    key1, key2 = _split_each(rng_key, 0), _split_each(rng_key, 1)    # key1, key2 = split(rng_key)
    root = root_fn(env, _split_each(key2, 0))                         # root_fn(env, split(key2, batch_size = 1)[0])
    invalid = ~valid_action(env).reshape(env.n, 4)
    return mcts.stochastic_muzero_policy(None, key1, root, make_recurrent_fn(env), make_recurrent_chance_fn(env), num_simulations,
                                         invalid_actions=invalid, max_depth=500, dirichlet_noise=dirichlet_noise,
                                         qtransform=functools.partial(mcts.qtransform_by_min_max, min_value=-1, max_value=1),
                                         graph_cache=graph_cache)

