"""Host mirror of TicTacToe/TicTacToeV2.py (and, with VARIANT = 0, TicTacToe/TicTacToe.py): env_reset :38, env_step :46,
valid_action_mask :81, policy_function :96, root_fn :121, recurrent_fn :128 — batched over games, one libdogstep.so kernel
per call.  The env doubles as the search embedding (18 floats), exactly as the reference stores the env pytree in the tree."""
import ctypes as C
import functools

import numpy as np
import torch

from .. import _lib, mcts
from ..MADN._state import BatchedEnv, to_dev

VARIANT = 1


class TicTacToeV2(BatchedEnv):
    LEAVES = {
        "board": (torch.int8, lambda s: (3, 3)),
        "current_player": (torch.int8, lambda s: ()),
        "reward": (torch.int8, lambda s: ()),
        "done": (torch.bool, lambda s: ()),
        "memory": (torch.int8, lambda s: (2, 3)),
    }

    def cstate(self):
        t = self._t
        return _lib.tag(_lib.TttState(*[C.c_void_p(t[k].data_ptr()) for k in ("board", "current_player", "reward", "done", "memory")]),
                        t["board"].device)


def env_reset(_, n=None, device="cuda", variant=None):
    """env_reset (:38-45); n=None gives the reference's single env, n=k a batch of k games"""
    env = TicTacToeV2(1 if n is None else n, dict(variant=VARIANT if variant is None else variant), torch.device(device), n is not None)
    env.alloc()
    st = env.cstate()
    _lib.check(_lib.lib().dogstep_ttt_reset(C.byref(st), C.c_int64(env.n), _lib.stream()), "ttt_reset")
    return env


def _v(env):
    return C.c_int32(env.static["variant"])


def env_step(env, action, inplace=False):
    """env_step (:46-79) -> (env, reward, done)"""
    if not inplace:
        env = env.clone()
    act = to_dev(action, torch.int8, env.device).reshape(-1)
    act = act.expand(env.n).contiguous() if act.numel() == 1 else act
    reward = torch.empty(env.n, dtype=torch.int8, device=env.device)
    done = torch.empty(env.n, dtype=torch.bool, device=env.device)
    st = env.cstate()
    _lib.check(_lib.lib().dogstep_ttt_step(C.byref(st), C.c_int64(env.n), _v(env), _lib.ptr(act), _lib.ptr(reward), _lib.ptr(done),
                                          _lib.stream()), "ttt_step")
    return env, (reward if env.batched else reward[0]), (done if env.batched else done[0])


def play_move(env, action_weights, plies=None):
    """One ply of the reference's match loops, in place (TicTacToe/eval.py:97-125, :151-176): every game that is not done plays
    get_mcts_action (:28-34) = argmax(where(board == 0, action_weights, -inf)) through env_step; finished games are left
    untouched.  Returns the moves (int8 [n], -1 for a finished game); `plies` (int32 [n]) counts the games that moved."""
    w = to_dev(action_weights, torch.float32, env.device).reshape(env.n, 9)
    action = torch.empty(env.n, dtype=torch.int8, device=env.device)
    st = env.cstate()
    _lib.check(_lib.lib().dogstep_ttt_play_move(C.byref(st), C.c_int64(env.n), _v(env), _lib.ptr(w), _lib.ptr(action),
                                               None if plies is None else _lib.ptr(plies), _lib.stream()), "ttt_play_move")
    return action


def valid_action_mask(env):
    """valid_action_mask (:81-82) -> bool [n,3,3]"""
    m = torch.empty((env.n, 9), dtype=torch.uint8, device=env.device)
    st = env.cstate()
    _lib.check(_lib.lib().dogstep_ttt_policy_function(C.byref(st), C.c_int64(env.n), _v(env), None, _lib.ptr(m), _lib.stream()), "ttt_valid")
    m = m.view(torch.bool).reshape(env.n, 3, 3)
    return m if env.batched else m[0]


def policy_function(env):
    """policy_function (:96-102) -> f32 [n,9]"""
    lg = torch.empty((env.n, 9), dtype=torch.float32, device=env.device)
    st = env.cstate()
    _lib.check(_lib.lib().dogstep_ttt_policy_function(C.byref(st), C.c_int64(env.n), _v(env), _lib.ptr(lg), None, _lib.stream()), "ttt_policy")
    return lg if env.batched else lg[0]


def root_fn(env, rng_key):
    """root_fn (:121-126); rng_key uint32 [n,2]"""
    n, dev = env.n, env.device
    prior = torch.empty((n, 9), dtype=torch.float32, device=dev)
    value = torch.empty(n, dtype=torch.float32, device=dev)
    emb = torch.empty((n, 18), dtype=torch.float32, device=dev)
    st = env.cstate()
    _lib.check(_lib.lib().dogstep_ttt_root_fn(C.byref(st), C.c_int64(n), _v(env), _lib.ptr(rng_key.contiguous()), _lib.ptr(prior),
                                             _lib.ptr(value), _lib.ptr(emb), _lib.stream()), "ttt_root_fn")
    return mcts.RootFnOutput(prior, value, emb)


@functools.lru_cache(maxsize=None)
def make_recurrent_fn(variant=VARIANT):
    def recurrent_fn(params, rng_key, action, embedding):
        """recurrent_fn (:128-140) on the 18-float env embedding"""
        n, dev = embedding.shape[0], embedding.device
        prior = torch.empty((n, 9), dtype=torch.float32, device=dev)
        value, reward, discount = (torch.empty(n, dtype=torch.float32, device=dev) for _ in range(3))
        nxt = torch.empty((n, 18), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().dogstep_ttt_recurrent_fn(C.c_int64(n), C.c_int32(variant), _lib.ptr(rng_key.contiguous()),
                                                      _lib.ptr(action.to(torch.int32).contiguous()), _lib.ptr(embedding.contiguous()),
                                                      _lib.ptr(prior), _lib.ptr(value), _lib.ptr(reward), _lib.ptr(discount),
                                                      _lib.ptr(nxt), _lib.stream()), "ttt_recurrent_fn")
        return mcts.RecurrentFnOutput(reward, discount, prior, value), nxt
    return recurrent_fn


recurrent_fn = make_recurrent_fn(VARIANT)
