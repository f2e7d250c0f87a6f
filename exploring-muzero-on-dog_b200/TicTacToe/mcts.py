"""Host mirror of TicTacToe/mcts.py: run_mcts :10-23 (mctx.muzero_policy, max_depth 9, qtransform_by_min_max(-1, 1),
dirichlet_fraction 0) and run_gumbel :25-38, on the true env with rollout values — batched over games (the reference runs
batch_size = 1), plus the lockstep self-play driver of BASELINE config 1."""
import ctypes as C
import functools

import torch

from .. import _lib, jaxrand
from .. import mcts as _mcts
from . import TicTacToeV2 as game


def _split_each(keys, index):
    out = torch.empty_like(keys)
    _lib.check(_lib.lib().dogstep_random_split_each(_lib.ptr(keys), C.c_int64(keys.shape[0]), C.c_uint32(index), _lib.ptr(out),
                                                   _lib.stream()), "random_split_each")
    return out


def _root(env, rng_key):
    key1, key2 = _split_each(rng_key, 0), _split_each(rng_key, 1)    # key1, key2 = split(rng_key)
    return key1, game.root_fn(env, _split_each(key2, 0))              # root_fn(env, split(key2, batch_size=1)[0])


_QT = functools.partial(_mcts.qtransform_by_min_max, min_value=-1, max_value=1)


def run_mcts(rng_key, env, num_simulations, graph_cache=None):
    """run_mcts (:10-23); rng_key uint32 [games, 2].  graph_cache: optional mcts.GraphCache — the 3 x num_simulations
    launches of one move are then replayed as one CUDA graph (the loop is launch-bound at 512 games)."""
    key1, root = _root(env, rng_key)
    return _mcts.muzero_policy(None, key1, root, game.make_recurrent_fn(env.static["variant"]), num_simulations, max_depth=9,
                               qtransform=_QT, dirichlet_fraction=0.0, graph_cache=graph_cache)


def run_gumbel(rng_key, env, num_simulations, graph_cache=None):
    """run_gumbel (:25-38)"""
    key1, root = _root(env, rng_key)
    return _mcts.gumbel_muzero_policy(None, key1, root, game.make_recurrent_fn(env.static["variant"]), num_simulations, max_depth=9,
                                      qtransform=_QT, graph_cache=graph_cache)


def play_mcts_games(n, rng_key, num_simulations=50, limit=30, variant=1, device="cuda", search=run_mcts, graph_cache=None):
    """BASELINE config 1: n lockstep games, both sides pick run_mcts(...).action every ply, until all are done or `limit`
    plies (TicTacToe/eval.py:97-125 with get_mcts_action on both seats).  Returns (env, plies played per game)."""
    env = game.env_reset(0, n=n, device=device, variant=variant)
    plies = torch.zeros(n, dtype=torch.int32, device=device)
    key = rng_key
    for step in range(limit):
        live = ~env.raw("done")
        if not bool(live.any()):
            break
        key, sub = jaxrand.split_host(key)                            # rng_key, action_key = split(rng_key)
        out = search(jaxrand.split(sub, n, device=device), env, num_simulations, **({} if graph_cache is None else {"graph_cache": graph_cache}))
        stepped, _, _ = game.env_step(env, out.action.to(torch.int8))
        merged = {k: torch.where(live.reshape((-1,) + (1,) * (stepped.raw(k).ndim - 1)), stepped.raw(k), env.raw(k))
                  for k in ("board", "current_player", "reward", "done", "memory")}
        env = env.replace(**merged)
        plies += live.to(torch.int32)
    return env, plies
