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


class FusedSearch:
    """The whole true-env search of one move as ONE launch (dogstep_ttt_search): root_fn, init, num_simulations x (select,
    recurrent_fn with its rollout, expand), policy output — one game per warp, no synchronisation between games.  Driven
    simulation by simulation, every simulation waits for the longest rollout of the batch (86 % of config 1); bit-identical to
    that path.  Buffers (tree + per-game work rows) are allocated once and reused move after move."""

    def __init__(self, policy, n, num_simulations, device):
        cfg = _mcts._cfg(policy, _QT, num_simulations, 9, 9, 0, 18, dirichlet_fraction=0.0)
        self.search = _mcts.Search(cfg, n, device)
        f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=device)
        self.rows = dict(prior_logits=f(n, 9), value=f(n), reward=f(n), discount=f(n), next_embedding=f(n, 18), root_prior_logits=f(n, 9),
                         root_value=f(n), root_embedding=f(n, 18))
        s = self.search
        self.scratch = _lib.tag(_lib.TttSearchScratch(*[C.c_void_p(t.data_ptr()) for t in (
            s.parent, s.action, s.embedding, s.expand_key, self.rows["prior_logits"], self.rows["value"], self.rows["reward"],
            self.rows["discount"], self.rows["next_embedding"], self.rows["root_prior_logits"], self.rows["root_value"],
            self.rows["root_embedding"])]), s.parent.device)
        self.action = torch.empty(n, dtype=torch.int32, device=device)
        self.weights = torch.empty((n, 9), dtype=torch.float32, device=device)
        self.value = torch.empty(n, dtype=torch.float32, device=device)

    def __call__(self, env, rng_key):
        key1, key2 = _split_each(rng_key, 0), _split_each(rng_key, 1)   # key1, key2 = split(rng_key)              (mcts.py:12)
        root_keys = _split_each(key2, 0)                                 # split(key2, batch_size = 1)[0]           (:15)
        s, st = self.search, env.cstate()
        _lib.check(_lib.lib().dogstep_ttt_search(C.byref(st), C.c_int64(env.n), C.c_int32(env.static["variant"]), C.byref(s._ct),
                                                C.byref(s.cfg), _lib.ptr(key1), _lib.ptr(root_keys), C.byref(self.scratch),
                                                _lib.ptr(self.action), _lib.ptr(self.weights), _lib.ptr(self.value), _lib.stream()),
                   "ttt_search")
        return _mcts.PolicyOutput(action=self.action, action_weights=self.weights, search_tree=s.tree)


def _fused(cache, policy, env, num_simulations):
    key = (policy, env.n, num_simulations, str(env.device))
    fs = cache.get(key)
    if fs is None:
        fs = cache[key] = FusedSearch(policy, env.n, num_simulations, env.device)
    return fs


def run_mcts(rng_key, env, num_simulations, graph_cache=None, fused=None):
    """run_mcts (:10-23); rng_key uint32 [games, 2].  fused: a dict the caller keeps (one launch per move, see FusedSearch);
    graph_cache: optional mcts.GraphCache — the 3 x num_simulations launches of one move replayed as one CUDA graph."""
    if fused is not None:
        return _fused(fused, _mcts.MUZERO, env, num_simulations)(env, rng_key)
    key1, root = _root(env, rng_key)
    return _mcts.muzero_policy(None, key1, root, game.make_recurrent_fn(env.static["variant"]), num_simulations, max_depth=9,
                               qtransform=_QT, dirichlet_fraction=0.0, graph_cache=graph_cache)


def run_gumbel(rng_key, env, num_simulations, graph_cache=None, fused=None):
    """run_gumbel (:25-38)"""
    if fused is not None:
        return _fused(fused, _mcts.GUMBEL, env, num_simulations)(env, rng_key)
    key1, root = _root(env, rng_key)
    return _mcts.gumbel_muzero_policy(None, key1, root, game.make_recurrent_fn(env.static["variant"]), num_simulations, max_depth=9,
                                      qtransform=_QT, graph_cache=graph_cache)


def play_mcts_games(n, rng_key, num_simulations=50, limit=30, variant=1, device="cuda", search=run_mcts, graph_cache=None, fused=None,
                    move="weights"):
    """BASELINE config 1: n lockstep games, both seats searched every ply, until all are done or `limit` plies — the
    reference's play_mcts_match (TicTacToe/eval.py:151-176).  move = "weights": the reference's get_mcts_action (:28-34),
    the largest action weight among the empty cells; "sample": PolicyOutput.action.  Returns (env, plies played per game)."""
    env = game.env_reset(0, n=n, device=device, variant=variant)
    plies = torch.zeros(n, dtype=torch.int32, device=device)
    key = rng_key
    for step in range(limit):
        # the termination test is a host round trip: every fourth ply (a ply on a finished batch changes nothing — finished
        # games are not stepped and not counted)
        if step % 4 == 0 and bool(env.raw("done").all()):
            break
        key, sub = jaxrand.split_host(key)                            # rng_key, action_key = split(rng_key)
        kw = {"fused": fused} if fused is not None else ({} if graph_cache is None else {"graph_cache": graph_cache})
        out = search(jaxrand.split(sub, n, device=device), env, num_simulations, **kw)
        if move == "weights":
            game.play_move(env, out.action_weights, plies)           # one launch: pick + env_step of the live games, in place
            continue
        live = ~env.raw("done")
        stepped, _, _ = game.env_step(env, out.action.to(torch.int8))
        merged = {k: torch.where(live.reshape((-1,) + (1,) * (stepped.raw(k).ndim - 1)), stepped.raw(k), env.raw(k))
                  for k in ("board", "current_player", "reward", "done", "memory")}
        env = env.replace(**merged)
        plies += live.to(torch.int32)
    return env, plies


def play_match(n, rng_key, num_simulations, bot_player=1, limit=30, variant=1, device="cuda", search=run_mcts, fused=None):
    """n lockstep games of the search bot (seat `bot_player`, +1 moves first) against the random bot — the reference's test /
    play_match with a search player (TicTacToe/eval.py:97-125, :252-275; get_random_action :51-55 = categorical over the
    empty cells with the ply's action key).  Returns (env, result per game: +1 the bot won, -1 it lost, 0 draw or limit)."""
    env = game.env_reset(0, n=n, device=device, variant=variant)
    key = rng_key
    for step in range(limit):
        if step % 4 == 0 and bool(env.raw("done").all()):
            break
        key, sub = jaxrand.split_host(key)
        if (1 if step % 2 == 0 else -1) == bot_player:                # seats alternate every ply in both variants
            weights = search(jaxrand.split(sub, n, device=device), env, num_simulations,
                             **({"fused": fused} if fused is not None else {})).action_weights
        else:
            # categorical(key, where(empty, 0, -inf)) over the batch: the empty cell with the largest Gumbel draw, i.e. with
            # the largest of the uniform draws behind it
            weights = jaxrand.uniform(sub, 9 * n, device=device).reshape(n, 9)
        game.play_move(env, weights)
    board = env.raw("board").reshape(n, 9).to(torch.int32)
    lines = torch.tensor([[0, 1, 2], [3, 4, 5], [6, 7, 8], [0, 3, 6], [1, 4, 7], [2, 5, 8], [0, 4, 8], [2, 4, 6]], device=board.device)
    sums = board[:, lines].sum(2)
    winner = (sums == 3).any(1).to(torch.int32) - (sums == -3).any(1).to(torch.int32)   # get_winner
    return env, torch.where(env.raw("done"), winner * bot_player, torch.zeros_like(winner))
