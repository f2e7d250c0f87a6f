"""Host mirror of the reference's MADN/deterministic_madn.py for the self-play hot path.

Same function names, argument order and return tuples as the reference
(/root/reference/MADN/deterministic_madn.py): env_reset :42, env_step :170, set_pins_on_board :259,
no_step :283, valid_action :299, encode_board :395, map_action :469 — but every function is
already "vmapped": it takes the batched env (leaves with a leading game axis, CUDA tensors) and runs
one libdogstep.so kernel over all games.  An env built from a scalar seed behaves like the
reference's single env.  Functions are pure by default (the input env is left intact, as in JAX);
pass `inplace=True` on the hot path to update the leaves in place (what jax.ffi does with
input_output_aliases).
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib, rules as _rules
from ._state import BatchedEnv, reuse_or_alloc, seeds_to_dev, to_dev

RULE_KEYS = ("enable_teams", "enable_initial_free_pin", "enable_circular_board", "enable_start_blocking",
             "enable_jump_in_goal_area", "enable_friendly_fire", "enable_start_on_1", "enable_bonus_turn_on_6",
             "must_traverse_start")


def _layout_mask(layout):
    if layout is None:
        return 0xF
    a = np.asarray(layout.cpu() if isinstance(layout, torch.Tensor) else layout).astype(bool).ravel()
    return int(sum(1 << i for i in range(min(4, a.size)) if a[i]))


def _geometry(num_players, layout_mask, distance):
    """start / target / goal rows exactly as env_reset builds them (:70-78)."""
    cnt = bin(layout_mask).count("1")
    if cnt != num_players or (layout_mask == 0xF and num_players < 4):
        layout_mask = (1 << num_players) - 1
    seats = [i for i in range(4) if (layout_mask >> i) & 1]
    bs = 4 * distance
    start = np.array([i * distance for i in seats], np.int8)
    target = ((start.astype(np.int32) - 1) % bs).astype(np.int8)
    goal = np.array([[bs + 4 * i + k for k in range(4)] for i in seats], np.int8)
    return start, target, goal


class deterministic_MADN(BatchedEnv):
    """Batched leaves of the reference dataclass (:24-40)."""

    LEAVES = {
        "board": (torch.int8, lambda s: (s["total_board_size"],)),
        "current_player": (torch.int8, lambda s: ()),
        "pins": (torch.int8, lambda s: (s["num_players"], 4)),
        "reward": (torch.int8, lambda s: ()),
        "done": (torch.bool, lambda s: ()),
        "action_set": (torch.int8, lambda s: (s["num_players"], 6)),
        "key": (torch.uint32, lambda s: (2,)),
    }

    # start / target / goal are per-env leaves in the reference; they are constants of the cfg here
    def _const(self, name):
        a = torch.as_tensor(self.static["_" + name], device=self.device)
        return a.expand((self.n,) + tuple(a.shape)) if self.batched else a

    @property
    def start(self):
        return self._const("start")

    @property
    def target(self):
        return self._const("target")

    @property
    def goal(self):
        return self._const("goal")

    def cfg(self):
        return self.memo("cfg", self._make_cfg)

    def _make_cfg(self):
        s = self.static
        return _lib.MadnCfg(s["num_players"], s["layout_mask"], s["board_size"] // 4, _rules.to_mask(s["rules"]))

    def cstate(self):
        return self.memo("cstate", self._make_cstate)

    def _make_cstate(self):
        t = self._t
        return _lib.tag(_lib.MadnDetState(*[C.c_void_p(t[k].data_ptr()) for k in
                                            ("board", "current_player", "pins", "reward", "done", "action_set", "key")]), t["board"].device)


def env_reset(_, num_players=4, layout=None, distance=10, starting_player=0, seed=42, enable_teams=False,
              enable_initial_free_pin=False, enable_circular_board=True, enable_start_blocking=False,
              enable_jump_in_goal_area=True, enable_friendly_fire=False, enable_start_on_1=True,
              enable_bonus_turn_on_6=True, must_traverse_start=False, device="cuda", out=None):
    """env_reset (:42-120).  `seed` may be a scalar (single env) or an int array [n] (what
    jax.vmap(env_reset_batched) receives, game_agent.py:24-44).  `out`: an env of the same size and configuration whose
    leaves are overwritten (no allocation: the steady state of a self-play iteration re-seeds the same buffers)."""
    num_players, distance = int(num_players), int(distance)
    batched, seeds = seeds_to_dev(seed, device)
    lm = _layout_mask(layout)
    start, target, goal = _geometry(num_players, lm, distance)
    rules = dict(enable_teams=bool(enable_teams) and num_players == 4,
                 enable_initial_free_pin=bool(enable_initial_free_pin),
                 enable_circular_board=bool(enable_circular_board), enable_start_blocking=bool(enable_start_blocking),
                 enable_jump_in_goal_area=bool(enable_jump_in_goal_area), enable_friendly_fire=bool(enable_friendly_fire),
                 enable_start_on_1=bool(enable_start_on_1), enable_bonus_turn_on_6=bool(enable_bonus_turn_on_6),
                 must_traverse_start=bool(must_traverse_start))
    static = dict(num_players=num_players, board_size=4 * distance, total_board_size=4 * distance + 16, rules=rules,
                  layout_mask=lm, _start=start, _target=target, _goal=goal)
    env = reuse_or_alloc(deterministic_MADN, out, int(seeds.numel()), static, device, batched)
    cfg, st = env.cfg(), env.cstate()
    _lib.check(_lib.lib().dogstep_madn_det_reset(C.byref(st), C.c_int64(env.n), C.byref(cfg), _lib.ptr(seeds),
                                                C.c_int32(int(starting_player)), _lib.stream()), "madn_det_reset")
    return env


def _out(env, t):
    return t if env.batched else t[0]


def valid_action(env):
    """valid_action (:299-393) -> bool [n, 4, 6] (or [4, 6] for a single env)."""
    mask = torch.empty((env.n, 4, 6), dtype=torch.uint8, device=env.device)
    cfg, st = env.cfg(), env.cstate()
    _lib.check(_lib.lib().dogstep_madn_det_valid_action(C.byref(st), C.c_int64(env.n), C.byref(cfg), _lib.ptr(mask),
                                                       _lib.stream()), "madn_det_valid_action")
    return _out(env, mask.view(torch.bool))


def env_step(env, action, inplace=False):
    """env_step (:170-257).  action = [pin, move] per game -> (env, reward int8, done bool)."""
    if not inplace:
        env = env.clone()
    act = to_dev(action, torch.int8, env.device).reshape(env.n, 2)
    reward = torch.empty(env.n, dtype=torch.int8, device=env.device)
    done = torch.empty(env.n, dtype=torch.bool, device=env.device)
    cfg, st = env.cfg(), env.cstate()
    _lib.check(_lib.lib().dogstep_madn_det_step(C.byref(st), C.c_int64(env.n), C.byref(cfg), _lib.ptr(act), _lib.ptr(reward),
                                               _lib.ptr(done), _lib.stream()), "madn_det_step")
    return env, _out(env, reward), _out(env, done)


def no_step(env, inplace=False):
    """no_step (:283-297) -> (env, 0, env.done)."""
    if not inplace:
        env = env.clone()
    reward = torch.empty(env.n, dtype=torch.int8, device=env.device)
    done = torch.empty(env.n, dtype=torch.bool, device=env.device)
    cfg, st = env.cfg(), env.cstate()
    _lib.check(_lib.lib().dogstep_madn_det_no_step(C.byref(st), C.c_int64(env.n), C.byref(cfg), _lib.ptr(reward),
                                                  _lib.ptr(done), _lib.stream()), "madn_det_no_step")
    return env, _out(env, reward), _out(env, done)


def set_pins_on_board(board, pins, distance=None):
    """set_pins_on_board (:259-271): board int8[..., total], pins int8[..., P, 4] -> new board."""
    dev = board.device if isinstance(board, torch.Tensor) and board.is_cuda else torch.device("cuda")
    board = to_dev(board, torch.int8, dev)
    pins = to_dev(pins, torch.int8, dev)
    total, P = board.shape[-1], pins.shape[-2]
    n = pins.numel() // (P * 4)
    out = torch.empty((n, total), dtype=torch.int8, device=dev)
    cfg = _lib.MadnCfg(P, (1 << P) - 1, (total - 16) // 4, 0)
    _lib.check(_lib.lib().dogstep_madn_set_pins_on_board(_lib.ptr(pins.reshape(n, P, 4).contiguous()), _lib.ptr(out),
                                                        C.c_int64(n), C.byref(cfg), _lib.stream()), "set_pins_on_board")
    return out.reshape(board.shape)


def encode_board(env, dtype=torch.int8):
    """encode_board (:395-438) -> [n, 8*P+2, total].  The reference's result is int32 by NumPy
    promotion; values are identical, int8 is the compact device layout (pass dtype to widen)."""
    P, T = env.static["num_players"], env.static["total_board_size"]
    obs = torch.empty((env.n, 8 * P + 2, T), dtype=torch.int8, device=env.device)
    cfg, st = env.cfg(), env.cstate()
    _lib.check(_lib.lib().dogstep_madn_det_encode_board(C.byref(st), C.c_int64(env.n), C.byref(cfg), _lib.ptr(obs),
                                                       _lib.stream()), "madn_det_encode_board")
    return _out(env, obs if dtype == torch.int8 else obs.to(dtype))


def map_action(action_index):
    """map_action (:469-479): idx -> [idx // 6, idx % 6 + 1] as int8."""
    a = action_index if isinstance(action_index, torch.Tensor) else torch.as_tensor(action_index)
    return torch.stack([(a // 6).to(torch.int8), (a % 6 + 1).to(torch.int8)], dim=-1)


def random_step(env, rng_key, game_offset=0, active_count=None):
    """One fused lockstep iteration of the random-legal-policy driver
    (MuZero_det_MADN/evaluate_agent.py:733-930, do_random), in place."""
    ac = None if active_count is None else active_count.data_ptr()
    prep = env.memo(("random_step", game_offset, ac), lambda: _lib.Prepared(
        _lib.lib().dogstep_madn_det_random_step,
        [C.byref(env.cstate()), C.c_int64(env.n), C.byref(env.cfg()), None, C.c_int64(game_offset), _lib.ptr(active_count), _lib.stream()]))
    prep.args[3] = _lib.host_key(rng_key)   # a jaxrand.KeyChain is passed as its buffer: no per-call conversion
    rc = prep()
    if rc:
        _lib.check(rc, "madn_det_random_step")
    return env


def random_steps(env, rng_key, iterations, game_offset=0, game_len=None, total_steps=None):
    """`iterations` lockstep iterations of the random-legal-policy driver in ONE launch (the state stays in registers between
    them), in place -> the loop key to continue with.  A host loop that calls this with 16-64 iterations at a time pays one
    launch + one state round trip per chunk instead of per iteration (random_step)."""
    play_random(env, rng_key, max_steps=iterations, game_offset=game_offset, game_len=game_len, total_steps=total_steps)
    from .. import jaxrand
    return jaxrand.key_chain_host(rng_key, iterations)


def play_random(env, rng_key, max_steps=2000, game_offset=0, game_len=None, total_steps=None):
    """The whole random-policy while_loop (evaluate_agent.py:733-930, cap :918) as one persistent
    kernel, in place.  Returns (env, game_len int32[n])."""
    if game_len is None:
        game_len = torch.empty(env.n, dtype=torch.int32, device=env.device)
    cfg, st = env.cfg(), env.cstate()
    _lib.check(_lib.lib().dogstep_madn_det_play_random(C.byref(st), C.c_int64(env.n), C.byref(cfg), _lib.host_key(rng_key),
                                                      C.c_int64(game_offset), C.c_int32(max_steps), _lib.ptr(game_len),
                                                      _lib.ptr(total_steps), _lib.stream()), "madn_det_play_random")
    return env, game_len


# ---- true-env mctx callbacks (:480-590) and the search that uses them (MADN/simulate_deterministicMADN.py:12-35) ---------------
def embed_dim(env):
    """floats of the env embedding: board, current_player, pins, reward, done, action_set"""
    return int(_lib.lib().dogstep_madn_det_embed_dim(C.byref(env.cfg())))


def policy_function(env):
    """policy_function (:495-507) -> f32 [n, 24]: 100 * valid_action + 200 * winning_action (:480-493)"""
    lg = torch.empty((env.n, 24), dtype=torch.float32, device=env.device)
    cfg, st = env.cfg(), env.cstate()
    _lib.check(_lib.lib().dogstep_madn_det_policy_function(C.byref(st), C.c_int64(env.n), C.byref(cfg), _lib.ptr(lg), _lib.stream()),
               "madn_det_policy_function")
    return _out(env, lg)


def root_fn(env, rng_key):
    """root_fn (:551-566); rng_key uint32 [n, 2].  value = rollout(env, key) (:509-541) as a scalar per game: the reference's
    float32[4] holds four equal entries (see include/dogstep.h)."""
    from .. import mcts
    n, dev = env.n, env.device
    prior = torch.empty((n, 24), dtype=torch.float32, device=dev)
    value = torch.empty(n, dtype=torch.float32, device=dev)
    emb = torch.empty((n, embed_dim(env)), dtype=torch.float32, device=dev)
    cfg, st = env.cfg(), env.cstate()
    _lib.check(_lib.lib().dogstep_madn_det_root_fn(C.byref(st), C.c_int64(n), C.byref(cfg), _lib.ptr(rng_key.contiguous()), _lib.ptr(prior),
                                                  _lib.ptr(value), _lib.ptr(emb), _lib.stream()), "madn_det_root_fn")
    return mcts.RootFnOutput(prior, value, emb)


def value_function(env, rng_key):
    """value_function (:543-549) = rollout(env, rng_key)"""
    return root_fn(env, rng_key).value


rollout = value_function


def make_recurrent_fn(env):
    """recurrent_fn (:568-590) for games with `env`'s static fields (players, layout, distance, rule dict), on the env embedding"""
    from .. import mcts
    cfg = env.cfg()

    def recurrent_fn(params, rng_key, action, embedding):
        n, dev = embedding.shape[0], embedding.device
        prior = torch.empty((n, 24), dtype=torch.float32, device=dev)
        value, reward, discount = (torch.empty(n, dtype=torch.float32, device=dev) for _ in range(3))
        nxt = torch.empty_like(embedding)
        _lib.check(_lib.lib().dogstep_madn_det_recurrent_fn(C.c_int64(n), C.byref(cfg), _lib.ptr(rng_key.contiguous()),
                                                           _lib.ptr(action.to(torch.int32).contiguous()), _lib.ptr(embedding.contiguous()),
                                                           _lib.ptr(prior), _lib.ptr(value), _lib.ptr(reward), _lib.ptr(discount),
                                                           _lib.ptr(nxt), _lib.stream()), "madn_det_recurrent_fn")
        return mcts.RecurrentFnOutput(reward, discount, prior, value), nxt
    return recurrent_fn


def run_gumbel(rng_key, env, num_simulations, graph_cache=None):
    """run_gumbel (MADN/simulate_deterministicMADN.py:12-35): mctx.gumbel_muzero_policy on the true env with rollout values,
    invalid_actions = ~valid_action, max_depth 350, qtransform_by_min_max(-1, 1) — batched over games (the reference runs
    batch_size = 1); rng_key uint32 [n, 2]"""
    import functools
    from .. import mcts
    from ..TicTacToe.mcts import _split_each
    key1, key2 = _split_each(rng_key, 0), _split_each(rng_key, 1)    # key1, key2 = split(rng_key)
    root = root_fn(env, _split_each(key2, 0))                         # root_fn(env, split(key2, batch_size = 1)[0])
    invalid = ~valid_action(env).reshape(env.n, 24)
    return mcts.gumbel_muzero_policy(None, key1, root, make_recurrent_fn(env), num_simulations, invalid_actions=invalid, max_depth=350,
                                     qtransform=functools.partial(mcts.qtransform_by_min_max, min_value=-1, max_value=1),
                                     graph_cache=graph_cache)

