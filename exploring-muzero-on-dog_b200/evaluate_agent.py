"""Evaluation loop: mirror of play_eval_loop_jitted / play_n_games_for_eval_jitted (MuZero_det_MADN/evaluate_agent.py:715-930)
and of the dice game's twin (MuZero_Classic_MADN/evaluate_agent_stochastic.py:738-905; pass a classic_MADN env).

Four seats, each played by an agent given as a params dict with a 'type' entry exactly as in the reference: 3 = random
legal policy, 2 = the rule-based scorer (do_rule_based :780-878), anything else = MuZero tree search with that seat's
network parameters.  One fused kernel per lockstep iteration (`dogstep_madn_det_eval_step`): legal mask, the seat's
policy, env_step / no_step, winner bookkeeping.  Search seats: `search_fn(params_tuple, step_keys [n,2], obs int8
[n,34,56], invalid bool [n,24], current_player int8 [n]) -> action int32 [n]` runs first on the whole batch (fixed shapes,
like the vmapped reference) and the kernel picks its action only where the seat to move searches.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, jaxrand
from .MADN import classic_madn as cm
from .MADN import deterministic_madn as dm

RULES = dict(enable_teams=True, enable_initial_free_pin=True, enable_circular_board=False, enable_friendly_fire=False,
             enable_start_blocking=False, enable_jump_in_goal_area=True, enable_start_on_1=True, enable_bonus_turn_on_6=True,
             must_traverse_start=False)  # evaluate_agent.py:932-942


def agent_types(params_tuple):
    """params['type'] per seat (a plain int also works); search seats get 0"""
    out = []
    for p in params_tuple:
        t = p.get("type", 0) if isinstance(p, dict) else int(p)
        out.append(int(t) if int(t) in (2, 3) else 0)
    return out


def eval_step(envs, types, rng_key, search_action=None, winners=None, game_offset=0, active_count=None, throw=True):
    """one lockstep iteration, in place (dice game: `throw` = roll the die of the live games first, as the reference's loop body does)"""
    cfg, st = envs.cfg(), envs.cstate()
    at = (C.c_int32 * 4)(*[int(t) for t in (list(types) + [3, 3, 3, 3])[:4]])
    det = isinstance(envs, dm.deterministic_MADN)
    if not det and throw:  # the dice game throws first (evaluate_agent_stochastic.py:757), live games only
        _lib.check(_lib.lib().dogstep_madn_cls_throw_die_active(C.byref(st), C.c_int64(envs.n), C.byref(cfg), _lib.stream()),
                   "throw_die_active")
    fn = _lib.lib().dogstep_madn_det_eval_step if det else _lib.lib().dogstep_madn_cls_eval_step
    _lib.check(fn(C.byref(st), C.c_int64(envs.n), C.byref(cfg), at, _lib.ptr(search_action), _lib.host_key(rng_key),
                  C.c_int64(game_offset), _lib.ptr(winners), _lib.ptr(active_count), _lib.stream()), "madn_eval_step")


def play_eval_loop(envs, params_tuple, rng_key, num_envs, search_fn=None, max_steps=2000, game_offset=0, poll_every=16):
    """play_eval_loop_jitted (:733-930): returns (final_envs, winners int32 [num_envs, 4]); `envs` is stepped in place.
    The termination test any(~done) is polled every `poll_every` iterations (iterations on finished games are no-ops)."""
    types = agent_types(params_tuple)
    needs_search = any(t not in (2, 3) for t in types)
    if needs_search and search_fn is None:
        raise ValueError("a seat plays by tree search: search_fn is required")
    dev = envs.device
    winners = torch.zeros((num_envs, 4), dtype=torch.int32, device=dev)
    # rng_key, *step_keys = split(rng_key, num_envs + 1): element 0 does not depend on the count and is advanced in place on the host
    key = jaxrand.KeyChain(rng_key.numpy() if isinstance(rng_key, jaxrand.KeyChain) else np.asarray(rng_key, dtype=np.uint32))
    classic = not isinstance(envs, dm.deterministic_MADN)
    step = 0
    while step < max_steps:
        if step % poll_every == 0 and bool(envs.raw("done").all()):
            break
        action = None
        if classic:  # throw_die comes before encode_board / valid_action / the search (evaluate_agent_stochastic.py:757-760)
            cfg, st = envs.cfg(), envs.cstate()
            _lib.check(_lib.lib().dogstep_madn_cls_throw_die_active(C.byref(st), C.c_int64(envs.n), C.byref(cfg), _lib.stream()),
                       "throw_die_active")
        if needs_search:
            mod = cm if classic else dm
            step_keys = jaxrand.split(key.numpy(), num_envs + 1, device=dev)[1:].contiguous()
            obs = mod.encode_board(envs)
            valid = mod.valid_action(envs).reshape(num_envs, -1)
            action = search_fn(params_tuple, step_keys, obs, ~valid, envs.raw("current_player")).to(torch.int32).contiguous()
        eval_step(envs, types, key, action, winners, game_offset, throw=False)
        key.advance()
        step += 1
    return envs, winners


def play_n_games_for_eval(params_list, rng_key, num_envs=20, search_fn=None, rules=RULES, device="cuda"):
    """play_n_games_for_eval_jitted (:715-731): num_envs games for each of the four starting seats, seeds and loop key as the
    reference draws them.  Returns (winners int32 [4 * num_envs, 4], final envs)."""
    rng_key, subkey = jaxrand.split_host(rng_key)
    n = num_envs * 4
    seeds = jaxrand.randint(subkey, n, 0, 1000000, device=device)
    envs = dm.env_reset(0, num_players=4, distance=10, starting_player=0, seed=seeds, device=device, **rules)
    # batch_reset(seeds, repeat(arange(4), num_envs)): the starting seat only sets current_player (deterministic_madn.py:102-108)
    envs.raw("current_player").copy_(torch.arange(4, device=device).repeat_interleave(num_envs).to(torch.int8))
    envs, winners = play_eval_loop(envs, tuple(params_list), subkey, n, search_fn=search_fn)
    return winners, envs
