"""Lockstep self-play loop: mirror of play_batch_of_games_jitted / play_n_games_v3 / run_muzero_mcts
(MuZero_det_MADN/game_agent.py:50-192, muzero_deterministic_madn.py:663-704) and of the dice variant
(MuZero_Classic_MADN/game_agent_stochastic.py:52-218, muzero_classic_madn.py:464-517).

Per lockstep iteration: split the loop key (device) -> [dice: throw_die on live games] -> encode_board -> valid_action ->
root network -> tree search (select / network / expand kernels) -> one fused agent-step kernel (env_step or no_step,
reward / discount class targets, trajectory row).  The networks are the caller's (`root_fn(params, obs)`,
`recurrent_fn(params, rng, action, embedding)`; Flax through DLPack in the reference's setting, torch in the tests).
Unlike the vmapped reference, finished games and games without a legal move cost nothing in the env kernels; the search
still runs on the full batch so that network calls keep a fixed shape.
"""
import ctypes as C
import functools

import numpy as np
import torch

from . import _lib, jaxrand, mcts
from .DOG import dog as dg
from .MADN import classic_madn as cm
from .MADN import deterministic_madn as dm

RULES = dict(enable_teams=True, enable_initial_free_pin=True, enable_circular_board=False, enable_friendly_fire=False,
             enable_start_blocking=False, enable_jump_in_goal_area=True, enable_start_on_1=True, enable_bonus_turn_on_6=True,
             must_traverse_start=False)  # MuZero_det_MADN/game_agent.py:12-22
DOG_RULES = dict(enable_teams=True, enable_initial_free_pin=False, enable_circular_board=True, enable_friendly_fire=True,
                 enable_start_blocking=True, enable_jump_in_goal_area=False, must_traverse_start=True)  # MuZero_DOG/game_agent.py:12-23


def _split_each(keys, index):
    out = torch.empty_like(keys)
    _lib.check(_lib.lib().dogstep_random_split_each(_lib.ptr(keys), C.c_int64(keys.shape[0]), C.c_uint32(index), _lib.ptr(out),
                                                   _lib.stream()), "random_split_each")
    return out


def run_muzero_mcts(params, rng_key, observations, invalid_actions, num_simulations, max_depth, temperature, *, root_fn,
                    recurrent_fn):
    """run_muzero_mcts (muzero_deterministic_madn.py:663-704), batched: rng_key uint32 [games, 2]."""
    key2 = _split_each(rng_key, 1)                      # key1, key2 = split(rng_key); key1 unused
    root = root_fn(params, observations)
    out = mcts.gumbel_muzero_policy(params, key2, root, recurrent_fn, num_simulations, invalid_actions=invalid_actions,
                                    max_depth=max_depth,
                                    qtransform=functools.partial(mcts.qtransform_completed_by_mix_value, value_scale=0.5),
                                    gumbel_scale=temperature)
    return out, out.search_tree.summary().value


def run_stochastic_muzero_mcts(params, rng_key, observations, invalid_actions, num_simulations, max_depth, temperature, *,
                               root_fn, decision_recurrent_fn, chance_recurrent_fn, dirichlet_noise=None):
    """run_stochastic_muzero_mcts (muzero_classic_madn.py:464-517), batched."""
    key2 = _split_each(rng_key, 1)
    root = root_fn(params, observations)
    out = mcts.stochastic_muzero_policy(params, key2, root, decision_recurrent_fn, chance_recurrent_fn, num_simulations,
                                        invalid_actions=invalid_actions, max_depth=max_depth,
                                        qtransform=mcts.qtransform_by_parent_and_siblings, temperature=temperature,
                                        dirichlet_noise=dirichlet_noise)
    return out, torch.clamp(out.search_tree.node_values[:, 0], -1.0, 1.0)


class Trajectories:
    """init_buffers of play_batch_of_games_jitted (game_agent.py:158-169) as device tensors"""

    def __init__(self, n, max_steps, obs_shape, action_dim, stochastic, device, obs_dtype=torch.float32):
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=device)
        self.n, self.max_steps, self.obs_shape, self.action_dim, self.stochastic = n, max_steps, tuple(obs_shape), action_dim, stochastic
        self.obs = z((n, max_steps, *obs_shape), obs_dtype)
        self.act, self.rew, self.player, self.discount = (z((n, max_steps), torch.int32) for _ in range(4))
        self.val, self.mask = z((n, max_steps), torch.float32), z((n, max_steps), torch.float32)
        self.pol = z((n, max_steps, action_dim), torch.float32)
        self.team = torch.full((n, max_steps), -1, dtype=torch.int32, device=device)
        self.idx = z((n,), torch.int32)
        self.dice = z((n, max_steps), torch.int32) if stochastic else None
        self.dice_dist = z((n, max_steps, 6), torch.float32) if stochastic else None

    def carrays(self):
        ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        return _lib.ReplayArrays(self.n, self.max_steps, int(np.prod(self.obs_shape)), self.action_dim,
                                 int(self.obs.dtype == torch.int8), int(self.stochastic), ptr(self.obs), ptr(self.act), ptr(self.rew),
                                 ptr(self.val), ptr(self.pol), ptr(self.mask), ptr(self.player), ptr(self.team), ptr(self.discount),
                                 ptr(self.idx), ptr(self.dice), ptr(self.dice_dist))

    def as_dict(self):
        d = dict(obs=self.obs, act=self.act, rew=self.rew, val=self.val, pol=self.pol, mask=self.mask, player=self.player,
                 team=self.team, discount=self.discount, idx=self.idx)
        if self.stochastic:
            d.update(dice=self.dice, dice_dist=self.dice_dist)
        return d


def agent_step(envs, traj, action, root_value, action_weights, obs):
    """the post-search part of one lockstep iteration (fused kernel), in place"""
    if isinstance(envs, dg.DOG):
        fn = _lib.lib().dogstep_dog_agent_step
    else:
        fn = _lib.lib().dogstep_madn_det_agent_step if isinstance(envs, dm.deterministic_MADN) else _lib.lib().dogstep_madn_cls_agent_step
    cfg, st, tr = envs.cfg(), envs.cstate(), traj.carrays()
    _lib.check(fn(C.byref(st), C.c_int64(envs.n), C.byref(cfg), _lib.ptr(action.to(torch.int32).contiguous()),
                  _lib.ptr(root_value.float().contiguous()), _lib.ptr(action_weights.float().contiguous()),
                  _lib.ptr(obs.contiguous()), C.byref(tr), _lib.stream()), "agent_step")


def play_batch_of_games(envs, num_envs, input_shape, params, rng_key, num_simulations, max_depth, max_steps, temp, *,
                        search_fn, obs_dtype=torch.float32):
    """play_batch_of_games_jitted (game_agent.py:50-183 / game_agent_stochastic.py:52-218).
    search_fn(params, step_keys [n,2], obs int8 [n,C,T], invalid bool [n,A]) -> (action [n], action_weights [n,A], root_value [n]).
    `envs` is stepped in place.  Returns the buffers dict the reference returns."""
    dog = isinstance(envs, dg.DOG)
    det = dog or isinstance(envs, dm.deterministic_MADN)
    mod = dm if det else cm
    dev = envs.device
    action_dim = dg.get_play_action_size(envs) + 14 if dog else (24 if det else 4)
    traj = Trajectories(num_envs, max_steps, input_shape, action_dim, not det, dev, obs_dtype)
    key = np.asarray(rng_key, dtype=np.uint32)
    step = 0
    while step < max_steps and not bool(envs.raw("done").all()):
        keys = jaxrand.split(key, num_envs + 1, device=dev)         # rng_key, *step_keys = split(rng_key, num_envs + 1)
        key = keys[0].cpu().numpy()
        step_keys = keys[1:].contiguous()
        if not det:
            cfg, st = envs.cfg(), envs.cstate()
            _lib.check(_lib.lib().dogstep_madn_cls_throw_die_active(C.byref(st), C.c_int64(envs.n), C.byref(cfg), _lib.stream()),
                       "throw_die_active")
        if dog:  # the reference has no DOG encoder (DOG/dog.py:1264-1272): raw mover-view leaves stand in for it
            obs, valid = dg.raw_observation(envs), dg.valid_actions(envs)
        else:
            obs, valid = mod.encode_board(envs), mod.valid_action(envs).reshape(num_envs, -1)
        action, weights, value = search_fn(params, step_keys, obs, ~valid)
        agent_step(envs, traj, action, value, weights, obs)
        step += 1
    return traj.as_dict()


def play_n_games_v3(params, rng_key, input_shape, num_envs, num_simulation, max_depth, max_steps, temp, *, root_fn, recurrent_fn,
                    rules=RULES, obs_dtype=torch.float32, device="cuda"):
    """play_n_games_v3 (game_agent.py:185-192): seeds = randint(subkey, (num_envs,), 0, 1e6); the SAME subkey drives the loop."""
    rng_key, subkey = jaxrand.split_host(rng_key)
    seeds = jaxrand.randint(subkey, num_envs, 0, 1000000, device=device)
    envs = dm.env_reset(0, num_players=4, distance=10, starting_player=0, seed=seeds, device=device, **rules)

    def search_fn(p, keys, obs, invalid):
        out, root_value = run_muzero_mcts(p, keys, obs.to(torch.float32), invalid, num_simulation, max_depth, temp, root_fn=root_fn,
                                          recurrent_fn=recurrent_fn)
        return out.action, out.action_weights, root_value

    return play_batch_of_games(envs, num_envs, input_shape, params, subkey, num_simulation, max_depth, max_steps, temp,
                               search_fn=search_fn, obs_dtype=obs_dtype)


def play_n_dog_games(params, rng_key, num_envs, num_simulation, max_depth, max_steps, temp, *, root_fn, recurrent_fn, rules=DOG_RULES,
                     obs_dtype=torch.int8, device="cuda", max_num_considered_actions=16, graph_cache=None):
    """BASELINE config 5: play_n_games_v3's shape (game_agent.py:185-192) on the DOG env (MuZero_DOG/game_agent.py:12-44 rules and
    batch_reset), Gumbel MuZero search over the 806 DOG actions (MuZero_DOG/muzero_dog.py:101-136).  The reference's DOG
    networks are stubs, so root_fn / recurrent_fn are the caller's."""
    rng_key, subkey = jaxrand.split_host(rng_key)
    seeds = jaxrand.randint(subkey, num_envs, 0, 1000000, device=device)
    envs = dg.env_reset(0, num_players=4, distance=10, starting_player=0, seed=seeds, device=device, **rules)

    def search_fn(p, keys, obs, invalid):
        key2 = _split_each(keys, 1)
        out = mcts.gumbel_muzero_policy(p, key2, root_fn(p, obs.to(torch.float32)), recurrent_fn, num_simulation,
                                        invalid_actions=invalid, max_depth=max_depth,
                                        qtransform=functools.partial(mcts.qtransform_completed_by_mix_value, value_scale=0.5),
                                        gumbel_scale=temp, max_num_considered_actions=max_num_considered_actions,
                                        **({} if graph_cache is None else {"graph_cache": graph_cache}))
        return out.action, out.action_weights, out.search_tree.summary().value

    return envs, play_batch_of_games(envs, num_envs, (dg.RAW_OBS_SIZE,), params, subkey, num_simulation, max_depth, max_steps, temp,
                                     search_fn=search_fn, obs_dtype=obs_dtype)
