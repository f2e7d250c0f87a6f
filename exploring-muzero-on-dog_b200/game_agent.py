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
                               root_fn, decision_recurrent_fn, chance_recurrent_fn, dirichlet_noise=None, dirichlet_fraction=0.25):
    """run_stochastic_muzero_mcts (muzero_classic_madn.py:464-517), batched."""
    key2 = _split_each(rng_key, 1)
    root = root_fn(params, observations)
    out = mcts.stochastic_muzero_policy(params, key2, root, decision_recurrent_fn, chance_recurrent_fn, num_simulations,
                                        invalid_actions=invalid_actions, max_depth=max_depth,
                                        qtransform=mcts.qtransform_by_parent_and_siblings, temperature=temperature,
                                        dirichlet_noise=dirichlet_noise, dirichlet_fraction=dirichlet_fraction)
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

    def clear(self):
        """back to init_buffers in place (a SelfPlayLoop that is run again keeps its buffers and its captured graph)"""
        for t in (self.obs, self.act, self.rew, self.player, self.discount, self.val, self.mask, self.pol, self.idx, self.dice,
                  self.dice_dist):
            if t is not None:
                t.zero_()
        self.team.fill_(-1)

    def carrays(self):
        ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        return _lib.tag(_lib.ReplayArrays(self.n, self.max_steps, int(np.prod(self.obs_shape)), self.action_dim,
                                          int(self.obs.dtype == torch.int8), int(self.stochastic), ptr(self.obs), ptr(self.act),
                                          ptr(self.rew), ptr(self.val), ptr(self.pol), ptr(self.mask), ptr(self.player), ptr(self.team),
                                          ptr(self.discount), ptr(self.idx), ptr(self.dice), ptr(self.dice_dist)), self.obs.device)

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


class SelfPlayLoop:
    """play_batch_of_games_jitted (game_agent.py:50-183 / game_agent_stochastic.py:52-218) without a host round trip per
    lockstep iteration.

    * the loop key lives on the device: `rng_key, *step_keys = split(rng_key, n + 1)` is dogstep_random_split_chain;
    * the termination test `any(~dones)` is a device counter copied to a ring of pinned host words behind an event; the host
      looks at the newest copy that has ALREADY arrived (never waits for the iteration in flight) and stays at most
      `lookahead` iterations ahead of the device.  Iterations enqueued after the last game ended change nothing: every env /
      trajectory kernel skips finished games, exactly like the reference's `lax.cond(~done, ...)`;
    * every buffer is allocated once; with `cuda_graph=True` one whole iteration (key split, [throw_die], encode_board,
      valid_action, root network, tree search, agent step, live-game count) is captured once and replayed — `search_fn`
      must then be capturable (CUDA work on the current stream only, no host synchronisation, params updated in place).
    """

    def __init__(self, envs, num_envs, input_shape, params, max_steps, *, search_fn, obs_dtype=torch.float32, cuda_graph=False,
                 lookahead=4):
        self.envs, self.n, self.params, self.max_steps, self.search_fn = envs, num_envs, params, max_steps, search_fn
        self.dog = isinstance(envs, dg.DOG)
        self.det = self.dog or isinstance(envs, dm.deterministic_MADN)
        self.mod = dm if self.det else cm
        dev = self.dev = envs.raw("done").device
        self.raw_dog_obs = self.dog and tuple(input_shape) == (dg.RAW_OBS_SIZE,)
        action_dim = dg.get_play_action_size(envs) + 14 if self.dog else (24 if self.det else 4)
        self.traj = Trajectories(num_envs, max_steps, input_shape, action_dim, not self.det, dev, obs_dtype)
        self.loop_key = torch.zeros(2, dtype=torch.uint32, device=dev)
        self.step_keys = torch.empty((num_envs, 2), dtype=torch.uint32, device=dev)
        self.live = torch.zeros(1, dtype=torch.int32, device=dev)
        self.lookahead = max(1, int(lookahead))
        self.live_host = torch.full((self.lookahead,), -1, dtype=torch.int32).pin_memory()
        self.events = [torch.cuda.Event() for _ in range(self.lookahead)]
        self.cuda_graph, self.graph = bool(cuda_graph), None
        self.iterations = 0          # lockstep iterations the reference's while_loop would have run
        self.enqueued = 0            # iterations actually enqueued (>= iterations: the poll lags)
        self._ran = False

    def _iteration(self):
        envs, n = self.envs, self.n
        _lib.check(_lib.lib().dogstep_random_split_chain(_lib.ptr(self.loop_key), C.c_int64(n), _lib.ptr(self.step_keys), _lib.stream()),
                   "random_split_chain")
        if not self.det:
            cfg, st = envs.cfg(), envs.cstate()
            _lib.check(_lib.lib().dogstep_madn_cls_throw_die_active(C.byref(st), C.c_int64(n), C.byref(cfg), _lib.stream()),
                       "throw_die_active")
        if self.dog:  # the reference has no DOG encoder (DOG/dog.py:1264-1272): dog.encode_board is this repo's design, the raw
            # 74-byte mover view remains for callers that ask for it by shape
            obs = dg.raw_observation(envs) if self.raw_dog_obs else dg.encode_board(envs)
            valid = dg.valid_actions(envs)
        else:
            obs, valid = self.mod.encode_board(envs), self.mod.valid_action(envs).reshape(n, -1)
        action, weights, value = self.search_fn(self.params, self.step_keys, obs, ~valid)
        agent_step(envs, self.traj, action, value, weights, obs)
        self.live.copy_((~envs.raw("done")).sum(dtype=torch.int32))

    def run(self, rng_key):
        """-> the buffers dict the reference returns; `envs` is stepped in place.  May be called again after the caller has
        re-seeded the SAME env object (env_reset(..., out=envs)): buffers and the captured graph are reused."""
        dev = self.dev
        with torch.cuda.device(dev):
            if self._ran:
                self.traj.clear()
            self._ran = True
            self.loop_key.copy_(torch.from_numpy(np.asarray(rng_key, dtype=np.uint32).copy()), non_blocking=False)
            if bool(self.envs.raw("done").all()) or self.max_steps <= 0:   # one test before the loop, like cond_fn on entry
                return self.traj.as_dict()
            if self.cuda_graph and self.graph is None:
                # capture only: the captured kernels do not run, so the first replay is iteration 0
                self.graph = torch.cuda.CUDAGraph()
                saved = self.loop_key.clone()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):  # warm-up outside capture (allocator pools, lazily created handles), then undone:
                    snap = self.envs.clone()   # env leaves, row counters and the key are restored; the trajectory rows the
                    idx0 = self.traj.idx.clone()  # warm-up wrote are the ones iteration 0 writes again
                    self._iteration()
                    for k, t in self.envs._t.items():
                        t.copy_(snap._t[k])
                    self.traj.idx.copy_(idx0)
                    self.loop_key.copy_(saved)
                torch.cuda.current_stream().wait_stream(side)
                with torch.cuda.graph(self.graph):
                    self._iteration()
            step, seen_done_at = 0, None
            self.live_host.fill_(-1)
            while step < self.max_steps:
                slot = step % self.lookahead
                if step >= self.lookahead:
                    self.events[slot].synchronize()          # iteration step - lookahead: long finished unless the host runs far ahead
                    if int(self.live_host[slot]) == 0:
                        seen_done_at = step - self.lookahead
                        break
                if self.graph is not None:
                    self.graph.replay()
                else:
                    self._iteration()
                self.live_host[slot:slot + 1].copy_(self.live, non_blocking=True)
                self.events[slot].record()
                step += 1
            self.enqueued = step
            torch.cuda.current_stream().synchronize()
            # the iteration after which no game was live (what the reference's step_count ends at)
            idx_max = int(self.traj.idx.max().item())
            self.iterations = min(step, max(idx_max, 0)) if bool(self.envs.raw("done").all()) else step
        return self.traj.as_dict()


def play_batch_of_games(envs, num_envs, input_shape, params, rng_key, num_simulations, max_depth, max_steps, temp, *,
                        search_fn, obs_dtype=torch.float32, cuda_graph=False, lookahead=4):
    """play_batch_of_games_jitted (game_agent.py:50-183 / game_agent_stochastic.py:52-218).
    search_fn(params, step_keys [n,2], obs int8 [n,C,T], invalid bool [n,A]) -> (action [n], action_weights [n,A], root_value [n]).
    `envs` is stepped in place.  Returns the buffers dict the reference returns.  No host synchronisation per lockstep
    iteration (SelfPlayLoop); cuda_graph=True additionally replays each iteration as one CUDA graph."""
    return SelfPlayLoop(envs, num_envs, input_shape, params, max_steps, search_fn=search_fn, obs_dtype=obs_dtype,
                        cuda_graph=cuda_graph, lookahead=lookahead).run(rng_key)


def play_n_games_v3(params, rng_key, input_shape, num_envs, num_simulation, max_depth, max_steps, temp, *, root_fn, recurrent_fn,
                    rules=RULES, obs_dtype=torch.float32, device="cuda", cuda_graph=False):
    """play_n_games_v3 (game_agent.py:185-192): seeds = randint(subkey, (num_envs,), 0, 1e6); the SAME subkey drives the loop."""
    rng_key, subkey = jaxrand.split_host(rng_key)
    seeds = jaxrand.randint(subkey, num_envs, 0, 1000000, device=device)
    envs = dm.env_reset(0, num_players=4, distance=10, starting_player=0, seed=seeds, device=device, **rules)

    def search_fn(p, keys, obs, invalid):
        out, root_value = run_muzero_mcts(p, keys, obs.to(torch.float32), invalid, num_simulation, max_depth, temp, root_fn=root_fn,
                                          recurrent_fn=recurrent_fn)
        return out.action, out.action_weights, root_value

    return play_batch_of_games(envs, num_envs, input_shape, params, subkey, num_simulation, max_depth, max_steps, temp,
                               search_fn=search_fn, obs_dtype=obs_dtype, cuda_graph=cuda_graph)


STOCHASTIC_RULES = dict(RULES, enable_dice_rethrow=True)  # MuZero_Classic_MADN/game_agent_stochastic.py:13-24


def play_n_games_v3_stochastic(params, rng_key, input_shape, num_envs, num_simulation, max_depth, max_steps, temp, *, root_fn,
                               decision_recurrent_fn, chance_recurrent_fn, rules=STOCHASTIC_RULES, obs_dtype=torch.float32,
                               device="cuda", cuda_graph=False, dirichlet_noise=None, dirichlet_fraction=0.25):
    """play_n_games_v3 of the dice game (game_agent_stochastic.py:220-244).  Per live game and iteration the reference draws
    `key1, key2 = split(step_key)` (:89) and hands key2 to run_stochastic_muzero_mcts, which splits once more (:476 of
    muzero_classic_madn.py): the search_fn below applies both splits, so the key stream equals the reference's."""
    rng_key, subkey = jaxrand.split_host(rng_key)
    seeds = jaxrand.randint(subkey, num_envs, 0, 1000000, device=device)
    envs = cm.env_reset(0, num_players=4, distance=10, starting_player=0, seed=seeds, device=device, **rules)

    def search_fn(p, keys, obs, invalid):
        key2 = _split_each(keys, 1)                     # key1, key2 = jax.random.split(key)  (game_agent_stochastic.py:89)
        out, root_value = run_stochastic_muzero_mcts(p, key2, obs.to(torch.float32), invalid, num_simulation, max_depth, temp,
                                                     root_fn=root_fn, decision_recurrent_fn=decision_recurrent_fn,
                                                     chance_recurrent_fn=chance_recurrent_fn, dirichlet_noise=dirichlet_noise,
                                                     dirichlet_fraction=dirichlet_fraction)
        return out.action, out.action_weights, root_value

    return play_batch_of_games(envs, num_envs, input_shape, params, subkey, num_simulation, max_depth, max_steps, temp,
                               search_fn=search_fn, obs_dtype=obs_dtype, cuda_graph=cuda_graph)


def play_n_dog_games(params, rng_key, num_envs, num_simulation, max_depth, max_steps, temp, *, root_fn, recurrent_fn, rules=DOG_RULES,
                     obs_dtype=torch.int8, device="cuda", max_num_considered_actions=16, graph_cache=None, cuda_graph=False,
                     return_loop=False, encoded_obs=False):
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

    loop = SelfPlayLoop(envs, num_envs, (dg.obs_planes(envs), envs.static["total_board_size"]) if encoded_obs else (dg.RAW_OBS_SIZE,),
                        params, max_steps, search_fn=search_fn, obs_dtype=obs_dtype, cuda_graph=cuda_graph)
    buffers = loop.run(subkey)
    return (envs, buffers, loop) if return_loop else (envs, buffers)
