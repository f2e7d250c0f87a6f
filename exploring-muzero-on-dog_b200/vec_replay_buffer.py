"""Device-resident stand-in for the reference's replay buffers.

Same constructor, attributes and methods as VectorizedReplayBuffer (MuZero_det_MADN/vec_replay_buffer.py:9-264; the copy
in MuZero_DOG/ is byte-identical) and VectorizedReplayBufferStochastic
(MuZero_Classic_MADN/vec_replay_buffer_stochastic.py): `.save_games_from_buffers(all_buffers)`, `.sample_batch()`,
`.size`, `.position`, `.bootstrap_value_target`.  The arrays live in HBM (one shard per GPU) and both operations are
libdogstep.so kernels; `sample_batch` returns CUDA tensors (zero-copy to JAX through DLPack).

Extras beyond the reference: `seed=` makes sampling reproducible (the reference draws from an unseeded np.random),
`sample_batch(plan=(ep_indices, t_starts))` replays a given draw, `obs_dtype=torch.int8` stores observations 4x
smaller, `sample_batch_global()` all-gathers the per-rank batches over NCCL (the one collective of the design), and
`prioritized=True` adds proportional prioritised sampling (`sample_batch_prioritized`, `update_priorities`).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, jaxrand

GAMMA = 0.997          # vec_replay_buffer.py:70
TERMINAL_RATIO = 0.25  # :71
_TRAJ = dict(observations="obs", actions="act", rewards="rew", root_values="val", child_visits="pol", masks="mask",
             players="player", teams="team", discounts="discount", dice_outcomes="dice", dice_distributions="dice_dist")


class VectorizedReplayBuffer:
    STOCHASTIC = False

    def __init__(self, capacity, batch_size, unroll_steps, td_steps, obs_shape=(14, 56), action_dim=24, max_episode_length=500,
                 bootstrap_value_target=True, device="cuda", obs_dtype=torch.float32, seed=0, prioritized=False,
                 priority_alpha=1.0, priority_eps=0.0):
        self.capacity, self.batch_size, self.unroll_steps, self.td_steps = capacity, batch_size, unroll_steps, td_steps
        self.obs_shape, self.action_dim, self.max_episode_length = tuple(obs_shape), action_dim, max_episode_length
        self.bootstrap_value_target = bootstrap_value_target
        self.device = torch.device(device)
        T, dev = max_episode_length, self.device
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)
        self.observations = z((capacity, T, *obs_shape), obs_dtype)
        self.actions = torch.full((capacity, T), -1 if self.STOCHASTIC else 0, dtype=torch.int32, device=dev)
        self.rewards, self.players, self.teams, self.discounts = (z((capacity, T), torch.int32) for _ in range(4))
        self.root_values, self.masks = z((capacity, T), torch.float32), z((capacity, T), torch.float32)
        self.child_visits = z((capacity, T, action_dim), torch.float32)
        self.episode_lengths = z((capacity,), torch.int32)
        self.dice_outcomes = torch.full((capacity, T), -1, dtype=torch.int32, device=dev) if self.STOCHASTIC else None
        self.dice_distributions = z((capacity, T, 6), torch.float32) if self.STOCHASTIC else None
        self.position, self.size = 0, 0
        self.prioritized, self.max_priority = bool(prioritized), 1.0
        self.priority_alpha, self.priority_eps = float(priority_alpha), float(priority_eps)  # stored = (|p| + eps) ** alpha
        if self.prioritized:  # fixed-point priorities (2^-20 units) + exact uint64 row sums, see include/dogstep.h
            self.priorities = z((capacity, T), torch.uint32)
            self.priority_row_sums = z((capacity,), torch.uint64)
            self._cdf_work = z((capacity,), torch.uint64)
        self._key = jaxrand.PRNGKey(seed)
        # GAMMA ** k in float64 exactly as NumPy evaluates it in the reference (:228, :234)
        self._gamma_pow = torch.from_numpy(np.float64(GAMMA) ** np.arange(T + 1, dtype=np.int64)).to(dev)

    # ------------------------------------------------------------------ helpers
    def _arrays(self, src, capacity, T):
        ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        obs = src["observations"]
        return _lib.tag(_lib.ReplayArrays(capacity, T, int(np.prod(self.obs_shape)), self.action_dim, int(obs.dtype == torch.int8),
                                          int(self.STOCHASTIC), *[ptr(src.get(k)) for k in
                                                                  ("observations", "actions", "rewards", "root_values", "child_visits",
                                                                   "masks", "players", "teams", "discounts", "episode_lengths",
                                                                   "dice_outcomes", "dice_distributions")]), obs.device)

    def _own(self):
        names = ("observations", "actions", "rewards", "root_values", "child_visits", "masks", "players", "teams", "discounts",
                 "episode_lengths", "dice_outcomes", "dice_distributions")
        return self._arrays({k: getattr(self, k) for k in names}, self.capacity, self.max_episode_length)

    def _as(self, x, dtype):
        t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
        return t.to(device=self.device, dtype=dtype).contiguous()

    # ------------------------------------------------------------------ API of the reference
    def save_games_from_buffers(self, all_buffers):
        """save_games_from_buffers (:36-61).  `all_buffers` is the dict play_batch_of_games returns (device tensors or
        host arrays): obs/act/rew/val/pol/mask/player/team/discount [games, T', ...] and idx [games]."""
        idx = self._as(all_buffers["idx"], torch.int32)
        n_games = int(idx.numel())
        obs = all_buffers["obs"]
        obs_dt = torch.int8 if (isinstance(obs, torch.Tensor) and obs.dtype == torch.int8) else torch.float32
        src = dict(observations=self._as(obs, obs_dt), actions=self._as(all_buffers["act"], torch.int32),
                   rewards=self._as(all_buffers["rew"], torch.int32), root_values=self._as(all_buffers["val"], torch.float32),
                   child_visits=self._as(all_buffers["pol"], torch.float32), masks=self._as(all_buffers["mask"], torch.float32),
                   players=self._as(all_buffers["player"], torch.int32), teams=self._as(all_buffers["team"], torch.int32),
                   discounts=self._as(all_buffers["discount"], torch.int32), episode_lengths=idx)
        if self.STOCHASTIC:
            src["dice_outcomes"] = self._as(all_buffers["dice"], torch.int32)
            src["dice_distributions"] = self._as(all_buffers["dice_dist"], torch.float32)
        T_traj = src["actions"].shape[1]
        nonzero = idx > 0
        rank = torch.cumsum(nonzero.to(torch.int32), 0, dtype=torch.int32) - 1
        # games are written one ring slot after the other; more games than slots -> sequential chunks so that "the later
        # game wins" exactly as in the reference's loop
        count = int(nonzero.sum().item())
        done = 0
        buf = self._own()
        while done < count:
            take = min(self.capacity, count - done)
            sel = nonzero & (rank >= done) & (rank < done + take)
            slot = torch.where(sel, (self.position + rank - done) % self.capacity, torch.full_like(rank, -1)).to(torch.int32).contiguous()
            traj = self._arrays(src, n_games, T_traj)
            _lib.check(_lib.lib().dogstep_replay_save(C.byref(buf), C.byref(traj), C.c_int64(n_games), _lib.ptr(slot), _lib.stream()),
                       "replay_save")
            if self.prioritized:  # new episodes enter at the running maximum priority
                _lib.check(_lib.lib().dogstep_replay_prio_fill(_lib.ptr(self.priorities), _lib.ptr(self.priority_row_sums),
                                                              _lib.ptr(self.episode_lengths), C.c_int32(self.max_episode_length),
                                                              _lib.ptr(slot), C.c_int32(n_games), C.c_float(self.max_priority),
                                                              _lib.stream()), "replay_prio_fill")
            self.position = (self.position + take) % self.capacity
            self.size = min(self.size + take, self.capacity)
            done += take

    def draw_plan(self):
        """the random part of sample_batch (:73-97) on the device, keyed by the buffer's own key chain"""
        self._key, sub = jaxrand.split_host(self._key)
        ep = torch.empty(self.batch_size, dtype=torch.int32, device=self.device)
        ts = torch.empty(self.batch_size, dtype=torch.int32, device=self.device)
        buf = self._own()
        _lib.check(_lib.lib().dogstep_replay_plan(C.byref(buf), C.c_int32(self.size), C.c_int32(self.batch_size),
                                                 C.c_int32(self.unroll_steps), C.c_float(TERMINAL_RATIO), _lib.host_key(sub),
                                                 _lib.ptr(ep), _lib.ptr(ts), _lib.stream()), "replay_plan")
        return ep, ts

    def sample_batch(self, plan=None):
        """sample_batch (:63-264) -> dict with the reference's keys (:255-264), CUDA tensors"""
        if self.size < 1:
            raise ValueError("sample_batch on an empty buffer")
        ep, ts = self.draw_plan() if plan is None else (self._as(plan[0], torch.int32), self._as(plan[1], torch.int32))
        B, K, A, dev = int(ep.numel()), self.unroll_steps + 1, self.action_dim, self.device
        e = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)
        out = dict(observations=e((B, *self.obs_shape), torch.float32), actions=e((B, K - 1), torch.int32),
                   rewards=e((B, K - 1), torch.int32), policies=e((B, K, A), torch.float32), values=e((B, K), torch.float32),
                   masks=e((B, K), torch.float32), target_values=e((B, K), torch.float32), discount_targets=e((B, K - 1), torch.int32))
        if self.STOCHASTIC:
            out["dice_outcomes"] = e((B, K - 1), torch.int32)
            out["dice_probs"] = e((B, K - 1, 6), torch.float32)
        ptr = lambda k: C.c_void_p(out[k].data_ptr()) if k in out else None
        cb = _lib.tag(_lib.ReplayBatch(*[ptr(k) for k in ("observations", "actions", "rewards", "policies", "values", "masks",
                                                          "target_values", "discount_targets", "dice_outcomes", "dice_probs")]),
                      out["observations"].device)
        buf = self._own()
        _lib.check(_lib.lib().dogstep_replay_gather(C.byref(buf), C.c_int32(B), C.c_int32(self.unroll_steps), C.c_int32(self.td_steps),
                                                   C.c_int32(int(bool(self.bootstrap_value_target))), _lib.ptr(self._gamma_pow),
                                                   _lib.ptr(ep), _lib.ptr(ts), C.byref(cb), _lib.stream()), "replay_gather")
        return out

    # ------------------------------------------------------------------ prioritised sampling (extension)
    def draw_plan_prioritized(self):
        """(ep_indices, t_starts, P(e, t) float64) drawn proportionally to the stored priorities"""
        if not self.prioritized:
            raise ValueError("buffer was built with prioritized=False")
        self._key, sub = jaxrand.split_host(self._key)
        B, dev = self.batch_size, self.device
        ep, ts = torch.empty(B, dtype=torch.int32, device=dev), torch.empty(B, dtype=torch.int32, device=dev)
        prob = torch.empty(B, dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().dogstep_replay_plan_prioritized(_lib.ptr(self.priorities), _lib.ptr(self.priority_row_sums),
                                                             _lib.ptr(self._cdf_work), C.c_int32(self.size),
                                                             C.c_int32(self.max_episode_length), C.c_int32(B), _lib.host_key(sub),
                                                             _lib.ptr(ep), _lib.ptr(ts), _lib.ptr(prob), _lib.stream()),
                   "replay_plan_prioritized")
        return ep, ts, prob

    def sample_batch_prioritized(self, beta=1.0):
        """sample_batch with windows drawn proportionally to priority; adds `ep_indices`, `t_starts` (hand them back to
        update_priorities) and the importance weights `weights` = (N * P)^-beta / max over the batch."""
        if self.size < 1:
            raise ValueError("sample_batch on an empty buffer")
        ep, ts, prob = self.draw_plan_prioritized()
        out = self.sample_batch(plan=(ep, ts))
        n_items = self.episode_lengths[:self.size].sum().clamp(min=1).to(torch.float64)
        w = (n_items * prob.clamp(min=1e-300)) ** (-float(beta))
        out["weights"] = (w / w.max()).to(torch.float32)
        out["ep_indices"], out["t_starts"] = ep, ts
        return out

    def update_priorities(self, ep_indices, t_starts, priorities):
        """set the priority of the sampled (episode, ply) pairs to (|priorities| + priority_eps) ** priority_alpha (e.g. from
        |search value - target|); stored with a floor of one fixed-point unit; pairs outside the stored episodes are ignored"""
        pr = self._as(priorities, torch.float32).abs()
        if self.priority_eps or self.priority_alpha != 1.0:
            pr = ((pr + self.priority_eps) ** self.priority_alpha).contiguous()
        ep, ts = self._as(ep_indices, torch.int32), self._as(t_starts, torch.int32)
        self.max_priority = max(self.max_priority, float(pr.max().item()))
        _lib.check(_lib.lib().dogstep_replay_prio_update(_lib.ptr(self.priorities), _lib.ptr(self.priority_row_sums),
                                                        _lib.ptr(self.episode_lengths), C.c_int32(self.capacity),
                                                        C.c_int32(self.max_episode_length), C.c_int32(int(ep.numel())), _lib.ptr(ep),
                                                        _lib.ptr(ts), _lib.ptr(pr), _lib.stream()), "replay_prio_update")

    def sample_batch_global(self, group=None):
        """Every rank samples `batch_size` windows from its own shard; the batches are all-gathered (NCCL over NVLink on
        the GPU box, gloo in the CPU tests) so that each rank trains on the world-size-times-larger global batch."""
        return allgather_batch(self.sample_batch(), group)


class VectorizedReplayBufferStochastic(VectorizedReplayBuffer):
    STOCHASTIC = True

    def __init__(self, capacity, batch_size, unroll_steps, td_steps, obs_shape=(11, 56), action_dim=4, max_episode_length=500,
                 bootstrap_value_target=True, **kw):
        super().__init__(capacity, batch_size, unroll_steps, td_steps, obs_shape, action_dim, max_episode_length,
                         bootstrap_value_target, **kw)


def allgather_batch(batch, group=None):
    """all-gather every leaf of a sampled batch along the batch axis (the replay-shard exchange of the design).  The leaves
    are packed into ONE byte buffer so that the exchange is a single collective (1-2 MB per rank: latency, not bandwidth)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return batch
    world = dist.get_world_size(group)
    keys = sorted(batch.keys())
    flat = [batch[k].contiguous().view(torch.uint8).reshape(-1) for k in keys]
    sizes = [int(f.numel()) for f in flat]
    pad = [(-sz) % 16 for sz in sizes]                      # keep every leaf 16-byte aligned inside the packed buffer
    packed = torch.cat([torch.cat([f, f.new_zeros(p)]) if p else f for f, p in zip(flat, pad)])
    full = torch.empty((world, packed.numel()), dtype=torch.uint8, device=packed.device)
    dist.all_gather_into_tensor(full.reshape(-1), packed, group=group)
    out, off = {}, 0
    for k, sz, p in zip(keys, sizes, pad):
        v = batch[k]
        out[k] = full[:, off:off + sz].contiguous().view(v.dtype).reshape((world * v.shape[0],) + tuple(v.shape[1:]))
        off += sz + p
    return out
