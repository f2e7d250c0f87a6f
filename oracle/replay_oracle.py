"""replay_oracle.py — TEST INFRASTRUCTURE ONLY (CPU oracle; never imported by the product).

Plain-loop NumPy restatement of the reference's host replay buffer:
  VectorizedReplayBuffer            /root/reference/MuZero_det_MADN/vec_replay_buffer.py:9-264
  VectorizedReplayBufferStochastic  /root/reference/MuZero_Classic_MADN/vec_replay_buffer_stochastic.py
written sample by sample (the reference is one block of fancy indexing) with the random draws made explicit
inputs (`ep_indices`, `t_starts`), because the reference draws them from an unseeded np.random.

PARITY STATUS: pinned by tests/golden/replay_reference.npz — outputs of the reference classes themselves (imported
from /root/reference in the build container, np.random seeded, the draw sequence replayed to recover the plan).
"""
import numpy as np

GAMMA = 0.997           # vec_replay_buffer.py:70
TERMINAL_RATIO = 0.25   # :71

FIELDS = ("observations", "actions", "rewards", "root_values", "child_visits", "masks", "players", "teams", "discounts")
TRAJ_KEYS = dict(observations="obs", actions="act", rewards="rew", root_values="val", child_visits="pol", masks="mask",
                 players="player", teams="team", discounts="discount", dice_outcomes="dice", dice_distributions="dice_dist")


class ReplayOracle:
    def __init__(self, capacity, batch_size, unroll_steps, td_steps, obs_shape, action_dim, max_episode_length,
                 bootstrap_value_target=True, stochastic=False):
        T = max_episode_length
        self.capacity, self.batch_size, self.unroll_steps, self.td_steps = capacity, batch_size, unroll_steps, td_steps
        self.stochastic, self.bootstrap_value_target, self.T = stochastic, bootstrap_value_target, T
        self.observations = np.zeros((capacity, T, *obs_shape), np.float32)
        self.actions = np.full((capacity, T), -1 if stochastic else 0, np.int32)   # stochastic variant initialises to -1
        self.rewards = np.zeros((capacity, T), np.int32)
        self.root_values = np.zeros((capacity, T), np.float32)
        self.child_visits = np.zeros((capacity, T, action_dim), np.float32)
        self.masks = np.zeros((capacity, T), np.float32)
        self.players = np.zeros((capacity, T), np.int32)
        self.teams = np.zeros((capacity, T), np.int32)
        self.discounts = np.zeros((capacity, T), np.int32)
        self.episode_lengths = np.zeros(capacity, np.int32)
        if stochastic:
            self.dice_outcomes = np.full((capacity, T), -1, np.int32)
            self.dice_distributions = np.zeros((capacity, T, 6), np.float32)
        self.position, self.size = 0, 0

    def fields(self):
        return FIELDS + (("dice_outcomes", "dice_distributions") if self.stochastic else ())

    def save_games_from_buffers(self, buffers):
        """:36-61 — sequential ring write, zero-length games skipped, only the first `length` rows replaced"""
        lengths = np.asarray(buffers["idx"])
        for i in range(lengths.shape[0]):
            n = int(lengths[i])
            if n == 0:
                continue
            for f in self.fields():
                getattr(self, f)[self.position, :n] = np.asarray(buffers[TRAJ_KEYS[f]][i, :n])
            self.episode_lengths[self.position] = n
            self.position = (self.position + 1) % self.capacity
            self.size = min(self.size + 1, self.capacity)

    def gather(self, ep_indices, t_starts):
        """:104-264 given the drawn (episode, t_start) pairs"""
        B, K, TD, A = len(ep_indices), self.unroll_steps + 1, self.td_steps, self.child_visits.shape[2]
        out = dict(observations=np.zeros((B,) + self.observations.shape[2:], np.float32), actions=np.zeros((B, K - 1), np.int32),
                   rewards=np.zeros((B, K - 1), np.int32), policies=np.zeros((B, K, A), np.float32), values=np.zeros((B, K), np.float32),
                   masks=np.zeros((B, K), np.float32), target_values=np.zeros((B, K), np.float32),
                   discount_targets=np.zeros((B, K - 1), np.int32))
        if self.stochastic:
            out["dice_outcomes"] = np.zeros((B, K - 1), np.int32)
            out["dice_probs"] = np.zeros((B, K - 1, 6), np.float32)
        for b in range(B):
            e, t0 = int(ep_indices[b]), int(t_starts[b])
            n = int(self.episode_lengths[e])
            last = n - 1
            out["observations"][b] = self.observations[e, t0]
            fr, fp, ft = self.rewards[e, last], self.players[e, last], self.teams[e, last]
            won = (fr > 0) if self.stochastic else (fr == 2)
            for k in range(K):
                idx = t0 + k
                valid = idx < n
                ci = min(idx, last)
                if valid:
                    out["policies"][b, k] = self.child_visits[e, ci]
                    out["values"][b, k] = self.root_values[e, ci]
                    out["masks"][b, k] = self.masks[e, ci]
                if k < K - 1:
                    out["actions"][b, k] = self.actions[e, ci] if valid else 0
                    out["rewards"][b, k] = self.rewards[e, ci] if valid else 1
                    out["discount_targets"][b, k] = self.discounts[e, ci] if valid else 1
                    if self.stochastic:
                        out["dice_outcomes"][b, k] = max((self.dice_outcomes[e, ci] if valid else 0) - 1, 0)
                        out["dice_probs"][b, k] = self.dice_distributions[e, ci] if valid else np.float32(1.0 / 6.0)
                if not valid:
                    continue
                player, team = self.players[e, ci], self.teams[e, ci]
                z = 0.0
                if won:
                    z = (1.0 if fp == player else -1.0) if team == -1 else (1.0 if ft == team else -1.0)
                until = last - idx
                bi = min(idx + TD, last)
                bv = np.float64(self.root_values[e, bi])
                same = (team == self.teams[e, bi]) if team != -1 else (player == self.players[e, bi])
                if not same:
                    bv = -bv
                z = z * GAMMA ** np.int64(max(until, 0))
                if z == 0 or (until >= TD and self.bootstrap_value_target):
                    target = bv * GAMMA ** np.int64(min(TD, until))
                else:
                    target = z
                out["target_values"][b, k] = np.float32(min(max(target, -1.0), 1.0))
        return out


# ---------------------------------------------------------------------------------------------------------------------
# Prioritised sampling (extension beyond the reference; include/dogstep.h "Prioritised sampling").  Fixed-point priorities
# and Python integers: exact, so the CUDA path must agree bit for bit.
PRIO_SHIFT = 20


def prio_to_fixed(p):
    """float32 priority -> stored fixed-point value; floor of one unit (a stored ply never becomes undrawable)"""
    p = np.asarray(p, np.float32)
    v = np.floor(p.astype(np.float64) * float(1 << PRIO_SHIFT) + 0.5)
    return np.where(p > 0, np.clip(v, 1.0, 4294967295.0), 1.0).astype(np.uint64).astype(np.uint32)


def plan_prioritized(prio_fixed, size, bits):
    """prio_fixed uint32 [capacity, T]; bits uint32 [2B] (threefry bits of the draw key) -> ep, t, prob"""
    prio = np.asarray(prio_fixed, np.uint64)[:size]
    row_sum = [int(x) for x in prio.sum(1, dtype=np.uint64)]
    cdf = np.cumsum(np.array(row_sum, dtype=object))
    total = int(cdf[-1])
    B = len(bits) // 2
    ep, ts, prob = np.zeros(B, np.int32), np.zeros(B, np.int32), np.zeros(B, np.float64)
    for b in range(B):
        b64 = (int(bits[2 * b]) << 32) | int(bits[2 * b + 1])
        target = (b64 * total) >> 64
        e = next(i for i in range(size) if int(cdf[i]) > target)
        r = target - (int(cdf[e - 1]) if e else 0)
        acc, t, pv = 0, 0, 0
        for k in range(prio.shape[1]):
            pv = int(prio[e, k])
            acc += pv
            t = k
            if acc > r:
                break
        ep[b], ts[b], prob[b] = e, t, (pv / total if total else 0.0)
    return ep, ts, prob
