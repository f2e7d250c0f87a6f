#!/usr/bin/env python
"""bench.py — BASELINE.json metric ("env steps/s & MCTS sims/s at 1/2/4/8 B200; % of HBM roofline").

Headline (`value`): config 2.  A "step" is one whole pass of the hot path over one batch: reset 65,536 deterministic-MADN games
per GPU from device-resident seeds, then play every game to termination with the reference's random legal policy
(MuZero_det_MADN/evaluate_agent.py:733-930 do_random: valid_action -> categorical -> env_step / no_step, cap 2000).
`value` = env steps/s (active (game, iteration) pairs) summed over all GPUs / max-over-ranks device time.

`extras`, at EVERY N (per rank, whole-box aggregate = sum of units / max-over-ranks device time):
  selfplay_cfg3   4,096 dice-MADN games per GPU, stochastic MuZero search with 64 simulations per move (sims/s)
  selfplay_cfg5   8,192 DOG games per GPU, Gumbel MuZero search with 100 simulations over 806 actions, then replay save +
                  sample + the one collective of the design (sample_batch_global) (sims/s)
  dog_cfg4        16,384 DOG games per GPU, random legal policy to termination (env steps/s)
and at N = 1 the tree-kernels-alone numbers, the per-call path, the evaluation loop, config 1 and a CPU baseline (the oracle
on the host cores) beside each.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--games G] [--impl reference]
N > 1 is launched by torchrun (one rank per GPU); games are sharded with no data-path collective (weak scaling).
"""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RULES = dict(enable_teams=True, enable_initial_free_pin=True, enable_circular_board=False, enable_friendly_fire=False,
             enable_start_blocking=False, enable_jump_in_goal_area=True, enable_start_on_1=True,
             enable_bonus_turn_on_6=True, must_traverse_start=False)  # MuZero_det_MADN/game_agent.py:12-22
DOG_RULES = dict(enable_teams=True, enable_initial_free_pin=False, enable_circular_board=True, enable_friendly_fire=True,
                 enable_start_blocking=True, enable_jump_in_goal_area=False, must_traverse_start=True)  # MuZero_DOG/game_agent.py:12-23
BYTES_PER_STEP = 226  # SURVEY.md 8(d) cfg 2: 99 B state read + 99 B written + action 2 + mask 24 + reward/done 2
DOG_BYTES_PER_STEP = 1228  # SURVEY.md 8(d) cfg 4
MAX_STEPS = 2000      # evaluate_agent.py:918
METRIC = "env steps/s (deterministic MADN, random legal policy to termination)"
SMS, SCHEDULERS = 148, 4


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def _profile_facts():
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler:
    """SM clock / throttle reasons read through NVML DURING the timed regions (same fields as the nvidia-smi line of
    B200_PROFILING.md), from a helper thread (pynvml is ctypes: the GIL is released inside a call).  Inline sampling from the
    enqueuing thread was measured and dropped: the first NVML query after the GPU leaves idle blocks for ~30 ms (seen in both
    timed regions), which then sits inside an e2e step.  Period 5 ms; the slowest call is reported (`nvml_call_ms_max`)."""

    def __init__(self, index, period=0.005):
        import threading
        self.rows, self.period, self._stop, self.th, self.slowest = [], period, False, None, 0.0
        self._threading = threading
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def start(self):
        if self.nv is not None:
            self.th = self._threading.Thread(target=self._run, daemon=True)
            self.th.start()

    def _run(self):
        nv = self.nv
        while not self._stop:
            t0 = time.perf_counter()
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM), self.max_mhz,
                                  nv.nvmlDeviceGetCurrentClocksEventReasons(self.h), nv.nvmlDeviceGetPowerUsage(self.h) / 1e3))
            except Exception:
                pass
            self.slowest = max(self.slowest, 1e3 * (time.perf_counter() - t0))
            time.sleep(self.period)

    def clear(self):
        self.rows.clear()
        self.slowest = 0.0

    def result(self):
        self._stop = True
        if self.th:
            self.th.join(timeout=1)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(k for k, bit in names.items() if any(r[2] & bit for r in self.rows))
        return {"sm_mhz": statistics.median(r[0] for r in self.rows), "sm_max_mhz": max(r[1] for r in self.rows),
                "reasons": reasons, "samples": len(self.rows), "power_w_max": max(r[3] for r in self.rows),
                "nvml_call_ms_max": round(self.slowest, 2)}


def pin_cores(local, world):
    """each rank keeps to its own slice of the host cores (8 ranks + their NCCL / CUDA helper threads otherwise migrate over
    each other's cores in the host-driven e2e region)"""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // max(world, 1)
        if world > 1 and per >= 2:
            mine = cores[local * per:(local + 1) * per]
            os.sched_setaffinity(0, mine)
            return len(mine)
        return len(cores)
    except Exception:
        return os.cpu_count() or 1


def _roofline(kernel, units, bytes_per_unit, ms, peak, bound="hbm", **more):
    ach = units * bytes_per_unit / (ms / 1e3) / 1e9
    return dict({"bound": bound, "kernel": kernel, "algorithmic_bytes_per_unit": bytes_per_unit, "achieved": ach, "peak": peak,
                 "unit": "GB/s", "frac": ach / peak}, **more)


def _timed(fn, reps):
    import torch
    ms = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    return ms, out


def _tree_depth(t, S):
    import torch
    depth = torch.zeros_like(t.parents)
    for node in range(1, S + 1):  # nodes are created in order, so a parent's depth is known
        par = t.parents[:, node].long().clamp(min=0)
        depth[:, node] = torch.where(t.parents[:, node] >= 0, depth.gather(1, par[:, None])[:, 0] + 1, 0)
    return float(depth[:, 1:].float().mean().item())


# ------------------------------------------------------------------------------------------------ stand-in networks
class StandInNet:
    """The networks are the caller's (Flax in the reference, exchanged through DLPack); the reference's DOG networks do not even
    exist (muzero_dog.py:85-99).  A minimal random-init latent-256 stand-in of the dynamics / prediction family keeps the bench
    about the path under test: one dense layer per function, every head written into a preallocated contiguous buffer."""

    def __init__(self, dev, obs_size, A, E, n, seed=0, chance=0):
        import torch
        g = torch.Generator(device=dev).manual_seed(seed)
        rnd = lambda *shape: torch.randn(*shape, device=dev, generator=g)
        self.A, self.E, self.Cn = A, E, chance
        self.Wr = rnd(obs_size, E) * (1.0 / obs_size ** 0.5)
        self.Wd, self.Wa = rnd(E, E) * (1.2 / E ** 0.5), rnd(A + chance, E)
        self.Wp, self.Wc = rnd(E, A) * (2.0 / E ** 0.5), rnd(E, max(chance, 1)) * (2.0 / E ** 0.5)
        self.n, self.ones = n, torch.ones(n, device=dev)
        self.after = torch.zeros((n, E + 2), device=dev)

    def root(self, params, obs):
        import torch
        from exploring_muzero_on_dog_b200 import mcts
        e = torch.tanh(obs.reshape(obs.shape[0], -1).float() @ self.Wr)
        return mcts.RootFnOutput(e @ self.Wp, torch.tanh(e[:, 0]), e)

    def _heads(self, e):
        import torch
        return 0.1 * e[:, 0], torch.copysign(self.ones, e[:, 1]), torch.tanh(e[:, 2])   # reward, discount (+-1), value

    def recurrent(self, params, rng, action, emb):
        import torch
        from exploring_muzero_on_dog_b200 import mcts
        e = torch.tanh(torch.addmm(self.Wa[action], emb, self.Wd))
        r, d, v = self._heads(e)
        return mcts.RecurrentFnOutput(r, d, e @ self.Wp, v), e

    # stochastic MuZero: afterstate embedding = [latent | reward | discount] like the reference (muzero_classic_madn.py:415-424)
    def decision(self, params, rng, action, emb):
        import torch
        from exploring_muzero_on_dog_b200 import mcts
        E = self.E
        after = self.after                                     # preallocated [n, E + 2]
        torch.tanh(torch.addmm(self.Wa[action], emb, self.Wd), out=after[:, :E])
        e = after[:, :E]
        after[:, E] = 0.1 * e[:, 0]
        torch.copysign(self.ones, e[:, 1], out=after[:, E + 1])
        return mcts.DecisionRecurrentFnOutput(e @ self.Wc, torch.tanh(e[:, 2])), after

    def chance(self, params, rng, outcome, after):
        import torch
        from exploring_muzero_on_dog_b200 import mcts
        E = self.E
        e = torch.tanh(torch.addmm(self.Wa[outcome + self.A], after[:, :E], self.Wd))
        return mcts.ChanceRecurrentFnOutput(e @ self.Wp, torch.tanh(e[:, 0]), after[:, E], after[:, E + 1]), e


def _gather_ranks(dev, world, my_ms, my_units):
    """-> (max ms over ranks, units summed over ranks, per-rank ms list)"""
    import torch
    import torch.distributed as dist
    if world == 1:
        return my_ms, my_units, [my_ms]
    t = torch.tensor([my_ms], dtype=torch.float64, device=dev)
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    c = torch.tensor([my_units], dtype=torch.int64, device=dev)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    per = [float(x.item()) for x in allt]
    return max(per), int(c.item()), per


class PrecomputedNet:
    """callbacks that launch NOTHING: they hand back preallocated random tensors (R sets, cycled).  The loop then costs what the
    path itself costs — env, observation, legal mask, tree kernels, trajectory rows — which is what `vs_tree_only` is about."""

    def __init__(self, dev, A, E, n, seed=0, chance=0, R=4):
        import torch
        g = torch.Generator(device=dev).manual_seed(seed)
        rnd = lambda *shape: torch.randn(*shape, device=dev, generator=g)
        self.R, self.k, self.A, self.E = R, 0, A, E
        self.prior, self.emb = [rnd(n, A) for _ in range(R)], [rnd(n, E) for _ in range(R)]
        self.val, self.rew = [torch.tanh(rnd(n)) for _ in range(R)], [0.1 * rnd(n) for _ in range(R)]
        self.disc = [torch.where(rnd(n) > 0, 1.0, -1.0) for _ in range(R)]
        self.chance = [rnd(n, max(chance, 1)) for _ in range(R)]
        self.after = [rnd(n, E + 2) for _ in range(R)]

    def _next(self):
        self.k = (self.k + 1) % self.R
        return self.k

    def root(self, params, obs):
        from exploring_muzero_on_dog_b200 import mcts
        return mcts.RootFnOutput(self.prior[0], self.val[0], self.emb[0])

    def recurrent(self, params, rng, action, emb):
        from exploring_muzero_on_dog_b200 import mcts
        k = self._next()
        return mcts.RecurrentFnOutput(self.rew[k], self.disc[k], self.prior[k], self.val[k]), self.emb[k]

    def decision(self, params, rng, action, emb):
        from exploring_muzero_on_dog_b200 import mcts
        k = self._next()
        return mcts.DecisionRecurrentFnOutput(self.chance[k], self.val[k]), self.after[k]

    def chance_fn(self, params, rng, outcome, after):
        from exploring_muzero_on_dog_b200 import mcts
        k = self.k
        return mcts.ChanceRecurrentFnOutput(self.prior[k], self.val[(k + 1) % self.R], self.rew[k], self.disc[k]), self.emb[k]


def make_selfplay(which, dev, rank, world, cuda_graph=True, plies=None, net="standin"):
    """-> (SelfPlayLoop, loop key, reseed()) for BASELINE config 3 or 5 on this rank's shard of the games"""
    import functools
    import torch
    from exploring_muzero_on_dog_b200 import game_agent, jaxrand, mcts
    from exploring_muzero_on_dog_b200.DOG import dog as dg
    from exploring_muzero_on_dog_b200.MADN import classic_madn as cm
    key = jaxrand.split_host(jaxrand.PRNGKey(0))[1]
    if which == "cfg3":
        n, S, A, Cn, E = 4096, 64, 4, 6, 256
        plies = plies or 128
        seeds = jaxrand.randint(key, n * world, 0, 1_000_000, device=dev)[rank * n:(rank + 1) * n].contiguous()
        reset = lambda out=None: cm.env_reset(0, seed=seeds, device=dev, out=out, **game_agent.STOCHASTIC_RULES)
        envs = reset()
        noise = torch.distributions.Dirichlet(torch.full((A,), 0.3, device=dev)).sample((n,))  # root noise sample: an input
        nn = StandInNet(dev, 11 * 56, A, E, n, seed=3, chance=Cn) if net == "standin" else PrecomputedNet(dev, A, E, n, seed=3, chance=Cn)
        dec, ch = (nn.decision, nn.chance) if net == "standin" else (nn.decision, nn.chance_fn)

        def search_fn(p, keys, obs, invalid):
            key2 = game_agent._split_each(keys, 1)
            out, rv = game_agent.run_stochastic_muzero_mcts(p, key2, obs, invalid, S, 50, 1.0, root_fn=nn.root, decision_recurrent_fn=dec,
                                                            chance_recurrent_fn=ch, dirichlet_noise=noise)
            return out.action, out.action_weights, rv

        loop = game_agent.SelfPlayLoop(envs, n, (11, 56), None, plies, search_fn=search_fn, obs_dtype=torch.int8, cuda_graph=cuda_graph)
    else:
        n, S, A, E = 8192, 100, 806, 256
        plies = plies or 64
        seeds = jaxrand.randint(key, n * world, 0, 1_000_000, device=dev)[rank * n:(rank + 1) * n].contiguous()
        reset = lambda out=None: dg.env_reset(0, seed=seeds, device=dev, out=out, **DOG_RULES)
        envs = reset()
        nn = StandInNet(dev, dg.RAW_OBS_SIZE, A, E, n, seed=5) if net == "standin" else PrecomputedNet(dev, A, E, n, seed=5)

        def search_fn(p, keys, obs, invalid):
            key2 = game_agent._split_each(keys, 1)
            out = mcts.gumbel_muzero_policy(p, key2, nn.root(p, obs), nn.recurrent, S, invalid_actions=invalid, max_depth=50,
                                            qtransform=functools.partial(mcts.qtransform_completed_by_mix_value, value_scale=0.5),
                                            gumbel_scale=1.0, max_num_considered_actions=16)
            return out.action, out.action_weights, out.search_tree.summary().value

        loop = game_agent.SelfPlayLoop(envs, n, (dg.RAW_OBS_SIZE,), None, plies, search_fn=search_fn, obs_dtype=torch.int8,
                                       cuda_graph=cuda_graph)
    loop.reseed = lambda: reset(out=envs)
    loop.shape = dict(n=n, S=S, A=A, E=E)
    return loop, key


def _run_selfplay(which, dev, rank, world, net, plies=None):
    import torch
    loop, key = make_selfplay(which, dev, rank, world, cuda_graph=True, plies=plies, net=net)
    cap = loop.max_steps
    loop.max_steps = 3
    loop.run(key)                                           # untimed: graph capture + first replays
    loop.max_steps = cap
    loop.reseed()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    buf = loop.run(key)
    e1.record()
    torch.cuda.synchronize()
    return loop, key, buf, e0.elapsed_time(e1)


def selfplay_cfg3(dev, rank, world, plies_cap=128):
    """BASELINE config 3 as a self-play loop (MuZero_Classic_MADN/game_agent_stochastic.py:52-244): 4,096 dice-MADN games per
    GPU, throw_die -> encode_board -> valid_action -> run_stochastic_muzero_mcts(64 sims, max_depth 50) -> env_step, whole
    iteration replayed as one CUDA graph, no host synchronisation per iteration."""
    out = {}
    for net in ("standin", "precomputed"):
        loop, key, buf, ms = _run_selfplay("cfg3", dev, rank, world, net, plies_cap)
        n, S, E = loop.shape["n"], loop.shape["S"], loop.shape["E"]
        plies = loop.enqueued
        max_ms, all_sims, per = _gather_ranks(dev, world, ms, n * plies * S)
        r = {"sims": all_sims, "ms": max_ms, "sims_per_s": all_sims / (max_ms / 1e3), "per_rank_ms": per, "iterations": plies,
             "iterations_needed": loop.iterations, "searched_moves_rank0": int((buf["mask"] > 0).sum().item()),
             "env_steps_rank0": int(buf["idx"].sum().item())}
        if net == "standin":
            out = dict({"workload": f"cfg3 self-play: 4,096 dice-MADN games per GPU x 64 sims per move (stochastic MuZero, A=4+6, latent {E}, "
                                    f"stand-in network), {plies} lockstep iterations (cap {plies_cap}); one CUDA graph per iteration, no host sync"},
                       **r, gpu_launches_per_iteration="1 graph replay (key split, throw_die, encode, mask, root net, init, 64 x (tree kernel + network), policy, agent step)")
        else:
            out["path_only"] = dict(r, note="same loop with callbacks that launch nothing (precomputed network outputs): the cost of the path itself")
        del loop, buf
    return out


def selfplay_cfg5(dev, rank, world, plies=64, with_exchange=True):
    """BASELINE config 5: 8,192 DOG games per GPU x 100 Gumbel simulations over 806 actions (latent 256), the det-MADN loop shape
    (MuZero_det_MADN/game_agent.py:50-192) on DOG/dog.py, then the replay shard: save, sample, and the all-gathered global batch."""
    import torch
    from exploring_muzero_on_dog_b200 import vec_replay_buffer
    from exploring_muzero_on_dog_b200.DOG import dog as dg
    out = {}
    for net in ("standin", "precomputed"):
        loop, key, traj, ms = _run_selfplay("cfg5", dev, rank, world, net, plies)
        n, S, A, E = (loop.shape[k] for k in ("n", "S", "A", "E"))
        max_ms, all_sims, per = _gather_ranks(dev, world, ms, n * loop.enqueued * S)
        r = {"sims": all_sims, "ms": max_ms, "sims_per_s": all_sims / (max_ms / 1e3), "per_rank_ms": per, "iterations": loop.enqueued,
             "env_steps_rank0": int(traj["idx"].sum().item())}
        if net == "standin":
            buf = vec_replay_buffer.VectorizedReplayBuffer(n, 128, 10, 50, obs_shape=(dg.RAW_OBS_SIZE,), action_dim=A, max_episode_length=plies,
                                                           device=dev, obs_dtype=torch.int8, prioritized=True, seed=rank)
            e1, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e1.record()
            buf.save_games_from_buffers(traj)
            batch = buf.sample_batch_prioritized(beta=0.5)
            if with_exchange and world > 1:
                batch = vec_replay_buffer.allgather_batch(batch)
            e2.record()
            torch.cuda.synchronize()
            rep_max, _, _ = _gather_ranks(dev, world, e1.elapsed_time(e2), 0)
            out = dict({"workload": f"cfg5 self-play: 8,192 DOG games per GPU x 100 sims per move (Gumbel MuZero, A=806, latent {E}, stand-in "
                                    f"network), {loop.enqueued} lockstep iterations; one CUDA graph per iteration; then replay save + prioritised "
                                    "sample(batch 128, unroll 10, td 50)" + (" + all-gather of the batch (NCCL)" if world > 1 else "")},
                       **r, replay_ms=rep_max, global_batch=int(batch["actions"].shape[0]), batch_keys=sorted(batch.keys()))
            del buf, batch
        else:
            out["path_only"] = dict(r, note="same loop with callbacks that launch nothing (precomputed network outputs): the cost of the path itself")
        del loop, traj
        torch.cuda.empty_cache()
    return out


def dog_cfg4(dev, rank, world, peak):
    import torch
    from exploring_muzero_on_dog_b200 import jaxrand
    from exploring_muzero_on_dog_b200.DOG import dog
    n = 16384
    key = jaxrand.split_host(jaxrand.PRNGKey(0))[1]
    all_seeds = jaxrand.randint(key, n * world, 0, 1_000_000, device=dev)
    seeds = all_seeds[rank * n:(rank + 1) * n].contiguous()
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    env = dog.env_reset(0, seed=seeds, device=dev, **DOG_RULES)

    def once():
        dog.env_reset(0, seed=seeds, device=dev, out=env, **DOG_RULES)
        total.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dog.play_random(env, key, max_steps=MAX_STEPS, game_offset=rank * n, total_steps=total)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    for _ in range(2):
        once()
    ms = once()
    steps = int(total.item())
    max_ms, all_steps, per = _gather_ranks(dev, world, ms, steps)
    return {"workload": "cfg4: DOG 2v2, 16,384 lockstep games per GPU, random legal policy (806 actions) to termination",
            "env_steps": all_steps, "kernel_ms": max_ms, "env_steps_per_s": all_steps / (max_ms / 1e3), "per_rank_ms": per,
            "roofline": _roofline("k_dog_play_random", all_steps / world, DOG_BYTES_PER_STEP, max_ms, peak, bound="issue",
                                  note="nominal: algorithmic bytes of the reference dataflow; the persistent kernel keeps a game in shared "
                                       "memory, its DRAM traffic is one read + one write of the state")}, env


# ------------------------------------------------------------------------------------------------ N = 1 side measurements
def tree_only(dev, key, peak):
    import torch
    from exploring_muzero_on_dog_b200 import jaxrand, mcts
    out = {}
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *shape: torch.randn(*shape, device=dev, generator=g)
    # config 3 shape
    n, S, A, Cn, E, R = 4096, 64, 4, 6, 258, 8
    cfg = mcts._cfg(mcts.STOCHASTIC, mcts.qtransform_by_parent_and_siblings, S, 50, A, Cn, E, dirichlet_fraction=0.0)
    srch = mcts.Search(cfg, n, dev)
    keys = jaxrand.split(key, n, device=dev)
    root = mcts.RootFnOutput(rnd(n, A), torch.zeros(n, device=dev), rnd(n, E))
    pl, cl, emb = [rnd(n, A) for _ in range(R)], [rnd(n, Cn) for _ in range(R)], [rnd(n, E) for _ in range(R)]
    val, rew = [torch.tanh(rnd(n)) for _ in range(R)], [0.1 * rnd(n) for _ in range(R)]
    disc = [torch.where(rnd(n) > 0, 1.0, -1.0) for _ in range(R)]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for rep in range(3):
        ev[0].record()
        srch.init(keys, root, None, None)
        ev[1].record()
        srch.select(0)
        for sim in range(S):
            k = sim % R
            (srch.expand_select if sim + 1 < S else srch.expand)(sim, pl[k], val[k], rew[k], disc[k], emb[k], cl[k], val[(k + 1) % R], emb[(k + 1) % R])
        ev[2].record()
        torch.cuda.synchronize()
    ms_init, ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    dbar = _tree_depth(srch.tree, S)
    bps = 2100 + 248 * dbar  # SURVEY.md 8(d) cfg 3
    out["mcts_cfg3_tree_only"] = {"workload": "cfg3 tree kernels alone (select + expand/backup, precomputed network outputs): 4,096 games x 64 sims, A=4+6, E=258",
                                  "sims": n * S, "ms": ms, "init_ms": ms_init, "sims_per_s": n * S / (ms / 1e3), "mean_expansion_depth": dbar,
                                  "gpu_launches": S + 1, "roofline": _roofline("k_mcts_expand_select", n * S, bps, ms, peak)}
    del srch
    # config 5 shape
    n, S, A, E, R = 8192, 100, 806, 256, 4
    cfg = mcts._cfg(mcts.GUMBEL, mcts.qtransform_completed_by_mix_value(value_scale=0.5), S, 50, A, 0, E)
    srch = mcts.Search(cfg, n, dev)
    keys = jaxrand.split(key, n, device=dev)
    root = mcts.RootFnOutput(rnd(n, A), torch.zeros(n, device=dev), rnd(n, E))
    pl, emb = [rnd(n, A) for _ in range(R)], [rnd(n, E) for _ in range(R)]
    val, rew = [torch.tanh(rnd(n)) for _ in range(R)], [0.1 * rnd(n) for _ in range(R)]
    disc = [torch.where(rnd(n) > 0, 1.0, -1.0) for _ in range(R)]
    for rep in range(2):
        ev[0].record()
        srch.init(keys, root, None, None)
        ev[1].record()
        srch.select(0)
        for sim in range(S):
            k = sim % R
            (srch.expand_select if sim + 1 < S else srch.expand)(sim, pl[k], val[k], rew[k], disc[k], emb[k])
        ev[2].record()
        torch.cuda.synchronize()
    ms_init, ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    t = srch.tree
    dbar = _tree_depth(t, S)
    bps = 5300 + 16200 * dbar  # SURVEY.md 8(d) cfg 5: five dense 806-wide child rows per visited level
    out["mcts_cfg5_tree_only"] = {
        "workload": "cfg5 tree kernels alone: Gumbel MuZero, 8,192 games x 100 sims, A=806, E=256, precomputed network outputs",
        "sims": n * S, "ms": ms, "init_ms": ms_init, "sims_per_s": n * S / (ms / 1e3), "mean_expansion_depth": dbar, "gpu_launches": S + 1,
        "tree_bytes": int(sum(getattr(t, k).numel() * getattr(t, k).element_size() for k in
                              ("children_index", "children_prior_logits", "children_visits", "children_rewards", "children_discounts",
                               "children_values", "embeddings"))),
        "roofline": _roofline("k_mcts_expand_select<8,2>", n * S, bps, ms, peak,
                              note="algorithmic bytes of the reference dataflow (5 dense child rows per level); the kernel reads one dense "
                                   "row per level plus the visited children (select cache), so its DRAM traffic is about a third of that")}
    del srch, t, pl, emb
    torch.cuda.empty_cache()
    return out


def side_measurements(dev, key, peak):
    """N = 1 only: the per-call path, the evaluation loop and config 1"""
    import numpy as np
    import torch
    from exploring_muzero_on_dog_b200 import evaluate_agent as ea, jaxrand, mcts
    from exploring_muzero_on_dog_b200.MADN import deterministic_madn as dm
    from exploring_muzero_on_dog_b200.TicTacToe import mcts as tm
    out = {}
    # config 2 driven per call: one fused lockstep iteration (legal mask + categorical draw + env_step / no_step) per launch,
    # i.e. what a jitted loop body that calls the drop-in functions every iteration pays; 842 iterations = the longest game
    n2, iters = 65536, 842
    seeds2 = jaxrand.randint(key, n2, 0, 1_000_000, device=dev)
    env2 = dm.env_reset(0, seed=seeds2, device=dev, **RULES)
    act = torch.zeros(1, dtype=torch.int64, device=dev)
    for rep in range(2):
        dm.env_reset(0, seed=seeds2, device=dev, out=env2, **RULES)
        act.zero_()
        k = jaxrand.KeyChain(key)          # rng_key, *step_keys = split(rng_key, N + 1): element 0, advanced on the host
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(iters):
            dm.random_step(env2, k, active_count=act)
            k.advance()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    steps2 = int(act.item())
    out["madn_cfg2_per_call"] = {"workload": "cfg2 through one launch per lockstep iteration (k_madn_det_random_step), 65,536 games, 842 iterations",
                                 "env_steps": steps2, "ms": ms, "env_steps_per_s": steps2 / (ms / 1e3), "gpu_launches": iters,
                                 "all_done": bool(env2.raw("done").all()),
                                 "roofline": _roofline("k_madn_det_random_step", steps2, BYTES_PER_STEP, ms, peak)}
    # the same host loop with 32 lockstep iterations per launch (random_steps: the state stays in registers inside a chunk)
    chunk = 32
    tot2 = torch.zeros(1, dtype=torch.int64, device=dev)
    for rep in range(2):
        dm.env_reset(0, seed=seeds2, device=dev, out=env2, **RULES)
        tot2.zero_()
        k = np.asarray(key, dtype=np.uint32)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(0, iters + chunk - 1, chunk):
            k = dm.random_steps(env2, k, chunk, total_steps=tot2)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    steps3 = int(tot2.item())
    out["madn_cfg2_chunked"] = {"workload": f"cfg2 driven from the host {chunk} lockstep iterations per launch (random_steps), 65,536 games",
                                "env_steps": steps3, "ms": ms, "env_steps_per_s": steps3 / (ms / 1e3), "gpu_launches": (iters + chunk - 1) // chunk,
                                "all_done": bool(env2.raw("done").all()), "same_steps_as_per_call": steps3 == steps2,
                                "roofline": _roofline("k_madn_det_play_cta", steps3, BYTES_PER_STEP, ms, peak, bound="issue")}
    # evaluation loop: rule-based team against random team, one fused launch per lockstep iteration
    n3 = 16384
    seeds3 = jaxrand.randint(key, n3, 0, 1_000_000, device=dev)
    for rep in range(2):
        env3 = dm.env_reset(0, seed=seeds3, device=dev, **RULES)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        env3, win3 = ea.play_eval_loop(env3, ({"type": 2}, {"type": 3}, {"type": 2}, {"type": 3}), key, n3, poll_every=64)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out["eval_loop_rule_vs_random"] = {"workload": "play_eval_loop: 16,384 det-MADN games, seats 0/2 rule-based scorer, seats 1/3 random, to termination "
                                                   "(k_madn_eval_step per lockstep iteration)",
                                       "ms": ms, "games_per_s": n3 / (ms / 1e3), "team_0_2_wins": int(win3[:, 0].sum().item()),
                                       "team_1_3_wins": int(win3[:, 1].sum().item())}
    # config 1: TicTacToeV2, 512 lockstep games x 50 simulations per ply, true-env callbacks with rollout, PUCT (TicTacToe/mcts.py:9-23)
    cache = mcts.GraphCache()
    for rep in range(3):  # the per-call path: root_fn / init / 50 x (select, recurrent_fn, expand) launches per move, as one CUDA graph
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _, plies = tm.play_mcts_games(512, jaxrand.PRNGKey(rep), num_simulations=50, limit=30, variant=1, device=dev, graph_cache=cache)
        e1.record()
        torch.cuda.synchronize()
    ms_calls = e0.elapsed_time(e1)
    fused = {}
    for rep in range(3):  # the whole search of a move as ONE launch, one game per warp (dogstep_ttt_search)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _, plies = tm.play_mcts_games(512, jaxrand.PRNGKey(rep), num_simulations=50, limit=30, variant=1, device=dev, fused=fused)
        e1.record()
        torch.cuda.synchronize()
    ms, moves = e0.elapsed_time(e1), int(plies.sum().item())
    out["ttt_cfg1"] = {"workload": "cfg1: TicTacToeV2 self-play, 512 lockstep games x 50 sims per ply (muzero_policy on the true env with rollouts), "
                                   "to termination; each move's whole search is one launch, one game per warp (k_ttt_search)",
                       "per_call_path_ms": ms_calls, "per_call_path_sims_per_s": moves * 50 / (ms_calls / 1e3),
                       "searched_moves": moves, "ms": ms, "sims_per_s": moves * 50 / (ms / 1e3), "env_steps_per_s": moves / (ms / 1e3),
                       "roofline": _roofline("k_ttt_search", moves * 50, 800, ms, peak, bound="latency",
                                             note="SURVEY 8(d) cfg 1: 0.8 KB per simulation at depth 3.  512 games are 512 warps on 148 SMs and "
                                                  "every expansion runs a random rollout whose key chain is serial (one Threefry per rollout "
                                                  "ply): bound by the latency of one game's instruction stream, not by bytes")}
    return out


# ------------------------------------------------------------------------------------------------ CPU baselines (the oracle = the checker, timed)
def cpu_baseline(games, nthreads, key, seeds_np):
    """the oracle port on the host cores: same workload, `games` games"""
    import oracle as O
    from exploring_muzero_on_dog_b200 import rules as R
    cfg = O.MadnCfg(4, 0xF, 10, R.to_mask(RULES))
    t0 = time.perf_counter()
    s = O.madn_reset(cfg, seeds_np[:games], 0)
    glen, total, _ = O.madn_det_play_random(s, key, MAX_STEPS, nthreads=nthreads)
    dt = time.perf_counter() - t0
    return total, dt, s, glen


def cpu_baselines_extras(key, cores, dog_ctx=None):
    """bounded samples of the other configurations on the oracle (kind "port"), a few seconds each"""
    import numpy as np
    import oracle as O
    from exploring_muzero_on_dog_b200 import rules as R
    out = {}
    # cfg 4: DOG random play, all host threads, ALL 16,384 games: the CPU baseline and the full-size leaf-for-leaf parity check
    n = 16384
    seeds = O.randint(key, n, 0, 1_000_000)
    s = O.dog_reset(O.DogCfg(4, 0xF, 10, R.to_mask(DOG_RULES)), seeds, 0)
    t0 = time.perf_counter()
    _, total, _ = O.dog_play_random(s, key, MAX_STEPS, nthreads=cores)
    dt = time.perf_counter() - t0
    out["dog_cfg4"] = {"value": total / dt, "unit": "env_steps/s", "cores": cores, "kind": "port",
                       "sample": f"all {n} games played to termination on the C oracle ({total} env steps, {dt:.1f} s)"}
    if dog_ctx is not None:
        got = dog_ctx.numpy()
        out["dog_cfg4"]["gpu_result_equals_oracle"] = all(
            np.array_equal(np.asarray(v).astype(np.int64), got[k].astype(np.int64)) for k, v in s.fields().items())
    # tree kernels alone on the oracle (single thread: the oracle's tree code is scalar C)
    rng = np.random.default_rng(0)
    for name, d, n in (("mcts_cfg3_tree_only", dict(policy=2, qtransform=1, num_simulations=64, max_depth=50, num_actions=4, num_chance=6, embed_dim=258), 512),
                       ("mcts_cfg5_tree_only", dict(policy=1, qtransform=2, num_simulations=100, max_depth=50, num_actions=806, num_chance=0, embed_dim=256), 96)):
        d.update(max_num_considered_actions=16, q_min=-1.0, q_max=1.0, value_scale=0.5, maxvisit_init=50.0, epsilon=1e-8, pb_c_init=1.25,
                 pb_c_base=19652.0, dirichlet_fraction=0.0, temperature=1.0, gumbel_scale=1.0)
        cfg = O.MctsCfg(**d)
        A, Cn, E, S = d["num_actions"], d["num_chance"], d["embed_dim"], d["num_simulations"]
        f = lambda *shape: rng.standard_normal(shape).astype(np.float32)
        pri, emb, cl = f(n, A), f(n, E), f(n, max(Cn, 1))
        val, rew, disc = np.tanh(f(n)), 0.1 * f(n), np.where(f(n) > 0, 1.0, -1.0).astype(np.float32)
        tree = O.MctsTree(cfg, n)
        keys = rng.integers(0, 2**32, (n, 2), dtype=np.uint64).astype(np.uint32)
        O.mcts_init(tree, keys, pri, np.zeros(n, np.float32), emb)
        t0 = time.perf_counter()
        for sim in range(S):
            p, a, _, _ = O.mcts_select(tree, sim)
            if Cn:
                O.mcts_expand(tree, sim, p, a, pri, val, rew, disc, emb, cl, val, emb)
            else:
                O.mcts_expand(tree, sim, p, a, pri, val, rew, disc, emb)
        dt = time.perf_counter() - t0
        out[name] = {"value": n * S / dt, "unit": "sims/s", "cores": 1, "kind": "port",
                     "sample": f"{n} games x {S} sims on the oracle's tree code with precomputed network outputs ({dt:.1f} s)"}
    # cfg 1: TicTacToeV2 lockstep self-play with the oracle's search and true-env callbacks (single thread)
    n, S, moves = 128, 50, 0
    st = O.TttState(n, 1)
    d = dict(policy=0, qtransform=0, num_simulations=S, max_depth=9, num_actions=9, num_chance=0, embed_dim=18, max_num_considered_actions=16,
             q_min=-1.0, q_max=1.0, value_scale=0.1, maxvisit_init=50.0, epsilon=1e-8, pb_c_init=1.25, pb_c_base=19652.0,
             dirichlet_fraction=0.0, temperature=1.0, gumbel_scale=1.0)
    t0 = time.perf_counter()
    for ply in range(30):
        live = st.done == 0
        if not live.any():
            break
        keys = rng.integers(0, 2**32, (n, 2), dtype=np.uint64).astype(np.uint32)
        pri, val, emb = O.ttt_root_fn(st, keys)
        tree = O.MctsTree(O.MctsCfg(**d), n)
        O.mcts_init(tree, keys, pri, val, emb)
        for sim in range(S):
            p, a, e, _ = O.mcts_select(tree, sim)
            rp, rv, rr, rd, re = O.ttt_recurrent_fn(1, tree.expand_key, a, e)
            O.mcts_expand(tree, sim, p, a, rp, rv, rr, rd, re)
        act, _, _ = O.mcts_policy_output(tree)
        nxt = st.copy()
        O.ttt_step(nxt, act.astype(np.int8))
        for k in st.FIELDS:
            getattr(st, k)[live] = getattr(nxt, k)[live]
        moves += int(live.sum())
    dt = time.perf_counter() - t0
    out["ttt_cfg1"] = {"value": moves * S / dt, "unit": "sims/s", "cores": 1, "kind": "port",
                       "sample": f"{n} lockstep TicTacToeV2 games x {S} sims per ply to termination on the oracle ({moves} searched moves, {dt:.1f} s)"}
    # evaluation loop: rule-based team against random team on the NumPy / C oracle
    from oracle import eval_oracle
    n = 256
    cfgm = O.MadnCfg(4, 0xF, 10, R.to_mask(RULES))
    s = O.madn_reset(cfgm, O.randint(key, 16384, 0, 1_000_000)[:n], 0)
    t0 = time.perf_counter()
    eval_oracle.play_eval_loop(s, (2, 3, 2, 3), key)
    dt = time.perf_counter() - t0
    out["eval_loop_rule_vs_random"] = {"value": n / dt, "unit": "games/s", "cores": 1, "kind": "port",
                                       "sample": f"first {n} of the 16,384 games on oracle/eval_oracle.py, to termination ({dt:.1f} s)"}
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  JAX is not installable here (no wheels, no network), so
    this times the oracle port of MADN/deterministic_madn.py + the do_random driver with every host thread, on the SAME
    configuration as the product arm: all `--games` games of a step, every step."""
    if rank != 0:
        return
    import oracle as O
    from exploring_muzero_on_dog_b200 import jaxrand
    cores = os.cpu_count() or 1
    key = jaxrand.split_host(jaxrand.PRNGKey(0))[1]
    seeds = O.randint(key, args.games, 0, 1_000_000)
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(min(args.games, 2048), cores, key, seeds)
    tot, dt, per = 0, 0.0, []
    for _ in range(args.steps):
        a, b, _, _ = cpu_baseline(args.games, cores, key, seeds)
        tot += a
        dt += b
        per.append(round(1e3 * b, 1))
    v = tot / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "env_steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": "cfg2: deterministic MADN 4 players, 65,536 lockstep games per GPU, random legal policy to termination",
                       "games_per_gpu": args.games, "max_steps": MAX_STEPS, "rules": "MuZero_det_MADN/game_agent.py:12-22",
                       "env_steps_per_pass_per_gpu": tot // args.steps},
            "cpu_baseline": {"value": v, "unit": "env_steps/s", "cores": cores, "kind": "port",
                             "sample": f"all {args.games} games of a step played to termination on the C oracle (oracle/madn_oracle.c), "
                                       f"{args.steps} steps, {cores} threads; the reference's own JAX path cannot run here (no jax wheel)",
                             "step_ms": per},
            "e2e": {"value": v, "unit": "env_steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--games", type=int, default=65536, help="lockstep games per GPU")
    ap.add_argument("--impl", default="dogstep", choices=["dogstep", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the self-play (sims/s), DOG and side measurements")
    ap.add_argument("--extras-only", default="", help="comma list of extras to run (selfplay_cfg3, selfplay_cfg5, dog_cfg4, tree, side)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist
    from exploring_muzero_on_dog_b200 import jaxrand
    from exploring_muzero_on_dog_b200.MADN import deterministic_madn as dm

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    my_cores = pin_cores(local, world)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.backends.cuda.matmul.allow_tf32 = True  # the stand-in networks of the self-play extras (the caller's side of the path)
    if world > 1:
        # NCCL may print its version banner on stdout while the communicator is created: keep stdout for the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    n = args.games
    offset = rank * n                                    # rank r owns global games [r*n, (r+1)*n)
    key = jaxrand.split_host(jaxrand.PRNGKey(0))[1]      # subkey = split(PRNGKey(0))[1]  (game_agent.py:187)
    all_seeds = jaxrand.randint(key, n * world, 0, 1_000_000, device=dev)  # seeds = randint(subkey,(N,),0,1e6) (:188)
    seeds = all_seeds[offset:offset + n].contiguous()
    seeds_host = seeds.cpu().pin_memory()
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    glen = torch.empty(n, dtype=torch.int32, device=dev)
    env = dm.env_reset(0, seed=seeds, **RULES, device=dev)   # the one set of leaves every step re-seeds in place
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    drain = torch.zeros(64 << 20, dtype=torch.int32, device=dev)   # 256 MiB read after the write: evicts the dirty lines

    def one_step(seed_t):
        dm.env_reset(0, seed=seed_t, **RULES, device=dev, out=env)
        dm.play_random(env, key, max_steps=MAX_STEPS, game_offset=offset, game_len=glen, total_steps=total)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    clocks.start()  # before the warm-up: NVML initialisation and the first (slow) queries stay out of the timed regions
    for _ in range(max(args.warmup, 3)):
        one_step(seeds)
    barrier()

    # ---- timed region 1: inputs resident in HBM; L2 flushed between steps; per-kernel events for the roofline
    total.zero_()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    clocks.clear()  # keep only the samples taken during the two timed regions
    # the whole region is enqueued behind a ~20 ms spin kernel, so the device runs the K steps back to back from its queue and
    # a host hiccup (GC, another tenant on the box) cannot leave it idle between a reset and its play kernel
    torch.cuda._sleep(40_000_000)
    for s in range(args.steps):
        flush.fill_(s & 0xFF)
        drain.max()  # read sweep: the flush's write-backs finish before the timed step instead of inside it
        ev[s][0].record()
        dm.env_reset(0, seed=seeds, **RULES, device=dev, out=env)
        ev[s][1].record()
        dm.play_random(env, key, max_steps=MAX_STEPS, game_offset=offset, game_len=glen, total_steps=total)
        ev[s][2].record()
    barrier()
    step_ms = [ev[s][0].elapsed_time(ev[s][2]) for s in range(args.steps)]
    play_ms = [ev[s][1].elapsed_time(ev[s][2]) for s in range(args.steps)]
    my_ms = sum(step_ms)
    my_steps = int(total.item())
    final_state = {k: v.cpu().numpy() for k, v in env._t.items()}   # for the full-size parity check against the oracle below
    final_glen = glen.cpu().numpy()

    # ---- timed region 2 (e2e): host seeds in pinned memory -> H2D -> public API -> D2H of the results
    shapes = {"game_len": ((n,), torch.int32), "reward": ((n,), torch.int8), "done": ((n,), torch.bool), "pins": ((n, 4, 4), torch.int8)}
    # double-buffered: the caller consumes step s while step s + 1 is already enqueued (a host sync per step would leave the
    # GPU idle for the launch latency of the next step); every step still pays its own H2D and D2H copies
    res_host = [{k: torch.empty(sh, dtype=d).pin_memory() for k, (sh, d) in shapes.items()} for _ in range(2)]
    res_dev = [{k: torch.empty(sh, dtype=d, device=dev) for k, (sh, d) in shapes.items()} for _ in range(2)]
    seeds_dev = [torch.empty_like(seeds) for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    staged = [torch.cuda.Event() for _ in range(2)]

    def e2e_step(b):
        """H2D seeds -> reset + play (in place) -> results staged on the device -> D2H on the copy stream, so that the next
        step's kernels do not queue behind this step's PCIe transfers"""
        seeds_dev[b].copy_(seeds_host, non_blocking=True)
        one_step(seeds_dev[b])
        res_dev[b]["game_len"].copy_(glen)
        res_dev[b]["reward"].copy_(env.raw("reward"))
        res_dev[b]["done"].copy_(env.raw("done"))
        res_dev[b]["pins"].copy_(env.raw("pins"))
        staged[b].record()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(staged[b])
            for k in shapes:
                res_host[b][k].copy_(res_dev[b][k], non_blocking=True)
            consumed[b].record()

    for b in range(2):  # untimed: first use of the pinned buffers and copy paths
        e2e_step(b)
    torch.cuda.synchronize()
    total.zero_()
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    e2e_ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    host_t = []
    t0.record()
    for s in range(args.steps):
        b = s & 1
        h0 = time.perf_counter()
        if s >= 2:
            consumed[b].synchronize()  # results of step s - 2 have arrived on the host before their buffers are reused
        e2e_step(b)
        e2e_ev[s].record()
        host_t.append(1e3 * (time.perf_counter() - h0))
    torch.cuda.current_stream().wait_stream(copy_stream)
    t1.record()
    torch.cuda.synchronize()
    barrier()
    clk = clocks.result()
    e2e_ms = t0.elapsed_time(t1)
    e2e_step_ms = [(t0 if s == 0 else e2e_ev[s - 1]).elapsed_time(e2e_ev[s]) for s in range(args.steps)]
    e2e_steps = int(total.item())
    assert all(bool(r["done"].all()) for r in res_host[:min(2, args.steps)]), "games did not terminate"
    h2d = seeds_host.numel() * 4
    d2h = sum(t.numel() * t.element_size() for t in res_host[0].values())

    if world > 1:
        t = torch.tensor([my_ms, e2e_ms], dtype=torch.float64, device=dev)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        c = torch.tensor([my_steps, e2e_steps], dtype=torch.int64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        per_rank_ms = [float(x[0].item()) for x in allt]
        per_rank_e2e = [float(x[1].item()) for x in allt]
        max_ms, e2e_max_ms = max(per_rank_ms), max(per_rank_e2e)
        all_steps, all_e2e_steps = c.tolist()
    else:
        max_ms, e2e_max_ms, all_steps, all_e2e_steps = my_ms, e2e_ms, my_steps, e2e_steps
        per_rank_ms, per_rank_e2e = [my_ms], [e2e_ms]

    peak, peak_src = _peaks()
    extras, dog_ctx = {}, None
    only = set(x for x in args.extras_only.split(",") if x)
    want = lambda name: not args.no_extras and (not only or name in only)
    del flush, drain
    torch.cuda.empty_cache()
    if want("selfplay_cfg3"):
        extras["selfplay_cfg3"] = selfplay_cfg3(dev, rank, world)
    if want("selfplay_cfg5"):
        extras["selfplay_cfg5"] = selfplay_cfg5(dev, rank, world)
        torch.cuda.empty_cache()
    if want("dog_cfg4"):
        extras["dog_cfg4"], dog_ctx = dog_cfg4(dev, rank, world, peak)
    if world == 1 and want("tree"):
        extras.update(tree_only(dev, key, peak))
        for cfgk, loopk in (("mcts_cfg3_tree_only", "selfplay_cfg3"), ("mcts_cfg5_tree_only", "selfplay_cfg5")):
            if loopk in extras:
                extras[loopk]["vs_tree_only"] = extras[loopk]["sims_per_s"] / extras[cfgk]["sims_per_s"]
                extras[loopk]["path_only"]["vs_tree_only"] = extras[loopk]["path_only"]["sims_per_s"] / extras[cfgk]["sims_per_s"]
    if world == 1 and want("side"):
        extras.update(side_measurements(dev, key, peak))

    if rank == 0:
        play_s = sum(play_ms) / 1e3
        achieved = my_steps * BYTES_PER_STEP / play_s / 1e9
        facts = _profile_facts()
        traffic = facts.get("k_madn_det_play_cta_dram_bytes_per_launch")
        inst = facts.get("k_madn_det_play_cta_warp_instructions_per_launch")
        sm_hz = (clk.get("sm_mhz") or 1965.0) * 1e6
        issue = (inst / (SMS * SCHEDULERS * sm_hz * (play_s / args.steps))) if inst else None
        line = {
            "metric": METRIC, "value": all_steps / (max_ms / 1e3), "unit": "env_steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": max_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": "cfg2: deterministic MADN 4 players, 65,536 lockstep games per GPU, random legal policy to termination",
                       "games_per_gpu": n, "max_steps": MAX_STEPS, "rules": "MuZero_det_MADN/game_agent.py:12-22",
                       "env_steps_per_pass_per_gpu": my_steps // args.steps, "l2": "flushed between timed steps (256 MiB write, then a 256 MiB read sweep so that no write-back of the flush is pending)",
                       "parallelism": f"games sharded x{world}, no collective on the stepping path", "host_cores_per_rank": my_cores},
            "e2e": {"value": all_e2e_steps / (e2e_max_ms / 1e3), "unit": "env_steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_max_ms / args.steps, "per_rank_ms": [round(x, 3) for x in per_rank_e2e],
                    "rank0_step_ms": [round(x, 3) for x in e2e_step_ms], "rank0_host_ms": [round(x, 3) for x in host_t]},
            "per_rank_ms": [round(x, 3) for x in per_rank_ms],
            "gpu_launches": 2 * args.steps,
            "roofline": {"bound": "issue", "kernel": "k_madn_det_play_cta", "achieved": achieved, "peak": peak,
                         "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "algorithmic_bytes_per_env_step": BYTES_PER_STEP, "kernel_ms_per_launch": sum(play_ms) / args.steps,
                         "issue_slot_frac": issue, "warp_instructions_per_launch": inst,
                         "step_ms": [round(x, 4) for x in step_ms], "kernel_ms": [round(x, 4) for x in play_ms],
                         "note": "achieved / peak / frac: NOMINAL HBM figure on the algorithmic bytes of the per-step reference dataflow (SURVEY 8d). "
                                 "The persistent kernel keeps a game in registers (DRAM traffic = one read + one write of the state), so what "
                                 "bounds it is the integer issue rate: issue_slot_frac = warp instructions per launch (ncu, profiles/) / "
                                 "(148 SMs x 4 schedulers x SM clock x kernel time)"},
            "clocks": clk,
        }
        if extras:
            line["extras"] = extras
        if world == 1 and not args.no_cpu_baseline:
            import oracle as O
            cores = os.cpu_count() or 1
            seeds_np = seeds.cpu().numpy()
            # all 65,536 games on the oracle: the CPU baseline AND the full-size leaf-for-leaf parity check of this very run
            tot, dt, s, olen = cpu_baseline(n, cores, key, seeds_np)
            same = bool(np.array_equal(olen, final_glen)) and all(
                np.array_equal(np.asarray(v).astype(np.int64), final_state[k].astype(np.int64)) for k, v in s.fields().items())
            line["cpu_baseline"] = {"value": tot / dt, "unit": "env_steps/s", "cores": cores, "kind": "port",
                                    "sample": f"all {n} games of one step played to termination on the C oracle ({tot} env steps, {dt:.1f} s)",
                                    "gpu_result_equals_oracle": same}
            line["full_size_parity"] = {"cfg2_65536_games_leaf_for_leaf": same}
            if not args.no_extras and not only:
                base = cpu_baselines_extras(key, cores, dog_ctx)
                for k, v in base.items():
                    if k in extras:
                        extras[k]["cpu_baseline"] = v
                if "gpu_result_equals_oracle" in base.get("dog_cfg4", {}):
                    line["full_size_parity"]["cfg4_16384_games_leaf_for_leaf"] = base["dog_cfg4"]["gpu_result_equals_oracle"]
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
