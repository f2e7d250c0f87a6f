#!/usr/bin/env python
"""bench.py — BASELINE.json metric on its config 2 workload.

A "step" is one whole pass of the hot path over one batch: reset 65,536 deterministic-MADN games per GPU from
device-resident seeds, then play every game to termination with the reference's random legal policy
(MuZero_det_MADN/evaluate_agent.py:733-930 do_random: valid_action -> categorical -> env_step / no_step, cap 2000).
`value` = env steps/s (active (game, iteration) pairs) summed over all GPUs / max-over-ranks device time.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--games G] [--impl reference]
N > 1 is launched by torchrun (one rank per GPU); games are sharded with no data-path collective (weak scaling).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RULES = dict(enable_teams=True, enable_initial_free_pin=True, enable_circular_board=False, enable_friendly_fire=False,
             enable_start_blocking=False, enable_jump_in_goal_area=True, enable_start_on_1=True,
             enable_bonus_turn_on_6=True, must_traverse_start=False)  # MuZero_det_MADN/game_agent.py:12-22
BYTES_PER_STEP = 226  # SURVEY.md 8(d) cfg 2: 99 B state read + 99 B written + action 2 + mask 24 + reward/done 2
MAX_STEPS = 2000      # evaluate_agent.py:918
METRIC = "env steps/s (deterministic MADN, random legal policy to termination)"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """SM clock / throttle reasons sampled through NVML DURING the timed region (same fields as the
    nvidia-smi line of B200_PROFILING.md; in-process so that a sub-second region still gets samples)."""

    def __init__(self, index, period=0.002):
        self.index, self.period, self.rows, self._stop, self.th = index, period, [], False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        except Exception:
            self.nv = None

    def start(self):
        if self.nv is None:
            return
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def _run(self):
        nv = self.nv
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1e3
                self.rows.append((sm, mx, rs, pw))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
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
                "reasons": reasons, "samples": len(self.rows), "power_w_max": max(r[3] for r in self.rows)}



def measure_extras(dev, key):
    """Side measurements on one GPU, reported under "extras" (not the headline): BASELINE config 4 (DOG, 16,384 lockstep
    games, random legal policy over 806 actions to termination) and config 3's tree work (stochastic MuZero search,
    4,096 games x 64 simulations, A = 4 + 6) with a minimal stand-in network, so the number is the tree kernels' rate."""
    import torch
    from exploring_muzero_on_dog_b200 import jaxrand, mcts
    from exploring_muzero_on_dog_b200.DOG import dog
    peak, _ = _peaks()
    out = {}
    DOG_RULES = dict(enable_teams=True, enable_initial_free_pin=False, enable_circular_board=True, enable_friendly_fire=True,
                     enable_start_blocking=True, enable_jump_in_goal_area=False, must_traverse_start=True)  # MuZero_DOG/game_agent.py:12-23
    n = 16384
    seeds = jaxrand.randint(key, n, 0, 1_000_000, device=dev)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    for rep in range(3):
        env = dog.env_reset(0, seed=seeds, device=dev, **DOG_RULES)
        total.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dog.play_random(env, key, max_steps=MAX_STEPS, total_steps=total)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    steps = int(total.item())
    out["dog_cfg4"] = {"workload": "cfg4: DOG 2v2, 16,384 lockstep games, random legal policy (806 actions) to termination",
                       "env_steps": steps, "kernel_ms": ms, "env_steps_per_s": steps / (ms / 1e3),
                       "roofline": {"bound": "hbm", "kernel": "k_dog_play_random", "algorithmic_bytes_per_env_step": 1228,
                                    "achieved": steps * 1228 / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                                    "frac": steps * 1228 / (ms / 1e3) / 1e9 / peak}}
    # config 3 tree work
    n, S, A, Cn, E = 4096, 64, 4, 6, 258
    g = torch.Generator(device=dev).manual_seed(0)
    Wd = torch.randn(E, E, device=dev, generator=g) * 0.05
    Wp, Wc = torch.randn(E, A, device=dev, generator=g), torch.randn(E, Cn, device=dev, generator=g)

    def dec(params, rng, action, emb):
        e = torch.tanh(emb @ Wd)
        return mcts.DecisionRecurrentFnOutput(e @ Wc, torch.tanh(e[:, 0])), e

    def ch(params, rng, outcome, emb):
        e = torch.tanh(emb @ Wd)
        return mcts.ChanceRecurrentFnOutput(e @ Wp, torch.tanh(e[:, 0]), 0.1 * e[:, 1], torch.where(e[:, 2] > 0, 1.0, -1.0)), e

    root = mcts.RootFnOutput(torch.randn(n, A, device=dev, generator=g), torch.zeros(n, device=dev), torch.randn(n, E, device=dev, generator=g))
    keys = jaxrand.split(key, n, device=dev)
    invalid = torch.zeros(n, A, dtype=torch.bool, device=dev)
    cache = mcts.GraphCache()  # the whole search (init, 64 x (select, both networks, expand), policy) replayed as one CUDA graph
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        po = mcts.stochastic_muzero_policy(None, keys, root, dec, ch, S, invalid_actions=invalid, max_depth=50, dirichlet_fraction=0.0,
                                           graph_cache=cache)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = po.search_tree
    depth = torch.zeros_like(t.parents)
    for node in range(1, S + 1):  # nodes are created in order, so a parent's depth is known
        par = t.parents[:, node].long().clamp(min=0)
        depth[:, node] = torch.where(t.parents[:, node] >= 0, depth.gather(1, par[:, None])[:, 0] + 1, 0)
    dbar = float(depth[:, 1:].float().mean().item())
    bytes_per_sim = 2100 + 248 * dbar  # SURVEY.md 8(d) cfg 3
    out["mcts_cfg3"] = {"workload": "cfg3 tree work: stochastic MuZero search, 4,096 games x 64 sims, A=4+6, E=258, stand-in network",
                        "sims": n * S, "ms": ms, "sims_per_s": n * S / (ms / 1e3), "mean_expansion_depth": dbar,
                        "roofline": {"bound": "hbm", "kernel": "k_mcts_select + k_mcts_expand", "algorithmic_bytes_per_sim": bytes_per_sim,
                                     "achieved": n * S * bytes_per_sim / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                                     "frac": n * S * bytes_per_sim / (ms / 1e3) / 1e9 / peak,
                                     "note": "device time of the whole search incl. the stand-in network (3 small matmuls per simulation), replayed as one CUDA graph"}}
    # the same search with precomputed network outputs: the tree kernels alone (select + expand/backup)
    cfg = mcts._cfg(mcts.STOCHASTIC, mcts.qtransform_by_parent_and_siblings, S, 50, A, Cn, E, dirichlet_fraction=0.0)
    srch = mcts.Search(cfg, n, dev)
    R = 8
    rnd = lambda *shape: torch.randn(*shape, device=dev, generator=g)
    pl, cl, emb = [rnd(n, A) for _ in range(R)], [rnd(n, Cn) for _ in range(R)], [rnd(n, E) for _ in range(R)]
    val, rew = [torch.tanh(rnd(n)) for _ in range(R)], [0.1 * rnd(n) for _ in range(R)]
    disc = [torch.where(rnd(n) > 0, 1.0, -1.0) for _ in range(R)]
    for rep in range(2):
        srch.init(keys, root, None, None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        srch.select(0)
        for sim in range(S):
            k = sim % R
            step = srch.expand_select if sim + 1 < S else srch.expand  # expand(sim) + select(sim + 1): one launch
            step(sim, pl[k], val[k], rew[k], disc[k], emb[k], cl[k], val[(k + 1) % R], emb[(k + 1) % R])
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out["mcts_cfg3_tree_only"] = {"workload": "cfg3 tree kernels alone (select + expand/backup, precomputed network outputs)",
                                  "sims": n * S, "ms": ms, "sims_per_s": n * S / (ms / 1e3), "gpu_launches": S + 1,
                                  "roofline": {"bound": "hbm", "kernel": "k_mcts_expand_select",
                                               "algorithmic_bytes_per_sim": bytes_per_sim, "achieved": n * S * bytes_per_sim / (ms / 1e3) / 1e9,
                                               "peak": peak, "unit": "GB/s", "frac": n * S * bytes_per_sim / (ms / 1e3) / 1e9 / peak}}
    out.update(measure_cfg5(dev, key, peak))
    # config 2 driven per call: one fused lockstep iteration (legal mask + categorical draw + env_step / no_step) per launch,
    # i.e. what a jitted loop body that calls the drop-in functions every iteration pays; 842 iterations = the longest game
    from exploring_muzero_on_dog_b200.MADN import deterministic_madn as dm
    n2 = 65536
    seeds2 = jaxrand.randint(key, n2, 0, 1_000_000, device=dev)
    iters = 842
    for rep in range(2):
        env2 = dm.env_reset(0, seed=seeds2, device=dev, **RULES)
        act = torch.zeros(1, dtype=torch.int64, device=dev)
        import numpy as np
        k = np.asarray(key, dtype=np.uint32)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(iters):
            dm.random_step(env2, k, active_count=act)
            k = jaxrand.split_host(k)[0]
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    steps2 = int(act.item())
    out["madn_cfg2_per_call"] = {"workload": "cfg2 through one launch per lockstep iteration (k_madn_det_random_step), 65,536 games, 842 iterations",
                                 "env_steps": steps2, "ms": ms, "env_steps_per_s": steps2 / (ms / 1e3), "gpu_launches": iters,
                                 "all_done": bool(env2.raw("done").all()),
                                 "roofline": {"bound": "hbm", "kernel": "k_madn_det_random_step", "algorithmic_bytes_per_env_step": BYTES_PER_STEP,
                                              "achieved": steps2 * BYTES_PER_STEP / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                                              "frac": steps2 * BYTES_PER_STEP / (ms / 1e3) / 1e9 / peak}}
    # evaluation loop: rule-based team against random team, one fused launch per lockstep iteration
    from exploring_muzero_on_dog_b200 import evaluate_agent as ea
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
    from exploring_muzero_on_dog_b200.TicTacToe import mcts as tm
    cache = mcts.GraphCache()
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _, plies = tm.play_mcts_games(512, jaxrand.PRNGKey(rep), num_simulations=50, limit=30, variant=1, device=dev, graph_cache=cache)
        e1.record()
        torch.cuda.synchronize()
    ms, moves = e0.elapsed_time(e1), int(plies.sum().item())
    out["ttt_cfg1"] = {"workload": "cfg1: TicTacToeV2 self-play, 512 lockstep games x 50 sims per ply (muzero_policy on the true env with rollouts), "
                                   "to termination; each ply's search replayed as one CUDA graph",
                       "searched_moves": moves, "ms": ms, "sims_per_s": moves * 50 / (ms / 1e3), "env_steps_per_s": moves / (ms / 1e3)}
    return out


def measure_replay_exchange(dev, rank, world):
    """N > 1 only: the one collective of the design — every rank samples a batch from its own replay shard and the batches are
    all-gathered over NCCL (VectorizedReplayBuffer.sample_batch_global).  Off the stepping path; reported as a side number."""
    import torch
    import torch.distributed as dist
    from exploring_muzero_on_dog_b200 import vec_replay_buffer
    n, T, A = 2048, 64, 24
    g = torch.Generator(device=dev).manual_seed(rank)
    traj = dict(obs=torch.randint(-1, 4, (n, T, 34, 56), device=dev, generator=g, dtype=torch.int8),
                act=torch.randint(0, A, (n, T), device=dev, generator=g, dtype=torch.int32),
                rew=torch.randint(0, 3, (n, T), device=dev, generator=g, dtype=torch.int32),
                val=torch.rand(n, T, device=dev, generator=g), pol=torch.rand(n, T, A, device=dev, generator=g),
                mask=torch.ones(n, T, device=dev), player=torch.randint(0, 4, (n, T), device=dev, generator=g, dtype=torch.int32),
                team=torch.randint(0, 2, (n, T), device=dev, generator=g, dtype=torch.int32),
                discount=torch.randint(0, 3, (n, T), device=dev, generator=g, dtype=torch.int32),
                idx=torch.full((n,), T, device=dev, dtype=torch.int32))
    buf = vec_replay_buffer.VectorizedReplayBuffer(n, 128, 10, 50, obs_shape=(34, 56), action_dim=A, max_episode_length=T, device=dev,
                                                   obs_dtype=torch.int8, seed=rank)
    buf.save_games_from_buffers(traj)
    ms = []
    for rep in range(6):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        batch = buf.sample_batch_global()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = torch.tensor([sorted(ms[1:])[len(ms[1:]) // 2]], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    nbytes = sum(v.numel() * v.element_size() for v in batch.values())
    return {"replay_exchange": {"workload": "sample_batch(128, unroll 10, td 50) on every rank + NCCL all-gather of the batch leaves",
                                "ms": float(t.item()), "global_batch": int(batch["actions"].shape[0]), "gathered_bytes_per_rank": int(nbytes),
                                "collective": "one ncclAllGather of the %d packed leaves" % len(batch)}}


def measure_cfg5(dev, key, peak):
    """BASELINE config 5 (8,192 DOG games per GPU x 100 Gumbel simulations over 806 actions, latent 256): the tree kernels alone
    with precomputed network outputs, and a short slice of the whole self-play loop (DOG env + search with a random-init
    stand-in network + trajectory rows + replay save / sample) — the reference has no DOG networks (muzero_dog.py:85-99)."""
    import torch
    from exploring_muzero_on_dog_b200 import game_agent, jaxrand, mcts, vec_replay_buffer
    out = {}
    n, S, A, E, R = 8192, 100, 806, 256, 4
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *shape: torch.randn(*shape, device=dev, generator=g)
    cfg = mcts._cfg(mcts.GUMBEL, mcts.qtransform_completed_by_mix_value(value_scale=0.5), S, 50, A, 0, E)
    srch = mcts.Search(cfg, n, dev)
    keys = jaxrand.split(key, n, device=dev)
    root = mcts.RootFnOutput(rnd(n, A), torch.zeros(n, device=dev), rnd(n, E))
    pl, emb = [rnd(n, A) for _ in range(R)], [rnd(n, E) for _ in range(R)]
    val, rew = [torch.tanh(rnd(n)) for _ in range(R)], [0.1 * rnd(n) for _ in range(R)]
    disc = [torch.where(rnd(n) > 0, 1.0, -1.0) for _ in range(R)]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
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
    depth = torch.zeros_like(t.parents)
    for node in range(1, S + 1):
        par = t.parents[:, node].long().clamp(min=0)
        depth[:, node] = torch.where(t.parents[:, node] >= 0, depth.gather(1, par[:, None])[:, 0] + 1, 0)
    dbar = float(depth[:, 1:].float().mean().item())
    bps = 5300 + 16200 * dbar  # SURVEY.md 8(d) cfg 5: five dense 806-wide child rows per visited level
    out["mcts_cfg5_tree_only"] = {
        "workload": "cfg5 tree kernels alone: Gumbel MuZero, 8,192 games x 100 sims, A=806, E=256, precomputed network outputs",
        "sims": n * S, "ms": ms, "init_ms": ms_init, "sims_per_s": n * S / (ms / 1e3), "mean_expansion_depth": dbar, "gpu_launches": S + 1,
        "tree_bytes": int(sum(getattr(t, k).numel() * getattr(t, k).element_size() for k in
                              ("children_index", "children_prior_logits", "children_visits", "children_rewards", "children_discounts",
                               "children_values", "embeddings"))),
        "roofline": {"bound": "hbm", "kernel": "k_mcts_expand_select<8,2>", "algorithmic_bytes_per_sim": bps,
                     "achieved": n * S * bps / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s", "frac": n * S * bps / (ms / 1e3) / 1e9 / peak,
                     "note": "algorithmic bytes of the reference dataflow (5 dense child rows per level); the kernel reads one dense "
                             "row per level plus the visited children (select cache), so its DRAM traffic is about a third of that"}}
    del srch, t, depth, pl, emb
    torch.cuda.empty_cache()
    # a slice of the whole loop
    plies = 4
    Wr, Wp, Wv = rnd(74, E) * 0.2, rnd(E, A), rnd(E)
    Wa, Wd = rnd(A, E), rnd(E, E) * 0.06

    def root_fn(params, obs):
        e = torch.tanh(obs.reshape(obs.shape[0], -1) @ Wr)
        return mcts.RootFnOutput(e @ Wp, torch.tanh(e @ Wv), e)

    def recurrent_fn(params, rng, action, e0):
        e = torch.tanh(e0 @ Wd + Wa[action])
        return mcts.RecurrentFnOutput(0.1 * e[:, 0], torch.where(e[:, 1] > 0, 1.0, -1.0), e @ Wp, torch.tanh(e @ Wv)), e

    buf = vec_replay_buffer.VectorizedReplayBuffer(n, 128, 10, 50, obs_shape=(74,), action_dim=A, max_episode_length=64, device=dev,
                                                   obs_dtype=torch.int8)
    cache = mcts.GraphCache()  # each ply's search (init, 100 x (tree kernel, network), policy) replayed as one CUDA graph
    for rep in range(2):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        envs, traj = game_agent.play_n_dog_games(None, jaxrand.PRNGKey(rep), n, S, 50, plies, 1.0, root_fn=root_fn, recurrent_fn=recurrent_fn,
                                                 device=dev, graph_cache=cache)
        e1.record()
        buf.save_games_from_buffers(traj)
        batch = buf.sample_batch()
        e2.record()
        torch.cuda.synchronize()
    ms_play, ms_rep = e0.elapsed_time(e1), e1.elapsed_time(e2)
    out["selfplay_cfg5_slice"] = {
        "workload": f"cfg5 loop slice: 8,192 DOG games x {plies} plies x 100 sims (env + Gumbel search + stand-in latent-256 network + "
                    "trajectory rows), then replay save + sample(batch 128, unroll 10, td 50)",
        "plies": plies, "play_ms": ms_play, "replay_ms": ms_rep, "sims_per_s": n * plies * S / (ms_play / 1e3),
        "env_steps_per_s": n * plies / (ms_play / 1e3), "batch_keys": sorted(batch.keys())}
    return out


def cpu_baseline(games, nthreads, key, seeds_np):
    """the oracle port on the host cores: same workload, bounded sample of `games` games"""
    import oracle as O
    from exploring_muzero_on_dog_b200 import rules as R
    cfg = O.MadnCfg(4, 0xF, 10, R.to_mask(RULES))
    t0 = time.perf_counter()
    s = O.madn_reset(cfg, seeds_np[:games], 0)
    _, total, _ = O.madn_det_play_random(s, key, MAX_STEPS, nthreads=nthreads)
    dt = time.perf_counter() - t0
    return total, dt


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  JAX is not installable here (no wheels,
    no network), so this times the oracle port of MADN/deterministic_madn.py with every host thread."""
    if rank != 0:
        return
    import numpy as np
    import oracle as O
    from exploring_muzero_on_dog_b200 import jaxrand
    cores = os.cpu_count() or 1
    key = jaxrand.split_host(jaxrand.PRNGKey(0))[1]
    sample = min(args.games, 4096)
    seeds = O.randint(key, sample, 0, 1_000_000)
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(min(sample, 512), cores, key, seeds)
    tot, dt = 0, 0.0
    for _ in range(args.steps):
        a, b = cpu_baseline(sample, cores, key, seeds)
        tot += a
        dt += b
    v = tot / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "env_steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": "cfg2: deterministic MADN 4 players, random legal policy to termination",
                       "games_per_step": sample, "max_steps": MAX_STEPS, "rules": "MuZero_det_MADN/game_agent.py:12-22"},
            "cpu_baseline": {"value": v, "unit": "env_steps/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} of {args.games} games per step, played to termination, {args.steps} steps"},
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
    ap.add_argument("--no-extras", action="store_true", help="skip the DOG (cfg 4) and MCTS (cfg 3) side measurements")
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
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
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
    env = dm.env_reset(0, seed=seeds, **RULES, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    drain = torch.zeros(64 << 20, dtype=torch.int32, device=dev)   # 256 MiB read after the write: evicts the dirty lines

    def one_step(seed_t):
        e = dm.env_reset(0, seed=seed_t, **RULES, device=dev)
        dm.play_random(e, key, max_steps=MAX_STEPS, game_offset=offset, game_len=glen, total_steps=total)
        return e

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    clocks.start()  # before the warm-up: NVML initialisation and thread start-up stay out of the timed regions
    for _ in range(max(args.warmup, 3)):
        e = one_step(seeds)  # bound like in the timed loops: the allocator then holds the two generations of leaves they cycle through
    barrier()

    # ---- timed region 1: inputs resident in HBM; L2 flushed between steps; per-kernel events for the roofline
    total.zero_()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    clocks.rows.clear()  # keep only the samples taken during the two timed regions
    # the whole region is enqueued behind a ~20 ms spin kernel, so the device runs the K steps back to back from its queue and
    # a host hiccup (GC, another tenant on the box) cannot leave it idle between a reset and its play kernel
    torch.cuda._sleep(40_000_000)
    for s in range(args.steps):
        flush.fill_(s & 0xFF)
        drain.max()  # read sweep: the flush's write-backs finish before the timed step instead of inside it
        ev[s][0].record()
        e = dm.env_reset(0, seed=seeds, **RULES, device=dev)
        ev[s][1].record()
        dm.play_random(e, key, max_steps=MAX_STEPS, game_offset=offset, game_len=glen, total_steps=total)
        ev[s][2].record()
    barrier()
    step_ms = [ev[s][0].elapsed_time(ev[s][2]) for s in range(args.steps)]
    play_ms = [ev[s][1].elapsed_time(ev[s][2]) for s in range(args.steps)]
    my_ms = sum(step_ms)
    my_steps = int(total.item())

    # ---- timed region 2 (e2e): host seeds in pinned memory -> H2D -> public API -> D2H of the results
    shapes = {"game_len": ((n,), torch.int32), "reward": ((n,), torch.int8), "done": ((n,), torch.bool), "pins": ((n, 4, 4), torch.int8)}
    # double-buffered: the caller consumes step s while step s + 1 is already enqueued (a host sync per step would leave the
    # GPU idle for the launch latency of the next step); every step still pays its own H2D and D2H copies
    res_host = [{k: torch.empty(sh, dtype=d).pin_memory() for k, (sh, d) in shapes.items()} for _ in range(2)]
    seeds_dev = [torch.empty_like(seeds) for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    for b in range(2):  # untimed: first use of the pinned buffers and copy paths
        seeds_dev[b].copy_(seeds_host, non_blocking=True)
        e = one_step(seeds_dev[b])
        for k, src in (("game_len", glen), ("reward", e.raw("reward")), ("done", e.raw("done")), ("pins", e.raw("pins"))):
            res_host[b][k].copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    total.zero_()
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for s in range(args.steps):
        b = s & 1
        if s >= 2:
            consumed[b].synchronize()  # results of step s - 2 have been read before their buffers are reused
        seeds_dev[b].copy_(seeds_host, non_blocking=True)
        e = one_step(seeds_dev[b])
        res_host[b]["game_len"].copy_(glen, non_blocking=True)
        res_host[b]["reward"].copy_(e.raw("reward"), non_blocking=True)
        res_host[b]["done"].copy_(e.raw("done"), non_blocking=True)
        res_host[b]["pins"].copy_(e.raw("pins"), non_blocking=True)
        consumed[b].record()
    torch.cuda.current_stream().synchronize()
    t1.record()
    barrier()
    clk = clocks.stop()
    e2e_ms = t0.elapsed_time(t1)
    e2e_steps = int(total.item())
    assert all(bool(r["done"].all()) for r in res_host[:min(2, args.steps)]), "games did not terminate"
    h2d = seeds_host.numel() * 4
    d2h = sum(t.numel() * t.element_size() for t in res_host[0].values())

    if world > 1:
        t = torch.tensor([my_ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        c = torch.tensor([my_steps, e2e_steps], dtype=torch.int64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        max_ms, e2e_max_ms = t.tolist()
        all_steps, all_e2e_steps = c.tolist()
    else:
        max_ms, e2e_max_ms, all_steps, all_e2e_steps = my_ms, e2e_ms, my_steps, e2e_steps

    extras = None
    if world == 1 and not args.no_extras:
        extras = measure_extras(dev, key)
    elif world > 1 and not args.no_extras:
        extras = measure_replay_exchange(dev, rank, world)

    if rank == 0:
        peak, peak_src = _peaks()
        play_s = sum(play_ms) / 1e3
        achieved = my_steps * BYTES_PER_STEP / play_s / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                traffic = json.load(f).get("k_madn_det_play_cta_dram_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": all_steps / (max_ms / 1e3), "unit": "env_steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": max_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": "cfg2: deterministic MADN 4 players, 65,536 lockstep games per GPU, random legal policy to termination",
                       "games_per_gpu": n, "max_steps": MAX_STEPS, "rules": "MuZero_det_MADN/game_agent.py:12-22",
                       "env_steps_per_pass_per_gpu": my_steps // args.steps, "l2": "flushed between timed steps (256 MiB write, then a 256 MiB read sweep so that no write-back of the flush is pending)",
                       "parallelism": f"games sharded x{world}, no collective on the stepping path"},
            "e2e": {"value": all_e2e_steps / (e2e_max_ms / 1e3), "unit": "env_steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_max_ms / args.steps},
            "gpu_launches": 2 * args.steps,
            "roofline": {"bound": "hbm", "kernel": "k_madn_det_play_cta", "achieved": achieved, "peak": peak,
                         "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "algorithmic_bytes_per_env_step": BYTES_PER_STEP, "kernel_ms_per_launch": sum(play_ms) / args.steps,
                         "step_ms": [round(x, 4) for x in step_ms], "kernel_ms": [round(x, 4) for x in play_ms],
                         "note": "algorithmic bytes of the per-step reference dataflow; the persistent kernel keeps a game in registers (DRAM traffic = one read + one write of the state), so it is integer-issue bound, not HBM bound"},
            "clocks": clk,
        }
        if extras:
            line["extras"] = extras
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            seeds_np = seeds.cpu().numpy()
            est_steps, est_dt = cpu_baseline(256, cores, key, seeds_np)
            rate = est_steps / est_dt
            sample = int(min(n, max(256, (12.0 * rate) / (est_steps / 256))))
            tot, dt = cpu_baseline(sample, cores, key, seeds_np)
            line["cpu_baseline"] = {"value": tot / dt, "unit": "env_steps/s", "cores": cores, "kind": "port",
                                    "sample": f"first {sample} of the {n} games, played to termination once ({tot} env steps, {dt:.1f} s)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
