#!/usr/bin/env python
"""Kernel-time breakdown of BASELINE config 1 (TicTacToeV2 self-play, 512 games x 50 sims per ply), eager launches, through
torch.profiler.  python scripts/prof_ttt.py [games]"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from exploring_muzero_on_dog_b200 import jaxrand, _lib  # noqa: E402
if os.environ.get("DOGSTEP_LIB"):
    _lib.LIB_PATH = os.environ["DOGSTEP_LIB"]  # instrumented build (scripts/build_trace_lib.sh)
from exploring_muzero_on_dog_b200.TicTacToe import mcts as tm  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
kw = {"fused": {}} if len(sys.argv) > 2 and sys.argv[2] == "fused" else {}   # fused: one launch per move (k_ttt_search)
tm.play_mcts_games(n, jaxrand.PRNGKey(0), num_simulations=50, limit=30, variant=1, **kw)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _, plies = tm.play_mcts_games(n, jaxrand.PRNGKey(1), num_simulations=50, limit=30, variant=1, **kw)
    e1.record()
    torch.cuda.synchronize()
print(f"wall (events) {e0.elapsed_time(e1):.1f} ms")
tot, cnt = collections.Counter(), collections.Counter()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        tot[e.name] += e.device_time_total
        cnt[e.name] += 1
total = sum(tot.values())
moves = int(plies.sum().item())
print(f"{n} games, {moves} searched moves, {int(plies.max())} plies, {total / 1e3:.1f} ms of kernel time, {moves * 50 / (total / 1e6) / 1e6:.2f} M sims/s on kernel time")
for name, us in tot.most_common(12):
    print(f"{us / 1e3:9.2f} ms  {cnt[name]:6d} launches  {us / cnt[name]:8.1f} us each  {name[:100]}")
