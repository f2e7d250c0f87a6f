import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from exploring_muzero_on_dog_b200 import jaxrand
from exploring_muzero_on_dog_b200.MADN import deterministic_madn as dm
R = dict(enable_teams=True, enable_initial_free_pin=True, enable_circular_board=False, enable_friendly_fire=False,
         enable_start_blocking=False, enable_jump_in_goal_area=True, enable_start_on_1=True,
         enable_bonus_turn_on_6=True, must_traverse_start=False)
n = 65536
key = jaxrand.split_host(jaxrand.PRNGKey(0))[1]
seeds = jaxrand.randint(key, n, 0, 1_000_000)
env = dm.env_reset(0, seed=seeds, **R)
for it in (1, 2, 4, 8, 32):
    e = env.clone()
    k = key
    for _ in range(5): k = dm.random_steps(e, k, it)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 64 // it if it < 64 else 1
    e0.record()
    for _ in range(reps): k = dm.random_steps(e, k, it)
    e1.record(); torch.cuda.synchronize()
    print("random_steps(%d): %.1f us per launch, %.2f us per iteration" % (it, e0.elapsed_time(e1) * 1e3 / reps, e0.elapsed_time(e1) * 1e3 / reps / it))
e = env.clone(); k = key
for _ in range(20):
    dm.random_step(e, k); k = jaxrand.split_host(k, 1)[0]
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    dm.random_step(e, k); k = jaxrand.split_host(k, 1)[0]
e1.record(); torch.cuda.synchronize()
print("random_step: %.2f us per call (host-driven)" % (e0.elapsed_time(e1) * 1e3 / 200))
g = torch.cuda.CUDAGraph()
ks = [k]
for _ in range(50): ks.append(jaxrand.split_host(ks[-1], 1)[0])
with torch.cuda.graph(g):
    for j in range(50): dm.random_step(e, ks[j])
torch.cuda.synchronize()
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print("random_step in a graph of 50: %.2f us per call" % (e0.elapsed_time(e1) * 1e3 / 50))
gl = torch.empty(n, dtype=torch.int32, device="cuda")
for it in (1, 2, 4):
    e = env.clone()
    g = torch.cuda.CUDAGraph()
    dm.play_random(e, ks[0], max_steps=it, game_len=gl); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for j in range(50): dm.play_random(e, ks[j], max_steps=it, game_len=gl)
    torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    print("play_random(max_steps=%d) in a graph of 50: %.2f us per launch" % (it, e0.elapsed_time(e1) * 1e3 / 50))
