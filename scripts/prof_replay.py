"""replay save / plan / gather and the fused agent-step kernel at config-5-like sizes, for ncu / quick timing (profiling helper)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exploring_muzero_on_dog_b200 import game_agent, jaxrand, vec_replay_buffer
from exploring_muzero_on_dog_b200.MADN import deterministic_madn as dm
dev = torch.device("cuda")
n, T, A = 8192, 64, 24
key = jaxrand.split_host(jaxrand.PRNGKey(0))[1]
seeds = jaxrand.randint(key, n, 0, 1_000_000)
envs = dm.env_reset(0, seed=seeds, **game_agent.RULES)
traj = game_agent.Trajectories(n, T, (34, 56), A, False, dev, torch.int8)
g = torch.Generator(device=dev).manual_seed(0)
ev = lambda: torch.cuda.Event(enable_timing=True)
t_step = []
for t in range(T):
    obs = dm.encode_board(envs)
    valid = dm.valid_action(envs).reshape(n, -1)
    score = torch.rand(n, A, device=dev, generator=g).masked_fill(~valid, -1.0)
    action = score.argmax(1).to(torch.int32)
    w = torch.softmax(score * 4, 1)
    val = torch.rand(n, device=dev, generator=g)
    e0, e1 = ev(), ev()
    e0.record(); game_agent.agent_step(envs, traj, action, val, w, obs); e1.record()
    t_step.append((e0, e1))
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in t_step)[len(t_step) // 2]
row = 34 * 56 + 24 * 4 + 7 * 4 + 99 * 2 + 34 * 56
print(f"agent_step: {ms*1e3:.1f} us per iteration of {n} games  ({n*row/ms/1e6:.0f} GB/s on {row} B per game)")
buf = vec_replay_buffer.VectorizedReplayBuffer(20000, 128, 10, 50, obs_shape=(34, 56), action_dim=A, max_episode_length=T, device=dev,
                                               obs_dtype=torch.int8)
for rep in range(3):
    e0, e1, e2 = ev(), ev(), ev()
    e0.record(); buf.save_games_from_buffers(traj.as_dict()); e1.record(); b = buf.sample_batch(); e2.record()
    torch.cuda.synchronize()
ep_bytes = T * (34 * 56 + A * 4 + 7 * 4)
print(f"replay save: {e0.elapsed_time(e1):.3f} ms for {n} episodes x {T} plies ({2*n*ep_bytes/e0.elapsed_time(e1)/1e6:.0f} GB/s read+write)   "
      f"sample_batch(128, unroll 10, td 50): {e1.elapsed_time(e2):.3f} ms")
