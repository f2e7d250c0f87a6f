"""tree-kernel-only MCTS run (select + expand with precomputed network outputs) for timing / ncu (profiling helper)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exploring_muzero_on_dog_b200 import jaxrand, mcts, _lib
if os.environ.get("DOGSTEP_LIB"):
    _lib.LIB_PATH = os.environ["DOGSTEP_LIB"]  # a variant build (scripts/microbench)
shape = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
key = jaxrand.split_host(jaxrand.PRNGKey(0))[1]
if shape == "cfg3":   # stochastic MuZero, 4096 games x 64 sims, A = 4 + 6, E = 258
    n, S, A, Cn, E, policy, qt = 4096, 64, 4, 6, 258, mcts.STOCHASTIC, mcts.qtransform_by_parent_and_siblings
elif shape == "cfg5":  # gumbel MuZero on DOG's 806 actions, 100 sims, latent 256 (games reduced by argv[2])
    n, S, A, Cn, E, policy, qt = int(sys.argv[2]) if len(sys.argv) > 2 else 2048, 100, 806, 0, 256, mcts.GUMBEL, mcts.qtransform_completed_by_mix_value(value_scale=0.5)
else:                  # det MADN gumbel: 24 actions, latent 256
    n, S, A, Cn, E, policy, qt = 8192, 100, 24, 0, 256, mcts.GUMBEL, mcts.qtransform_completed_by_mix_value(value_scale=0.5)
cfg = mcts._cfg(policy, qt, S, 50, A, Cn, E, dirichlet_fraction=0.0)
s = mcts.Search(cfg, n, dev)
keys = jaxrand.split(key, n, device=dev)
root = mcts.RootFnOutput(torch.randn(n, A, device=dev, generator=g), torch.zeros(n, device=dev), torch.randn(n, E, device=dev, generator=g))
R = 8  # a few distinct precomputed "network outputs"
pl = [torch.randn(n, A, device=dev, generator=g) for _ in range(R)]
cl = [torch.randn(n, max(Cn, 1), device=dev, generator=g) for _ in range(R)]
val = [torch.tanh(torch.randn(n, device=dev, generator=g)) for _ in range(R)]
rew = [0.1 * torch.randn(n, device=dev, generator=g) for _ in range(R)]
disc = [torch.where(torch.randn(n, device=dev, generator=g) > 0, 1.0, -1.0) for _ in range(R)]
emb = [torch.randn(n, E, device=dev, generator=g) for _ in range(R)]
for rep in range(3):
    s.init(keys, root, None, None)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * S + 1)]
    ev[0].record()
    fused = os.environ.get("FUSED", "1") == "1"
    if fused:
        s.select(0)
    for sim in range(S):
        if not fused:
            s.select(sim)
        ev[2 * sim + 1].record()
        k = sim % R
        step = s.expand_select if (fused and sim + 1 < S) else s.expand
        if policy == mcts.STOCHASTIC:
            step(sim, pl[k], val[k], rew[k], disc[k], emb[k], cl[k], val[(k + 1) % R], emb[(k + 1) % R])
        else:
            step(sim, pl[k], val[k], rew[k], disc[k], emb[k])
        ev[2 * sim + 2].record()
    torch.cuda.synchronize()
    tot = ev[0].elapsed_time(ev[-1])
    print(f"{shape} n={n} S={S} A'={A+Cn} fused={int(fused)}: total {tot:.2f} ms -> {n*S/tot/1e3:.2f} M sims/s (tree kernels only)")
if os.environ.get("DOGSTEP_LIB") and hasattr(_lib.lib(), "dogstep_trace_wide_decided"):
    import ctypes
    c3 = (ctypes.c_ulonglong * 5)()
    _lib.lib().dogstep_trace_wide_decided(c3)
    tot = max(1, c3[3] + c3[4])
    print(f"wide interior levels: decided without the row {c3[3] / tot:.3f}; with the row: from bounds {c3[1] / tot:.3f}, exact evaluation "
          f"{c3[0] / tot:.3f}, two near-maximal children {c3[2] / tot:.3f} (of {tot})")

