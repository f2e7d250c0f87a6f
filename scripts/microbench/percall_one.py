"""One lockstep iteration per launch (k_madn_det_random_step) on 65,536 games: 40 launches, for an ncu capture of one of them."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from exploring_muzero_on_dog_b200 import jaxrand
from exploring_muzero_on_dog_b200.MADN import deterministic_madn as dm
import bench
n = 65536
key = jaxrand.split_host(jaxrand.PRNGKey(0))[1]
env = dm.env_reset(0, seed=jaxrand.randint(key, n, 0, 1_000_000), **bench.RULES)
k = key
for _ in range(40):
    dm.random_step(env, k); k = jaxrand.split_host(k, 1)[0]
torch.cuda.synchronize()
