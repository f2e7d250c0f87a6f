"""det-MADN play_random run for ncu / quick timing (profiling helper, not part of the product)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exploring_muzero_on_dog_b200 import jaxrand, _lib
if os.environ.get('DOGSTEP_LIB'):
    _lib.LIB_PATH = os.environ['DOGSTEP_LIB']  # instrumented build (scripts/microbench)
from exploring_muzero_on_dog_b200.MADN import deterministic_madn as dm
R = dict(enable_teams=True, enable_initial_free_pin=True, enable_circular_board=False, enable_friendly_fire=False,
         enable_start_blocking=False, enable_jump_in_goal_area=True, enable_start_on_1=True,
         enable_bonus_turn_on_6=True, must_traverse_start=False)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
key = jaxrand.split_host(jaxrand.PRNGKey(0))[1]
seeds = jaxrand.randint(key, n, 0, 1_000_000)
for rep in range(reps):
    env = dm.env_reset(0, seed=seeds, **R)
    tot = torch.zeros(1, dtype=torch.int64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env2, glen = dm.play_random(env, key, max_steps=2000, total_steps=tot); e1.record(); torch.cuda.synchronize()
    if os.environ.get('DOGSTEP_PLAY_TRACE') is not None:  # trace the CTA that holds the longest game next
        per = -(-n // 148)
        os.environ['DOGSTEP_PLAY_TRACE'] = str(int(glen.argmax().item()) // per)
        print("longest game", int(glen.max().item()), "in CTA", os.environ['DOGSTEP_PLAY_TRACE'])
    print("steps", int(tot.item()), "ms", e0.elapsed_time(e1), "Gsteps/s", tot.item() / e0.elapsed_time(e1) / 1e6)
