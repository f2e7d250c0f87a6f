"""small DOG play_random run for ncu (profiling helper, not part of the product)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exploring_muzero_on_dog_b200 import jaxrand, _lib
if os.environ.get('DOGSTEP_LIB'):
    _lib.LIB_PATH = os.environ['DOGSTEP_LIB']  # instrumented build (scripts/build_trace_lib.sh)
from exploring_muzero_on_dog_b200.DOG import dog
R = dict(enable_teams=True, enable_initial_free_pin=False, enable_circular_board=True, enable_friendly_fire=True,
         enable_start_blocking=True, enable_jump_in_goal_area=False, must_traverse_start=True)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 150
key = jaxrand.split_host(jaxrand.PRNGKey(0))[1]
seeds = jaxrand.randint(key, n, 0, 1_000_000)
for rep in range(2):
    env = dog.env_reset(0, seed=seeds, **R)
    tot = torch.zeros(1, dtype=torch.int64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dog.play_random(env, key, max_steps=steps, total_steps=tot); e1.record(); torch.cuda.synchronize()
print("steps", int(tot.item()), "ms", e0.elapsed_time(e1), "Msteps/s", tot.item() / e0.elapsed_time(e1) / 1e3)
