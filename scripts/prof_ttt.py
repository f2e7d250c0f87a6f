"""BASELINE config 1 timing: TicTacToe, 512 lockstep games x 50 simulations per ply (profiling helper)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exploring_muzero_on_dog_b200 import jaxrand, mcts
from exploring_muzero_on_dog_b200.TicTacToe import mcts as tm
n, S = 512, 50
for name, cache in (("eager", None), ("graph", mcts.GraphCache())):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        env, plies = tm.play_mcts_games(n, jaxrand.PRNGKey(rep), num_simulations=S, limit=30, variant=1, graph_cache=cache)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        moves = int(plies.sum().item())
        print(f"{name}: {moves} searched moves x {S} sims in {dt*1e3:.1f} ms -> {moves*S/dt/1e6:.2f} M sims/s, {moves/dt/1e3:.1f} k env steps/s")
