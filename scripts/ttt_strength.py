#!/usr/bin/env python
"""Playing strength of the true-env TicTacToeV2 search on the GPU next to what the reference recorded with mctx
(TicTacToe/results.md:12-15 search bot against the random bot, :53-68 search against search; 1,000 games each there).

    python scripts/ttt_strength.py [games]          # default 10,000 games per row
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from exploring_muzero_on_dog_b200 import jaxrand  # noqa: E402
from exploring_muzero_on_dog_b200.TicTacToe import mcts as tm  # noqa: E402

VS_RANDOM = {5: (97.1, 2.7, 0.2), 10: (98.3, 1.7, 0.0), 30: (99.1, 0.9, 0.0), 100: (99.4, 0.6, 0.0)}
SELFPLAY = {("run_mcts", 5): (48.0, 47.2, 4.8), ("run_mcts", 10): (50.4, 46.4, 3.2), ("run_mcts", 30): (58.1, 40.7, 1.2),
            ("run_mcts", 100): (64.1, 20.3, 15.6), ("run_gumbel", 5): (52.0, 40.5, 7.5), ("run_gumbel", 10): (54.7, 40.3, 5.0),
            ("run_gumbel", 30): (52.2, 46.7, 1.1), ("run_gumbel", 100): (57.7, 41.5, 0.8)}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    fused = {}
    print(f"search bot against the random bot, {n} games per row (half on each seat): win / loss / tie %   [recorded, 1,000 games]")
    for name in ("run_mcts", "run_gumbel"):
        for S, ref in VS_RANDOM.items():
            res = torch.cat([tm.play_match(n // 2, jaxrand.PRNGKey(1000 * S + seat + 1), S, bot_player=seat, search=getattr(tm, name), fused=fused)[1]
                             for seat in (1, -1)])
            w, l = float((res == 1).float().mean()) * 100, float((res == -1).float().mean()) * 100
            print(f"  {name:10s} {S:4d} sims: {w:5.1f} / {l:4.1f} / {100 - w - l:4.1f}   [{ref[0]} / {ref[1]} / {ref[2]}]")
    print(f"search against search, {n} games per row: first player / second player / draw %   [recorded, 1,000 games]")
    lines = torch.tensor([[0, 1, 2], [3, 4, 5], [6, 7, 8], [0, 3, 6], [1, 4, 7], [2, 5, 8], [0, 4, 8], [2, 4, 6]], device="cuda")
    for (name, S), ref in SELFPLAY.items():
        env, _ = tm.play_mcts_games(n, jaxrand.PRNGKey(77 + S), num_simulations=S, limit=30, variant=1, search=getattr(tm, name), fused=fused)
        sums = env.raw("board").reshape(n, 9).to(torch.int32)[:, lines].sum(2)
        a, b = float((sums == 3).any(1).float().mean()) * 100, float((sums == -3).any(1).float().mean()) * 100
        print(f"  {name:10s} {S:4d} sims: {a:5.1f} / {b:4.1f} / {100 - a - b:4.1f}   [{ref[0]} / {ref[1]} / {ref[2]}]")


if __name__ == "__main__":
    main()
