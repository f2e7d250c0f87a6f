"""N > 1 host logic on CPU: two gloo ranks shard the games exactly as bench.py / the GPU path does (rank r owns global
games [r*n, (r+1)*n), keys indexed by the GLOBAL game index, no collective while stepping) and exchange sampled replay
batches with the product's all-gather helper."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    from helpers import TRAIN_RULES, mask_of
    from exploring_muzero_on_dog_b200.vec_replay_buffer import allgather_batch
    n = n_total // world
    key = O.split(O.prng_key(0))[1]
    seeds = O.randint(key, n_total, 0, 1_000_000)[rank * n:(rank + 1) * n]
    s = O.madn_reset(O.MadnCfg(4, 0xF, 10, mask_of(TRAIN_RULES)), seeds, 0)
    glen, total, _ = O.madn_det_play_random(s, key, 2000, game_offset=rank * n)
    # whole-job aggregate exactly as bench.py reduces it
    t = torch.tensor([total], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    ms = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    batch = {"game_len": torch.from_numpy(glen.copy()), "pins": torch.from_numpy(s.pins.copy())}
    full = allgather_batch(batch)
    if rank == 0:
        np.savez(os.path.join(out_dir, "gathered.npz"), game_len=full["game_len"].numpy(), pins=full["pins"].numpy(),
                 total=t.item(), ms=ms.item())
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process(tmp_path):
    n_total = 512
    mp.spawn(_worker, args=(2, _free_port(), n_total, str(tmp_path)), nprocs=2, join=True)
    z = np.load(tmp_path / "gathered.npz")
    sys.path.insert(0, ROOT)
    import oracle as O
    from helpers import TRAIN_RULES, mask_of
    key = O.split(O.prng_key(0))[1]
    s = O.madn_reset(O.MadnCfg(4, 0xF, 10, mask_of(TRAIN_RULES)), O.randint(key, n_total, 0, 1_000_000), 0)
    glen, total, _ = O.madn_det_play_random(s, key, 2000, nthreads=4)
    assert np.array_equal(z["game_len"], glen) and np.array_equal(z["pins"], s.pins)
    assert int(z["total"]) == total and float(z["ms"]) == 11.0
