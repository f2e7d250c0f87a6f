#!/usr/bin/env python
"""Kernel-time breakdown of ONE lockstep iteration of the self-play loops of bench.py (cfg 3 / cfg 5), eager launches, through
torch.profiler (CUPTI).  Prints the top kernels by total time and the share of libdogstep's own kernels.

    python scripts/prof_selfplay.py cfg5 [iterations]
"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    torch.cuda.set_device(0)
    torch.backends.cuda.matmul.allow_tf32 = True
    dev = torch.device("cuda", 0)
    loop, key = bench.make_selfplay(which, dev, 0, 1, cuda_graph=False, plies=iters + 2 + (int(sys.argv[4]) if len(sys.argv) > 4 else 0), net=(sys.argv[3] if len(sys.argv) > 3 else "standin"))
    skip = int(sys.argv[4]) if len(sys.argv) > 4 else 0   # lockstep iterations played before the profiled ones
    loop.max_steps = max(skip, 2)
    loop.run(key)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(iters):   # the next iterations of the SAME games (no re-seed, no buffer reset)
            loop._iteration()
        torch.cuda.synchronize()
    tot = collections.Counter()
    cnt = collections.Counter()
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            tot[e.name] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
            cnt[e.name] += 1
    total = sum(tot.values())
    own = sum(v for k, v in tot.items() if "dogstep" in k)
    print(f"{which}: {iters} iterations, {total / 1e3 / iters:.3f} ms of kernel time per iteration, libdogstep kernels {100 * own / total:.1f} %")
    for name, us in tot.most_common(25):
        print(f"{us / iters:10.1f} us/iter  {cnt[name] / iters:7.1f} launches/iter  {name[:110]}")
    from exploring_muzero_on_dog_b200 import _lib
    if os.environ.get("DOGSTEP_LIB") and hasattr(_lib.lib(), "dogstep_trace_wide_decided"):  # instrumented build only
        import ctypes
        c = (ctypes.c_ulonglong * 5)()
        _lib.lib().dogstep_trace_wide_decided(c)
        tot_l = max(1, c[3] + c[4])
        print(f"wide interior levels since start: decided without the row {c[3] / tot_l:.3f}; with the row: from bounds {c[1] / tot_l:.3f}, "
              f"exact evaluation {c[0] / tot_l:.3f}, two near-maximal children {c[2] / tot_l:.3f} (of {tot_l})")


if __name__ == "__main__":
    main()
