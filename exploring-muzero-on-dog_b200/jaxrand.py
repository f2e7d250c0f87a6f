"""jax.random stand-ins used by the self-play drivers around the env kernels
(MuZero_det_MADN/game_agent.py:60,187-188; evaluate_agent.py:741).

Keys are raw uint32[2] (what jax.random.PRNGKey returns with the default threefry impl).
Scalar key chaining (PRNGKey / split of ONE key) is a few dozen integer ops and runs on the host;
anything of size n (split into n keys, randint / uniform / bits of shape (n,)) runs on the GPU in
libdogstep.so (dogstep_random_*).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

_M = 0xFFFFFFFF
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
    return ((x << r) | (x >> (32 - r))) & _M


def threefry2x32(key, c0, c1):
    """Threefry-2x32-20 on Python ints (host-side scalar key chaining only)."""
    k0, k1 = int(key[0]) & _M, int(key[1]) & _M
    ks = (k0, k1, k0 ^ k1 ^ 0x1BD11BDA)
    x0, x1 = (c0 + ks[0]) & _M, (c1 + ks[1]) & _M
    for g in range(1, 6):
        for r in _ROT[(g - 1) & 1]:
            x0 = (x0 + x1) & _M
            x1 = _rotl(x1, r) ^ x0
        x0 = (x0 + ks[g % 3]) & _M
        x1 = (x1 + ks[(g + 1) % 3] + g) & _M
    return x0, x1


def PRNGKey(seed):
    return np.array([0, int(seed) & _M], dtype=np.uint32)


def split_host(key, num=2):
    """jax.random.split(key, num) for small num, on the host -> uint32 [num, 2]"""
    return np.array([threefry2x32(key, 0, i) for i in range(num)], dtype=np.uint32)


def split(key, num, device="cuda"):
    """jax.random.split(key, num) on the GPU -> uint32 tensor [num, 2]"""
    out = torch.empty((num, 2), dtype=torch.uint32, device=device)
    _lib.check(_lib.lib().dogstep_random_split(_lib.host_key(key), C.c_int64(num), _lib.ptr(out), _lib.stream()), "random_split")
    return out


def randint(key, n, minval, maxval, device="cuda"):
    out = torch.empty(n, dtype=torch.int32, device=device)
    _lib.check(_lib.lib().dogstep_random_randint(_lib.host_key(key), C.c_int64(n), C.c_int32(minval), C.c_int32(maxval),
                                                _lib.ptr(out), _lib.stream()), "random_randint")
    return out


def uniform(key, n, minval=0.0, maxval=1.0, device="cuda"):
    out = torch.empty(n, dtype=torch.float32, device=device)
    _lib.check(_lib.lib().dogstep_random_uniform(_lib.host_key(key), C.c_int64(n), C.c_float(minval), C.c_float(maxval),
                                                _lib.ptr(out), _lib.stream()), "random_uniform")
    return out


def bits(key, n, device="cuda"):
    out = torch.empty(n, dtype=torch.uint32, device=device)
    _lib.check(_lib.lib().dogstep_random_bits(_lib.host_key(key), C.c_int64(n), _lib.ptr(out), _lib.stream()), "random_bits")
    return out
