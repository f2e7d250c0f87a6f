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
    """jax.random.split(key, num) for small num, on the host (libdogstep's host-side Threefry, ~1 us) -> uint32 [num, 2]"""
    out = np.empty((num, 2), dtype=np.uint32)
    k = key if (isinstance(key, np.ndarray) and key.dtype == np.uint32 and key.flags.c_contiguous) else np.ascontiguousarray(key, dtype=np.uint32)
    rc = _host_fn("dogstep_host_split")(k.ctypes.data, num, out.ctypes.data)
    if rc:
        _lib.check(rc, "host_split")
    return out


class KeyChain:
    """The loop key of a host-driven lockstep loop, kept in a ctypes buffer: `rng_key, *step_keys = split(rng_key, N + 1)` every
    iteration (game_agent.py:60, evaluate_agent.py:741) needs only element 0 on the host, and advance() computes it in place
    in ~1 us (a NumPy round trip per iteration costs more than the kernel launch it feeds).  Accepted wherever a host key is."""

    def __init__(self, key):
        k = np.asarray(key, dtype=np.uint32).reshape(2)
        self.buf = (C.c_uint32 * 2)(int(k[0]), int(k[1]))
        self._fn = _host_fn("dogstep_host_key_chain")
        self._p = C.addressof(self.buf)

    def advance(self, steps=1):
        self._fn(self._p, steps, self._p)
        return self

    def numpy(self):
        return np.array([self.buf[0], self.buf[1]], dtype=np.uint32)

    def __iter__(self):
        return iter((self.buf[0], self.buf[1]))

    def __len__(self):
        return 2

    def __getitem__(self, i):
        return self.buf[i]


_HOST = {}


def _host_fn(name):
    """raw ctypes function with argtypes set (no per-call argument objects): the host-side key arithmetic sits in per-iteration
    host loops, its cost is all call overhead"""
    f = _HOST.get(name)
    if f is None:
        f = getattr(_lib.lib()._cdll, name)
        f.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        f.restype = C.c_int
        _HOST[name] = f
    return f


def key_chain_host(key, steps):
    """the loop key after `steps` lockstep iterations: rng <- split(rng, N + 1)[0], `steps` times (game_agent.py:60)"""
    out = np.empty(2, dtype=np.uint32)
    k = key if (isinstance(key, np.ndarray) and key.dtype == np.uint32 and key.flags.c_contiguous) else np.ascontiguousarray(key, dtype=np.uint32)
    rc = _host_fn("dogstep_host_key_chain")(k.ctypes.data, int(steps), out.ctypes.data)
    if rc:
        _lib.check(rc, "host_key_chain")
    return out


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
