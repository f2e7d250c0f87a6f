"""`jax.random` (threefry2x32, partitionable) on NumPy uint32 arrays.  TEST INFRASTRUCTURE ONLY.
Same published algorithm as oracle/jaxrand_oracle.h, written independently in vectorised NumPy; the two are
cross-checked in tests/test_oracle_threefry.py when /root/reference-side goldens are generated."""
import numpy as _np

from ._core import wrap

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _tf(key, c0, c1):
    k0, k1 = _np.uint32(key[0]), _np.uint32(key[1])
    ks = (k0, k1, _np.uint32(k0 ^ k1 ^ _np.uint32(0x1BD11BDA)))
    with _np.errstate(over="ignore"):
        x0 = (c0.astype(_np.uint32) + ks[0]).astype(_np.uint32)
        x1 = (c1.astype(_np.uint32) + ks[1]).astype(_np.uint32)
        for g in range(1, 6):
            for r in _ROT[(g - 1) & 1]:
                x0 = (x0 + x1).astype(_np.uint32)
                x1 = ((x1 << _np.uint32(r)) | (x1 >> _np.uint32(32 - r))).astype(_np.uint32)
                x1 = x1 ^ x0
            x0 = (x0 + ks[g % 3]).astype(_np.uint32)
            x1 = (x1 + ks[(g + 1) % 3] + _np.uint32(g)).astype(_np.uint32)
    return x0, x1


def _counts(shape):
    n = int(_np.prod(shape)) if len(shape) else 1
    lo = _np.arange(n, dtype=_np.uint32)
    return _np.zeros(n, _np.uint32), lo


def _shape(shape):
    if isinstance(shape, (int, _np.integer)):
        return (int(shape),)
    return tuple(int(s) for s in shape)


def PRNGKey(seed):
    s = int(_np.asarray(seed)) & 0xFFFFFFFFFFFFFFFF
    return wrap(_np.array([0, s & 0xFFFFFFFF], dtype=_np.uint32))


key = PRNGKey


def split(key, num=2):
    key = _np.asarray(key)
    shape = _shape(num)
    hi, lo = _counts(shape)
    a, b = _tf(key, hi, lo)
    return wrap(_np.stack([a, b], axis=-1).reshape(shape + (2,)))


def bits(key, shape=(), dtype=_np.uint32):
    shape = _shape(shape)
    hi, lo = _counts(shape)
    a, b = _tf(_np.asarray(key), hi, lo)
    return wrap((a ^ b).reshape(shape))


def uniform(key, shape=(), dtype=_np.float32, minval=0.0, maxval=1.0):
    b = _np.asarray(bits(key, shape))
    f = ((b >> _np.uint32(9)) | _np.uint32(0x3F800000)).view(_np.float32) - _np.float32(1.0)
    minval, maxval = _np.float32(minval), _np.float32(maxval)
    v = f * (maxval - minval) + minval
    return wrap(_np.maximum(minval, v).astype(_np.float32).reshape(_shape(shape)))


def randint(key, shape, minval, maxval, dtype=_np.int32):
    k1, k2 = split(key)
    hb, lb = _np.asarray(bits(k1, shape)), _np.asarray(bits(k2, shape))
    lo, hi = int(_np.asarray(minval)), int(_np.asarray(maxval))
    span = _np.uint32(hi - lo) if hi > lo else _np.uint32(1)
    with _np.errstate(over="ignore"):
        mult = _np.uint32(65536) % span
        mult = _np.uint32((mult * mult) % span)
        off = ((hb % span) * mult + (lb % span)).astype(_np.uint32) % span
        return wrap((_np.int32(lo) + off.astype(_np.int32)).astype(_np.int32))


def gumbel(key, shape=(), dtype=_np.float32):
    u = _np.asarray(uniform(key, shape, minval=_np.finfo(_np.float32).tiny, maxval=1.0))
    # float contract of DESIGN.md: log evaluated in float64 and rounded to float32 once per call (XLA's own logf
    # polynomial cannot be reproduced here; any two faithful logf differ by <= 1 ulp)
    inner = _np.log(u.astype(_np.float64)).astype(_np.float32)
    return wrap((-_np.log((-inner).astype(_np.float64)).astype(_np.float32)).astype(_np.float32))


def categorical(key, logits, axis=-1, shape=None):
    logits = _np.asarray(logits, dtype=_np.float32)
    g = _np.asarray(gumbel(key, logits.shape))
    return wrap(_np.asarray(_np.argmax(g + logits, axis=axis)).astype(_np.int32))


def _cumsum6_cpu(p):
    """XLA:CPU association of cumsum for 6 elements."""
    p = p.astype(_np.float32)
    s01, s23, s45 = p[0] + p[1], p[2] + p[3], p[4] + p[5]
    c3 = s01 + s23
    return _np.array([p[0], s01, s01 + p[2], c3, c3 + p[4], c3 + s45], dtype=_np.float32)


def choice(key, a, shape=(), replace=True, p=None):
    a = _np.asarray(a)
    if a.ndim == 0:
        a = _np.arange(int(a), dtype=_np.int32)
    if p is None:
        ind = _np.asarray(randint(key, shape, 0, a.shape[0]))
        return wrap(a[ind])
    p = _np.asarray(p, dtype=_np.float32)
    c = _cumsum6_cpu(p) if p.shape[0] == 6 else _np.cumsum(p, dtype=_np.float32)
    u = _np.asarray(uniform(key, shape))
    r = c[-1] * (_np.float32(1.0) - u)
    ind = _np.searchsorted(c, r, side="left")
    ind = _np.clip(ind, 0, a.shape[0] - 1)
    return wrap(a[ind])


def permutation(key, x):
    raise NotImplementedError


def normal(key, shape=(), dtype=_np.float32):
    raise NotImplementedError
