"""Batched env pytree stand-in shared by the MADN / DOG host mirrors.

The reference's envs are flax.struct dataclasses whose leaves gain a leading game axis under
jax.vmap (MuZero_det_MADN/game_agent.py:44-48).  `BatchedEnv` keeps exactly those leaves as
contiguous CUDA tensors (structure of arrays) so libdogstep.so can update them in place.
An env made from a scalar seed behaves like the reference's un-vmapped env: attribute reads
return the leaf without the game axis (a view, so in-place kernel updates stay visible).
"""
import numpy as np
import torch


def to_dev(x, dtype, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype).contiguous()
    a = np.asarray(x)
    if dtype == torch.uint32:
        return torch.from_numpy(a.astype(np.uint32)).to(device)
    return torch.as_tensor(a, device=device).to(dtype).contiguous()


def seeds_to_dev(seed, device):
    """(batched?, int32 [n] device tensor) from a scalar / array / tensor seed.  A tensor that already lives on the
    device is used as is — no host round trip, no synchronisation on the hot path."""
    if isinstance(seed, torch.Tensor):
        return seed.dim() > 0, seed.reshape(-1).to(device=device, dtype=torch.int32).contiguous()
    a = np.asarray(seed)
    return a.ndim > 0, to_dev(np.atleast_1d(a), torch.int32, device)


def reuse_or_alloc(cls, out, n, static, device, batched):
    """env_reset(out=...): overwrite the leaves of an existing env of the same size and configuration instead of
    allocating new ones (the steady state of a self-play iteration re-seeds the same buffers)"""
    if out is None:
        env = cls(n, static, torch.device(device), batched)
        env.alloc()
        return env
    pub = lambda d: {k: v for k, v in d.items() if not k.startswith("_")}
    if not isinstance(out, cls) or out.n != n or pub(out.static) != pub(static):
        raise ValueError("env_reset(out=...): size or configuration differs from the env to be overwritten")
    return out


class BatchedEnv:
    """Leaves: name -> (torch dtype, trailing shape).  Static fields live in `static`."""

    LEAVES = {}

    def __init__(self, n, static, device, batched):
        object.__setattr__(self, "_t", {})
        object.__setattr__(self, "n", n)
        object.__setattr__(self, "static", dict(static))
        object.__setattr__(self, "device", device)
        object.__setattr__(self, "batched", batched)

    def leaf_shape(self, name):
        return tuple(self.LEAVES[name][1](self.static))

    def alloc(self):
        for name, (dt, _) in self.LEAVES.items():
            self._t[name] = torch.empty((self.n,) + self.leaf_shape(name), dtype=dt, device=self.device)  # env_reset writes every leaf

    def __getattr__(self, name):
        t = object.__getattribute__(self, "_t")
        if name in t:
            return t[name] if self.batched else t[name][0]
        st = object.__getattribute__(self, "static")
        if name in st:
            return st[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        raise AttributeError("envs are immutable like the reference's dataclasses; use .replace()")

    def memo(self, key, make):
        """per-env cache of the ctypes structs handed to the C-ABI (leaves are never rebound after alloc(), static fields
        only change through replace(), which builds a new env): keeps the per-call host cost of the drop-in functions low"""
        c = self.__dict__.setdefault("_memo", {})
        v = c.get(key)
        if v is None:
            v = c[key] = make()
        return v

    def raw(self, name):
        """leaf WITH the game axis, whatever `batched` says"""
        return self._t[name]

    def replace(self, **kw):
        """dataclass-style functional update (returns a new env sharing untouched leaves' storage
        semantics of the reference: every leaf is copied, so the old env stays valid)."""
        new = self.__class__.__new__(self.__class__)
        BatchedEnv.__init__(new, self.n, self.static, self.device, self.batched)
        for name, t in self._t.items():
            if name in kw:
                dt = self.LEAVES[name][0]
                v = to_dev(kw.pop(name), dt, self.device)
                new._t[name] = v.reshape((self.n,) + self.leaf_shape(name)).clone()
            else:
                new._t[name] = t.clone()
        for k in list(kw):
            if k in self.static:
                new.static[k] = kw.pop(k)
        if kw:
            raise TypeError(f"unknown fields {sorted(kw)}")
        return new

    def clone(self):
        return self.replace()

    def numpy(self):
        """host copy of every leaf (with game axis) for comparisons in tests"""
        out = {}
        for k, t in self._t.items():
            out[k] = t.cpu().numpy()
        return out
