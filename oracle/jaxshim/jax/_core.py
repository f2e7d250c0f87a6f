"""Array type with JAX indexing / dtype rules, pytree helpers, vmap, jit."""
import dataclasses
import functools

import numpy as np

_DOWN = {np.dtype("int64"): np.int32, np.dtype("uint64"): np.uint32, np.dtype("float64"): np.float32}


def _narrow(x):
    """JAX runs with x64 disabled: every 64-bit result becomes its 32-bit sibling."""
    if isinstance(x, np.ndarray) and x.dtype in _DOWN:
        x = x.astype(_DOWN[x.dtype])
    elif isinstance(x, np.generic) and x.dtype in _DOWN:
        x = np.asarray(x).astype(_DOWN[x.dtype])
    return x


def wrap(x):
    if isinstance(x, Arr):
        return x
    if isinstance(x, (np.ndarray, np.generic)):
        return _narrow(np.asarray(x)).view(Arr)
    return x


def _is_index_array(i):
    return isinstance(i, (np.ndarray, np.generic)) and np.asarray(i).dtype != np.bool_ or \
        (isinstance(i, (int, np.integer)) and not isinstance(i, (bool, np.bool_)))


def _norm_index(idx, shape, clamp):
    """JAX gather: negative indices wrap once, then clamp.  For scatters (`clamp=False`) returns (index, in_bounds)."""
    if not isinstance(idx, tuple):
        idx = (idx,)
    out, ok, dim = [], True, 0
    if any(i is Ellipsis for i in idx):  # expand `...` into full slices (the loop files index obs[None, ...])
        k = next(j for j, i in enumerate(idx) if i is Ellipsis)
        consumed = sum(1 for i in idx if i is not None and i is not Ellipsis)
        idx = idx[:k] + (slice(None),) * (len(shape) - consumed) + idx[k + 1:]
    for i in idx:
        if i is None:
            out.append(None)
            continue
        if isinstance(i, slice):
            out.append(i)
            dim += 1
            continue
        a = np.asarray(i)
        if a.dtype == np.bool_:
            out.append(a)
            dim += a.ndim
            continue
        size = shape[dim]
        a = a.astype(np.int64)
        a = np.where(a < 0, a + size, a)
        inb = (a >= 0) & (a < size)
        if clamp:
            a = np.clip(a, 0, size - 1)
        else:
            ok = ok & inb
        out.append(a if a.ndim else int(a))
        dim += 1
    return tuple(out), ok


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def _scatter(self, values, op):
        base = np.array(self.arr, copy=True)
        idx, ok = _norm_index(self.idx, base.shape, clamp=False)
        values = np.asarray(values)
        if ok is True or bool(np.all(ok)):
            if op == "set":
                base[idx] = values.astype(base.dtype)
            else:
                np.add.at(base, idx, values.astype(base.dtype))
            return wrap(base)
        # some updates are out of bounds -> drop them, one index tuple at a time
        if any(isinstance(i, slice) or i is None for i in idx):
            raise NotImplementedError("jaxshim: out-of-bounds scatter mixed with slices")
        arrays = [np.asarray(i) for i in idx]
        b = np.broadcast_arrays(*arrays, np.asarray(ok))
        okb = b[-1]
        tail = base.shape[len(arrays):]
        vals = np.broadcast_to(values, okb.shape + tail)
        for pos in np.ndindex(okb.shape):
            if okb[pos]:
                tgt = tuple(int(x[pos]) for x in b[:-1])
                if op == "set":
                    base[tgt] = vals[pos]
                else:
                    base[tgt] += vals[pos]
        return wrap(base)

    def set(self, values, mode=None, **kw):
        return self._scatter(values, "set")

    def add(self, values, mode=None, **kw):
        return self._scatter(values, "add")

    def get(self, mode=None, **kw):
        return self.arr[self.idx]


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


class Arr(np.ndarray):
    """np.ndarray with JAX semantics for indexing, `.at`, and 32-bit results."""

    __array_priority__ = 100

    def __array_finalize__(self, obj):
        pass

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        args = [np.asarray(i) if isinstance(i, Arr) else i for i in inputs]
        if out is not None:
            kwargs["out"] = tuple(np.asarray(o) if isinstance(o, Arr) else o for o in out)
        with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
            res = getattr(ufunc, method)(*args, **kwargs)
        if isinstance(res, tuple):
            return tuple(wrap(r) for r in res)
        return wrap(res)

    def __getitem__(self, idx):
        base = np.asarray(self)
        idx2, _ = _norm_index(idx, base.shape, clamp=True)
        return wrap(base[idx2])

    def __setitem__(self, idx, v):
        raise TypeError("JAX arrays are immutable; use .at[].set()")

    @property
    def at(self):
        return _At(self)

    def astype(self, dtype, *a, **k):
        with np.errstate(over="ignore", invalid="ignore"):
            return wrap(np.asarray(self).astype(dtype))

    def flatten(self, *a, **k):
        return wrap(np.asarray(self).flatten())

    def reshape(self, *a, **k):
        return wrap(np.asarray(self).reshape(*a, **k))

    @property
    def T(self):
        return wrap(np.asarray(self).T)

    def __iter__(self):
        base = np.asarray(self)
        if base.ndim == 0:
            raise TypeError("iteration over a 0-d array")
        return (wrap(base[i]) for i in range(base.shape[0]))

    def __bool__(self):
        return bool(np.asarray(self))

    def __index__(self):
        return int(np.asarray(self))

    def __hash__(self):
        return id(self)

    def item(self, *a):
        return np.asarray(self).item(*a)

    def block_until_ready(self):
        return self


# ------------------------------------------------------------------ pytrees
def _is_struct(x):
    return dataclasses.is_dataclass(x) and not isinstance(x, type)


def _node_fields(x):
    return [f for f in dataclasses.fields(x) if f.metadata.get("pytree_node", True)]


def tree_map(fn, tree, *rest):
    if _is_struct(tree):
        kw = {f.name: tree_map(fn, getattr(tree, f.name), *[getattr(r, f.name) for r in rest]) for f in _node_fields(tree)}
        return dataclasses.replace(tree, **kw)
    if isinstance(tree, dict):
        return {k: tree_map(fn, v, *[r[k] for r in rest]) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)) and not hasattr(tree, "_fields"):
        return type(tree)(tree_map(fn, v, *[r[i] for r in rest]) for i, v in enumerate(tree))
    if hasattr(tree, "_fields"):  # namedtuple
        return type(tree)(*[tree_map(fn, v, *[r[i] for r in rest]) for i, v in enumerate(tree)])
    if tree is None:
        return None
    return fn(tree, *rest)


def tree_leaves(tree):
    out = []
    tree_map(lambda x: out.append(x), tree)
    return out


def _take(tree, axis, i):
    if axis is None:
        return tree
    return tree_map(lambda x: wrap(np.take(np.asarray(x), i, axis=axis)), tree)


def _axis_size(tree, axis):
    if axis is None:
        return None
    leaves = tree_leaves(tree)
    return np.asarray(leaves[0]).shape[axis]


def vmap(fn, in_axes=0, out_axes=0):
    """jax.vmap as a Python loop over the mapped axis + stack of the results (pytrees supported)."""

    @functools.wraps(fn)
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        sizes = [s for s in (_axis_size(a, ax) for a, ax in zip(args, axes)) if s is not None]
        n = sizes[0]
        outs = [fn(*[_take(a, ax, i) for a, ax in zip(args, axes)]) for i in range(n)]
        return tree_map(lambda *xs: wrap(np.stack([np.asarray(x) for x in xs], axis=out_axes)), outs[0], *outs[1:])

    return mapped


def jit(fn=None, **kw):
    if fn is None:
        return lambda f: f
    return fn
