"""`jax.lax` control flow as plain Python (TEST INFRASTRUCTURE ONLY)."""
import numpy as _np

from ._core import wrap


_NO_OPERAND = object()


def cond(pred, true_fun, false_fun, *operands, operand=_NO_OPERAND, **kw):
    if operand is not _NO_OPERAND and not operands:  # jax's legacy keyword: `operand=None` passes None as the one operand
        operands = (operand,)
    return true_fun(*operands) if bool(_np.asarray(pred)) else false_fun(*operands)


def fori_loop(lower, upper, body_fun, init_val):
    val = init_val
    for i in range(int(_np.asarray(lower)), int(_np.asarray(upper))):
        val = body_fun(wrap(_np.asarray(i, dtype=_np.int32)), val)
    return val


def while_loop(cond_fun, body_fun, init_val):
    val = init_val
    while bool(_np.asarray(cond_fun(val))):
        val = body_fun(val)
    return val


def switch(index, branches, *operands):
    i = int(_np.clip(int(_np.asarray(index)), 0, len(branches) - 1))
    return branches[i](*operands)


def select(pred, a, b):
    return wrap(_np.where(_np.asarray(pred), _np.asarray(a), _np.asarray(b)))


def stop_gradient(x):
    return x


def scan(f, init, xs, length=None):
    raise NotImplementedError("jaxshim: lax.scan is not needed for the env path")
