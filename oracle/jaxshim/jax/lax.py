"""jax.lax stand-in: control flow as plain Python (static trip counts, exactly what tracing would unroll/scan)."""
import torch as _torch

from . import tree_util
from ._array import T as _T, W as _W


def fori_loop(lower, upper, body_fun, init_val):
    val = init_val
    for i in range(int(lower), int(upper)):
        val = body_fun(i, val)
    return val


def scan(f, init, xs=None, length=None):
    carry = init
    n = length if xs is None else _T(tree_util.tree_leaves(xs)[0]).shape[0]
    ys = []
    for i in range(int(n)):
        x = None if xs is None else tree_util.tree_map(lambda l: _W(_T(l)[i]), xs)
        carry, y = f(carry, x)
        ys.append(y)
    if ys and ys[0] is not None:
        ys = tree_util.tree_map(lambda *ls: _W(_torch.stack([_T(l) for l in ls])), ys[0], *ys[1:])
    else:
        ys = None
    return carry, ys


def stop_gradient(x):
    return tree_util.tree_map(lambda l: _W(_T(l).detach()), x)


def cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def pmean(x, axis_name=None):
    return x
