"""jax.numpy stand-in (torch-backed).  TEST INFRASTRUCTURE ONLY -- see oracle/jaxshim/README.md.

Differentiation conventions that differ between torch and JAX are written the JAX way:
  * clip(a, lo, hi) = minimum(hi, maximum(lo, a))  (jax/_src/numpy/lax_numpy.py) -> a tie with a bound
    passes HALF the cotangent (lax.max/min "balanced" JVP); torch.clamp would pass all of it;
  * linalg.norm = sqrt(sum(x*x)) -> NaN gradient at exactly 0 (torch.linalg.norm returns 0 there).
"""
import math as _math

import numpy as _np
import torch as _torch

from .._array import Array, T as _T, W as _W, _dtype, float_dtype as _float_dtype
from . import linalg  # noqa: F401

ndarray = Array
float32 = _torch.float32
float64 = _torch.float64
int32 = _torch.int32
int64 = _torch.int64
uint32 = _torch.int64
bool_ = _torch.bool
inf = float("inf")
pi = _math.pi
newaxis = None


def array(x, dtype=None):
    t = _T(x)
    if dtype is not None:
        t = t.to(_dtype(dtype))
    return _W(t)


asarray = array


def _shape(s):
    if isinstance(s, (int, _np.integer)):
        return (int(s),)
    return tuple(int(i) for i in s)


def zeros(shape, dtype=None):
    return _W(_torch.zeros(_shape(shape), dtype=_dtype(dtype) or _float_dtype()))


def ones(shape, dtype=None):
    return _W(_torch.ones(_shape(shape), dtype=_dtype(dtype) or _float_dtype()))


def full(shape, fill_value, dtype=None):
    v = _T(fill_value)
    return _W(_torch.full(_shape(shape), v.item(), dtype=_dtype(dtype) or v.dtype))


def zeros_like(x):
    return _W(_torch.zeros_like(_T(x)))


def ones_like(x):
    return _W(_torch.ones_like(_T(x)))


def eye(n, dtype=None):
    return _W(_torch.eye(int(n), dtype=_dtype(dtype) or _float_dtype()))


def arange(*a, dtype=None):
    r = _np.arange(*[(int(v) if float(v).is_integer() else float(v)) for v in a])
    return array(r, dtype)


def linspace(a, b, n):
    return _W(_torch.linspace(float(a), float(b), int(n), dtype=_float_dtype()))


def indices(dims):
    return tuple(_W(_torch.from_numpy(g).to(_torch.int32)) for g in _np.indices(_shape(dims)))


def concatenate(xs, axis=0):
    ts = [_T(x) for x in xs]
    dt = ts[0].dtype
    for t in ts[1:]:
        dt = _torch.promote_types(dt, t.dtype)
    return _W(_torch.cat([t.to(dt) for t in ts], dim=axis))


def stack(xs, axis=0):
    ts = [_T(x) for x in xs]
    dt = ts[0].dtype
    for t in ts[1:]:
        dt = _torch.promote_types(dt, t.dtype)
    return _W(_torch.stack([t.to(dt) for t in ts], dim=axis))


def hstack(xs):
    return concatenate(xs, axis=1 if _T(xs[0]).dim() > 1 else 0)


def where(c, a, b):
    c = _T(c)
    if c.dtype != _torch.bool:
        c = c != 0
    a, b = _T(a), _T(b)
    dt = _torch.result_type(a, b)
    return _W(_torch.where(c, a.to(dt), b.to(dt)))


def maximum(a, b):
    a, b = _T(a), _T(b)
    dt = _torch.result_type(a, b)
    return _W(_torch.maximum(a.to(dt), b.to(dt)))       # ties: cotangent split evenly, as lax.max


def minimum(a, b):
    a, b = _T(a), _T(b)
    dt = _torch.result_type(a, b)
    return _W(_torch.minimum(a.to(dt), b.to(dt)))


def clip(a, a_min=None, a_max=None):
    if a_min is not None:
        a = maximum(a_min, a)
    if a_max is not None:
        a = minimum(a_max, a)
    return a if isinstance(a, Array) else array(a)


def _unary(fn):
    def f(x):
        t = _T(x)
        if not t.is_floating_point():
            t = t.to(_float_dtype())
        return _W(fn(t))
    return f


sqrt = _unary(_torch.sqrt)
exp = _unary(_torch.exp)
log = _unary(_torch.log)
sin = _unary(_torch.sin)
cos = _unary(_torch.cos)
tanh = _unary(_torch.tanh)
floor = _unary(_torch.floor)
square = _unary(_torch.square)


def abs(x):  # noqa: A001
    return _W(_T(x).abs())


absolute = abs


def isnan(x):
    return _W(_torch.isnan(_T(x)))


def nan_to_num(x, nan=0.0, posinf=None, neginf=None):
    return _W(_torch.nan_to_num(_T(x), nan=nan, posinf=posinf, neginf=neginf))


def sum(x, axis=None, keepdims=False):  # noqa: A001
    return array(x).sum(axis, keepdims)


def prod(x, axis=None):
    return array(x).prod(axis)


def mean(x, axis=None):
    return array(x).mean(axis)


def max(x, axis=None):  # noqa: A001
    return array(x).max(axis)


def min(x, axis=None):  # noqa: A001
    return array(x).min(axis)


def argsort(x, axis=-1):
    return _W(_torch.argsort(_T(x), dim=axis, stable=True).to(_torch.int32))


def argmin(x, axis=None):
    t = _T(x)
    return _W((t.argmin() if axis is None else t.argmin(dim=axis)).to(_torch.int32))


def dot(a, b):
    a, b = _T(a), _T(b)
    dt = _torch.result_type(a, b)
    a, b = a.to(dt), b.to(dt)
    if b.dim() == 1:
        return _W((a * b).sum(-1))
    return _W(_torch.tensordot(a, b, dims=([a.dim() - 1], [b.dim() - 2])))


def matmul(a, b):
    a, b = _T(a), _T(b)
    dt = _torch.result_type(a, b)
    return _W(_torch.matmul(a.to(dt), b.to(dt)))


def outer(a, b):
    return _W(_torch.outer(_T(a).flatten(), _T(b).flatten()))


def cross(a, b):
    a, b = _T(a), _T(b)
    shape = _torch.broadcast_shapes(a.shape, b.shape)
    return _W(_torch.linalg.cross(a.expand(shape), b.expand(shape), dim=-1))


def einsum(spec, *ops):
    return _W(_torch.einsum(spec, *[_T(o) for o in ops]))


def expand_dims(x, axis):
    return _W(_T(x).unsqueeze(axis))


def transpose(x, axes=None):
    return array(x).transpose(*([axes] if axes is not None else []))


def conjugate(x):
    return array(x)


conj = conjugate


def nonzero(x):
    return tuple(_W(i.to(_torch.int32)) for i in _torch.nonzero(_T(x), as_tuple=True))


def reshape(x, shape):
    return array(x).reshape(shape)


def power(a, b):
    return array(a) ** b


def isfinite(x):
    return _W(_torch.isfinite(_T(x)))


def float_(x):
    return array(x, dtype=float32)
