"""jax.numpy.linalg stand-in (see ../__init__.py for the differentiation conventions)."""
import torch as _torch

from .._array import T as _T, W as _W


def norm(x, ord=None, axis=None, keepdims=False):
    # jnp.linalg.norm (2-norm): sqrt(sum(real(x * conj(x)))) -> NaN cotangent at exactly 0, as in JAX
    t = _T(x)
    if not t.is_floating_point():
        from .._array import float_dtype
        t = t.to(float_dtype())
    assert ord in (None, 2)
    s = (t * t).sum() if axis is None else (t * t).sum(dim=axis, keepdim=keepdims)
    return _W(_torch.sqrt(s))


def svd(A, full_matrices=True, compute_uv=True):
    # LAPACK gesdd on CPU, the same family jnp.linalg.svd dispatches to on the CPU backend
    U, S, Vh = _torch.linalg.svd(_T(A), full_matrices=full_matrices)
    if not compute_uv:
        return _W(S)
    return _W(U), _W(S), _W(Vh)


def inv(A):
    return _W(_torch.linalg.inv(_T(A)))


def det(A):
    return _W(_torch.linalg.det(_T(A)))
