"""Minimal torch-backed stand-in for the JAX APIs used by DaXBench's simulator step.

TEST INFRASTRUCTURE ONLY.  Purpose: execute the UNMODIFIED reference sources
(/root/reference/DaXBench/daxbench/core/engine/*.py) in a container without jax/jaxlib, so that golden
vectors come from the reference's own Python (control flow, formulas, index arithmetic, custom_vjp rules)
rather than from a restatement.  torch autograd is the reverse-mode engine behind jax.grad / custom_vjp.
See oracle/jaxshim/README.md for the list of mirrored semantics and known differences.
"""
import torch as _torch

from . import tree_util  # noqa: F401
from ._array import Array, T as _T, W as _W, set_float, float_dtype  # noqa: F401
from . import numpy  # noqa: F401
from . import lax  # noqa: F401
from . import random  # noqa: F401
from ._src.lax import control_flow as _cf  # noqa: F401

__version__ = "0.0-shim"


def jit(f, *a, **k):
    """No compilation, but the one observable thing jax.jit does to its inputs: NumPy arrays arrive as jax arrays
    (float64 -> float32, int64 -> int32 with x64 disabled), e.g. shape_rope_env.py:129 feeds np actions to a jitted
    step_diff that then uses `.at[]`."""
    import functools

    import numpy as _np

    def conv(l):
        if isinstance(l, _np.ndarray):
            t = _torch.from_numpy(_np.ascontiguousarray(l))
            if t.dtype == _torch.float64:
                t = t.to(_torch.float32)
            elif t.dtype == _torch.int64:
                t = t.to(_torch.int32)
            return _W(t)
        return l

    @functools.wraps(f)
    def wrapped(*args, **kw):
        return f(*[tree_util.tree_map(conv, x) for x in args], **{n: tree_util.tree_map(conv, x) for n, x in kw.items()})
    return wrapped


def _rebuild(tree):
    """Fresh containers with the same leaves: what every jax transformation boundary does to a pytree
    (in-place list edits inside a traced function never reach the caller's containers)."""
    leaves, td = tree_util.tree_flatten(tree)
    return tree_util.tree_unflatten(td, leaves)


def vmap(f, in_axes=0, out_axes=0):
    def batched(*args, **kwargs):
        # jax.vmap maps keyword arguments over axis 0 (cloth_env_para.py:191 passes eval_min_max_stiff by keyword)
        kw_names = list(kwargs)
        n_pos = len(args)
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * n_pos
        args = tuple(args) + tuple(kwargs[k] if isinstance(kwargs[k], Array) else numpy.array(kwargs[k]) for k in kw_names)
        axes = tuple(axes) + (0,) * len(kw_names)
        B = None
        for a, ax in zip(args, axes):
            if ax is None:
                continue
            for leaf in tree_util.tree_leaves(a):
                B = _T(leaf).shape[ax]
                break
            if B is not None:
                break
        outs = []
        for b in range(B):
            sl = [a if ax is None else tree_util.tree_map(lambda l, ax=ax: _W(_T(l).select(ax, b)), a)
                  for a, ax in zip(args, axes)]
            outs.append(f(*sl[:n_pos], **dict(zip(kw_names, sl[n_pos:]))))
        return tree_util.tree_map(lambda *ls: _W(_torch.stack([_T(l) for l in ls], dim=out_axes)), outs[0], *outs[1:])
    return batched


def _is_float_leaf(x):
    return isinstance(x, (Array, _torch.Tensor)) and _T(x).is_floating_point()


def grad(fun, argnums=0, allow_int=False, has_aux=False):
    def g(*args):
        x = args[argnums]
        leaves, td = tree_util.tree_flatten(x)
        new_leaves, diff = [], []
        for l in leaves:
            if _is_float_leaf(l):
                t = _T(l).detach().clone().requires_grad_(True)
                diff.append(t)
                new_leaves.append(_W(t))
            else:
                new_leaves.append(l)
        args2 = list(args)
        args2[argnums] = tree_util.tree_unflatten(td, new_leaves)
        with _torch.enable_grad():
            out = fun(*args2)
            aux = None
            if has_aux:
                out, aux = out
            gs = _torch.autograd.grad(_T(out), diff, allow_unused=True)
        it = iter(gs)
        res = []
        for l in leaves:
            if _is_float_leaf(l):
                gi = next(it)
                res.append(_W(gi if gi is not None else _torch.zeros_like(_T(l))))
            else:
                # float0 cotangent of an integer input: behaves as zeros
                res.append(_W(_torch.zeros_like(_T(l), dtype=float_dtype())) if isinstance(l, (Array, _torch.Tensor))
                           else 0.0)
        gtree = tree_util.tree_unflatten(td, res)
        return (gtree, aux) if has_aux else gtree
    return g


def value_and_grad(fun, argnums=0, has_aux=False):
    def vg(*args):
        raise NotImplementedError
    return vg


_FLOAT_SLOT = object()


class _CustomVJPFn(_torch.autograd.Function):
    """One autograd node per custom_vjp call.  forward = the user's fwd rule on detached primals (nothing
    is recorded, as in JAX); backward = the user's bwd rule on (residuals, cotangent pytree)."""

    @staticmethod
    def forward(ctx, cv, nondiff, td_args, slots, holder, *tensors):
        it = iter(tensors)
        leaves = [(_W(next(it).detach()) if s is _FLOAT_SLOT else s) for s in slots]
        args = tree_util.tree_unflatten(td_args, leaves)
        out, res = cv.fwd(*_merge_args(cv.nondiff_argnums, nondiff, args))
        out_leaves, td_out = tree_util.tree_flatten(out)
        is_f = [_is_float_leaf(l) for l in out_leaves]
        ctx.cv, ctx.nondiff, ctx.res = cv, nondiff, res
        ctx.td_out, ctx.out_leaves, ctx.is_f = td_out, out_leaves, is_f
        ctx.in_leaves, ctx.slots = leaves, slots
        holder["td_out"], holder["out_leaves"], holder["is_f"] = td_out, out_leaves, is_f
        return tuple(_T(l).clone() for l, f in zip(out_leaves, is_f) if f)

    @staticmethod
    def backward(ctx, *gouts):
        it = iter(gouts)
        g_leaves = []
        for l, f in zip(ctx.out_leaves, ctx.is_f):
            if f:
                gi = next(it)
                g_leaves.append(_W(gi if gi is not None else _torch.zeros_like(_T(l))))
            elif isinstance(l, (Array, _torch.Tensor)):
                g_leaves.append(_W(_torch.zeros(_T(l).shape, dtype=float_dtype())))   # float0 cotangent
            else:
                g_leaves.append(0.0)
        g = tree_util.tree_unflatten(ctx.td_out, g_leaves)
        with _torch.enable_grad():
            cts = ctx.cv.bwd(*ctx.nondiff, ctx.res, g)
        ct_leaves = tree_util.tree_leaves(tuple(cts))
        assert len(ct_leaves) == len(ctx.in_leaves), (len(ct_leaves), len(ctx.in_leaves))
        grads = [_T(c).detach().to(_T(l).dtype).reshape(_T(l).shape)
                 for l, c, s in zip(ctx.in_leaves, ct_leaves, ctx.slots) if s is _FLOAT_SLOT]
        return (None, None, None, None, None, *grads)


def _merge_args(nondiff_argnums, nondiff, diff_args):
    n = len(nondiff) + len(diff_args)
    full, di, ni = [], iter(diff_args), iter(nondiff)
    for i in range(n):
        full.append(next(ni) if i in nondiff_argnums else next(di))
    return full


class custom_vjp:
    """jax.custom_vjp(fun, nondiff_argnums) with .defvjp(fwd, bwd)."""

    def __init__(self, fun, nondiff_argnums=()):
        self.fun = fun
        self.nondiff_argnums = tuple(nondiff_argnums)
        self.fwd = self.bwd = None
        self.__name__ = getattr(fun, "__name__", "custom_vjp")

    def defvjp(self, fwd, bwd):
        self.fwd, self.bwd = fwd, bwd

    def __call__(self, *args, **kwargs):
        if kwargs:
            raise NotImplementedError("jaxshim custom_vjp: keyword arguments")
        import inspect
        params = list(inspect.signature(self.fun).parameters.values())
        args = list(args) + [p.default for p in params[len(args):] if p.default is not inspect.Parameter.empty]
        nondiff = tuple(a for i, a in enumerate(args) if i in self.nondiff_argnums)
        diff_args = tuple(_rebuild(a) for i, a in enumerate(args) if i not in self.nondiff_argnums)
        leaves, td = tree_util.tree_flatten(diff_args)
        tensors = [_T(l) for l in leaves if _is_float_leaf(l)]
        needs = _torch.is_grad_enabled() and any(t.requires_grad for t in tensors)
        if not needs or self.fwd is None:
            return self.fun(*_merge_args(self.nondiff_argnums, nondiff, diff_args))
        slots = [(_FLOAT_SLOT if _is_float_leaf(l) else l) for l in leaves]
        holder = {}
        outs = _CustomVJPFn.apply(self, nondiff, td, slots, holder, *tensors)
        it = iter(outs)
        out_leaves = [(_W(next(it)) if f else l) for l, f in zip(holder["out_leaves"], holder["is_f"])]
        return tree_util.tree_unflatten(holder["td_out"], out_leaves)


def device_put(x, *a, **k):
    return x


def devices(*a):
    return ["cpu"]


def local_device_count():
    return 1
