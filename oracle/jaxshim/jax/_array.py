"""Array type of the jax shim: a thin wrapper around torch.Tensor with numpy/JAX method semantics.

TEST INFRASTRUCTURE ONLY (see oracle/jaxshim/README.md).  JAX rules mirrored here, each relied on by
the reference path (DaXBench/daxbench/core/engine/*.py):
  * x64 disabled: floats are float32, ints int32 (set_float(torch.float64) switches the float type for
    the fp64 noise-floor runs);
  * python scalars are weakly typed (0-dim tensors in torch's promotion lattice);
  * arrays are immutable: `a += b` rebinds, `.at[...]` is functional;
  * gather (`a[idx]`): negative indices wrap once, out-of-bounds indices CLAMP;
  * scatter (`a.at[idx].set/add`): negative indices wrap once, out-of-bounds updates are DROPPED;
  * `astype(int32)` truncates toward zero; `trace()` runs over axes (0, 1).
"""
import numpy as np
import torch

_FLOAT = torch.float32


def set_float(dtype):
    """Float type standing in for jnp.float32 (torch.float64 for fp64 noise-floor runs)."""
    global _FLOAT
    _FLOAT = dtype
    torch.set_default_dtype(dtype)


def float_dtype():
    return _FLOAT


def _dtype(d):
    """Map a dtype spec (torch / numpy / python type / DType alias) to torch under x64-disabled rules."""
    if d is None:
        return None
    if isinstance(d, torch.dtype):
        if d in (torch.float64, torch.float32, torch.float16):
            return _FLOAT
        if d == torch.int64:
            return torch.int32
        return d
    if d in (float, np.float32, np.float64, "float32", "float64"):
        return _FLOAT
    if d in (int, np.int32, np.int64, "int32", "int64"):
        return torch.int32
    if d in (bool, np.bool_, "bool"):
        return torch.bool
    if d in (np.uint32, "uint32"):
        return torch.int64  # keys only
    return _dtype(torch.from_numpy(np.zeros((), dtype=d)).dtype)


def _canon(t):
    if t.dtype in (torch.float64, torch.float32, torch.float16) and t.dtype != _FLOAT:
        return t.to(_FLOAT)
    if t.dtype == torch.int64:
        return t.to(torch.int32)
    return t


def T(x):
    """Anything array-like -> torch.Tensor (canonical dtypes; python scalars become 0-dim = weakly typed)."""
    if isinstance(x, Array):
        return x.t
    if isinstance(x, torch.Tensor):
        return _canon(x)
    if isinstance(x, (bool, np.bool_)):
        return torch.tensor(bool(x))
    if isinstance(x, (int, np.integer)):
        return torch.tensor(int(x), dtype=torch.int32)
    if isinstance(x, (float, np.floating)):
        return torch.tensor(float(x), dtype=_FLOAT)
    if isinstance(x, np.ndarray):
        return _canon(torch.from_numpy(np.ascontiguousarray(x)))
    if isinstance(x, (list, tuple)):
        if any(isinstance(e, (Array, torch.Tensor)) for e in x) or any(isinstance(e, (list, tuple)) and any(
                isinstance(f, (Array, torch.Tensor)) for f in e) for e in x):
            elems = [T(e) for e in x]
            dt = elems[0].dtype
            for e in elems[1:]:
                dt = torch.promote_types(dt, e.dtype)
            shape = torch.broadcast_shapes(*[e.shape for e in elems])
            return torch.stack([e.to(dt).expand(shape) for e in elems])
        return T(np.asarray(x))
    raise TypeError(f"jaxshim: cannot convert {type(x)} to an array")


def W(t):
    return Array(t)


def _index_tensor(k):
    if isinstance(k, Array):
        return k.t
    if isinstance(k, torch.Tensor):
        return k
    if isinstance(k, np.ndarray):
        return torch.from_numpy(k)
    if isinstance(k, (list, tuple)):
        return torch.as_tensor(np.asarray(k))
    return None


def _norm_key(t, key):
    """Expand a numpy-style key against t.shape.  Returns a list of (kind, value, dim)."""
    if not isinstance(key, tuple):
        key = (key,)
    n_consumed = 0
    for k in key:
        if k is None or k is Ellipsis:
            continue
        kt = _index_tensor(k) if not isinstance(k, (int, np.integer, slice)) else None
        n_consumed += kt.dim() if (kt is not None and kt.dtype == torch.bool) else 1
    out, dim = [], 0
    for k in key:
        if k is None:
            out.append(("new", None, None))
        elif k is Ellipsis:
            n_fill = t.dim() - n_consumed
            for _ in range(n_fill):
                out.append(("slice", slice(None), dim))
                dim += 1
        elif isinstance(k, slice):
            out.append(("slice", k, dim))
            dim += 1
        elif isinstance(k, (int, np.integer)):
            out.append(("int", int(k), dim))
            dim += 1
        else:
            kt = _index_tensor(k)
            if kt is None:
                raise TypeError(f"jaxshim: unsupported index {type(k)}")
            if kt.dtype == torch.bool:
                out.append(("bool", kt, dim))
                dim += kt.dim()
            elif kt.dim() == 0:
                out.append(("int", int(kt.item()), dim))
                dim += 1
            else:
                out.append(("arr", kt.to(torch.int64), dim))
                dim += 1
    return out


def _gather_key(t, key):
    """Torch index implementing JAX gather rules: wrap negatives once, clamp out-of-bounds."""
    res = []
    for kind, v, dim in _norm_key(t, key):
        if kind == "new":
            res.append(None)
        elif kind in ("slice", "bool"):
            res.append(v)
        elif kind == "int":
            n = t.shape[dim]
            i = v + n if v < 0 else v
            res.append(min(max(i, 0), n - 1))
        else:
            n = t.shape[dim]
            i = torch.where(v < 0, v + n, v)
            res.append(i.clamp(0, n - 1))
    return tuple(res)


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, key):
        return _AtKey(self.arr, key)


class _AtKey:
    def __init__(self, arr, key):
        self.arr, self.key = arr, key

    def _apply(self, value, accumulate):
        t = self.arr.t
        val = T(value)
        if val.dtype != t.dtype:
            val = val.to(t.dtype)
        items = _norm_key(t, self.key)
        arrs = [(i, it) for i, it in enumerate(items) if it[0] == "arr"]
        idx = []
        valid = None
        for kind, v, dim in items:
            if kind == "new":
                raise NotImplementedError("jaxshim: None inside .at[] key")
            if kind in ("slice", "bool"):
                idx.append(v)
            elif kind == "int":
                n = t.shape[dim]
                i = v + n if v < 0 else v
                if i < 0 or i >= n:
                    return self.arr          # out-of-bounds scatter: dropped
                idx.append(i)
            else:
                n = t.shape[dim]
                i = torch.where(v < 0, v + n, v)
                ok = (i >= 0) & (i < n)
                valid = ok if valid is None else (valid & ok)
                idx.append(i)
        new = t.clone()
        if valid is not None and not bool(valid.all()):
            # supported drop pattern: every advanced index is 1-D of the same length and leads the key
            lead = all(items[i][0] == "arr" for i in range(len(arrs))) and all(it[1].dim() == 1 for _, it in arrs)
            if not lead:
                raise NotImplementedError("jaxshim: out-of-bounds scatter with a mixed key")
            keep = valid
            idx = [(i[keep] if isinstance(i, torch.Tensor) and i.dtype == torch.int64 else i) for i in idx]
            if val.dim() > 0 and val.shape[0] == keep.shape[0]:
                val = val[keep]
        leading = all(it[0] == "arr" for it in items[:len(arrs)]) and all(it[0] in ("arr", "slice") for it in items)
        if arrs and accumulate and not leading:
            # slices before the index array: supported when the index array has no duplicates (then add == set of sum)
            for _, it in arrs:
                assert it[1].numel() == it[1].unique().numel(), "jaxshim: duplicate indices in a mixed .at[].add key"
            new[tuple(idx)] = new[tuple(idx)] + val
        elif arrs and accumulate:
            ia = tuple(i for i in idx[:len(arrs)])
            shape = torch.broadcast_shapes(*[i.shape for i in ia]) + new.shape[len(arrs):]
            new.index_put_(ia, val.expand(shape), accumulate=True)
        elif accumulate:
            new[tuple(idx)] = new[tuple(idx)] + val
        else:
            new[tuple(idx)] = val
        return W(new)

    def set(self, value):
        return self._apply(value, False)

    def add(self, value):
        return self._apply(value, True)


def _binop(fn, reverse=False):
    def op(self, other):
        if isinstance(other, (dict, str)) or other is None:
            return NotImplemented
        a, b = self.t, T(other)
        return W(fn(b, a) if reverse else fn(a, b))
    return op


def _truediv(a, b):
    return torch.true_divide(a, b)


def _mul(a, b):
    if a.dtype == torch.bool and b.dtype == torch.bool:
        return a & b
    return a * b


def _pow(a, b):
    return torch.pow(a, b)


class Array:
    """Immutable array with jax.numpy's ndarray surface (only what the reference path touches)."""
    __slots__ = ("t",)
    __array_priority__ = 1000

    def __init__(self, t):
        self.t = t

    # ---- metadata
    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def dtype(self):
        return self.t.dtype

    @property
    def ndim(self):
        return self.t.dim()

    @property
    def size(self):
        return self.t.numel()

    @property
    def T(self):
        return W(self.t.permute(*reversed(range(self.t.dim()))))

    @property
    def at(self):
        return _At(self)

    def __len__(self):
        return self.t.shape[0]

    def __iter__(self):
        for i in range(self.t.shape[0]):
            yield W(self.t[i])

    def __repr__(self):
        return f"ShimArray({self.t.detach().cpu().numpy()!r})"

    def __array__(self, dtype=None, copy=None):
        a = self.t.detach().cpu().numpy()
        return a.astype(dtype) if dtype is not None else a

    def __float__(self):
        return float(self.t.item())

    def __int__(self):
        return int(self.t.item())

    def __index__(self):
        return int(self.t.item())

    def __bool__(self):
        return bool(self.t.item())

    def __hash__(self):
        return id(self)

    def item(self):
        return self.t.item()

    def block_until_ready(self):
        return self

    # ---- indexing (gather rules)
    def __getitem__(self, key):
        return W(self.t[_gather_key(self.t, key)])

    # ---- arithmetic
    __add__ = _binop(torch.add)
    __radd__ = _binop(torch.add, True)
    __sub__ = _binop(torch.sub)
    __rsub__ = _binop(torch.sub, True)
    __mul__ = _binop(_mul)
    __rmul__ = _binop(_mul, True)
    __truediv__ = _binop(_truediv)
    __rtruediv__ = _binop(_truediv, True)
    __floordiv__ = _binop(torch.floor_divide)
    __mod__ = _binop(torch.remainder)
    __pow__ = _binop(_pow)
    __rpow__ = _binop(_pow, True)
    __matmul__ = _binop(torch.matmul)
    __rmatmul__ = _binop(torch.matmul, True)
    __and__ = _binop(torch.bitwise_and)
    __rand__ = _binop(torch.bitwise_and, True)
    __or__ = _binop(torch.bitwise_or)
    __ror__ = _binop(torch.bitwise_or, True)
    __xor__ = _binop(torch.bitwise_xor)
    __lt__ = _binop(torch.lt)
    __le__ = _binop(torch.le)
    __gt__ = _binop(torch.gt)
    __ge__ = _binop(torch.ge)
    __eq__ = _binop(torch.eq)
    __ne__ = _binop(torch.ne)

    def __neg__(self):
        return W(-self.t)

    def __pos__(self):
        return self

    def __abs__(self):
        return W(self.t.abs())

    def __invert__(self):
        return W(~self.t)

    # ---- methods
    def astype(self, d):
        return W(self.t.to(_dtype(d)))

    def reshape(self, *shape, order="C"):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return W(self.t.reshape(tuple(int(s) for s in shape)))

    def repeat(self, repeats, axis=None):
        if axis is None:
            return W(self.t.flatten().repeat_interleave(int(repeats)))
        return W(self.t.repeat_interleave(int(repeats), dim=axis))

    def transpose(self, *axes):
        if len(axes) == 1 and isinstance(axes[0], (tuple, list)):
            axes = tuple(axes[0])
        if not axes:
            return self.T
        return W(self.t.permute(*axes))

    def swapaxes(self, a, b):
        return W(self.t.transpose(a, b))

    def squeeze(self, axis=None):
        return W(self.t.squeeze() if axis is None else self.t.squeeze(axis))

    def flatten(self):
        return W(self.t.flatten())

    def ravel(self):
        return W(self.t.flatten())

    def sum(self, axis=None, keepdims=False):
        t = self.t.to(torch.int32) if self.t.dtype == torch.bool else self.t
        return W(t.sum() if axis is None else t.sum(dim=axis, keepdim=keepdims))

    def prod(self, axis=None):
        return W(self.t.prod() if axis is None else self.t.prod(dim=axis))

    def mean(self, axis=None):
        return W(self.t.mean() if axis is None else self.t.mean(dim=axis))

    def max(self, axis=None):
        return W(self.t.max() if axis is None else self.t.amax(dim=axis))

    def min(self, axis=None):
        return W(self.t.min() if axis is None else self.t.amin(dim=axis))

    def clip(self, a_min=None, a_max=None):
        from .numpy import clip
        return clip(self, a_min, a_max)

    def dot(self, other):
        from .numpy import dot
        return dot(self, other)

    def trace(self, offset=0, axis1=0, axis2=1):
        return W(torch.diagonal(self.t, offset, axis1, axis2).sum(-1))

    def conj(self):
        return self

    def copy(self):
        return W(self.t.clone())

    def tolist(self):
        return self.t.tolist()
