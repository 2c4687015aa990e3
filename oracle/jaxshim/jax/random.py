"""jax.random stand-in.  PRNGKey / split / uniform / normal are the threefry2x32 streams of the jax 0.3.14 the reference
pins (unidom_b200.jaxrng: NumPy, pinned to the Random123 vectors and JAX's documented draws), so a reset of the unmodified
reference run under this shim draws the numbers a JAX run draws (normal's last bit aside, see jaxrng).  randint is
NOT jax's (nothing on the path draws from it)."""
import os as _os
import sys as _sys

import numpy as _np
import torch as _torch

from ._array import Array, T as _T, W as _W, float_dtype as _fd

_sys.path.insert(0, _os.path.abspath(_os.path.join(_os.path.dirname(__file__), "..", "..", "..")))
from unidom_b200 import jaxrng as _R  # noqa: E402

KeyArray = Array


def _key(key):
    return _np.asarray(_T(key).tolist(), dtype=_np.int64).astype(_np.uint32)


def PRNGKey(seed):
    return _W(_torch.from_numpy(_R.PRNGKey(seed).astype(_np.int64)))


def split(key, num=2):
    return _W(_torch.from_numpy(_R.split(_key(key), int(num)).astype(_np.int64)))


def uniform(key, shape=(), dtype=None, minval=0.0, maxval=1.0):
    return _W(_torch.from_numpy(_np.array(_R.uniform(_key(key), tuple(shape), minval, maxval), dtype=_np.float32)).to(_fd()))


def normal(key, shape=(), dtype=None):
    return _W(_torch.from_numpy(_np.array(_R.normal(_key(key), tuple(shape)), dtype=_np.float32)).to(_fd()))


def _rng(key):
    k = _key(key).tolist()
    return _np.random.Generator(_np.random.PCG64([int(k[0]), int(k[1])]))


def randint(key, shape, minval, maxval, dtype=None):
    return _W(_torch.from_numpy(_rng(key).integers(minval, maxval, tuple(shape))).to(_torch.int32))
