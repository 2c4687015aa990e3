"""jax.random stand-in.  NOT threefry: keys are (2,) integer arrays advanced by a splitmix-style hash and
draws come from numpy's PCG64 seeded by the key.  The step only passes keys through
(cloth_simulator.py:172 splits and discards), so no draw reaches the golden vectors of the hot path."""
import numpy as _np
import torch as _torch

from ._array import Array, T as _T, W as _W, float_dtype as _fd

KeyArray = Array


def PRNGKey(seed):
    return _W(_torch.tensor([0, int(seed) & 0xFFFFFFFF], dtype=_torch.int64))


def _mix(a):
    a = (int(a) + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    a = ((a ^ (a >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    a = ((a ^ (a >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return (a ^ (a >> 31)) & 0xFFFFFFFF


def split(key, num=2):
    k = _T(key).tolist()
    base = (int(k[0]) << 32) | int(k[1])
    out = [[_mix(base + 2 * i + 1), _mix(base + 2 * i + 2)] for i in range(int(num))]
    return _W(_torch.tensor(out, dtype=_torch.int64))


def _rng(key):
    k = _T(key).tolist()
    return _np.random.Generator(_np.random.PCG64([int(k[0]), int(k[1])]))


def uniform(key, shape=(), dtype=None, minval=0.0, maxval=1.0):
    u = _rng(key).random(tuple(shape))
    return _W(_torch.from_numpy(_np.asarray(u * (maxval - minval) + minval)).to(_fd()))


def normal(key, shape=(), dtype=None):
    return _W(_torch.from_numpy(_np.asarray(_rng(key).standard_normal(tuple(shape)))).to(_fd()))


def randint(key, shape, minval, maxval, dtype=None):
    return _W(_torch.from_numpy(_rng(key).integers(minval, maxval, tuple(shape))).to(_torch.int32))
