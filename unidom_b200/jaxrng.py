"""JAX's default PRNG (threefry2x32, the non-partitionable layout of jax 0.3.14 that the reference pins in
DaXBench/env.yml:62) in NumPy, so that resets and scene construction draw the SAME numbers as the reference:

  jax.random.PRNGKey / split / uniform / normal
  used at  core/envs/basic/cloth_env.py:181-185   reset: key, _ = split(key); x[..., [0, 2]] += normal(key, (2,)) * 0.05
           core/engine/mpm_simulator.py:89        add_box(material=0): uniform(conf.key, (n_points, 3))
           core/engine/mpm_simulator.py:169       reset_jax: split(key_global, batch_size)

Integer parts (the block cipher, key splitting, bits -> [0, 1) floats) are exact and pinned to the Random123 known-answer
vectors and to the values JAX's documentation prints (tests/test_jaxrng_cpu.py).  `normal` goes through erf_inv, which
XLA evaluates with the float32 polynomial restated in _erfinv_f32; its log1p is XLA's own, so the last bit of a normal
draw is not guaranteed (the documented draws are reproduced to 1e-7).
Host-side scene construction only (a few thousand numbers per reset): NumPy, not a kernel."""
import numpy as np

_U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
    return (x << _U32(r)) | (x >> _U32(32 - r))


def threefry2x32(key, x0, x1):
    """Threefry-2x32, 20 rounds.  key: (k0, k1) uint32; x0, x1: uint32 arrays of one shape.  Returns (y0, y1)."""
    with np.errstate(over="ignore"):
        k0, k1 = _U32(key[0]), _U32(key[1])
        ks = (k0, k1, _U32(k0 ^ k1 ^ _U32(0x1BD11BDA)))
        x0 = np.asarray(x0, _U32) + ks[0]
        x1 = np.asarray(x1, _U32) + ks[1]
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r) ^ x0
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + ks[(i + 2) % 3] + _U32(i + 1)
    return x0, x1


def _threefry_2x32_counts(key, counts):
    """jax._src.prng.threefry_2x32: the count vector is split in HALVES (first half -> lane 0, second -> lane 1), padded
    with one zero when its length is odd; the outputs are concatenated back."""
    counts = np.asarray(counts, _U32).ravel()
    odd = counts.size % 2
    if odd:
        counts = np.concatenate([counts, np.zeros(1, _U32)])
    h = counts.size // 2
    y0, y1 = threefry2x32(key, counts[:h], counts[h:])
    out = np.concatenate([y0, y1])
    return out[:-1] if odd else out


def PRNGKey(seed):
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], _U32)


def split(key, num=2):
    return _threefry_2x32_counts(key, np.arange(2 * num, dtype=_U32)).reshape(num, 2)


def random_bits(key, shape):
    n = int(np.prod(shape)) if len(shape) else 1
    return _threefry_2x32_counts(key, np.arange(n, dtype=_U32)).reshape(shape)


def uniform(key, shape=(), minval=0.0, maxval=1.0):
    """jax.random.uniform, float32: 23 random mantissa bits -> [1, 2) - 1, scaled, clamped from below."""
    bits = random_bits(key, tuple(shape))
    f = ((bits >> _U32(9)) | _U32(0x3F800000)).view(np.float32) - np.float32(1.0)
    lo, hi = np.float32(minval), np.float32(maxval)
    return np.maximum(lo, (f * np.float32(hi - lo) + lo).astype(np.float32))


def _erfinv_f32(x):
    """XLA's float32 erf_inv (Giles' single-precision approximation), evaluated op by op in float32."""
    x = np.asarray(x, np.float32)
    w = (-np.log1p((-(x * x)).astype(np.float32))).astype(np.float32)
    small = w < np.float32(5.0)
    ws = (w - np.float32(2.5)).astype(np.float32)
    wl = (np.sqrt(np.maximum(w, np.float32(5.0))).astype(np.float32) - np.float32(3.0)).astype(np.float32)
    cs = (2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503, -0.00417768164,
          0.246640727, 1.50140941)
    cl = (-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613, 0.00943887047,
          1.00167406, 2.83297682)
    ps = np.full_like(x, np.float32(cs[0]))
    pl = np.full_like(x, np.float32(cl[0]))
    for a, b in zip(cs[1:], cl[1:]):
        ps = (np.float32(a) + (ps * ws).astype(np.float32)).astype(np.float32)
        pl = (np.float32(b) + (pl * wl).astype(np.float32)).astype(np.float32)
    out = (np.where(small, ps, pl).astype(np.float32) * x).astype(np.float32)
    return np.where(np.abs(x) == 1, np.float32(np.inf) * x, out).astype(np.float32)


def normal(key, shape=()):
    """jax.random.normal, float32: sqrt(2) * erf_inv(uniform(key, shape, nextafter(-1, 0), 1))."""
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0), dtype=np.float32)
    u = uniform(key, shape, lo, 1.0)
    return (np.float32(np.sqrt(2)) * _erfinv_f32(u)).astype(np.float32)
