"""JAX's threefry2x32 PRNG restated in NumPy (oracle; test infrastructure only).

The arithmetic lives in a third-party dependency of the reference that is absent here:
``jax`` / ``jaxlib`` (``/root/reference/requirements.txt:7-8``, ``jax>=0.4.16``, no lock
file; effective window 0.4.16 .. ~0.4.34 because ``geomjax/progress_bar.py:19`` imports
``jax.experimental.host_callback``).  In that window ``jax_threefry_partitionable``
defaults to False, so ``mode="legacy"`` is the bit-exact target; ``"partitionable"`` is
provided as well.

Reference call sites this module serves:
  ``jax.random.split(rng_key, 2)``      rmhmc/rmhmc.py:158, lmcmc/lmc.py:164, lmcmonge/lmc.py:196
  ``jax.random.normal`` via             util.py:81-82 (generate_gaussian_noise)
  ``jax.random.bernoulli``              mcmc/proposal.py:178
  ``split(key, num_samples|num_chains)`` examples/funnel/main.py:18,22

Published algorithm restated (Salmon et al., Random123 threefry2x32-20; jax/_src/prng.py
``threefry_2x32``, ``_threefry_split``, ``_threefry_random_bits``; jax/_src/random.py
``_uniform``, ``_normal_real``, ``_bernoulli``; XLA ``ErfInv32``).
"""
from __future__ import annotations

import numpy as np

LEGACY = "legacy"
PARTITIONABLE = "partitionable"

_ROT_A = (13, 15, 26, 6)
_ROT_B = (17, 29, 16, 24)
_U32 = np.uint32


def _rotl(x, r):
    return (x << _U32(r)) | (x >> _U32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    """One threefry2x32-20 block per element. All args uint32 arrays (broadcastable)."""
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, dtype=_U32)
        k1 = np.asarray(k1, dtype=_U32)
        x0 = np.array(x0, dtype=_U32, copy=True)
        x1 = np.array(x1, dtype=_U32, copy=True)
        ks = (k0, k1, k0 ^ k1 ^ _U32(0x1BD11BDA))
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for g in range(1, 6):
            for r in (_ROT_A if g % 2 == 1 else _ROT_B):
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[g % 3]
            x1 = x1 + ks[(g + 1) % 3] + _U32(g)
    return x0, x1


def key(seed: int) -> np.ndarray:
    """``jax.random.PRNGKey(seed)`` / ``jax.random.key(seed)`` raw key data: [hi32, lo32]."""
    seed = int(seed)
    if seed < 0:  # x64 disabled (the reference's configuration): the seed is an int32, its high word is 0
        return np.array([0, seed & 0xFFFFFFFF], dtype=_U32)
    seed &= 0xFFFFFFFFFFFFFFFF
    return np.array([seed >> 32, seed & 0xFFFFFFFF], dtype=_U32)


PRNGKey = key


def _hash_counts_legacy(k, count):
    """jax/_src/prng.py ``threefry_2x32(keypair, count)``: pad odd counts with one 0,
    hash (first half, second half) pairwise, concatenate, drop the pad.
    ``k``: (..., 2) uint32; ``count``: (n,) uint32.  Returns (..., n)."""
    k = np.asarray(k, dtype=_U32)
    count = np.asarray(count, dtype=_U32)
    n = count.shape[0]
    if n % 2:
        count = np.concatenate([count, np.zeros(1, _U32)])
    h = count.shape[0] // 2
    o0, o1 = threefry2x32(k[..., 0:1], k[..., 1:2], count[:h], count[h:])
    out = np.concatenate([o0, o1], axis=-1)
    return out[..., :n]


def split(k, num: int = 2, mode: str = LEGACY) -> np.ndarray:
    """``jax.random.split``.  ``k``: (..., 2) -> (..., num, 2)."""
    k = np.asarray(k, dtype=_U32)
    if mode == LEGACY:
        bits = _hash_counts_legacy(k, np.arange(2 * num, dtype=_U32))
        return bits.reshape(k.shape[:-1] + (num, 2))
    o0, o1 = threefry2x32(k[..., 0:1], k[..., 1:2], np.zeros(num, _U32),
                          np.arange(num, dtype=_U32))
    return np.stack([o0, o1], axis=-1)


def split_index(k, num: int, idx, mode: str = LEGACY) -> np.ndarray:
    """``split(k, num)[idx]`` without materialising the other ``num - 1`` keys.
    ``k``: (2,), ``idx``: integer array -> idx.shape + (2,)."""
    k = np.asarray(k, dtype=_U32)
    idx = np.asarray(idx, dtype=np.int64)
    if mode == PARTITIONABLE:
        o0, o1 = threefry2x32(k[0], k[1], np.zeros(idx.shape, _U32), idx.astype(_U32))
        return np.stack([o0, o1], axis=-1)

    def bit(j):  # element j of random_bits(k, (2*num,)) in legacy mode; 2*num is even
        lo = j < num
        c0 = np.where(lo, j, j - num).astype(_U32)
        o0, o1 = threefry2x32(k[0], k[1], c0, c0 + _U32(num))
        return np.where(lo, o0, o1)

    return np.stack([bit(2 * idx), bit(2 * idx + 1)], axis=-1)


def random_bits(k, shape, mode: str = LEGACY) -> np.ndarray:
    """32-bit ``jax.random.bits``-style raw draw.  ``k``: (..., 2) -> (...,) + shape."""
    k = np.asarray(k, dtype=_U32)
    shape = tuple(int(s) for s in np.atleast_1d(shape)) if shape != () else ()
    size = int(np.prod(shape)) if shape else 1
    if mode == LEGACY:
        bits = _hash_counts_legacy(k, np.arange(size, dtype=_U32))
    else:
        o0, o1 = threefry2x32(k[..., 0:1], k[..., 1:2], np.zeros(size, _U32),
                              np.arange(size, dtype=_U32))
        bits = o0 ^ o1
    return bits.reshape(k.shape[:-1] + shape)


def _bits_to_unit_float(bits) -> np.ndarray:
    """jax/_src/random.py ``_uniform``: mantissa trick -> float32 in [0, 1)."""
    fb = (bits >> _U32(9)) | _U32(0x3F800000)
    return fb.view(np.float32) - np.float32(1.0)


def uniform(k, shape=(), mode: str = LEGACY, minval=0.0, maxval=1.0) -> np.ndarray:
    f = _bits_to_unit_float(random_bits(k, shape, mode))
    lo = np.float32(minval)
    hi = np.float32(maxval)
    return np.maximum(lo, f * (hi - lo) + lo).astype(np.float32)


_ERFINV_LT5 = np.array([2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06,
                        0.00021858087, -0.00125372503, -0.00417768164, 0.246640727,
                        1.50140941], dtype=np.float32)
_ERFINV_GE5 = np.array([-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844,
                        0.00573950773, -0.0076224613, 0.00943887047, 1.00167406,
                        2.83297682], dtype=np.float32)


_LG1, _LG2, _LG3, _LG4 = (np.float32(0xAAAAAA / 2.0 ** 24), np.float32(0xCCCE13 / 2.0 ** 25),
                          np.float32(0x91E9EE / 2.0 ** 25), np.float32(0xF89E26 / 2.0 ** 26))
_LN2_HI, _LN2_LO = np.float32(6.9313812256e-01), np.float32(9.0580006145e-06)


def log1p_f32(x) -> np.ndarray:
    """float32 log1p, restated from the public-domain fdlibm/musl ``log1pf`` (argument reduction of
    1+x to [sqrt(2)/2, sqrt(2)) with the rounding error of 1+x carried as a correction term, degree-4
    polynomial in s^2, s = f/(2+f); < 1 ulp).  Every operation is an individually rounded float32
    op, so the CUDA kernels (csrc/common.cuh::log1p_f32) reproduce it bit for bit.  It agrees with
    the correctly rounded log1p on ~94% of the erfinv domain and is within 1 ulp elsewhere -- the same
    class of approximation XLA itself uses (XLA's log1p is not correctly rounded either).
    Domain used here: -1 < x <= 0."""
    f32, u32 = np.float32, np.uint32
    x = np.asarray(x, f32)
    ix = x.view(u32)
    with np.errstate(all="ignore"):
        small = (ix < u32(0x3ED413D0)) | ((ix >> u32(31)) == 1)          # 1 + x < sqrt(2)
        tiny = small & ((ix << u32(1)) < u32(0x67000000))                # |x| < 2^-24: log1p(x) = x
        k0 = small & (ix <= u32(0xBE95F619))                             # sqrt(2)/2 <= 1 + x: no reduction
        u = (f32(1) + x).astype(f32)
        iu = u.view(u32) + u32(0x3F800000 - 0x3F3504F3)
        k = (iu >> u32(23)).astype(np.int32) - 0x7F
        c = np.where(k >= 2, (f32(1) - (u - x).astype(f32)).astype(f32),
                     (x - (u - f32(1)).astype(f32)).astype(f32)).astype(f32)
        c = np.where(k < 25, (c / u).astype(f32), f32(0)).astype(f32)
        f = (((iu & u32(0x007FFFFF)) + u32(0x3F3504F3)).view(f32) - f32(1)).astype(f32)
        k = np.where(k0, 0, k)
        c = np.where(k0, f32(0), c).astype(f32)
        f = np.where(k0, x, f).astype(f32)
        s = (f / (f32(2) + f).astype(f32)).astype(f32)
        z = (s * s).astype(f32)
        w = (z * z).astype(f32)
        t1 = (w * (_LG2 + (w * _LG4).astype(f32)).astype(f32)).astype(f32)
        t2 = (z * (_LG1 + (w * _LG3).astype(f32)).astype(f32)).astype(f32)
        R = (t2 + t1).astype(f32)
        hfsq = ((f32(0.5) * f).astype(f32) * f).astype(f32)
        dk = k.astype(f32)
        r = (s * (hfsq + R).astype(f32)).astype(f32)
        r = (r + ((dk * _LN2_LO).astype(f32) + c).astype(f32)).astype(f32)
        r = (r - hfsq).astype(f32)
        r = (r + f).astype(f32)
        r = (r + (dk * _LN2_HI).astype(f32)).astype(f32)
    return np.where(tiny, x, r).astype(f32)


def erfinv_f32(x) -> np.ndarray:
    """XLA ``ErfInv32`` (Giles' single-precision polynomial): w = -log1p(-x*x); two degree-8 Horner
    branches; p*x.  ``log1p`` is ``log1p_f32`` above; every op is a separately rounded float32 op (no
    FMA contraction) -- the CUDA path does exactly the same, so the two agree bit for bit; XLA:CPU's
    own log1p may differ from this by <= 1 ulp."""
    x = np.asarray(x, dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (x * x).astype(np.float32)
        w = (-log1p_f32(-t)).astype(np.float32)
        lt = w < np.float32(5.0)
        w = np.where(lt, w - np.float32(2.5), np.sqrt(w) - np.float32(3.0)).astype(np.float32)
        p = np.where(lt, _ERFINV_LT5[0], _ERFINV_GE5[0]).astype(np.float32)
        for i in range(1, 9):
            c = np.where(lt, _ERFINV_LT5[i], _ERFINV_GE5[i]).astype(np.float32)
            p = (c + (p * w).astype(np.float32)).astype(np.float32)
        r = (p * x).astype(np.float32)
        r = np.where(np.abs(x) == np.float32(1.0), x * np.finfo(np.float32).max, r)
    return r.astype(np.float32)


def normal(k, shape=(), mode: str = LEGACY) -> np.ndarray:
    """``jax.random.normal(key, shape, float32)`` (jax/_src/random.py ``_normal_real``)."""
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0))
    u = uniform(k, shape, mode, minval=lo, maxval=1.0)
    return (np.float32(np.sqrt(2.0)) * erfinv_f32(u)).astype(np.float32)


def bernoulli(k, p, mode: str = LEGACY):
    """``jax.random.bernoulli(key, p)`` = ``uniform(key, shape(p)) < p``; returns (accept, u)."""
    p = np.asarray(p, dtype=np.float32)
    k = np.asarray(k, dtype=_U32)
    u = uniform(k, (), mode)
    return u < p, u


def randint(k, minval, maxval, mode: str = LEGACY) -> np.ndarray:
    """``jax.random.randint(key, (), minval, maxval)`` per key (jax/_src/random.py ``_randint``): two 32-bit draws from
    ``split(key)``, combined modulo the span with the multiplier 2**32 % span (uint32 wrap-around arithmetic).  The
    default ``integration_steps_fn`` of the dynamic kernels (rmhmc/rmhmc.py:183).  Not pinned against a JAX output
    (none is documented): restated from the published algorithm."""
    k = np.asarray(k, dtype=_U32)
    ks = split(k, 2, mode)
    hi = random_bits(ks[..., 0, :], (), mode).astype(np.uint64)
    lo = random_bits(ks[..., 1, :], (), mode).astype(np.uint64)
    span = np.uint64(max(int(maxval) - int(minval), 1) if maxval > minval else 1)
    mult = np.uint64(((1 << 16) % int(span)) ** 2 % int(span))
    m32 = np.uint64(0xFFFFFFFF)
    off = ((((hi % span) * mult) & m32) + (lo % span)) & m32
    return (int(minval) + (off % span).astype(np.int64)).astype(np.int32)
