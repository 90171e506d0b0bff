"""jax.random-compatible key handling on the GPU (threefry2x32, bit-exact with JAX's
``legacy`` mode by default; see oracle/prng.py for the restated specification)."""
from __future__ import annotations

import numpy as np
import torch

from . import _native as N

_MODE = N.LEGACY


def set_threefry_partitionable(flag: bool):
    """Equivalent of ``jax.config.update('jax_threefry_partitionable', flag)``."""
    global _MODE
    _MODE = N.PARTITIONABLE if flag else N.LEGACY


def threefry_mode() -> int:
    return _MODE


def PRNGKey(seed: int) -> np.ndarray:
    """Raw key data ``[hi32(seed), lo32(seed)]`` (host array; root keys are tiny).  As in JAX with x64 disabled (the
    reference's configuration) a negative seed is an int32: ``PRNGKey(-1) == [0, 0xFFFFFFFF]``; seeds >= 2**32 keep
    their high word (JAX's x64 behaviour; JAX without x64 raises for them)."""
    seed = int(seed)
    if seed < 0:
        if seed < -(1 << 31):
            raise OverflowError("negative seeds must fit int32 (JAX with x64 disabled)")
        return np.array([0, seed & 0xFFFFFFFF], dtype=np.uint32)
    seed &= 0xFFFFFFFFFFFFFFFF
    return np.array([seed >> 32, seed & 0xFFFFFFFF], dtype=np.uint32)


key = PRNGKey


def _keys_tensor(keys, device=None) -> torch.Tensor:
    if isinstance(keys, np.ndarray):
        keys = torch.from_numpy(np.ascontiguousarray(keys.astype(np.uint32)))
    if not keys.is_cuda:
        keys = keys.to(device or "cuda")
    if keys.dtype not in (torch.uint32, torch.int32):
        raise TypeError("keys must be uint32")
    return keys.contiguous()


def split(keys, num: int = 2, device=None) -> torch.Tensor:
    """``jax.random.split`` for a batch of keys: (..., 2) -> (..., num, 2) uint32 on the GPU."""
    k = _keys_tensor(keys, device)
    n = k.numel() // 2
    out = torch.empty(k.shape[:-1] + (num, 2), dtype=torch.uint32, device=k.device)
    with torch.cuda.device(k.device):
        N.check(N.lib().gb200_threefry_split(N.ptr(k), N.ptr(out), n, num, _MODE, N.stream_ptr()))
    return out


def _draw(fn, keys, count, dtype, device):
    k = _keys_tensor(keys, device)
    n = k.numel() // 2
    out = torch.empty(k.shape[:-1] + (count,), dtype=dtype, device=k.device)
    with torch.cuda.device(k.device):
        N.check(fn(N.ptr(k), N.ptr(out), n, count, _MODE, N.stream_ptr()))
    return out


def bits(keys, count: int, device=None) -> torch.Tensor:
    return _draw(N.lib().gb200_random_bits, keys, count, torch.uint32, device)


def uniform(keys, count: int = 1, device=None) -> torch.Tensor:
    return _draw(N.lib().gb200_uniform_f32, keys, count, torch.float32, device)


def normal(keys, count: int, device=None) -> torch.Tensor:
    return _draw(N.lib().gb200_normal_f32, keys, count, torch.float32, device)


def randint(keys, minval: int, maxval: int, device=None) -> torch.Tensor:
    """``jax.random.randint(key, (), minval, maxval)`` for a batch of keys (..., 2) -> (...,) int32
    (jax/_src/random.py ``_randint``: two 32-bit draws from ``split(key)``, combined modulo the span with the
    multiplier 2**32 % span so that the result is uniform over the 64-bit draw).  The default
    ``integration_steps_fn`` of the dynamic kernels (rmhmc/rmhmc.py:183, lmcmc/lmc.py:189)."""
    k = _keys_tensor(keys, device)
    ks = split(k, 2)
    hi = bits(ks[..., 0, :].contiguous(), 1)[..., 0].to(torch.int64)
    lo = bits(ks[..., 1, :].contiguous(), 1)[..., 0].to(torch.int64)
    span = max(int(maxval) - int(minval), 1) if maxval > minval else 1
    mult = ((1 << 16) % span) ** 2 % span
    M32 = (1 << 32) - 1
    off = ((((hi % span) * mult) & M32) + (lo % span)) & M32  # uint32 wrap-around, as in lax
    return (int(minval) + off % span).to(torch.int32)


def chain_keys(root_key, t: int, total_transitions: int, num_chains: int, chain_offset: int = 0,
               total_chains: int | None = None, device="cuda") -> torch.Tensor:
    """``split(split(root, T)[t], C_total)[offset:offset+C]`` (examples/funnel/main.py:18,22)."""
    import ctypes as C
    total_chains = num_chains if total_chains is None else total_chains
    root = (C.c_uint32 * 2)(int(root_key[0]), int(root_key[1]))
    out = torch.empty((num_chains, 2), dtype=torch.uint32, device=device)
    with torch.cuda.device(out.device):
        N.check(N.lib().gb200_chain_keys(root, t, total_transitions, chain_offset, total_chains,
                                         N.ptr(out), num_chains, _MODE, N.stream_ptr()))
    return out
