"""``rhat`` / ``ess`` with the reference signatures (geomjax/diagnostics.py:25-75, :78-209),
computed on the GPU as chain-summed sufficient statistics and finalised on the host.

When the chain axis is sharded over ranks (one process per GPU), pass ``process_group`` (or
rely on the default group): the statistics are plain sums over chains, so ONE all-reduce(sum)
of 3D+1 (R-hat) / num_lags*D (ESS) float64 values gives every rank the global answer.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as N

__all__ = ["potential_scale_reduction", "effective_sample_size", "rhat", "ess"]


def _canon(x: torch.Tensor, chain_axis: int, sample_axis: int) -> torch.Tensor:
    """-> contiguous (T, C, D) with trailing dims flattened."""
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise N.NativeError("diagnostics operate on CUDA tensors only (no CPU fallback)")
    nd = x.ndim
    chain_axis %= nd
    sample_axis %= nd
    if chain_axis == sample_axis:
        raise ValueError("chain_axis and sample_axis must differ")
    rest = [a for a in range(nd) if a not in (chain_axis, sample_axis)]
    x = x.permute(sample_axis, chain_axis, *rest)
    out_shape = tuple(x.shape[2:])
    return x.reshape(x.shape[0], x.shape[1], -1).contiguous(), out_shape


def _dtype(x):
    return N.F32 if x.dtype == torch.float32 else N.F64


def _allreduce(t, process_group):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=process_group)
    return t


def _rhat_stats(x3, process_group):
    T_, C_, D = x3.shape
    stats = torch.empty(3 * D + 1, dtype=torch.float64, device=x3.device)
    with torch.cuda.device(x3.device):
        N.check(N.lib().gb200_rhat_partial(N.ptr(x3), T_, C_, D, N.ptr(stats), _dtype(x3), N.stream_ptr()))
    return _allreduce(stats, process_group)


def potential_scale_reduction(input_array, chain_axis: int = 0, sample_axis: int = 1, process_group=None):
    x3, out_shape = _canon(input_array, chain_axis, sample_axis)
    T_, C_, D = x3.shape
    stats = _rhat_stats(x3, process_group).cpu().numpy()
    assert stats[3 * D] > 1, "potential_scale_reduction as implemented only works for two or more chains."
    out = np.empty(D, np.float64)
    N.check(N.lib().gb200_rhat_finalize(stats.ctypes.data_as(C.POINTER(C.c_double)), T_, D,
                                        out.ctypes.data_as(C.POINTER(C.c_double))))
    return torch.from_numpy(out.reshape(out_shape).astype(np.float32 if x3.dtype == torch.float32 else np.float64))


def effective_sample_size(input_array, chain_axis: int = 0, sample_axis: int = 1, process_group=None,
                          initial_lags: int = 64):
    """Geyer's truncation makes every lag beyond the first non-positive pair irrelevant, so the
    autocovariance is computed for ``initial_lags`` lags first and doubled until no dimension
    is still positive at the last computed pair (identical result to the all-lags FFT)."""
    x3, out_shape = _canon(input_array, chain_axis, sample_axis)
    T_, C_, D = x3.shape
    stats = _rhat_stats(x3, process_group).cpu().numpy()
    c_total = int(round(stats[3 * D]))
    assert c_total > 1, "effective_sample_size as implemented only works for two or more chains."
    lags = min(max(4, initial_lags), T_)
    while True:
        acov = torch.empty(lags * D, dtype=torch.float64, device=x3.device)
        with torch.cuda.device(x3.device):
            N.check(N.lib().gb200_ess_partial(N.ptr(x3), T_, C_, D, lags, N.ptr(acov), _dtype(x3), N.stream_ptr()))
        acov = _allreduce(acov, process_group).cpu().numpy()
        out = np.empty(D, np.float64)
        trunc = np.zeros(D, np.uint8)
        N.check(N.lib().gb200_ess_finalize(acov.ctypes.data_as(C.POINTER(C.c_double)),
                                           stats.ctypes.data_as(C.POINTER(C.c_double)), T_, c_total, D, lags,
                                           out.ctypes.data_as(C.POINTER(C.c_double)),
                                           trunc.ctypes.data_as(C.POINTER(C.c_uint8))))
        if not trunc.any() or lags >= T_:
            break
        lags = min(2 * lags, T_)
    return torch.from_numpy(out.reshape(out_shape).astype(np.float32 if x3.dtype == torch.float32 else np.float64))


rhat = potential_scale_reduction
ess = effective_sample_size
