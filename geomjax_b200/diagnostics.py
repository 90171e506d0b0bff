"""``rhat`` / ``ess`` with the reference signatures (geomjax/diagnostics.py:25-75, :78-209),
computed on the GPU as chain-summed sufficient statistics and finalised on the host.

When the chain axis is sharded over ranks (one process per GPU), pass ``process_group`` (or
rely on the default group): the statistics are plain sums over chains, so ONE all-reduce(sum)
of 3D+1 (R-hat) / num_lags*D (ESS) float64 values gives every rank the global answer.

The results are O(D) host-finalised values and are returned as CPU tensors (the reference returns device arrays);
call ``.to(device)`` to combine them with CUDA tensors.  The batch ``ess`` holds a whole series in shared memory
(T <= ~51k samples); longer runs go through ``StreamingDiagnostics`` / ``sample_streaming`` below.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as N

__all__ = ["potential_scale_reduction", "effective_sample_size", "rhat", "ess", "StreamingDiagnostics", "sample_streaming"]


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


class StreamingDiagnostics:
    """``rhat`` / ``ess`` of a run whose samples are never held as one (T, C, D) tensor: feed blocks ``(Tb, C, D)``
    with ``update`` (e.g. the reused sample buffer of successive fused launches), read ``rhat()`` / ``ess()`` at the
    end.  Same sufficient statistics as the batch functions above (geomjax/diagnostics.py:25-209), same ONE
    all-reduce per statistic when chains are sharded over ranks.  Autocovariances are kept for ``max_lags`` lags
    (memory: 8 * max_lags bytes per chain and dimension); if Geyer's sequence is still positive there, ``ess``
    raises unless ``allow_truncated`` (the batch version would have doubled the lag count instead)."""

    def __init__(self, chains: int, dim: int, max_lags: int = 64, device="cuda", dtype=torch.float32):
        self.C, self.D, self.K = int(chains), int(dim), (int(max_lags) + 7) // 8 * 8
        self.device, self.dtype, self.T = torch.device(device), dtype, 0
        n = int(N.lib().gb200_stream_diag_workspace(self.C, self.D, self.K))
        if n < 0:
            raise ValueError("StreamingDiagnostics needs chains >= 1, dim >= 1, max_lags >= 8")
        self._ws = torch.empty(n + 256, dtype=torch.uint8, device=self.device)
        self._base = (self._ws.data_ptr() + 255) // 256 * 256

    def update(self, block: torch.Tensor):
        if not block.is_cuda or block.ndim != 3 or block.shape[1:] != (self.C, self.D) or block.dtype != self.dtype:
            raise ValueError(f"block must be a CUDA {self.dtype} tensor of shape (Tb, {self.C}, {self.D})")
        block = block.contiguous()
        step = 512  # samples per launch (shared-memory tile)
        with torch.cuda.device(self.device):
            for t0 in range(0, block.shape[0], step):
                part = block[t0:t0 + step]
                N.check(N.lib().gb200_stream_diag_update(C.c_void_p(self._base), N.ptr(part), part.shape[0], self.C, self.D,
                                                         self.K, self.T, _dtype(block), N.stream_ptr()))
                self.T += part.shape[0]
        return self

    def _partial(self, lags, process_group):
        stats = torch.empty(3 * self.D + 1, dtype=torch.float64, device=self.device)
        acov = torch.empty(max(lags, 1) * self.D, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            N.check(N.lib().gb200_stream_diag_partial(C.c_void_p(self._base), self.T, self.C, self.D, self.K, lags,
                                                      N.ptr(stats), N.ptr(acov) if lags > 0 else None, N.stream_ptr()))
        stats = _allreduce(stats, process_group).cpu().numpy()
        return stats, (_allreduce(acov, process_group).cpu().numpy() if lags > 0 else None)

    def rhat(self, process_group=None) -> torch.Tensor:
        stats, _ = self._partial(0, process_group)
        assert stats[3 * self.D] > 1, "potential_scale_reduction as implemented only works for two or more chains."
        out = np.empty(self.D, np.float64)
        N.check(N.lib().gb200_rhat_finalize(stats.ctypes.data_as(C.POINTER(C.c_double)), self.T, self.D,
                                            out.ctypes.data_as(C.POINTER(C.c_double))))
        return torch.from_numpy(out.astype(np.float32 if self.dtype == torch.float32 else np.float64))

    def ess(self, process_group=None, allow_truncated: bool = False) -> torch.Tensor:
        lags = min(self.K, self.T)
        stats, acov = self._partial(lags, process_group)
        c_total = int(round(stats[3 * self.D]))
        assert c_total > 1, "effective_sample_size as implemented only works for two or more chains."
        out = np.empty(self.D, np.float64)
        trunc = np.zeros(self.D, np.uint8)
        N.check(N.lib().gb200_ess_finalize(acov.ctypes.data_as(C.POINTER(C.c_double)),
                                           stats.ctypes.data_as(C.POINTER(C.c_double)), self.T, c_total, self.D, lags,
                                           out.ctypes.data_as(C.POINTER(C.c_double)),
                                           trunc.ctypes.data_as(C.POINTER(C.c_uint8))))
        self.truncated = trunc.astype(bool)
        if trunc.any() and lags < self.T and not allow_truncated:
            raise RuntimeError(f"autocorrelation still positive at lag {lags} for {int(trunc.sum())} dimension(s): "
                               "construct StreamingDiagnostics with a larger max_lags")
        return torch.from_numpy(out.astype(np.float32 if self.dtype == torch.float32 else np.float64))


def sample_streaming(step_fn, rng_key, state, num_samples: int, *, block: int = 64, max_lags: int = 64,
                     chain_offset: int = 0, total_chains=None, on_block=None):
    """The driver loop of examples/funnel/main.py:7-25 followed by its rhat / ess (:77-80) with O(block) sample
    memory: ``num_samples`` transitions in fused launches of ``block`` transitions whose sample buffer is reused and
    folded into a ``StreamingDiagnostics``.  Bit-identical chains to one ``run_fused`` of ``num_samples`` (same key
    tree).  ``on_block(first, samples_block)`` may consume the block (e.g. accumulate posterior moments).
    Returns (final state, StreamingDiagnostics, mean acceptance rate per chain [C])."""
    from .samplers import run_fused
    q = state[0]
    C_, D = q.shape
    diag = StreamingDiagnostics(C_, D, max_lags, q.device, q.dtype)
    buf = torch.empty((block, C_, D), dtype=q.dtype, device=q.device)
    acc_sum = torch.zeros((C_,), dtype=q.dtype, device=q.device)
    acc = torch.empty((C_,), dtype=q.dtype, device=q.device)
    done = 0
    while done < num_samples:
        tb = min(block, num_samples - done)
        out = buf[:tb]
        state, _, a = run_fused(step_fn, rng_key, state, tb, first=done, total=num_samples, chain_offset=chain_offset,
                                total_chains=total_chains, out_samples=out, return_accept="mean", out_accept=acc)
        acc_sum += a * tb
        diag.update(out)
        if on_block is not None:
            on_block(done, out)
        done += tb
    return state, diag, acc_sum / max(num_samples, 1)
