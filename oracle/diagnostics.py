"""``geomjax.rhat`` / ``geomjax.ess`` restated in NumPy (oracle; TEST INFRASTRUCTURE ONLY).

Follows geomjax/diagnostics.py:25-75 (potential_scale_reduction) and :78-209
(effective_sample_size: FFT autocovariance, Geyer initial positive + monotone sequence).
"""
from __future__ import annotations

import numpy as np
from scipy.fftpack import next_fast_len


def potential_scale_reduction(x, chain_axis=0, sample_axis=1):
    x = np.asarray(x)
    assert x.shape[chain_axis] > 1
    n = x.shape[sample_axis]
    m = x.mean(axis=sample_axis, keepdims=True)
    v = x.var(axis=sample_axis, ddof=1, keepdims=True)
    B = n * m.var(axis=chain_axis, ddof=1, keepdims=True)
    W = v.mean(axis=chain_axis, keepdims=True)
    return np.sqrt((B / W + n - 1) / n).squeeze()


def effective_sample_size(x, chain_axis=0, sample_axis=1):
    x = np.asarray(x)
    shape = x.shape
    sample_axis = sample_axis if sample_axis >= 0 else len(shape) + sample_axis
    M, N = shape[chain_axis], shape[sample_axis]
    assert M > 1
    mean_c = x.mean(axis=sample_axis, keepdims=True)
    cen = x - mean_c
    m = next_fast_len(2 * N)
    f = np.fft.rfft(cen, n=m, axis=sample_axis)
    f = f * np.conjugate(f)
    acov = np.fft.irfft(f, n=m, axis=sample_axis)
    acov = np.take(acov, np.arange(N), axis=sample_axis) / N
    mean_acov = acov.mean(chain_axis, keepdims=True)
    return _geyer(mean_acov, mean_c.var(axis=chain_axis, ddof=1, keepdims=True), M, N, sample_axis)


def _geyer(mean_acov, var_of_means, M, N, sample_axis):
    """diagnostics.py:133-209 given the chain-averaged autocovariance and the variance of the
    per-chain means (both keepdims arrays)."""
    mean_var0 = np.take(mean_acov, [0], axis=sample_axis) * N / (N - 1.0)
    weighted_var = mean_var0 * (N - 1.0) / N + var_of_means
    n_even = N - N % 2
    tp1 = np.take(mean_acov, np.arange(1, n_even), axis=sample_axis)
    rho = np.concatenate([np.ones_like(mean_var0), 1.0 - (mean_var0 - tp1) / weighted_var],
                         axis=sample_axis)
    rho = np.moveaxis(rho, sample_axis, 0)
    even, odd = rho[0::2], rho[1::2]
    mask0 = (even + odd) > 0.0
    carry = np.ones_like(mask0[0])
    max_t = np.zeros(mask0[0].shape, dtype=int)
    mask = np.zeros_like(mask0)
    for t in range(mask0.shape[0]):
        carry = carry & mask0[t]
        max_t = np.where(carry, t, max_t)
        mask[t] = carry
    idx = np.indices(max_t.shape)
    # JAX semantics: out-of-bounds gather clamps, out-of-bounds scatter is dropped
    in_bounds = (max_t + 1) < even.shape[0]
    indices = tuple([np.minimum(max_t + 1, even.shape[0] - 1)] + [idx[i] for i in range(max_t.ndim)])
    odd = np.where(mask, odd, 0.0)
    mask_even = mask.copy()
    mask_even[indices] = np.where(in_bounds, even[indices] > 0, mask_even[indices])
    even = np.where(mask_even, even, 0.0)
    s = even + odd
    prev = s[0]
    upd_mask = np.zeros_like(mask0)
    upd_val = np.zeros_like(s)
    for t in range(s.shape[0]):
        um = s[t] > prev
        prev = np.where(um, prev, s[t])
        upd_mask[t], upd_val[t] = um, prev
    even_f = np.where(upd_mask, upd_val / 2.0, even)
    odd_f = np.where(upd_mask, upd_val / 2.0, odd)
    ess_raw = M * N
    tau = -1.0 + 2.0 * np.sum(even_f + odd_f, axis=0) - even_f[indices]
    tau = np.maximum(tau, 1 / np.log10(ess_raw))
    return (ess_raw / tau).squeeze()
