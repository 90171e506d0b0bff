"""Warm-up drivers with the reference signatures, batched over chains.

  step_size_adaptation  geomjax/adaptation/step_size_adaptation.py:100-203 (rmhmc / lmc / lmcmonge)
  window_adaptation     geomjax/adaptation/window_adaptation.py:245-450 (+ base :45-243, schedule
                        :360-450, Welford geomjax/adaptation/mass_matrix.py:60-239) for lmcmonge's
                        diagonal ``inverse_mass_matrix`` (for rmhmc/lmc the reference itself passes
                        the mass matrix where ``metric_fn`` is expected -- SURVEY F9)
  dual_averaging        geomjax/optimizers/dual_averaging.py:53-129

``run(rng_key, position, num_steps)`` is the ``jax.vmap`` of the reference's single-chain ``run``:
``rng_key`` is ``(C, 2)`` (one key per chain; a single ``(2,)`` key is first ``split`` into C),
``position`` is ``(C, D)``; every chain adapts its own step size (and mass matrix).  The per-chain
dual-averaging update runs as the epilogue of the fused transition kernel
(``gb200_run_opts.dual_averaging``), so a warm-up transition is still ONE launch.

NEW (not in the reference; modelled on the cross-chain statistics of
``chees_adaptation_riemanian.py:145-166``): ``pooled=True`` adapts ONE step size for all chains from
the mean acceptance rate, all-reduced over ranks when ``torch.distributed`` is initialised.
"""
from __future__ import annotations

from typing import NamedTuple

import numpy as np
import torch

from . import _native as N
from . import random as grandom
from .base import AdaptationAlgorithm, AdaptationResults
from .samplers import _Engine, _merge_target, lmc, lmcmonge, rmhmc

__all__ = ["step_size_adaptation", "window_adaptation", "dual_averaging", "build_schedule"]


class DualAveragingState(NamedTuple):  # optimizers/dual_averaging.py:40-50
    log_x: torch.Tensor
    log_x_avg: torch.Tensor
    step: torch.Tensor
    avg_error: torch.Tensor
    mu: torch.Tensor


def dual_averaging(t0: int = 10, gamma: float = 0.05, kappa: float = 0.75):
    """optimizers/dual_averaging.py:53-129 on ``(C,)`` CUDA tensors; returns (init, update, final)."""

    def _pack(st):
        return torch.stack([st.log_x, st.log_x_avg, st.step.to(st.log_x.dtype), st.avg_error, st.mu], dim=1).contiguous()

    def _unpack(da):
        return DualAveragingState(da[:, 0], da[:, 1], da[:, 2], da[:, 3], da[:, 4])

    def init(x_init: torch.Tensor) -> DualAveragingState:
        _require_f32(x_init)
        x = x_init.contiguous()
        da = torch.empty((x.shape[0], 5), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            N.check(N.lib().gb200_dual_averaging_init(N.ptr(da), N.ptr(x), x.shape[0], N.F32, N.stream_ptr()))
        return _unpack(da)

    def update(da_state: DualAveragingState, gradient: torch.Tensor) -> DualAveragingState:
        da = _pack(da_state)
        acc = (-gradient).contiguous()  # the native update takes (target - acceptance) with target = 0
        with torch.cuda.device(da.device):
            N.check(N.lib().gb200_dual_averaging_update(N.ptr(da), N.ptr(acc), 0.0, float(t0), float(gamma),
                                                        float(kappa), da.shape[0], N.F32, N.stream_ptr()))
        return _unpack(da)

    def final(da_state: DualAveragingState) -> torch.Tensor:
        return torch.exp(da_state.log_x_avg)

    return init, update, final


def _chain_keys(rng_key, C, device):
    if isinstance(rng_key, np.ndarray) and rng_key.shape == (2,):
        return grandom.split(rng_key[None], C, device=device)[0].contiguous()
    k = grandom._keys_tensor(rng_key, device)
    if k.shape == (2,):
        return grandom.split(k[None], C)[0].contiguous()
    if k.shape != (C, 2):
        raise ValueError(f"rng_key must be (2,) or (C, 2) = ({C}, 2); got {tuple(k.shape)}")
    return k


def _engine_for(algorithm, logdensity_fn, step_size, extra):
    extra = dict(extra)
    L = extra.pop("num_integration_steps")
    if algorithm is lmcmonge:
        return _Engine(N.LMCMONGE, logdensity_fn, step_size, L,
                       inverse_mass_matrix=extra.pop("inverse_mass_matrix", None),
                       alpha2=extra.pop("alpha2", 0.001), **extra)
    if algorithm in (rmhmc, lmc):
        tgt = _merge_target(logdensity_fn, extra.pop("metric_fn", None))
        return _Engine(N.RMHMC if algorithm is rmhmc else N.LMC, tgt, step_size, L, **extra)
    raise NotImplementedError("adaptation is available for geomjax_b200.rmhmc, lmc and lmcmonge")


def _require_f32(position):
    # the fused dual-averaging epilogue indexes its state in the chain state's dtype; the drivers keep float32
    # adaptation state, so float64 chains are refused here (and by gb200_step) instead of being mis-read
    if position.dtype != torch.float32:
        raise TypeError(f"adaptation runs on float32 chains; got {position.dtype}")


def _da_init(C, eps0, device):
    da = torch.empty((C, 5), dtype=torch.float32, device=device)
    x = torch.full((C,), float(eps0), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        N.check(N.lib().gb200_dual_averaging_init(N.ptr(da), N.ptr(x), C, N.F32, N.stream_ptr()))
    return da


def _run_opts(da, target_acceptance_rate):
    opts = N.RunOpts()
    opts.dual_averaging = N.ptr(da)
    opts.da_target, opts.da_t0, opts.da_gamma, opts.da_kappa = float(target_acceptance_rate), 10.0, 0.05, 0.75
    return opts


def _mean_accept(acc, process_group):
    s = torch.stack([acc.sum(), torch.tensor(float(acc.numel()), device=acc.device)])
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1:
        dist.all_reduce(s, group=process_group)  # the only collective of the warm-up: 2 floats
    return (s[0] / s[1]).reshape(1)


def step_size_adaptation(algorithm, logdensity_fn, initial_step_size: float = 1.0,
                         target_acceptance_rate: float = 0.80, progress_bar: bool = False,
                         lower_bound: float = 1e-3, pooled: bool = False, process_group=None,
                         **extra_parameters) -> AdaptationAlgorithm:
    """adaptation/step_size_adaptation.py:100-203."""
    del progress_bar  # host UI of the reference (fastprogress through host_callback); no-op here

    def run(rng_key, position: torch.Tensor, num_steps: int = 1000):
        _require_f32(position)
        C_ = position.shape[0]
        dev = position.device
        eng = _engine_for(algorithm, logdensity_fn, initial_step_size, extra_parameters)
        state = algorithm.init(position, eng.target)
        keys = _chain_keys(rng_key, C_, dev)
        all_keys = grandom.split(keys, num_steps).transpose(0, 1).contiguous()  # [t, c] = split(key_c, num_steps)[t]
        accept_hist = torch.empty((num_steps, C_), dtype=torch.float32, device=dev)
        fields = list(state)
        if not pooled:
            da = _da_init(C_, initial_step_size, dev)
            opts = _run_opts(da, target_acceptance_rate)
        else:
            da = _da_init(1, initial_step_size, dev)
            opts = N.RunOpts()
        for t in range(num_steps):
            ks = N.KeySource()
            ks.keys = N.ptr(all_keys[t])
            ks.num_transitions = 1
            opts.sample_accept = N.ptr(accept_hist[t])
            if pooled:
                eng.step_size = torch.exp(da[:, 0]).expand(C_).contiguous()
            fields, _ = eng.launch(fields, ks, want_info=False, opts=opts, out_state=fields)
            if pooled:
                acc = _mean_accept(accept_hist[t], process_group)
                with torch.cuda.device(dev):
                    N.check(N.lib().gb200_dual_averaging_update(N.ptr(da), N.ptr(acc), float(target_acceptance_rate),
                                                                10.0, 0.05, 0.75, 1, N.F32, N.stream_ptr()))
        step_size = torch.clamp(torch.exp(da[:, 1]), min=lower_bound)
        if pooled:
            step_size = step_size.expand(C_).contiguous()
        parameters = {"step_size": step_size, **extra_parameters}
        info = {"acceptance_rate": accept_hist, "dual_averaging": da}
        return AdaptationResults(eng.make_state(fields), parameters), info

    return AdaptationAlgorithm(run)


def _slow_windows(start: int, stop: int, first: int):
    """End indices (exclusive) of the doubling slow windows that tile [start, stop): a window doubles as long as
    the one after it would still fit three of itself; the last one absorbs the remainder."""
    ends, size = [], first
    while start < stop:
        if stop - start >= 3 * size:
            start, size = start + size, 2 * size
        else:
            start = stop
        ends.append(start)
    return ends


def build_schedule(num_steps: int, initial_buffer_size: int = 75, final_buffer_size: int = 50,
                   first_window_size: int = 25):
    """Stan's fast / slow / fast warm-up schedule as a list of ``(stage, is_middle_window_end)`` per transition:
    stage 0 = step size only, stage 1 = step size + mass matrix, the flag marks the last transition of a slow
    window.  Same list as adaptation/window_adaptation.py:360-450 (checked element by element against the oracle's
    restatement in tests/test_cabi_and_host.py)."""
    if num_steps < 20:  # too short for mass-matrix adaptation
        return [(0, False)] * num_steps
    if initial_buffer_size + first_window_size + final_buffer_size > num_steps:
        initial_buffer_size, final_buffer_size = int(0.15 * num_steps), int(0.1 * num_steps)
        first_window_size = num_steps - initial_buffer_size - final_buffer_size
    slow_stop = num_steps - final_buffer_size
    window_ends = set(_slow_windows(initial_buffer_size, slow_stop, first_window_size))
    return [(1, (t + 1) in window_ends) if initial_buffer_size <= t < slow_stop else (0, False)
            for t in range(num_steps)]


def window_adaptation(algorithm, logdensity_fn, is_mass_matrix_diagonal: bool = True,
                      initial_step_size: float = 1.0, target_acceptance_rate: float = 0.80,
                      progress_bar: bool = False, **extra_parameters) -> AdaptationAlgorithm:
    """adaptation/window_adaptation.py:245-357 for ``lmcmonge`` (diagonal mass matrix, per chain)."""
    del progress_bar
    if algorithm is not lmcmonge:
        raise NotImplementedError(
            "window_adaptation passes inverse_mass_matrix as the 5th kernel argument "
            "(window_adaptation.py:303-310), which only lmcmonge accepts among the Riemannian kernels; "
            "use step_size_adaptation for rmhmc / lmc")
    if not is_mass_matrix_diagonal:
        raise ValueError("The mass matrix has the wrong number of dimensions: expected 1, got 2.")  # lmcmonge/metrics.py:148-153

    def run(rng_key, position: torch.Tensor, num_steps: int = 1000):
        _require_f32(position)
        C_, D = position.shape
        dev = position.device
        inv_mass = torch.ones((C_, D), dtype=torch.float32, device=dev)  # mm_init mass_matrix.py:95-100
        eng = _engine_for(algorithm, logdensity_fn, initial_step_size,
                          dict(extra_parameters, inverse_mass_matrix=inv_mass))
        eng.inverse_mass_matrix = inv_mass  # keep the (C, D) tensor even while it is all ones
        state = algorithm.init(position, eng.target)
        keys = _chain_keys(rng_key, C_, dev)
        all_keys = grandom.split(keys, num_steps).transpose(0, 1).contiguous()
        da = _da_init(C_, initial_step_size, dev)
        opts = _run_opts(da, target_acceptance_rate)
        accept_hist = torch.empty((num_steps, C_), dtype=torch.float32, device=dev)
        mean = torch.zeros((C_, D), dtype=torch.float32, device=dev)   # Welford mass_matrix.py:200-215
        m2 = torch.zeros_like(mean)
        count = 0
        fields = list(state)
        for t, (stage, is_middle_window_end) in enumerate(build_schedule(num_steps)):
            ks = N.KeySource()
            ks.keys = N.ptr(all_keys[t])
            ks.num_transitions = 1
            opts.sample_accept = N.ptr(accept_hist[t])
            fields, _ = eng.launch(fields, ks, want_info=False, opts=opts, out_state=fields)
            if stage == 1:  # slow_update :148-168 (the dual-averaging update already ran in-kernel)
                count += 1
                delta = fields[0] - mean
                mean = mean + delta / count
                m2 = m2 + delta * (fields[0] - mean)
            if is_middle_window_end:  # slow_final :170-190, mm_final mass_matrix.py:129-150
                cov = m2 / (count - 1)
                inv_mass.copy_((count / (count + 5.0)) * cov + 1e-3 * (5.0 / (count + 5.0)))
                mean.zero_()
                m2.zero_()
                count = 0
                eps = torch.exp(da[:, 1]).contiguous()  # da_init(da_final(ss_state))
                with torch.cuda.device(dev):
                    N.check(N.lib().gb200_dual_averaging_init(N.ptr(da), N.ptr(eps), C_, N.F32, N.stream_ptr()))
        parameters = {"step_size": torch.exp(da[:, 1]), "inverse_mass_matrix": inv_mass, **extra_parameters}
        info = {"acceptance_rate": accept_hist, "dual_averaging": da}
        return AdaptationResults(eng.make_state(fields), parameters), info

    return AdaptationAlgorithm(run)
