"""ChEES adaptation for the Riemannian kernels with the reference signature
(geomjax/adaptation/chees_adaptation_riemanian.py:56-466): step size by dual averaging on the harmonic mean of the
chains' acceptance probabilities, trajectory length by an optimiser step on the ChEES criterion, both from statistics
pooled over ALL chains.

What runs where: every transition is one launch of the fused lmc / rmhmc kernel with that transition's (jittered)
number of integration steps; the cross-chain statistics are two reduction kernels (csrc/chees.cu) whose outputs are
plain sums, all-reduced over ranks when ``torch.distributed`` is initialised (chains sharded across GPUs: the pooled
adaptation is the design's only collective besides the diagnostics); the scalar recurrences (dual averaging, Adam on
the log trajectory length, moving averages) are a few host flops per transition.

``optim`` follows the optax ``GradientTransformation`` protocol (``init(params)``, ``update(grads, state, params) ->
(updates, state)``); optax is not installed here, ``adam`` below restates ``optax.adam``.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Callable, NamedTuple, Optional

import numpy as np
import torch

from . import _native as N
from . import random as grandom
from .base import AdaptationAlgorithm, AdaptationResults
from .samplers import _Engine, _merge_target, lmc, rmhmc

__all__ = ["chees_adaptation", "adam", "halton_sequence", "ChEESAdaptationState", "OPTIMAL_TARGET_ACCEPTANCE_RATE"]

OPTIMAL_TARGET_ACCEPTANCE_RATE = 0.651  # chees_adaptation_riemanian.py:24


class ChEESAdaptationState(NamedTuple):  # chees_adaptation_riemanian.py:27-53
    step_size: float
    log_step_size_moving_average: float
    trajectory_length: float
    log_trajectory_length_moving_average: float
    da_state: tuple
    optim_state: object
    random_generator_arg: object
    step: int


class _Adam(NamedTuple):
    init: Callable
    update: Callable


def adam(learning_rate: float, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8) -> _Adam:
    """``optax.adam(learning_rate)`` for a scalar parameter: scale_by_adam (bias-corrected moments, eps outside the
    square root) followed by ``scale(-learning_rate)``."""

    def init(params):
        return (0, 0.0, 0.0)

    def update(grad, state, params=None):
        count, mu, nu = state
        count += 1
        mu = b1 * mu + (1.0 - b1) * grad
        nu = b2 * nu + (1.0 - b2) * grad * grad
        mu_hat = mu / (1.0 - b1 ** count)
        nu_hat = nu / (1.0 - b2 ** count)
        return -learning_rate * mu_hat / (math.sqrt(nu_hat) + eps), (count, mu, nu)

    return _Adam(init, update)


def halton_sequence(i: int, max_bits: int = 10) -> float:
    """chees_adaptation_riemanian.py:469-471: base-2 radical inverse of i + 1 over `max_bits` bits."""
    return float(sum((((i + 1) >> b) & 1) * 0.5 / (1 << b) for b in range(int(max_bits))))


def _da_init(x):  # optimizers/dual_averaging.py:87-99 (scalar, float64 on the host)
    return (math.log(x), 0.0, 1, 0.0, math.log(10.0 * x))


def _da_update(st, gradient, t0=10, gamma=0.05, kappa=0.75):  # :101-123
    log_x, log_x_avg, step, avg_error, mu = st
    reg_step = step + t0
    eta_t = step ** (-kappa)
    avg_error = (1.0 - 1.0 / reg_step) * avg_error + gradient / reg_step
    new_log_x = mu - (math.sqrt(step) / gamma) * avg_error
    new_log_x_avg = eta_t * log_x + (1.0 - eta_t) * log_x_avg  # the PREVIOUS log_x, as the reference writes it (:120)
    return (new_log_x, new_log_x_avg, step + 1, avg_error, mu)


def _allreduce(t, process_group):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=process_group)
    return t


def pooled_statistics(proposed_positions, proposed_velocities, initial_positions, acceptance_probabilities, is_divergent,
                      process_group=None):
    """The chain-axis reductions of ``compute_parameters`` (:145-187) on the GPU: returns (harmonic mean of the
    acceptance probabilities over the non-divergent chains, acceptance-weighted mean of the per-chain ChEES gradient
    factors).  Two kernels, two small all-reduces."""
    q = proposed_positions.contiguous()
    C_, D = q.shape
    dev = q.device
    div = is_divergent.to(torch.uint8).contiguous()
    acc = acceptance_probabilities.contiguous()
    out1 = torch.empty(4 * D + 2, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().gb200_chees_moments(N.ptr(q), N.ptr(initial_positions.contiguous()), N.ptr(acc), N.ptr(div), C_, D,
                                            N.ptr(out1), N.stream_ptr()))
    out1 = _allreduce(out1, process_group)
    means = torch.cat([out1[:D] / out1[D:2 * D], out1[2 * D:3 * D] / out1[3 * D:4 * D]]).contiguous()
    out2 = torch.empty(2, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().gb200_chees_gradient(N.ptr(q), N.ptr(proposed_velocities.contiguous()),
                                             N.ptr(initial_positions.contiguous()), N.ptr(acc), N.ptr(div), N.ptr(means), C_, D,
                                             N.ptr(out2), N.stream_ptr()))
    out2 = _allreduce(out2, process_group)
    h = out1[4 * D:].cpu().numpy()
    g2 = out2.cpu().numpy()
    with np.errstate(all="ignore"):
        harmonic_mean = float(1.0 / (h[0] / h[1]))
        weighted = float(g2[0] / g2[1])
    return harmonic_mean, weighted


def base(jitter_generator: Callable, next_random_arg_fn: Callable, optim, target_acceptance_rate: float,
         decay_rate: float, process_group=None):
    """chees_adaptation_riemanian.py:56-273: (init, update) of the adaptation state."""

    def init(random_generator_arg, step_size: float) -> ChEESAdaptationState:
        return ChEESAdaptationState(step_size, 0.0, step_size, 0.0, _da_init(step_size), optim.init(step_size),
                                    random_generator_arg, 1)

    def update(st: ChEESAdaptationState, proposed_positions, proposed_velocities, initial_positions,
               acceptance_probabilities, is_divergent) -> ChEESAdaptationState:
        harmonic_mean, weighted = pooled_statistics(proposed_positions, proposed_velocities, initial_positions,
                                                    acceptance_probabilities, is_divergent, process_group)
        da_ = _da_update(st.da_state, target_acceptance_rate - harmonic_mean)
        try:
            step_size_ = math.exp(da_[0])
        except OverflowError:
            step_size_ = math.inf
        if math.isfinite(step_size_):
            new_step_size, new_da, new_log_step_size = step_size_, da_, da_[0]
        else:
            new_step_size, new_da, new_log_step_size = st.step_size, st.da_state, st.da_state[0]
        w = st.step ** (-decay_rate)
        new_log_step_size_ma = (1.0 - w) * st.log_step_size_moving_average + w * new_log_step_size
        trajectory_gradient = jitter_generator(st.random_generator_arg) * st.trajectory_length * weighted
        log_tl = math.log(st.trajectory_length)
        updates, optim_state_ = optim.update(trajectory_gradient, st.optim_state, log_tl)
        log_tl_ = log_tl + updates
        if math.isfinite(log_tl_):
            new_log_tl, new_optim = log_tl_, optim_state_
        else:
            new_log_tl, new_optim = log_tl, st.optim_state
        new_log_tl_ma = (1.0 - w) * st.log_trajectory_length_moving_average + w * new_log_tl
        return ChEESAdaptationState(new_step_size, new_log_step_size_ma, math.exp(new_log_tl_ma), new_log_tl_ma, new_da,
                                    new_optim, next_random_arg_fn(st.random_generator_arg), st.step + 1)

    return init, update


def chees_adaptation(logprob_fn, metric_fn, num_chains: int, *, jitter_generator: Optional[Callable] = None,
                     jitter_amount: float = 1.0, target_acceptance_rate: float = OPTIMAL_TARGET_ACCEPTANCE_RATE,
                     decay_rate: float = 0.5, dynamics: str = "lmc", process_group=None, chain_offset: int = 0,
                     total_chains: Optional[int] = None) -> AdaptationAlgorithm:
    """chees_adaptation_riemanian.py:276-466.  ``num_chains`` is the LOCAL chain count; with chains sharded over ranks
    pass ``chain_offset`` / ``total_chains`` (keys come from the global chain index, statistics are all-reduced)."""
    if dynamics not in ("lmc", "rmhmc"):
        raise ValueError("dynamics must be 'lmc' or 'rmhmc'")
    algorithm = lmc if dynamics == "lmc" else rmhmc
    sampler_id = N.LMC if dynamics == "lmc" else N.RMHMC
    total_chains = num_chains + chain_offset if total_chains is None else total_chains

    def run(rng_key, positions: torch.Tensor, step_size: float, optim, num_steps: int = 1000, *,
            max_sampling_steps: int = 1000):
        assert positions.shape[0] == num_chains, "initial `positions` leading dimension must be equal to the `num_chains`"
        if positions.dtype != torch.float32:
            raise TypeError("chees_adaptation runs on float32 chains")
        dev = positions.device
        key = grandom._keys_tensor(rng_key, dev).reshape(2)
        key_init, key_step = grandom.split(key[None], 2)[0]
        if jitter_generator is not None:
            jitter_gn = lambda k: jitter_generator(k) * jitter_amount + (1.0 - jitter_amount)
            next_random_arg_fn = lambda k: grandom.split(k[None], 2)[0, 1].contiguous()
            init_random_arg = key_init
        else:
            bits_ = math.ceil(math.log2(num_steps + max_sampling_steps))
            jitter_gn = lambda i: halton_sequence(i, bits_) * jitter_amount + (1.0 - jitter_amount)
            next_random_arg_fn = lambda i: i + 1
            init_random_arg = 0

        def integration_steps_fn(random_generator_arg, trajectory_length_adjusted):
            return int(math.ceil(float(jitter_gn(random_generator_arg)) * trajectory_length_adjusted))

        init, update = base(jitter_gn, next_random_arg_fn, optim, target_acceptance_rate, decay_rate, process_group)
        target = _merge_target(logprob_fn, metric_fn)
        state = algorithm.init(positions, target)
        adaptation_state = init(init_random_arg, float(step_size))
        keys_step = grandom.split(key_step[None], num_steps)[0]  # (num_steps, 2)
        history = {"step_size": [], "trajectory_length": [], "num_integration_steps": [], "acceptance_rate": []}
        for t in range(num_steps):
            L_t = integration_steps_fn(adaptation_state.random_generator_arg,
                                       adaptation_state.trajectory_length / adaptation_state.step_size)
            eng = _Engine(sampler_id, target, adaptation_state.step_size, L_t)
            keys = grandom.split(keys_step[t][None], total_chains)[0][chain_offset:chain_offset + num_chains].contiguous()
            new_state, info = eng.step(keys, state)
            adaptation_state = update(adaptation_state, info.proposal.state.position, info.proposal.state.velocity,
                                      state.position, info.acceptance_rate, info.is_divergent)
            state = new_state
            history["step_size"].append(adaptation_state.step_size)
            history["trajectory_length"].append(adaptation_state.trajectory_length)
            history["num_integration_steps"].append(L_t)
            history["acceptance_rate"].append(info.acceptance_rate)
        trajectory_length_adjusted = math.exp(adaptation_state.log_trajectory_length_moving_average
                                              - adaptation_state.log_step_size_moving_average)
        parameters = {"step_size": math.exp(adaptation_state.log_step_size_moving_average), "metric_fn": metric_fn,
                      "next_random_arg_fn": next_random_arg_fn,
                      "integration_steps_fn": lambda arg: integration_steps_fn(
                          int(arg.reshape(-1)[0]) if isinstance(arg, torch.Tensor) and jitter_generator is None else arg,
                          trajectory_length_adjusted)}
        history["acceptance_rate"] = torch.stack(history["acceptance_rate"]) if history["acceptance_rate"] else None
        history["adaptation_state"] = adaptation_state
        # the dynamic kernels carry the generator argument in their state (rmhmc/rmhmc.py:44-55)
        from .samplers import DynamicLMCState, DynamicRMHMCState, _random_arg
        arg = _random_arg(adaptation_state.random_generator_arg if jitter_generator is None else
                          adaptation_state.random_generator_arg, num_chains, dev)
        last = DynamicLMCState(*state, arg) if dynamics == "lmc" else DynamicRMHMCState(*state, arg)
        return AdaptationResults(last, parameters), history

    return AdaptationAlgorithm(run)
