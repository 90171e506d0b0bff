"""Dual-averaging step-size warm-up restated in NumPy (oracle; TEST INFRASTRUCTURE ONLY).

Follows optimizers/dual_averaging.py:53-129, adaptation/step_size.py:65-150 and
adaptation/step_size_adaptation.py:39-203 (one independent adaptation per chain = the
``jax.vmap`` of the reference's single-chain ``run``).
"""
from __future__ import annotations

import numpy as np


def da_init(x_init, dtype=np.float32):  # optimizers/dual_averaging.py:87-99
    x = np.asarray(x_init, dtype)
    return dict(log_x=np.log(x), log_x_avg=np.zeros_like(x), step=np.ones(x.shape, np.int32),
                avg_error=np.zeros_like(x), mu=np.log(dtype(10) * x))


def da_update(st, gradient, t0=10, gamma=0.05, kappa=0.75):  # :101-123
    dt = st["log_x"].dtype
    step = st["step"].astype(dt)
    reg = step + dt.type(t0)
    eta = step ** dt.type(-kappa)
    avg_error = (dt.type(1) - dt.type(1) / reg) * st["avg_error"] + np.asarray(gradient, dt) / reg
    log_x = st["mu"] - (np.sqrt(step) / dt.type(gamma)) * avg_error
    log_x_avg = eta * st["log_x"] + (dt.type(1) - eta) * st["log_x_avg"]  # uses the PREVIOUS log_x
    return dict(log_x=log_x.astype(dt), log_x_avg=log_x_avg.astype(dt), step=st["step"] + 1,
                avg_error=avg_error.astype(dt), mu=st["mu"])


def step_size_adaptation(step_fn, init_state, keys_fn, num_steps, num_chains,
                         initial_step_size=1.0, target_acceptance_rate=0.8, lower_bound=1e-3,
                         dtype=np.float32):
    """adaptation/step_size_adaptation.py:143-201; ``step_fn(keys, state, step_size[C])``."""
    st = da_init(np.full(num_chains, initial_step_size, dtype), dtype)
    eps = np.full(num_chains, initial_step_size, dtype)
    state = init_state
    for t in range(num_steps):
        state, info = step_fn(keys_fn(t), state, eps)
        st = da_update(st, dtype(target_acceptance_rate) - info.acceptance_rate)
        eps = np.exp(st["log_x"]).astype(dtype)
    final = np.maximum(np.exp(st["log_x_avg"]), dtype(lower_bound)).astype(dtype)
    return state, final, st


def build_schedule(num_steps, initial_buffer_size=75, final_buffer_size=50, first_window_size=25):
    """adaptation/window_adaptation.py:360-450 restated statement by statement (the specification the product's
    arithmetic formulation in geomjax_b200/adaptation.py is compared with)."""
    schedule = []
    if num_steps < 20:
        schedule += [(0, False)] * num_steps
    else:
        if initial_buffer_size + first_window_size + final_buffer_size > num_steps:
            initial_buffer_size = int(0.15 * num_steps)
            final_buffer_size = int(0.1 * num_steps)
            first_window_size = num_steps - initial_buffer_size - final_buffer_size
        schedule += [(0, False)] * (initial_buffer_size - 1)
        schedule.append((0, False))
        final_buffer_start = num_steps - final_buffer_size
        next_window_size = first_window_size
        next_window_start = initial_buffer_size
        while next_window_start < final_buffer_start:
            current_start, current_size = next_window_start, next_window_size
            if 3 * current_size <= final_buffer_start - current_start:
                next_window_size = 2 * current_size
            else:
                current_size = final_buffer_start - current_start
            next_window_start = current_start + current_size
            schedule += [(1, False)] * (next_window_start - 1 - current_start)
            schedule.append((1, True))
        schedule += [(0, False)] * (num_steps - 1 - final_buffer_start)
        schedule.append((0, False))
    return schedule
