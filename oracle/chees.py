"""ChEES adaptation restated in NumPy (oracle; TEST INFRASTRUCTURE ONLY).

Follows adaptation/chees_adaptation_riemanian.py:56-466 (``base.compute_parameters``, ``chees_adaptation.run`` with the
default Halton jitter) with ``optax.adam`` restated (scale_by_adam + scale(-lr)); the transitions are the oracle's
``lmc_step`` / ``rmhmc_step`` with that transition's number of integration steps for every chain."""
from __future__ import annotations

import numpy as np

from . import adaptation as A
from . import prng as P
from . import samplers as S


def halton(i, max_bits=10):  # :469-471
    masks = 2 ** np.arange(int(max_bits), dtype=np.int64)
    return float((((i + 1) // masks) % 2 * (0.5 / masks)).sum())


def adam_init():
    return dict(count=0, mu=0.0, nu=0.0)


def adam_update(grad, st, lr, b1=0.9, b2=0.999, eps=1e-8):
    c = st["count"] + 1
    mu = b1 * st["mu"] + (1 - b1) * grad
    nu = b2 * st["nu"] + (1 - b2) * grad * grad
    upd = -lr * (mu / (1 - b1 ** c)) / (np.sqrt(nu / (1 - b2 ** c)) + eps)
    return upd, dict(count=c, mu=mu, nu=nu)


def compute_parameters(st, prop_q, prop_v, init_q, acc, div, jitter, lr, target, decay_rate):
    """:102-219 for one transition; `st` is a dict of the ChEESAdaptationState fields."""
    ok = ~div
    with np.errstate(all="ignore"):
        harmonic = 1.0 / np.mean(1.0 / acc[ok].astype(np.float64))
        da_ = A.da_update(st["da"], np.float64(target - harmonic))
        step_size_ = float(np.exp(da_["log_x"]))
        if np.isfinite(step_size_):
            new_eps, new_da, new_log_eps = step_size_, da_, float(da_["log_x"])
        else:
            new_eps, new_da, new_log_eps = st["step_size"], st["da"], float(st["da"]["log_x"])
        w = st["step"] ** (-decay_rate)
        log_eps_ma = (1 - w) * st["log_step_size_ma"] + w * new_log_eps
        pc = prop_q - np.nanmean(prop_q, axis=0)
        ic = init_q - np.nanmean(init_q, axis=0)
        g = jitter * st["trajectory_length"] * ((pc * pc).sum(1) - (ic * ic).sum(1)) * (pc * prop_v).sum(1)
        grad = float((acc[ok].astype(np.float64) * g[ok]).sum() / acc[ok].astype(np.float64).sum())
    log_tl = np.log(st["trajectory_length"])
    upd, opt_ = adam_update(grad, st["optim"], lr)
    log_tl_ = log_tl + upd
    if np.isfinite(log_tl_):
        new_log_tl, new_opt = log_tl_, opt_
    else:
        new_log_tl, new_opt = log_tl, st["optim"]
    log_tl_ma = (1 - w) * st["log_trajectory_length_ma"] + w * new_log_tl
    return dict(step_size=new_eps, log_step_size_ma=log_eps_ma, trajectory_length=float(np.exp(log_tl_ma)),
                log_trajectory_length_ma=log_tl_ma, da=new_da, optim=new_opt, arg=st["arg"] + 1, step=st["step"] + 1)


def run(rng_key, positions, target, step_size, lr, num_steps, *, dynamics="lmc", max_sampling_steps=1000,
        target_acceptance_rate=0.651, decay_rate=0.5, jitter_amount=1.0):
    """:353-464 (Halton jitter); returns the per-transition history (step size, trajectory length, L)."""
    C = positions.shape[0]
    key_init, key_step = P.split(np.asarray(rng_key, np.uint32), 2)
    bits = int(np.ceil(np.log2(num_steps + max_sampling_steps)))
    jit = lambda i: halton(i, bits) * jitter_amount + (1.0 - jitter_amount)
    st = dict(step_size=float(step_size), log_step_size_ma=0.0, trajectory_length=float(step_size),
              log_trajectory_length_ma=0.0, da=A.da_init(np.float64(step_size), np.float64), optim=adam_init(), arg=0, step=1)
    state = (S.lmc_init if dynamics == "lmc" else S.rmhmc_init)(positions, target)
    keys_step = P.split(key_step, num_steps)
    hist = dict(step_size=[], trajectory_length=[], num_integration_steps=[])
    for t in range(num_steps):
        L = int(np.ceil(jit(st["arg"]) * st["trajectory_length"] / st["step_size"]))
        keys = P.split(keys_step[t], C)
        with np.errstate(all="ignore"):
            if dynamics == "lmc":
                new, info = S.lmc_step(keys, state, target, st["step_size"], L)
            else:
                new, info = S.rmhmc_step(keys, state, target, st["step_size"], L)
        st = compute_parameters(st, info.proposal["position"], info.proposal["velocity"], state[0], info.acceptance_rate,
                                info.is_divergent, jit(st["arg"]), lr, target_acceptance_rate, decay_rate)
        state = new
        hist["step_size"].append(st["step_size"])
        hist["trajectory_length"].append(st["trajectory_length"])
        hist["num_integration_steps"].append(L)
    return state, st, hist
