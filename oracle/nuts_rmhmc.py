"""``geomjax.nutsrmhmc`` restated for ONE chain (oracle regression only; no CUDA NUTS exists).

TEST INFRASTRUCTURE ONLY.  Purpose: the reference's single author-produced number,
``tests/test_samplers.py:10-19`` (``nutsrmhmc`` one step from ``zeros(2)`` with
``jr.key(42)``, step_size 1e-2, funnel metric -> ``[-0.73879963, 1.2370402]``), runs through
the same threefry PRNG, momentum draw, implicit-midpoint integrator, funnel target and
``hmc_energy`` as the static ``rmhmc`` kernel, so reproducing it pins those oracle pieces.

Follows rmhmc/nuts.py:114-161,253-348; mcmc/trajectory.py:149-317 (dynamic progressive
integration), :517-668 (multiplicative expansion); mcmc/termination.py:30-119
(numpyro-style iterative U-turn); mcmc/proposal.py:198-254 (progressive sampling);
rmhmc/metrics.py:76-118 (``is_turning``, criterion "euc").
"""
from __future__ import annotations

import numpy as np

from . import prng
from .samplers import _rmhmc_kinetic, _mv, _solve, implicit_midpoint_step


def _logaddexp(a, b):
    with np.errstate(invalid="ignore"):
        return np.logaddexp(a, b).astype(np.result_type(a, b))


def _is_turning(v_left, v_right, v_sum):  # criterion "euc": rmhmc/metrics.py:104-108
    return bool((np.dot(v_left, v_sum) <= 0) | (np.dot(v_right, v_sum) <= 0))


def _ckpt_idxs(n):  # mcmc/termination.py:72-87
    idx_max = bin(n >> 1).count("1")
    num_subtrees = 0
    m = n
    while m & 1:
        m >>= 1
        num_subtrees += 1
    return idx_max - num_subtrees + 1, idx_max


def nuts_rmhmc_step(key, position, target, step_size, *, max_num_doublings=10,
                    divergence_threshold=1000, mode=prng.LEGACY):
    dt = target.dtype
    D = target.D
    q0 = np.asarray(position, dt).reshape(1, D)

    def energy(s):  # hmc_energy, mcmc/metrics.py:160-166
        return (-s["l"] + _rmhmc_kinetic(target, s["q"], s["p"]))[0].astype(dt)

    def integrate(s, eps):
        q, p, v, l, g, _ = implicit_midpoint_step(target, s["q"], s["p"], eps)
        return dict(q=q, p=p, v=v, l=l, g=g)

    k_m, k_i = prng.split(key, 2, mode)                       # rmhmc/nuts.py:144
    z = prng.normal(k_m, (D,), mode).astype(dt)[None]
    G = target.metric(q0)
    p0 = _mv(np.linalg.cholesky(G), z).astype(dt)
    s0 = dict(q=q0, p=p0, v=_solve(G, p0).astype(dt), l=target.logp(q0), g=target.grad(q0))
    H0 = energy(s0)
    minus_inf = dt.type(-np.inf)

    # termination checkpoints: mcmc/termination.py:33-43
    r_ck = np.zeros((max_num_doublings, D), dt)
    rs_ck = np.zeros_like(r_ck)
    v_ck = np.zeros_like(r_ck)
    vs_ck = np.zeros_like(r_ck)

    def new_proposal(s):  # mcmc/proposal.py:87-121
        with np.errstate(invalid="ignore"):
            delta = (H0 - energy(s)).astype(dt)
        if np.isnan(delta):
            delta = minus_inf
        return dict(state=s, energy=energy(s), weight=delta, slpa=min(delta, dt.type(0)))

    def subtree(k_traj, start, direction, max_steps):
        """mcmc/trajectory.py:182-315."""
        nonlocal r_ck, rs_ck, v_ck, vs_ck
        step = 0
        prop = new_proposal(start)
        traj = dict(left=start, right=start, psum=start["p"][0], vsum=start["v"][0], n=0)
        diverging = turned = False
        carry = k_traj
        while step < max_steps and not turned and not diverging:
            carry, k_p = prng.split(carry, 2, mode)
            new = integrate(traj["right"], dt.type(direction) * dt.type(step_size))
            npz = new_proposal(new)
            diverging = bool(-npz["weight"] > divergence_threshold)
            if step == 0:
                traj = dict(left=new, right=new, psum=new["p"][0], vsum=new["v"][0], n=1)
                prop = npz
            else:
                traj = dict(left=traj["left"], right=new, psum=traj["psum"] + new["p"][0],
                            vsum=traj["vsum"] + new["v"][0], n=traj["n"] + 1)
                # progressive_uniform_sampling mcmc/proposal.py:198-222
                with np.errstate(over="ignore"):
                    p_acc = dt.type(1) / (dt.type(1) + np.exp(-(npz["weight"] - prop["weight"])))
                acc = prng.uniform(k_p, (), mode) < p_acc
                w = _logaddexp(prop["weight"], npz["weight"])
                sl = _logaddexp(prop["slpa"], npz["slpa"])
                src = npz if acc else prop
                prop = dict(state=src["state"], energy=src["energy"], weight=w, slpa=sl)
            # update_criterion_state mcmc/termination.py:45-70
            idx_min, idx_max = _ckpt_idxs(step)
            if step % 2 == 0:
                r_ck[idx_max] = new["p"][0]
                rs_ck[idx_max] = traj["psum"]
                v_ck[idx_max] = new["v"][0]
                vs_ck[idx_max] = traj["vsum"]
            # _is_iterative_turning :89-117
            turned = False
            i = idx_max
            while i >= idx_min and not turned:
                sub_vsum = traj["vsum"] - vs_ck[i] + v_ck[i]
                turned = _is_turning(v_ck[i], new["v"][0], sub_vsum)
                i -= 1
            step += 1
        if direction < 0:
            traj = dict(left=traj["right"], right=traj["left"], psum=traj["psum"],
                        vsum=traj["vsum"], n=traj["n"])
        return prop, traj, diverging, turned

    # mcmc/trajectory.py:558-664
    prop = dict(state=s0, energy=H0, weight=dt.type(0), slpa=minus_inf)
    traj = dict(left=s0, right=s0, psum=s0["p"][0], vsum=s0["v"][0], n=0)
    step = 0
    diverging = turning = False
    carry = k_i
    while step < max_num_doublings and not diverging and not turning:
        ks = prng.split(carry, 4, mode)
        carry, k_dir, k_traj, k_prop = ks[0], ks[1], ks[2], ks[3]
        direction = 1 if prng.uniform(k_dir, (), mode) < np.float32(0.5) else -1
        start = traj["right"] if direction > 0 else traj["left"]
        nprop, ntraj, diverging, turn_sub = subtree(k_traj, start, direction, 2 ** step)
        if diverging or turn_sub:
            prop = dict(prop, slpa=_logaddexp(prop["slpa"], nprop["slpa"]))
        else:  # progressive_biased_sampling mcmc/proposal.py:225-254
            with np.errstate(over="ignore"):
                p_acc = min(np.exp(nprop["weight"] - prop["weight"]), dt.type(1))
            acc = prng.uniform(k_prop, (), mode) < p_acc
            w = _logaddexp(prop["weight"], nprop["weight"])
            sl = _logaddexp(prop["slpa"], nprop["slpa"])
            src = nprop if acc else prop
            prop = dict(state=src["state"], energy=src["energy"], weight=w, slpa=sl)
        left, right = (traj, ntraj) if direction > 0 else (ntraj, traj)
        traj = dict(left=left["left"], right=right["right"], psum=left["psum"] + right["psum"],
                    vsum=left["vsum"] + right["vsum"], n=left["n"] + right["n"])
        turn_full = _is_turning(traj["left"]["v"][0], traj["right"]["v"][0], traj["vsum"])
        turning = turn_sub or turn_full
        step += 1
    info = dict(momentum=p0[0], num_doublings=step, num_states=traj["n"],
                is_divergent=diverging, is_turning=turning)
    return prop["state"]["q"][0], info
