"""The three static Riemannian transition kernels of geomjax, restated in NumPy and
batched over chains (oracle; TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``).

Follows, line by line:
  rmhmc     rmhmc/rmhmc.py:131-174,416-462; rmhmc/integrators.py:53-156; rmhmc/metrics.py:42-129
  lmc       lmcmc/lmc.py:135-180,451-499;   lmcmc/integrators.py:51-144;  lmcmc/metrics.py:42-221
  lmcmonge  lmcmonge/lmc.py:151-235,512-565; lmcmonge/integrators.py:52-230; lmcmonge/metrics.py:44-284
  shared    mcmc/trajectory.py:121-139 (static_integration); mcmc/proposal.py:43-121,168-185;
            mcmc/metrics.py:160-166 (hmc_energy); util.py:58-83,120-121
  hmc       mcmc/hmc.py + mcmc/integrators.py:47 velocity_verlet, only for the reference
            test's cross-sampler equivalence (tests/test_samplers.py:21-57).

Dense linear algebra throughout, exactly as the reference (Cholesky / LU / solve per call);
autodiff is replaced by the analytic derivatives in ``oracle/targets.py``.
Everything is evaluated in ``target.dtype`` (float32 by default = JAX without x64).
"""
from __future__ import annotations

from typing import NamedTuple

import numpy as np
import scipy.linalg as sla

from . import prng

# ----------------------------------------------------------------------------- helpers


def _mv(A, x):
    return np.einsum("cij,cj->ci", A, x)


def _dot(a, b):
    return np.einsum("ci,ci->c", a, b)


def _solve(A, b):
    try:
        return np.linalg.solve(A, b[..., None])[..., 0]
    except np.linalg.LinAlgError:  # singular / non-finite member of the batch: NaN for that chain (JAX semantics)
        out = np.full(b.shape, np.nan, b.dtype)
        for c in range(A.shape[0]):
            try:
                out[c] = np.linalg.solve(A[c], b[c])
            except np.linalg.LinAlgError:
                pass
        return out


def _chol(A):
    """Batched lower Cholesky; a non-positive-definite member yields NaN (as jnp does) instead of raising."""
    try:
        return np.linalg.cholesky(A)
    except np.linalg.LinAlgError:
        out = np.full(A.shape, np.nan, A.dtype)
        for c in range(A.shape[0]):
            try:
                out[c] = np.linalg.cholesky(A[c])
            except np.linalg.LinAlgError:
                pass
        return out


class Info(NamedTuple):
    """RMHMCInfo / LMCInfo (rmhmc/rmhmc.py:58-93, lmcmc/lmc.py:60-95, lmcmonge/lmc.py:63-98).
    ``momentum`` holds the initial draw: momentum for rmhmc, velocity for the LMC kernels.
    ``proposal`` is the flipped end-of-trajectory integrator state as a dict plus
    ``energy``/``weight``/``sum_log_p_accept`` (mcmc/proposal.py:22-40)."""
    momentum: np.ndarray
    acceptance_rate: np.ndarray
    is_accepted: np.ndarray
    is_divergent: np.ndarray
    energy: np.ndarray
    proposal: dict
    num_integration_steps: int
    extra: dict


def _mh(key_accept, H0, H1, divergence_threshold, mode, u_override=None):
    """mcmc/proposal.py:87-121 (proposal_from_energy_diff) + :168-185 (static_binomial_sampling)."""
    dt = H0.dtype
    with np.errstate(invalid="ignore", over="ignore"):
        delta = (H0 - H1).astype(dt)
        delta = np.where(np.isnan(delta), dt.type(-np.inf), delta)
        weight = delta
        sum_log_p_accept = np.minimum(delta, dt.type(0))
        p_accept = np.minimum(np.exp(weight), dt.type(1)).astype(dt)
        is_div = (-weight) > dt.type(divergence_threshold)
    if u_override is None:
        u = prng.uniform(key_accept, (), mode).astype(dt)
    else:
        u = np.asarray(u_override, dt)
    do_accept = u < p_accept
    return do_accept, p_accept, is_div, weight, sum_log_p_accept, u


def _draw_keys(keys, mode):
    """``key_a, key_b = jax.random.split(rng_key, 2)`` per chain."""
    ks = prng.split(np.asarray(keys, np.uint32), 2, mode)
    return ks[..., 0, :], ks[..., 1, :]


def _normal(key, D, dt, mode, z_override):
    if z_override is not None:
        return np.asarray(z_override, dt)
    return prng.normal(key, (D,), mode).astype(dt)


# ----------------------------------------------------------------------------- RMHMC


class RMHMCState(NamedTuple):  # rmhmc/rmhmc.py:30-41
    position: np.ndarray
    logdensity: np.ndarray
    logdensity_grad: np.ndarray


def rmhmc_init(position, target):  # rmhmc/rmhmc.py:96-98
    q = np.asarray(position, target.dtype)
    return RMHMCState(q, target.logp(q), target.grad(q))


def _rmhmc_kinetic(target, q, p):
    """rmhmc/metrics.py:60-74: -multivariate_normal.logpdf(p; 0, G(q)) through Cholesky."""
    dt = target.dtype
    G = target.metric(q)
    Lc = _chol(G)
    y = np.stack([sla.solve_triangular(Lc[c], p[c], lower=True, check_finite=False) for c in range(q.shape[0])])
    D = q.shape[1]
    logpdf = (dt.type(-0.5) * _dot(y, y) - dt.type(D / 2.0 * np.log(2 * np.pi))
              - np.log(np.diagonal(Lc, axis1=1, axis2=2)).sum(-1))
    return (-logpdf).astype(dt)


def _rmhmc_kinetic_grad(target, q, p):
    """jax.grad of the kinetic energy wrt (q, p) (rmhmc/integrators.py:114-116):
    dT/dp = G^-1 p;  dT/dq_i = 1/2 tr(G^-1 d_i G) - 1/2 (G^-1 p)^T d_i G (G^-1 p)."""
    dt = target.dtype
    G = target.metric(q)
    v = _solve(G, p)
    Ginv = np.linalg.inv(G)
    if getattr(target, "structured_dmetric", False):
        tr, quad = target.contract_dmetric(q, Ginv, v)   # same contractions, dG never materialised
    else:
        dG = target.dmetric(q)
        tr = np.einsum("cjl,cjli->ci", Ginv, dG)
        quad = np.einsum("cj,cjli,cl->ci", v, dG, v)
    return (dt.type(0.5) * tr - dt.type(0.5) * quad).astype(dt), v.astype(dt)


def implicit_midpoint_step(target, q0, p0, step_size, *, convergence_tol=1e-6,
                           divergence_tol=1e10, max_iters=100):
    """rmhmc/integrators.py:92-156 with solve_fixed_point_iteration :53-89 (vmapped
    while_loop = per-chain masked iteration)."""
    dt = target.dtype
    eps = dt.type(step_size)
    half = dt.type(0.5)

    def update(q, p, dUdq, init_q, init_p):  # _update :119-136
        dTdq, dHdp = _rmhmc_kinetic_grad(target, q, p)
        dHdq = dTdq - dUdq
        return (init_q + half * eps * dHdp).astype(dt), (init_p - half * eps * dHdq).astype(dt)

    def f(q, p):  # _step :139-142
        return update(q, p, target.grad(q), q0, p0)

    def norm(qa, pa, qb, pb):  # :57-60, max|.| over the ravelled (q, p) tuple
        with np.errstate(invalid="ignore"):
            return np.maximum(np.abs(qa - qb).max(-1), np.abs(pa - pb).max(-1)).astype(dt)

    with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
        q, p = f(q0, p0)
        nrm = norm(q, p, q0, p0)
        n = np.zeros(q0.shape[0], np.int32)
        while True:
            active = ((n < max_iters) & np.isfinite(nrm) & (nrm < dt.type(divergence_tol))
                      & (nrm > dt.type(convergence_tol)))
            if not active.any():
                break
            qn, pn = f(q, p)
            nn = norm(qn, pn, q, p)
            a = active[:, None]
            q, p = np.where(a, qn, q), np.where(a, pn, p)
            nrm = np.where(active, nn, nrm)
            n = n + active.astype(np.int32)
        success = np.isfinite(nrm) & (nrm <= dt.type(convergence_tol))
        # explicit update from the midpoint :147-148
        q, p = update(q, p, target.grad(q), q, p)
        v = _solve(target.metric(q), p).astype(dt)  # :150
    return q, p, v, target.logp(q), target.grad(q), dict(iters=n, norm=nrm, success=success)


def rmhmc_step(keys, state, target, step_size, num_integration_steps, *,
               divergence_threshold=1000, mode=prng.LEGACY, z=None, u=None,
               convergence_tol=1e-6, max_iters=100):
    """rmhmc/rmhmc.py:131-174 (kernel) + :416-438 (generate) for a batch of chains."""
    dt = target.dtype
    q0, l0, g0 = (np.asarray(a, dt) for a in state)
    C, D = q0.shape
    k_m, k_a = _draw_keys(keys, mode)
    zz = _normal(k_m, D, dt, mode, z)
    G = target.metric(q0)
    p0 = _mv(_chol(G), zz).astype(dt)      # rmhmc/metrics.py:45-58
    v0 = _solve(G, p0).astype(dt)                        # :120-127
    H0 = (-l0 + _rmhmc_kinetic(target, q0, p0)).astype(dt)  # mcmc/metrics.py:160-166
    q, p, v, l, g = q0, p0, v0, l0, g0
    iters = np.zeros(C, np.int64)
    for _ in range(num_integration_steps):               # mcmc/trajectory.py:137
        q, p, v, l, g, fp = implicit_midpoint_step(target, q, p, step_size,
                                                   convergence_tol=convergence_tol,
                                                   max_iters=max_iters)
        iters += fp["iters"]
    p, v = -p, -v                                        # flip_momentum :443-462
    with np.errstate(invalid="ignore"):
        H1 = (-l + _rmhmc_kinetic(target, q, p)).astype(dt)
    acc, p_acc, is_div, w, slpa, uu = _mh(k_a, H0, H1, divergence_threshold, mode, u)
    a = acc[:, None]
    new = RMHMCState(np.where(a, q, q0), np.where(acc, l, l0), np.where(a, g, g0))
    prop = dict(position=q, momentum=p, velocity=v, logdensity=l, logdensity_grad=g,
                energy=H1, weight=w, sum_log_p_accept=slpa)
    info = Info(p0, p_acc, acc, is_div, H1, prop, num_integration_steps,
                dict(z=zz, u=uu, H0=H0, velocity0=v0, fp_iters=iters))
    return new, info


# ----------------------------------------------------------------------------- LMC (Lan et al.)


class LMCState(NamedTuple):  # lmcmc/lmc.py:30-42, lmcmonge/lmc.py:32-44
    position: np.ndarray
    logdensity: np.ndarray
    logdensity_grad: np.ndarray
    volume_adjustment: np.ndarray


def lmc_init(position, target):  # lmcmc/lmc.py:98-101
    q = np.asarray(position, target.dtype)
    return LMCState(q, target.logp(q), target.grad(q), np.zeros(q.shape[0], target.dtype))


def _omega_tilde(target, q, v, step_size):
    """lmcmc/metrics.py:158-179 (dense branch)."""
    dt = target.dtype
    G = target.metric(q)
    dG = target.dmetric(q)
    p1 = np.einsum("ci,cjli->clj", v, dG)
    p2 = np.einsum("ci,cilj->clj", v, dG)
    p3 = np.einsum("ci,cijl->clj", v, dG)
    Om = dt.type(0.5) * (p1 + p2 - p3)
    return (G + dt.type(0.5) * dt.type(step_size) * Om).astype(dt)


def _grad_logdet_metric(target, q):
    """lmcmc/metrics.py:181-189: grad of slogdet(metric_fn) = tr(G^-1 d_i G)."""
    Ginv = np.linalg.inv(target.metric(q))
    return np.einsum("cjl,cjli->ci", Ginv, target.dmetric(q)).astype(target.dtype)


def _lu_logdet(A):
    lu, piv = sla.lu_factor(A, check_finite=False)  # NaN/inf are data (-> rejected), as in JAX
    d = np.diagonal(lu, axis1=-2, axis2=-1)
    with np.errstate(divide="ignore"):
        return (lu, piv), np.log(np.abs(d)).sum(-1).astype(A.dtype)


def _lmc_half_step(target, q, v, J, g, step_size):
    """lmcmc/integrators.py:61-91."""
    dt = target.dtype
    eps = dt.type(step_size)
    (lu, piv), ld = _lu_logdet(_omega_tilde(target, q, v, eps))
    J = (J - ld).astype(dt)
    dphi = -g + dt.type(0.5) * _grad_logdet_metric(target, q)
    v_temp = (_mv(target.metric(q), v) - dt.type(0.5) * eps * dphi).astype(dt)
    v_new = np.stack([sla.lu_solve((lu[c], piv[c]), v_temp[c], check_finite=False) for c in range(q.shape[0])]).astype(dt)
    _, ld2 = _lu_logdet(_omega_tilde(target, q, v_new, -eps))
    J = (J + ld2).astype(dt)
    return v_new, J


def lan_step(target, q, v, l, g, J, step_size):
    """lmcmc/integrators.py:93-142."""
    dt = target.dtype
    with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
        v, J = _lmc_half_step(target, q, v, J, g, step_size)
        q = (q + dt.type(step_size) * v).astype(dt)
        l, g = target.logp(q), target.grad(q)
        v, J = _lmc_half_step(target, q, v, J, g, step_size)
        p = _mv(target.metric(q), v).astype(dt)
    return q, p, v, l, g, J


def _lmc_kinetic(target, q, v):
    """lmcmc/metrics.py:93-112 (dense branch, symmetrised G, slogdet)."""
    dt = target.dtype
    G = target.metric(q)
    G = dt.type(0.5) * (G + G.transpose(0, 2, 1))
    with np.errstate(invalid="ignore", divide="ignore"):
        _, logdet = np.linalg.slogdet(G)
    return (dt.type(-0.5) * logdet.astype(dt) + dt.type(0.5) * _dot(_mv(G, v), v)).astype(dt)


def lmc_step(keys, state, target, step_size, num_integration_steps, *,
             divergence_threshold=1000, mode=prng.LEGACY, z=None, u=None):
    """lmcmc/lmc.py:135-180 (kernel) + :451-473 (generate)."""
    dt = target.dtype
    q0, l0, g0, J0 = (np.asarray(a, dt) for a in state)
    C, D = q0.shape
    k_v, k_a = _draw_keys(keys, mode)
    zz = _normal(k_v, D, dt, mode, z)
    # velocity_generator lmcmc/metrics.py:75-91: sigma = L^-T via solve_triangular(L, I, trans)
    G = target.metric(q0)
    G = dt.type(0.5) * (G + G.transpose(0, 2, 1))
    Lc = _chol(G)
    eye = np.eye(D, dtype=dt)
    sig = np.stack([sla.solve_triangular(Lc[c], eye, lower=True, trans=1, check_finite=False) for c in range(C)])
    v0 = _mv(sig.astype(dt), zz).astype(dt)
    p0 = _mv(target.metric(q0), v0).astype(dt)           # lmcmc/lmc.py:168
    H0 = (-l0 + _lmc_kinetic(target, q0, v0) - J0).astype(dt)  # lmc_energy :209-221
    q, p, v, l, g, J = q0, p0, v0, l0, g0, J0
    for _ in range(num_integration_steps):
        q, p, v, l, g, J = lan_step(target, q, v, l, g, J, step_size)
    p, v = -p, -v                                        # flip_velocity :478-499
    with np.errstate(invalid="ignore"):
        H1 = (-l + _lmc_kinetic(target, q, v) - J).astype(dt)
    acc, p_acc, is_div, w, slpa, uu = _mh(k_a, H0, H1, divergence_threshold, mode, u)
    a = acc[:, None]
    new = LMCState(np.where(a, q, q0), np.where(acc, l, l0), np.where(a, g, g0), np.where(acc, J, J0))
    prop = dict(position=q, momentum=p, velocity=v, logdensity=l, logdensity_grad=g,
                volume_adjustment=J, energy=H1, weight=w, sum_log_p_accept=slpa)
    info = Info(v0, p_acc, acc, is_div, H1, prop, num_integration_steps,
                dict(z=zz, u=uu, H0=H0, momentum0=p0))
    return new, info


# ----------------------------------------------------------------------------- LMC Monge


def lmcmonge_init(position, target):  # lmcmonge/lmc.py:101-109
    return lmc_init(position, target)


def _half_step_omega(dt, alpha2, v, J, dl, Hv, L, dl_ig, Hdl_ig, ig_Hdl_ig, eps, variant):
    """lmcmonge/integrators.py:158-194 (``omega``: as written, no alpha2 on the
    Christoffel correction -- SURVEY F8), ``omega_fixed`` (alpha2 restored), and
    :197-230 (``omegatilde``)."""
    a2 = dt.type(alpha2)
    half = dt.type(0.5)
    one = dt.type(1)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        det1 = one + half * eps * a2 * _dot(Hv, dl_ig)
        J = J - np.log(np.abs(det1))
        if variant in ("omega", "omega_fixed"):
            sL = np.sqrt(L)
            dphi = -dl + a2 * Hdl_ig / sL[:, None]
            dphi_ig = -dl_ig + a2 * ig_Hdl_ig / sL[:, None]
            v_temp = v - (half * eps * sL)[:, None] * (dphi_ig - (a2 * _dot(dphi, dl_ig))[:, None] * dl_ig)
            c = half * eps * _dot(v_temp, Hv) / det1
            if variant == "omega_fixed":
                c = a2 * c
            v_new = v_temp - c[:, None] * dl_ig
        elif variant == "omegatilde":
            v = v + (a2 * L * _dot(dl, v) + half * eps * np.sqrt(L))[:, None] * dl_ig \
                - half * a2 * eps * ig_Hdl_ig
            num = a2 * (_dot(dl, v) + half * eps * _dot(Hv, v))
            v_new = v - (num / det1)[:, None] * dl_ig
        else:
            raise ValueError(variant)
        J = J + np.log(np.abs(one - half * eps * a2 * _dot(Hdl_ig, v_new)))
    return v_new.astype(dt), J.astype(dt)


def _monge_det(dt, alpha2, inv_mass, g):  # normalizing_constant lmcmonge/metrics.py:193-197
    return (dt.type(1) + dt.type(alpha2) * _dot(inv_mass * g, g)).astype(dt)


def _monge_mvp(dt, alpha2, mass, v, g, L):  # metric_vector_product :243-255
    return (v * mass + (dt.type(alpha2) * L * _dot(v, g))[:, None] * g).astype(dt)


def _monge_kinetic(dt, alpha2, mass, v, dl, L):  # kinetic_energy :168-186
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        v_g = np.sqrt(mass) * v
        return (dt.type(-0.5) * (np.log(L) + np.log(mass).sum())
                + dt.type(0.5) * _dot(v_g, v_g)
                + dt.type(0.5) * L * dt.type(alpha2) * _dot(v, dl) ** dt.type(2.0)).astype(dt)


def monge_lan_step(target, alpha2, inv_mass, st, step_size, variant):
    """lmcmonge/integrators.py:63-153; ``st`` = dict of RiemannianIntegratorState fields."""
    dt = target.dtype
    eps = dt.type(step_size)
    mass = (dt.type(1) / inv_mass).astype(dt)
    with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
        v, J = _half_step_omega(dt, alpha2, st["velocity"], st["volume_adjustment"],
                                st["dl"], st["Hv"], st["L"], st["dl_ig"], st["Hdl_ig"],
                                st["ig_Hdl_ig"], eps, variant)
        q = (st["position"] + eps * v).astype(dt)
        l, g = target.logp(q), target.grad(q)
        L = _monge_det(dt, alpha2, inv_mass, g)
        sL = np.sqrt(L)[:, None]
        dl = (g / sL).astype(dt)
        Hv = (target.hvp(q, v) / sL).astype(dt)
        dl_ig = (inv_mass * dl).astype(dt)
        Hdl_ig = (target.hvp(q, dl_ig) / sL).astype(dt)
        ig_Hdl_ig = (inv_mass * Hdl_ig).astype(dt)
        v, J = _half_step_omega(dt, alpha2, v, J, dl, Hv, L, dl_ig, Hdl_ig, ig_Hdl_ig, eps, variant)
        Hv = (target.hvp(q, v) / sL).astype(dt)
        p = _monge_mvp(dt, alpha2, mass, v, g, L)
    return dict(position=q, momentum=p, velocity=v, logdensity=l, dl=dl, dl_ig=dl_ig,
                Hdl_ig=Hdl_ig, ig_Hdl_ig=ig_Hdl_ig, Hv=Hv, L=L, volume_adjustment=J)


def lmcmonge_step(keys, state, target, step_size, inverse_mass_matrix, num_integration_steps,
                  alpha2=0.001, *, divergence_threshold=1000, mode=prng.LEGACY,
                  half_step="omega", z=None, u=None):
    """lmcmonge/lmc.py:151-235 (kernel) + :512-534 (generate)."""
    dt = target.dtype
    q0, l0, g0, J0 = (np.asarray(a, dt) for a in state)
    C, D = q0.shape
    inv_mass = np.asarray(inverse_mass_matrix, dt)
    mass = (dt.type(1) / inv_mass).astype(dt)            # set_inverse_mass :44-60 ("fixed")
    with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
        L0 = _monge_det(dt, alpha2, inv_mass, g0)        # :177
        sL = np.sqrt(L0)[:, None]
        dl = (g0 / sL).astype(dt)                        # :180
        k_v, k_a = _draw_keys(keys, mode)                # :196
        dl_ig = (inv_mass * dl).astype(dt)               # :198
        Hdl_ig = (target.hvp(q0, dl_ig) / sL).astype(dt)  # :199
        ig_Hdl_ig = (inv_mass * Hdl_ig).astype(dt)       # :200
        zz = _normal(k_v, D, dt, mode, z)
        # velocity_generator lmcmonge/metrics.py:155-166: dense Cholesky of G^-1
        inv_metric = (np.einsum("i,ij->ij", inv_mass, np.eye(D, dtype=dt))[None]
                      - dt.type(alpha2) * np.einsum("ci,cj->cij", dl_ig, dl_ig)).astype(dt)
        Lc = _chol(inv_metric)
        v0 = _mv(Lc, zz).astype(dt)
        p0 = _monge_mvp(dt, alpha2, mass, v0, g0, L0)    # :202-204 (un-normalised grad, as written)
        Hv = (target.hvp(q0, v0) / sL).astype(dt)        # :206-208
        st0 = dict(position=q0, momentum=p0, velocity=v0, logdensity=l0, dl=dl, dl_ig=dl_ig,
                   Hdl_ig=Hdl_ig, ig_Hdl_ig=ig_Hdl_ig, Hv=Hv, L=L0, volume_adjustment=J0)
        # lmcmonge_energy :267-284
        H0 = (-l0 + _monge_kinetic(dt, alpha2, mass, v0, dl, L0) - J0).astype(dt)
        st = st0
        for _ in range(num_integration_steps):
            st = monge_lan_step(target, alpha2, inv_mass, st, step_size, half_step)
        st = dict(st, momentum=-st["momentum"], velocity=-st["velocity"])  # flip_velocity :539-565
        H1 = (-st["logdensity"] + _monge_kinetic(dt, alpha2, mass, st["velocity"], st["dl"], st["L"])
              - st["volume_adjustment"]).astype(dt)
    acc, p_acc, is_div, w, slpa, uu = _mh(k_a, H0, H1, divergence_threshold, mode, u)
    a = acc[:, None]
    with np.errstate(invalid="ignore", over="ignore"):
        # :226-233: logdensity_grad = logdensity_grad_norm * sqrt(determinant_metric) of the
        # *sampled* integrator state (on rejection this is dl0*sqrt(L0), not bitwise g0)
        g_prop = (st["dl"] * np.sqrt(st["L"])[:, None]).astype(dt)
        g_init = (dl * np.sqrt(L0)[:, None]).astype(dt)
    new = LMCState(np.where(a, st["position"], q0), np.where(acc, st["logdensity"], l0),
                   np.where(a, g_prop, g_init), np.where(acc, st["volume_adjustment"], J0))
    prop = dict(st, energy=H1, weight=w, sum_log_p_accept=slpa, logdensity_grad=g_prop)
    info = Info(v0, p_acc, acc, is_div, H1, prop, num_integration_steps,
                dict(z=zz, u=uu, H0=H0, momentum0=p0, L0=L0))
    return new, info


# ----------------------------------------------------------------------------- Euclidean HMC
# Only for the reference test's equivalence check (tests/test_samplers.py:27-32,56-57).


def hmc_step(keys, state, target, step_size, inverse_mass_matrix, num_integration_steps, *,
             divergence_threshold=1000, mode=prng.LEGACY):
    """mcmc/hmc.py kernel with mcmc/metrics.py gaussian_euclidean (diagonal) and
    mcmc/integrators.py:47 velocity_verlet."""
    dt = target.dtype
    q0, l0, g0 = (np.asarray(a, dt) for a in state[:3])
    D = q0.shape[1]
    im = np.asarray(inverse_mass_matrix, dt)
    k_m, k_a = _draw_keys(keys, mode)
    zz = prng.normal(k_m, (D,), mode).astype(dt)
    p0 = (zz * (dt.type(1) / np.sqrt(im))).astype(dt)
    eps = dt.type(step_size)
    half = dt.type(0.5)

    def kin(p):
        return (half * _dot(p, im * p)).astype(dt)

    q, p, l, g = q0, p0, l0, g0
    for _ in range(num_integration_steps):
        p = p + half * eps * g
        q = q + eps * (im * p)
        l, g = target.logp(q), target.grad(q)
        p = p + half * eps * g
    p = -p
    H0 = -l0 + kin(p0)
    H1 = -l + kin(p)
    acc, p_acc, is_div, w, slpa, uu = _mh(k_a, H0, H1, divergence_threshold, mode)
    a = acc[:, None]
    return RMHMCState(np.where(a, q, q0), np.where(acc, l, l0), np.where(a, g, g0)), \
        Info(p0, p_acc, acc, is_div, H1, dict(position=q, momentum=p), num_integration_steps, dict(u=uu))


# ----------------------------------------------------------------------------- driver loop


def chain_keys(root_key, num_samples, t, num_chains, chain_idx=None, mode=prng.LEGACY):
    """examples/funnel/main.py:18,22: ``split(split(root, T)[t], C)[c]``."""
    k_t = prng.split_index(root_key, num_samples, np.asarray(t), mode)
    idx = np.arange(num_chains) if chain_idx is None else np.asarray(chain_idx)
    return prng.split_index(k_t, num_chains, idx, mode)


def inference_loop(root_key, step_fn, init_state, num_samples, mode=prng.LEGACY):
    """examples/funnel/main.py:7-25 (inference_loop_multiple_chains); returns stacked positions
    (T, C, D) and the acceptance rates (T, C)."""
    state = init_state
    C = state[0].shape[0]
    pos, acc = [], []
    for t in range(num_samples):
        keys = chain_keys(root_key, num_samples, t, C, mode=mode)
        state, info = step_fn(keys, state)
        pos.append(state[0].copy())
        acc.append(info.acceptance_rate.copy())
    return state, np.stack(pos), np.stack(acc)
