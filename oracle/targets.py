"""Targets (log-density + metric) for the oracle, batched over a leading chain axis.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference takes arbitrary ``logdensity_fn`` / ``metric_fn`` callables and differentiates
them with JAX autodiff (``jax.value_and_grad``, ``jax.jacfwd``, ``jax.jvp(jax.grad)``:
rmhmc/integrators.py:113-116, lmcmc/metrics.py:164,188, util.py:120-121).  Without JAX the
oracle carries the analytic derivatives of each built-in target instead; they are verified
against central finite differences in ``tests/test_oracle_targets.py``.

Every target exposes, for ``q`` of shape (C, D):
  logp(q)      -> (C,)          logdensity_fn
  grad(q)      -> (C, D)        jax.grad(logdensity_fn)
  hvp(q, u)    -> (C, D)        util.hvp (Hessian-vector product)
  metric(q)    -> (C, D, D)     metric_fn  (dense)
  dmetric(q)   -> (C, D, D, D)  jax.jacfwd(metric_fn): out[c, j, l, i] = d G_jl / d q_i
Only Neal's funnel with its pull-back metric exists in the reference
(examples/funnel/main.py:28-54); Gaussian, banana and logistic regression are NEW targets
named by BASELINE.json's north_star and specified in SURVEY.md Appendix B.
"""
from __future__ import annotations

import numpy as np


class NealFunnel:
    """examples/funnel/main.py:28-54 (``neal_funnel``; D=2, sigma=3 by default)."""

    name = "funnel"

    def __init__(self, D=2, sigma=3.0, dtype=np.float32):
        self.D = int(D)
        self.sigma = float(sigma)
        self.dtype = np.dtype(dtype)

    # logp: examples/funnel/main.py:34-39, with jax.scipy.stats.norm.logpdf =
    # -(log(2*pi*scale^2) + (x-loc)^2/scale^2)/2
    def logp(self, q):
        dt = self.dtype
        q = np.asarray(q, dt)
        v = q[..., -1]
        x = q[..., :-1]
        two_pi = dt.type(2.0 * np.pi)
        s2 = dt.type(self.sigma) ** 2
        top = -(np.log(two_pi * s2) + v * v / s2) / dt.type(2)
        scale = np.exp(dt.type(0.5) * v)
        sc2 = scale * scale
        per = -(np.log(two_pi * sc2)[..., None] + x * x / sc2[..., None]) / dt.type(2)
        return (top + per.sum(-1)).astype(dt)

    def grad(self, q):
        dt = self.dtype
        q = np.asarray(q, dt)
        v = q[..., -1]
        x = q[..., :-1]
        e = np.exp(-v)
        g = np.empty_like(q)
        g[..., :-1] = -x * e[..., None]
        g[..., -1] = (-v / dt.type(self.sigma ** 2) - dt.type(0.5 * (self.D - 1))
                      + dt.type(0.5) * e * (x * x).sum(-1))
        return g

    def hvp(self, q, u):
        dt = self.dtype
        q = np.asarray(q, dt)
        u = np.asarray(u, dt)
        v = q[..., -1]
        x = q[..., :-1]
        e = np.exp(-v)
        ux, uv = u[..., :-1], u[..., -1]
        out = np.empty_like(q)
        out[..., :-1] = e[..., None] * (-ux + x * uv[..., None])
        hvv = -dt.type(1.0 / self.sigma ** 2) - dt.type(0.5) * e * (x * x).sum(-1)
        out[..., -1] = e * (x * ux).sum(-1) + hvv * uv
        return out

    def hessian(self, q):
        D = self.D
        eye = np.eye(D, dtype=self.dtype)
        return np.stack([self.hvp(q, np.broadcast_to(eye[i], q.shape)) for i in range(D)], -1)

    def dhessian(self, q):
        """out[c, j, l, i] = d^3 logp / dq_j dq_l dq_i (what jacfwd(hessian) yields)."""
        dt = self.dtype
        q = np.asarray(q, dt)
        C, D = q.shape
        x = q[:, :-1]
        e = np.exp(-q[:, -1])
        T3 = np.zeros((C, D, D, D), dt)
        k = np.arange(D - 1)
        V = D - 1
        # H_kk = -e, H_kv = x_k e, H_vv = -1/sigma^2 - e |x|^2 / 2
        T3[:, k, k, V] = e[:, None]            # d_v H_kk
        T3[:, k, V, k] = e[:, None]            # d_xk H_kv
        T3[:, V, k, k] = e[:, None]
        T3[:, k, V, V] = -(x * e[:, None])     # d_v H_kv
        T3[:, V, k, V] = -(x * e[:, None])
        T3[:, V, V, k] = -(x * e[:, None])     # d_xk H_vv
        T3[:, V, V, V] = dt.type(0.5) * e * (x * x).sum(-1)
        return T3

    # examples/funnel/main.py:41-50
    def inverse_jacobian(self, q):
        dt = self.dtype
        q = np.asarray(q, dt)
        C, D = q.shape
        A = np.zeros((C, D, D), dt)
        s = np.exp(dt.type(-0.5) * q[:, -1])
        idx = np.arange(D - 1)
        A[:, idx, idx] = s[:, None]
        A[:, :-1, -1] = dt.type(-0.5) * s[:, None] * q[:, :-1]
        A[:, -1, -1] = dt.type(1.0 / self.sigma)
        return A

    # examples/funnel/main.py:52-54
    def metric(self, q):
        A = self.inverse_jacobian(q)
        G = np.matmul(A.transpose(0, 2, 1), A)
        return (self.dtype.type(0.5) * (G + G.transpose(0, 2, 1))).astype(self.dtype)

    def dmetric(self, q):
        dt = self.dtype
        q = np.asarray(q, dt)
        C, D = q.shape
        x = q[:, :-1]
        e = np.exp(-q[:, -1])
        dG = np.zeros((C, D, D, D), dt)
        k = np.arange(D - 1)
        # d/dx_k
        dG[:, k, D - 1, k] = dt.type(-0.5) * e[:, None]
        dG[:, D - 1, k, k] = dt.type(-0.5) * e[:, None]
        dG[:, D - 1, D - 1, :-1] = dt.type(0.5) * e[:, None] * x
        # d/dv
        dG[:, k, k, D - 1] = -e[:, None]
        dG[:, :-1, D - 1, D - 1] = dt.type(0.5) * e[:, None] * x
        dG[:, D - 1, :-1, D - 1] = dt.type(0.5) * e[:, None] * x
        dG[:, D - 1, D - 1, D - 1] = dt.type(-0.25) * e * (x * x).sum(-1)
        return dG


class Gaussian:
    """NEW: l = -1/2 (q-mu)^T diag(prec) (q-mu); metric = diag(prec)."""

    name = "gaussian"

    def __init__(self, mean, precision_diag, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.mean = np.asarray(mean, self.dtype)
        self.prec = np.asarray(precision_diag, self.dtype)
        self.D = self.mean.shape[0]

    def logp(self, q):
        d = np.asarray(q, self.dtype) - self.mean
        return (self.dtype.type(-0.5) * (d * d * self.prec).sum(-1)).astype(self.dtype)

    def grad(self, q):
        return (-(np.asarray(q, self.dtype) - self.mean) * self.prec).astype(self.dtype)

    def hvp(self, q, u):
        return (-np.asarray(u, self.dtype) * self.prec).astype(self.dtype)

    def metric(self, q):
        C = q.shape[0]
        return np.broadcast_to(np.diag(self.prec), (C, self.D, self.D)).astype(self.dtype)

    def dmetric(self, q):
        return np.zeros((q.shape[0], self.D, self.D, self.D), self.dtype)


class Banana:
    """NEW, D=2: l = -x1^2/(2 s1^2) - (x2 - b (x1^2 - s1^2))^2 / 2;
    metric = identity."""

    name = "banana"

    def __init__(self, sigma1_sq=100.0, b=0.03, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.s1 = float(sigma1_sq)
        self.b = float(b)
        self.D = 2

    def _r(self, q):
        dt = self.dtype
        return q[..., 1] - dt.type(self.b) * (q[..., 0] ** 2 - dt.type(self.s1))

    def logp(self, q):
        dt = self.dtype
        q = np.asarray(q, dt)
        r = self._r(q)
        return (-q[..., 0] ** 2 / dt.type(2 * self.s1) - dt.type(0.5) * r * r).astype(dt)

    def grad(self, q):
        dt = self.dtype
        q = np.asarray(q, dt)
        r = self._r(q)
        g = np.empty_like(q)
        g[..., 0] = -q[..., 0] / dt.type(self.s1) + dt.type(2 * self.b) * q[..., 0] * r
        g[..., 1] = -r
        return g

    def hvp(self, q, u):
        dt = self.dtype
        q = np.asarray(q, dt)
        u = np.asarray(u, dt)
        r = self._r(q)
        x = q[..., 0]
        b = dt.type(self.b)
        h00 = -dt.type(1.0 / self.s1) + dt.type(2) * b * r - dt.type(4) * b * b * x * x
        h01 = dt.type(2) * b * x
        out = np.empty_like(q)
        out[..., 0] = h00 * u[..., 0] + h01 * u[..., 1]
        out[..., 1] = h01 * u[..., 0] - u[..., 1]
        return out

    def metric(self, q):
        return np.broadcast_to(np.eye(2, dtype=self.dtype), (q.shape[0], 2, 2)).copy()

    def dmetric(self, q):
        return np.zeros((q.shape[0], 2, 2, 2), self.dtype)


def make_logreg_data(N, D, seed=0, dtype=np.float32):
    """Frozen synthetic design: X[:,0]=1, X[:,1:]~N(0,1) column-standardised,
    theta*~N(0,1), y~Bernoulli(sigmoid(X theta*/sqrt(D)))."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, D))
    X[:, 1:] = (X[:, 1:] - X[:, 1:].mean(0)) / X[:, 1:].std(0)
    X[:, 0] = 1.0
    theta = rng.standard_normal(D)
    p = 1.0 / (1.0 + np.exp(-(X @ theta) / np.sqrt(D)))
    y = (rng.random(N) < p).astype(np.float64)
    return X.astype(dtype), y.astype(dtype)


class LogisticRegression:
    """NEW: Bayesian logistic regression, N(0, 1/alpha) prior, with the
    Fisher-information-plus-prior metric G = X^T diag(s(1-s)) X + alpha I."""

    name = "logreg"

    def __init__(self, X, y, prior_precision=0.01, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.X = np.asarray(X, self.dtype)
        self.y = np.asarray(y, self.dtype)
        self.alpha = float(prior_precision)
        self.N, self.D = self.X.shape

    def _eta(self, q):
        return np.asarray(q, self.dtype) @ self.X.T  # (C, N)

    def logp(self, q):
        dt = self.dtype
        q = np.asarray(q, dt)
        eta = self._eta(q)
        sp = np.logaddexp(dt.type(0), eta)
        return ((self.y * eta - sp).sum(-1) - dt.type(0.5 * self.alpha) * (q * q).sum(-1)).astype(dt)

    def _s(self, eta):
        return (self.dtype.type(1) / (self.dtype.type(1) + np.exp(-eta))).astype(self.dtype)

    def grad(self, q):
        dt = self.dtype
        q = np.asarray(q, dt)
        s = self._s(self._eta(q))
        return ((self.y - s) @ self.X - dt.type(self.alpha) * q).astype(dt)

    def hvp(self, q, u):
        dt = self.dtype
        q = np.asarray(q, dt)
        u = np.asarray(u, dt)
        s = self._s(self._eta(q))
        w = s * (dt.type(1) - s)
        return (-(w * (u @ self.X.T)) @ self.X - dt.type(self.alpha) * u).astype(dt)

    def metric(self, q):
        dt = self.dtype
        s = self._s(self._eta(q))
        w = s * (dt.type(1) - s)
        G = np.einsum("cn,ni,nj->cij", w, self.X, self.X, optimize=True)
        return (G + dt.type(self.alpha) * np.eye(self.D, dtype=dt)).astype(dt)

    def dmetric(self, q):
        dt = self.dtype
        s = self._s(self._eta(q))
        wp = s * (dt.type(1) - s) * (dt.type(1) - dt.type(2) * s)
        return np.einsum("cn,nj,nl,ni->cjli", wp, self.X, self.X, self.X, optimize=True).astype(dt)


def _logreg_contract_dmetric(self, q, Ginv, v):
    """tr(G^-1 d_i G) and v^T d_i G v for all i without materialising dG (C, D, D, D): the same einsum
    contractions as the dense path of oracle/samplers.py::_rmhmc_kinetic_grad, re-associated through
    d_i G = X^T diag(w' x_.i) X.  Needed for D = 100, N = 10,000 (dG would be 4 MB per chain and N D^3
    flops); equality with the dense path is tested at small sizes."""
    dt = self.dtype
    s = self._s(self._eta(q))
    wp = s * (dt.type(1) - s) * (dt.type(1) - dt.type(2) * s)          # (C, N)
    h = np.einsum("nj,cjl,nl->cn", self.X, Ginv.astype(dt), self.X, optimize=True)
    u = np.einsum("nj,cj->cn", self.X, v.astype(dt))
    tr = np.einsum("cn,ni->ci", wp * h, self.X)
    quad = np.einsum("cn,ni->ci", wp * u * u, self.X)
    return tr.astype(dt), quad.astype(dt)


LogisticRegression.contract_dmetric = _logreg_contract_dmetric


class WithMetric:
    """Wrap a target with a different ``metric_fn`` (e.g. ``lambda x: jnp.eye(2)`` of
    tests/test_samplers.py:25,37)."""

    def __init__(self, base, metric, dmetric=None):
        self.base = base
        self.D = base.D
        self.dtype = base.dtype
        self.logp, self.grad, self.hvp = base.logp, base.grad, base.hvp
        self._metric, self._dmetric = metric, dmetric

    def metric(self, q):
        return self._metric(q)

    def dmetric(self, q):
        if self._dmetric is None:
            return np.zeros((q.shape[0], self.D, self.D, self.D), self.dtype)
        return self._dmetric(q)


def identity_metric(target):
    D, dt = target.D, target.dtype
    return WithMetric(target, lambda q: np.broadcast_to(np.eye(D, dtype=dt), (q.shape[0], D, D)).copy())


def softabs_metric(target, alpha=1e6):
    """SoftAbs metric (Betancourt 2013): with H = -hessian(logp) = Q diag(lam) Q^T,
    G = Q diag(f(lam)) Q^T, f(lam) = lam coth(alpha lam), f(0) = 1/alpha.  The derivative is what autodiff
    through ``eigh`` computes (Daleckii-Krein): d_k G = Q (J o (Q^T d_k H Q)) Q^T with
    J_ij = (f_i - f_j)/(lam_i - lam_j), J_ii = f'(lam_i) = coth(a lam) - a lam / sinh^2(a lam).
    The eigen-decomposition runs in float64 and the results are rounded to the target dtype once
    (the spec the CUDA closed form for D = 2 is held to).  NEW metric, not in the reference; needs
    ``target.hessian`` and ``target.dhessian``."""
    dt = target.dtype

    def f_and_fp(lam):
        x = alpha * lam
        small = np.abs(x) < 1e-4
        big = np.abs(x) > 30.0
        xs = np.where(small | big, 1.0, x)
        coth = np.where(big, np.sign(x), 1.0 / np.tanh(xs))
        csch2 = np.where(big, 0.0, 1.0 / np.sinh(xs) ** 2)
        # series near 0: x coth x = 1 + x^2/3, d/dlam = (2/3) alpha x
        f = np.where(small, (1.0 + x * x / 3.0) / alpha, lam * coth)
        fp = np.where(small, (2.0 / 3.0) * x, coth - x * csch2)
        return f, fp

    def eig(q):
        H = -np.asarray(target.hessian(q), np.float64)
        H = 0.5 * (H + H.transpose(0, 2, 1))
        lam, Q = np.linalg.eigh(H)
        return lam, Q

    def metric(q):
        lam, Q = eig(q)
        f, _ = f_and_fp(lam)
        return np.einsum("cij,cj,ckj->cik", Q, f, Q).astype(dt)

    def dmetric(q):
        lam, Q = eig(q)
        f, fp = f_and_fp(lam)
        dl = lam[:, :, None] - lam[:, None, :]
        df = f[:, :, None] - f[:, None, :]
        # near-degenerate pairs use the derivative at the mean eigenvalue
        _, fpm = f_and_fp(0.5 * (lam[:, :, None] + lam[:, None, :]))
        close = np.abs(dl) <= 1e-9 * np.maximum(np.abs(lam[:, :, None]), np.abs(lam[:, None, :])) + 1e-300
        J = np.where(close, fpm, df / np.where(close, 1.0, dl))
        dH = -np.asarray(target.dhessian(q), np.float64)          # [c, j, l, i]
        M = np.einsum("cja,cjli,clb->cabi", Q, dH, Q)             # Q^T d_i H Q
        return np.einsum("cja,cab,cabi,clb->cjli", Q, J, M, Q).astype(dt)

    w = WithMetric(target, metric, dmetric)
    w.name = "softabs_" + getattr(target, "name", "target")
    w.softabs_alpha = float(alpha)
    return w
