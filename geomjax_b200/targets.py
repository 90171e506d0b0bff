"""Target descriptors: the built-in replacements for the reference's ``logdensity_fn`` /
``metric_fn`` callables (the engine evaluates targets inside the fused CUDA kernels, so an
arbitrary Python callable cannot be accepted -- there is no CPU / tracing fallback).

Only Neal's funnel exists in the reference tree (examples/funnel/main.py:28-54); the others
are the NEW built-ins named by BASELINE.json's north_star.
"""
from __future__ import annotations

from . import _native as N


class TargetDescriptor:
    """Opaque handle passed wherever the reference takes ``logdensity_fn`` (and ``metric_fn``)."""

    def __init__(self, kind: int, D: int, params=(), metric: int = N.METRIC_TARGET, N_rows: int = 0,
                 X=None, y=None, vec0=None, vec1=None, name: str = "target"):
        self.kind, self.D, self.metric, self.name = int(kind), int(D), int(metric), name
        self.params = tuple(float(p) for p in params)
        self.N = int(N_rows)
        self._keep = (X, y, vec0, vec1)  # keep the device tensors alive

    def c_struct(self) -> N.TargetDesc:
        d = N.TargetDesc()
        d.kind, d.metric, d.D, d.N = self.kind, self.metric, self.D, self.N
        for i, p in enumerate(self.params):
            d.params[i] = p
        X, y, v0, v1 = self._keep
        d.X, d.y, d.vec0, d.vec1 = N.ptr(X), N.ptr(y), N.ptr(v0), N.ptr(v1)
        return d

    def evaluate_metric(self, position):
        """``jax.vmap(metric_fn)(position)``: (C, D) -> (C, D, D).  Built for the logistic-regression
        target, where it runs as one tcgen05 (3xTF32) GEMM over the chain dimension."""
        import ctypes as C
        import torch
        if self.kind != N.TARGET_LOGREG:
            raise NotImplementedError("evaluate_metric() is built for the logistic-regression target")
        q = position.to(torch.float32).contiguous()
        if q.ndim != 2 or q.shape[1] != self.D:
            raise ValueError(f"position must have shape (C, {self.D})")
        d = self.c_struct()
        ws_bytes = N.lib().gb200_logreg_fisher_metric_workspace(C.byref(d), q.shape[0])
        ws = torch.empty(max(int(ws_bytes), 4) // 4, dtype=torch.float32, device=q.device)
        G = torch.empty((q.shape[0], self.D, self.D), dtype=torch.float32, device=q.device)
        with torch.cuda.device(q.device):
            N.check(N.lib().gb200_logreg_fisher_metric(C.byref(d), N.ptr(q), N.ptr(G), N.ptr(ws), ws_bytes,
                                                       q.shape[0], N.F32, N.stream_ptr()))
        return G

    def quadratic_forms(self, matrices):
        """``h[c, n] = x_n^T A_c x_n`` for per-chain symmetric matrices ``A`` (C, D, D) -> (C, N): with ``A = G^-1``
        the ``h_n`` of rmhmc's ``dT/dq`` on the logistic-regression target, as one tcgen05 (3xTF32) GEMM over the
        chain dimension (the twin of ``evaluate_metric``)."""
        import ctypes as C
        import torch
        if self.kind != N.TARGET_LOGREG:
            raise NotImplementedError("quadratic_forms() is built for the logistic-regression target")
        A = matrices.to(torch.float32).contiguous()
        if A.ndim != 3 or A.shape[1:] != (self.D, self.D):
            raise ValueError(f"matrices must have shape (C, {self.D}, {self.D})")
        d = self.c_struct()
        ws_bytes = N.lib().gb200_logreg_quadform_workspace(C.byref(d), A.shape[0])
        ws = torch.empty(max(int(ws_bytes), 16) // 4, dtype=torch.float32, device=A.device)
        h = torch.empty((A.shape[0], self.N), dtype=torch.float32, device=A.device)
        with torch.cuda.device(A.device):
            N.check(N.lib().gb200_logreg_quadform(C.byref(d), N.ptr(A), N.ptr(h), self.N, N.ptr(ws), ws_bytes,
                                                  A.shape[0], N.F32, N.stream_ptr()))
        return h

    def with_metric(self, metric: str) -> "TargetDescriptor":
        """``metric='identity'`` == ``metric_fn=lambda x: jnp.eye(D)`` (tests/test_samplers.py:25)."""
        m = {"target": N.METRIC_TARGET, "identity": N.METRIC_IDENTITY}[metric]
        t = TargetDescriptor(self.kind, self.D, self.params, m, self.N, *self._keep, name=self.name)
        return t

    @property
    def base_params(self):
        """Parameters of the log-density (params[7] is reserved for the metric: softabs alpha)."""
        return tuple(self.params[:7])

    # the reference passes the same object's bound methods as logdensity_fn / metric_fn
    @property
    def logp(self):
        return self

    @property
    def fisher_metric_fn(self):
        return self

    metric_fn = fisher_metric_fn

    def __repr__(self):
        return f"TargetDescriptor({self.name}, D={self.D})"


def neal_funnel(D: int = 2, mean: float = 0.0, sigma: float = 3.0) -> TargetDescriptor:
    """examples/funnel/main.py:28-54 (``class neal_funnel``); ``mean`` is unused there as well."""
    if D < 2:
        raise ValueError("neal_funnel needs D >= 2")
    return TargetDescriptor(N.TARGET_FUNNEL, D, (sigma,), name="neal_funnel")


def softabs(target: TargetDescriptor, alpha: float = 1e6) -> TargetDescriptor:
    """NEW metric (the metric BASELINE.json configs[0] names for rmhmc): the SoftAbs
    map of the Hessian, ``G = Q diag(lam coth(alpha lam)) Q^T`` with ``-hessian(logp) = Q diag(lam) Q^T``.
    Use as ``metric_fn``: ``rmhmc(funnel, eps, softabs(funnel), L)``.  Built for Neal's funnel, D = 2."""
    if target.kind != N.TARGET_FUNNEL or target.D != 2:
        raise NotImplementedError("softabs() is built for neal_funnel(D=2)")
    params = list(target.base_params) + [0.0] * (7 - len(target.base_params)) + [float(alpha)]
    return TargetDescriptor(target.kind, target.D, params, N.METRIC_SOFTABS, target.N, *target._keep,
                            name="softabs(" + target.name + ")")


def gaussian(mean, precision_diag) -> TargetDescriptor:
    """NEW built-in: l = -1/2 sum_j prec_j (q_j - mean_j)^2 with the constant
    metric diag(prec) (a (D,)-shaped ``metric_fn`` in the reference's terms).  CUDA tensors (D,)."""
    import torch
    mean = mean.to(torch.float32).contiguous()
    prec = precision_diag.to(torch.float32).contiguous()
    if mean.ndim != 1 or prec.shape != mean.shape:
        raise ValueError("gaussian needs mean (D,) and precision_diag (D,)")
    return TargetDescriptor(N.TARGET_GAUSSIAN, mean.shape[0], (), vec0=mean, vec1=prec, name="gaussian")


def banana(sigma1_sq: float = 100.0, b: float = 0.03) -> TargetDescriptor:
    """NEW built-in, D = 2: l = -x1^2/(2 s1) - (x2 - b (x1^2 - s1))^2 / 2; identity metric."""
    return TargetDescriptor(N.TARGET_BANANA, 2, (sigma1_sq, b), metric=N.METRIC_IDENTITY, name="banana")


def logistic_regression(X, y, prior_precision: float = 0.01) -> TargetDescriptor:
    """NEW built-in (BASELINE.json north_star): Bayesian logistic regression
    l(theta) = sum_n [y_n eta_n - softplus(eta_n)] - alpha/2 |theta|^2, eta = X theta, with the
    Fisher-information + prior metric G = X^T diag(s(1-s)) X + alpha I.  ``X``: (N, D) CUDA tensor,
    ``y``: (N,) in {0, 1}.  The kernels read the transposed, 16-byte-padded copy made here."""
    import torch
    if X.ndim != 2 or y.ndim != 1 or y.shape[0] != X.shape[0]:
        raise ValueError("logistic_regression needs X (N, D) and y (N,)")
    Nrows, D = X.shape
    ldx = (Nrows + 3) // 4 * 4
    Xt = torch.zeros((D, ldx), dtype=torch.float32, device=X.device)
    Xt[:, :Nrows] = X.t().to(torch.float32)
    X = X.to(torch.float32).contiguous()
    y = y.to(torch.float32).contiguous()
    return TargetDescriptor(N.TARGET_LOGREG, D, (prior_precision, float(ldx)), N_rows=Nrows, X=X, y=y, vec0=Xt,
                            name="logistic_regression")


def as_target(obj) -> TargetDescriptor:
    if isinstance(obj, TargetDescriptor):
        return obj
    raise NotImplementedError(
        "geomjax_b200 evaluates log-densities and metrics inside fused CUDA kernels: pass a "
        "TargetDescriptor (e.g. geomjax_b200.targets.neal_funnel(D)) where the reference takes "
        f"logdensity_fn / metric_fn; got {type(obj).__name__}. There is no CPU fallback.")
