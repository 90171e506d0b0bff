"""rmhmc on the logistic-regression target with every evaluation of the implicit-midpoint map run for ALL
chains in lock-step on the warp-specialised tcgen05 pipeline (``gb200_logreg_midpoint_map`` /
``gb200_logreg_state_eval``): both D^2 N products of an evaluation are chain-batched 3xTF32 GEMMs.

Semantics are the reference's under ``jax.vmap``: rmhmc/rmhmc.py:131-174 (kernel), :416-462 (proposal,
flip), rmhmc/integrators.py:53-89 (``solve_fixed_point_iteration``: a vmapped ``while_loop`` runs until every
chain is done, converged chains keep their iterate -- masked commits), :92-156 (implicit midpoint),
mcmc/proposal.py:87-121,168-185 (energy difference, accept).  The host loop below only sequences launches and
does the O(C D) book-keeping (masks, norms, selects, compaction of the still-active chains) with torch
element-wise ops; everything O(N D) and above is in the CUDA library.  One device->host read per fixed-point
iteration (which chains are still active).
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _native as N
from . import random as grandom
from .base import SamplingAlgorithm
from .samplers import Proposal, RMHMCInfo, RMHMCIntegratorState, RMHMCState, _merge_target

__all__ = ["rmhmc_lockstep"]


class _Pipe:
    """Workspace + the two native entry points for one (target, C)."""

    def __init__(self, target, C_, device):
        self.t, self.C, self.dev = target, C_, device
        self.desc = target.c_struct()
        self.ws_bytes = int(N.lib().gb200_logreg_midpoint_map_workspace(C.byref(self.desc), C_))
        self.ws = torch.empty(self.ws_bytes // 4 + 64, dtype=torch.float32, device=device)
        self.wsp = C.c_void_p(self.ws.data_ptr() + (-self.ws.data_ptr()) % 256)

    def _new(self, *shape):
        return torch.empty(shape, dtype=torch.float32, device=self.dev)

    def state_eval(self, q, p=None, z=None):
        D = self.t.D
        assert q.shape[0] == self.C
        out = dict(logdensity=self._new(self.C), logdensity_grad=self._new(self.C, D), velocity=self._new(self.C, D),
                   logdet=self._new(self.C), momentum=self._new(self.C, D) if z is not None else p)
        with torch.cuda.device(self.dev):
            N.check(N.lib().gb200_logreg_state_eval(C.byref(self.desc), N.ptr(q), N.ptr(p), N.ptr(z),
                                                    N.ptr(out["momentum"]) if z is not None else None,
                                                    N.ptr(out["logdensity"]), N.ptr(out["logdensity_grad"]),
                                                    N.ptr(out["velocity"]), N.ptr(out["logdet"]), self.wsp, self.ws_bytes,
                                                    self.C, N.F32, N.stream_ptr()))
        return out

    def midpoint_map(self, q, p, qi, pi, he):
        """One evaluation for the ``q.shape[0] <= C`` chains passed in (a compacted sub-batch reuses the workspace)."""
        D, Ca = self.t.D, q.shape[0]
        qn, pn = self._new(Ca, D), self._new(Ca, D)
        lp, g, v, ld = self._new(Ca), self._new(Ca, D), self._new(Ca, D), self._new(Ca)
        with torch.cuda.device(self.dev):
            N.check(N.lib().gb200_logreg_midpoint_map(C.byref(self.desc), N.ptr(q), N.ptr(p), N.ptr(qi), N.ptr(pi), float(he),
                                                      N.ptr(qn), N.ptr(pn), N.ptr(lp), N.ptr(g), N.ptr(v), N.ptr(ld), None,
                                                      self.wsp, self.ws_bytes, Ca, N.F32, N.stream_ptr()))
        return qn, pn


def _norm(qa, pa, qb, pb):
    """max |x_{n+1} - x_n| over the ravelled (q, p) tuple per chain (rmhmc/integrators.py:57-60); inf if not finite."""
    d = torch.maximum((qa - qb).abs().amax(dim=1), (pa - pb).abs().amax(dim=1))
    return torch.where(torch.isnan(d), torch.full_like(d, float("inf")), d)


class rmhmc_lockstep:
    """Same constructor as ``geomjax_b200.rmhmc`` (geomjax/rmhmc/rmhmc.py:286-311) for the logistic-regression
    target; ``step(rng_keys[C, 2], state) -> (RMHMCState, RMHMCInfo)``."""

    @staticmethod
    def init(position, logdensity_fn):
        from .samplers import rmhmc
        return rmhmc.init(position, logdensity_fn)

    def __new__(cls, logdensity_fn, step_size, metric_fn, num_integration_steps, *, divergence_threshold: int = 1000,
                convergence_tol: float = 1e-6, divergence_tol: float = 1e10, max_iters: int = 100):
        target = _merge_target(logdensity_fn, metric_fn)
        if target.kind != N.TARGET_LOGREG or target.metric != N.METRIC_TARGET:
            raise NotImplementedError("rmhmc_lockstep is built for the logistic-regression target with its Fisher metric")
        L, eps = int(num_integration_steps), float(step_size)
        pipes = {}

        def step(rng_key, state):
            q0 = state.position.to(torch.float32).contiguous()
            C_, D = q0.shape
            dev = q0.device
            pipe = pipes.get((C_, dev))
            if pipe is None:
                pipe = pipes[(C_, dev)] = _Pipe(target, C_, dev)
            keys = grandom._keys_tensor(rng_key, dev)
            if keys.shape != (C_, 2):
                raise ValueError(f"rng_key must have shape (C, 2) = ({C_}, 2)")
            ks = grandom.split(keys, 2)                       # rmhmc/rmhmc.py:158
            z = grandom.normal(ks[:, 0].contiguous(), D)       # util.py:81-82
            u = grandom.uniform(ks[:, 1].contiguous(), 1).reshape(C_)  # mcmc/proposal.py:177-178
            s0 = pipe.state_eval(q0, z=z)                      # p = chol(G) z, v = G^-1 p
            p0 = s0["momentum"]
            half_log_2pi_D = 0.5 * D * math.log(2.0 * math.pi)
            H0 = -state.logdensity + 0.5 * (p0 * s0["velocity"]).sum(1) + 0.5 * s0["logdet"] + half_log_2pi_D
            q, p = q0, p0
            he = 0.5 * eps
            iters = torch.zeros(C_, dtype=torch.int32, device=dev)
            inf = float("inf")
            for _ in range(L):                                 # mcmc/trajectory.py:137
                qi, pi = q, p
                q, p = pipe.midpoint_map(qi, pi, qi, pi, he)   # x1 = f(x0)
                nrm = _norm(q, p, qi, pi)
                n = torch.zeros(C_, dtype=torch.int32, device=dev)
                while True:
                    active = (n < max_iters) & (nrm < inf) & (nrm < divergence_tol) & (nrm > convergence_tol)
                    idx = active.nonzero().squeeze(1)          # vmapped while_loop: until every chain is done
                    if idx.numel() == 0:
                        break
                    # The fixed-point count is heavy-tailed (a float32 iterate stalling just above tol runs to
                    # max_iters): evaluate the map only for the chains that are still iterating.  Chains are
                    # independent, so compaction changes no chain's numbers -- only how many ride along.
                    full = idx.numel() == C_
                    qa, pa = (q, p) if full else (q[idx], p[idx])
                    qia, pia = (qi, pi) if full else (qi[idx], pi[idx])
                    qc, pc = pipe.midpoint_map(qa.contiguous(), pa.contiguous(), qia.contiguous(), pia.contiguous(), he)
                    nc = _norm(qc, pc, qa, pa)
                    if full:
                        q, p, nrm = qc, pc, nc
                    else:
                        q = q.index_copy(0, idx, qc)
                        p = p.index_copy(0, idx, pc)
                        nrm = nrm.index_copy(0, idx, nc)
                    n = n + active.to(torch.int32)
                iters += n
                q, p = pipe.midpoint_map(q, p, q, p, he)       # explicit update from the midpoint :147-148
            s1 = pipe.state_eval(q, p=p)
            H1 = -s1["logdensity"] + 0.5 * (p * s1["velocity"]).sum(1) + 0.5 * s1["logdet"] + half_log_2pi_D
            delta = H0 - H1
            delta = torch.where(torch.isnan(delta), torch.full_like(delta, -inf), delta)
            p_accept = torch.clamp(torch.exp(delta), max=1.0)
            accept = u < p_accept
            a = accept[:, None]
            new = RMHMCState(torch.where(a, q, q0), torch.where(accept, s1["logdensity"], state.logdensity),
                             torch.where(a, s1["logdensity_grad"], state.logdensity_grad))
            prop = Proposal(RMHMCIntegratorState(q, -p, -s1["velocity"], s1["logdensity"], s1["logdensity_grad"]), H1, delta,
                            torch.clamp(delta, max=0.0))
            info = RMHMCInfo(p0, p_accept, accept, (-delta) > divergence_threshold, H1, prop, L)
            step.last_fp_iters = iters
            return new, info

        return SamplingAlgorithm(lambda position: cls.init(position, target), step)
