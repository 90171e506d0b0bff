"""Integrator sentinels.  The reference passes integrator *functions*
(rmhmc/integrators.py:92 ``implicit_midpoint``, lmcmc/integrators.py:51 ``lan_integrator``,
lmcmonge/integrators.py:52 ``lan_integrator`` with ``half_step_omega`` / ``half_step_omegatilde``
:158,:197); here the integrators are compiled into the fused kernels, so the ``integrator=``
argument selects one of these descriptors instead (anything else raises)."""
from __future__ import annotations

from . import _native as N


class Integrator:
    def __init__(self, name, samplers, **kwargs):
        self.name, self.samplers, self.kwargs = name, samplers, kwargs

    def __call__(self, **solver_kwargs):
        """``implicit_midpoint(convergence_tol=..., max_iters=...)``-style overrides
        (rmhmc/integrators.py:53-61 solver kwargs)."""
        allowed = {"convergence_tol": "fp_convergence_tol", "divergence_tol": "fp_divergence_tol",
                   "max_iters": "fp_max_iters", "half_step": "half_step"}
        kw = dict(self.kwargs)
        for k, v in solver_kwargs.items():
            if k not in allowed:
                raise TypeError(f"unknown integrator option {k!r}")
            kw[allowed[k]] = v
        return Integrator(self.name, self.samplers, **kw)

    def __repr__(self):
        return f"<integrator {self.name} {self.kwargs}>"


implicit_midpoint = Integrator("implicit_midpoint", (N.RMHMC,))
lan_integrator = Integrator("lan_integrator", (N.LMC,))
lan_integrator_monge = Integrator("lan_integrator_monge", (N.LMCMONGE,), half_step="omega")
half_step_omega = lan_integrator_monge
half_step_omegatilde = Integrator("lan_integrator_monge", (N.LMCMONGE,), half_step="omegatilde")
half_step_omega_fixed = Integrator("lan_integrator_monge", (N.LMCMONGE,), half_step="omega_fixed")


def resolve(integrator, sampler_id) -> Integrator:
    if not isinstance(integrator, Integrator):
        raise NotImplementedError(
            "integrators are compiled into the fused CUDA kernels; pass one of "
            "geomjax_b200.integrators.{implicit_midpoint, lan_integrator, lan_integrator_monge, "
            "half_step_omegatilde, half_step_omega_fixed}")
    if sampler_id not in integrator.samplers:
        raise ValueError(f"integrator {integrator.name} cannot be used with this sampler")
    return integrator
