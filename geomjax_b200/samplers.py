"""rmhmc / lmc / lmcmonge with the reference's ``build_kernel / init / step`` API, batched over a
leading chain axis and executed by the fused CUDA kernels behind the C ABI.

Mirrors (names, argument order, defaults, NamedTuple fields):
  geomjax/rmhmc/rmhmc.py:30-41,58-93,96-98,110-176,247-311
  geomjax/lmcmc/lmc.py:30-42,60-95,98-101,114-182,255-346
  geomjax/lmcmonge/lmc.py:32-44,63-98,101-109,130-237,312-405
  geomjax/mcmc/proposal.py:22-40 (Proposal)
Differences, all forced by "no tracing compiler, no CPU fallback":
  * ``logdensity_fn`` / ``metric_fn`` are ``TargetDescriptor`` objects (geomjax_b200.targets);
  * every array carries a leading chain axis: ``init(position[C, D])``,
    ``step(rng_keys[C, 2] uint32, state)`` == ``jax.vmap(kernel)(keys, states)``
    (examples/funnel/main.py:13,19);
  * ``integrator`` must be one of the sentinels in ``geomjax_b200.integrators``.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, NamedTuple, Optional

import numpy as np
import torch

from . import _native as N
from . import integrators
from . import random as grandom
from .base import SamplingAlgorithm
from .targets import as_target

__all__ = ["rmhmc", "lmc", "lmcmonge", "dynamic_rmhmc", "dynamic_lmc", "dynamic_lmcmonge", "RMHMCState", "RMHMCInfo",
           "LMCState", "LMCInfo", "DynamicRMHMCState", "DynamicLMCState", "Proposal", "run_fused"]


class RMHMCState(NamedTuple):
    position: torch.Tensor
    logdensity: torch.Tensor
    logdensity_grad: torch.Tensor


class LMCState(NamedTuple):
    position: torch.Tensor
    logdensity: torch.Tensor
    logdensity_grad: torch.Tensor
    volume_adjustment: torch.Tensor


class DynamicRMHMCState(NamedTuple):  # rmhmc/rmhmc.py:44-55
    position: torch.Tensor
    logdensity: torch.Tensor
    logdensity_grad: torch.Tensor
    random_generator_arg: torch.Tensor


class DynamicLMCState(NamedTuple):  # lmcmc/lmc.py:45-57, lmcmonge/lmc.py:47-60
    position: torch.Tensor
    logdensity: torch.Tensor
    logdensity_grad: torch.Tensor
    volume_adjustment: torch.Tensor
    random_generator_arg: torch.Tensor


class RMHMCIntegratorState(NamedTuple):  # rmhmc/integrators.py:25-36
    position: torch.Tensor
    momentum: torch.Tensor
    velocity: torch.Tensor
    logdensity: torch.Tensor
    logdensity_grad: torch.Tensor


class LMCIntegratorState(NamedTuple):  # lmcmc/integrators.py:26-38
    position: torch.Tensor
    momentum: torch.Tensor
    velocity: torch.Tensor
    logdensity: torch.Tensor
    logdensity_grad: torch.Tensor
    volume_adjustment: torch.Tensor


class Proposal(NamedTuple):  # mcmc/proposal.py:22-40
    state: NamedTuple
    energy: torch.Tensor
    weight: torch.Tensor
    sum_log_p_accept: torch.Tensor


class RMHMCInfo(NamedTuple):
    momentum: torch.Tensor
    acceptance_rate: torch.Tensor
    is_accepted: torch.Tensor
    is_divergent: torch.Tensor
    energy: torch.Tensor
    proposal: Proposal
    num_integration_steps: int


class LMCInfo(NamedTuple):
    velocity: torch.Tensor
    acceptance_rate: torch.Tensor
    is_accepted: torch.Tensor
    is_divergent: torch.Tensor
    energy: torch.Tensor
    proposal: Proposal
    num_integration_steps: int


# ------------------------------------------------------------------------------------ engine


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return N.F32
    if t.dtype == torch.float64:
        return N.F64
    raise TypeError(f"unsupported dtype {t.dtype}")


def _check_position(position, target):
    if not isinstance(position, torch.Tensor):
        raise TypeError("position must be a torch tensor on a CUDA device, shape (C, D)")
    if position.ndim != 2 or position.shape[1] != target.D:
        raise ValueError(f"position must have shape (C, {target.D}); got {tuple(position.shape)}")
    if not position.is_cuda:
        raise N.NativeError("geomjax_b200 operates on CUDA tensors only (no CPU fallback)")
    return position.contiguous()


def _check_field(name, t, shape, ref):
    if not isinstance(t, torch.Tensor) or tuple(t.shape) != tuple(shape) or t.dtype != ref.dtype or t.device != ref.device:
        got = (tuple(t.shape), t.dtype, t.device) if isinstance(t, torch.Tensor) else type(t).__name__
        raise ValueError(f"{name} must be a {ref.dtype} tensor of shape {tuple(shape)} on {ref.device}; got {got}")


def _init(position, logdensity_fn, with_volume: bool):
    target = as_target(logdensity_fn)
    q = _check_position(position, target)
    C_ = q.shape[0]
    logp = torch.empty(C_, dtype=q.dtype, device=q.device)
    grad = torch.empty_like(q)
    vol = torch.empty(C_, dtype=q.dtype, device=q.device) if with_volume else None
    st = N.State(N.ptr(q), N.ptr(logp), N.ptr(grad), N.ptr(vol))
    desc = target.c_struct()
    with torch.cuda.device(q.device):
        N.check(N.lib().gb200_init(C.byref(desc), st, C_, _dtype_code(q), N.stream_ptr()))
    return (q, logp, grad, vol)


class _Engine:
    """One configured transition kernel (sampler id + target + parameters)."""

    def __init__(self, sampler: int, target, step_size, num_integration_steps, *,
                 divergence_threshold=1000, inverse_mass_matrix=None, alpha2=0.001,
                 half_step="omega", fp_convergence_tol=1e-6, fp_divergence_tol=1e10, fp_max_iters=100,
                 lanes_per_chain=0, logreg_path="lockstep"):
        self.sampler = sampler
        self.target = as_target(target)
        self.step_size = step_size
        # a (C,) integer tensor = per-chain step counts of the dynamic kernels; the scalar passed down is their bound
        self.steps_per_chain = None
        if isinstance(num_integration_steps, torch.Tensor) and num_integration_steps.ndim >= 1:
            self.steps_per_chain = num_integration_steps
            num_integration_steps = int(num_integration_steps.max()) if num_integration_steps.numel() else 0
        self.L = int(num_integration_steps)
        self.divergence_threshold = float(divergence_threshold)
        if isinstance(inverse_mass_matrix, (torch.Tensor, np.ndarray)) and inverse_mass_matrix.ndim == 1 \
                and bool((torch.as_tensor(inverse_mass_matrix) == 1).all()):
            inverse_mass_matrix = None  # ones: the C ABI takes NULL and runs the unit-mass kernels
        self.inverse_mass_matrix = inverse_mass_matrix
        self.alpha2 = float(alpha2)
        self.half_step = N.HALF_STEP[half_step]
        self.fp = (float(fp_convergence_tol), float(fp_divergence_tol), int(fp_max_iters))
        self.lanes_per_chain = int(lanes_per_chain)
        self.with_volume = sampler != N.RMHMC
        if logreg_path not in ("lockstep", "per_chain"):
            raise ValueError("logreg_path must be 'lockstep' or 'per_chain'")
        self.logreg_path = logreg_path

    def plan(self, chains: int, device):
        """Lock-step plan (workspace + CUDA graph) for `chains` chains on `device`, created once and shared."""
        from .plan import cached_plan
        return cached_plan(self.target, chains, device)

    def _params(self, ref: torch.Tensor):
        p = N.KernelParams()
        keep = []
        ss = self.step_size
        if isinstance(ss, torch.Tensor) and ss.ndim >= 1:
            ss = ss.to(device=ref.device, dtype=ref.dtype).contiguous()
            if ss.shape != (ref.shape[0],):
                raise ValueError("per-chain step_size must have shape (C,)")
            p.step_size_per_chain = N.ptr(ss)
            keep.append(ss)
            p.step_size = 0.0
        else:
            p.step_size = float(ss)
        p.num_integration_steps = self.L
        if self.steps_per_chain is not None:
            spc = self.steps_per_chain.to(device=ref.device, dtype=torch.int32).contiguous()
            if spc.shape != (ref.shape[0],):
                raise ValueError("per-chain num_integration_steps must have shape (C,)")
            p.num_integration_steps_per_chain = N.ptr(spc)
            keep.append(spc)
        p.threefry_mode = grandom.threefry_mode()
        p.divergence_threshold = self.divergence_threshold
        p.fp_convergence_tol, p.fp_divergence_tol, p.fp_max_iters = self.fp
        p.half_step = self.half_step
        p.alpha2 = self.alpha2
        if self.inverse_mass_matrix is not None:
            im = self.inverse_mass_matrix
            if isinstance(im, np.ndarray):
                im = torch.from_numpy(im)
            im = im.to(device=ref.device, dtype=ref.dtype).contiguous()
            if im.ndim == 2 and im.shape == (ref.shape[0], self.target.D):
                # batched (vmapped) form: one diagonal inverse mass matrix per chain
                p.inverse_mass_per_chain = 1
            elif im.ndim != 1:
                # lmcmonge/metrics.py:145-153: only a diagonal (1-d) mass matrix is accepted
                raise ValueError("The mass matrix has the wrong number of dimensions:"
                                 f" expected 1, got {im.ndim}.")
            elif im.shape[0] != self.target.D:
                raise ValueError("inverse_mass_matrix must have shape (D,)")
            p.inverse_mass_matrix = N.ptr(im)
            keep.append(im)
        p.dtype = _dtype_code(ref)
        p.lanes_per_chain = self.lanes_per_chain
        return p, keep

    def launch(self, state, key_source: N.KeySource, *, want_info=True, out_state=None, opts=None,
               extra_info=False):
        q = _check_position(state[0], self.target)
        C_, D = q.shape
        dev, dt = q.device, q.dtype
        fields = [q] + [t.contiguous() for t in state[1:] if t is not None]  # rmhmc carries no volume_adjustment
        want = 4 if self.with_volume else 3
        if len(fields) < want:
            raise ValueError(f"state needs {want} fields (position, logdensity, logdensity_grad"
                             + (", volume_adjustment)" if self.with_volume else ")"))
        for name, t, shape in zip(("position", "logdensity", "logdensity_grad", "volume_adjustment"), fields[:want],
                                  ((C_, D), (C_,), (C_, D), (C_,))):
            _check_field(name, t, shape, q)
        if out_state is not None:
            for name, t, shape in zip(("position", "logdensity", "logdensity_grad", "volume_adjustment"),
                                      [t for t in out_state if t is not None][:want], ((C_, D), (C_,), (C_, D), (C_,))):
                _check_field("out_state." + name, t, shape, q)
                if not t.is_contiguous():
                    raise ValueError(f"out_state.{name} must be contiguous")
        if out_state is None:
            out = [torch.empty_like(t) for t in fields]
        else:
            out = [t for t in out_state if t is not None]
        if not self.with_volume:
            fields = fields[:3] + [None]
            out = out[:3] + [None]
        st_in = N.State(*[N.ptr(t) for t in fields])
        st_out = N.State(*[N.ptr(t) for t in out])
        info_c = N.Info()
        info_t = {}
        if want_info:
            def new(name, shape, dtype=dt):
                t = torch.empty(shape, dtype=dtype, device=dev)
                info_t[name] = t
                setattr(info_c, name, N.ptr(t))
            new("momentum", (C_, D)); new("acceptance_rate", (C_,))
            new("is_accepted", (C_,), torch.uint8); new("is_divergent", (C_,), torch.uint8)
            new("energy", (C_,)); new("proposal_position", (C_, D)); new("proposal_momentum", (C_, D))
            new("proposal_velocity", (C_, D)); new("proposal_logdensity", (C_,))
            new("proposal_logdensity_grad", (C_, D)); new("proposal_weight", (C_,))
            if self.with_volume:
                new("proposal_volume_adjustment", (C_,))
            if extra_info:
                new("initial_energy", (C_,)); new("accept_uniform", (C_,)); new("noise", (C_, D))
                if self.sampler == N.RMHMC:
                    new("fp_iters", (C_,), torch.int32)
        p, keep = self._params(q)
        desc = self.target.c_struct()
        if self.target.kind == N.TARGET_LOGREG and self.sampler == N.RMHMC:
            if opts is None:
                opts = N.RunOpts()
            if self.logreg_path == "lockstep" and dt == torch.float32:
                # the product path: every map evaluation for all chains at once on the tcgen05 GEMMs
                opts.plan = self.plan(C_, dev).handle
            else:
                # CTA-per-chain FP32 kernels; the scratch word is their dynamic chain hand-out counter
                ws = torch.empty(4, dtype=torch.int32, device=dev)
                keep.append(ws)
                opts.workspace, opts.workspace_bytes = N.ptr(ws), ws.numel() * 4
        with torch.cuda.device(dev):
            N.check(N.lib().gb200_step(self.sampler, C.byref(p), C.byref(desc), C.byref(key_source), st_in, st_out,
                                       C.byref(info_c) if want_info else None,
                                       C.byref(opts) if opts is not None else None, C_, N.stream_ptr()))
        return out, info_t

    # -- public pieces ---------------------------------------------------------------
    def make_state(self, fields):
        return LMCState(*fields) if self.with_volume else RMHMCState(*fields[:3])

    def make_info(self, t):
        if self.with_volume:
            ist = LMCIntegratorState(t["proposal_position"], t["proposal_momentum"], t["proposal_velocity"],
                                     t["proposal_logdensity"], t["proposal_logdensity_grad"],
                                     t["proposal_volume_adjustment"])
            cls = LMCInfo
        else:
            ist = RMHMCIntegratorState(t["proposal_position"], t["proposal_momentum"], t["proposal_velocity"],
                                       t["proposal_logdensity"], t["proposal_logdensity_grad"])
            cls = RMHMCInfo
        w = t["proposal_weight"]
        prop = Proposal(ist, t["energy"], w, torch.clamp(w, max=0.0))
        return cls(t["momentum"], t["acceptance_rate"], t["is_accepted"].bool(), t["is_divergent"].bool(),
                   t["energy"], prop, self.L)

    def step(self, rng_key, state):
        keys = grandom._keys_tensor(rng_key, state[0].device)
        if keys.shape != (state[0].shape[0], 2):
            raise ValueError(f"rng_key must have shape (C, 2) = ({state[0].shape[0]}, 2); got {tuple(keys.shape)} "
                             "(one key per chain, as in jax.vmap(kernel)(keys, states))")
        ks = N.KeySource()
        ks.keys = N.ptr(keys)
        ks.num_transitions = 1
        out, info = self.launch(state, ks, want_info=True)
        return self.make_state(out), self.make_info(info)


class _StepFn:
    """``SamplingAlgorithm.step``; also exposes the fused multi-transition launch."""

    def __init__(self, engine: _Engine):
        self.engine = engine

    def __call__(self, rng_key, state):
        return self.engine.step(rng_key, state)


def run_fused(step_fn, rng_key, state, num_samples: int, *, first: int = 0, total: Optional[int] = None,
              chain_offset: int = 0, total_chains: Optional[int] = None, return_samples: bool = False,
              return_accept=False, dual_averaging=None, da_target: float = 0.8, inplace: bool = False,
              out_samples=None, out_accept=None):
    """Run ``num_samples`` transitions in ONE launch with in-kernel key derivation
    ``split(split(rng_key, total)[t], total_chains)[chain_offset + c]`` -- the driver loop of
    examples/funnel/main.py:7-25 (``inference_loop_multiple_chains``) without the host round trip.
    Identical, bit for bit, to calling ``step`` ``num_samples`` times with those keys.
    Returns (state, samples[T, C, D] | None, acceptance_rate[T, C] | None); with ``return_accept="mean"`` the third
    item is the per-chain mean acceptance rate [C], accumulated inside the kernels (no [T, C] buffer).
    ``out_samples`` / ``out_accept``: caller-owned result buffers (allocation stays outside a timed region)."""
    eng = step_fn.engine if isinstance(step_fn, _StepFn) else step_fn
    q = state[0]
    C_, D = q.shape
    total = num_samples + first if total is None else total
    total_chains = C_ + chain_offset if total_chains is None else total_chains
    ks = N.KeySource()
    ks.keys = None
    ks.root_key[0], ks.root_key[1] = int(rng_key[0]), int(rng_key[1])
    ks.first_transition, ks.num_transitions, ks.total_transitions = first, num_samples, total
    ks.chain_offset, ks.total_chains = chain_offset, total_chains
    opts = N.RunOpts()
    samples = acc = None
    if return_samples or out_samples is not None:
        samples = _result_buffer(out_samples, (num_samples, C_, D), q)
        opts.samples = N.ptr(samples)
    if return_accept == "mean":
        acc = _result_buffer(out_accept, (C_,), q)
        acc.zero_()
        opts.accept_sum = N.ptr(acc)
    elif return_accept or out_accept is not None:
        acc = _result_buffer(out_accept, (num_samples, C_), q)
        opts.sample_accept = N.ptr(acc)
    if dual_averaging is not None:
        opts.dual_averaging = N.ptr(dual_averaging)
        opts.da_target, opts.da_t0, opts.da_gamma, opts.da_kappa = da_target, 10.0, 0.05, 0.75
    out, _ = eng.launch(state, ks, want_info=False, opts=opts, out_state=list(state) if inplace else None)
    if return_accept == "mean":
        acc.mul_(1.0 / num_samples)
    return eng.make_state(out), samples, acc


def _result_buffer(buf, shape, like):
    if buf is None:
        return torch.empty(shape, dtype=like.dtype, device=like.device)
    if tuple(buf.shape) != tuple(shape) or buf.dtype != like.dtype or buf.device != like.device or not buf.is_contiguous():
        raise ValueError(f"result buffer must be a contiguous {like.dtype} tensor of shape {tuple(shape)} on {like.device}")
    return buf


# ------------------------------------------------------------------------------------ façades


def _kernel_fn(sampler_id, kind):
    """build_kernel(...) -> kernel(rng_key, state, logdensity_fn, step_size, metric_fn|inverse_mass_matrix,
    num_integration_steps[, alpha2])."""

    def build_kernel(integrator: Callable = None, divergence_threshold: float = 1000):
        default = {N.RMHMC: integrators.implicit_midpoint, N.LMC: integrators.lan_integrator,
                   N.LMCMONGE: integrators.lan_integrator_monge}[sampler_id]
        integ = integrators.resolve(default if integrator is None else integrator, sampler_id)

        if sampler_id == N.LMCMONGE:
            def kernel(rng_key, state, logdensity_fn, step_size, inverse_mass_matrix, num_integration_steps,
                       alpha2):
                eng = _Engine(sampler_id, logdensity_fn, step_size, num_integration_steps,
                              divergence_threshold=divergence_threshold, inverse_mass_matrix=inverse_mass_matrix,
                              alpha2=alpha2, **integ.kwargs)
                return eng.step(rng_key, state)
        else:
            def kernel(rng_key, state, logdensity_fn, step_size, metric_fn, num_integration_steps):
                tgt = _merge_target(logdensity_fn, metric_fn)
                eng = _Engine(sampler_id, tgt, step_size, num_integration_steps,
                              divergence_threshold=divergence_threshold, **integ.kwargs)
                return eng.step(rng_key, state)
        return kernel

    return build_kernel


def _merge_target(logdensity_fn, metric_fn):
    t = as_target(logdensity_fn)
    if metric_fn is None or metric_fn is t:
        return t
    if isinstance(metric_fn, str):
        return t.with_metric(metric_fn)
    m = as_target(metric_fn)
    pad = lambda ps: tuple(ps) + (0.0,) * (7 - len(ps))
    if (m.kind, m.D, pad(m.base_params)) != (t.kind, t.D, pad(t.base_params)):
        raise ValueError("metric_fn must be the metric of the same target descriptor (or 'identity')")
    return m


class rmhmc:
    """geomjax/rmhmc/rmhmc.py:247-311."""

    @staticmethod
    def init(position, logdensity_fn):
        return RMHMCState(*_init(position, logdensity_fn, False)[:3])

    build_kernel = staticmethod(_kernel_fn(N.RMHMC, "rmhmc"))

    def __new__(cls, logdensity_fn, step_size, metric_fn, num_integration_steps, *,
                divergence_threshold: int = 1000, integrator: Callable = None, lanes_per_chain: int = 0,
                logreg_path: str = "lockstep"):
        integ = integrators.resolve(integrators.implicit_midpoint if integrator is None else integrator, N.RMHMC)
        eng = _Engine(N.RMHMC, _merge_target(logdensity_fn, metric_fn), step_size, num_integration_steps,
                      divergence_threshold=divergence_threshold, lanes_per_chain=lanes_per_chain,
                      logreg_path=logreg_path, **integ.kwargs)
        return SamplingAlgorithm(lambda position: cls.init(position, eng.target), _StepFn(eng))


class lmc:
    """geomjax/lmcmc/lmc.py:255-346."""

    @staticmethod
    def init(position, logdensity_fn):
        return LMCState(*_init(position, logdensity_fn, True))

    build_kernel = staticmethod(_kernel_fn(N.LMC, "lmc"))

    def __new__(cls, logdensity_fn, step_size, metric_fn, num_integration_steps, *,
                divergence_threshold: int = 1000, integrator: Callable = None, lanes_per_chain: int = 0):
        integ = integrators.resolve(integrators.lan_integrator if integrator is None else integrator, N.LMC)
        eng = _Engine(N.LMC, _merge_target(logdensity_fn, metric_fn), step_size, num_integration_steps,
                      divergence_threshold=divergence_threshold, lanes_per_chain=lanes_per_chain, **integ.kwargs)
        return SamplingAlgorithm(lambda position: cls.init(position, eng.target), _StepFn(eng))


class lmcmonge:
    """geomjax/lmcmonge/lmc.py:312-405 (``alpha2=0.001`` default :385)."""

    @staticmethod
    def init(position, logdensity_fn):
        return LMCState(*_init(position, logdensity_fn, True))

    build_kernel = staticmethod(_kernel_fn(N.LMCMONGE, "lmcmonge"))

    def __new__(cls, logdensity_fn, step_size, inverse_mass_matrix, num_integration_steps, *,
                alpha2: float = 0.001, divergence_threshold: int = 1000, integrator: Callable = None,
                lanes_per_chain: int = 0):
        integ = integrators.resolve(integrators.lan_integrator_monge if integrator is None else integrator,
                                    N.LMCMONGE)
        eng = _Engine(N.LMCMONGE, logdensity_fn, step_size, num_integration_steps,
                      divergence_threshold=divergence_threshold, inverse_mass_matrix=inverse_mass_matrix,
                      alpha2=alpha2, lanes_per_chain=lanes_per_chain, **integ.kwargs)
        return SamplingAlgorithm(lambda position: cls.init(position, eng.target), _StepFn(eng))


# ------------------------------------------------------------------------------------ dynamic kernels
# rmhmc/rmhmc.py:179-244, lmcmc/lmc.py:185-252, lmcmonge/lmc.py:240-309: the number of integration steps of each
# transition is drawn by ``integration_steps_fn(state.random_generator_arg)`` and the argument advanced by
# ``next_random_arg_fn``.  Batched over chains: ``random_generator_arg`` is (C, 2) uint32 keys (the reference's
# default, every chain its own key) or a (C,) / scalar integer counter (the Halton jitter of ChEES); the step counts
# go to the kernels as a per-chain array (gb200_kernel_params.num_integration_steps_per_chain) -- a warp runs to its
# largest count with the finished chains masked.


def _default_next_random_arg(arg):
    return grandom.split(arg, 2)[..., 1, :].contiguous()  # lambda key: jax.random.split(key)[1]


def _default_integration_steps(arg):
    return grandom.randint(arg, 1, 10)  # lambda key: jax.random.randint(key, (), 1, 10)


def _steps_tensor(steps, C_, device):
    if isinstance(steps, torch.Tensor):
        st = steps.to(device=device, dtype=torch.int32).reshape(-1)
        return st.expand(C_).contiguous() if st.numel() == 1 else st
    return int(steps)


def _dynamic_kernel_fn(sampler_id):
    def build_dynamic_kernel(integrator: Callable = None, divergence_threshold: float = 1000,
                             next_random_arg_fn: Callable = _default_next_random_arg,
                             integration_steps_fn: Callable = _default_integration_steps):
        base = _kernel_fn(sampler_id, "dynamic")(integrator, divergence_threshold)

        def finish(new, info, state):
            return type(state)(*new, next_random_arg_fn(state.random_generator_arg)), info

        if sampler_id == N.LMCMONGE:
            def kernel(rng_key, state, logdensity_fn, step_size, inverse_mass_matrix, alpha2=0.001,
                       **integration_steps_kwargs):
                steps = _steps_tensor(integration_steps_fn(state.random_generator_arg, **integration_steps_kwargs),
                                      state.position.shape[0], state.position.device)
                new, info = base(rng_key, LMCState(*state[:4]), logdensity_fn, step_size, inverse_mass_matrix, steps, alpha2)
                return finish(new, info, state)
        else:
            nfields = 3 if sampler_id == N.RMHMC else 4
            cls = RMHMCState if sampler_id == N.RMHMC else LMCState

            def kernel(rng_key, state, logdensity_fn, step_size, metric_fn, **integration_steps_kwargs):
                steps = _steps_tensor(integration_steps_fn(state.random_generator_arg, **integration_steps_kwargs),
                                      state.position.shape[0], state.position.device)
                new, info = base(rng_key, cls(*state[:nfields]), logdensity_fn, step_size, metric_fn, steps)
                return finish(new, info, state)
        return kernel

    return build_dynamic_kernel


def _random_arg(arg, C_, device):
    """one generator argument per chain: (C, 2) keys, a single (2,) key (split into C), or an integer counter"""
    if isinstance(arg, int):
        return torch.full((C_,), arg, dtype=torch.int64, device=device)
    if isinstance(arg, np.ndarray):
        arg = torch.from_numpy(np.ascontiguousarray(arg))
    arg = arg.to(device)
    if arg.dtype in (torch.uint32, torch.int32) and arg.shape == (2,):
        return grandom.split(arg[None], C_)[0].contiguous()
    return arg


class dynamic_rmhmc:
    """geomjax/rmhmc/rmhmc.py:314-376."""

    @staticmethod
    def init(position, logdensity_fn, random_generator_arg):
        q, l, gr = _init(position, logdensity_fn, False)[:3]
        return DynamicRMHMCState(q, l, gr, _random_arg(random_generator_arg, q.shape[0], q.device))

    build_kernel = staticmethod(_dynamic_kernel_fn(N.RMHMC))

    def __new__(cls, logdensity_fn, step_size, metric_fn, *, divergence_threshold: int = 1000, integrator: Callable = None,
                next_random_arg_fn: Callable = _default_next_random_arg,
                integration_steps_fn: Callable = _default_integration_steps):
        kernel = cls.build_kernel(integrator, divergence_threshold, next_random_arg_fn, integration_steps_fn)
        return SamplingAlgorithm(lambda position, random_generator_arg: cls.init(position, logdensity_fn, random_generator_arg),
                                 lambda rng_key, state: kernel(rng_key, state, logdensity_fn, step_size, metric_fn))


class dynamic_lmc:
    """geomjax/lmcmc/lmc.py:349-408."""

    @staticmethod
    def init(position, logdensity_fn, random_generator_arg):
        f = _init(position, logdensity_fn, True)
        return DynamicLMCState(*f, _random_arg(random_generator_arg, f[0].shape[0], f[0].device))

    build_kernel = staticmethod(_dynamic_kernel_fn(N.LMC))

    def __new__(cls, logdensity_fn, step_size, metric_fn, *, divergence_threshold: int = 1000, integrator: Callable = None,
                next_random_arg_fn: Callable = _default_next_random_arg,
                integration_steps_fn: Callable = _default_integration_steps):
        kernel = cls.build_kernel(integrator, divergence_threshold, next_random_arg_fn, integration_steps_fn)
        return SamplingAlgorithm(lambda position, random_generator_arg: cls.init(position, logdensity_fn, random_generator_arg),
                                 lambda rng_key, state: kernel(rng_key, state, logdensity_fn, step_size, metric_fn))


class dynamic_lmcmonge:
    """geomjax/lmcmonge/lmc.py dynamic_lmc (exported as geomjax.dynamic_lmcmonge)."""

    @staticmethod
    def init(position, logdensity_fn, random_generator_arg):
        f = _init(position, logdensity_fn, True)
        return DynamicLMCState(*f, _random_arg(random_generator_arg, f[0].shape[0], f[0].device))

    build_kernel = staticmethod(_dynamic_kernel_fn(N.LMCMONGE))

    def __new__(cls, logdensity_fn, step_size, inverse_mass_matrix, *, alpha2: float = 0.001, divergence_threshold: int = 1000,
                integrator: Callable = None, next_random_arg_fn: Callable = _default_next_random_arg,
                integration_steps_fn: Callable = _default_integration_steps):
        kernel = cls.build_kernel(integrator, divergence_threshold, next_random_arg_fn, integration_steps_fn)
        return SamplingAlgorithm(lambda position, random_generator_arg: cls.init(position, logdensity_fn, random_generator_arg),
                                 lambda rng_key, state: kernel(rng_key, state, logdensity_fn, step_size, inverse_mass_matrix, alpha2))
