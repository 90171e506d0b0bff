"""Plan of the lock-step rmhmc sampler on the logistic-regression target (``gb200_plan`` in include/geomb200.h):
the device workspace and the captured CUDA graph (one round of the rolling batch as the body of a WHILE node).

Reference semantics: rmhmc/rmhmc.py:131-174 and rmhmc/integrators.py:53-156 under ``jax.vmap``; see
csrc/rmhmc_lockstep.cu.  PyTorch only supplies the device memory."""
from __future__ import annotations

import ctypes as C
import os
import weakref

import torch

from . import _native as N


_PLANS: "dict[tuple, LockstepPlan]" = {}
_MAX_PLANS = 4


def cached_plan(target, chains: int, device) -> "LockstepPlan":
    """One plan per (design matrix, responses, shape, prior, chain count, device): ``build_kernel()``-style callers
    construct an engine per call and must not re-capture the graph every transition.  Oldest plan evicted."""
    dev = torch.device(device)
    _, y, xt, _ = target._keep
    key = (int(xt.data_ptr()), int(y.data_ptr()), int(target.N), int(target.D),
           tuple(float(x) for x in target.params), int(chains), dev.index)
    pl = _PLANS.pop(key, None)
    if pl is None:
        # GEOMB200_LOCKSTEP_LOOP=1: host-sequenced rounds instead of the graph WHILE node (profilers that cannot
        # see inside conditional graph nodes); read once per plan, never on the launch path
        pl = LockstepPlan(target, chains, dev, loop_mode=int(os.environ.get("GEOMB200_LOCKSTEP_LOOP", "0")))
        while len(_PLANS) >= _MAX_PLANS:
            _PLANS.pop(next(iter(_PLANS)))
    _PLANS[key] = pl  # most recently used last
    return pl


class LockstepPlan:
    def __init__(self, target, chains: int, device, loop_mode: int = 0):
        if target.kind != N.TARGET_LOGREG or target.metric != N.METRIC_TARGET:
            raise NotImplementedError("the lock-step plan is built for the logistic-regression target with its Fisher metric")
        self.target, self.C, self.device = target, int(chains), torch.device(device)
        self.desc = target.c_struct()
        lib = N.lib()
        nbytes = int(lib.gb200_rmhmc_logreg_plan_workspace(C.byref(self.desc), self.C))
        if nbytes < 0:
            raise N.NativeError(f"geomb200: {lib.gb200_last_error().decode()}")
        self.workspace_bytes = nbytes
        self._ws = torch.zeros(nbytes + 256, dtype=torch.uint8, device=self.device)
        base = self._ws.data_ptr()
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            N.check(lib.gb200_rmhmc_logreg_plan_create(C.byref(self.desc), self.C, C.c_void_p((base + 255) // 256 * 256), nbytes,
                                                       int(loop_mode), N.stream_ptr(), C.byref(handle)))
        self.handle = handle
        self._fin = weakref.finalize(self, lib.gb200_plan_destroy, handle)

    @property
    def loop_mode(self) -> str:
        return N.lib().gb200_plan_loop_mode(self.handle).decode()

    def stats(self):
        """(rounds, chain-evaluations mod 2^31) of the last launch; synchronises the current stream."""
        r, e = C.c_int64(), C.c_int64()
        with torch.cuda.device(self.device):
            N.check(N.lib().gb200_plan_stats(self.handle, C.byref(r), C.byref(e), N.stream_ptr()))
        return int(r.value), int(e.value)

    def evaluate(self, mode: int, q, p, qi=None, pi=None, half_step: float = 0.0):
        """One round for explicit inputs (``gb200_logreg_lockstep_eval``): mode 0 = the implicit-midpoint map
        (rmhmc/integrators.py:119-142), 1 = end-of-trajectory state, 2 = map with the momentum drawn first (``p`` = z).
        Returns a dict of the outputs that mode defines."""
        f = lambda t: None if t is None else t.to(device=self.device, dtype=torch.float32).contiguous()
        q, p, qi, pi = f(q), f(p), f(qi), f(pi)
        C_, D = q.shape
        if D != self.target.D or C_ > self.C or p.shape != q.shape:
            raise ValueError(f"q, p must have shape (C <= {self.C}, {self.target.D})")
        new = lambda *s: torch.empty(s, dtype=torch.float32, device=self.device)
        out = {"velocity": new(C_, D), "logdet": new(C_)}
        if mode == 1:
            out.update(logdensity=new(C_), logdensity_grad=new(C_, D))
        else:
            out.update(q=new(C_, D), p=new(C_, D), dHdq=new(C_, D))
            if mode == 2:
                out["momentum"] = new(C_, D)
        g = lambda k: N.ptr(out.get(k))
        with torch.cuda.device(self.device):
            N.check(N.lib().gb200_logreg_lockstep_eval(self.handle, int(mode), N.ptr(q), N.ptr(p), N.ptr(qi), N.ptr(pi),
                                                       float(half_step), g("q"), g("p"), g("momentum"), g("logdensity"),
                                                       g("logdensity_grad"), g("velocity"), g("logdet"), g("dHdq"), C_,
                                                       N.stream_ptr()))
        return out
