"""ctypes binding of the C++/OpenMP CPU restatement (oracle/cpp/) -- the CPU baseline BASELINE.md section 4(a) and
SURVEY 8(d) specify.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): imported by tests/, smoke() and bench.py's
cpu_baseline / --impl reference legs; the NumPy oracle (oracle/samplers.py) is the specification it is tested against."""
from __future__ import annotations

import ctypes as C

import numpy as np

LMCMONGE, LMC, RMHMC = 0, 1, 2
FUNNEL, LOGREG = 0, 1
HALF_STEP = {"omega": 0, "omega_fixed": 1, "omegatilde": 2}
_fp = C.POINTER(C.c_float)


class Problem(C.Structure):
    _fields_ = [("sampler", C.c_int32), ("target", C.c_int32), ("D", C.c_int32), ("L", C.c_int32), ("N", C.c_int32),
                ("half_step", C.c_int32), ("fp_max_iters", C.c_int32), ("threads", C.c_int32),
                ("step_size", C.c_double), ("sigma", C.c_double), ("alpha2", C.c_double),
                ("prior_precision", C.c_double), ("divergence_threshold", C.c_double), ("fp_tol", C.c_double),
                ("fp_div_tol", C.c_double), ("X", C.c_void_p), ("y", C.c_void_p), ("inv_mass", C.c_void_p)]


class Info(C.Structure):
    _fields_ = [("draw", C.c_void_p), ("acceptance_rate", C.c_void_p), ("is_accepted", C.c_void_p),
                ("energy", C.c_void_p), ("initial_energy", C.c_void_p), ("proposal_position", C.c_void_p),
                ("accept_uniform", C.c_void_p), ("fp_iters", C.c_void_p)]


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        from .cpp.build import LIB, build
        try:
            path = build()
        except Exception:  # no compiler on this box: use the prebuilt library that travelled with the tree
            path = LIB
        l = C.CDLL(str(path))
        l.ocpu_max_threads.restype = C.c_int
        l.ocpu_uniform.restype = C.c_float
        _LIB = l
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class CpuSampler:
    """One configured transition kernel on the host cores (float32)."""

    def __init__(self, sampler: str, D: int, step_size: float, num_integration_steps: int, *, sigma=3.0, alpha2=1e-3,
                 half_step="omega", inverse_mass_matrix=None, X=None, y=None, prior_precision=0.01, threads=0,
                 divergence_threshold=1000.0, fp_tol=1e-6, fp_div_tol=1e10, fp_max_iters=100):
        p = Problem()
        p.sampler = {"lmcmonge": LMCMONGE, "lmc": LMC, "rmhmc": RMHMC}[sampler]
        p.target = LOGREG if X is not None else FUNNEL
        p.D, p.L, p.threads = int(D), int(num_integration_steps), int(threads)
        p.half_step, p.fp_max_iters = HALF_STEP[half_step], int(fp_max_iters)
        p.step_size, p.sigma, p.alpha2, p.prior_precision = float(step_size), float(sigma), float(alpha2), float(prior_precision)
        p.divergence_threshold, p.fp_tol, p.fp_div_tol = float(divergence_threshold), float(fp_tol), float(fp_div_tol)
        self._keep = []
        if X is not None:
            X = np.ascontiguousarray(X, np.float32)
            y = np.ascontiguousarray(y, np.float32)
            p.N, p.X, p.y = X.shape[0], _ptr(X), _ptr(y)
            self._keep += [X, y]
        if inverse_mass_matrix is not None:
            im = np.ascontiguousarray(inverse_mass_matrix, np.float32)
            p.inv_mass = _ptr(im)
            self._keep.append(im)
        self.p, self.D, self.with_volume = p, int(D), sampler != "rmhmc"

    @property
    def threads(self):
        return self.p.threads or lib().ocpu_max_threads()

    def init(self, position):
        q = np.ascontiguousarray(position, np.float32).copy()
        Cn = q.shape[0]
        st = [q, np.empty(Cn, np.float32), np.empty_like(q), np.empty(Cn, np.float32) if self.with_volume else None]
        rc = lib().ocpu_init(C.byref(self.p), C.c_int64(Cn), _ptr(q), _ptr(st[1]), _ptr(st[2]), _ptr(st[3]))
        if rc:
            raise RuntimeError(f"ocpu_init: {rc}")
        return st

    def step(self, keys, state, want_info=True):
        """One transition per chain, in place on `state`; returns the Info arrays as a dict."""
        keys = np.ascontiguousarray(keys, np.uint32)
        Cn = state[0].shape[0]
        out, info = {}, Info()
        if want_info:
            out = dict(draw=np.empty((Cn, self.D), np.float32), acceptance_rate=np.empty(Cn, np.float32),
                       is_accepted=np.empty(Cn, np.uint8), energy=np.empty(Cn, np.float32),
                       initial_energy=np.empty(Cn, np.float32), proposal_position=np.empty((Cn, self.D), np.float32),
                       accept_uniform=np.empty(Cn, np.float32), fp_iters=np.empty(Cn, np.int32))
            for k, v in out.items():
                setattr(info, k, _ptr(v))
        rc = lib().ocpu_step(C.byref(self.p), C.c_int64(Cn), _ptr(keys), *[_ptr(a) for a in state],
                             C.byref(info) if want_info else None)
        if rc:
            raise RuntimeError(f"ocpu_step: {rc}")
        return out

    def run(self, root_key, state, T, *, first=0, total=None, chain_offset=0, total_chains=None):
        """T transitions with the example's key tree; returns the mean acceptance rate."""
        Cn = state[0].shape[0]
        root = np.ascontiguousarray(root_key, np.uint32)
        acc = C.c_double()
        rc = lib().ocpu_run(C.byref(self.p), C.c_int64(Cn), _ptr(root), C.c_int64(first), C.c_int64(T),
                            C.c_int64(first + T if total is None else total), C.c_int64(chain_offset),
                            C.c_int64(Cn + chain_offset if total_chains is None else total_chains),
                            *[_ptr(a) for a in state], C.byref(acc))
        if rc:
            raise RuntimeError(f"ocpu_run: {rc}")
        return acc.value
