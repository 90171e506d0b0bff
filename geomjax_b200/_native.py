"""ctypes binding of the C ABI in ``include/geomb200.h`` (libgeomb200.so).

PyTorch is only the plumbing here (device memory, streams); the signatures carry plain
pointers and sizes.  There is NO CPU fallback: if the library is missing or a tensor is not
on a CUDA device the call raises.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

_LIB = None
LIB_PATH = Path(__file__).resolve().parent / "_lib" / "libgeomb200.so"
HEADER_PATH = Path(__file__).resolve().parent.parent / "include" / "geomb200.h"

F32, F64 = 0, 1
LEGACY, PARTITIONABLE = 0, 1
RMHMC, LMC, LMCMONGE = 0, 1, 2
HALF_STEP = {"omega": 0, "omega_fixed": 1, "omegatilde": 2}
TARGET_FUNNEL, TARGET_GAUSSIAN, TARGET_BANANA, TARGET_LOGREG = 0, 1, 2, 3
METRIC_TARGET, METRIC_IDENTITY, METRIC_SOFTABS = 0, 1, 2

vp = C.c_void_p


class TargetDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("metric", C.c_int32), ("D", C.c_int32), ("reserved", C.c_int32),
                ("N", C.c_int64), ("params", C.c_double * 8),
                ("X", vp), ("y", vp), ("vec0", vp), ("vec1", vp)]


class KernelParams(C.Structure):
    _fields_ = [("step_size", C.c_double), ("step_size_per_chain", vp),
                ("num_integration_steps", C.c_int32), ("threefry_mode", C.c_int32),
                ("divergence_threshold", C.c_double), ("fp_convergence_tol", C.c_double),
                ("fp_divergence_tol", C.c_double), ("fp_max_iters", C.c_int32), ("half_step", C.c_int32),
                ("alpha2", C.c_double), ("inverse_mass_matrix", vp),
                ("dtype", C.c_int32), ("lanes_per_chain", C.c_int32),
                ("inverse_mass_per_chain", C.c_int32), ("reserved", C.c_int32),
                ("num_integration_steps_per_chain", vp)]


class State(C.Structure):
    _fields_ = [("position", vp), ("logdensity", vp), ("logdensity_grad", vp), ("volume_adjustment", vp)]


INFO_FIELDS = ["momentum", "acceptance_rate", "is_accepted", "is_divergent", "energy",
               "proposal_position", "proposal_momentum", "proposal_velocity", "proposal_logdensity",
               "proposal_logdensity_grad", "proposal_volume_adjustment", "proposal_weight",
               "initial_energy", "fp_iters", "accept_uniform", "noise"]


class Info(C.Structure):
    _fields_ = [(n, vp) for n in INFO_FIELDS]


class KeySource(C.Structure):
    _fields_ = [("keys", vp), ("root_key", C.c_uint32 * 2), ("first_transition", C.c_int64),
                ("num_transitions", C.c_int64), ("total_transitions", C.c_int64),
                ("chain_offset", C.c_int64), ("total_chains", C.c_int64)]


class RunOpts(C.Structure):
    _fields_ = [("samples", vp), ("sample_accept", vp), ("noise_override", vp), ("uniform_override", vp),
                ("dual_averaging", vp), ("da_target", C.c_double), ("da_t0", C.c_double),
                ("da_gamma", C.c_double), ("da_kappa", C.c_double),
                ("workspace", vp), ("workspace_bytes", C.c_int64), ("plan", vp), ("accept_sum", vp)]


_i32, _i64, _dbl = C.c_int32, C.c_int64, C.c_double
_P = C.POINTER

PROTOTYPES = {
    "gb200_version": (C.c_int, []),
    "gb200_kernel_launches": (C.c_longlong, []),
    "gb200_last_error": (C.c_char_p, []),
    "gb200_threefry_split": (C.c_int, [vp, vp, _i64, _i32, _i32, vp]),
    "gb200_random_bits": (C.c_int, [vp, vp, _i64, _i32, _i32, vp]),
    "gb200_uniform_f32": (C.c_int, [vp, vp, _i64, _i32, _i32, vp]),
    "gb200_normal_f32": (C.c_int, [vp, vp, _i64, _i32, _i32, vp]),
    "gb200_chain_keys": (C.c_int, [_P(C.c_uint32), _i64, _i64, _i64, _i64, vp, _i64, _i32, vp]),
    "gb200_init": (C.c_int, [_P(TargetDesc), State, _i64, _i32, vp]),
    "gb200_step": (C.c_int, [_i32, _P(KernelParams), _P(TargetDesc), _P(KeySource), State, State,
                             _P(Info), _P(RunOpts), _i64, vp]),
    "gb200_rmhmc_step": (C.c_int, [_P(KernelParams), _P(TargetDesc), _P(KeySource), State, State,
                                   _P(Info), _P(RunOpts), _i64, vp]),
    "gb200_lmc_step": (C.c_int, [_P(KernelParams), _P(TargetDesc), _P(KeySource), State, State,
                                 _P(Info), _P(RunOpts), _i64, vp]),
    "gb200_lmcmonge_step": (C.c_int, [_P(KernelParams), _P(TargetDesc), _P(KeySource), State, State,
                                      _P(Info), _P(RunOpts), _i64, vp]),
    "gb200_dual_averaging_init": (C.c_int, [vp, vp, _i64, _i32, vp]),
    "gb200_dual_averaging_update": (C.c_int, [vp, vp, _dbl, _dbl, _dbl, _dbl, _i64, _i32, vp]),
    "gb200_rhat_partial": (C.c_int, [vp, _i64, _i64, _i32, vp, _i32, vp]),
    "gb200_rhat_finalize": (C.c_int, [_P(_dbl), _i64, _i32, _P(_dbl)]),
    "gb200_ess_partial": (C.c_int, [vp, _i64, _i64, _i32, _i32, vp, _i32, vp]),
    "gb200_ess_finalize": (C.c_int, [_P(_dbl), _P(_dbl), _i64, _i64, _i32, _i32, _P(_dbl), _P(C.c_uint8)]),
    "gb200_logreg_fisher_metric": (C.c_int, [_P(TargetDesc), vp, vp, vp, _i64, _i64, _i32, vp]),
    "gb200_logreg_fisher_metric_workspace": (_i64, [_P(TargetDesc), _i64]),
    "gb200_logreg_quadform": (C.c_int, [_P(TargetDesc), vp, vp, _i64, vp, _i64, _i64, _i32, vp]),
    "gb200_logreg_quadform_workspace": (_i64, [_P(TargetDesc), _i64]),
    "gb200_rmhmc_logreg_plan_workspace": (_i64, [_P(TargetDesc), _i64]),
    "gb200_rmhmc_logreg_plan_create": (C.c_int, [_P(TargetDesc), _i64, vp, _i64, _i32, vp, _P(vp)]),
    "gb200_plan_destroy": (C.c_int, [vp]),
    "gb200_plan_loop_mode": (C.c_char_p, [vp]),
    "gb200_plan_stats": (C.c_int, [vp, _P(_i64), _P(_i64), vp]),
    "gb200_logreg_lockstep_eval": (C.c_int, [vp, _i32, vp, vp, vp, vp, _dbl, vp, vp, vp, vp, vp, vp, vp, vp, _i64, vp]),
    "gb200_stream_diag_workspace": (_i64, [_i64, _i32, _i32]),
    "gb200_stream_diag_update": (C.c_int, [vp, vp, _i64, _i64, _i32, _i32, _i64, _i32, vp]),
    "gb200_stream_diag_partial": (C.c_int, [vp, _i64, _i64, _i32, _i32, _i32, vp, vp, vp]),
    "gb200_chees_moments": (C.c_int, [vp, vp, vp, vp, _i64, _i32, vp, vp]),
    "gb200_chees_gradient": (C.c_int, [vp, vp, vp, vp, vp, vp, _i64, _i32, vp, vp]),
    "gb200_fp32_peak_kernel": (C.c_int, [vp, _i32, _i32, _i64, vp]),
    "gb200_flops_per_chain_step": (_dbl, [_i32, _P(TargetDesc)]),
    "gb200_flops_per_transition": (_dbl, [_i32, _P(TargetDesc)]),
}


def header_symbols() -> list[str]:
    """Every function ``include/geomb200.h`` declares."""
    text = HEADER_PATH.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gb200_[a-z0-9_]+)\s*\(", text)))


class NativeLibraryMissing(RuntimeError):
    pass


def lib():
    """Load libgeomb200.so (built by ``python -m geomjax_b200.build`` / ``__graft_entry__.build``)."""
    global _LIB
    if _LIB is None:
        if not LIB_PATH.exists():
            raise NativeLibraryMissing(
                f"{LIB_PATH} is missing: build it with `python -m geomjax_b200.build`. "
                "geomjax_b200 has no CPU fallback.")
        l = C.CDLL(str(LIB_PATH))
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = l
    return _LIB


class NativeError(RuntimeError):
    pass


def check(rc: int):
    if rc != 0:
        raise NativeError(f"geomb200 error {rc}: {lib().gb200_last_error().decode()}")


def ptr(t):
    """Device pointer of a CUDA tensor (or NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise NativeError("geomjax_b200 operates on CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise NativeError("tensor must be contiguous")
    return t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
