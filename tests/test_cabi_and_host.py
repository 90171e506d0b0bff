"""CPU-side checks: the C-ABI library loads and exports every symbol of include/geomb200.h; host-only
entry points (diagnostics finalisation, flop model) against the oracle; facade error behaviour.
No kernel is launched here."""
import ctypes as C

import numpy as np
import pytest

from oracle import diagnostics as OD


def _lib():
    from geomjax_b200 import _native as N
    return N, N.lib()


def test_library_exports_every_header_symbol():
    N, lib = _lib()
    syms = N.header_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert set(N.PROTOTYPES) == set(syms)
    assert lib.gb200_version() == 100


def test_struct_sizes_match_header_layout():
    N, _ = _lib()
    assert C.sizeof(N.State) == 32 and C.sizeof(N.Info) == 16 * 8
    assert C.sizeof(N.KeySource) == 8 + 8 + 5 * 8
    assert C.sizeof(N.TargetDesc) == 16 + 8 + 64 + 32
    assert C.sizeof(N.KernelParams) == 8 + 8 + 8 + 8 + 8 + 8 + 8 + 8 + 8 + 8 + 8 + 8  # + num_integration_steps_per_chain


def _ar1(T, Cn, D, rho, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((T, Cn, D))
    for t in range(1, T):
        x[t] = rho * x[t - 1] + x[t]
    return x + rng.standard_normal((1, Cn, 1)) * 0.05


def partial_stats(x):
    """NumPy restatement of k_rhat_partial / k_ess_partial (chain-summed sufficient statistics)."""
    T, Cn, D = x.shape
    m = x.mean(0)
    v = x.var(0, ddof=1)
    stats = np.concatenate([m.sum(0), (m * m).sum(0), v.sum(0), [float(Cn)]])
    cen = x - m[None]
    acov = np.stack([(cen[: T - l] * cen[l:]).sum(0).sum(0) / T for l in range(T)])
    return stats, acov


@pytest.mark.parametrize("T,Cn,D,rho", [(200, 4, 3, 0.7), (101, 6, 2, 0.2), (64, 2, 1, 0.95)])
def test_rhat_ess_finalize_match_oracle(T, Cn, D, rho):
    N, lib = _lib()
    x = _ar1(T, Cn, D, rho, seed=T)
    stats, acov = partial_stats(x)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rhat = np.empty(D)
    assert lib.gb200_rhat_finalize(dp(stats), T, D, dp(rhat)) == 0
    np.testing.assert_allclose(rhat, np.atleast_1d(OD.potential_scale_reduction(x, 1, 0)), rtol=1e-10)
    want = np.atleast_1d(OD.effective_sample_size(x, 1, 0))
    ess = np.empty(D)
    trunc = np.zeros(D, np.uint8)
    a = np.ascontiguousarray(acov)
    assert lib.gb200_ess_finalize(dp(a), dp(stats), T, Cn, D, T, dp(ess), trunc.ctypes.data_as(C.POINTER(C.c_uint8))) == 0
    np.testing.assert_allclose(ess, want, rtol=1e-8)
    # a truncated lag window gives the same answer as long as Geyer's sequence ended inside it
    for lags in (8, 16, 32):
        a = np.ascontiguousarray(acov[:lags])
        assert lib.gb200_ess_finalize(dp(a), dp(stats), T, Cn, D, lags, dp(ess), trunc.ctypes.data_as(C.POINTER(C.c_uint8))) == 0
        ok = trunc == 0
        np.testing.assert_allclose(ess[ok], want[ok], rtol=1e-8)
    assert lib.gb200_rhat_finalize(dp(stats), 1, D, dp(rhat)) < 0  # bad argument -> negative status + message
    assert b"rhat_finalize" in lib.gb200_last_error()


def test_flop_model_and_facade_errors():
    import torch
    import geomjax_b200 as g
    N, lib = _lib()
    t = g.neal_funnel(20)
    d = t.c_struct()
    assert lib.gb200_flops_per_chain_step(N.LMCMONGE, C.byref(d)) == 59 * 20 + 60
    assert lib.gb200_flops_per_chain_step(N.LMC, C.byref(d)) == 25 * 20 + 140
    assert lib.gb200_flops_per_transition(N.LMCMONGE, C.byref(d)) == 79 * 20 + 40
    assert lib.gb200_flops_per_transition(N.LMC, C.byref(d)) == 58 * 20 + 30
    with pytest.raises(ValueError):
        g.neal_funnel(1)
    with pytest.raises(NotImplementedError):
        g.rmhmc(lambda x: -0.5 * (x ** 2).sum(), 0.1, lambda x: torch.eye(2), 4)
    with pytest.raises(NotImplementedError):
        g.lmc(t, 0.1, t, 4, integrator=lambda *a: None)
    alg = g.lmc(t, 0.1, t.fisher_metric_fn, 4)
    with pytest.raises(g._native.NativeError):
        alg.init(torch.ones((3, 20)))  # CPU tensor: there is no CPU fallback
    with pytest.raises(ValueError):
        g.lmc(t, 0.1, g.neal_funnel(21), 4)  # metric of a different target


def test_build_schedule_matches_reference_shape():
    """adaptation/window_adaptation.py:360-450: Stan windows for the default 1000 steps."""
    from geomjax_b200.adaptation import build_schedule
    s = build_schedule(1000)
    assert len(s) == 1000
    assert s[:75] == [(0, False)] * 75 and s[-50:] == [(0, False)] * 50
    ends = [i for i, (st, e) in enumerate(s) if e]
    assert ends == [99, 149, 249, 449, 949]
    assert all(st == 1 for st, _ in s[75:950])
    assert build_schedule(10) == [(0, False)] * 10
    s = build_schedule(100)
    assert len(s) == 100 and [i for i, (_, e) in enumerate(s) if e] == [89]
    # element by element against the oracle's statement-by-statement restatement of the reference
    from oracle.adaptation import build_schedule as ref
    for n in list(range(0, 260)) + [300, 500, 999, 1000, 1001, 2500, 10000]:
        assert build_schedule(n) == ref(n), n
    for args in [(400, 100, 40, 10), (1000, 10, 10, 5), (150, 75, 50, 25), (2000, 200, 100, 50)]:
        assert build_schedule(*args) == ref(*args), args


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU oracle port the driver times beside the GPU arm) prints ONE JSON
    line with the contract's keys; it must run without a GPU."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "c1",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=240, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "chain-leapfrog-steps/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["unit"] == "chain-steps/s" and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["gpu_launches"] == 0


def test_lockstep_plan_rejects_other_targets():
    """The lock-step plan is built for the logistic-regression target; anything else is refused up front."""
    import geomjax_b200 as g
    with pytest.raises(NotImplementedError):
        g.LockstepPlan(g.neal_funnel(4), 8, "cpu")
    with pytest.raises(ValueError):
        g.rmhmc(g.neal_funnel(4), 0.1, g.neal_funnel(4), 4, logreg_path="nope")


def test_bench_data_equals_the_oracle_generator():
    from bench.data import make_logreg_data
    from oracle.targets import make_logreg_data as ref
    for N_, D_, seed in [(40, 3, 0), (1000, 25, 0), (333, 17, 4)]:
        X, y = make_logreg_data(N_, D_, seed)
        Xo, yo = ref(N_, D_, seed)
        assert (X == Xo).all() and (y == yo).all()


def test_chees_host_recurrences_match_oracle():
    """Halton jitter, Adam and dual averaging of the ChEES driver (host scalars) == the oracle's restatement."""
    from geomjax_b200 import chees
    from oracle import adaptation as OA, chees as OC
    import numpy as np
    for i in range(40):
        assert chees.halton_sequence(i, 11) == OC.halton(i, 11)
    opt = chees.adam(0.05)
    st, ost = opt.init(0.3), OC.adam_init()
    for gr in [0.5, -2.0, 1e-3, 7.0, -0.1]:
        u, st = opt.update(gr, st, 0.0)
        uo, ost = OC.adam_update(gr, ost, 0.05)
        assert abs(u - uo) < 1e-15
    da, oda = chees._da_init(0.2), OA.da_init(np.float64(0.2), np.float64)
    for gr in [0.3, -0.2, 0.05, 0.6]:
        da = chees._da_update(da, gr)
        oda = OA.da_update(oda, np.float64(gr))
        assert abs(da[0] - float(oda["log_x"])) < 1e-12 and abs(da[1] - float(oda["log_x_avg"])) < 1e-12


def test_lockstep_plan_workspace_is_host_arithmetic():
    """gb200_rmhmc_logreg_plan_workspace only sizes the carve (no device call): positive, growing with the chain
    count, ~28 GB for c5 as named (131,072 chains, N = 10,000, D = 100), and refused for shapes the plan does not take."""
    import ctypes as Ct
    from geomjax_b200 import _native as N

    def ws(Nrows, D, chains, ldx=None):
        t = N.TargetDesc()
        t.kind, t.metric, t.D, t.N = N.TARGET_LOGREG, N.METRIC_TARGET, D, Nrows
        t.params[0], t.params[1] = 0.01, float(ldx if ldx is not None else (Nrows + 3) // 4 * 4)
        t.y = t.vec0 = Ct.c_void_p(256)  # never dereferenced by the size computation
        return int(N.lib().gb200_rmhmc_logreg_plan_workspace(Ct.byref(t), chains))

    sizes = [ws(1000, 25, c) for c in (1, 3, 256, 257, 2048, 16384)]
    assert all(s > 0 for s in sizes) and sizes == sorted(sizes)
    assert 20e9 < ws(10000, 100, 131072) < 40e9
    assert ws(1000, 125, 64) < 0        # D > 124
    assert ws(1000, 25, 0) < 0          # no chains
    assert ws(1000, 25, 64, ldx=1001) < 0  # ldx must be a multiple of 4
