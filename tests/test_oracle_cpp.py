"""The C++/OpenMP CPU baseline (oracle/cpp/) against the NumPy oracle, which is pinned to
the reference (tests/test_oracle_pins.py): draws bit-exact, one transition within float32 round-off."""
import numpy as np
import pytest

from oracle import cpu, prng as P, samplers as S, targets as T


def test_prng_bit_exact():
    l = cpu.lib()
    for seed, n in [(0, 1), (1, 2), (42, 20), (7, 25), (9, 101)]:
        k = P.split(P.key(seed), 3)[1]
        out = np.empty(n, np.float32)
        l.ocpu_normal(cpu._ptr(np.ascontiguousarray(k)), n, cpu._ptr(out))
        np.testing.assert_array_equal(out.view(np.uint32), P.normal(k, (n,)).view(np.uint32))
        assert np.float32(l.ocpu_uniform(cpu._ptr(np.ascontiguousarray(k)))) == P.uniform(k, ())
        got = np.empty(2, np.uint32)
        l.ocpu_split_index(cpu._ptr(np.ascontiguousarray(k)), 1000, 37, cpu._ptr(got))
        np.testing.assert_array_equal(got, P.split(k, 1000)[37])


def _funnel_start(C, D, seed):
    rng = np.random.default_rng(seed)
    v = 0.3 * rng.standard_normal((C, 1))
    q = np.concatenate([np.exp(0.5 * v) * 0.5 * rng.standard_normal((C, D - 1)), v], 1).astype(np.float32)
    keys = rng.integers(0, 2 ** 32, size=(C, 2), dtype=np.uint64).astype(np.uint32)
    return q, keys


@pytest.mark.parametrize("D,L,eps,variant", [(20, 8, 0.001016, "omega"), (20, 4, 0.3, "omega_fixed"), (5, 3, 0.1, "omegatilde"),
                                               (2, 8, 0.2, "omega_fixed")])
def test_lmcmonge_vs_numpy_oracle(D, L, eps, variant):
    C = 64
    q, keys = _funnel_start(C, D, D)
    tgt = T.NealFunnel(D)
    onew, oi = S.lmcmonge_step(keys, S.lmcmonge_init(q, tgt), tgt, eps, np.ones(D, np.float32), L, half_step=variant)
    smp = cpu.CpuSampler("lmcmonge", D, eps, L, half_step=variant, inverse_mass_matrix=np.ones(D, np.float32))
    st = smp.init(q)
    np.testing.assert_allclose(st[1], tgt.logp(q), rtol=1e-6, atol=1e-5)
    info = smp.step(keys, st)
    np.testing.assert_array_equal(info["accept_uniform"], oi.extra["u"])
    np.testing.assert_allclose(info["draw"], oi.momentum, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(info["proposal_position"], oi.proposal["position"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(info["energy"], oi.energy, rtol=1e-5, atol=2e-4)
    np.testing.assert_allclose(info["acceptance_rate"], oi.acceptance_rate, atol=5e-4)
    clear = np.abs(oi.extra["u"] - oi.acceptance_rate) > 2e-3
    np.testing.assert_array_equal(info["is_accepted"][clear].astype(bool), oi.is_accepted[clear])
    same = info["is_accepted"].astype(bool) == oi.is_accepted
    np.testing.assert_allclose(st[0][same], onew.position[same], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(st[2][same], onew.logdensity_grad[same], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("D,L,eps", [(2, 8, 0.1), (20, 4, 0.05), (100, 8, 0.05)])
def test_lmc_vs_numpy_oracle(D, L, eps):
    """closed (arrow-matrix) forms of the funnel's pull-back metric against the oracle's dense LU path"""
    C = 16
    q, keys = _funnel_start(C, D, D + 1)
    tgt = T.NealFunnel(D)
    onew, oi = S.lmc_step(keys, S.lmc_init(q, tgt), tgt, eps, L)
    smp = cpu.CpuSampler("lmc", D, eps, L)
    st = smp.init(q)
    info = smp.step(keys, st)
    tol = 1e-4 if D <= 20 else 2e-3  # the dense float32 LAPACK path is itself only ~D eps accurate
    np.testing.assert_allclose(info["draw"], oi.momentum, rtol=tol, atol=tol)
    np.testing.assert_allclose(info["proposal_position"], oi.proposal["position"], rtol=tol, atol=tol)
    np.testing.assert_allclose(info["acceptance_rate"], oi.acceptance_rate, atol=20 * tol)
    clear = np.abs(oi.extra["u"] - oi.acceptance_rate) > 40 * tol
    np.testing.assert_array_equal(info["is_accepted"][clear].astype(bool), oi.is_accepted[clear])


@pytest.mark.parametrize("N,D,C,L,eps", [(40, 3, 9, 1, 0.1), (200, 8, 12, 2, 0.1), (1000, 25, 6, 2, 0.1)])
def test_rmhmc_logreg_vs_numpy_oracle(N, D, C, L, eps):
    X, y = T.make_logreg_data(N, D, seed=1)
    tgt = T.LogisticRegression(X, y, 0.01)
    rng = np.random.default_rng(D)
    q = (0.1 * rng.standard_normal((C, D))).astype(np.float32)
    keys = rng.integers(0, 2 ** 32, size=(C, 2), dtype=np.uint64).astype(np.uint32)
    onew, oi = S.rmhmc_step(keys, S.rmhmc_init(q, tgt), tgt, eps, L)
    smp = cpu.CpuSampler("rmhmc", D, eps, L, X=X, y=y, prior_precision=0.01)
    st = smp.init(q)
    np.testing.assert_allclose(st[1], tgt.logp(q), rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(st[2], tgt.grad(q), rtol=1e-4, atol=1e-3)
    info = smp.step(keys, st)
    ok = oi.extra["fp_iters"] < 50 * L
    scale = float(np.abs(oi.momentum).max())
    np.testing.assert_allclose(info["draw"], oi.momentum, rtol=1e-5, atol=3e-5 * scale)
    np.testing.assert_allclose(info["proposal_position"][ok], oi.proposal["position"][ok], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(info["energy"][ok], oi.energy[ok], rtol=1e-5, atol=3e-3)
    np.testing.assert_allclose(info["acceptance_rate"][ok], oi.acceptance_rate[ok], atol=5e-3)


def test_run_equals_stepwise_and_is_partition_invariant():
    D, C, Tn = 20, 40, 3
    q, _ = _funnel_start(C, D, 3)
    root = P.key(5)
    smp = cpu.CpuSampler("lmcmonge", D, 0.002, 4, inverse_mass_matrix=np.ones(D, np.float32))
    a = smp.init(q)
    for t in range(Tn):
        smp.step(S.chain_keys(root, Tn, t, C), a, want_info=False)
    b = smp.init(q)
    acc = smp.run(root, b, Tn)
    assert 0.0 <= acc <= 1.0
    np.testing.assert_array_equal(a[0], b[0])
    lo = [x[:16].copy() if x is not None else None for x in smp.init(q)]
    hi = [x[16:].copy() if x is not None else None for x in smp.init(q)]
    smp.run(root, lo, Tn, total_chains=C)
    smp.run(root, hi, Tn, chain_offset=16, total_chains=C)
    np.testing.assert_array_equal(np.concatenate([lo[0], hi[0]]), b[0])
