"""Fused CUDA lmc (Lan integrator, arrow-matrix closed forms) and rmhmc (implicit midpoint)
against the dense NumPy oracle (which follows the reference line by line)."""
import numpy as np
import pytest

from oracle import prng as P
from oracle import samplers as S
from oracle import targets as T

pytestmark = pytest.mark.gpu


def _t(x, dev):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def _close(got, want, rtol=1e-5, atol=1e-6, what=""):
    """Element-wise rtol plus a NORM-WISE floor: the oracle (like the reference) pushes every vector
    through dense float32 LAPACK (Cholesky / LU / solve), whose forward error is ~ D * eps_f32
    relative to the vector norm, not to each element; the closed forms on the GPU are more
    accurate than that (see test_lmc_closer_to_float64_than_the_float32_oracle)."""
    got = got.cpu().numpy() if hasattr(got, "cpu") else np.asarray(got)
    want = np.asarray(want)
    floor = 3e-5 * float(np.abs(want[np.isfinite(want)]).max()) if want.ndim == 2 and want.size and want.shape[1] >= 33 else 0.0
    np.testing.assert_allclose(got, want, rtol=rtol, atol=max(atol, floor), err_msg=what)


def _setup(D, C, seed=0, scale=0.5):
    rng = np.random.default_rng(seed)
    v = 0.7 * scale * rng.standard_normal((C, 1))
    q = np.concatenate([np.exp(0.5 * v) * rng.standard_normal((C, D - 1)) * scale, v], axis=1).astype(np.float32)
    keys = rng.integers(0, 2 ** 32, size=(C, 2), dtype=np.uint64).astype(np.uint32)
    return q, keys


CASES = [(2, 1, 128), (2, 2, 128), (5, 4, 128), (20, 1, 96), (20, 2, 96), (20, 4, 96), (33, 32, 32), (100, 4, 12), (100, 8, 12)]


@pytest.mark.parametrize("L,rtol", [(1, 1e-5), (8, 1e-4)])
@pytest.mark.parametrize("D,lpc,C", CASES)
def test_lmc_vs_oracle(cuda, D, lpc, C, L, rtol):
    import geomjax_b200 as g
    q, keys = _setup(D, C, seed=D + 1)
    tgt = T.NealFunnel(D)
    eps = 0.2 / np.sqrt(D) / L ** 0.5
    ost = S.lmc_init(q, tgt)
    onew, oinfo = S.lmc_step(keys, ost, tgt, eps, L)
    target = g.neal_funnel(D)
    alg = g.lmc(target, eps, target, L, lanes_per_chain=lpc)
    st = alg.init(_t(q, cuda))
    new, info = alg.step(_t(keys, cuda), st)
    with np.errstate(invalid="ignore"):
        tame = np.isfinite(oinfo.proposal["weight"]) & (np.abs(oinfo.proposal["weight"]) < 50)
    assert tame.mean() > 0.9 and oinfo.is_accepted.mean() > 0.3
    tm = _t(tame, cuda)
    ps = info.proposal.state
    _close(info.velocity, oinfo.momentum, rtol=1e-5, atol=1e-6, what="velocity draw")
    _close(ps.position[tm], oinfo.proposal["position"][tame], rtol=rtol, atol=1e-6, what="position")
    _close(ps.velocity[tm], oinfo.proposal["velocity"][tame], rtol=rtol, atol=1e-5, what="velocity")
    _close(ps.momentum[tm], oinfo.proposal["momentum"][tame], rtol=rtol, atol=1e-4, what="momentum")
    _close(ps.logdensity[tm], oinfo.proposal["logdensity"][tame], rtol=rtol, atol=1e-4)
    # the oracle's volume term is a sum of 4 float32 LU log-dets per step (error ~ D * eps_f32 each)
    vol_atol = max(2e-5, 1e-6 * D * L)
    _close(ps.volume_adjustment[tm], oinfo.proposal["volume_adjustment"][tame], rtol=rtol, atol=vol_atol, what="volume")
    _close(info.energy[tm], oinfo.energy[tame], rtol=rtol, atol=2e-4, what="energy")
    _close(info.acceptance_rate[tm], oinfo.acceptance_rate[tame], rtol=10 * rtol, atol=1e-3)
    got_acc = info.is_accepted.cpu().numpy()
    dp = np.abs(info.acceptance_rate.cpu().numpy() - oinfo.acceptance_rate)[tame].max()
    clear = np.abs(oinfo.extra["u"] - oinfo.acceptance_rate) > 4 * dp + 1e-6  # 4x the observed |delta p_accept|
    assert clear.mean() > 0.95, dp
    np.testing.assert_array_equal(got_acc[clear], oinfo.is_accepted[clear])
    same = (got_acc == oinfo.is_accepted) & tame
    _close(new.position[_t(same, cuda)], onew.position[same], rtol=rtol)
    _close(new.volume_adjustment[_t(same, cuda)], onew.volume_adjustment[same], rtol=rtol, atol=vol_atol)


@pytest.mark.parametrize("L,rtol", [(1, 2e-5), (4, 1e-4)])
@pytest.mark.parametrize("D,lpc,C", CASES)
def test_rmhmc_vs_oracle(cuda, D, lpc, C, L, rtol):
    import geomjax_b200 as g
    q, keys = _setup(D, C, seed=D + 2)
    tgt = T.NealFunnel(D)
    eps = 0.1 / np.sqrt(D)
    ost = S.rmhmc_init(q, tgt)
    onew, oinfo = S.rmhmc_step(keys, ost, tgt, eps, L)
    target = g.neal_funnel(D)
    alg = g.rmhmc(target, eps, target, L, lanes_per_chain=lpc)
    st = alg.init(_t(q, cuda))
    _close(st.logdensity_grad, ost.logdensity_grad)
    new, info = alg.step(_t(keys, cuda), st)
    with np.errstate(invalid="ignore"):
        tame = np.isfinite(oinfo.proposal["weight"]) & (np.abs(oinfo.proposal["weight"]) < 50) \
            & (oinfo.extra["fp_iters"] < 60 * L)
    assert tame.mean() > 0.8 and oinfo.is_accepted.mean() > 0.3
    tm = _t(tame, cuda)
    ps = info.proposal.state
    _close(info.momentum, oinfo.momentum, rtol=1e-5, atol=1e-6, what="momentum draw")
    # fixed-point tolerance 1e-6 is absolute: both sides stop within ~1e-6 of the same fixed point
    _close(ps.position[tm], oinfo.proposal["position"][tame], rtol=rtol, atol=5e-6, what="position")
    _close(ps.momentum[tm], oinfo.proposal["momentum"][tame], rtol=rtol, atol=2e-5, what="momentum")
    _close(ps.velocity[tm], oinfo.proposal["velocity"][tame], rtol=rtol, atol=2e-5, what="velocity")
    _close(ps.logdensity[tm], oinfo.proposal["logdensity"][tame], rtol=rtol, atol=2e-4)
    _close(info.energy[tm], oinfo.energy[tame], rtol=rtol, atol=3e-4, what="energy")
    got_acc = info.is_accepted.cpu().numpy()
    dp = np.abs(info.acceptance_rate.cpu().numpy() - oinfo.acceptance_rate)[tame].max()
    clear = np.abs(oinfo.extra["u"] - oinfo.acceptance_rate) > 4 * dp + 1e-6  # 4x the observed |delta p_accept|
    assert clear[tame].mean() > 0.95, dp
    np.testing.assert_array_equal(got_acc[clear & tame], oinfo.is_accepted[clear & tame])


def test_lmc_closer_to_float64_than_the_float32_oracle(cuda):
    """D=100: |CUDA - f64 oracle| <= |f32 oracle - f64 oracle| (+ small slack) for one Lan step."""
    import geomjax_b200 as g
    D, C, eps = 100, 8, 0.02
    q, keys = _setup(D, C, seed=7)
    o32, i32 = S.lmc_step(keys, S.lmc_init(q, T.NealFunnel(D)), T.NealFunnel(D), eps, 1)
    t64 = T.NealFunnel(D, dtype=np.float64)
    o64, i64 = S.lmc_step(keys, S.lmc_init(q, t64), t64, eps, 1, z=i32.extra["z"], u=i32.extra["u"])
    target = g.neal_funnel(D)
    alg = g.lmc(target, eps, target, 1)
    new, info = alg.step(_t(keys, cuda), alg.init(_t(q, cuda)))
    for name, got in (("position", info.proposal.state.position), ("velocity", info.proposal.state.velocity)):
        e_gpu = np.abs(got.cpu().numpy() - i64.proposal[name]).max()
        e_o32 = np.abs(i32.proposal[name] - i64.proposal[name]).max()
        assert e_gpu <= 2.0 * e_o32 + 2e-6, (name, e_gpu, e_o32)
    e_gpu = np.abs(info.proposal.state.volume_adjustment.cpu().numpy() - i64.proposal["volume_adjustment"]).max()
    e_o32 = np.abs(i32.proposal["volume_adjustment"] - i64.proposal["volume_adjustment"]).max()
    assert e_gpu <= 2.0 * e_o32 + 2e-6, ("volume", e_gpu, e_o32)
    e_gpu = np.abs(info.energy.cpu().numpy() - i64.energy).max()
    e_o32 = np.abs(i32.energy - i64.energy).max()
    assert e_gpu <= 2.0 * e_o32 + 1e-4, (e_gpu, e_o32)


def test_reference_test_equivalences_on_gpu(cuda):
    """tests/test_samplers.py:21-57 on the CUDA path: G = I => rmhmc ~ lmc (~ hmc) at rtol 1e-4,
    and the SURVEY Appendix A.2 candidate values for key 42."""
    import geomjax_b200 as g
    target = g.neal_funnel(2)
    key = _t(P.key(42)[None], cuda)
    z = _t(np.zeros((1, 2), np.float32), cuda)
    s1, _ = g.rmhmc(target, 1e-2, "identity", 10).step(key, g.rmhmc.init(z, target))
    s3, _ = g.lmc(target, 1e-2, "identity", 10).step(key, g.lmc.init(z, target))
    np.testing.assert_allclose(s1.position.cpu().numpy(), s3.position.cpu().numpy(), rtol=1e-4)
    _close(s3.position, np.array([[0.06480624, -0.06973797]], np.float32), rtol=1e-5)
    _close(s1.position, np.array([[0.064804554, -0.06973776]], np.float32), rtol=2e-5)
    # funnel metric, same key: lmc and rmhmc follow the same dynamics
    s4, i4 = g.rmhmc(target, 1e-2, target, 10).step(key, g.rmhmc.init(z, target))
    s5, i5 = g.lmc(target, 1e-2, target, 10).step(key, g.lmc.init(z, target))
    _close(i4.proposal.state.position, np.array([[0.058600806, -0.20142354]], np.float32), rtol=2e-5)
    _close(i5.proposal.state.position, np.array([[0.058600716, -0.20142427]], np.float32), rtol=2e-5)
    _close(i5.proposal.state.volume_adjustment, np.array([-0.20142305], np.float32), rtol=1e-4)
    _close(i4.momentum, np.array([[0.6491706, -0.22417447]], np.float32), rtol=1e-6)
    _close(i5.velocity, np.array([[0.6491706, -2.01757]], np.float32), rtol=1e-6)


def test_shipped_example_first_transition(cuda):
    """examples/funnel/main.py as shipped (lmc, funnel metric, eps=0.1, L=8, ones(2), PRNGKey(0), 8 chains)."""
    import torch
    import geomjax_b200 as g
    target = g.neal_funnel(2)
    alg = g.lmc(target, 0.1, target.fisher_metric_fn, 8)
    st0 = alg.init(torch.ones((8, 2), device=cuda))
    root = g.random.PRNGKey(0)
    st, samples, acc = g.run_fused(alg.step, root, st0, 3, total=1000, return_samples=True)
    keys = g.random.chain_keys(root, 0, 1000, 8)
    st1, info = alg.step(keys, st0)
    assert bool((samples[0] == st1.position).all())
    _close(info.velocity[0], np.array([-0.022572383, -0.44815361], np.float32), rtol=2e-6)
    _close(info.proposal.state.position[0], np.array([0.61434513, 0.37447447], np.float32), rtol=2e-5)
    _close(info.proposal.state.volume_adjustment[0], np.float32(-0.62510157), rtol=1e-4)
    assert bool(info.is_accepted[0])
    tgt = T.NealFunnel(2)
    ost = S.lmc_init(np.ones((8, 2), np.float32), tgt)
    onew, oinfo = S.lmc_step(S.chain_keys(P.key(0), 1000, 0, 8), ost, tgt, 0.1, 8)
    _close(st1.position, onew.position, rtol=1e-4)
    np.testing.assert_array_equal(info.is_accepted.cpu().numpy(), oinfo.is_accepted)


@pytest.mark.parametrize("sampler,D", [("lmc", 100), ("lmc", 20), ("lmc", 2), ("rmhmc", 20), ("rmhmc", 2),
                                       ("lmcmonge", 20), ("lmcmonge", 2), ("lmcmonge", 100)])
def test_lean_fused_kernels_equal_stepwise(cuda, sampler, D):
    """The lean instantiations (no Info / overrides / adaptation, compile-time half-step: what a fused
    sampling launch runs) are bit-identical to T single steps through the full kernels."""
    import torch
    import geomjax_b200 as g
    import geomjax_b200.random as R
    C, T_ = 70, 4
    target = g.neal_funnel(D)
    if sampler == "lmc":
        alg = g.lmc(target, 0.3 / np.sqrt(D), target, 3)
    elif sampler == "rmhmc":
        alg = g.rmhmc(target, 0.3 / np.sqrt(D), target, 3)
    else:
        alg = g.lmcmonge(target, 0.2 / np.sqrt(D), torch.ones(D, device=cuda), 3,
                         integrator=g.integrators.half_step_omega_fixed)
    root = P.key(3)
    st0 = alg.init(torch.ones((C, D), device=cuda))
    st, pos = st0, []
    for t in range(T_):
        st, info = alg.step(R.chain_keys(root, t, T_, C), st)
        pos.append(st.position.clone())
    fst, samples, _ = g.run_fused(alg.step, root, st0, T_, return_samples=True)
    assert bool((samples == torch.stack(pos)).all())
    for a_, b_ in zip(fst, st):
        assert bool((a_ == b_).all())
    assert float((samples[-1] != samples[0]).float().mean()) > 0.5  # chains actually moved


@pytest.mark.parametrize("sampler", ["lmc", "rmhmc", "lmcmonge"])
@pytest.mark.parametrize("D,C", [(2, 64), (8, 33), (20, 40), (27, 9)])
def test_float64_single_step_rel_1e10(cuda, sampler, D, C):
    """jax_enable_x64 parity (north-star tolerance rel 1e-10): the float64 instantiations against the
    float64 oracle, one integrator step, same keys (a float64 launch widens the float32 draws)."""
    import torch
    import geomjax_b200 as g
    q32, keys = _setup(D, C, seed=3 * D + 1)
    q = q32.astype(np.float64)
    t32, t64 = T.NealFunnel(D), T.NealFunnel(D, dtype=np.float64)
    eps = 0.05 / np.sqrt(D)
    km, ka = S._draw_keys(keys, P.LEGACY)
    z = np.stack([P.normal(km[c], (D,)) for c in range(C)]).astype(np.float64)
    u = np.array([P.uniform(ka[c]) for c in range(C)], np.float64)
    target = g.neal_funnel(D)
    qd = _t(q, cuda)
    if sampler == "lmc":
        onew, oi = S.lmc_step(keys, S.lmc_init(q, t64), t64, eps, 1, z=z, u=u)
        alg = g.lmc(target, eps, target, 1)
    elif sampler == "rmhmc":
        onew, oi = S.rmhmc_step(keys, S.rmhmc_init(q, t64), t64, eps, 1, z=z, u=u)
        alg = g.rmhmc(target, eps, target, 1)
    else:
        onew, oi = S.lmcmonge_step(keys, S.lmcmonge_init(q, t64), t64, eps, np.ones(D), 1, z=z, u=u,
                                   half_step="omega_fixed")
        alg = g.lmcmonge(target, eps, torch.ones(D, device=cuda, dtype=torch.float64), 1,
                         integrator=g.integrators.half_step_omega_fixed)
    st = alg.init(qd)
    assert st.position.dtype == torch.float64 and st.logdensity.dtype == torch.float64
    new, info = alg.step(_t(keys, cuda), st)
    n = lambda t: t.cpu().numpy()
    ps = info.proposal.state
    draw = info.momentum if sampler == "rmhmc" else info.velocity
    np.testing.assert_allclose(n(draw), oi.momentum, rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(n(ps.position), oi.proposal["position"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(n(ps.logdensity), oi.proposal["logdensity"], rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(n(info.energy), oi.energy, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(n(info.acceptance_rate), oi.acceptance_rate, rtol=1e-9, atol=1e-10)
    np.testing.assert_array_equal(n(info.is_accepted), oi.is_accepted)
    np.testing.assert_allclose(n(new.position), onew.position, rtol=1e-10, atol=1e-12)
    if sampler != "rmhmc":
        np.testing.assert_allclose(n(ps.volume_adjustment), oi.proposal["volume_adjustment"], rtol=1e-9, atol=1e-11)
