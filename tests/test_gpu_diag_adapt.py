"""GPU diagnostics (rhat / ess) and warm-up drivers against the oracle."""
import numpy as np
import pytest

from oracle import adaptation as OA
from oracle import diagnostics as OD
from oracle import prng as P
from oracle import samplers as S
from oracle import targets as T
from tests.test_cabi_and_host import _ar1

pytestmark = pytest.mark.gpu


def _t(x, dev):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


@pytest.mark.parametrize("T_,Cn,D,rho", [(200, 8, 3, 0.7), (1000, 64, 5, 0.3), (333, 5, 1, 0.9), (64, 2, 20, 0.0)])
def test_rhat_ess_vs_oracle(cuda, T_, Cn, D, rho):
    import geomjax_b200 as g
    x = _ar1(T_, Cn, D, rho, seed=T_).astype(np.float32)
    xt = _t(x, cuda)
    np.testing.assert_allclose(g.rhat(xt, chain_axis=1, sample_axis=0).numpy().reshape(-1),
                               np.atleast_1d(OD.potential_scale_reduction(x.astype(np.float64), 1, 0)), rtol=1e-5)
    np.testing.assert_allclose(g.ess(xt, chain_axis=1, sample_axis=0).numpy().reshape(-1),
                               np.atleast_1d(OD.effective_sample_size(x.astype(np.float64), 1, 0)), rtol=2e-4)
    # default axes of the reference signature: chain_axis=0, sample_axis=1
    xc = _t(np.ascontiguousarray(x.transpose(1, 0, 2)), cuda)
    np.testing.assert_allclose(g.ess(xc).numpy().reshape(-1),
                               np.atleast_1d(OD.effective_sample_size(x.astype(np.float64), 1, 0)), rtol=2e-4)
    with pytest.raises(AssertionError):
        g.rhat(xt[:, :1], chain_axis=1, sample_axis=0)  # needs two or more chains


def test_dual_averaging_kernels_vs_oracle(cuda):
    import torch
    import geomjax_b200 as g
    init, update, final = g.dual_averaging()
    rng = np.random.default_rng(0)
    eps0 = (0.05 + rng.random(33)).astype(np.float32)
    st = init(_t(eps0, cuda))
    ost = OA.da_init(eps0)
    for _ in range(40):
        grad = (0.8 - rng.random(33)).astype(np.float32)
        st = update(st, _t(grad, cuda))
        ost = OA.da_update(ost, grad)
    np.testing.assert_allclose(st.log_x.cpu().numpy(), ost["log_x"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(final(st).cpu().numpy(), np.exp(ost["log_x_avg"]), rtol=2e-5)


def test_step_size_adaptation_vs_oracle(cuda):
    """adaptation/step_size_adaptation.py:143-201, one independent adaptation per chain."""
    import geomjax_b200 as g
    D, C, W = 6, 12, 12
    tgt = T.NealFunnel(D)
    rng = np.random.default_rng(2)
    q = (0.3 * rng.standard_normal((C, D))).astype(np.float32)
    keys = rng.integers(0, 2 ** 32, size=(C, 2), dtype=np.uint64).astype(np.uint32)
    target = g.neal_funnel(D)
    res, info = g.step_size_adaptation(g.lmc, target, initial_step_size=0.5, metric_fn=target,
                                       num_integration_steps=3).run(_t(keys, cuda), _t(q, cuda), W)
    # oracle: chain by chain (the oracle kernels take a scalar step size)
    eps_final = np.empty(C, np.float32)
    acc0 = np.empty(C, np.float32)
    for c in range(C):
        kc = P.split(keys[c], W)
        st = S.lmc_init(q[c:c + 1], tgt)
        da = OA.da_init(np.array([0.5], np.float32))
        eps = np.float32(0.5)
        for t in range(W):
            st, oi = S.lmc_step(kc[t][None], st, tgt, float(eps), 3)
            if t == 0:
                acc0[c] = oi.acceptance_rate[0]
            da = OA.da_update(da, np.float32(0.8) - oi.acceptance_rate)
            eps = np.exp(da["log_x"])[0]
        eps_final[c] = max(np.exp(da["log_x_avg"])[0], 1e-3)
    np.testing.assert_allclose(info["acceptance_rate"][0].cpu().numpy(), acc0, rtol=1e-4, atol=1e-4)
    got = res.parameters["step_size"].cpu().numpy()
    # trajectories are chaotic in the accept decisions; the adapted step sizes still agree closely
    assert np.median(np.abs(got - eps_final) / eps_final) < 1e-3
    assert res.parameters["num_integration_steps"] == 3 and res.state.position.shape == (C, D)


def test_pooled_step_size_adaptation_reaches_target(cuda):
    import torch
    import geomjax_b200 as g
    D, C = 10, 4096
    target = g.neal_funnel(D)
    res, info = g.step_size_adaptation(g.lmc, target, initial_step_size=0.1, metric_fn=target,
                                       num_integration_steps=4, pooled=True).run(
        g.random.PRNGKey(1), torch.ones((C, D), device=cuda), 150)
    eps = res.parameters["step_size"]
    assert float(eps.std()) == 0.0 and 1e-3 < float(eps[0]) < 2.0
    tail = float(info["acceptance_rate"][-40:].mean())
    assert 0.7 < tail < 0.9, tail


def test_window_adaptation_lmcmonge(cuda):
    import torch
    import geomjax_b200 as g
    D, C, W = 5, 256, 200
    target = g.neal_funnel(D)
    wa = g.window_adaptation(g.lmcmonge, target, initial_step_size=0.05, num_integration_steps=4, alpha2=1e-3)
    res, info = wa.run(g.random.PRNGKey(0), 0.1 * torch.ones((C, D), device=cuda), W)
    im = res.parameters["inverse_mass_matrix"]
    assert im.shape == (C, D) and bool((im > 0).all()) and not bool((im == 1).all())
    assert res.parameters["step_size"].shape == (C,) and bool(torch.isfinite(res.parameters["step_size"]).all())
    # the adapted parameters plug straight back into the sampler (per-chain step size and mass)
    alg = g.lmcmonge(target, res.parameters["step_size"], im, 4, alpha2=1e-3)
    st, info2 = alg.step(g.random.chain_keys(g.random.PRNGKey(5), 0, 1, C), res.state)
    assert 0.3 < float(info2.acceptance_rate.mean()) <= 1.0
    with pytest.raises(NotImplementedError):
        g.window_adaptation(g.rmhmc, target, num_integration_steps=4)


@pytest.mark.parametrize("T_,Cn,D,rho,block", [(300, 40, 3, 0.6, 37), (1000, 130, 5, 0.9, 256), (64, 7, 2, 0.0, 64), (50, 33, 1, 0.5, 8)])
def test_streaming_diagnostics_equal_batch(cuda, T_, Cn, D, rho, block):
    """N2: rhat / ess from blocks of samples (no (T, C, D) tensor) == the batch functions on the whole tensor."""
    import torch
    import geomjax_b200 as g
    rng = np.random.default_rng(T_ + Cn)
    x = np.empty((T_, Cn, D), np.float32)
    x[0] = rng.standard_normal((Cn, D))
    for t in range(1, T_):
        x[t] = rho * x[t - 1] + np.sqrt(1 - rho * rho) * rng.standard_normal((Cn, D))
    x += np.array([0.0, 100.0, -3.0, 7.0, 1e3][:D], np.float32)  # means far from zero: the shift matters
    xt = torch.from_numpy(x).to(cuda)
    lags = 256 if rho > 0.8 else (64 if T_ >= 64 else 48)  # enough lags for Geyer's truncation (no doubling in streaming)
    sd = g.StreamingDiagnostics(Cn, D, max_lags=lags, device=cuda)
    for t0 in range(0, T_, block):
        sd.update(xt[t0:t0 + block])
    np.testing.assert_allclose(sd.rhat().numpy(), g.rhat(xt, chain_axis=1, sample_axis=0).numpy(), rtol=1e-5)
    want = g.ess(xt, chain_axis=1, sample_axis=0, initial_lags=lags)
    got = sd.ess(allow_truncated=True)
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=2e-4)
    np.testing.assert_allclose(got.numpy(), OD.effective_sample_size(x.astype(np.float64), chain_axis=1, sample_axis=0),
                               rtol=5e-3)


def test_sample_streaming_equals_fused_run(cuda):
    """the blocked driver reproduces one fused launch bit for bit and its diagnostics match the batch ones"""
    import torch
    import geomjax_b200 as g
    D, Cn, Tn = 5, 256, 200
    mean = torch.arange(D, dtype=torch.float32, device=cuda)
    target = g.gaussian(mean, torch.tensor([0.25, 1.0, 4.0, 1.0, 2.0], device=cuda))
    alg = g.lmc(target, 0.3, target, 5)
    root = g.random.PRNGKey(7)
    st0 = alg.init(mean.repeat(Cn, 1).contiguous())
    st_a, samples, acc_a = g.run_fused(alg.step, root, st0, Tn, return_samples=True, return_accept=True)
    seen = []
    st_b, diag, acc_b = g.sample_streaming(alg.step, root, st0, Tn, block=48, on_block=lambda f, blk: seen.append(blk.clone()))
    assert bool((st_a.position == st_b.position).all()) and bool((torch.cat(seen) == samples).all())
    np.testing.assert_allclose(acc_b.cpu().numpy(), acc_a.mean(0).cpu().numpy(), rtol=1e-5)
    np.testing.assert_allclose(diag.rhat().numpy(), g.rhat(samples, chain_axis=1, sample_axis=0).numpy(), rtol=1e-5)
    np.testing.assert_allclose(diag.ess(allow_truncated=True).numpy(), g.ess(samples, chain_axis=1, sample_axis=0).numpy(), rtol=1e-3)
