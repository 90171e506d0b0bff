"""The NEW built-in targets (Gaussian, banana) through all three samplers,
against the dense oracle."""
import numpy as np
import pytest

from oracle import samplers as S
from oracle import targets as T

pytestmark = pytest.mark.gpu


def _t(x, dev):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def _close(got, want, rtol=1e-5, atol=1e-6, what=""):
    got = got.cpu().numpy() if hasattr(got, "cpu") else np.asarray(got)
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol, err_msg=what)


def _keys(C, seed):
    return np.random.default_rng(seed).integers(0, 2 ** 32, size=(C, 2), dtype=np.uint64).astype(np.uint32)


def _check(info, oinfo, new, onew, cuda, rtol, vel_field="velocity"):
    with np.errstate(invalid="ignore"):
        ok = np.isfinite(oinfo.proposal["weight"]) & (np.abs(oinfo.proposal["weight"]) < 50)  # well-conditioned chains
    assert ok.mean() > 0.75
    okt = _t(ok, cuda)
    ps = info.proposal.state
    _close(ps.position[okt], oinfo.proposal["position"][ok], rtol, 1e-5, "position")
    _close(ps.logdensity[okt], oinfo.proposal["logdensity"][ok], rtol, 1e-4, "logdensity")
    _close(info.energy[okt], oinfo.energy[ok], rtol, 2e-4, "energy")
    _close(info.acceptance_rate[okt], oinfo.acceptance_rate[ok], 10 * rtol, 1e-3, "acceptance")
    got = info.is_accepted.cpu().numpy()
    clear = (np.abs(oinfo.extra["u"] - oinfo.acceptance_rate) > 4e-3) & ok
    np.testing.assert_array_equal(got[clear], oinfo.is_accepted[clear])
    same = (got == oinfo.is_accepted) & ok
    _close(new.position[_t(same, cuda)], onew.position[same], rtol, 1e-5)


@pytest.mark.parametrize("D,lpc", [(3, 1), (7, 2), (20, 4), (50, 8)])
def test_gaussian_all_samplers(cuda, D, lpc):
    import geomjax_b200 as g
    C, L, eps = 64, 4, 0.15
    rng = np.random.default_rng(D)
    mean = rng.standard_normal(D).astype(np.float32)
    prec = (0.5 + 2 * rng.random(D)).astype(np.float32)
    q = (mean + rng.standard_normal((C, D)) / np.sqrt(prec)).astype(np.float32)
    keys = _keys(C, D)
    tgt = T.Gaussian(mean, prec)
    target = g.gaussian(_t(mean, cuda), _t(prec, cuda))
    # lmc with the target's (diagonal) metric
    alg = g.lmc(target, eps, target, L, lanes_per_chain=lpc)
    new, info = alg.step(_t(keys, cuda), alg.init(_t(q, cuda)))
    onew, oinfo = S.lmc_step(keys, S.lmc_init(q, tgt), tgt, eps, L)
    _close(info.velocity, oinfo.momentum, 1e-5, 1e-6, "lmc draw")
    _check(info, oinfo, new, onew, cuda, 5e-5)
    assert float(new.volume_adjustment.abs().max()) < 1e-5  # constant metric: no volume change
    # rmhmc with the diagonal metric: a constant metric makes the implicit midpoint converge fast
    alg = g.rmhmc(target, eps, target, L, lanes_per_chain=lpc)
    new, info = alg.step(_t(keys, cuda), alg.init(_t(q, cuda)))
    onew, oinfo = S.rmhmc_step(keys, S.rmhmc_init(q, tgt), tgt, eps, L)
    _close(info.momentum, oinfo.momentum, 1e-5, 1e-6, "rmhmc draw")
    _check(info, oinfo, new, onew, cuda, 5e-5)
    # lmcmonge
    im = (0.5 + rng.random(D)).astype(np.float32)
    em = 0.1 / D  # the as-written omega half step is only stable for eps ~ 1/D
    alg = g.lmcmonge(target, em, _t(im, cuda), L, alpha2=0.01, lanes_per_chain=lpc)
    new, info = alg.step(_t(keys, cuda), alg.init(_t(q, cuda)))
    onew, oinfo = S.lmcmonge_step(keys, S.lmcmonge_init(q, tgt), tgt, em, im, L, alpha2=0.01)
    _close(info.velocity, oinfo.momentum, 1e-5, 2e-6, "monge draw")
    _check(info, oinfo, new, onew, cuda, 1e-4)


def test_banana_all_samplers(cuda):
    import geomjax_b200 as g
    C, L, eps = 128, 5, 0.2
    rng = np.random.default_rng(0)
    q = np.stack([3 * rng.standard_normal(C), rng.standard_normal(C)], 1).astype(np.float32)
    keys = _keys(C, 1)
    tgt = T.Banana()
    target = g.banana()
    alg = g.lmc(target, eps, target, L)
    st = alg.init(_t(q, cuda))
    ost = S.lmc_init(q, tgt)
    _close(st.logdensity, ost.logdensity, 1e-5, 1e-5)
    _close(st.logdensity_grad, ost.logdensity_grad, 1e-5, 1e-5)
    new, info = alg.step(_t(keys, cuda), st)
    onew, oinfo = S.lmc_step(keys, ost, tgt, eps, L)
    _check(info, oinfo, new, onew, cuda, 5e-5)
    alg = g.rmhmc(target, eps, target, L)
    new, info = alg.step(_t(keys, cuda), alg.init(_t(q, cuda)))
    onew, oinfo = S.rmhmc_step(keys, S.rmhmc_init(q, tgt), tgt, eps, L)
    _check(info, oinfo, new, onew, cuda, 5e-5)
    alg = g.lmcmonge(target, 0.1, _t(np.ones(2, np.float32), cuda), L, alpha2=0.05)
    new, info = alg.step(_t(keys, cuda), alg.init(_t(q, cuda)))
    onew, oinfo = S.lmcmonge_step(keys, S.lmcmonge_init(q, tgt), tgt, 0.1, np.ones(2, np.float32), L, alpha2=0.05)
    _check(info, oinfo, new, onew, cuda, 1e-4)


def test_gaussian_sampling_moments(cuda):
    """Statistical check (BASELINE.json: posterior moments within MCSE, R-hat < 1.01) on a target with
    known moments: 4096 chains x 300 lmc transitions."""
    import torch
    import geomjax_b200 as g
    D = 5
    mean = torch.arange(D, dtype=torch.float32, device=cuda)
    prec = torch.tensor([0.25, 1.0, 4.0, 1.0, 2.0], device=cuda)
    target = g.gaussian(mean, prec)
    alg = g.lmc(target, 0.3, target, 5)
    st = alg.init(mean.repeat(4096, 1).contiguous())
    st, samples, acc = g.run_fused(alg.step, g.random.PRNGKey(7), st, 300, return_samples=True, return_accept=True)
    x = samples[50:]
    rhat = g.rhat(x, chain_axis=1, sample_axis=0)
    ess = g.ess(x, chain_axis=1, sample_axis=0)
    assert float(rhat.max()) < 1.01
    m = x.mean(dim=(0, 1)).cpu()
    v = x.var(dim=(0, 1)).cpu()
    mcse = (1.0 / prec.cpu() / ess).sqrt()
    assert bool(((m - mean.cpu()).abs() < 5 * mcse).all()), (m, mcse)
    np.testing.assert_allclose(v.numpy(), (1.0 / prec).cpu().numpy(), rtol=0.03)
    assert float(acc.mean()) > 0.8
