"""N4 of SURVEY 8(f): the dynamic (jittered-L) kernels and the ChEES adaptation.
rmhmc/rmhmc.py:179-244, lmcmc/lmc.py:185-252, lmcmonge/lmc.py:240-309, adaptation/chees_adaptation_riemanian.py:56-466."""
import numpy as np
import pytest

from oracle import chees as OC, prng as P, samplers as S, targets as T

pytestmark = pytest.mark.gpu


def _t(x, dev):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def _funnel_start(C, D, seed):
    rng = np.random.default_rng(seed)
    v = 0.3 * rng.standard_normal((C, 1))
    q = np.concatenate([np.exp(0.5 * v) * 0.5 * rng.standard_normal((C, D - 1)), v], 1).astype(np.float32)
    keys = rng.integers(0, 2 ** 32, size=(C, 2), dtype=np.uint64).astype(np.uint32)
    return q, keys


def test_randint_matches_oracle(cuda):
    import geomjax_b200 as g
    keys = P.split(P.key(3), 4096)
    for lo, hi in [(1, 10), (0, 7), (5, 6), (1, 1000)]:
        got = g.random.randint(_t(keys, cuda), lo, hi).cpu().numpy()
        np.testing.assert_array_equal(got, P.randint(keys, lo, hi))
        assert got.min() >= lo and got.max() < hi
    assert len(np.unique(g.random.randint(_t(keys, cuda), 1, 10).cpu().numpy())) == 9
    assert (g.random.PRNGKey(-1) == np.array([0, 0xFFFFFFFF], np.uint32)).all()


@pytest.mark.parametrize("sampler,D", [("lmcmonge", 20), ("lmcmonge", 5), ("lmc", 20), ("lmc", 100), ("rmhmc", 20), ("rmhmc", 2),
                                       ("rmhmc_logreg", 8)])
def test_per_chain_steps_equal_static_launches(cuda, sampler, D):
    """A launch with a per-chain number of integration steps == for every chain, the static launch with that chain's
    count, BIT FOR BIT (a finished chain is masked / steps by zero while its warp runs on)."""
    import torch
    import geomjax_b200 as g
    C = 96
    rng = np.random.default_rng(D)
    steps = rng.integers(0 if sampler != "rmhmc_logreg" else 1, 6, size=C).astype(np.int32)
    if sampler == "rmhmc_logreg":
        X, y = T.make_logreg_data(200, D, seed=1)
        target = g.logistic_regression(_t(X, cuda), _t(y, cuda), 0.01)
        q = (0.1 * rng.standard_normal((C, D))).astype(np.float32)
        keys = rng.integers(0, 2 ** 32, size=(C, 2), dtype=np.uint64).astype(np.uint32)
        make = lambda L: g.rmhmc(target, 0.1, target, L)
    else:
        target = g.neal_funnel(D)
        q, keys = _funnel_start(C, D, D + 1)
        if sampler == "lmcmonge":
            make = lambda L: g.lmcmonge(target, 0.05, torch.ones(D, device=cuda), L, integrator=g.integrators.half_step_omega_fixed)
        elif sampler == "lmc":
            make = lambda L: g.lmc(target, 0.05, target, L)
        else:
            make = lambda L: g.rmhmc(target, 0.05, target, L)
    kt = _t(keys, cuda)
    dyn = make(_t(steps, cuda))
    st0 = dyn.init(_t(q, cuda))
    new, info = dyn.step(kt, st0)
    for L in np.unique(steps):
        sel = _t(steps == L, cuda)
        ref_new, ref_info = make(int(L)).step(kt, st0)
        if L == 0 and sampler != "rmhmc":
            # zero steps: the masked half-steps still refresh log-density / gradient from the (unchanged) position,
            # where the static launch reuses the input state's values: equal up to that round-off, not bit for bit
            np.testing.assert_allclose(info.proposal.state.position[sel].cpu().numpy(),
                                       ref_info.proposal.state.position[sel].cpu().numpy(), rtol=0, atol=0)
            np.testing.assert_allclose(info.energy[sel].cpu().numpy(), ref_info.energy[sel].cpu().numpy(), rtol=2e-6, atol=1e-4)
            assert float(info.acceptance_rate[sel].min()) > 0.999
            continue
        assert bool((new.position[sel] == ref_new.position[sel]).all()), L
        assert bool((info.proposal.state.position[sel] == ref_info.proposal.state.position[sel]).all()), L
        assert bool((info.energy[sel] == ref_info.energy[sel]).all() or torch.isnan(info.energy[sel]).any()), L
        assert bool((info.is_accepted[sel] == ref_info.is_accepted[sel]).all()), L


def test_dynamic_lmc_vs_oracle(cuda):
    """dynamic_lmc with the reference defaults: L = randint(random_generator_arg, 1, 10) per chain, argument advanced
    by split(key)[1]; against the oracle's static lmc_step run per group of chains with equal L."""
    import geomjax_b200 as g
    D, C, eps = 5, 64, 0.1
    q, keys = _funnel_start(C, D, 11)
    target = g.neal_funnel(D)
    alg = g.dynamic_lmc(target, eps, target)
    arg = P.split(P.key(9), C)
    st = alg.init(_t(q, cuda), _t(arg, cuda))
    new, info = alg.step(_t(keys, cuda), st)
    Ls = P.randint(arg, 1, 10)
    np.testing.assert_array_equal(new.random_generator_arg.cpu().numpy(), P.split(arg, 2)[:, 1])
    tgt = T.NealFunnel(D)
    got = info.proposal.state.position.cpu().numpy()
    for L in np.unique(Ls):
        sel = Ls == L
        _, oi = S.lmc_step(keys[sel], S.lmc_init(q[sel], tgt), tgt, eps, int(L))
        np.testing.assert_allclose(got[sel], oi.proposal["position"], rtol=2e-4, atol=2e-5)
    assert info.num_integration_steps == int(Ls.max())
    # a single (2,) key is split into one generator argument per chain; integer counters are broadcast
    st2 = alg.init(_t(q, cuda), P.key(1))
    assert st2.random_generator_arg.shape == (C, 2)
    st3 = g.dynamic_rmhmc.init(_t(q, cuda), target, 0)
    assert st3.random_generator_arg.shape == (C,)


@pytest.mark.parametrize("dynamics", ["lmc", "rmhmc"])
def test_chees_adaptation_vs_oracle(cuda, dynamics):
    """ChEES warm-up (Halton jitter): the step size / trajectory length recurrences driven by the pooled cross-chain
    statistics follow the oracle's restatement of chees_adaptation_riemanian.py over the first transitions (later the
    float32 chains decorrelate and only the statistics agree)."""
    import torch
    import geomjax_b200 as g
    D, C, n = 4, 512, 40
    mean = np.arange(D, dtype=np.float32)
    prec = np.array([0.25, 1.0, 4.0, 1.0], np.float32)
    tgt = T.Gaussian(mean, prec)
    rng = np.random.default_rng(0)
    q0 = (mean + rng.standard_normal((C, D)) / np.sqrt(prec)).astype(np.float32)
    with np.errstate(all="ignore"):
        _, ost, oh = OC.run(P.key(4), q0, tgt, 0.2, 0.05, n, dynamics=dynamics)
    target = g.gaussian(_t(mean, cuda), _t(prec, cuda))
    warm = g.chees_adaptation_riemanian(target, target, C, dynamics=dynamics)
    (last, params), hist = warm.run(g.random.PRNGKey(4), _t(q0, cuda), 0.2, g.chees.adam(0.05), n)
    assert hist["num_integration_steps"][:6] == oh["num_integration_steps"][:6]
    np.testing.assert_allclose(hist["step_size"][:6], oh["step_size"][:6], rtol=2e-3)
    np.testing.assert_allclose(hist["trajectory_length"][:6], oh["trajectory_length"][:6], rtol=2e-3)
    np.testing.assert_allclose(hist["step_size"], oh["step_size"], rtol=0.1)
    assert params["step_size"] > 0 and params["integration_steps_fn"](0) >= 1
    assert last.random_generator_arg.shape == (C,) and int(last.random_generator_arg[0]) == n
    # the tuned parameters plug into the dynamic kernel (docstring example of chees_adaptation_riemanian.py:313-323)
    cls = g.dynamic_lmc if dynamics == "lmc" else g.dynamic_rmhmc
    alg = cls(target, params["step_size"], params["metric_fn"], next_random_arg_fn=params["next_random_arg_fn"],
              integration_steps_fn=params["integration_steps_fn"])
    new, info = alg.step(g.random.chain_keys(g.random.PRNGKey(8), 0, 1, C), last)
    assert bool(torch.isfinite(new.position).all())
    if dynamics == "lmc":  # (rmhmc: after 40 warm-up transitions the moving-average step size is still dominated by
        assert float(info.acceptance_rate.mean()) > 0.05  # dual averaging's first large trials; the fixed point diverges)
    assert int(new.random_generator_arg[0]) == n + 1
