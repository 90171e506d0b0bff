"""rmhmc on Bayesian logistic regression with the Fisher metric (dense per-chain metric, CTA per chain)
against the dense NumPy oracle (same reference semantics, analytic derivatives)."""
import numpy as np
import pytest

from oracle import samplers as S
from oracle import targets as T

pytestmark = pytest.mark.gpu


def _t(x, dev):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def _close(got, want, rtol, atol, what=""):
    got = got.cpu().numpy() if hasattr(got, "cpu") else np.asarray(got)
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol, err_msg=what)


_SHAPES = [(40, 3, 9, 1), (200, 8, 20, 2), (1000, 25, 6, 1), (1000, 25, 4, 3), (333, 17, 5, 2),
           (40, 3, 64, 1), (200, 8, 40, 2), (1000, 25, 70, 1), (1000, 25, 33, 2), (333, 17, 65, 2),
           # D > 32 (blocked 8x8 factorisation; rmhmc_logreg_big.cu for per_chain); last = c5's shape
           (300, 40, 5, 2), (50, 33, 4, 1), (4000, 30, 3, 1), (10000, 100, 3, 1)]
_PER_CHAIN = [(40, 3, 9, 1), (1000, 25, 6, 1), (333, 17, 65, 2), (300, 40, 5, 2), (4000, 30, 3, 1)]


@pytest.mark.parametrize("Nrows,D,C,L,path", [s + ("lockstep",) for s in _SHAPES] + [s + ("per_chain",) for s in _PER_CHAIN])
def test_rmhmc_logreg_vs_oracle(cuda, Nrows, D, C, L, path):
    """`path` = the product path (lock-step rolling batch on the tcgen05 GEMMs, rmhmc_lockstep.cu) and the
    CTA-per-chain FP32 kernels kept beside it (rmhmc_logreg.cu / rmhmc_logreg_big.cu)."""
    import geomjax_b200 as g
    X, y = T.make_logreg_data(Nrows, D, seed=1)
    tgt = T.LogisticRegression(X, y, 0.01)
    tgt.structured_dmetric = Nrows * D ** 3 > 1e9  # same contractions without the (C, D, D, D) tensor
    rng = np.random.default_rng(D)
    q = (0.1 * rng.standard_normal((C, D))).astype(np.float32)
    keys = rng.integers(0, 2 ** 32, size=(C, 2), dtype=np.uint64).astype(np.uint32)
    eps = 0.1 if D <= 40 else 0.05
    ost = S.rmhmc_init(q, tgt)
    onew, oinfo = S.rmhmc_step(keys, ost, tgt, eps, L)
    target = g.logistic_regression(_t(X, cuda), _t(y, cuda), 0.01)
    alg = g.rmhmc(target, eps, target, L, logreg_path=path)
    st = alg.init(_t(q, cuda))
    _close(st.logdensity, ost.logdensity, 1e-5, 1e-3, "init logdensity")
    _close(st.logdensity_grad, ost.logdensity_grad, 1e-4, 1e-3, "init grad")
    eng = alg.step.engine
    ks = g._native.KeySource()
    kt = _t(keys, cuda)
    ks.keys, ks.num_transitions = g._native.ptr(kt), 1
    out, inf = eng.launch(st, ks, want_info=True, extra_info=True)
    new, info = eng.make_state(out), eng.make_info(inf)
    fp = inf["fp_iters"].cpu().numpy()
    # chains whose float32 fixed point stalls at tol = 1e-6 (the noise level of |p| ~ sqrt(N) in float32) iterate until
    # they hit the tolerance by chance, so their iteration COUNT is round-off-dependent on both sides (measured: oracle
    # 107 iterations, GPU 13 for the same chain) and their end point is only defined to ~1e-4.  They are not dropped:
    # they must land on the same proposal to that looser tolerance, and every chain must respect max_iters.
    ok = oinfo.extra["fp_iters"] < 50 * L
    assert oinfo.is_accepted.mean() > 0.5 and ok.mean() > 0.7
    slow = ~ok
    assert (fp >= 0).all() and (fp <= 100 * L).all()
    if slow.any():
        _close(info.proposal.state.position[_t(slow, cuda)], oinfo.proposal["position"][slow], 2e-3, 2e-4, "stalled chains")
    if ok.sum() >= 16:  # the typical iteration count agrees (a single round-off-stalled chain would move a tiny mean)
        assert abs(np.median(fp[ok]) - np.median(oinfo.extra["fp_iters"][ok])) <= 0.25 * np.median(oinfo.extra["fp_iters"][ok]) + 1
    okt = _t(ok, cuda)
    scale = float(np.abs(oinfo.momentum).max())
    _close(info.momentum, oinfo.momentum, 1e-5, 3e-5 * scale, "momentum draw")
    ps = info.proposal.state
    _close(ps.position[okt], oinfo.proposal["position"][ok], 1e-4, 1e-5, "position")
    _close(ps.momentum[okt], oinfo.proposal["momentum"][ok], 1e-4, 1e-4 * scale, "momentum")
    _close(ps.velocity[okt], oinfo.proposal["velocity"][ok], 1e-4, 1e-5, "velocity")
    _close(ps.logdensity[okt], oinfo.proposal["logdensity"][ok], 1e-5, 2e-3, "logdensity")
    _close(info.energy[okt], oinfo.energy[ok], 1e-5, 3e-3, "energy")
    _close(info.acceptance_rate[okt], oinfo.acceptance_rate[ok], 1e-2, 5e-3, "acceptance")
    # accept decisions: identical wherever the uniform is further from the acceptance probability than 4x the
    # largest observed |p_accept(GPU) - p_accept(oracle)| (same uniforms: the accept stream is bit-exact)
    _close(inf["accept_uniform"], oinfo.extra["u"], 0, 0, "accept uniforms")
    got_acc = info.is_accepted.cpu().numpy()
    dp = np.abs(info.acceptance_rate.cpu().numpy() - oinfo.acceptance_rate)[ok].max()
    clear = np.abs(oinfo.extra["u"] - oinfo.acceptance_rate) > 4 * dp + 1e-6
    assert clear[ok].mean() > 0.9, dp
    np.testing.assert_array_equal(got_acc[clear & ok], oinfo.is_accepted[clear & ok])
    same = (got_acc == oinfo.is_accepted) & ok
    _close(new.position[_t(same, cuda)], onew.position[same], 1e-4, 1e-5)


@pytest.mark.parametrize("Nrows,D,C", [(200, 8, 40), (1000, 25, 130), (10000, 100, 6)])
def test_logreg_midpoint_map_closer_to_float64(cuda, Nrows, D, C):
    """North-star tolerance for a single integrator map evaluation (rel 1e-5, float32) in the form that does not
    depend on the float32 oracle's own round-off: |CUDA - float64 oracle| <= max(2 |float32 oracle - float64 oracle|,
    1e-5), norm-wise per chain, for every output of the implicit-midpoint map (rmhmc/integrators.py:119-142)."""
    import geomjax_b200 as g
    X, y = T.make_logreg_data(Nrows, D, seed=5)
    t32 = T.LogisticRegression(X, y, 0.01)
    t64 = T.LogisticRegression(X.astype(np.float64), y.astype(np.float64), 0.01, dtype=np.float64)
    t32.structured_dmetric = t64.structured_dmetric = Nrows * D ** 3 > 1e9
    rng = np.random.default_rng(D)
    q = (0.2 * rng.standard_normal((C, D))).astype(np.float32)
    p = (np.sqrt(Nrows) * 0.3 * rng.standard_normal((C, D))).astype(np.float32)
    he = 0.05
    want, o32 = {}, {}
    for tgt, dst, dt in ((t64, want, np.float64), (t32, o32, np.float32)):
        dT, v = S._rmhmc_kinetic_grad(tgt, q.astype(dt), p.astype(dt))
        gr = tgt.grad(q.astype(dt))
        dst.update(velocity=v, dHdq=dT - gr, q=q.astype(dt) + dt(he) * v, p=p.astype(dt) - dt(he) * (dT - gr),
                   logdet=np.linalg.slogdet(tgt.metric(q.astype(dt)))[1])
    target = g.logistic_regression(_t(X, cuda), _t(y, cuda), 0.01)
    out = g.LockstepPlan(target, C, cuda).evaluate(0, _t(q, cuda), _t(p, cuda), half_step=he)
    report = {}
    for k in ("velocity", "dHdq", "q", "p"):
        sc = np.abs(want[k]).max(axis=1, keepdims=True)
        e_gpu = (np.abs(out[k].cpu().numpy().astype(np.float64) - want[k]) / sc).max()
        e_o32 = (np.abs(o32[k].astype(np.float64) - want[k]) / sc).max()
        report[k] = (e_gpu, e_o32)
    e_gpu = np.abs(out["logdet"].cpu().numpy() - want["logdet"]).max() / np.abs(want["logdet"]).max()
    e_o32 = np.abs(o32["logdet"] - want["logdet"]).max() / np.abs(want["logdet"]).max()
    report["logdet"] = (e_gpu, e_o32)
    for k, (eg, eo) in report.items():
        assert eg <= max(2.0 * eo, 1e-5), report


def test_logreg_fused_equals_stepwise_and_limits(cuda):
    import torch
    import geomjax_b200 as g
    X, y = T.make_logreg_data(120, 6, seed=3)
    target = g.logistic_regression(_t(X, cuda), _t(y, cuda), 0.01)
    alg = g.rmhmc(target, 0.1, target, 2)
    st0 = alg.init(torch.zeros((33, 6), device=cuda))
    root = g.random.PRNGKey(0)
    st = st0
    for t in range(3):
        st, _ = alg.step(g.random.chain_keys(root, t, 3, 33), st)
    fst, samples, _ = g.run_fused(alg.step, root, st0, 3, return_samples=True)
    assert bool((fst.position == st.position).all()) and bool((samples[-1] == st.position).all())
    # lmc / lmcmonge on logreg and designs that do not fit shared memory are refused loudly
    with pytest.raises(g._native.NativeError):
        g.lmc(target, 0.1, target, 2).step(g.random.chain_keys(root, 0, 3, 33), g.lmc.init(st0.position, target))
    Xb, yb = T.make_logreg_data(500, 125, seed=3)  # D > 124 is not built
    big = g.logistic_regression(_t(Xb, cuda), _t(yb, cuda), 0.01)
    with pytest.raises(g._native.NativeError):
        g.rmhmc.init(torch.zeros((2, 125), device=cuda), big)


@pytest.mark.parametrize("Nrows,D,C", [(1000, 25, 300), (64, 4, 7), (333, 17, 129), (2000, 32, 130)])
def test_fisher_metric_tcgen05_vs_oracle(cuda, Nrows, D, C):
    """vmap(metric_fn)(position) as one tcgen05 3xTF32 GEMM over the chain dimension vs the float64
    oracle: error must stay at FP32 level (a single-pass TF32 product would be ~5e-4)."""
    import geomjax_b200 as g
    X, y = T.make_logreg_data(Nrows, D, seed=2)
    t64 = T.LogisticRegression(X.astype(np.float64), y.astype(np.float64), 0.01, dtype=np.float64)
    q = (0.3 * np.random.default_rng(C).standard_normal((C, D))).astype(np.float32)
    want = t64.metric(q.astype(np.float64))
    target = g.logistic_regression(_t(X, cuda), _t(y, cuda), 0.01)
    got = target.evaluate_metric(_t(q, cuda)).cpu().numpy()
    scale = np.abs(want).max(axis=(1, 2), keepdims=True)
    err = np.abs(got - want) / scale
    assert err.max() < 5e-6, err.max()  # FP32-level (two-level accumulation, see fisher_tc.cu)
    np.testing.assert_array_equal(got, got.transpose(0, 2, 1))


@pytest.mark.parametrize("Nrows,D", [(150, 6), (200, 40)])
def test_step_size_adaptation_on_logreg(cuda, Nrows, D):
    """c5's warm-up: vmapped dual averaging fused into the rmhmc logistic-regression kernels (per-chain step
    size from the dual-averaging state, update in the kernel epilogue); D = 40 runs rmhmc_logreg_big.cu."""
    import torch
    import geomjax_b200 as g
    X, y = T.make_logreg_data(Nrows, D, seed=4)
    target = g.logistic_regression(_t(X, cuda), _t(y, cuda), 0.01)
    C, W = 12, 40
    res, info = g.step_size_adaptation(g.rmhmc, target, initial_step_size=0.5, metric_fn=target,
                                       num_integration_steps=2).run(g.random.PRNGKey(2), torch.zeros((C, D), device=cuda), W)
    eps = res.parameters["step_size"]
    assert eps.shape == (C,) and bool(torch.isfinite(eps).all()) and bool((eps >= 1e-3).all())
    assert float(eps.std()) > 0  # per-chain adaptation
    acc = info["acceptance_rate"]
    assert acc.shape == (W, C) and 0.5 < float(acc[-15:].mean()) <= 1.0
    # first transition against the oracle (same keys: split(key_c, W)[0], step size 0.5)
    import geomjax_b200.random as R
    from oracle import prng as P
    keys = P.split(P.split(P.key(2), C), W)[:, 0]
    tgt = T.LogisticRegression(X, y, 0.01)
    _, oi = S.rmhmc_step(keys, S.rmhmc_init(np.zeros((C, D), np.float32), tgt), tgt, 0.5, 2)
    ok = oi.extra["fp_iters"] < 100
    np.testing.assert_allclose(acc[0].cpu().numpy()[ok], oi.acceptance_rate[ok], rtol=0, atol=5e-3)
    # adapted parameters plug back into the sampler
    alg = g.rmhmc(target, eps, target, 2)
    st, i2 = alg.step(g.random.chain_keys(g.random.PRNGKey(9), 0, 1, C), res.state)
    assert bool(torch.isfinite(st.position).all())


@pytest.mark.parametrize("Nrows,D,C", [(1000, 25, 300), (70, 4, 7), (333, 17, 129), (10000, 100, 130)])
def test_quadratic_forms_tcgen05_vs_float64(cuda, Nrows, D, C):
    """h[c, n] = x_n^T A_c x_n (GEMM 5 of SURVEY Appendix B.1) on the warp-specialised tcgen05 pipeline vs
    float64 NumPy, with A_c = the inverse Fisher metric of random positions; error at FP32 level."""
    import geomjax_b200 as g
    X, y = T.make_logreg_data(Nrows, D, seed=2)
    t64 = T.LogisticRegression(X.astype(np.float64), y.astype(np.float64), 0.01, dtype=np.float64)
    q = (0.3 * np.random.default_rng(C).standard_normal((C, D)))
    A = np.linalg.inv(t64.metric(q))
    A = 0.5 * (A + A.transpose(0, 2, 1))
    want = np.einsum("ni,cij,nj->cn", X.astype(np.float64), A, X.astype(np.float64), optimize=True)
    target = g.logistic_regression(_t(X, cuda), _t(y, cuda), 0.01)
    got = target.quadratic_forms(_t(A.astype(np.float32), cuda)).cpu().numpy()
    A32 = A.astype(np.float32).astype(np.float64)  # the kernel sees the float32-rounded matrices
    want32 = np.einsum("ni,cij,nj->cn", X.astype(np.float64), A32, X.astype(np.float64), optimize=True)
    scale = np.abs(want).max(axis=1, keepdims=True)
    err = np.abs(got - want32) / scale
    assert err.max() < 5e-6, err.max()


@pytest.mark.parametrize("Nrows,D,C", [(200, 8, 40), (1000, 25, 130), (10000, 100, 6)])
def test_lockstep_midpoint_map_vs_oracle(cuda, Nrows, D, C):
    """One evaluation of the implicit-midpoint map for all chains at once, both D^2 N products on the tcgen05
    GEMMs (the unit of a lock-step fixed-point loop), against the oracle's dense evaluation
    (rmhmc/integrators.py:119-142 via oracle.samplers._rmhmc_kinetic_grad)."""
    import geomjax_b200 as g
    X, y = T.make_logreg_data(Nrows, D, seed=5)
    tgt = T.LogisticRegression(X, y, 0.01)
    tgt.structured_dmetric = Nrows * D ** 3 > 1e9
    t64 = T.LogisticRegression(X.astype(np.float64), y.astype(np.float64), 0.01, dtype=np.float64)
    t64.structured_dmetric = tgt.structured_dmetric
    rng = np.random.default_rng(D)
    q = (0.2 * rng.standard_normal((C, D))).astype(np.float32)
    p = (np.sqrt(Nrows) * 0.3 * rng.standard_normal((C, D))).astype(np.float32)
    qi = (q + 0.01 * rng.standard_normal((C, D))).astype(np.float32)
    pi = (p + 0.1 * rng.standard_normal((C, D))).astype(np.float32)
    he = 0.05
    dT64, v64 = S._rmhmc_kinetic_grad(t64, q.astype(np.float64), p.astype(np.float64))
    g64 = t64.grad(q.astype(np.float64))
    target = g.logistic_regression(_t(X, cuda), _t(y, cuda), 0.01)
    plan = g.LockstepPlan(target, C, cuda)
    out = plan.evaluate(0, _t(q, cuda), _t(p, cuda), _t(qi, cuda), _t(pi, cuda), he)
    end = plan.evaluate(1, _t(q, cuda), _t(p, cuda))
    n = lambda k: (end if k in ("logdensity", "logdensity_grad") else out)[k].cpu().numpy().astype(np.float64)
    vs = np.abs(v64).max(axis=1, keepdims=True)
    assert (np.abs(n("velocity") - v64) / vs).max() < 2e-4
    gs = np.abs(g64).max(axis=1, keepdims=True)
    assert (np.abs(n("logdensity_grad") - g64) / gs).max() < 2e-5
    np.testing.assert_allclose(n("logdensity"), t64.logp(q.astype(np.float64)), rtol=2e-5)
    np.testing.assert_allclose(n("logdet"), np.linalg.slogdet(t64.metric(q.astype(np.float64)))[1], rtol=1e-5, atol=1e-3)
    ds = np.abs(dT64).max(axis=1, keepdims=True)
    dH64 = dT64 - g64
    assert (np.abs(n("dHdq") - dH64) / np.maximum(ds, gs)).max() < 5e-4
    assert (np.abs(end["velocity"].cpu().numpy() - v64) / vs).max() < 2e-4
    qn64 = qi.astype(np.float64) + he * v64
    pn64 = pi.astype(np.float64) - he * (dT64 - g64)
    assert (np.abs(n("q") - qn64) / np.abs(qn64).max(axis=1, keepdims=True)).max() < 2e-4
    assert (np.abs(n("p") - pn64) / np.abs(pn64).max(axis=1, keepdims=True)).max() < 2e-4


def test_lockstep_divergent_chains_are_data_not_errors(cuda):
    """A step size far too large: the fixed point diverges / overflows for most chains.  As under jax.vmap, that is
    data -- the launch terminates (round bound), diverged proposals are rejected (NaN energy -> delta = -inf), the
    returned state stays finite and equals the input for rejected chains, and the oracle rejects the same chains."""
    import torch
    import geomjax_b200 as g
    X, y = T.make_logreg_data(200, 8, seed=1)
    tgt = T.LogisticRegression(X, y, 0.01)
    rng = np.random.default_rng(3)
    C = 64
    q = (0.1 * rng.standard_normal((C, 8))).astype(np.float32)
    keys = rng.integers(0, 2 ** 32, size=(C, 2), dtype=np.uint64).astype(np.uint32)
    target = g.logistic_regression(_t(X, cuda), _t(y, cuda), 0.01)
    alg = g.rmhmc(target, 20.0, target, 2)
    st = alg.init(_t(q, cuda))
    new, info = alg.step(_t(keys, cuda), st)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(new.position).all()) and bool(torch.isfinite(new.logdensity).all())
    rej = ~info.is_accepted
    assert bool((new.position[rej] == st.position[rej]).all())
    with np.errstate(all="ignore"):
        _, oinfo = S.rmhmc_step(keys, S.rmhmc_init(q, tgt), tgt, 20.0, 2)
    # the bulk of the chains is rejected on both sides; individual borderline chains may differ
    assert float(rej.float().mean()) > 0.5 and (~oinfo.is_accepted).mean() > 0.5
    agree = (info.is_accepted.cpu().numpy() == oinfo.is_accepted).mean()
    assert agree > 0.8, agree
    # and a fused multi-transition launch with such chains still terminates and keeps the state finite
    fst, samples, _ = g.run_fused(alg.step, g.random.PRNGKey(1), st, 3, return_samples=True)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(fst.position).all()) and bool(torch.isfinite(samples).all())
