"""Fused CUDA lmcmonge transition vs the NumPy oracle (same keys, same state)."""
import numpy as np
import pytest

from oracle import prng as P
from oracle import samplers as S
from oracle import targets as T

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # BASELINE.json north_star: single-step outputs within rel 1e-5 (fp32)


def _t(x, dev):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def _close(got, want, rtol=RTOL, atol=1e-6, what=""):
    got = got.cpu().numpy() if hasattr(got, "cpu") else np.asarray(got)
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol, err_msg=what)


def _setup(D, C, seed=0, scale=1.0):
    """Positions drawn from a (narrowed) funnel so that trajectories are well conditioned."""
    rng = np.random.default_rng(seed)
    v = 0.7 * scale * rng.standard_normal((C, 1))
    q = np.concatenate([np.exp(0.5 * v) * rng.standard_normal((C, D - 1)) * scale, v], axis=1).astype(np.float32)
    keys = rng.integers(0, 2 ** 32, size=(C, 2), dtype=np.uint64).astype(np.uint32)
    return q, keys


@pytest.mark.parametrize("L,rtol", [(1, 1e-5), (8, 1e-4)])
@pytest.mark.parametrize("D,lpc", [(2, 1), (2, 2), (5, 4), (20, 1), (20, 2), (20, 4), (20, 8), (33, 32), (100, 4), (100, 8)])
def test_init_and_transition_vs_oracle(cuda, D, lpc, L, rtol):
    """L=1: the single-integrator-step contract (rel 1e-5, BASELINE.json north_star).
    L=8: a full trajectory; rounding differences grow along it, so 1e-4 on the chains whose
    trajectory is tame (|H0-H1| < 100); wild chains must be rejected/flagged identically."""
    import geomjax_b200 as g
    C = 300
    q, keys = _setup(D, C, seed=D, scale=0.5)
    tgt = T.NealFunnel(D)
    ost = S.lmcmonge_init(q, tgt)
    im = (0.5 + np.random.default_rng(1).random(D)).astype(np.float32)
    target = g.neal_funnel(D)
    # the as-written "omega" half step has an O(eps * D) energy error: keep eps ~ 1/(D L)
    eps = 0.1 / (D * L)
    alg = g.lmcmonge(target, eps, _t(im, cuda), L, alpha2=1e-3, lanes_per_chain=lpc)
    st = alg.init(_t(q, cuda))
    _close(st.logdensity, ost.logdensity, what="init logdensity")
    _close(st.logdensity_grad, ost.logdensity_grad, what="init grad")
    assert float(st.volume_adjustment.abs().max()) == 0.0
    new, info = alg.step(_t(keys, cuda), st)
    onew, oinfo = S.lmcmonge_step(keys, ost, tgt, eps, im, L, alpha2=1e-3)
    _close(info.velocity, oinfo.momentum, what="velocity draw")
    with np.errstate(invalid="ignore"):
        tame = np.isfinite(oinfo.proposal["weight"]) & (np.abs(oinfo.proposal["weight"]) < 100)
    assert tame.mean() > 0.9 and oinfo.is_accepted.mean() > 0.2
    tm = _t(tame, cuda)
    ps = info.proposal.state
    _close(ps.position[tm], oinfo.proposal["position"][tame], rtol=rtol, what="proposal position")
    _close(ps.velocity[tm], oinfo.proposal["velocity"][tame], rtol=rtol, atol=1e-5, what="proposal velocity")
    _close(ps.momentum[tm], oinfo.proposal["momentum"][tame], rtol=rtol, atol=1e-5, what="proposal momentum")
    _close(ps.logdensity[tm], oinfo.proposal["logdensity"][tame], rtol=rtol, atol=1e-4)
    _close(ps.logdensity_grad[tm], oinfo.proposal["logdensity_grad"][tame], rtol=rtol, atol=1e-4)
    _close(ps.volume_adjustment[tm], oinfo.proposal["volume_adjustment"][tame], rtol=rtol, atol=1e-5)
    _close(info.energy[tm], oinfo.energy[tame], rtol=rtol, atol=1e-4, what="energy")
    _close(info.acceptance_rate[tm], oinfo.acceptance_rate[tame], rtol=10 * rtol, atol=5e-4)
    # accept decisions: identical except where |u - p| is at rounding level
    got_acc = info.is_accepted.cpu().numpy()
    # window = 4x the largest observed |p_accept(GPU) - p_accept(oracle)|, not a fixed constant
    dp = np.abs(info.acceptance_rate.cpu().numpy() - oinfo.acceptance_rate)[tame].max()
    clear = np.abs(oinfo.extra["u"] - oinfo.acceptance_rate) > 4 * dp + 1e-6
    assert clear.mean() > 0.97, dp
    np.testing.assert_array_equal(got_acc[clear], oinfo.is_accepted[clear])
    np.testing.assert_array_equal(info.is_divergent.cpu().numpy()[~tame | clear], oinfo.is_divergent[~tame | clear])
    same = (got_acc == oinfo.is_accepted) & tame
    _close(new.position[_t(same, cuda)], onew.position[same], rtol=rtol)
    _close(new.logdensity_grad[_t(same, cuda)], onew.logdensity_grad[same], rtol=rtol, atol=1e-4)
    _close(new.volume_adjustment[_t(same, cuda)], onew.volume_adjustment[same], rtol=rtol, atol=1e-5)


@pytest.mark.parametrize("half_step", ["omega", "omega_fixed", "omegatilde"])
def test_half_step_variants(cuda, half_step):
    import geomjax_b200 as g
    from geomjax_b200 import integrators as I
    D, C = 6, 64
    q, keys = _setup(D, C, seed=5, scale=0.7)
    tgt = T.NealFunnel(D)
    ost = S.lmcmonge_init(q, tgt)
    integ = {"omega": I.half_step_omega, "omega_fixed": I.half_step_omega_fixed, "omegatilde": I.half_step_omegatilde}[half_step]
    alg = g.lmcmonge(g.neal_funnel(D), 0.1, _t(np.ones(D, np.float32), cuda), 5, alpha2=0.3, integrator=integ)
    new, info = alg.step(_t(keys, cuda), alg.init(_t(q, cuda)))
    onew, oinfo = S.lmcmonge_step(keys, ost, tgt, 0.1, np.ones(D, np.float32), 5, alpha2=0.3, half_step=half_step)
    with np.errstate(invalid="ignore"):
        tame = np.isfinite(oinfo.proposal["weight"]) & (np.abs(oinfo.proposal["weight"]) < 20)
    assert tame.mean() > 0.7
    _close(info.proposal.state.position[_t(tame, cuda)], oinfo.proposal["position"][tame], rtol=1e-4, atol=1e-5)
    _close(info.energy[_t(tame, cuda)], oinfo.energy[tame], rtol=1e-4, atol=1e-4)


def test_reference_test_setup_monge(cuda):
    """tests/test_samplers.py:39-53 setup: zeros(2), eps=1e-2, L=10, alpha2=0, key 42."""
    import geomjax_b200 as g
    alg = g.lmcmonge(g.neal_funnel(2), 1e-2, _t(np.ones(2, np.float32), cuda), 10, alpha2=0.0)
    st = alg.init(_t(np.zeros((1, 2), np.float32), cuda))
    new, info = alg.step(_t(P.key(42)[None], cuda), st)
    _close(new.position, np.array([[0.06474835, -0.07099186]], np.float32), rtol=1e-5)
    _close(info.acceptance_rate, np.array([0.9829441], np.float32), rtol=1e-5)


def test_build_kernel_signature(cuda):
    """lmcmonge/lmc.py:151-159: kernel(rng_key, state, logdensity_fn, step_size, inverse_mass_matrix, L, alpha2)."""
    import torch
    import geomjax_b200 as g
    D, C = 4, 16
    q, keys = _setup(D, C)
    target = g.neal_funnel(D)
    im = _t(np.ones(D, np.float32), cuda)
    kernel = g.lmcmonge.build_kernel()
    st = g.lmcmonge.init(_t(q, cuda), target)
    a, ia = kernel(_t(keys, cuda), st, target, 0.1, im, 3, 0.001)
    b, ib = g.lmcmonge(target, 0.1, im, 3).step(_t(keys, cuda), st)
    assert bool((a.position == b.position).all()) and torch.equal(ia.energy.nan_to_num(), ib.energy.nan_to_num())
    with pytest.raises(NotImplementedError):
        g.lmcmonge(lambda x: 0.0, 0.1, im, 3)
    with pytest.raises(ValueError):
        g.lmcmonge(target, 0.1, _t(np.eye(D, dtype=np.float32), cuda), 3).step(_t(keys, cuda), st)


def test_fused_equals_stepwise_and_sharding(cuda):
    """In-kernel key derivation == examples/funnel/main.py:18,22; results independent of how the
    global chain set is partitioned (multi-GPU determinism)."""
    import torch
    import geomjax_b200 as g
    import geomjax_b200.random as R
    D, C, T_ = 20, 96, 5
    target = g.neal_funnel(D)
    alg = g.lmcmonge(target, 0.1, _t(np.ones(D, np.float32), cuda), 4)
    root = P.key(0)
    st0 = alg.init(torch.ones((C, D), device=cuda))
    st = st0
    pos = []
    for t in range(T_):
        st, _ = alg.step(R.chain_keys(root, t, T_, C), st)
        pos.append(st.position.clone())
    fst, samples, acc = g.run_fused(alg.step, root, st0, T_, return_samples=True, return_accept=True)
    assert bool((fst.position == st.position).all())
    assert bool((samples == torch.stack(pos)).all())
    assert bool((fst.logdensity_grad == st.logdensity_grad).all())
    # two shards of the same global chain set
    half = C // 2
    for lo, hi in ((0, half), (half, C)):
        sub = g.LMCState(*[x[lo:hi].contiguous() for x in st0])
        sst, ssamples, _ = g.run_fused(alg.step, root, sub, T_, chain_offset=lo, total_chains=C, return_samples=True)
        assert bool((ssamples == samples[:, lo:hi]).all())
    # oracle for the first transition
    tgt = T.NealFunnel(D)
    ost = S.lmcmonge_init(np.ones((C, D), np.float32), tgt)
    onew, _ = S.lmcmonge_step(S.chain_keys(root, T_, 0, C), ost, tgt, 0.1, np.ones(D, np.float32), 4)
    _close(pos[0], onew.position)


def test_edge_cases(cuda):
    import torch
    import geomjax_b200 as g
    D = 3
    target = g.neal_funnel(D)
    alg = g.lmcmonge(target, 0.1, _t(np.ones(D, np.float32), cuda), 2)
    # empty batch
    st = alg.init(torch.empty((0, D), device=cuda))
    new, info = alg.step(torch.empty((0, 2), dtype=torch.uint32, device=cuda), st)
    assert new.position.shape == (0, D)
    # a NaN / overflowing start is data, not an error: rejected, flagged, state unchanged
    q = torch.tensor([[1.0, 1.0, 1.0], [0.0, 0.0, -200.0]], device=cuda)
    st = alg.init(q)
    keys = _t(np.array([[1, 2], [3, 4]], np.uint32), cuda)
    new, info = alg.step(keys, st)
    assert not bool(info.is_accepted[1])
    assert bool((new.position[1] == q[1]).all())
    # wrong shapes raise
    with pytest.raises(ValueError):
        alg.init(torch.ones((4, D + 1), device=cuda))
    with pytest.raises(ValueError):
        alg.step(keys[:1], st)
    with pytest.raises(Exception):
        alg.init(torch.ones((4, D)))  # CPU tensor: no CPU fallback
