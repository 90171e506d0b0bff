"""CUDA threefry PRNG vs the oracle: bits, splits, uniforms AND normals bit-exact."""
import numpy as np
import pytest

from oracle import prng as P

pytestmark = pytest.mark.gpu


def _np(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("mode", ["legacy", "partitionable"])
def test_bits_uniform_normal_bit_exact(cuda, mode):
    import geomjax_b200.random as R
    R.set_threefry_partitionable(mode == "partitionable")
    try:
        rng = np.random.default_rng(0)
        keys = rng.integers(0, 2 ** 32, size=(257, 2), dtype=np.uint64).astype(np.uint32)
        keys[0] = P.key(0)
        keys[1] = P.key(42)
        for count in (1, 2, 3, 20, 25, 100, 101):
            want_bits = P.random_bits(keys, (count,), mode)
            np.testing.assert_array_equal(_np(R.bits(keys, count)).view(np.uint32), want_bits)
            np.testing.assert_array_equal(_np(R.uniform(keys, count)), P.uniform(keys, (count,), mode))
            got = _np(R.normal(keys, count))
            want = P.normal(keys, (count,), mode)
            np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
        for num in (1, 2, 4, 7, 1000):
            np.testing.assert_array_equal(_np(R.split(keys[:5], num)).view(np.uint32), P.split(keys[:5], num, mode))
    finally:
        R.set_threefry_partitionable(False)


def test_documented_jax_values(cuda):
    import geomjax_b200.random as R
    assert float(_np(R.uniform(P.key(0)[None], 1))[0, 0]) == np.float32(0.41845703)
    assert _np(R.normal(P.key(0)[None], 1))[0, 0] == np.float32(-0.20584226)
    np.testing.assert_array_equal(_np(R.normal(P.key(0)[None], 3))[0],
                                  np.array([1.8160863, -0.48262316, 0.33988908], np.float32))


@pytest.mark.parametrize("mode", ["legacy", "partitionable"])
def test_chain_keys(cuda, mode):
    import geomjax_b200.random as R
    from oracle.samplers import chain_keys
    R.set_threefry_partitionable(mode == "partitionable")
    try:
        root = P.key(0)
        got = _np(R.chain_keys(root, 3, 1000, 8)).view(np.uint32)
        np.testing.assert_array_equal(got, chain_keys(root, 1000, 3, 8, mode=mode))
        # a shard of a larger global chain set
        got = _np(R.chain_keys(root, 999, 1000, 5, chain_offset=11, total_chains=64)).view(np.uint32)
        np.testing.assert_array_equal(got, chain_keys(root, 1000, 999, 64, np.arange(11, 16), mode=mode))
    finally:
        R.set_threefry_partitionable(False)


def test_extreme_normal_tails(cuda):
    """bits near 0 / 2^32-1 hit the w >= 5 branch and the clamp at lo."""
    import ctypes as C
    import torch
    from geomjax_b200 import _native as N
    # drive the conversion through keys whose first bits element we know from the oracle
    keys = np.random.default_rng(3).integers(0, 2 ** 32, size=(200000, 2), dtype=np.uint64).astype(np.uint32)
    import geomjax_b200.random as R
    got = _np(R.normal(keys, 2))
    want = P.normal(keys, (2,))
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.abs(want).max() > 4.0


@pytest.mark.parametrize("sampler,D,lpc", [("lmc", 100, 4), ("lmc", 20, 4), ("lmc", 20, 2), ("lmc", 20, 1), ("lmc", 33, 32),
                                           ("lmc", 7, 1), ("lmcmonge", 20, 2), ("rmhmc", 20, 4), ("lmc", 100, 8)])
def test_in_kernel_noise_and_uniform_bit_exact(cuda, sampler, D, lpc):
    """The z ~ normal(k_m, (D,)) and u ~ uniform(k_a) a fused transition actually uses (every noise-loop
    variant: paired counters in one lane, pairs split across lanes, generic) == jax.random semantics."""
    import torch
    import geomjax_b200 as g
    from geomjax_b200 import _native as N
    from oracle import prng as P
    C = 37
    rng = np.random.default_rng(D * 7 + lpc)
    keys = rng.integers(0, 2 ** 32, size=(C, 2), dtype=np.uint64).astype(np.uint32)
    target = g.neal_funnel(D)
    if sampler == "lmc":
        alg = g.lmc(target, 0.01, target, 1, lanes_per_chain=lpc)
    elif sampler == "rmhmc":
        alg = g.rmhmc(target, 0.01, target, 1, lanes_per_chain=lpc)
    else:
        alg = g.lmcmonge(target, 0.001, torch.ones(D, device=cuda), 1, lanes_per_chain=lpc)
    st = alg.init(0.1 * torch.ones((C, D), device=cuda))
    ks = N.KeySource()
    kt = torch.from_numpy(keys).to(cuda)
    ks.keys, ks.num_transitions = N.ptr(kt), 1
    _, info = alg.step.engine.launch(st, ks, want_info=True, extra_info=True)
    km, ka = P.split(keys, 2)[:, 0], P.split(keys, 2)[:, 1]
    z = np.stack([P.normal(km[c], (D,)) for c in range(C)])
    u = np.array([P.uniform(ka[c]) for c in range(C)], np.float32)
    np.testing.assert_array_equal(info["noise"].cpu().numpy(), z)
    np.testing.assert_array_equal(info["accept_uniform"].cpu().numpy(), u)
