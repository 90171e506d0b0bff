"""Analytic derivatives in oracle/targets.py against central finite differences (float64)."""
import numpy as np
import pytest

from oracle import targets as T


def _targets():
    X, y = T.make_logreg_data(40, 5, dtype=np.float64)
    return [T.NealFunnel(2, dtype=np.float64), T.NealFunnel(6, dtype=np.float64),
            T.Gaussian(np.arange(4.0), 1.0 + np.arange(4.0), dtype=np.float64),
            T.Banana(dtype=np.float64), T.LogisticRegression(X, y, dtype=np.float64)]


@pytest.mark.parametrize("tg", _targets(), ids=lambda t: f"{t.name}{t.D}")
def test_derivatives(tg):
    rng = np.random.default_rng(1)
    q = rng.standard_normal((3, tg.D)) * 0.7
    u = rng.standard_normal((3, tg.D))
    h = 1e-6
    g = tg.grad(q)
    Hu = tg.hvp(q, u)
    dG = tg.dmetric(q)
    for i in range(tg.D):
        e = np.zeros(tg.D)
        e[i] = h
        fd = (tg.logp(q + e) - tg.logp(q - e)) / (2 * h)
        np.testing.assert_allclose(g[:, i], fd, rtol=1e-6, atol=1e-7)
        fdG = (tg.metric(q + e) - tg.metric(q - e)) / (2 * h)
        np.testing.assert_allclose(dG[..., i], fdG, rtol=1e-6, atol=1e-7)
    fdH = (tg.grad(q + h * u) - tg.grad(q - h * u)) / (2 * h)
    np.testing.assert_allclose(Hu, fdH, rtol=1e-5, atol=1e-6)


def test_funnel_metric_closed_forms():
    """chol(G) = (J^-1)^T, logdet G = -(D-1) v - 2 log sigma."""
    f = T.NealFunnel(5, dtype=np.float64)
    q = np.random.default_rng(0).standard_normal((4, 5))
    G = f.metric(q)
    A = f.inverse_jacobian(q)
    np.testing.assert_allclose(np.linalg.cholesky(G), A.transpose(0, 2, 1), atol=1e-12)
    np.testing.assert_allclose(np.linalg.slogdet(G)[1], -(5 - 1) * q[:, -1] - 2 * np.log(3.0), atol=1e-12)


def test_logreg_structured_dmetric_contraction_equals_dense():
    """The D = 100 / N = 10,000 oracle path contracts d_i G without materialising it; same numbers."""
    from oracle import samplers as S
    X, y = T.make_logreg_data(50, 6, dtype=np.float64)
    t = T.LogisticRegression(X, y, dtype=np.float64)
    rng = np.random.default_rng(2)
    q, p = 0.4 * rng.standard_normal((3, 6)), rng.standard_normal((3, 6))
    dense, v1 = S._rmhmc_kinetic_grad(t, q, p)
    t.structured_dmetric = True
    fast, v2 = S._rmhmc_kinetic_grad(t, q, p)
    np.testing.assert_allclose(fast, dense, rtol=1e-11, atol=1e-13)
    np.testing.assert_array_equal(v1, v2)
