"""Frozen synthetic inputs of the benchmark workloads.  Input data only: the
product arm of bench.py and the tools read it from here, never from oracle/ (tests/test_cabi_and_host.py checks
that the oracle's generator yields the same arrays)."""
import numpy as np


def make_logreg_data(N, D, seed=0, dtype=np.float32):
    """Design matrix with an intercept column and standardised N(0, 1) features, responses drawn from the
    logistic model at theta* ~ N(0, 1) / sqrt(D)."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, D))
    feats = X[:, 1:]
    X[:, 1:] = (feats - feats.mean(0)) / feats.std(0)
    X[:, 0] = 1.0
    theta = rng.standard_normal(D)
    prob = 1.0 / (1.0 + np.exp(-(X @ theta) / np.sqrt(D)))
    y = (rng.random(N) < prob).astype(np.float64)
    return X.astype(dtype), y.astype(dtype)
