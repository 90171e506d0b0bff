#!/usr/bin/env python
"""Regenerate the golden fixtures under tests/golden/.

    python tests/golden/make_golden.py

The reference (williwilliams3/geomjax) is Python on JAX; jax/jaxlib are not installable in the
build image, so the fixtures cannot be produced by importing it.  They are produced by the NumPy
oracle (``oracle/``), which is itself pinned to the reference's only author-produced vector and to
JAX's documented PRNG outputs (tests/test_oracle_pins.py), and cross-checked here against the
independent scratch restatement recorded in SURVEY.md Appendix A.2 (``appendix_a2.json``; those
numbers were typed in from the survey, not produced by this script).

``static_kernels.npz``: for every static kernel x target on the hot path, the inputs (positions, per-chain
keys) and the full (State, Info) of one transition, float32, legacy threefry.  Used by
tests/test_golden.py: the oracle must reproduce them (CPU, regression) and the CUDA path must
match them through the C ABI (GPU).
"""
from __future__ import annotations

import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import prng as P  # noqa: E402
from oracle import samplers as S  # noqa: E402
from oracle import targets as T  # noqa: E402

# name -> (sampler, target kind, D, C, L, eps, extra)
CASES = {
    "rmhmc_funnel_d2": ("rmhmc", "funnel", 2, 16, 4, 0.1, {}),
    "rmhmc_funnel_d20": ("rmhmc", "funnel", 20, 16, 3, 0.02, {}),
    "lmc_funnel_d2": ("lmc", "funnel", 2, 16, 4, 0.1, {}),
    "lmc_funnel_d20": ("lmc", "funnel", 20, 16, 3, 0.05, {}),
    "lmc_funnel_d100": ("lmc", "funnel", 100, 8, 2, 0.02, {}),
    "lmcmonge_funnel_d2": ("lmcmonge", "funnel", 2, 16, 4, 0.01, {"half_step": "omega"}),
    "lmcmonge_funnel_d20": ("lmcmonge", "funnel", 20, 16, 4, 0.002, {"half_step": "omega"}),
    "lmcmonge_fixed_funnel_d20": ("lmcmonge", "funnel", 20, 16, 4, 0.1, {"half_step": "omega_fixed"}),
    "rmhmc_logreg_d5": ("rmhmc", "logreg", 5, 8, 2, 0.1, {"N": 64}),
    "rmhmc_logreg_d25": ("rmhmc", "logreg", 25, 4, 2, 0.1, {"N": 200}),
    "rmhmc_softabs_funnel_d2": ("rmhmc", "softabs", 2, 16, 4, 0.05, {}),
}

# The exact shapes BASELINE.json names (bench/configs.json: init position, step size, L, data seed 0) with the bench's
# key tree split(split(PRNGKey(0), T)[0], C)[c]; written to baseline_shapes.npz.
BASELINE_CASES = {
    "c1_lmc_funnel_d2_as_shipped": ("lmc", "funnel", 2, 8, 8, 0.1, {"init": "ones", "total_transitions": 1000}),
    "c2_lmcmonge_funnel_d20": ("lmcmonge", "funnel", 20, 64, 8, 0.001016, {"half_step": "omega", "init": "ones"}),
    "c2_lmcmonge_fixed_funnel_d20": ("lmcmonge", "funnel", 20, 64, 8, 0.3509, {"half_step": "omega_fixed", "init": "ones"}),
    "c3_lmc_funnel_d100": ("lmc", "funnel", 100, 8, 8, 0.05, {"init": "ones"}),
    "c4_rmhmc_logreg_d25_n1000": ("rmhmc", "logreg", 25, 16, 6, 0.1, {"N": 1000, "init": "zeros"}),
}
CASES.update(BASELINE_CASES)


def make_target(kind, D, extra):
    if kind == "funnel":
        return T.NealFunnel(D)
    if kind == "softabs":
        return T.softabs_metric(T.NealFunnel(D), alpha=1e6)
    if kind == "logreg":
        X, y = T.make_logreg_data(extra["N"], D, 0)
        tgt = T.LogisticRegression(X, y, 0.01)
        # the same contractions re-associated so that dG (C, D, D, D) is never materialised (minutes otherwise)
        tgt.structured_dmetric = extra["N"] * D ** 3 > 1e7
        return tgt
    raise ValueError(kind)


def make_inputs(kind, D, C, seed, extra=None):
    if extra and "init" in extra:
        q = (np.ones if extra["init"] == "ones" else np.zeros)((C, D), np.float32)
        return q, S.chain_keys(P.key(0), extra.get("total_transitions", 1 << 20), 0, C)
    rng = np.random.default_rng(seed)
    if kind == "logreg":
        q = (0.3 * rng.standard_normal((C, D))).astype(np.float32)
    else:
        # SoftAbs with alpha = 1e6 is |H|, singular where the funnel's Hessian changes signature
        # (x^2 e^-v = 2 / 9): start those chains inside the positive-definite region
        sx = 0.2 if kind == "softabs" else 0.5
        v = 0.35 * rng.standard_normal((C, 1))
        q = np.concatenate([np.exp(0.5 * v) * sx * rng.standard_normal((C, D - 1)), v], 1).astype(np.float32)
    keys = S.chain_keys(P.key(seed), 7, 3, C)
    return q, keys


def run_case(name):
    sampler, kind, D, C, L, eps, extra = CASES[name]
    tgt = make_target(kind, D, extra)
    q, keys = make_inputs(kind, D, C, seed=sum(map(ord, name)), extra=extra)
    if sampler == "rmhmc":
        st, info = S.rmhmc_step(keys, S.rmhmc_init(q, tgt), tgt, eps, L)
    elif sampler == "lmc":
        st, info = S.lmc_step(keys, S.lmc_init(q, tgt), tgt, eps, L)
    else:
        st, info = S.lmcmonge_step(keys, S.lmcmonge_init(q, tgt), tgt, eps, np.ones(D, np.float32), L,
                                   half_step=extra["half_step"])
    out = {"in_position": q, "in_keys": keys, "step_size": np.float32(eps), "L": np.int32(L),
           "position": st.position, "logdensity": st.logdensity, "logdensity_grad": st.logdensity_grad,
           "draw": info.momentum, "acceptance_rate": info.acceptance_rate, "is_accepted": info.is_accepted,
           "is_divergent": info.is_divergent, "energy": info.energy, "z": info.extra["z"], "u": info.extra["u"],
           "H0": info.extra["H0"], "proposal_position": info.proposal["position"],
           "proposal_logdensity": info.proposal["logdensity"]}
    if sampler != "rmhmc":
        out["volume_adjustment"] = st.volume_adjustment
        out["proposal_volume_adjustment"] = info.proposal["volume_adjustment"]
    else:
        out["fp_iters"] = info.extra["fp_iters"]
    return out


def main():
    try:
        rev = subprocess.run(["git", "-C", ROOT, "rev-parse", "HEAD"], capture_output=True, text=True).stdout.strip()
    except Exception:
        rev = "unknown"
    only_baseline = "--baseline-only" in sys.argv
    for fname, names in (("static_kernels.npz", [n for n in CASES if n not in BASELINE_CASES]),
                         ("baseline_shapes.npz", list(BASELINE_CASES))):
        if only_baseline and fname != "baseline_shapes.npz":
            continue
        flat = {}
        for name in names:
            for k, v in run_case(name).items():
                flat[f"{name}/{k}"] = np.asarray(v)
        flat["_meta/oracle_git_rev"] = np.array(rev)
        np.savez_compressed(os.path.join(HERE, fname), **flat)
        print(f"wrote {fname}: {len(flat)} arrays, oracle rev {rev}")


if __name__ == "__main__":
    main()
