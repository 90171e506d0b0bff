"""c5-shaped rmhmc (logistic regression D=100, N=10,000): chain-leapfrog-steps/s of the lock-step tcgen05 sampler
(geomjax_b200.rmhmc_lockstep) beside the CTA-per-chain FP32 kernel (geomjax_b200.rmhmc); GPU only."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import geomjax_b200 as g
from oracle.targets import make_logreg_data

N, D, L = int(os.environ.get("ROWS", 10000)), int(os.environ.get("DIM", 100)), 6
eps = float(os.environ.get("EPS", 0.05))
C = int(os.environ.get("CHAINS", 2048))
X, y = make_logreg_data(N, D, 0)
dev = torch.device("cuda:0")
t = g.logistic_regression(torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev), 0.01)
root = g.random.PRNGKey(0)
for name, alg, T_ in (("lockstep tcgen05", g.rmhmc_lockstep(t, eps, t, L), 2), ("CTA-per-chain fp32", g.rmhmc(t, eps, t, L), 1)):
    st = alg.init(torch.zeros((C, D), device=dev))
    st, info = alg.step(g.random.chain_keys(root, 0, 100, C), st)  # warm-up (and moves off the start point)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    acc = []
    for k in range(T_):
        st, info = alg.step(g.random.chain_keys(root, 1 + k, 100, C), st)
        acc.append(float(info.acceptance_rate.mean()))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name:20s} C={C} L={L}: {dt / T_ * 1e3:.1f} ms per transition, {C * L * T_ / dt:.0f} chain-leapfrog-steps/s, "
          f"mean acceptance {sum(acc) / len(acc):.3f}", flush=True)
