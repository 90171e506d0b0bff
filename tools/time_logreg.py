"""Per-evaluation cost of the logreg rmhmc kernels: vary fp max_iters (GPU only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import geomjax_b200 as g
from oracle.targets import make_logreg_data
N, D, C, L = 1000, 25, int(os.environ.get("CHAINS", 9472)), 6
X, y = make_logreg_data(N, D, 0)
dev = torch.device("cuda:0")
t = g.logistic_regression(torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev), 0.01)
root = g.random.PRNGKey(0)
for mi in (0, 4, 8, 100):
    alg = g.rmhmc(t, 0.1, t, L, integrator=g.integrators.implicit_midpoint(max_iters=mi))
    st = alg.init(torch.zeros((C, D), device=dev))
    st, _, _ = g.run_fused(alg.step, root, st, 1, total=100, inplace=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    st, _, _ = g.run_fused(alg.step, root, st, 1, first=1, total=100, inplace=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    evals = L * (2 + mi) + 2
    print(f"TC={os.environ.get('GB200_LOGREG_TC','1')} C={C} max_iters={mi}: {ms:.1f} ms/transition; if every step used max_iters: {ms*1e3/evals:.0f} us per lock-step evaluation round", flush=True)
