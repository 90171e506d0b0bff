"""Time the two warp-specialised tcgen05 GEMMs of the logistic-regression target (CUDA events); GPU only:
   evaluate_metric   vec(G)[P, C] = Z^T[P, N] . W[N, C]       (Fisher metric of every chain)
   quadratic_forms   h[N, C]      = Z[N, P] . vecsym(A)[P, C] (x_n^T A_c x_n, A = G^-1 in rmhmc's dT/dq)
Both are 3xTF32: tensor flops = 3 x algorithmic flops.  Timings include the operand pre-kernels."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import geomjax_b200 as g
from bench.data import make_logreg_data


def timed(fn, reps=10):
    for _ in range(3):
        out = fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


for (N, D, C) in ((1000, 25, 16384), (10000, 100, 2048)):
    X, y = make_logreg_data(N, D, 0)
    dev = torch.device("cuda:0")
    t = g.logistic_regression(torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev), 0.01)
    q = 0.1 * torch.randn((C, D), device=dev)
    P = D * (D + 1) // 2
    ms, G = timed(lambda: t.evaluate_metric(q))
    print(f"evaluate_metric  N={N} D={D} C={C}: {ms:.3f} ms per call; algorithmic {2.0 * N * P * C / ms / 1e9:.2f} TFLOP/s "
          f"(x3 TF32 passes = {6.0 * N * P * C / ms / 1e9:.2f} tensor TFLOP/s)")
    A = torch.linalg.inv(G)
    ms, h = timed(lambda: t.quadratic_forms(A))
    print(f"quadratic_forms  N={N} D={D} C={C}: {ms:.3f} ms per call; algorithmic {2.0 * N * P * C / ms / 1e9:.2f} TFLOP/s "
          f"(x3 TF32 passes = {6.0 * N * P * C / ms / 1e9:.2f} tensor TFLOP/s)")
    # one lock-step round for explicit inputs (both GEMMs + batched Cholesky / inverse + O(ND) kernels)
    p = torch.randn((C, D), device=dev) * (N ** 0.5) * 0.3
    plan = g.LockstepPlan(t, C, dev)
    ms, out = timed(lambda: plan.evaluate(0, q, p, half_step=0.05), reps=5)
    feval = 2.0 * N * D * (D + 1) + 6.0 * N * D + D ** 3
    print(f"lockstep round   N={N} D={D} C={C}: {ms:.3f} ms per call = {C / ms * 1e3:.0f} chain-evaluations/s; "
          f"algorithmic {feval * C / ms / 1e9:.2f} TFLOP/s")
