"""Time the tcgen05 batched Fisher-metric GEMM (CUDA events); GPU only."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import geomjax_b200 as g
from oracle.targets import make_logreg_data

for (N, D, C) in ((1000, 25, 16384), (10000, 100, 2048)):
    X, y = make_logreg_data(N, D, 0)
    dev = torch.device("cuda:0")
    t = g.logistic_regression(torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev), 0.01)
    q = 0.1 * torch.randn((C, D), device=dev)
    for _ in range(3):
        G = t.evaluate_metric(q)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        G = t.evaluate_metric(q)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    P = D * (D + 1) // 2
    print(f"N={N} D={D} C={C}: {ms:.3f} ms per call; algorithmic {2.0 * N * P * C / ms / 1e9:.2f} TFLOP/s "
          f"(x3 TF32 passes = {6.0 * N * P * C / ms / 1e9:.2f} tensor TFLOP/s)")
