"""Timing sweep (CUDA events) of the fused kernels over L / lanes-per-chain / chains; GPU only."""
import argparse
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import geomjax_b200 as g

ap = argparse.ArgumentParser()
ap.add_argument("--sampler", default="lmcmonge")
ap.add_argument("--D", type=int, default=20)
ap.add_argument("--C", type=int, default=65536)
ap.add_argument("--Ls", default="1,8,32")
ap.add_argument("--lpcs", default="1,2,4")
ap.add_argument("--T", type=int, default=16)
ap.add_argument("--eps", type=float, default=0.001)
a = ap.parse_args()
dev = torch.device("cuda:0")
tgt = g.neal_funnel(a.D)
root = g.random.PRNGKey(0)
for lpc in [int(x) for x in a.lpcs.split(",")]:
    for L in [int(x) for x in a.Ls.split(",")]:
        if a.sampler == "lmcmonge":
            alg = g.lmcmonge(tgt, a.eps, torch.ones(a.D, device=dev), L, lanes_per_chain=lpc)
        elif a.sampler == "lmc":
            alg = g.lmc(tgt, a.eps, tgt, L, lanes_per_chain=lpc)
        else:
            alg = g.rmhmc(tgt, a.eps, tgt, L, lanes_per_chain=lpc)
        st = alg.init(torch.ones((a.C, a.D), device=dev))
        for _ in range(3):
            st, _, _ = g.run_fused(alg.step, root, st, a.T, total=1 << 20, inplace=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        reps = 5
        for r in range(reps):
            st, _, _ = g.run_fused(alg.step, root, st, a.T, first=(r + 3) * a.T, total=1 << 20, inplace=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{a.sampler} D={a.D} C={a.C} lpc={lpc} L={L} T={a.T}: {ms:.3f} ms/launch, "
              f"{ms * 1e3 / a.T:.2f} us/transition, {a.C * L * a.T / ms / 1e6:.1f} M chain-steps/s", flush=True)
