"""Experiment tool: per-CTA time stamps of the two tcgen05 GEMMs of one lock-step evaluation.
Needs a library built with the stamps compiled in:
    GEOMB200_EXTRA_NVCC_FLAGS=-DGB_FT_TIMING python -m geomjax_b200.build
    GEOMB200_EXTRA_NVCC_FLAGS=-DGB_FT_TIMING PYTHONPATH=. python tools/ft_stamps.py N D C
(build again without the variable afterwards: the flag is part of the build digest).  Output: profiles/r02_ft_stamps.txt."""
import ctypes as Ct
import sys

import numpy as np
import torch

import geomjax_b200 as g
from geomjax_b200 import _native as Nn
from geomjax_b200.plan import LockstepPlan
from bench.data import make_logreg_data


def main():
    N, D, C = (int(a) for a in sys.argv[1:4])
    X, y = make_logreg_data(N, D, seed=0)
    dev = torch.device("cuda:0")
    target = g.logistic_regression(torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev), 0.01)
    plan = LockstepPlan(target, C, dev)
    gen = torch.Generator(device="cpu").manual_seed(1)
    q = (0.1 * torch.randn((C, D), generator=gen)).to(dev)
    p = torch.randn((C, D), generator=gen).to(dev)
    for _ in range(3):
        plan.evaluate(0, q, p, q, p, 0.05)
    torch.cuda.synchronize()
    buf = np.zeros(8192 * 16, np.uint64)
    lib = Nn.lib()
    lib.gb200_debug_ft_stamps.argtypes = [Ct.c_void_p]
    rc = lib.gb200_debug_ft_stamps(buf.ctypes.data)
    assert rc == 0, rc
    st = buf.reshape(8192, 16).astype(np.int64)
    names = ["entry->setup", "setup->first MMA", "first->last MMA issue", "(producer) entry->loop end", "final drain", "epilogue", "tail sync"]
    for half, name in ((0, "metric GEMM"), (1, "quad GEMM")):
        s = st[half * 4096:(half + 1) * 4096]
        s = s[s[:, 0] > 0]
        if not len(s):
            continue
        t0 = s[:, 0].min()
        print(f"{name}: {len(s)} CTAs, kernel span {(s[:, 7].max() - t0) / 1e3:.1f} us")
        d = {"entry->setup": s[:, 1] - s[:, 0], "setup->first MMA": s[:, 2] - s[:, 1], "first->last MMA issue": s[:, 3] - s[:, 2],
             "last MMA issue->loop end (tid 0)": s[:, 4] - s[:, 3], "final drain": s[:, 5] - s[:, 4], "epilogue": s[:, 6] - s[:, 5],
             "tail sync": s[:, 7] - s[:, 6], "CTA total": s[:, 7] - s[:, 0]}
        if half:
            d.update({"  epi: masks + issue s copies": s[:, 8] - s[:, 5], "  epi: wait s": s[:, 9] - s[:, 8], "  epi: R": s[:, 10] - s[:, 9], "  epi: X^T R": s[:, 6] - s[:, 10]})
        for k, v in d.items():
            print(f"  {k:34s} mean {v.mean() / 1e3:8.2f} us   p10 {np.percentile(v, 10) / 1e3:8.2f}   p90 {np.percentile(v, 90) / 1e3:8.2f}")
        start = np.sort(s[:, 0] - t0) / 1e3
        print("  CTA start times (us) quantiles:", np.percentile(start, [0, 25, 50, 75, 100]).round(1))


if __name__ == "__main__":
    main()
