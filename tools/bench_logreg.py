"""Time rmhmc on the logistic-regression target through the product path (geomjax_b200.rmhmc -> gb200_step ->
lock-step rolling batch) and, for comparison, the CTA-per-chain kernels.

    python tools/bench_logreg.py N D C L eps T [path]
"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import geomjax_b200 as g  # noqa: E402
from bench.data import make_logreg_data  # noqa: E402


def main():
    N, D, C, L = (int(x) for x in sys.argv[1:5])
    eps, T = float(sys.argv[5]), int(sys.argv[6])
    path = sys.argv[7] if len(sys.argv) > 7 else "lockstep"
    dev = torch.device("cuda:0")
    X, y = make_logreg_data(N, D, seed=0)
    target = g.logistic_regression(torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev), 0.01)
    alg = g.rmhmc(target, eps, target, L, logreg_path=path)
    st = alg.init(torch.zeros((C, D), device=dev))
    root = g.random.PRNGKey(0)
    st, _, acc = g.run_fused(alg.step, root, st, 2, total=2 + T, return_accept=True)  # warm-up (2 transitions)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    st, _, acc = g.run_fused(alg.step, root, st, T, first=2, total=2 + T, return_accept=True)
    e1.record()
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    line = f"{path} N={N} D={D} C={C} L={L} eps={eps} T={T}: {ms:.1f} ms ({t_enq * 1e3:.1f} ms to enqueue), " \
           f"{C * L * T / ms * 1e3:.0f} chain-leapfrog-steps/s, accept {float(acc.mean()):.3f}"
    if path == "lockstep":
        pl = alg.step.engine.plan(C, dev)
        rounds, evals = pl.stats()
        line += f", {rounds} rounds, {evals} chain-evaluations ({evals / (C * L * T):.2f} per chain-step, " \
                f"{evals / ms * 1e3:.0f} /s), loop: {pl.loop_mode}"
    print(line, flush=True)


def split_experiment():
    """Two independent half-size chain sets on two CUDA streams (their graphs interleave on the GPU: one half's FP32
    kernels run under the other half's GEMMs) against one full-size launch.  python tools/bench_logreg.py split N D C L eps T"""
    N, D, C, L = (int(x) for x in sys.argv[2:6])
    eps, T = float(sys.argv[6]), int(sys.argv[7])
    parts = int(sys.argv[8]) if len(sys.argv) > 8 else 2
    dev = torch.device("cuda:0")
    X, y = make_logreg_data(N, D, seed=0)
    target = g.logistic_regression(torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev), 0.01)
    alg = g.rmhmc(target, eps, target, L)
    root = g.random.PRNGKey(0)
    Ch = C // parts
    streams = [torch.cuda.Stream(device=dev) for _ in range(parts)]
    states = [alg.init(torch.zeros((Ch, D), device=dev)) for _ in range(parts)]
    from geomjax_b200.plan import LockstepPlan
    plans = [LockstepPlan(target, Ch, dev) for _ in range(parts)]  # one plan (workspace + graph) per part
    from geomjax_b200 import _native as Nn

    def launch(i, first, Tn):
        eng = alg.step.engine
        ks = Nn.KeySource()
        ks.keys = None
        ks.root_key[0], ks.root_key[1] = int(root[0]), int(root[1])
        ks.first_transition, ks.num_transitions, ks.total_transitions = first, Tn, 2 + T
        ks.chain_offset, ks.total_chains = i * Ch, C
        opts = Nn.RunOpts()
        opts.plan = plans[i].handle
        import ctypes as Ct
        q = states[i][0]
        p, keep = eng._params(q)
        desc = eng.target.c_struct()
        st = Nn.State(*[Nn.ptr(t) for t in states[i]], None)
        with torch.cuda.stream(streams[i]):
            Nn.check(Nn.lib().gb200_step(eng.sampler, Ct.byref(p), Ct.byref(desc), Ct.byref(ks), st, st, None, Ct.byref(opts),
                                         Ch, Nn.stream_ptr()))

    for i in range(parts):
        launch(i, 0, 2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s_ in streams:
        s_.wait_stream(torch.cuda.current_stream())
    for i in range(parts):
        launch(i, 2, T)
    for s_ in streams:
        torch.cuda.current_stream().wait_stream(s_)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"split x{parts} N={N} D={D} C={C} L={L} T={T}: {ms:.1f} ms, {C * L * T / ms * 1e3:.0f} chain-leapfrog-steps/s", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "split":
        split_experiment()
    else:
        main()
