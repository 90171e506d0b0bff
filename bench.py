#!/usr/bin/env python
"""Headline benchmark: chain-leapfrog-steps/s (and min-ESS/s) of the fused transition kernels.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2] [--sub ...]

Workload at N=1 = BASELINE.json configs[1] ("c2"): Neal's funnel D=20, lmcmonge, 65,536 chains.
N>1 (torchrun, one rank per GPU): chains are sharded -- every rank owns 65,536 chains of one GLOBAL chain set
(keys derived from the global chain index), no data-path collective; weak scaling.

A "step" = ONE fused launch that advances every chain by `transitions_per_step` transitions (each = key derivation
+ draw + L integrator steps + MH accept), inputs resident in HBM; the default 2048 transitions per step make the
K = 20 timed steps of the driver last > 1 s.  `e2e` = the same work through the public API with HOST buffers:
pinned host positions -> H2D -> init -> fused transitions -> D2H of the final positions and of the per-chain mean
acceptance rate (reduced inside the kernels), all inside the timed region, every step.

The same JSON line carries, under `workloads`, sub-records for the logistic-regression half of the metric measured
in the same process (c4; c5_shard = one GPU's 16,384-chain share of c5), the as-written / omega_fixed pair for c2,
and, for N > 1, a `collectives` record (pooled-adaptation all-reduce, sharded R-hat / ESS reduction on c3's shape).

`--impl reference` times the CPU restatement of the reference semantics (oracle/cpp, C++ / OpenMP, all host cores;
the reference itself needs JAX, which this image does not have) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CONFIGS = json.load(open(os.path.join(ROOT, "bench", "configs.json")))
METRIC = "chain-leapfrog-steps/s"
TOTAL_TRANSITIONS = 1 << 20  # width of the outer split(root, T): fixed, so keys do not depend on K


# --------------------------------------------------------------------------- CPU legs (oracle/; checker only)
def _cpu_sampler(cfg, threads=0, half_step=None, step_size=None):
    """The C++/OpenMP restatement (oracle/cpp) configured for a workload."""
    from oracle import cpu
    D = cfg["D"]
    eps = step_size or cfg["step_size"]
    if cfg["target"] == "logreg":
        from bench.data import make_logreg_data
        X, y = make_logreg_data(cfg["N"], D, cfg.get("data_seed", 0))
        return cpu.CpuSampler("rmhmc", D, eps, cfg["num_integration_steps"], X=X, y=y,
                              prior_precision=cfg["prior_precision"], threads=threads)
    if cfg["sampler"] == "rmhmc":
        return None  # rmhmc on the funnel (c1_softabs) is restated in NumPy only
    return cpu.CpuSampler(cfg["sampler"], D, eps, cfg["num_integration_steps"], sigma=cfg.get("sigma", 3.0),
                          alpha2=cfg.get("alpha2", 1e-3), half_step=half_step or cfg.get("half_step", "omega"),
                          inverse_mass_matrix=np.ones(D, np.float32), threads=threads)


def host_threads():
    """Every host core this process may run on (torchrun exports OMP_NUM_THREADS=1: the OpenMP default is not it)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def cpu_throughput(cfg, chains, transitions, threads=0, first=0, **kw):
    """chain-leapfrog-steps/s of the C++ restatement on `threads` OpenMP threads (0 = all host cores);
    (value, seconds, threads, accept)."""
    from oracle import prng as P
    smp = _cpu_sampler(cfg, threads or host_threads(), **kw)
    q0 = (np.ones if cfg["init_position"] == "ones" else np.zeros)((chains, cfg["D"]), np.float32)
    st = smp.init(q0)
    t0 = time.perf_counter()
    acc = smp.run(P.key(cfg["root_key"]), st, transitions, first=first, total=TOTAL_TRANSITIONS, total_chains=chains)
    dt = time.perf_counter() - t0
    return chains * cfg["num_integration_steps"] * transitions / dt, dt, smp.threads, acc


def numpy_throughput(cfg, chains, transitions):
    """The NumPy oracle (single process): kept as a second field beside the C++ figure."""
    from oracle import prng as P, samplers as S, targets as Tg
    D = cfg["D"]
    tgt = Tg.NealFunnel(D, cfg.get("sigma", 3.0))
    if cfg.get("metric") == "softabs":
        tgt = Tg.softabs_metric(tgt, cfg["softabs_alpha"])
    q0 = (np.ones if cfg["init_position"] == "ones" else np.zeros)((chains, D), np.float32)
    if cfg["sampler"] == "lmcmonge":
        st = S.lmcmonge_init(q0, tgt)
        step = lambda k, s: S.lmcmonge_step(k, s, tgt, cfg["step_size"], np.ones(D, np.float32),
                                            cfg["num_integration_steps"], alpha2=cfg["alpha2"], half_step=cfg["half_step"])
    elif cfg["sampler"] == "lmc":
        st = S.lmc_init(q0, tgt)
        step = lambda k, s: S.lmc_step(k, s, tgt, cfg["step_size"], cfg["num_integration_steps"])
    else:
        st = S.rmhmc_init(q0, tgt)
        step = lambda k, s: S.rmhmc_step(k, s, tgt, cfg["step_size"], cfg["num_integration_steps"])
    root = P.key(cfg["root_key"])
    t0 = time.perf_counter()
    for t in range(transitions):
        st, _ = step(S.chain_keys(root, TOTAL_TRANSITIONS, t, chains), st)
    dt = time.perf_counter() - t0
    return chains * cfg["num_integration_steps"] * transitions / dt, dt


def _cpu_sample_size(cfg, cores):
    """(chains, transitions) of one bounded CPU step: a few seconds of work on `cores` threads."""
    if cfg["target"] == "logreg":
        return (2 * cores, 1) if cfg["D"] > 32 else (16 * cores, 2)
    return (4096 * cores, 8) if cfg["D"] <= 32 else (2048 * cores, 8)


def run_reference(args, cfg):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = host_threads()
    L = cfg["num_integration_steps"]
    if _cpu_sampler(cfg) is None:  # funnel rmhmc: NumPy port only
        chains, tr = cfg["chains_per_gpu"], 50
        run = lambda: numpy_throughput(cfg, chains, tr)[0]
        kind_note, cores = "NumPy oracle port, 1 process", 1
    else:
        chains, tr = _cpu_sample_size(cfg, cores)
        v, dt, _, _ = cpu_throughput(cfg, chains, 1)
        tr = max(1, min(64, int(tr * 3.0 / max(dt * tr, 1e-3))))  # ~3 s per step
        run = lambda: cpu_throughput(cfg, chains, tr)[0]
        kind_note = "C++/OpenMP restatement of the reference semantics (oracle/cpp), closed-form funnel algebra / dense logreg"
    for _ in range(min(args.warmup, 2)):
        run()
    t0 = time.perf_counter()
    vals = [run() for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    value = float(np.mean(vals))
    sample = f"{chains} chains x {tr} transitions x L={L} per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "chain-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["name"], "sampler": cfg["sampler"], "D": cfg["D"],
                   "num_integration_steps": L, "step_size": cfg["step_size"]},
        "cpu_baseline": {"value": value, "unit": "chain-steps/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": kind_note + " (not JAX: jax/jaxlib are not installable in this image)"},
        "e2e": {"value": value, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- clocks sampler
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def bind_rank_to_cores(local, local_world):
    """Disjoint host-core set per rank (all ranks default to the same affinity mask): the e2e arm's host work
    (pinned-buffer copies, result reads) of one rank no longer competes with the others'."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // max(local_world, 1))
        mine = cores[local * per:(local + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        return len(mine)
    except Exception:
        return None


# --------------------------------------------------------------------------- our arm
class Ctx:
    pass


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "MEASURED_PEAKS.json"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def make_target(g, torch, cfg, dev):
    if cfg["target"] == "logreg":
        from bench.data import make_logreg_data  # frozen synthetic design (input data only)
        Xh, yh = make_logreg_data(cfg["N"], cfg["D"], cfg.get("data_seed", 0))
        return g.logistic_regression(torch.from_numpy(Xh).to(dev), torch.from_numpy(yh).to(dev), cfg["prior_precision"])
    return g.neal_funnel(cfg["D"], sigma=cfg.get("sigma", 3.0))


def make_alg(g, torch, N, cfg, target, dev, *, half_step=None, step_size=None, lanes=0):
    D, L = cfg["D"], cfg["num_integration_steps"]
    if cfg["sampler"] == "lmcmonge":
        hs = half_step or cfg["half_step"]
        integ = {"omega": g.integrators.half_step_omega, "omega_fixed": g.integrators.half_step_omega_fixed,
                 "omegatilde": g.integrators.half_step_omegatilde}[hs]
        eps = step_size or (cfg["step_size_omega_fixed"] if hs == "omega_fixed" else cfg["step_size"])
        return g.lmcmonge(target, eps, torch.ones(D, device=dev), L, alpha2=cfg["alpha2"], integrator=integ,
                          lanes_per_chain=lanes), eps, N.LMCMONGE
    eps = step_size or cfg["step_size"]
    if cfg["sampler"] == "lmc":
        return g.lmc(target, eps, target, L, lanes_per_chain=lanes), eps, N.LMC
    metric = g.softabs(target, cfg["softabs_alpha"]) if cfg.get("metric") == "softabs" else target
    return g.rmhmc(target, eps, metric, L, lanes_per_chain=lanes), eps, N.RMHMC


def timed_fused(cx, alg, state, C, TPS, K, W, t0_idx=0):
    """W warm-up + K timed fused launches (device-resident); returns (state, per-launch ms, first->last ms, launches)."""
    g, torch, N = cx.g, cx.torch, cx.N

    def fused(st, first):
        s, _, _ = g.run_fused(alg.step, cx.root, st, TPS, first=first, total=TOTAL_TRANSITIONS,
                              chain_offset=cx.rank * C, total_chains=C * cx.world, inplace=True)
        return s

    t_idx = t0_idx
    for _ in range(W):
        state = fused(state, t_idx)
        t_idx += TPS
    cx.barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    l0 = N.lib().gb200_kernel_launches()
    for k in range(K):
        evs[k][0].record()
        state = fused(state, t_idx)
        evs[k][1].record()
        t_idx += TPS
    launches = N.lib().gb200_kernel_launches() - l0
    cx.barrier()
    per = [a.elapsed_time(b) for a, b in evs]
    return state, per, cx.max_over_ranks(evs[0][0].elapsed_time(evs[-1][1])), int(launches), t_idx


def timed_e2e(cx, alg, C, D, TPS, K, W, start_position, min_warm=2):
    """Host buffers -> H2D -> init -> fused transitions -> D2H (positions + per-chain mean acceptance), every step,
    double-buffered over two streams; returns (ms max over ranks, mean acceptance, h2d bytes, d2h bytes)."""
    g, torch = cx.g, cx.torch
    dev = cx.dev
    host_q = start_position.detach().to("cpu").pin_memory()  # every step starts from the same (warmed-up) positions
    streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
    host_out = [torch.empty((C, D)).pin_memory() for _ in streams]
    host_acc = [torch.empty((C,)).pin_memory() for _ in streams]
    dev_acc = [torch.empty((C,), device=dev) for _ in streams]
    done = [torch.cuda.Event() for _ in streams]

    def enqueue(i, first):
        with torch.cuda.stream(streams[i]):
            st = alg.init(host_q.to(dev, non_blocking=True))
            s, _, acc = g.run_fused(alg.step, cx.root, st, TPS, first=first, total=TOTAL_TRANSITIONS,
                                    chain_offset=cx.rank * C, total_chains=C * cx.world, return_accept="mean",
                                    out_accept=dev_acc[i], inplace=True)
            host_out[i].copy_(s.position, non_blocking=True)
            host_acc[i].copy_(acc, non_blocking=True)
            done[i].record(streams[i])

    def collect(i):
        done[i].synchronize()
        return float(host_acc[i].mean())  # the step's result is read on the host

    def run(n):
        for k in range(n):
            enqueue(k % 2, k * TPS)
            if k > 0:
                collect((k - 1) % 2)
        return collect((n - 1) % 2)

    run(max(W, min_warm))
    cx.barrier()
    cur = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s_ in streams:
        s_.wait_stream(cur)
    acc = run(K)
    for s_ in streams:
        cur.wait_stream(s_)
    e1.record()
    cx.barrier()
    return cx.max_over_ranks(e0.elapsed_time(e1)), acc, C * D * 4, C * D * 4 + C * 4


def ess_record(cx, alg, C, D, Tn, burnin, init_fill, thin=1):
    """A sampling run that keeps every (thin-th) sample on the device (buffer allocated BEFORE the timed region), then
    the sharded R-hat / ESS.  min_ess_per_s is only reported when the chains have mixed (max R-hat < 1.01).  thin > 1:
    one fused launch of `thin` transitions per recorded sample; the ESS of the thinned chain is a lower bound of the
    full chain's."""
    g, torch = cx.g, cx.torch
    Tn = max(16, min(Tn, int(24e9 // (C * D * 4))))  # the sample tensor stays <= 24 GB (c3: 131,072 x 100 per GPU)
    st = alg.init(init_fill((C, D), device=cx.dev))
    samples = torch.empty((Tn, C, D), device=cx.dev)
    acc = torch.empty((C,), device=cx.dev)
    total = burnin + Tn * thin
    kw = dict(total=total, chain_offset=cx.rank * C, total_chains=C * cx.world)
    if burnin > 0:
        st, _, _ = g.run_fused(alg.step, cx.root, st, burnin, first=0, inplace=True, **kw)
    cx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if thin == 1:
        st, _, _ = g.run_fused(alg.step, cx.root, st, Tn, first=burnin, out_samples=samples, return_accept="mean",
                               out_accept=acc, **kw)
    else:
        for t in range(Tn):
            st, _, _ = g.run_fused(alg.step, cx.root, st, thin, first=burnin + t * thin, inplace=True,
                                   return_accept="mean", out_accept=acc, **kw)
            samples[t].copy_(st.position)
    e1.record()
    torch.cuda.synchronize()
    samp_ms = cx.max_over_ranks(e0.elapsed_time(e1))
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d0.record()
    rhat = g.rhat(samples, chain_axis=1, sample_axis=0)
    ess = g.ess(samples, chain_axis=1, sample_axis=0)
    d1.record()
    torch.cuda.synchronize()
    max_rhat, min_ess = float(rhat.max()), float(ess.min())
    rec = {"max_rhat": max_rhat, "min_ess": min_ess, "samples_per_chain": Tn, "thinning": thin, "burnin": burnin,
           "sampling_ms": samp_ms, "diagnostics_ms": d0.elapsed_time(d1), "mean_acceptance": float(acc.mean()),
           "min_ess_valid": bool(max_rhat < 1.01)}
    if rec["min_ess_valid"]:
        rec["min_ess_per_s"] = min_ess / (samp_ms * 1e-3)
    else:
        rec["min_ess_invalid_reason"] = f"max R-hat {max_rhat:.4f} >= 1.01 after {burnin}+{Tn * thin} transitions: chains have not mixed"
    del samples
    return rec


def streaming_record(cx, alg, C, D, Tn, init_fill, block=32, max_lags=64):
    """N2: Tn transitions in fused launches of `block` whose sample buffer is reused and folded into the streaming
    R-hat / ESS accumulators -- no (T, C, D) tensor (c3 at T = 1000 would need 52 GB per GPU); statistics all-reduced
    over ranks exactly like the batch diagnostics."""
    g, torch = cx.g, cx.torch
    st = alg.init(init_fill((C, D), device=cx.dev))
    cx.barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    st, diag, acc = g.sample_streaming(alg.step, cx.root, st, Tn, block=block, max_lags=max_lags, chain_offset=cx.rank * C,
                                       total_chains=C * cx.world)
    e1.record()
    rhat = diag.rhat()
    ess = diag.ess(allow_truncated=True)
    e2.record()
    torch.cuda.synchronize()
    run_ms = cx.max_over_ranks(e0.elapsed_time(e1))
    max_rhat, min_ess = float(rhat.max()), float(ess.min())
    rec = {"samples_per_chain": Tn, "block": block, "max_lags": max_lags, "sampling_plus_update_ms": run_ms,
           "finalize_ms": cx.max_over_ranks(e1.elapsed_time(e2)), "max_rhat": max_rhat, "min_ess": min_ess,
           "lags_truncated_dims": int(diag.truncated.sum()), "mean_acceptance": float(acc.mean()),
           "accumulator_bytes": int(diag._ws.numel()), "sample_tensor_bytes_avoided": int(Tn) * C * D * 4,
           "allreduce_bytes": (3 * D + 1) * 8 + max_lags * D * 8, "min_ess_valid": bool(max_rhat < 1.01)}
    if rec["min_ess_valid"]:
        rec["min_ess_per_s"] = min_ess / (run_ms * 1e-3)
    else:
        rec["min_ess_invalid_reason"] = f"max R-hat {max_rhat:.4f} >= 1.01 after {Tn} transitions"
    return rec


def fp32_peak(cx):
    """FFMA issue rate of this GPU, same process, same clocks (256 FFMA per loop trip, 16 chains per thread)."""
    torch, N = cx.torch, cx.N
    out = torch.zeros(1, device=cx.dev)
    iters, grid, block = 1 << 18, 148 * 8, 256
    for _ in range(2):
        N.check(N.lib().gb200_fp32_peak_kernel(N.ptr(out), grid, block, iters, N.stream_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    N.check(N.lib().gb200_fp32_peak_kernel(N.ptr(out), grid, block, iters, N.stream_ptr()))
    e1.record()
    torch.cuda.synchronize()
    return 2.0 * grid * block * iters / (e0.elapsed_time(e1) * 1e-3) / 1e12


def funnel_flops(cx, cfg, alg, sampler_id, target, state, C):
    """Algorithmic flops per chain-leapfrog-step: the integrator step + the per-transition
    work amortised over L steps; rmhmc: (2 + measured fixed-point iterations) map evaluations per step."""
    N, g = cx.N, cx.g
    L, D = cfg["num_integration_steps"], cfg["D"]
    fstep = N.lib().gb200_flops_per_chain_step(sampler_id, target.c_struct())
    ftrans = N.lib().gb200_flops_per_transition(sampler_id, target.c_struct())
    rec = {"flops_per_integrator_step": fstep, "flops_per_transition_outside_steps": ftrans}
    unit = fstep + ftrans / L
    if sampler_id == N.RMHMC:
        ks = N.KeySource()
        kk = g.random.chain_keys(cx.root, 0, TOTAL_TRANSITIONS, C, chain_offset=cx.rank * C, total_chains=C * cx.world)
        ks.keys, ks.num_transitions = N.ptr(kk), 1
        _, inf = alg.step.engine.launch(state, ks, want_info=True, extra_info=True)
        it = float(inf["fp_iters"].float().mean()) / L
        rec["fp_iters_per_step"] = it
        unit = (2.0 + it) * (21.0 * D + 50.0) + ftrans / L
    rec["flops_per_chain_step"] = unit
    return unit, rec


def bench_funnel(cx, args, cfg, wl_key, *, half_step=None, step_size=None, C=None, TPS=None, K=None, W=None,
                 with_e2e=True, ess_samples=0, burnin=0, thin=1):
    """Device-resident + e2e + roofline (+ R-hat / ESS) for one funnel workload."""
    g, torch, N = cx.g, cx.torch, cx.N
    D, L = cfg["D"], cfg["num_integration_steps"]
    C = C or args.chains or cfg["chains_per_gpu"]
    TPS, K, W = TPS or args.transitions_per_step, K or args.steps, W or args.warmup
    target = make_target(g, torch, cfg, cx.dev)
    alg, eps, sid = make_alg(g, torch, N, cfg, target, cx.dev, half_step=half_step, step_size=step_size,
                             lanes=args.lanes_per_chain)
    init_fill = torch.ones if cfg["init_position"] == "ones" else torch.zeros
    state = alg.init(init_fill((C, D), device=cx.dev))
    state, per_ms, dev_ms, launches, _ = timed_fused(cx, alg, state, C, TPS, K, W)
    total_chains = C * cx.world
    value = total_chains * L * TPS * K / (dev_ms * 1e-3)
    rec = {"value": value, "unit": "chain-steps/s", "ms_per_step": dev_ms / K, "timed_region_s": dev_ms * 1e-3,
           "gpu_launches": launches,
           "config": {"workload": cfg["name"], "sampler": cfg["sampler"], "D": D, "chains_per_gpu": C,
                      "total_chains": total_chains, "num_integration_steps": L, "transitions_per_step": TPS,
                      "step_size": eps, "half_step": half_step or cfg.get("half_step"),
                      "lanes_per_chain": args.lanes_per_chain,
                      "l2_policy": "no flush needed: per-launch HBM traffic is the chain state only (compute-bound "
                                   "kernel; the state re-read per transition is L1/L2 resident by design)",
                      "parallelism": f"chains sharded over {cx.world} GPU(s), no data-path collective"}}
    if with_e2e:
        e2e_ms, acc, h2d, d2h = timed_e2e(cx, alg, C, D, TPS, K, W, state.position)
        rec["e2e"] = {"value": total_chains * L * TPS * K / (e2e_ms * 1e-3), "unit": "chain-steps/s",
                      "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "mean_acceptance": acc,
                      "timed_region_s": e2e_ms * 1e-3,
                      "pipelining": "2 CUDA streams, double-buffered pinned host buffers; acceptance reduced in-kernel; "
                                    "the fused launches of consecutive steps overlap on the GPU (one launch is a single "
                                    "wave of ~28 warps/SM), which is why e2e can exceed the one-launch-at-a-time value"}
    if cx.rank == 0:
        unit, frec = funnel_flops(cx, cfg, alg, sid, target, state, C)
        med = float(np.median(per_ms))
        ach = unit * C * L * TPS / (med * 1e-3) / 1e12
        clk = cx.clocks.summary()["sm_mhz"] if cx.clocks else None
        nominal = 148 * 128 * 2 * (clk or 1965.0) * 1e6 / 1e12
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(wl_key, {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        rec["roofline"] = dict(
            bound="fp32", achieved=ach, peak=cx.fp32_peak, unit="TFLOP/s", frac=ach / cx.fp32_peak, traffic=traffic,
            traffic_note="dram bytes of a 16-transition launch (ncu --set full capture, profiles/traffic.json): the chain "
                         "state read once; HBM is idle on this path",
            peak_source="measured in this process: FFMA microbenchmark (148x8 CTAs x 256 threads, 16 independent "
                        "chains, 256 FFMA per loop trip)",
            nominal_peak=nominal, frac_of_nominal=ach / nominal,
            nominal_source=f"148 SMs x 128 lanes x 2 x {clk or 1965.0:.0f} MHz (median SM clock of this run)",
            kernel_ms_median=med, **frec)
        # issue-slot view: the bit-exact threefry stream is mandatory integer work the flop count leaves out
        # (75 integer ops per block; per chain-transition 6 key-tree blocks + 1 accept uniform + D / 2 normal blocks)
        int_ops = 75.0 * (7 + (D + 1) // 2)
        slots = (unit * L / 2.0 + int_ops) * C * TPS / (med * 1e-3)
        slot_peak = 148 * 128 * (clk or 1965.0) * 1e6
        rec["roofline"]["issue_slots"] = {"mandatory_thread_instructions_per_chain_transition": unit * L / 2.0 + int_ops,
                                          "of_which_integer_prng": int_ops, "achieved_per_s": slots, "peak_per_s": slot_peak,
                                          "frac": slots / slot_peak,
                                          "note": "FMA-equivalent slots (flops / 2) + threefry integer ops against 148 SMs x 128 lanes x clock"}
    if ess_samples > 0:
        rec.update(ess_record(cx, alg, C, D, ess_samples, burnin, init_fill, thin))
    return rec


def logreg_flops_per_eval(cfg):
    """Algorithmic flops of ONE evaluation of the implicit-midpoint map for one chain: the two D^2 N
    products (metric vec(G) = Z^T w and the quadratic forms h = Z vecsym(A), 2 N P each, P = D(D+1)/2), the four
    O(N D) products (eta, X^T r, u = X v folded into A, X^T t: 2 N D each) and the O(D^3) factorisation + inverse."""
    Nr, D = cfg["N"], cfg["D"]
    P = D * (D + 1) // 2
    gemm = 2 * (2.0 * Nr * P)
    return gemm + 3 * 2.0 * Nr * D + D ** 3, gemm


def bench_logreg(cx, args, cfg, *, C, T, K, W, label, ess_samples=0, burnin=0, adapt_transitions=0):
    """rmhmc on logistic regression through geomjax_b200.rmhmc (lock-step rolling batch on the tcgen05 GEMMs):
    device-resident value, e2e, fixed-point histogram, tensor-pipe roofline."""
    g, torch, N = cx.g, cx.torch, cx.N
    D, L = cfg["D"], cfg["num_integration_steps"]
    target = make_target(g, torch, cfg, cx.dev)
    alg, eps, sid = make_alg(g, torch, N, cfg, target, cx.dev)
    state = alg.init(torch.zeros((C, D), device=cx.dev))
    plan = alg.step.engine.plan(C, cx.dev)
    state, per_ms, dev_ms, launches, t_idx = timed_fused(cx, alg, state, C, T, K, W)
    rounds, evals = plan.stats()  # of the last launch
    total_chains = C * cx.world
    value = total_chains * L * T * K / (dev_ms * 1e-3)
    e2e_ms, acc, h2d, d2h = timed_e2e(cx, alg, C, D, T, max(K // 2, 1), 1, state.position, min_warm=1)
    Ke = max(K // 2, 1)
    rec = {"value": value, "unit": "chain-steps/s", "ms_per_step": dev_ms / K, "timed_region_s": dev_ms * 1e-3,
           "steps": K, "warmup": W, "gpu_launches": launches,
           "e2e": {"value": total_chains * L * T * Ke / (e2e_ms * 1e-3), "unit": "chain-steps/s",
                   "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "mean_acceptance": acc, "steps": Ke},
           "config": {"workload": label, "sampler": "rmhmc", "D": D, "N": cfg["N"], "chains_per_gpu": C,
                      "total_chains": total_chains, "num_integration_steps": L, "transitions_per_step": T,
                      "step_size": eps, "loop": plan.loop_mode,
                      "l2_policy": "working set per round (W / B operand tiles, s, packed G, partials) exceeds L2 "
                                   "for C >= 4096 at D=100; no flush"}}
    if cx.rank == 0:
        # fixed-point iteration histogram of one transition (Info.fp_iters, summed over the L steps)
        ks = N.KeySource()
        kk = g.random.chain_keys(cx.root, t_idx, TOTAL_TRANSITIONS, C, chain_offset=cx.rank * C, total_chains=total_chains)
        ks.keys, ks.num_transitions = N.ptr(kk), 1
        _, inf = alg.step.engine.launch(state, ks, want_info=True, extra_info=True)
        it = inf["fp_iters"].float() / L
        qs = torch.quantile(it, torch.tensor([0.5, 0.9, 0.99, 1.0], device=it.device)).tolist()
        rec["fp_iters_per_step"] = {"mean": float(it.mean()), "p50": qs[0], "p90": qs[1], "p99": qs[2], "max": qs[3]}
        f_eval, f_gemm = logreg_flops_per_eval(cfg)
        last_ms = per_ms[-1]
        peaks, src = _peaks()
        peak = peaks["bf16_tflops_sustained"] / 2.0 / 3.0
        ach = f_eval * evals / (last_ms * 1e-3) / 1e12
        rec["chain_evaluations_per_s"] = evals / (last_ms * 1e-3)
        rec["evaluations_per_chain_step"] = evals / (C * L * T)
        rec["rounds_last_launch"] = rounds
        rec["kernels_in_graph_last_launch"] = 9 * rounds  # gpu_launches counts the graph launch once; a round is 9 kernels
        rec["roofline"] = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                           "traffic": None, "flops_per_chain_evaluation": f_eval, "gemm_flops_per_chain_evaluation": f_gemm,
                           "peak_source": f"tensor_3xtf32 = bf16_tflops_sustained ({src}) / 2 (tf32) / 3 (three MMAs per "
                                          "product: error-compensated split); algorithmic flops of every map evaluation "
                                          "of the launch / launch time, ALL kernels of the round included"}
    if ess_samples > 0:
        rec.update(ess_record(cx, alg, C, D, ess_samples, burnin, torch.zeros))
    if adapt_transitions > 0:
        # c5's warm-up: per-chain dual averaging fused into the transition's epilogue (step_size_adaptation), timed here
        # for `adapt_transitions` transitions from the initial position (adaptation/step_size_adaptation.py:143-201)
        cx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res, ainfo = g.step_size_adaptation(g.rmhmc, target, initial_step_size=eps, metric_fn=target,
                                            num_integration_steps=L).run(g.random.chain_keys(cx.root, 0, 1, C, chain_offset=cx.rank * C, total_chains=total_chains),
                                                                         torch.zeros((C, D), device=cx.dev), adapt_transitions)
        e1.record()
        torch.cuda.synchronize()
        ss = res.parameters["step_size"]
        rec["step_size_adaptation"] = {"transitions": adapt_transitions,
                                       "ms_per_transition": cx.max_over_ranks(e0.elapsed_time(e1)) / adapt_transitions,
                                       "mean_acceptance": float(ainfo["acceptance_rate"].mean()),
                                       "step_size_after": [float(ss.min()), float(ss.median()), float(ss.max())],
                                       "note": "per-chain dual averaging in the kernel epilogue; initial step size = the config's"}
    return rec


def collectives_record(cx, args):
    """The only NCCL traffic of the design, timed inside this run: (1) the pooled step-size adaptation all-reduce
    (2 floats per warm-up transition), (2) the sharded R-hat / ESS reduction on c3's shape."""
    g, torch = cx.g, cx.torch
    import torch.distributed as dist
    rec = {"world": cx.world}
    s = torch.ones(2, device=cx.dev)
    for _ in range(20):
        dist.all_reduce(s)
    cx.barrier()
    n = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        dist.all_reduce(s)
    e1.record()
    torch.cuda.synchronize()
    rec["pooled_adaptation_allreduce"] = {"us_per_call": cx.max_over_ranks(e0.elapsed_time(e1)) / n * 1e3, "calls": n,
                                         "bytes": 8, "note": "one call per warm-up transition (adaptation.py::_mean_accept)"}
    cfg = CONFIGS["c3"]
    C, D, Tn = cfg["chains_per_gpu"], cfg["D"], 64
    target = make_target(g, torch, cfg, cx.dev)
    alg, _, _ = make_alg(g, torch, cx.N, cfg, target, cx.dev)
    st = alg.init(torch.ones((C, D), device=cx.dev))
    samples = torch.empty((Tn, C, D), device=cx.dev)
    g.run_fused(alg.step, cx.root, st, Tn, total=Tn, chain_offset=cx.rank * C, total_chains=C * cx.world, out_samples=samples)
    g.rhat(samples, chain_axis=1, sample_axis=0)
    cx.barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    rhat = g.rhat(samples, chain_axis=1, sample_axis=0)
    e1.record()
    ess = g.ess(samples, chain_axis=1, sample_axis=0)
    e2.record()
    torch.cuda.synchronize()
    rec["sharded_diagnostics_c3_shape"] = {
        "chains_per_gpu": C, "D": D, "samples_per_chain": Tn,
        "rhat_ms": cx.max_over_ranks(e0.elapsed_time(e1)), "ess_ms": cx.max_over_ranks(e1.elapsed_time(e2)),
        "rhat_allreduce_bytes": (3 * D + 1) * 8, "ess_allreduce_bytes_per_round": 64 * D * 8,
        "max_rhat": float(rhat.max()), "min_ess": float(ess.min()),
        "note": "per-rank partial sums over the local chains + ONE all-reduce(sum) of the sufficient statistics "
                "(float64); Geyer truncation redundantly per rank"}
    del samples
    return rec


def run_ours(args, cfg):
    import torch
    import geomjax_b200 as g
    from geomjax_b200 import _native as N

    cx = Ctx()
    cx.g, cx.torch, cx.N = g, torch, N
    cx.rank = int(os.environ.get("RANK", "0"))
    cx.world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    host_cores = bind_rank_to_cores(local, int(os.environ.get("LOCAL_WORLD_SIZE", cx.world)))
    torch.cuda.set_device(local)
    cx.dev = torch.device("cuda", local)
    dist = None
    if cx.world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=cx.dev)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=cx.dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    cx.barrier, cx.max_over_ranks = barrier, max_over_ranks
    cx.root = g.random.PRNGKey(cfg["root_key"])
    cx.clocks = None
    cx.fp32_peak = fp32_peak(cx)
    subs = [] if args.sub == "none" else [s for s in args.sub.split(",") if s]
    t_wall = time.perf_counter()

    # ---- the headline workload
    cx.clocks = ClockSampler(local)
    cx.clocks.start()
    if cfg["target"] == "logreg":
        C = args.chains or cfg["chains_per_gpu"]
        T = args.transitions_per_step if args.transitions_per_step != 2048 else (16 if cfg["D"] <= 32 else 2)
        main = bench_logreg(cx, args, cfg, C=C, T=T, K=args.steps, W=args.warmup, label=cfg["name"])
    else:
        main = bench_funnel(cx, args, cfg, args.workload, half_step=args.half_step, step_size=args.step_size or None,
                            ess_samples=args.ess_samples, burnin=args.ess_burnin)
    cx.clocks.stop_flag = True
    clocks = cx.clocks.summary()

    # ---- sub-records: the other half of the metric, measured in the same process
    workloads = {}
    for s in subs:
        try:
            if s == "c2_omega_fixed" and args.workload == "c2":
                workloads[s] = bench_funnel(cx, args, CONFIGS["c2"], "c2", half_step="omega_fixed", TPS=256, K=5, W=3,
                                            with_e2e=False, ess_samples=4 * args.ess_samples, burnin=16384, thin=64)
                workloads[s]["note"] = ("c2 with alpha2 restored on the Christoffel correction (half_step_omega_fixed, eps "
                                        "from the same warm-up recipe): the variant that can mix; the headline runs the "
                                        "reference AS WRITTEN (lmcmonge/integrators.py:186-189, SURVEY F8), which only "
                                        "accepts at eps ~ 1e-3 and therefore cannot mix in 1000 transitions.  The nearly "
                                        "Euclidean Monge metric (alpha2 = 1e-3) decorrelates the funnel's v over ~2,600 "
                                        "transitions: this record runs 262,144 transitions per chain and keeps every 64th sample "
                                        "(R-hat(v) 1.02); four times as many transitions made it WORSE (1.11: chains that enter "
                                        "the funnel's neck stay for very long), so no run length passes the 1.01 gate for this "
                                        "sampler on this target -- min-ESS/s stays flagged invalid")
            elif s in ("c1", "c1_softabs"):
                # the reference's own CPU-runnable case: examples/funnel as shipped / BASELINE configs[0], 1000 samples
                workloads[s] = bench_funnel(cx, args, CONFIGS[s], s, TPS=1000, K=3, W=3, with_e2e=False, ess_samples=1000, burnin=0)
            elif s == "c3_shard":
                c3 = CONFIGS["c3"]
                rec = bench_funnel(cx, args, c3, "c3", TPS=64, K=5, W=3, with_e2e=False)
                tg3 = make_target(g, torch, c3, cx.dev)
                alg3, _, _ = make_alg(g, torch, N, c3, tg3, cx.dev)
                rec["streaming_diagnostics"] = streaming_record(cx, alg3, c3["chains_per_gpu"], c3["D"], 1000, torch.ones)
                workloads[s] = rec
            elif s == "c4":
                workloads[s] = bench_logreg(cx, args, CONFIGS["c4"], C=CONFIGS["c4"]["chains_per_gpu"], T=16, K=3, W=1,
                                            label=CONFIGS["c4"]["name"], ess_samples=min(args.ess_samples, 640), burnin=40)
            elif s == "c5_shard":
                c5 = CONFIGS["c5_shard"]
                workloads[s] = bench_logreg(cx, args, c5, C=c5["chains_per_gpu"], T=2, K=2, W=1, label=c5["name"],
                                            adapt_transitions=2)
        except Exception as e:  # a sub-record must never take the headline line down
            workloads[s] = {"error": f"{type(e).__name__}: {e}"[:300]}
    coll = None
    if cx.world > 1 and args.collectives:
        try:
            coll = collectives_record(cx, args)
        except Exception as e:
            coll = {"error": f"{type(e).__name__}: {e}"[:300]}

    if cx.rank == 0:
        line = {"metric": METRIC, "value": main["value"], "unit": "chain-steps/s", "n_gpus": cx.world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
        line.update({k: v for k, v in main.items() if k not in ("value", "unit", "ms_per_step", "steps", "warmup")})
        line["clocks"] = clocks
        line["host_cores_bound_per_rank"] = host_cores
        if workloads:
            line["workloads"] = workloads
        if coll is not None:
            line["collectives"] = coll
        if not args.no_cpu_baseline and cx.world == 1:
            line["cpu_baseline"] = cpu_baseline_record(cfg)
        line["wall_s"] = time.perf_counter() - t_wall
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline_record(cfg):
    """The C++/OpenMP restatement on the box's host cores, bounded sample (~10-20 s); NumPy figure as a second field."""
    cores = host_threads()
    L = cfg["num_integration_steps"]
    if _cpu_sampler(cfg) is None:
        v, w = numpy_throughput(cfg, cfg["chains_per_gpu"], 200)
        return {"value": v, "unit": "chain-steps/s", "cores": 1, "kind": "port",
                "sample": f"{cfg['chains_per_gpu']} chains x 200 transitions x L={L} (NumPy oracle, 1 process, {w:.1f} s)"}
    chains, tr = _cpu_sample_size(cfg, cores)
    v1, d1, _, _ = cpu_throughput(cfg, chains, 1)
    tr = max(1, min(256, int(10.0 / max(d1, 1e-3))))
    v, d, thr, acc = cpu_throughput(cfg, chains, tr, first=1)
    rec = {"value": v, "unit": "chain-steps/s", "cores": thr, "kind": "port",
           "sample": f"{chains} chains x {tr} transitions x L={L} (C++/OpenMP restatement oracle/cpp, {thr} threads, {d:.1f} s)",
           "mean_acceptance": acc, "host_cores_available": os.cpu_count()}
    if cfg["target"] != "logreg":
        vn, wn = numpy_throughput(cfg, 2048, 8)
        rec["numpy_port"] = {"value": vn, "cores": 1, "sample": f"2048 chains x 8 transitions (NumPy oracle, 1 process, {wn:.1f} s)"}
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(k for k in CONFIGS if not k.startswith("_")))
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default: the workload's)")
    ap.add_argument("--transitions-per-step", type=int, default=2048)
    ap.add_argument("--lanes-per-chain", type=int, default=0)
    ap.add_argument("--half-step", default=None)
    ap.add_argument("--step-size", type=float, default=0.0)
    ap.add_argument("--ess-samples", type=int, default=1000)
    ap.add_argument("--ess-burnin", type=int, default=200)
    ap.add_argument("--sub", default="default", help="comma list of sub-records (c1,c1_softabs,c2_omega_fixed,c3_shard,c4,c5_shard) or none")
    ap.add_argument("--no-collectives", dest="collectives", action="store_false")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.sub == "default":
        args.sub = "c1,c1_softabs,c2_omega_fixed,c3_shard,c4,c5_shard" if args.workload == "c2" else "none"
    cfg = CONFIGS[args.workload]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
