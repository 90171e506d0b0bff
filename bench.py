#!/usr/bin/env python
"""Headline benchmark: chain-leapfrog-steps/s (and min-ESS/s) of the fused transition kernels.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]

Workload at N=1 = BASELINE.json configs[1] ("c2"): Neal's funnel D=20, lmcmonge, 65,536 chains.
N>1 (torchrun, one rank per GPU): chains are sharded -- every rank owns 65,536 chains of one
GLOBAL chain set (keys derived from the global chain index), no data-path collective; weak scaling.

A "step" = ONE fused launch that advances every chain by `transitions_per_step` transitions
(each = key derivation + velocity draw + L integrator steps + MH accept), inputs resident in HBM.
`e2e` = the same work through the public API with HOST buffers: pinned host positions -> H2D ->
init -> fused transitions -> D2H of the final positions and acceptance rates, all inside the timed
region.  `--impl reference` times the CPU oracle port (the reference itself needs JAX, which this
image does not have) on all host cores, on a bounded sample of the same workload.
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


# --------------------------------------------------------------------------- CPU oracle leg
def _oracle_chunk(args):
    """One worker: advance `C` chains by `T` transitions with the NumPy oracle; returns seconds."""
    cfg, C, T, chain_offset, total_chains = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import prng as P, samplers as S, targets as Tg
    D = cfg["D"]
    if cfg["target"] == "logreg":
        X, y = Tg.make_logreg_data(cfg["N"], D, cfg.get("data_seed", 0))
        tgt = Tg.LogisticRegression(X, y, cfg["prior_precision"])
        tgt.structured_dmetric = cfg["N"] * D ** 3 > 1e9  # same contractions without the (C, D, D, D) tensor
    else:
        tgt = Tg.NealFunnel(D, cfg.get("sigma", 3.0))
        if cfg.get("metric") == "softabs":
            tgt = Tg.softabs_metric(tgt, cfg["softabs_alpha"])
    root = P.key(cfg["root_key"])
    q0 = (np.ones if cfg["init_position"] == "ones" else np.zeros)((C, D), np.float32)
    idx = np.arange(chain_offset, chain_offset + C)
    if cfg["sampler"] == "lmcmonge":
        st = S.lmcmonge_init(q0, tgt)
        step = lambda k, s: S.lmcmonge_step(k, s, tgt, cfg["step_size"], np.ones(D, np.float32),
                                            cfg["num_integration_steps"], alpha2=cfg["alpha2"],
                                            half_step=cfg["half_step"])
    elif cfg["sampler"] == "lmc":
        st = S.lmc_init(q0, tgt)
        step = lambda k, s: S.lmc_step(k, s, tgt, cfg["step_size"], cfg["num_integration_steps"])
    else:
        st = S.rmhmc_init(q0, tgt)
        step = lambda k, s: S.rmhmc_step(k, s, tgt, cfg["step_size"], cfg["num_integration_steps"])
    t0 = time.perf_counter()
    for t in range(T):
        keys = S.chain_keys(root, 1 << 20, t, total_chains, idx)
        st, _ = step(keys, st)
    return time.perf_counter() - t0


def oracle_throughput(cfg, chains_per_worker, transitions, workers):
    """chain-leapfrog-steps/s of the oracle port on `workers` processes (1 = in-process)."""
    L = cfg["num_integration_steps"]
    total = chains_per_worker * workers
    jobs = [(cfg, chains_per_worker, transitions, w * chains_per_worker, total) for w in range(workers)]
    t0 = time.perf_counter()
    if workers == 1:
        _oracle_chunk(jobs[0])
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            pool.map(_oracle_chunk, jobs)
    wall = time.perf_counter() - t0
    return total * L * transitions / wall, wall


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cpw = 2048 if cfg["sampler"] == "lmcmonge" else (8 if cfg["target"] == "logreg" else 64)
    # calibrate the per-step sample so that warmup + steps finish within a few minutes
    v, wall = oracle_throughput(cfg, cpw, 1, cores)
    per_transition = wall
    tps = max(1, int(4.0 / max(per_transition, 1e-3)))  # ~4 s of wall per step: K = 20 steps stay within ~2 minutes
    tps = min(tps, 64)
    for _ in range(args.warmup):
        oracle_throughput(cfg, cpw, 1, cores)
    t0 = time.perf_counter()
    vals = []
    for _ in range(args.steps):
        v, _ = oracle_throughput(cfg, cpw, tps, cores)
        vals.append(v)
    wall = time.perf_counter() - t0
    value = float(np.mean(vals))
    sample = f"{cpw * cores} chains ({cpw}/process x {cores} processes) x {tps} transitions x L={cfg['num_integration_steps']} per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "chain-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["name"], "sampler": cfg["sampler"], "D": cfg["D"],
                   "num_integration_steps": cfg["num_integration_steps"], "step_size": cfg["step_size"]},
        "cpu_baseline": {"value": value, "unit": "chain-steps/s", "cores": cores, "kind": "port",
                         "sample": sample,
                         "note": "NumPy oracle port of the reference semantics (not JAX: jax/jaxlib are not installable in this image)"},
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


# --------------------------------------------------------------------------- our arm
def run_ours(args, cfg):
    import torch
    import geomjax_b200 as g
    from geomjax_b200 import _native as N

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    D, L = cfg["D"], cfg["num_integration_steps"]
    C = args.chains or cfg["chains_per_gpu"]
    TPS = args.transitions_per_step
    K, W = args.steps, args.warmup
    total_chains = C * world
    total_transitions = 1 << 20  # width of the outer split(root, T); fixed so that keys do not depend on K
    if cfg["target"] == "logreg":
        from oracle.targets import make_logreg_data  # frozen synthetic design (input data only)
        Xh, yh = make_logreg_data(cfg["N"], D, cfg.get("data_seed", 0))
        target = g.logistic_regression(torch.from_numpy(Xh).to(dev), torch.from_numpy(yh).to(dev), cfg["prior_precision"])
    else:
        target = g.neal_funnel(D, sigma=cfg.get("sigma", 3.0))
    init_fill = torch.ones if cfg["init_position"] == "ones" else torch.zeros
    root = g.random.PRNGKey(cfg["root_key"])
    if cfg["sampler"] == "lmcmonge":
        integ = {"omega": g.integrators.half_step_omega, "omega_fixed": g.integrators.half_step_omega_fixed,
                 "omegatilde": g.integrators.half_step_omegatilde}[args.half_step or cfg["half_step"]]
        eps = args.step_size or (cfg["step_size_omega_fixed"] if (args.half_step == "omega_fixed") else cfg["step_size"])
        alg = g.lmcmonge(target, eps, torch.ones(D, device=dev), L, alpha2=cfg["alpha2"], integrator=integ,
                         lanes_per_chain=args.lanes_per_chain)
        sampler_id = N.LMCMONGE
    elif cfg["sampler"] == "lmc":
        eps = args.step_size or cfg["step_size"]
        alg = g.lmc(target, eps, target, L, lanes_per_chain=args.lanes_per_chain)
        sampler_id = N.LMC
    else:
        eps = args.step_size or cfg["step_size"]
        metric = g.softabs(target, cfg["softabs_alpha"]) if cfg.get("metric") == "softabs" else target
        alg = g.rmhmc(target, eps, metric, L, lanes_per_chain=args.lanes_per_chain)
        sampler_id = N.RMHMC

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- FP32 roofline denominator: measured FFMA peak on this GPU, same process, same clocks
    out = torch.zeros(1, device=dev)
    iters = 1 << 16
    grid, block = 148 * 8, 256
    for _ in range(2):
        N.check(N.lib().gb200_fp32_peak_kernel(N.ptr(out), grid, block, iters, N.stream_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    N.check(N.lib().gb200_fp32_peak_kernel(N.ptr(out), grid, block, iters, N.stream_ptr()))
    e1.record()
    torch.cuda.synchronize()
    fp32_peak_tflops = 2.0 * grid * block * iters / (e0.elapsed_time(e1) * 1e-3) / 1e12

    # ---- device-resident arm
    state = alg.init(init_fill((C, D), device=dev))
    out_state = [torch.empty_like(t) for t in state]

    def fused(st, first):
        s, _, _ = g.run_fused(alg.step, root, st, TPS, first=first, total=total_transitions,
                              chain_offset=rank * C, total_chains=total_chains, inplace=True)
        return s

    t_idx = 0
    for _ in range(W):
        state = fused(state, t_idx)
        t_idx += TPS
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches0 = N.lib().gb200_kernel_launches()
    t0 = time.perf_counter()
    for k in range(K):
        evs[k][0].record()
        state = fused(state, t_idx)
        evs[k][1].record()
        t_idx += TPS
    launches = N.lib().gb200_kernel_launches() - launches0
    barrier()
    wall = time.perf_counter() - t0
    clocks.stop_flag = True
    kernel_ms = [a.elapsed_time(b) for a, b in evs]
    dev_ms = evs[0][0].elapsed_time(evs[-1][1])
    t_dev = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    dev_ms = float(t_dev.item())
    value = total_chains * L * TPS * K / (dev_ms * 1e-3)
    accept_now = None

    # ---- e2e arm: host buffers, H2D + init + fused transitions + D2H inside the timed region.
    # Steps are double-buffered over two CUDA streams so that step k's copies overlap step k+1's
    # kernel; every step still uploads its inputs from pinned memory and reads its results on the host.
    host_q = init_fill((C, D)).pin_memory()
    streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
    host_out = [torch.empty((C, D)).pin_memory() for _ in streams]
    host_acc = [torch.empty((TPS, C)).pin_memory() for _ in streams]
    done = [torch.cuda.Event() for _ in streams]

    def e2e_enqueue(i, first):
        with torch.cuda.stream(streams[i]):
            q = host_q.to(dev, non_blocking=True)
            st = alg.init(q)
            s, _, acc = g.run_fused(alg.step, root, st, TPS, first=first, total=total_transitions,
                                    chain_offset=rank * C, total_chains=total_chains, return_accept=True)
            host_out[i].copy_(s.position, non_blocking=True)
            host_acc[i].copy_(acc, non_blocking=True)
            done[i].record(streams[i])

    def e2e_collect(i):
        done[i].synchronize()
        return float(host_acc[i].mean())

    def e2e_run(n):
        acc = None
        for k in range(n):
            e2e_enqueue(k % 2, k * TPS)
            if k > 0:
                acc = e2e_collect((k - 1) % 2)
        return e2e_collect((n - 1) % 2)

    e2e_run(max(W, 2))
    barrier()
    t0 = time.perf_counter()
    cur = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for st_ in streams:
        st_.wait_stream(cur)
    accept_now = e2e_run(K)
    for st_ in streams:
        cur.wait_stream(st_)
    e1.record()
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = total_chains * L * TPS * K / (float(e2e_ms.item()) * 1e-3)

    # ---- min-ESS/s: a sampling run that keeps every sample on the device, then sharded R-hat / ESS
    ess_info = {}
    if args.ess_samples > 0:
        Tn = args.ess_samples
        st = alg.init(init_fill((C, D), device=dev))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st, samples, acc = g.run_fused(alg.step, root, st, Tn, first=0, total=Tn, chain_offset=rank * C,
                                       total_chains=total_chains, return_samples=True, return_accept=True)
        e1.record()
        torch.cuda.synchronize()
        samp_ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(samp_ms, op=dist.ReduceOp.MAX)
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        rhat = g.rhat(samples, chain_axis=1, sample_axis=0)
        ess = g.ess(samples, chain_axis=1, sample_axis=0)
        d1.record()
        torch.cuda.synchronize()
        ess_info = {"min_ess_per_s": float(ess.min()) / (float(samp_ms.item()) * 1e-3), "min_ess": float(ess.min()),
                    "max_rhat": float(rhat.max()), "samples_per_chain": Tn,
                    "sampling_ms": float(samp_ms.item()), "diagnostics_ms": d0.elapsed_time(d1),
                    "mean_acceptance": float(acc.mean())}
        del samples

    if rank == 0:
        flops_step = N.lib().gb200_flops_per_chain_step(sampler_id, target.c_struct())
        flops_transition = N.lib().gb200_flops_per_transition(sampler_id, target.c_struct())
        flops_unit = flops_step + flops_transition / L   # per-transition work amortised over the L steps
        fp_iters_per_step = None
        if sampler_id == N.RMHMC:
            # implicit midpoint: (2 + iters) evaluations of the fixed-point map per step; measure iters
            ks = N.KeySource()
            kk = g.random.chain_keys(root, 0, total_transitions, C, chain_offset=rank * C, total_chains=total_chains)
            ks.keys, ks.num_transitions = N.ptr(kk), 1
            _, inf = alg.step.engine.launch(state, ks, want_info=True, extra_info=True)
            fp_iters_per_step = float(inf["fp_iters"].float().mean()) / L
            if cfg["target"] == "logreg":
                Nr = cfg["N"]
                feval = 2.0 * Nr * D * D + 10.0 * Nr * D + D ** 3
            else:
                feval = 21.0 * D + 50.0
            flops_unit = (2.0 + fp_iters_per_step) * feval + flops_transition / L
        med_ms = float(np.median(kernel_ms))
        ach = flops_unit * C * L * TPS / (med_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # dram bytes per launch of the dominant kernel from the committed `ncu --set full` capture
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tr.get(args.workload, {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        state_bytes = (2 * (2 * D + 2) * 4) * C * TPS  # read + write of (q, grad, logp, vol) per transition
        line = {
            "metric": METRIC, "value": value, "unit": "chain-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["name"], "sampler": cfg["sampler"], "D": D, "chains_per_gpu": C,
                       "total_chains": total_chains, "num_integration_steps": L, "transitions_per_step": TPS,
                       "step_size": eps, "half_step": args.half_step or cfg.get("half_step"),
                       "lanes_per_chain": args.lanes_per_chain,
                       "l2_policy": "no flush needed: per-launch HBM traffic is the chain state only "
                                    "(compute-bound kernel; state re-read per transition is L1/L2 resident by design)",
                       "parallelism": f"chains sharded over {world} GPU(s), no data-path collective"},
            "roofline": {"bound": "fp32", "achieved": ach, "peak": fp32_peak_tflops, "unit": "TFLOP/s",
                         "frac": ach / fp32_peak_tflops,
                         "traffic": traffic,
                         "flops_per_chain_step": flops_unit, "flops_per_integrator_step": flops_step,
                         "flops_per_transition_outside_steps": flops_transition,
                         "fp_iters_per_step": fp_iters_per_step, "kernel_ms_median": med_ms,
                         "peak_source": "measured in this process: FFMA microbenchmark kernel (148x8 CTAs x 256 thr, 8 independent FMA chains)",
                         "hbm": {"achieved_gbs": state_bytes / (med_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                                 "note": "state traffic only; HBM is not the bound"}},
            "e2e": {"value": e2e_value, "unit": "chain-steps/s", "h2d_bytes_per_step": C * D * 4,
                    "d2h_bytes_per_step": C * D * 4 + TPS * C * 4, "mean_acceptance": accept_now,
                    "pipelining": "2 CUDA streams, double-buffered pinned host buffers"},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
            "wall_s": wall,
        }
        line.update(ess_info)
        if not args.no_cpu_baseline and world == 1:
            cpw = 2048 if cfg["sampler"] == "lmcmonge" else (8 if cfg["target"] == "logreg" else 64)
            v1, w1 = oracle_throughput(cfg, cpw, 1, 1)
            tr = max(1, min(64, int(12.0 / max(w1, 1e-3))))
            v, w = oracle_throughput(cfg, cpw, tr, 1)
            line["cpu_baseline"] = {"value": v, "unit": "chain-steps/s", "cores": 1, "kind": "port",
                                    "sample": f"{cpw} chains x {tr} transitions x L={L} (NumPy oracle, 1 process, {w:.1f} s)",
                                    "host_cores_available": os.cpu_count()}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(k for k in CONFIGS if not k.startswith("_")))
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default: the workload's)")
    ap.add_argument("--transitions-per-step", type=int, default=16)
    ap.add_argument("--lanes-per-chain", type=int, default=0)
    ap.add_argument("--half-step", default=None)
    ap.add_argument("--step-size", type=float, default=0.0)
    ap.add_argument("--ess-samples", type=int, default=1000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    cfg = CONFIGS[args.workload]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
