"""Summarise an .ncu-rep (captured with --set full --import-source on) into a small text file
that can be committed under profiles/: headline metrics, executed-instruction mix, stall mix.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/r01_xxx.txt
"""
import collections
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max",
]


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"== kernel: {name}")
        for h, u, v in zip(hdr, units, row):
            if h in WANT:
                print(f"  {h:70s} {v} {u}")
        stalls = [(h, float(v)) for h, v in zip(hdr, row)
                  if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")
                  or h.startswith("smsp__average_warp_latency_issue_stalled") and h.endswith(".ratio")]
        for h, v in sorted(stalls, key=lambda x: -x[1])[:8]:
            print(f"  stall {h:75s} {v:.3f}")
    sass = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    hdr = next((r for r in sass if "Instructions Executed" in r), None)
    if hdr is None:
        return
    ia, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
    byop, sampop, stall_tot = collections.Counter(), collections.Counter(), collections.Counter()
    n = tot = 0
    for r in sass:
        if len(r) <= ia or r is hdr:
            continue
        try:
            e, s = int(r[ia]), int(r[isamp])
        except ValueError:
            continue
        toks = r[isrc].split()
        op = (toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")).split(".")[0]
        byop[op] += e
        sampop[op] += s
        tot += e
        n += 1
        for i in stall_cols:
            try:
                stall_tot[hdr[i]] += int(r[i])
            except (ValueError, IndexError):
                pass
    ssum = max(1, sum(sampop.values()))
    print(f"== SASS: {n} static instructions, {tot} warp-instructions executed, {ssum} samples")
    for op, c in byop.most_common(22):
        print(f"  {op:10s} {100 * c / max(tot, 1):5.1f}% of executed   {100 * sampop[op] / ssum:5.1f}% of samples")
    st = max(1, sum(stall_tot.values()))
    print("== stall reasons (sampled, all warps)")
    for k, v in stall_tot.most_common(10):
        print(f"  {k:28s} {100 * v / st:5.1f}%")


if __name__ == "__main__":
    main()
