"""Per-source-line hot spots of an .ncu-rep (needs -lineinfo and --import-source on):
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, items = "?", None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        isamp, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
        continue
    if hdr is None or not r or not r[0].isdigit():
        continue
    try:
        items.append((cur_file, int(r[0]), int(r[isamp]), int(r[ie]), r[1].strip()[:100]))
    except (ValueError, IndexError):
        pass
ts = max(1, sum(i[2] for i in items))
te = max(1, sum(i[3] for i in items))
print(f"samples {ts}  warp-instructions {te}")
for f, ln, s, e, src in sorted(items, key=lambda x: -x[2])[:top]:
    print(f"{100 * s / ts:5.1f}% smp {100 * e / te:5.1f}% ins  {f}:{ln:<4d} {src}")
