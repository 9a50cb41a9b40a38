"""Per-source-line summary of an ncu report's source page (needs -lineinfo + --import-source on):
python profiles/ncu_lines.py report.ncu-rep [top]   -> per kernel: lines by stall samples, with warp instructions and active threads."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
kernel = None
rows = {}
hdr = None
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        kernel = r[1][:60]
        rows.setdefault(kernel, [])
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if kernel and hdr and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            rows[kernel].append((int(d["# Samples"]), int(d["Instructions Executed"]), int(d["Thread Instructions Executed"]), fname, int(r[0]), r[1].strip()[:110]))
        except (ValueError, KeyError):
            pass
for k, rs in rows.items():
    ts = sum(x[0] for x in rs) or 1
    ti = sum(x[1] for x in rs) or 1
    print(f"=== {k}: {ts} samples, {ti} warp instructions")
    for s, wi, thr, f, ln, src in sorted(rs, reverse=True)[:top]:
        print(f"  {100 * s / ts:5.1f}% smp {100 * wi / ti:5.1f}% inst  act {thr / max(wi, 1):4.1f}  {f}:{ln}  {src}")
