#!/usr/bin/env python
"""Per-source-line view of an ncu report: instructions executed and stall samples by CUDA source line.
usage: python dev/ncu_lines.py report.ncu-rep [warp_steps] [top]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
W = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
inst, samp, text = collections.Counter(), collections.Counter(), {}
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}; src_i = r.index("Source"); continue
    if r[0] == "Function Name" or hdr is None:
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    key = (cur_file, ln)
    text.setdefault(key, r[src_i].strip()[:110])
    try:
        inst[key] += int(r[hdr["Instructions Executed"]]); samp[key] += int(r[hdr["# Samples"]])
    except (ValueError, IndexError):
        pass
ti, ts = sum(inst.values()), sum(samp.values())
print(f"total inst {ti} ({ti / W:.1f} per warp-step), samples {ts}")
for key, n in sorted(samp.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{key[0]}:{key[1]:5d}  inst/step {inst[key] / W:8.1f}  samples {100.0 * n / ts:5.1f}%  | {text[key]}")
