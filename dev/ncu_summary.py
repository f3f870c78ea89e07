#!/usr/bin/env python
"""Summarise an ncu --set full report: key metrics per launch + instruction mix -> CSV/markdown under profiles/.
usage: python dev/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_xxx  [warp_steps_per_launch]"""
import collections
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
W = float(sys.argv[3]) if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']
with open(out + "_summary.csv", "w") as f:
    f.write("metric,unit," + ",".join(f"launch{i}" for i in range(len(rows) - 2)) + "\n")
    for i, h in enumerate(hdr):
        if h in keep:
            line = f"{h},{units[i]}," + ",".join(r[i].replace(",", "") for r in rows[2:])
            f.write(line + "\n")
            print(line)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
if his:
    hi = his[0]
    end = his[1] - 1 if len(his) > 1 else len(rows)
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    ops, samp = collections.Counter(), collections.Counter()
    tot = 0
    for r in rows[hi + 1:end]:
        if len(r) < len(hdr) or r[0] == 'Address':
            continue
        toks = r[ix['Source']].strip().split()
        op = toks[1] if toks[0].startswith('@') else toks[0]
        if not op.startswith(('LDS', 'STS', 'LDL', 'STL', 'LDG', 'STG', 'LDTM', 'STTM')):
            op = op.split('.')[0]
        n = int(r[ix['Instructions Executed']])
        ops[op] += n
        tot += n
        samp[op] += int(r[ix['# Samples']])
    with open(out + "_instmix.csv", "w") as f:
        f.write("opcode,warp_instructions,per_warp_step,stall_samples\n")
        for op, n in ops.most_common(40):
            f.write(f"{op},{n},{n / W if W else ''},{samp[op]}\n")
    print("total warp instructions", tot, "static", end - hi - 1, "per warp-step", tot / W if W else None)
    for op, n in ops.most_common(16):
        print(f"  {op:12s} {n / W if W else n:10.1f} samples {samp[op]}")
