#!/usr/bin/env python
"""ncu report -> entry of profiles/kernel_profile.json (what bench.py's roofline object reads).

usage: python dev/make_kernel_profile.py REPORT.ncu-rep VARIANT_KEY FILTERS CHUNK_STEPS [--launch I] [--reset]

VARIANT_KEY is str(rbis_batch_last_kernel_variant) of the captured launch ("2" = decoupled lane-per-filter kernel,
"0" dense, "66:imu_only" = 4-lane warp-group kernel on the IMU-only program of configs[1], ...).  The entry holds the FP64
instruction counts of the launch per 32 filter-steps (= lane instructions per filter-step: DFMA, DMUL, DADD from the
source page of the report) and its DRAM traffic; the file records the sha256 of the kernel sources it was captured
from, and bench.py ignores it when the sources have changed since (kernel_source_sha).
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    a = [x for x in sys.argv[1:] if not x.startswith("--")]
    rep, key, filters, steps = a[0], a[1], int(a[2]), int(a[3])
    li = int(sys.argv[sys.argv.index("--launch") + 1]) if "--launch" in sys.argv else 0
    import bench

    sha = bench.kernel_source_sha()
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    r = rows[2 + li]
    g = lambda name: float(r[hdr.index(name)].replace(",", ""))
    units = rows[1]
    unit = lambda name: units[hdr.index(name)]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    dram = g("dram__bytes_read.sum") * scale[unit("dram__bytes_read.sum")] + g("dram__bytes_write.sum") * scale[unit("dram__bytes_write.sum")]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    his = [i for i, rr in enumerate(rows) if rr and rr[0] == "Address"]
    hi = his[li]
    end = his[li + 1] - 1 if len(his) > li + 1 else len(rows)
    ix = {h: i for i, h in enumerate(rows[hi])}
    ops = collections.Counter()
    for rr in rows[hi + 1:end]:
        if len(rr) < len(ix) or rr[0] == "Address":
            continue
        toks = rr[ix["Source"]].strip().split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        ops[op] += int(rr[ix["Instructions Executed"]])
    per = 32.0 / (filters * steps)  # warp instructions -> per 32 filter-steps
    entry = {"dfma": ops["DFMA"] * per, "dmul": ops["DMUL"] * per, "dadd": ops["DADD"] * per,
             "warp_instructions": sum(ops.values()) * per, "dram_bytes": dram, "filters": filters, "chunk_steps": steps,
             "kernel": r[hdr.index("Kernel Name")], "duration_under_ncu": f'{g("gpu__time_duration.sum")} {unit("gpu__time_duration.sum")}',
             "fp64_pipe_pct_ncu": g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
             "registers": g("launch__registers_per_thread"), "source": os.path.basename(rep)}
    path = os.path.join(ROOT, "profiles", "kernel_profile.json")
    prof = {"source_sha": sha, "variants": {}}
    if os.path.exists(path) and "--reset" not in sys.argv:
        with open(path) as f:
            old = json.load(f)
        if old.get("source_sha") == sha:
            prof = old
    prof["variants"][key] = entry
    with open(path, "w") as f:
        json.dump(prof, f, indent=1)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
