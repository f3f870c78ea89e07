#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> markdown summary under profiles/.
usage: python dev/launch_list_summary.py gpurun_out/launches_X.csv profiles/r2_launches.md "command line that was profiled"

Per-launch times under ncu are cold-cache and serialised; what must agree with bench.py is the SHARE of the step the fused kernel
takes: computed over the launches from the first to the last fused launch (the warm-up + timed region of the profiled run)."""
import collections
import csv
import sys

src, dst, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hi]
ix = {h: i for i, h in enumerate(H)}
L = []
for r in rows[hi + 1:]:
    if len(r) < len(H):
        continue
    unit = r[ix["Metric Unit"]]
    v = float(r[ix["Metric Value"]].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    L.append((int(r[ix["ID"]]), r[ix["Kernel Name"]], r[ix["Grid Size"]], r[ix["Block Size"]], v))
ours = lambda name: name.startswith(("rbis", "void rbis")) or "rbisk" in name
fused = [k for k, l in enumerate(L) if "rbis_fused_kernel" in l[1] or "rbis_group_kernel" in l[1]]
with open(dst, "w") as f:
    f.write(f"# ncu launch list: `{cmd}`\n\n{len(L)} launches captured (`{src.split('/')[-1]}`); durations are per launch under ncu (serialised, cold caches).\n\n")
    if fused:
        lo, hi_ = fused[0], len(L) - 1  # through the statistics kernels that close the run
        span = L[lo:hi_ + 1]
        tot = sum(l[4] for l in span)
        tf = sum(l[4] for l in span if "rbis_fused_kernel" in l[1] or "rbis_group_kernel" in l[1])
        nf = sum(1 for l in span if "rbis_fused_kernel" in l[1] or "rbis_group_kernel" in l[1])
        f.write(f"From the first fused launch to the end of the run (warm-up + timed steps + the closing statistics reduction): {len(span)} launches, "
                f"{tot / 1e3:.2f} ms in kernels, of which the {nf} fused-kernel launches {tf / 1e3:.2f} ms = **{100 * tf / tot:.2f} %** "
                f"(bench.py: the fused launches are {{kernel_ms x launches}} of the step, the statistics tail ~0.13 ms of 181 ms). "
                f"A fused launch of the whole ensemble is split over the launch groups, hence the 29- and 26-CTA grids.\n\n")
    agg = collections.OrderedDict()
    for l in L:
        key = (l[1][:90], l[2], l[3])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += l[4]
    f.write("| kernel | grid | block | launches | mean us | total ms | ours |\n|---|---|---|---:|---:|---:|---|\n")
    for (name, g, b), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{name}` | {g} | {b} | {n} | {t / n:.1f} | {t / 1e3:.2f} | {'yes' if ours(name) else 'torch (workload setup)'} |\n")
print(open(dst).read()[:1500])
