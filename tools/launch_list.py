"""Markdown summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel of this library launches, total and mean
time, share.  Usage: python tools/launch_list.py gpurun_out/r02_launches_step.csv > profiles/r02_launches_step.md"""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
rows = []
for r in rd:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    u = r[ix["Metric Unit"]]
    us = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u.startswith("ms") else v)
    rows.append((r[ix["Kernel Name"]], us, r[ix["Grid Size"]], r[ix["Block Size"]]))
torch_rows = [r for r in rows if "native::" in r[0] or "at_cuda_detail" in r[0] or "at::" in r[0] or "cub" in r[0].lower()]
ours = [r for r in rows if r not in torch_rows]
agg = collections.OrderedDict()
for n, us, g, b in ours:
    a = agg.setdefault(re.sub(r"\(.*", "", n), [0, 0.0, g, b])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print("# round 2 - ncu launch list of three sampler steps (`tools/prof_step.py 416`: 416 Drugs-shaped molecules x 2, 37 426 atoms, global steps at i = 1500..1498)\n")
print("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_step.csv python tools/prof_step.py 416`,")
print("summarised by `tools/launch_list.py`. Plain launches, no graph; ncu serialises the launches and every kernel starts with cold caches, so")
print("absolute times are above the CUDA-event times of `bench.py` - the SHARES are what is compared. %d launches of this library's kernels," % len(ours))
print("%.2f ms in total; %d launches / %.2f ms of torch kernels belong to the host-side batch preparation before the first step (index sorts," % (tot / 1e3, len(torch_rows), sum(r[1] for r in torch_rows) / 1e3))
print("scans; not on the step path) and are left out.\n")
print("| kernel | launches | total us | mean us | share | grid x block |\n|---|---|---|---|---|---|")
for k, (c, t, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.1f | %.1f %% | %s x %s |" % (k.replace("void ", "").replace("agd::", ""), c, t, t / c, 100 * t / tot, g, b))
cf = [k for k in agg if "tc_cfconv_kernel" in k]
if cf:
    print("\n`tc_cfconv_kernel`: %.1f %% of the step under ncu; live CUDA-event timing in `bench.py` (`roofline.share_of_forward`, which leaves out" % (100 * agg[cf[0]][1] / tot))
    print("the step kernel and the edge export): see `profiles/r02_bench_default_1gpu.json`.")
