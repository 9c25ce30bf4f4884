"""Markdown rows of the per-kernel roofline table from a bench.py JSON line.  Usage: python tools/roofline_table.py profiles/r02_bench_default_1gpu.json"""
import json
import sys

d = json.load(open(sys.argv[1]))
print("| kernel | launches | ms/launch | share of the evaluation | algorithmic TFLOP/s (frac of peak) | algorithmic GB/s (frac of HBM peak) |")
print("|---|---|---|---|---|---|")
for k in d["kernel_roofline"]:
    print("| `%s` | %d | %.3f | %.1f %% | %.1f (%.3f) | %.0f (%.2f) |" % (k["kernel"], k["launches"], k["ms_per_launch"], 100 * k["share_of_forward"],
                                                                   k["tflops"], k["frac_tensor"], k["gbs"], k["frac_hbm"]))
print("\nvalue %.2f %s, e2e %.2f, cpu_baseline %.4f, roofline.frac %.4f (of attainable %.3f), traffic %d B/launch, edges profiled %d" % (
    d["value"], d["unit"], d["e2e"]["value"], d.get("cpu_baseline", {}).get("value", float("nan")), d["roofline"]["frac"],
    d["roofline"].get("frac_of_attainable", float("nan")), d["roofline"].get("traffic") or 0, d.get("edges_profiled", 0)))
