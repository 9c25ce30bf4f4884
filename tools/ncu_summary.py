"""Summarise .ncu-rep captures (ncu --set full) into a markdown table + the CSV bench.py reads for roofline.traffic.
Usage: python tools/ncu_summary.py <edges> <atoms> <local_edges> out.md out.csv rep1.ncu-rep [rep2 ...]"""
import csv, io, subprocess, sys

LABEL = [("tc_cfconv_kernel", "schnet.cfconv_f16"), ("tc_node16_kernel", "schnet.node_f16"), ("tc_gin_kernel", "gin.layer_tc"),
         ("tc_encoder16_kernel<0", "encoder.global_f16"), ("tc_encoder16_kernel<(int)0", "encoder.global_f16"),
         ("tc_encoder16_kernel<1", "encoder.local_f16"), ("tc_encoder16_kernel<(int)1", "encoder.local_f16"),
         ("tc_pair16_kernel<0", "pair.global_f16"), ("tc_pair16_kernel<(int)0", "pair.global_f16"),
         ("tc_pair16_kernel<1", "pair.local_f16"), ("tc_pair16_kernel<(int)1", "pair.local_f16"),
         ("langevin_step_kernel", "step.langevin"), ("edge_weight_kernel", "schnet.edge_weights")]
METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum"]
UNIT = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}

edges, atoms, local = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
out_md, out_csv, reps = sys.argv[4], sys.argv[5], sys.argv[6:]
best = {}
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        lab = next((l for k, l in LABEL if k in name), None)
        if lab is None:
            continue
        rec = {"kernel": name}
        for m in METRICS:
            v = float(r[ix[m]])
            u = units[ix[m]]
            rec[m] = v * UNIT.get(u, 1.0) if "bytes" in m else v
            rec[m + ".unit"] = "byte" if "bytes" in m else u
        # keep the longest launch of each kernel (the first node launch only embeds, ...)
        if lab not in best or rec["gpu__time_duration.sum"] > best[lab]["gpu__time_duration.sum"]:
            best[lab] = rec
with open(out_csv, "w") as f:
    w = csv.writer(f)
    w.writerow(["label", "kernel", "time_us", "dram_bytes_per_launch", "tensor_pct", "issue_pct", "dram_pct", "edges", "atoms", "local_edges"])
    for lab, rec in best.items():
        w.writerow([lab, rec["kernel"], "%.1f" % rec["gpu__time_duration.sum"], "%.0f" % (rec["dram__bytes_read.sum"] + rec["dram__bytes_write.sum"]),
                    "%.1f" % rec["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"],
                    "%.1f" % rec["smsp__issue_active.avg.pct_of_peak_sustained_active"],
                    "%.1f" % rec["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"], edges, atoms, local])
with open(out_md, "w") as f:
    labs = list(best)
    f.write("| metric | " + " | ".join(labs) + " |\n|---|" + "---|" * len(labs) + "\n")
    f.write("| kernel | " + " | ".join(best[l]["kernel"][:40] for l in labs) + " |\n")
    for m in METRICS:
        f.write("| %s [%s] | " % (m, best[labs[0]][m + ".unit"]) + " | ".join("%.4g" % best[l][m] for l in labs) + " |\n")
print(open(out_md).read())
