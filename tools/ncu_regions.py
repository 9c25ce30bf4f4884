"""Summarise an .ncu-rep: headline metrics per kernel and warp-stall samples per code region (40-instruction blocks,
labelled by the marker instructions they contain).  Usage: python tools/ncu_regions.py file.ncu-rep [kernel index]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for w in want:
    for i, h in enumerate(hdr):
        if h == w:
            print(w, [r[i] for r in rows[2:]])
for i, h in enumerate(hdr):
    if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h:
        vals = [r[i] for r in rows[2:]]
        if any(float(v) > 0.3 for v in vals):
            print(h.replace("smsp__average_warps_issue_stalled_", "stall ").replace("_per_issue_active.ratio", ""), vals)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
s, e = secs[which], secs[which + 1]
print("==", rows[s][1])
hdr = rows[s + 1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[s + 2:e]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
blk = 40
marks_all = ["LDTM", "STTM", "UTCHMMA", "LDG", "STG", "STS", "LDS", "MUFU.EX2", "MUFU.LG2", "SYNCS", "BAR.SYNC", "UTCBAR", "F2FP", "ATOM",
             "EXIT", "UBLKCP", "FFMA"]
for b in range(0, len(data), blk):
    chunk = data[b:b + blk]
    n = sum(int(r[ix["# Samples"]] or 0) for r in chunk)
    ex = sum(int(r[ix["Instructions Executed"]] or 0) for r in chunk)
    marks = [m for m in marks_all if any(m in r[ix["Source"]] for r in chunk)]
    st = {}
    for k in ("stall_long_sb", "stall_mio", "stall_short_sb", "stall_lg", "stall_barrier", "stall_wait", "stall_math", "stall_not_selected"):
        v = sum(int(r[ix[k]] or 0) for r in chunk)
        if v > 0.2 * max(n, 1):
            st[k[6:]] = v
    if n > 0.004 * tot:
        print("%4d %5.1f%% exec %9d  %-60s %s" % (b, 100 * n / tot, ex, ",".join(marks), st))
