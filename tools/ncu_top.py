"""Top-N SASS instructions of one kernel in an .ncu-rep by warp-stall samples, with their dominant stall reasons.
Usage: python tools/ncu_top.py file.ncu-rep [kernel index] [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
s, e = secs[which], secs[which + 1]
hdr = rows[s + 1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[s + 2:e]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
stalls = [h for h in hdr if h.startswith("stall_")]
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:topn]
print("kernel", rows[s][1], "total samples", tot)
for i in sorted(order):
    r = data[i]
    n = int(r[ix["# Samples"]] or 0)
    st = sorted(((int(r[ix[k]] or 0), k[6:]) for k in stalls), reverse=True)[:2]
    print("%5d %5.2f%% %-70s %s" % (i, 100.0 * n / tot, r[ix["Source"]][:70], " ".join("%s=%d" % (k, v) for v, k in st if v)))
