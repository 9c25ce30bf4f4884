"""Opcode evidence per kernel from the built library (cuobjdump -sass / -res-usage): tcgen05 / bulk-copy / setmaxnreg / mbarrier / spill
counts of the default-path kernels as a markdown table.  Usage: python tools/sass_summary.py > profiles/r02_sass_table.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "agdiff_b200", "libagdiff_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
res = subprocess.run("cuobjdump -res-usage %s | c++filt" % LIB, shell=True, capture_output=True, text=True).stdout
pat = re.compile(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
cur, counts = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = pat.match(line)
    if m and cur:
        counts[cur][m.group(1).split(".")[0]] += 1
        counts[cur]["full:" + m.group(1)] += 1


def dem(n):
    return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0].replace("agd::", "").replace("void ", "")


regs, name = {}, None
for line in res.splitlines():
    m = re.search(r"Function (.*?):$", line.strip())
    if m:
        name = m.group(1).split("(")[0].replace("agd::", "").replace("void ", "")
    m = re.search(r"REG:(\d+) STACK:(\d+)", line)
    if m and name:
        regs[name] = (m.group(1), m.group(2))
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UBLKPF", "UTMALDG", "UTCCP", "USETMAXREG", "SYNCS", "ELECT", "MUFU", "LDL", "STL"]
DEFAULT = ["tc_cfconv_kernel", "tc_node16_kernel", "tc_encoder16_kernel<true>", "tc_encoder16_kernel<false>", "tc_pair16_kernel<true>",
           "tc_pair16_kernel<false>", "tc_gin_kernel<true>", "tc_gin_kernel<false>", "edge_weight_kernel", "langevin_step_kernel<4>",
           "langevin_step_kernel<8>", "tc_filter16_kernel<128>", "tc_filter16_kernel<64>", "kabsch_rmsd_kernel"]
print("| kernel | instructions | " + " | ".join(KEYS) + " | REG | STACK B |")
print("|---|---|" + "---|" * len(KEYS) + "---|---|")
for k, c in sorted(counts.items(), key=lambda kv: -sum(v for kk, v in kv[1].items() if not kk.startswith("full:"))):
    d = dem(k)
    if d not in DEFAULT:
        continue
    tot = sum(v for kk, v in c.items() if not kk.startswith("full:"))
    r = regs.get(d, ("?", "?"))
    print("| `%s` | %d | " % (d, tot) + " | ".join(str(c[x]) for x in KEYS) + " | %s | %s |" % r)
cf = [k for k in counts if "tc_cfconv_kernel" in k][0]
mm = sorted((k[5:], v) for k, v in counts[cf].items() if k.startswith("full:") and any(t in k for t in ("UTCHMMA", "LDTM", "STTM", "UBLK", "USETMAXREG", "UTCBAR")))
print("\n`tc_cfconv_kernel` tensor-memory / bulk-copy opcodes in full: " + ", ".join("`%s` x %d" % (a, b) for a, b in mm))
