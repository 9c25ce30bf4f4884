"""Per-role / per-phase cycle breakdown of the warp-specialised CFConv kernel (tc_cfconv.cu: clock64 instrumentation read through
agd_debug_timing).  GPU box only; needs a build with AGD_BUILD_DEFS=-DAGD_F16_TIMING (python -m agdiff_b200.build --force)."""
import ctypes as C
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import agdiff_b200
from agdiff_b200 import _lib, graph, synth
from bench import CFG

n_mols = int(sys.argv[1]) if len(sys.argv) > 1 else 416
torch.manual_seed(2021)
m = agdiff_b200.get_model(SimpleNamespace(**CFG)).eval().to("cuda:0")
mols = [graph.extend_bond_order_host(x) for x in synth.drugs_like(n_mols, seed=2021)]
z, bi, bt, b, G = graph.collate(mols, 2)
pos = torch.randn(z.numel(), 3, generator=torch.Generator().manual_seed(0)) * 1.5
dev = "cuda:0"
args = (z.to(dev), pos.to(dev), bi.to(dev), bt.to(dev), b.to(dev), None)
out = m(*args, return_edges=True, extend_order=False)
E = out[2].size(1)
tiles = (E + 127) // 128
ROLES = {
    "E  epilogue 1 (warp 0)": ["wait D1x", "epi1 x", "wait D1y", "epi1 y"],
    "D  drain (warp 8)": ["tile top (cw loads)", "wait slab empty", "wait D2", "drain + publish"],
    "R  reduce (team 0, warp 16)": ["tile top", "wait slab full", "consume + finish + release", "next x request"],
    "L  loader (warp 12)": ["first loads in flight", "wait A1 free", "stage + remaining loads", "bookkeeping"],
    "M  MMA issue (warp 26)": ["wait A1 full", "issue L1 x+y", "wait A2x", "wait X2 free", "issue L2 x", "wait A2y", "wait Y2 free", "issue L2 y"],
}
print("edges", E, "tiles", tiles, "(6 launches summed; cycles per tile of one CTA)")
for rep in (0, 1):
    m.set_option("f16_timing", 1)
    out = m(*args, return_edges=True, extend_order=False)
    buf = (C.c_uint64 * 64)()
    _lib.check(_lib.load().agd_debug_timing(m._native_handle(), buf))
    t = np.array(list(buf), dtype=np.float64)[:40].reshape(5, 8) / (6 * tiles)
    print("== run", rep)
    for (role, names), row in zip(ROLES.items(), t):
        print("%-24s total %6.0f | " % (role, row.sum()) + ", ".join("%s %.0f" % (n, v) for n, v in zip(names, row) if n != "-"))
m.set_option("f16_timing", 0)
