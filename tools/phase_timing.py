"""Per-phase cycle breakdown of the fp16 CFConv kernels (clock64 instrumentation, agd_debug_timing).  GPU box only."""
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

n_mols = int(sys.argv[1]) if len(sys.argv) > 1 else 208
torch.manual_seed(2021)
m = agdiff_b200.get_model(SimpleNamespace(**CFG)).eval().to("cuda:0")
mols = [graph.extend_bond_order_host(x) for x in synth.drugs_like(n_mols, seed=2021)]
z, bi, bt, b, G = graph.collate(mols, 2)
pos = torch.randn(z.numel(), 3, generator=torch.Generator().manual_seed(0)) * 1.5
dev = "cuda:0"
args = (z.to(dev), pos.to(dev), bi.to(dev), bt.to(dev), b.to(dev), None)
out = m(*args, return_edges=True, extend_order=False)
E = out[2].size(1)
names = ["sync A", "L1 issue+cw+wait", "epilogue 1", "L2 issue+wait", "stage+epi2 ld/st+sync", "aggregate (own)", "sync after agg", "-"]
tiles = (E + 127) // 128
print("edges", E, "tiles", tiles, "(12 launches: 6 x F=128 + 6 x F=64 summed; needs a build with AGD_BUILD_DEFS=-DAGD_F16_TIMING)")
for rep in (0, 1):
    m.set_option("f16_timing", 1)
    out = m(*args, return_edges=True, extend_order=False)
    buf = (C.c_uint64 * 64)()
    _lib.check(_lib.load().agd_debug_timing(m._native_handle(), buf))
    t = np.array(list(buf), dtype=np.float64)[:32].reshape(2, 2, 8)
    print("== run", rep)
    for g in range(2):
        for o, who in enumerate(("warp 0", "warp 7")):
            per = t[g, o] / (12 * tiles / 2)   # each group handles half the tiles of each of the 12 launches
            print("group %d %s: total %.0f cycles/tile | " % (g, who, per.sum()) +
                  ", ".join("%s %.0f" % (n, v) for n, v in zip(names[:7], per[:7])))
m.set_option("f16_timing", 0)
