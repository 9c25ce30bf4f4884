"""Per-phase cycles of the SchNet node kernel (tc_node16.cu: clock64 of CTA 0 / thread 0, read through agd_debug_timing).
GPU box only; needs a build with AGD_BUILD_DEFS=-DAGD_F16_TIMING (python -m agdiff_b200.build --force)."""
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
m(*args, return_edges=True, extend_order=False)
PH = ["prologue", "agg load + split", "layer L2a", "epilogue L2a (ssp)", "layer LINa", "xp + agg2 load", "layer L2b", "epilogue L2b (ssp)",
      "layer LINb", "epilogue LINb", "layer A1", "gate + adaptive scaling", "h store + split", "layer next L1a", "epilogue L1a",
      "layer next L1b", "epilogue L1b + tile sync"]
for rep in (0, 1):
    m.set_option("f16_timing", 1)
    m(*args, return_edges=True, extend_order=False)
    buf = (C.c_uint64 * 64)()
    _lib.check(_lib.load().agd_debug_timing(m._native_handle(), buf))
    t = np.array(list(buf), dtype=np.float64)[40:57]
    print("== run", rep, " atoms", z.numel(), " (cycles of CTA 0, summed over the 7 launches of one forward; total %d)" % t.sum())
    for name, v in zip(PH, t):
        print("  %-28s %8d" % (name, v))
    g = np.array(list(buf), dtype=np.float64)[57:64]
    print("  GIN (4 launches summed, CTA 0; total %d)" % g.sum())
    for name, v in zip(["prologue + epilogue 2", "own gather", "wait for the slowest warp", "lift into TMEM", "layer 1", "epilogue 1", "layer 2"], g):
        print("    %-28s %8d" % (name, v))
