"""Profiling driver: a few forward passes (edge build + both encoders + pair MLPs) on a compact
Drugs-shaped batch, for `ncu` launch lists and full captures.  Not a benchmark."""
import os
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import agdiff_b200
from agdiff_b200 import graph, synth
from bench import CFG

n_mols = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(2021)
m = agdiff_b200.get_model(SimpleNamespace(**CFG)).eval().to("cuda:0")
mols = [graph.extend_bond_order_host(x) for x in synth.drugs_like(n_mols, seed=2021)]
z, bi, bt, b, G = graph.collate(mols, 2)
pos = torch.randn(z.numel(), 3, generator=torch.Generator().manual_seed(0)) * 1.5
dev = "cuda:0"
args = (z.to(dev), pos.to(dev), bi.to(dev), bt.to(dev), b.to(dev), None)
for _ in range(reps):
    out = m(*args, return_edges=True, extend_order=False)
torch.cuda.synchronize()
print("atoms", z.numel(), "edges", out[2].size(1), "local", out[1].size(0), "launches", m.launch_count())
