"""Measured per-step time of the captured sampling graphs (global and local-only steps) on a compact Drugs batch."""
import os
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import agdiff_b200
from agdiff_b200 import graph, synth
from bench import CFG, SAMPLER, set_regime

n_mols = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(2021)
m = agdiff_b200.get_model(SimpleNamespace(**CFG)).eval()
set_regime(m, "compact")
m = m.to("cuda:0")
mols = [graph.extend_bond_order_host(x) for x in synth.drugs_like(n_mols, seed=2021)]
z, bi, bt, b, G = graph.collate(mols, 2)
dev = "cuda:0"
pos = (torch.randn(z.numel(), 3, generator=torch.Generator().manual_seed(0)) * 1.5).to(dev)
args = (z.to(dev), pos, bi.to(dev), bt.to(dev), b.to(dev), G)
for name, t_start, n in (("global", 1500, 480), ("local-only", 5000, 1600)):   # long enough to amortise graph capture + instantiation
    for graph_on in (True, False):
        kw = dict(SAMPLER, n_steps=n, t_start=t_start, scale_init=False, return_traj=False, seed=1, use_cuda_graph=graph_on)
        m.langevin_dynamics_sample_diffusion(*args, **dict(kw, n_steps=10))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m.langevin_dynamics_sample_diffusion(*args, **kw)
        e1.record()
        torch.cuda.synchronize()
        print("%-10s steps, cuda graph %-5s: %.3f ms/step (%d atoms)" % (name, graph_on, e0.elapsed_time(e1) / n, z.numel()))
