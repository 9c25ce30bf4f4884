"""Profiling driver: a few sampler steps (global + local) on a compact Drugs-shaped batch, for `ncu` captures of the step kernel."""
import os
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import agdiff_b200
from agdiff_b200 import graph, synth
from bench import CFG, SAMPLER, set_regime

n_mols = int(sys.argv[1]) if len(sys.argv) > 1 else 416
torch.manual_seed(2021)
m = agdiff_b200.get_model(SimpleNamespace(**CFG)).eval()
set_regime(m, "compact")
m = m.to("cuda:0")
mols = [graph.extend_bond_order_host(x) for x in synth.drugs_like(n_mols, seed=2021)]
z, bi, bt, b, G = graph.collate(mols, 2)
dev = "cuda:0"
pos = (torch.randn(z.numel(), 3, generator=torch.Generator().manual_seed(0)) * 1.5).to(dev)
kw = dict(SAMPLER, n_steps=3, t_start=1500, scale_init=False, return_traj=False, seed=1, use_cuda_graph=False)
m.langevin_dynamics_sample_diffusion(z.to(dev), pos, bi.to(dev), bt.to(dev), b.to(dev), G, **kw)
torch.cuda.synchronize()
print("atoms", z.numel())
