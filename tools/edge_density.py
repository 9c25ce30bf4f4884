"""Edges per atom of the bench workload along the sampling trajectory (compact regime), at the steps the CPU arm of bench.py
times (i = 4999, 3000, 2012, 0).  GPU box only.  Output feeds bench.EDGE_DENSITY (committed under profiles/)."""
import json
import os
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import agdiff_b200
from agdiff_b200 import graph
from bench import SAMPLER, build_workload, cfg_for, set_regime

kind = sys.argv[1] if len(sys.argv) > 1 else "drugs"
n_mols = int(sys.argv[2]) if len(sys.argv) > 2 else 416
dev = "cuda:0"
torch.manual_seed(2021)
m = agdiff_b200.get_model(SimpleNamespace(**cfg_for(kind))).eval()
set_regime(m, "compact")
m = m.to(dev)
mols = build_workload(kind, n_mols)
z, bi, bt, b, G = graph.collate(mols, 2)
pos = torch.randn(z.numel(), 3, generator=torch.Generator().manual_seed(1000))
d = [t.to(dev) for t in (z, pos, bi, bt, b)]
out = {}
kw = dict(SAMPLER, return_traj=False, seed=1)
sig = m.step_schedule(1, 1e-6, 0.5)[0]
p = d[1] * sig[-1].to(dev)           # what the sampler starts from (pos_init * sigma_max)
t = 5000
for stop in (4999, 3000, 2012, 0):
    n = t - 1 - stop if stop != 4999 else 0
    if n > 0:
        p, _ = m.langevin_dynamics_sample_diffusion(d[0], p, d[2], d[3], d[4], G, n_steps=n, t_start=t, scale_init=False, **kw)
        t -= n
    ei, et, _ = m.build_edges(p, d[2], d[3], d[4], extend_order=False)
    out[str(stop)] = {"edges_per_atom": ei.size(1) / z.numel(), "rms_radius": float((p - 0).pow(2).sum(-1).mean().sqrt())}
    print(stop, out[str(stop)])
print(json.dumps({"workload": kind, "molecules": n_mols, "atoms": int(z.numel()), "density": out}))
