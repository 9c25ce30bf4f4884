"""Per-kernel device times of one forward (CUDA events behind every launch) on a compact Drugs batch."""
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

n_mols = int(sys.argv[1]) if len(sys.argv) > 1 else 256
kind = sys.argv[2] if len(sys.argv) > 2 else "drugs"
torch.manual_seed(2021)
m = agdiff_b200.get_model(SimpleNamespace(**CFG)).eval().to("cuda:0")
mols = [graph.extend_bond_order_host(x) for x in (synth.drugs_like(n_mols, seed=2021) if kind == "drugs" else synth.qm9_like(n_mols, seed=2021))]
z, bi, bt, b, G = graph.collate(mols, 2)
pos = (torch.randn(z.numel(), 3, generator=torch.Generator().manual_seed(0)) * 1.5).to("cuda:0")
m._sync_weights()
nb = m._prepare(z.to("cuda:0"), bi.to("cuda:0"), bt.to("cuda:0"), b.to("cuda:0"), False)
lib = _lib.load()
acc = {}
reps = 5
for r in range(reps + 2):
    labels = C.create_string_buffer(1 << 14)
    ms = (C.c_float * 256)()
    n, ne = C.c_int32(0), C.c_int32(0)
    _lib.check(lib.agd_profile_forward(m._native_handle(), nb.handle, C.c_void_p(pos.data_ptr()), 1, labels, len(labels), ms, 256,
                                       C.byref(n), C.byref(ne), m._stream()))
    if r >= 2:
        for lab, t in zip(labels.value.decode().split("\n"), list(ms)[: n.value]):
            acc.setdefault(lab, []).append(t)
tot = 0
for k, v in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
    print("%-24s %7.3f ms  (%d launches)" % (k, sum(v) / reps, len(v) // reps))
    tot += sum(v) / reps
print("total %.3f ms, atoms %d, edges %d" % (tot, z.numel(), ne.value))
