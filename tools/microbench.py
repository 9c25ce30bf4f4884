"""BASELINE.json config 4: gather/scatter microbenchmarks over an edge-count sweep.

Synthetic CSR graphs, in-degree 32, N = E/32, uniform features in [-1, 1], seed 0:
  (i)  CFConv message + aggregate, F in {128, 64, 192}: agg_i = sum_e x[src_e] * W_e        (agd_op_cfconv_aggregate)
  (ii) GIN message + aggregate: out_i = (1+eps) x_i + sum_e relu(x[src_e] + ea_e), 128 features  (agd_op_gin_message)
  (iii) eq_transform: out[row] += dd*s, out[col] -= dd*s, with atomics                       (agd_op_eq_transform)
        and in the atomics-free sorted-segment form the step kernel uses                     (agd_op_eq_transform_segments)
Prints achieved GB/s with the algorithmic bytes of SURVEY 8d (indices as int32) against MEASURED_PEAKS.json.
Each timing: 3 warm-up + 10 timed launches with CUDA events; inputs are larger than L2 from E >= 3e5 (F=128).
"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from agdiff_b200 import _lib

lib = _lib.load()
dev = "cuda:0"
peak = 6650.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
P = lambda t: C.c_void_p(t.data_ptr())


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


rows = []
gen = torch.Generator().manual_seed(0)
for E in (100_000, 300_000, 1_000_000, 3_000_000, 10_000_000):
    N = E // 32
    E = N * 32
    src = torch.randint(0, N, (E,), generator=gen).to(torch.int32).to(dev)
    in_ptr = (torch.arange(N + 1, dtype=torch.int32) * 32).to(dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for F in (128, 64, 192):
        x = (torch.rand(N, F, generator=gen) * 2 - 1).to(dev)
        W = (torch.rand(E, F, generator=gen) * 2 - 1).to(dev)
        out = torch.empty(N, F, device=dev)
        t = timeit(lambda: _lib.check(lib.agd_op_cfconv_aggregate(P(x), P(W), P(src), P(in_ptr), N, F, P(out), st)))
        byts = E * (4 * F + 4 * F + 4) + N * (4 * F + 4)
        rows.append(("cfconv_aggregate F=%d" % F, E, t * 1e6, byts / t / 1e9, (E * (4 * F + 4) + N * (8 * F + 4)) / t / 1e9))
        del x, W, out
    pos = torch.randn(N, 3, generator=gen).to(dev)
    dst = torch.arange(N, dtype=torch.int32).repeat_interleave(32).to(dev)
    ln = (torch.rand(E, generator=gen) + 0.5).to(dev)
    sc = torch.randn(E, generator=gen).to(dev)
    out = torch.empty(N, 3, device=dev)
    t = timeit(lambda: _lib.check(lib.agd_op_eq_transform(P(sc), P(pos), P(src), P(dst), P(ln), E, N, P(out), st)))
    byts = E * (4 + 8 + 4 + 24) + N * 24
    rows.append(("eq_transform (atomics)", E, t * 1e6, byts / t / 1e9, (E * 16 + N * 24) / t / 1e9))
    # (iii b) the same edge list as sorted segments: CSC as built (sorted by col = dst) and re-sorted by row for the out-segments
    order = torch.argsort(src.long(), stable=True)
    col_of_out = dst[order].contiguous()
    out_ptr = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(torch.bincount(src.long(), minlength=N), 0)]).to(torch.int32)
    sc_out = sc[order].contiguous()
    t = timeit(lambda: _lib.check(lib.agd_op_eq_transform_segments(P(pos), P(sc_out), P(col_of_out), P(out_ptr), P(sc), P(src), P(in_ptr), N, P(out), st)))
    byts = 2 * E * (4 + 4 + 12) + N * (12 + 8 + 12)   # every edge is visited from both ends: score + index + the far position
    rows.append(("eq_transform (segments)", E, t * 1e6, byts / t / 1e9, (2 * E * 8 + N * 32) / t / 1e9))
    del order, col_of_out, out_ptr, sc_out
    # (ii) GIN message
    x = (torch.rand(N, 128, generator=gen) * 2 - 1).to(dev)
    ea = (torch.rand(E, 128, generator=gen) * 2 - 1).to(dev)
    o = torch.empty(N, 128, device=dev)
    t = timeit(lambda: _lib.check(lib.agd_op_gin_message(P(x), P(ea), P(src), P(in_ptr), N, C.c_float(0.1), P(o), st)))
    byts = E * (512 + 512 + 4) + N * (512 + 512 + 4)
    rows.append(("gin_message", E, t * 1e6, byts / t / 1e9, (E * 516 + N * 1028) / t / 1e9))
    del x, ea, o

# algorithmic: SURVEY 8d's bytes (the gathered operand counted per edge - with N = E/32 it is served by L2, so this is a mixed
# HBM + L2 rate); compulsory: edge stream + indices + the gathered table once + the output, as a fraction of the measured HBM peak
print("| kernel | edges | us/launch | algorithmic GB/s | compulsory HBM GB/s | frac of measured HBM peak (%.0f GB/s) |" % peak)
print("|---|---|---|---|---|---|")
for name, E, us, gbs, comp in rows:
    print("| %s | %d | %.1f | %.0f | %.0f | %.2f |" % (name, E, us, gbs, comp, comp / peak))
