#!/usr/bin/env python
"""Benchmark of the AGDIFF sampling hot path (BASELINE.json metric: conformers/sec for the full
5000-step sampler on GEOM-Drugs-shaped synthetic molecules).

    python bench.py --gpus N --steps K --warmup W                 # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference algorithm on host cores

A bench "step" is ONE full sampling call (``langevin_dynamics_sample_diffusion`` with
``n_steps=5000`` and the reference scripts' arguments, scripts/test.py:147-164) over this rank's
shard of the batch: ``--mols`` Drugs-shaped molecules x 2 samples per GPU (weak scaling).
``value`` times it with the inputs already in HBM; ``e2e`` times the same call from pinned HOST
tensors including the host->device copies and the device->host read of the final positions.

Weights: random init (seed 2021).  With ``--regime compact`` (default) the last layer of the local
score MLP is replaced by a constant attraction so that the geometry stays compact like under a
trained model (dense radius graph, ~34 edges per atom); pure random-init dynamics fly apart after a
few steps, which empties the radius graph and would understate the per-step work ~3x
(``--regime random_init`` measures that case).  Both arms use identical weights and arguments.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(type="diffusion", network="dualenc", hidden_dim=128, num_convs=6, num_convs_local=4, cutoff=10.0,
           mlp_act="relu", beta_schedule="sigmoid", beta_start=1.e-7, beta_end=2.e-3, num_diffusion_timesteps=5000,
           edge_order=3, edge_encoder="mlp", smooth_conv=True)                    # configs/drugs_default.yml
SAMPLER = dict(extend_order=False, step_lr=1e-6, w_global=1.0, global_start_sigma=0.5, clip=1000.0, clip_local=20.0)
# FLOPs per edge actually executed by the kernels (2 per MAC, SURVEY.md 8a after the host-side merges)
FLOP_FILTER128 = 2 * (128 * 128 + 128 * 128)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mols", type=int, default=416, help="molecules per GPU (x2 samples each); 416 ~ 37k atoms / 0.94 M edges per evaluation. Throughput keeps rising with the batch (208: 44, 416: ~50, 832: ~54 conformers/s): fixed per-launch costs amortise")
    ap.add_argument("--sampler-steps", type=int, default=5000)
    ap.add_argument("--regime", default="compact", choices=["compact", "random_init"])
    ap.add_argument("--workload", default="drugs", choices=["drugs", "qm9"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=200, help="oracle steps timed per CPU sample (~10 s of CPU work)")
    return ap.parse_args()


def set_regime(model, regime):
    if regime == "compact":
        with torch.no_grad():
            model.grad_local_dist_mlp.layers[2].weight.zero_()
            model.grad_local_dist_mlp.layers[2].bias.fill_(-1.0)


def build_workload(kind, n_mols, seed=2021):
    from agdiff_b200 import graph, synth
    mols = synth.drugs_like(n_mols, seed=seed, force_max=(n_mols >= 16)) if kind == "drugs" else synth.qm9_like(n_mols, seed=seed)
    return [graph.extend_bond_order_host(m) for m in mols]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# --------------------------------------------------------------------------------------------
# CPU baseline: the reference algorithm (oracle port) on the host cores, bounded sample
# --------------------------------------------------------------------------------------------
def cpu_reference_rate(args, n_timed_steps, mols_in_sample=1):
    """conformers/s of the reference's CPU path extrapolated to the full sampler.  The reference
    evaluates both encoders on every step (dualenc.py:486-504) and with a compact geometry the edge
    count is stationary, so the per-step time is constant and 5000 steps = 5000 x (mean step time)."""
    from agdiff_b200 import graph
    from oracle import agdiff_oracle as O
    import agdiff_b200
    torch.manual_seed(2021)
    model = agdiff_b200.get_model(SimpleNamespace(**CFG)).eval()
    set_regime(model, args.regime)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    mols = build_workload(args.workload, 16)
    sizes = [m.num_nodes for m in mols]
    order = np.argsort(np.abs(np.asarray(sizes) - (44 if args.workload == "drugs" else 18)))
    pick = [mols[int(i)] for i in order[:mols_in_sample]]
    z, bi, bt, b, G = graph.collate(pick, 2)
    gen = torch.Generator().manual_seed(0)
    scale = 1.5 if args.regime == "compact" else 12.0
    pos = O.center_pos(torch.randn(z.numel(), 3, generator=gen) * scale, b)
    kw = dict(SAMPLER)
    kw.pop("extend_order")
    t_start = 2012                          # global branch active; the reference computes it on every step anyway
    with torch.no_grad():
        O.sample(sd, CFG, z, pos, bi, bt, b, G, False, n_steps=2, t_start=t_start, scale_init=False, keep_traj=False, **kw)
        t0 = time.perf_counter()
        O.sample(sd, CFG, z, pos, bi, bt, b, G, False, n_steps=n_timed_steps, t_start=t_start, scale_init=False,
                 keep_traj=False, **kw)
        dt = time.perf_counter() - t0
    per_step = dt / n_timed_steps
    rate = G / (per_step * args.sampler_steps)
    desc = ("%d molecule(s) (%s atoms) x 2 samples, %d oracle steps at i=%d after 2 warm-up, %.1f ms/step, "
            "extrapolated x%d" % (mols_in_sample, "+".join(str(m.num_nodes) for m in pick), n_timed_steps, t_start - 1,
                                  per_step * 1e3, args.sampler_steps))
    return rate, desc, per_step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = torch.get_num_threads()
    rates, ms = [], []
    desc = ""
    # the CPU's best batching measured so far (8 molecules x 2 samples per call beats the scripts' one-molecule batches)
    for _ in range(args.warmup):
        cpu_reference_rate(args, 2, 8)
    for _ in range(args.steps):
        r, desc, per_step = cpu_reference_rate(args, args.cpu_steps, 8)
        rates.append(r)
        ms.append(per_step * 1e3 * args.sampler_steps)
    value = float(np.mean(rates))
    line = {"impl": "reference", "metric": "conformers/sec (full 5000-step sampling, Drugs shape)", "value": value,
            "unit": "conformers/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(np.mean(ms)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "gpu_launches": 0,
            "config": workload_config(args, n_conf=16),
            "cpu_baseline": {"value": value, "unit": "conformers/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": "conformers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, n_conf):
    return {"workload": "GEOM-%s-shape synthetic molecules x 2 samples, full %d-step Langevin sampling, "
                        "scripts/test.py arguments (w_global=1, global_start_sigma=0.5, clip=1000, clip_local=20)"
                        % ("Drugs" if args.workload == "drugs" else "QM9", args.sampler_steps),
            "conformers_per_step": n_conf, "molecules_per_gpu": args.mols, "regime": args.regime,
            "weights": "random init seed 2021" + ("; final local-score layer = constant attraction (compact, trained-like "
                                                  "geometry)" if args.regime == "compact" else ""),
            "model_config": "configs/drugs_default.yml", "l2": "working set per step >> 126 MB L2 (no flush needed)"}


# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import agdiff_b200
    from agdiff_b200 import _lib, graph
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: agdiff_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(2021)
    model = agdiff_b200.get_model(SimpleNamespace(**CFG)).eval()
    set_regime(model, args.regime)
    model = model.to(dev)

    mols = build_workload(args.workload, args.mols * world)
    parts = graph.shard_molecules([m.num_nodes for m in mols], world)
    mine = [mols[i] for i in parts[rank]]
    z, bi, bt, b, G = graph.collate(mine, 2)
    gid = torch.tensor([2 * i + s for i in parts[rank] for s in range(2)], dtype=torch.long)
    gen = torch.Generator().manual_seed(1000 + rank)
    pos_init = torch.randn(z.numel(), 3, generator=gen)
    host = [t.pin_memory() for t in (z, pos_init, bi, bt, b)]
    devt = [t.to(dev) for t in host]
    n_conf_total = 2 * len(mols)
    kw = dict(SAMPLER, n_steps=args.sampler_steps, return_traj=False, mol_gid=gid.to(dev))
    max_atoms = max(len(p) and sum(mols[i].num_nodes for i in p) for p in parts) * 2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def gather_final(pos):
        """the only collective: final positions to every rank, padded to the largest shard"""
        if world == 1:
            return pos
        pad = torch.zeros(max_atoms, 3, device=dev)
        pad[: pos.size(0)] = pos
        out = torch.empty(world * max_atoms, 3, device=dev)
        dist.all_gather_into_tensor(out, pad)
        return out

    def step_resident(seed):
        pos, _ = model.langevin_dynamics_sample_diffusion(devt[0], devt[1], devt[2], devt[3], devt[4], G, seed=seed, **kw)
        return gather_final(pos)

    def step_e2e(seed):
        a = [t.to(dev, non_blocking=True) for t in host]
        pos, _ = model.langevin_dynamics_sample_diffusion(a[0], a[1], a[2], a[3], a[4], G, seed=seed, **kw)
        return gather_final(pos).cpu()

    def timed(fn, k):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = model.launch_count()
        ev0.record()
        for s in range(k):
            fn(100 + s)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), model.launch_count() - l0

    for s in range(args.warmup):
        step_resident(s)
    clocks = ClockSampler(local_rank) if rank == 0 else None
    ms_res, launches = timed(step_resident, args.steps)
    clk = clocks.stop() if clocks else None
    ms_e2e, _ = timed(step_e2e, args.steps)
    value = n_conf_total * args.steps / (ms_res / 1e3)
    e2e = n_conf_total * args.steps / (ms_e2e / 1e3)
    h2d = int(sum(t.numel() * t.element_size() for t in host))
    d2h = int((world * max_atoms if world > 1 else z.numel()) * 3 * 4)

    # ---- live per-kernel timing (CUDA events behind every launch, on the launching stream)
    roof = None
    extra = {}
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        pos_fin, _ = model.langevin_dynamics_sample_diffusion(devt[0], devt[1], devt[2], devt[3], devt[4], G, seed=7,
                                                             **dict(kw, n_steps=min(args.sampler_steps, 200), t_start=2012 + 100))
        nb = model._prepare(devt[0], devt[2], devt[3], devt[4], False)
        try:
            lib = _lib.load()
            acc = {}
            E = 0
            reps = 5
            for r in range(reps + 2):
                labels = C.create_string_buffer(1 << 14)
                ms = (C.c_float * 256)()
                n = C.c_int32(0)
                ne = C.c_int32(0)
                _lib.check(lib.agd_profile_forward(model._native_handle(), nb.handle, C.c_void_p(pos_fin.data_ptr()), 1, labels,
                                                   len(labels), ms, 256, C.byref(n), C.byref(ne), model._stream()))
                if r < 2:
                    continue
                E = int(ne.value)
                for lab, t in zip(labels.value.decode().split("\n"), list(ms)[: n.value]):
                    acc.setdefault(lab, []).append(t)
            per_kernel = {k: (float(np.sum(v)) / reps, len(v) // reps) for k, v in acc.items()}   # ms per forward, launches
            total = sum(v[0] for v in per_kernel.values())
            top = max(per_kernel, key=lambda k: per_kernel[k][0])
            extra["kernel_ms_per_forward"] = {k: round(v[0], 4) for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1][0])}
            extra["edges_profiled"] = E
            extra["atoms"] = int(z.numel())
            # dominant kernel: the CFConv filter network of conv1 (F = 128), in whichever arithmetic mode is active
            variants = [
                ("schnet.cfconv128_f16", "tc_filter16_kernel<128, fused> (CFConv filter net + aggregation, tcgen05 kind::f16, fp16 hi/lo' split, "
                                         "two edge tiles in flight per SM)", 3, "3 fp16 MMAs per product at the bf16 rate"),
                ("schnet.cfconv128_f16ws", "tc_filter16_ws_kernel<128> (warp-specialised CFConv filter net + aggregation, tcgen05 kind::f16, "
                                           "fp16 hi/lo' split; opt-in AGD_F16_WS=1)", 3, "3 fp16 MMAs per product at the bf16 rate"),
                ("schnet.filter128_f16", "tc_filter16_kernel<128> (CFConv filter net, tcgen05 kind::f16, fp16 hi/lo' split)", 3,
                 "3 fp16 MMAs per product at the bf16 rate"),
                ("schnet.filter128_tc", "tc_filter_kernel<128> (CFConv filter net, tcgen05 kind::tf32, 3xTF32 split)", 6,
                 "3 TF32 MMAs per product and kind::tf32 runs at half the bf16 rate"),
                ("schnet.filter128", "filter_kernel<128> (CFConv filter net, fp32 FFMA)", None, "fp32 FFMA pipe"),
            ]
            for label, desc, mult, why in variants:
                if label not in per_kernel:
                    continue
                f128_ms, f128_n = per_kernel[label]
                # algorithmic FLOPs (one fp32-equivalent product per MAC), not the split-emulation work
                ach = FLOP_FILTER128 * E / (f128_ms / f128_n * 1e-3) / 1e12
                peak = float(peaks.get("bf16_tflops_sustained", 1400.8))
                roof = {"kernel": desc, "bound": "tensor", "achieved": round(ach, 3), "peak": peak, "unit": "TFLOP/s",
                        "frac": round(ach / peak, 5), "traffic": None, "share_of_forward": round(f128_ms / total, 3),
                        "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (dense bf16 cuBLAS); fp32-faithful math costs "
                                       + why + (", so the attainable ceiling for this kernel is peak/%d" % mult if mult else ""),
                        "dominant_by_time": top}
                if mult:
                    roof["achieved_tensor_tflops_issued"] = round(mult / (2 if mult == 6 else 1) * ach, 3)
                    roof["attainable_peak"] = round(peak / mult, 1)
                    roof["frac_of_attainable"] = round(ach / (peak / mult), 4)
                if label in ("schnet.cfconv128_f16", "schnet.cfconv128_f16ws"):
                    # HBM side of the same launch: the pre-split g2 tile stream in (512 B / edge) + agg rows out; x gathers are L2 hits
                    byts = E * 512 + int(z.numel()) * 512
                    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
                    roof["hbm_algorithmic_gbs"] = round(byts / (f128_ms / f128_n * 1e-3) / 1e9, 1)
                    roof["hbm_frac"] = round(roof["hbm_algorithmic_gbs"] / hbm_peak, 4)
                    roof["traffic_note"] = ("ncu --set full (profiles/): dram read+write = 548 B per edge per launch vs 512 B algorithmic "
                                            "(760 803-edge profiling batch)")
                break
            ag_ms, ag_n = per_kernel.get("schnet.aggregate", (0.0, 1))
            if ag_ms > 0:
                byts = E * (4 * 192 + 4 * 192 + 4) + int(z.numel()) * (4 * 192 + 4)
                ach = byts / (ag_ms / ag_n * 1e-3) / 1e9
                peak = float(peaks.get("hbm_gbs", 6650.0))
                extra["roofline_aggregate"] = {"kernel": "cfconv_aggregate_kernel<192>", "bound": "hbm", "achieved": round(ach, 1),
                                               "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": None}
        finally:
            nb.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r1, d1, _ = cpu_reference_rate(args, args.cpu_steps, 1)
        r8, d8, _ = cpu_reference_rate(args, max(2, args.cpu_steps // 2), 8)
        best = max((r1, d1), (r8, d8))
        cpu = {"value": best[0], "unit": "conformers/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": best[1], "natural_batch": {"value": r1, "sample": d1}, "large_batch": {"value": r8, "sample": d8}}

    if rank == 0:
        line = {"metric": "conformers/sec (full 5000-step sampling, Drugs shape)", "value": value, "unit": "conformers/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args, n_conf_total), "gpu_launches": int(launches),
                "e2e": {"value": e2e, "unit": "conformers/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
                "range_fallbacks": int(model.range_fallbacks)}   # sampler calls re-run on the 3xTF32 kernels (0 expected)
        line.update(extra)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
