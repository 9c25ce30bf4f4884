#!/usr/bin/env python
"""Benchmark of the AGDIFF sampling hot path (BASELINE.json metric: conformers/sec for the full
5000-step sampler on GEOM-Drugs-shaped synthetic molecules).

    python bench.py --gpus N --steps K --warmup W                   # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference algorithm on host cores
    python bench.py --workload qm9                                  # BASELINE config 2 (1024 QM9-shaped molecules)
    torchrun ... bench.py --gpus N --total-mols 10000               # BASELINE config 5 (fixed set, strong scaling)

A bench "step" is ONE full sampling call (``langevin_dynamics_sample_diffusion`` with
``n_steps=5000`` and the reference scripts' arguments, scripts/test.py:147-164) over this rank's
shard of the batch: ``--mols`` molecules x 2 samples per GPU (weak scaling), or an LPT shard of
``--total-mols`` molecules (strong scaling).  ``value`` times it with the inputs already in HBM;
``e2e`` times the same call from pinned HOST tensors including the host->device copies and the
device->host read of the final positions.

Weights: random init (seed 2021).  With ``--regime compact`` (default) the last layer of the local
score MLP is replaced by a constant attraction so that the geometry stays compact like under a
trained model (dense radius graph, ~34 edges per atom at the end); pure random-init dynamics fly
apart after a few steps, which empties the radius graph and would understate the per-step work ~3x
(``--regime random_init`` measures that case).  Both arms use identical weights and arguments.

CPU arm (``--impl reference`` and the ``cpu_baseline`` object): the oracle port of the reference
algorithm (kind "port": the reference is Python + wheels that are absent offline), all host
threads, on a bounded sample of THE SAME molecule distribution (seeded random subset of the
workload, tails included), timed in windows at i = 4999, 3000, 2012 and 0 at the edge
density the GPU trajectory has there (EDGE_DENSITY below, measured by tools/edge_density.py) and
integrated piecewise-linearly over the 5000 steps.
"""
from __future__ import annotations

import argparse
import csv
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

BASE_CFG = dict(type="diffusion", network="dualenc", hidden_dim=128, num_convs=6, num_convs_local=4, cutoff=10.0,
                mlp_act="relu", beta_schedule="sigmoid", beta_start=1.e-7, beta_end=2.e-3, num_diffusion_timesteps=5000,
                edge_order=3, edge_encoder="mlp")
CFG = dict(BASE_CFG, smooth_conv=True)                                            # configs/drugs_default.yml
SAMPLER = dict(extend_order=False, step_lr=1e-6, w_global=1.0, global_start_sigma=0.5, clip=1000.0, clip_local=20.0)
WORKLOADS = {
    "drugs": dict(cfg=dict(BASE_CFG, smooth_conv=True), yml="configs/drugs_default.yml", name="Drugs", mean_atoms=44,
                  metric="conformers/sec (full 5000-step sampling, Drugs shape)", default_mols=416),
    "qm9": dict(cfg=dict(BASE_CFG, smooth_conv=False), yml="configs/qm9_default.yml", name="QM9", mean_atoms=18,
                metric="conformers/sec (full 5000-step sampling, QM9 shape)", default_mols=1024),
}
# Edges per atom of the GPU trajectory (compact regime) at the CPU arm's time windows, measured on a B200 with
# tools/edge_density.py (profiles/r02_edge_density.json); the CPU sample's geometry is scaled to these densities.
EDGE_DENSITY = {"drugs": {4999: 12.86, 3000: 33.85, 2012: 33.85, 0: 33.85}, "qm9": {4999: 9.47, 3000: 17.57, 2012: 17.57, 0: 17.57}}
WINDOWS = (4999, 3000, 2012, 0)

# ---- algorithmic work per unit executed by each timed kernel (2 FLOP per MAC of the fp32-equivalent product; bytes = mandatory
# HBM streams), keyed by the launch labels of agd_profile_forward.  unit: E = radius-graph edges, L = local edges, N = atoms.
KERNEL_WORK = {
    "schnet.cfconv_f16": ("E", 2 * (128 * 128 + 128 * 128 + 128 * 64 + 64 * 64) + 2 * 192, 512 + 8 + 12, "tensor",
                          "tc_cfconv_kernel: conv1 + conv2 filter nets + aggregation, one launch per block"),
    "schnet.cfconv128_f16": ("E", 2 * (128 * 128 + 128 * 128) + 2 * 128, 512 + 12, "tensor", "tc_filter16_kernel<128, fused> (round 1)"),
    "schnet.cfconv64_f16": ("E", 2 * (128 * 64 + 64 * 64) + 2 * 64, 512 + 12, "tensor", "tc_filter16_kernel<64, fused> (round 1)"),
    "schnet.filter128_f16": ("E", 2 * (128 * 128 + 128 * 128), 512 + 512, "tensor", "tc_filter16_kernel<128> (unfused)"),
    "schnet.filter64_f16": ("E", 2 * (128 * 64 + 64 * 64), 512 + 256, "tensor", "tc_filter16_kernel<64> (unfused)"),
    "schnet.aggregate": ("E", 2 * 192, 4 * 192 + 4 * 192 + 4, "hbm", "cfconv_aggregate_kernel<192>"),
    "schnet.node_f16": ("N", 2 * (128 * 128 * 4 + 64 * 128 * 3 + 128 * 8 * 2), 4 * (192 + 128 + 128 + 192), "tensor", "tc_node16_kernel"),
    "encoder.global_f16": ("E", 2 * (128 + 128 * 128 + 128 * 128), 16 + 512, "tensor", "tc_encoder16_kernel<global>"),
    "encoder.local_f16": ("L", 2 * (128 + 128 * 128 * 3), 16 + 512, "tensor", "tc_encoder16_kernel<local>"),
    "pair.global_f16": ("E", 2 * (128 * 128 * 2 + 128 * 64 + 64), 512 + 16 + 2 * 512, "tensor", "tc_pair16_kernel<global>"),
    "pair.local_f16": ("L", 2 * (128 * 128 * 2 + 128 * 64 + 64), 512 + 16 + 2 * 512, "tensor", "tc_pair16_kernel<local>"),
    "gin.layer_tc": ("N", 2 * (128 * 128 * 2), 4 * 128 * 3, "hbm", "tc_gin_kernel (+ 1032 B per local edge)"),
    "schnet.edge_weights": ("E", 12 * 200, 4 + 48, "hbm", "edge_weight_kernel"),
}


def cfg_for(kind):
    return WORKLOADS[kind]["cfg"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mols", type=int, default=0, help="molecules per GPU (x2 samples each); default 416 (drugs: ~37k atoms / "
                    "0.94 M edges per evaluation) or 1024 (qm9, BASELINE config 2)")
    ap.add_argument("--total-mols", type=int, default=0, help="fixed total number of molecules, LPT-sharded over the ranks "
                    "(strong scaling, BASELINE config 5: 10000); overrides --mols")
    ap.add_argument("--sampler-steps", type=int, default=5000)
    ap.add_argument("--regime", default="compact", choices=["compact", "random_init"])
    ap.add_argument("--workload", default="drugs", choices=["drugs", "qm9"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=4, help="oracle steps timed per window of the CPU sample")
    ap.add_argument("--cpu-mols", type=int, default=8, help="molecules (x2 samples) in the CPU sample")
    a = ap.parse_args()
    if a.mols <= 0:
        a.mols = WORKLOADS[a.workload]["default_mols"]
    return a


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
# CPU arm: the reference algorithm (oracle port) on the host cores, bounded sample
# --------------------------------------------------------------------------------------------
def host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm always uses every host core."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    torch.set_num_threads(n)
    return n


class CpuSample:
    """A seeded random subset of the workload's molecules and, per time window, a geometry whose radius-graph density equals
    the GPU trajectory's at that step (EDGE_DENSITY).  ``rate()`` = conformers/s of the reference algorithm on it."""

    def __init__(self, args):
        from agdiff_b200 import graph
        from oracle import agdiff_oracle as O
        import agdiff_b200
        self.O, self.args = O, args
        W = WORKLOADS[args.workload]
        self.cfg = W["cfg"]
        torch.manual_seed(2021)
        model = agdiff_b200.get_model(SimpleNamespace(**self.cfg)).eval()
        set_regime(model, args.regime)
        self.sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        pool = build_workload(args.workload, WORKLOADS[args.workload]["default_mols"])
        sizes = np.asarray([m.num_nodes for m in pool])
        self.pool_mean_atoms = float(sizes.mean())
        k = min(args.cpu_mols, len(pool))
        for self.seed in range(7, 7 + 256):     # first seeded draw whose mean size is within 3 % of the workload's (tails included)
            pick = sorted(np.random.default_rng(self.seed).choice(len(pool), size=k, replace=False).tolist())
            if abs(sizes[pick].mean() - self.pool_mean_atoms) <= 0.03 * self.pool_mean_atoms:
                break
        self.mols = [pool[i] for i in pick]
        self.z, self.bi, self.bt, self.b, self.G = graph.collate(self.mols, 2)
        self.kw = {k: v for k, v in SAMPLER.items() if k != "extend_order"}
        gen = torch.Generator().manual_seed(0)
        unit = O.center_pos(torch.randn(self.z.numel(), 3, generator=gen), self.b)
        sig_max = float(((1.0 - self.sd["alphas"]).sqrt() / self.sd["alphas"].sqrt())[-1])
        self.pos, self.density = {}, {}
        for i in WINDOWS:
            target = EDGE_DENSITY[args.workload][i] if args.regime == "compact" else None
            if i == WINDOWS[0] or target is None:
                scale = sig_max                    # the sampler's own start: pos_init * sigma_max
            else:                                  # bisection on the spread until the radius graph has the target density
                lo, hi = 0.3, sig_max
                for _ in range(12):
                    mid = 0.5 * (lo + hi)
                    if self._density(unit * mid) > target:
                        lo = mid
                    else:
                        hi = mid
                scale = 0.5 * (lo + hi)
            self.pos[i] = unit * scale
            self.density[i] = self._density(self.pos[i])

    def _density(self, pos):
        ei, _ = self.O.build_edges(pos, self.bi, self.bt, self.b, self.cfg, extend_order=False)
        return ei.size(1) / self.z.numel()

    def rate(self, n_timed, n_warm=1):
        per_step = {}
        with torch.no_grad():
            for i in WINDOWS:
                a = (self.sd, self.cfg, self.z, self.pos[i], self.bi, self.bt, self.b, self.G, False)
                kw = dict(self.kw, t_start=i + 1, scale_init=False, keep_traj=False)
                if n_warm:
                    self.O.sample(*a, n_steps=n_warm, **kw)
                t0 = time.perf_counter()
                self.O.sample(*a, n_steps=n_timed, **kw)
                per_step[i] = (time.perf_counter() - t0) / n_timed
        total = 0.0                                # piecewise-linear integral of the step time over i = 4999 .. 0
        for hi, lo in zip(WINDOWS[:-1], WINDOWS[1:]):
            total += 0.5 * (per_step[hi] + per_step[lo]) * (hi - lo)
        total += per_step[WINDOWS[-1]]
        total *= self.args.sampler_steps / 5000.0
        rate = self.G / total
        desc = ("%d molecules drawn at random (seed %d) from the workload (%s atoms; sample mean %.1f vs workload mean %.1f) x 2 samples; "
                "%d oracle steps (after %d) in each of the windows i = %s at %s edges/atom, %s ms/step, integrated over %d steps"
                % (len(self.mols), self.seed, "+".join(str(m.num_nodes) for m in self.mols), self.z.numel() / self.G, self.pool_mean_atoms,
                   n_timed, n_warm, "/".join(str(i) for i in WINDOWS), "/".join("%.1f" % self.density[i] for i in WINDOWS),
                   "/".join("%.0f" % (per_step[i] * 1e3) for i in WINDOWS), self.args.sampler_steps))
        return rate, desc, total


def workload_config(args, n_conf, mols_per_gpu, scaling):
    W = WORKLOADS[args.workload]
    return {"workload": "GEOM-%s-shape synthetic molecules x 2 samples, full %d-step Langevin sampling, "
                        "scripts/test.py arguments (w_global=1, global_start_sigma=0.5, clip=1000, clip_local=20)"
                        % (W["name"], args.sampler_steps),
            "conformers_per_step": n_conf, "molecules_per_gpu": mols_per_gpu, "regime": args.regime, "sharding": scaling,
            "weights": "random init seed 2021" + ("; final local-score layer = constant attraction (compact, trained-like "
                                                  "geometry)" if args.regime == "compact" else ""),
            "model_config": W["yml"], "l2": "working set per step >> 126 MB L2 (no flush needed)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    W = WORKLOADS[args.workload]
    sample = CpuSample(args)
    rates, secs, desc = [], [], ""
    for _ in range(args.warmup):
        sample.rate(1, 0)
    for _ in range(args.steps):
        r, desc, total = sample.rate(args.cpu_steps)
        rates.append(r)
        secs.append(total)
    value = float(np.mean(rates))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {"impl": "reference", "metric": W["metric"], "value": value, "unit": "conformers/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True,
            "scaling": "strong" if args.total_mols else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "gpu_launches": 0,
            "config": workload_config(args, sample.G, args.total_mols // max(world, 1) if args.total_mols else args.mols,
                                      "strong (fixed molecule set)" if args.total_mols else "weak (fixed molecules per GPU)"),
            "cpu_baseline": {"value": value, "unit": "conformers/s", "cores": cores, "kind": "port", "sample": desc},
            "same_work": {"edges_per_atom_windows": {str(i): round(sample.density[i], 2) for i in WINDOWS},
                          "atoms_per_conformer": sample.z.numel() / sample.G, "workload_atoms_per_conformer": sample.pool_mean_atoms},
            "e2e": {"value": value, "unit": "conformers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
def ncu_traffic(label):
    """dram read + write bytes per launch and the edge count of the profiling batch, from the committed ncu summary."""
    path = os.path.join(ROOT, "profiles", "r02_ncu_kernels.csv")
    try:
        with open(path) as f:
            for row in csv.DictReader(f):
                if row["label"] == label:
                    return float(row["dram_bytes_per_launch"]), int(row["edges"]), int(row["atoms"])
    except (OSError, KeyError, ValueError):
        pass
    return None


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
    W = WORKLOADS[args.workload]

    torch.manual_seed(2021)
    model = agdiff_b200.get_model(SimpleNamespace(**W["cfg"])).eval()
    set_regime(model, args.regime)
    model = model.to(dev)

    strong = args.total_mols > 0
    mols = build_workload(args.workload, args.total_mols if strong else args.mols * world)
    parts = graph.shard_molecules([m.num_nodes for m in mols], world)
    cost = [sum(mols[i].num_nodes * min(mols[i].num_nodes - 1, 32) for i in p) for p in parts]
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

    # ---- live per-kernel timing (CUDA events behind every launch, on the launching stream) + roofline table
    roof = None
    extra = {}
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tf_peak = float(peaks.get("bf16_tflops_sustained", 1400.8))
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = ("MEASURED_PEAKS.json (bf16_tflops_sustained, hbm_gbs)" if peaks else
                    "B200_PROFILING.md fallback (1400.8 TFLOP/s sustained dense bf16, 6650 GB/s)")
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
            N, L = int(z.numel()), int(nb.n_local)
            units = {"E": E, "L": L, "N": N}
            extra["kernel_ms_per_forward"] = {k: round(v[0], 4) for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1][0])}
            extra["edges_profiled"] = E
            extra["atoms"] = N
            extra["local_edges"] = L
            table = []
            for label, (ms_k, n_k) in sorted(per_kernel.items(), key=lambda kv: -kv[1][0]):
                if label not in KERNEL_WORK:
                    continue
                unit, flop, byts, bound, desc = KERNEL_WORK[label]
                extra_b = 1032 * L if label == "gin.layer_tc" else 0
                sec = ms_k / n_k * 1e-3
                tfl = flop * units[unit] / sec / 1e12
                gbs = (byts * units[unit] + extra_b) / sec / 1e9
                table.append({"kernel": label, "launches": n_k, "ms_per_launch": round(ms_k / n_k, 4), "share_of_forward": round(ms_k / total, 3),
                              "bound": bound, "tflops": round(tfl, 1), "frac_tensor": round(tfl / tf_peak, 4),
                              "gbs": round(gbs, 1), "frac_hbm": round(gbs / hbm_peak, 4)})
            extra["kernel_roofline"] = table
            top = max(per_kernel, key=lambda k: per_kernel[k][0])
            if top in KERNEL_WORK:
                unit, flop, byts, bound, desc = KERNEL_WORK[top]
                ms_k, n_k = per_kernel[top]
                sec = ms_k / n_k * 1e-3
                ach = flop * units[unit] / sec / 1e12
                roof = {"kernel": desc, "bound": bound, "achieved": round(ach, 3), "peak": tf_peak, "unit": "TFLOP/s",
                        "frac": round(ach / tf_peak, 5), "traffic": None, "share_of_forward": round(ms_k / total, 3),
                        "peak_source": peak_src + "; fp32-faithful math costs 3 fp16 MMAs per product at the bf16 rate, so the "
                                       "attainable ceiling of the split-precision kernels is peak/3",
                        "algorithmic_flop_per_unit": flop, "algorithmic_bytes_per_unit": byts, "units_per_launch": units[unit],
                        "frac_of_attainable": round(3 * ach / tf_peak, 4),
                        "hbm_algorithmic_gbs": round(byts * units[unit] / sec / 1e9, 1),
                        "hbm_frac": round(byts * units[unit] / sec / 1e9 / hbm_peak, 4)}
                tr = ncu_traffic(top)
                if tr is not None:    # ncu --set full capture of the same kernel (committed): bytes scale with the edge count
                    roof["traffic"] = round(tr[0] * units[unit] / (tr[1] if unit != "N" else tr[2]))
                    roof["traffic_source"] = ("profiles/r02_ncu_kernels.csv: dram__bytes_read.sum + dram__bytes_write.sum of one launch at "
                                              "%d edges, scaled to this batch's %d" % (tr[1], E))
        finally:
            nb.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_threads()
        sample = CpuSample(args)
        sample.rate(1, 0)
        r, d, _ = sample.rate(args.cpu_steps)
        cpu = {"value": r, "unit": "conformers/s", "cores": cores, "kind": "port", "sample": d}

    if rank == 0:
        line = {"metric": W["metric"], "value": value, "unit": "conformers/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res / args.steps,
                "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": workload_config(args, n_conf_total, len(mine), "strong (fixed molecule set, LPT shards)" if strong else
                                          "weak (fixed molecules per GPU)"),
                "gpu_launches": int(launches),
                "e2e": {"value": e2e, "unit": "conformers/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
                "shard_imbalance": round(max(cost) / (sum(cost) / len(cost)), 4) if cost and sum(cost) else 1.0,
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
