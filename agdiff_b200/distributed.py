"""Multi-GPU sampling: independent molecules shard across ranks with NO collective on the per-step
path (no edge crosses a molecule: reference common.py:217 passes ``batch`` to radius_graph, and
``center_pos`` is per molecule, dualenc.py:581-583); the only exchange is one final gather of the
positions.  One process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo
in the CPU tests) for the plumbing.

Noise streams are keyed by the *global* conformer id, so the result does not depend on the number
of ranks.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist

from .graph import collate, shard_molecules
from .synth import Molecule


def gather_positions(pos_local: torch.Tensor, my_mols: Sequence[int], sizes: Sequence[int], repeats: int,
                     parts: Sequence[Sequence[int]], group=None) -> torch.Tensor:
    """all-gather the per-rank final positions and put them back into global molecule order
    (molecule i, sample s -> rows of conformer i*repeats+s).  Every rank gets the full tensor."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n_rows = [sum(int(sizes[i]) for i in p) * repeats for p in parts]
    if world == 1:
        chunks = [pos_local]
    else:
        pad_to = max(n_rows)
        pad = pos_local.new_zeros((pad_to, 3))
        pad[: pos_local.size(0)] = pos_local
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
        chunks = [bufs[r][: n_rows[r]] for r in range(world)]
    # scatter the rank-major rows back to global conformer order
    offsets = [0]
    for i in range(len(sizes)):
        offsets.append(offsets[-1] + int(sizes[i]) * repeats)
    out = pos_local.new_empty((offsets[-1], 3))
    for r, p in enumerate(parts):
        o = 0
        for i in p:
            n = int(sizes[i]) * repeats
            out[offsets[i]: offsets[i] + n] = chunks[r][o: o + n]
            o += n
    return out


def sample_sharded(sampler: Callable, mols: List[Molecule], repeats: int, pos_init: torch.Tensor, device,
                   group=None, *, seed: int = None, **sampler_kwargs) -> torch.Tensor:
    """Shard ``mols`` (each sampled ``repeats`` times) over the ranks of ``group``, run ``sampler`` (a
    bound ``langevin_dynamics_sample_diffusion``) on the local shard and gather the final positions.
    ``pos_init`` is the full (sum n_i * repeats, 3) initial noise in global conformer order.

    ``seed`` keys the Langevin noise streams (together with the global conformer ids) and must be the same on every rank for
    the result to be independent of the rank count: when it is not given, rank 0 draws one from its torch generator and
    broadcasts it."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if seed is None and "noise" not in sampler_kwargs:
        t = torch.empty((), dtype=torch.int64).random_()
        if world > 1:
            t = t.to(device) if dist.get_backend(group) == "nccl" else t
            dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        seed = int(t.item()) & 0x7FFFFFFFFFFFFFFF
    if seed is not None:
        sampler_kwargs = dict(sampler_kwargs, seed=int(seed))
    sizes = [m.num_nodes for m in mols]
    parts = shard_molecules(sizes, world)
    mine = parts[rank]
    offsets = [0]
    for n in sizes:
        offsets.append(offsets[-1] + n * repeats)
    z, bi, bt, b, G = collate([mols[i] for i in mine], repeats)
    rows = torch.cat([torch.arange(offsets[i], offsets[i + 1]) for i in mine]) if mine else torch.zeros(0, dtype=torch.long)
    gid = torch.tensor([i * repeats + s for i in mine for s in range(repeats)], dtype=torch.long)
    if len(mine):
        pos, _ = sampler(z.to(device), pos_init[rows].to(device), bi.to(device), bt.to(device), b.to(device), G,
                         mol_gid=gid.to(device), return_traj=False, **sampler_kwargs)
    else:
        pos = torch.zeros((0, 3), device=device)
    return gather_positions(pos, mine, sizes, repeats, parts, group)
