"""Sampling front-end: the loop of the reference's ``scripts/test.py:130-181`` (one molecule at a time through
``repeat_data`` -> ``langevin_dynamics_sample_diffusion`` -> retry with ``clip_local=20`` on ``FloatingPointError`` ->
``--save_traj``), re-stated for a GPU that wants tens of thousands of atoms per call.

The reference samples ``num_samples`` copies of ONE molecule per call (``repeat_data``, ``utils/misc.py:88-90``).  Molecules
are independent on this path (no edge crosses a molecule, ``center_pos`` is per molecule), the product's reductions walk
sorted segments (bit-reproducible, independent of batch composition) and its noise streams are keyed by the global conformer
id, so MANY molecules' sample chains can share one sampler call without changing any molecule's result.  What is kept from
the reference, per molecule:

* ``extend_order=False`` with the order-3 extension done offline (``AddHigherOrderEdges``, ``scripts/test.py:99,155``);
* ``pos_init ~ N(0, 1)`` of shape (num_samples * n_atoms, 3) (``:147``) - drawn from a generator keyed by the molecule's index,
  so that it does not depend on how molecules were grouped into calls;
* at most one retry with ``clip_local=20`` after a ``FloatingPointError`` (``:144-181``).  A NaN in one molecule of a batched
  call raises for the whole call: the batch is bisected until the offending molecules are alone, and only those are retried
  with the clip - the others keep the unclipped result the reference would have given them;
* ``save_traj``: the stacked ``pos_gen_traj`` instead of the final positions (``:166-169``);
* results in input order, molecules listed in ``done`` skipped (``--resume``, ``:121-134``).

Caveat: the fp16-range fallback of the native library (``AGD_ERR_RANGE`` -> the span is repeated on the 3xTF32 kernels) is per
sampler CALL.  As long as no call falls back - the normal case - a molecule's result is bit-identical however the molecules are
grouped (``test_frontend_batched_equals_one_molecule_per_call``); when one does, the molecules that shared the call are computed
by the other (equally fp32-faithful) arithmetic and differ from their one-per-call result at the 1e-6 level.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Iterable, List, Optional, Sequence, Union

import torch

from .graph import collate, extend_bond_order_host, molecule_cost
from .synth import Molecule


@dataclass
class SampleResult:
    index: int                    # position in the input list
    pos_gen: torch.Tensor         # (num_samples * n_atoms, 3) final positions, or (n_steps, num_samples * n_atoms, 3) with save_traj
    num_samples: int
    clip_local: Optional[float]   # None: sampled without local clipping; 20: the retry path was taken
    failed: bool = False          # NaN even with the clip (the reference gives up on the molecule after the second attempt)


def plan_batches(sizes: Sequence[int], samples: Sequence[int], max_atoms: int) -> List[List[int]]:
    """Greedy, order-preserving grouping of molecules into sampler calls of at most ``max_atoms`` atoms (a molecule larger
    than the budget gets a call of its own)."""
    batches, cur, load = [], [], 0
    for i, (n, s) in enumerate(zip(sizes, samples)):
        need = int(n) * int(s)
        if cur and load + need > max_atoms:
            batches.append(cur)
            cur, load = [], 0
        cur.append(i)
        load += need
    if cur:
        batches.append(cur)
    return batches


def sample_conformers(model, mols: Sequence[Molecule], num_samples: Union[int, Sequence[int], Callable[[int], int]] = 2, *,
                      n_steps: int = 5000, w_global: float = 1.0, global_start_sigma: float = 0.5, clip: float = 1000.0,
                      step_lr: float = 1e-6, sampling_type: str = "ld", eta: float = 1.0, save_traj: bool = False,
                      extend_order_offline: bool = True, edge_order: Optional[int] = None, max_atoms_per_call: int = 40000,
                      seed: int = 2021, device=None, done: Iterable[int] = (), retry_clip_local: float = 20.0,
                      on_result: Optional[Callable[[SampleResult], None]] = None) -> List[Optional[SampleResult]]:
    """Sample ``num_samples`` conformers of every molecule (defaults = ``scripts/test.py:46-60``).  ``mols`` carry the plain
    bond graph when ``extend_order_offline`` (the order-3 extension is applied here, like the dataset transform does), or the
    already extended one.  Returns one ``SampleResult`` per molecule (``None`` for indices listed in ``done``)."""
    device = device if device is not None else next(model.parameters()).device
    order = int(edge_order if edge_order is not None else getattr(model.config, "edge_order", 3))
    ext = [extend_bond_order_host(m, order) if extend_order_offline else m for m in mols]
    if callable(num_samples):
        ns = [int(num_samples(i)) for i in range(len(ext))]
    elif isinstance(num_samples, int):
        ns = [num_samples] * len(ext)
    else:
        ns = [int(v) for v in num_samples]
    done = set(int(i) for i in done)
    todo = [i for i in range(len(ext)) if i not in done and ns[i] > 0]
    gid0 = [0]
    for s in ns:
        gid0.append(gid0[-1] + s)          # global conformer ids: molecule i owns [gid0[i], gid0[i+1])
    results: List[Optional[SampleResult]] = [None] * len(ext)

    def pos_init_of(i: int) -> torch.Tensor:
        g = torch.Generator().manual_seed(int(seed) * 1000003 + i)
        return torch.randn(ext[i].num_nodes * ns[i], 3, generator=g)

    def run(group: List[int], clip_local):
        """one sampler call on the conformers of ``group`` -> list of per-molecule tensors"""
        zs, bis, bts, bs, g, off = [], [], [], [], 0, 0
        for i in group:                                      # repeat_data: consecutive identical graphs
            z, bi, bt, b, G = collate([ext[i]], ns[i])
            zs.append(z); bis.append(bi + off); bts.append(bt); bs.append(b + g)
            off += z.numel()
            g += G
        z, bi, bt, b = torch.cat(zs), torch.cat(bis, 1), torch.cat(bts), torch.cat(bs)
        pos0 = torch.cat([pos_init_of(i) for i in group])
        gid = torch.cat([torch.arange(gid0[i], gid0[i + 1]) for i in group])
        pos, traj = model.langevin_dynamics_sample_diffusion(
            atom_type=z.to(device), pos_init=pos0.to(device), bond_index=bi.to(device), bond_type=bt.to(device),
            batch=b.to(device), num_graphs=g, extend_order=False, n_steps=n_steps, step_lr=step_lr, w_global=w_global,
            global_start_sigma=global_start_sigma, clip=clip, clip_local=clip_local, sampling_type=sampling_type, eta=eta,
            seed=seed, mol_gid=gid.to(device), return_traj=save_traj)
        full = torch.stack(list(traj)) if save_traj else pos.cpu()
        out, a = [], 0
        for i in group:
            n = ext[i].num_nodes * ns[i]
            out.append(full[..., a:a + n, :].clone())
            a += n
        return out

    def finish(i: int, t: torch.Tensor, clip_local, failed=False):
        results[i] = SampleResult(i, t, ns[i], clip_local, failed)
        if on_result is not None:
            on_result(results[i])

    def solve(group: List[int]):
        try:
            for i, t in zip(group, run(group, None)):
                finish(i, t, None)
            return
        except FloatingPointError:
            pass
        if len(group) > 1:                                   # isolate the molecule(s) that produced NaN
            mid = len(group) // 2
            solve(group[:mid])
            solve(group[mid:])
            return
        i = group[0]
        try:                                                 # the reference's single retry, scripts/test.py:144-181
            finish(i, run([i], retry_clip_local)[0], retry_clip_local)
        except FloatingPointError:
            finish(i, torch.empty(0, 3), retry_clip_local, failed=True)

    sizes = [ext[i].num_nodes for i in todo]
    for batch in plan_batches(sizes, [ns[i] for i in todo], max_atoms_per_call):
        solve([todo[k] for k in batch])
    return results


def estimated_cost(mols: Sequence[Molecule], num_samples: int = 2) -> int:
    """per-step work proxy of a job (directed edges after radius extension), for sizing ``max_atoms_per_call`` / sharding"""
    return sum(molecule_cost(m.num_nodes) for m in mols) * num_samples
