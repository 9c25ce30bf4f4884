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


def initial_positions(index: int, n_rows: int, seed: int = 2021) -> torch.Tensor:
    """pos_init ~ N(0, 1) of one molecule's conformers (scripts/test.py:147), from a generator keyed by the molecule's index in
    the job so that it does not depend on how molecules are grouped into sampler calls"""
    g = torch.Generator().manual_seed(int(seed) * 1000003 + int(index))
    return torch.randn(n_rows, 3, generator=g)


def sample_conformers(model, mols: Sequence[Molecule], num_samples: Union[int, Sequence[int], Callable[[int], int]] = 2, *,
                      n_steps: int = 5000, w_global: float = 1.0, global_start_sigma: float = 0.5, clip: float = 1000.0,
                      step_lr: float = 1e-6, sampling_type: str = "ld", eta: float = 1.0, save_traj: bool = False,
                      extend_order_offline: bool = True, edge_order: Optional[int] = None, max_atoms_per_call: int = 40000,
                      seed: int = 2021, device=None, done: Iterable[int] = (), retry_clip_local: float = 20.0,
                      on_result: Optional[Callable[[SampleResult], None]] = None,
                      noise_fn: Optional[Callable[[int, int], torch.Tensor]] = None) -> List[Optional[SampleResult]]:
    """Sample ``num_samples`` conformers of every molecule (defaults = ``scripts/test.py:46-60``).  ``mols`` carry the plain
    bond graph when ``extend_order_offline`` (the order-3 extension is applied here, like the dataset transform does), or the
    already extended one.  Returns one ``SampleResult`` per molecule (``None`` for indices listed in ``done``).
    ``noise_fn(index, n_rows) -> (n_steps, n_rows, 3)`` injects the Langevin noise of one molecule's conformers (parity tests);
    by default it is drawn on the device from Philox streams keyed by the global conformer id."""
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
        return initial_positions(i, ext[i].num_nodes * ns[i], seed)

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
        extra = {}
        if noise_fn is not None:
            extra["noise"] = torch.cat([noise_fn(i, ext[i].num_nodes * ns[i]) for i in group], dim=1)
        pos, traj = model.langevin_dynamics_sample_diffusion(
            atom_type=z.to(device), pos_init=pos0.to(device), bond_index=bi.to(device), bond_type=bt.to(device),
            batch=b.to(device), num_graphs=g, extend_order=False, n_steps=n_steps, step_lr=step_lr, w_global=w_global,
            global_start_sigma=global_start_sigma, clip=clip, clip_local=clip_local, sampling_type=sampling_type, eta=eta,
            seed=seed, mol_gid=gid.to(device), return_traj=save_traj, **extra)
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
        except FloatingPointError as err:
            bad_graphs = getattr(err, "bad_graphs", None)
        if len(group) > 1:                                   # isolate the molecule(s) that produced NaN
            if bad_graphs:
                # the sampler reports which conformers went NaN (agd_nan_steps): the others are repeated together without
                # them - one more call instead of a bisection - and only the offenders take the retry path one by one
                first, bad = 0, set()
                for i in group:
                    if any(first <= g < first + ns[i] for g in bad_graphs):
                        bad.add(i)
                    first += ns[i]
                good = [i for i in group if i not in bad]
                if bad and good:
                    solve(good)
                    for i in group:
                        if i in bad:
                            solve([i])
                    return
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


def repeat_data(data, num_repeat: int):
    """``repeat_data`` of the reference (utils/misc.py:88-90: ``Batch.from_data_list([data.clone()] * n)``) for any object
    that carries ``atom_type (n,)``, ``edge_index (2, e)`` and ``edge_type (e,)`` - a PyG ``Data`` or a plain namespace.
    Returns a namespace with the ``Batch`` fields the sampler call reads (scripts/test.py:147-164): ``atom_type, edge_index,
    edge_type, batch, num_graphs, num_nodes``."""
    from types import SimpleNamespace
    z, ei, et = data.atom_type, data.edge_index, data.edge_type
    n = int(z.numel())
    offs = torch.arange(num_repeat, device=ei.device).repeat_interleave(ei.size(1)) * n
    return SimpleNamespace(atom_type=z.repeat(num_repeat), edge_index=ei.repeat(1, num_repeat) + offs, edge_type=et.repeat(num_repeat),
                           batch=torch.arange(num_repeat, device=z.device).repeat_interleave(n), num_graphs=num_repeat,
                           num_nodes=n * num_repeat)


def molecules_from_batch(batch) -> List[Molecule]:
    """The graphs of a PyG-style ``Batch`` / ``Data`` (``atom_type, edge_index, edge_type[, batch]``; bond graph directed both
    ways and, as the datasets store it, possibly already order-extended) as ``Molecule`` records for ``sample_conformers``."""
    z = batch.atom_type.cpu()
    ei, et = batch.edge_index.cpu(), batch.edge_type.cpu()
    b = getattr(batch, "batch", None)
    b = torch.zeros(z.numel(), dtype=torch.long) if b is None else b.cpu()
    out = []
    counts = torch.bincount(b, minlength=int(b.max()) + 1 if b.numel() else 0)
    starts = torch.cumsum(counts, 0) - counts
    eb = b[ei[0]] if ei.numel() else torch.zeros(0, dtype=torch.long)
    for g in range(counts.numel()):
        a0, n = int(starts[g]), int(counts[g])
        m = eb == g
        e, t = ei[:, m] - a0, et[m]
        order = torch.argsort(e[0] * n + e[1])               # Molecule keeps its bond list sorted by row * n + col
        out.append(Molecule(z[a0:a0 + n].numpy().astype("int64"), e[:, order].numpy().astype("int64"), t[order].numpy().astype("int64")))
    return out


def estimated_cost(mols: Sequence[Molecule], num_samples: int = 2) -> int:
    """per-step work proxy of a job (directed edges after radius extension), for sizing ``max_atoms_per_call`` / sharding"""
    return sum(molecule_cost(m.num_nodes) for m in mols) * num_samples
