"""Host-side graph plumbing for the sampling path (runs once per call, never per step).

* ``extend_bond_order_host`` -- the reference's offline ``AddHigherOrderEdges(order=3)``
  transform (reference ``src/agdiff/utils/transforms.py:44-71``; batch twin
  ``src/agdiff/models/common.py:135-205``) restated as a breadth-first search: a pair at
  shortest directed path length k (2 <= k <= order) gets type ``num_types + k - 1`` (23, 24);
  bonds keep their own type; output sorted by ``row * n + col``.
* ``collate`` -- what ``Batch.from_data_list`` / ``repeat_data`` produce for the fields the
  sampler reads (reference ``src/agdiff/utils/misc.py:88-90``): concatenated atom types,
  bond lists with node offsets, and the sorted ``batch`` vector.
* ``shard_molecules`` -- balanced (LPT) partition of independent molecules over ranks
  (SURVEY.md section 8e); there is no collective on the step path.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from .synth import Molecule

NUM_BOND_TYPES = 22  # len(BOND_TYPES), reference src/agdiff/utils/chem.py:17


def extend_bond_order_host(mol: Molecule, order: int = 3, num_types: int = NUM_BOND_TYPES) -> Molecule:
    n = mol.num_nodes
    row, col, typ = mol.bond_index[0], mol.bond_index[1], mol.bond_type
    nbrs: List[List[int]] = [[] for _ in range(n)]
    tmat = {}
    for a, b, t in zip(row.tolist(), col.tolist(), typ.tolist()):
        nbrs[a].append(b)
        tmat[(a, b)] = tmat.get((a, b), 0) + t       # duplicates are summed (to_dense_adj)
    out_r, out_c, out_t = [], [], []
    for a in range(n):
        dist = {a: 0}
        frontier = [a]
        for k in range(1, order + 1):
            nxt = []
            for u in frontier:
                for v in nbrs[u]:
                    if v not in dist:
                        dist[v] = k
                        nxt.append(v)
            frontier = nxt
        for b in sorted(set(dist) | {c for (r, c) in tmat if r == a}):
            if b == a and (a, a) not in tmat:
                continue
            k = dist.get(b, 0)
            t = tmat.get((a, b), 0) + (num_types + k - 1 if k > 1 else 0)
            if t != 0:
                out_r.append(a)
                out_c.append(b)
                out_t.append(t)
    return Molecule(mol.atom_type.copy(), np.asarray([out_r, out_c], np.int64).reshape(2, -1),
                    np.asarray(out_t, np.int64))


def collate(mols: Sequence[Molecule], repeats: int = 1):
    """-> atom_type (N,), bond_index (2,E), bond_type (E,), batch (N,), num_graphs; each
    molecule repeated ``repeats`` times consecutively (``repeat_data`` semantics)."""
    zs, bis, bts, bs = [], [], [], []
    off = 0
    g = 0
    for m in mols:
        for _ in range(repeats):
            zs.append(m.atom_type)
            bis.append(m.bond_index + off)
            bts.append(m.bond_type)
            bs.append(np.full(m.num_nodes, g, np.int64))
            off += m.num_nodes
            g += 1
    return (torch.from_numpy(np.concatenate(zs)), torch.from_numpy(np.concatenate(bis, axis=1)),
            torch.from_numpy(np.concatenate(bts)), torch.from_numpy(np.concatenate(bs)), g)


def molecule_cost(n_atoms: int) -> int:
    """Per-step work proxy: directed edges after radius extension, n * min(n - 1, 32)."""
    return n_atoms * min(max(n_atoms - 1, 0), 32)


def shard_molecules(sizes: Sequence[int], world_size: int) -> List[List[int]]:
    """Longest-processing-time greedy partition of molecule indices over ``world_size`` ranks.
    Deterministic (ties broken by index); every rank's list is ascending."""
    order = sorted(range(len(sizes)), key=lambda i: (-molecule_cost(int(sizes[i])), i))
    load = [0] * world_size
    parts: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        parts[r].append(i)
        load[r] += molecule_cost(int(sizes[i]))
    return [sorted(p) for p in parts]


def mol_ptr_from_batch(batch: torch.Tensor, num_graphs: int | None = None) -> torch.Tensor:
    """CSR pointer over atoms from the sorted ``batch`` vector (int64, on batch.device)."""
    g = int(batch.max().item()) + 1 if num_graphs is None and batch.numel() else (num_graphs or 0)
    counts = torch.bincount(batch, minlength=g)
    ptr = torch.zeros(g + 1, dtype=torch.long, device=batch.device)
    ptr[1:] = torch.cumsum(counts, 0)
    return ptr
