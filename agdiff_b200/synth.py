"""Synthetic molecular graphs with QM9 / GEOM-Drugs shapes.

No dataset or checkpoint is available offline, so throughput and parity are measured on
seeded synthetic molecules: a random tree over the heavy atoms (max degree 4) plus a few
ring closures, hydrogens attached to free valences, bond types drawn from the RDKit codes
the reference uses (SINGLE=1, DOUBLE=2, TRIPLE=3, AROMATIC=12; reference
``src/agdiff/utils/chem.py:17``).  The bond list is directed, both directions present, and
sorted by ``row * n + col`` exactly like ``rdmol_to_data`` does
(reference ``src/agdiff/utils/datasets.py:358-360``).

``Molecule`` carries the *bond* graph; ``agdiff_b200.graph.extend_bond_order_host`` adds the
2-/3-hop edges (the reference's offline ``AddHigherOrderEdges`` transform).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

ALANINE_Z = [1, 6, 1, 1, 6, 8, 7, 1, 6, 1, 6, 1, 1, 1, 6, 8, 7, 1, 6, 1, 1, 1]
# 0-based bonds of ACE-ALA-NME (reference examples/alanine_dipeptide.pdb:2-23 has no CONECT
# records; the two C=O are typed DOUBLE here -- a fixed choice shared by oracle and product).
ALANINE_BONDS = [(1, 0, 1), (1, 2, 1), (1, 3, 1), (1, 4, 1), (4, 5, 2), (4, 6, 1), (6, 7, 1), (6, 8, 1),
                 (8, 9, 1), (8, 10, 1), (10, 11, 1), (10, 12, 1), (10, 13, 1), (8, 14, 1), (14, 15, 2),
                 (14, 16, 1), (16, 17, 1), (16, 18, 1), (18, 19, 1), (18, 20, 1), (18, 21, 1)]
ALANINE_POS = [
    (2.000, 1.000, -0.000), (2.000, 2.090, 0.000), (1.486, 2.454, 0.890), (1.486, 2.454, -0.890),
    (3.427, 2.641, -0.000), (4.391, 1.877, -0.000), (3.555, 3.970, -0.000), (2.733, 4.556, -0.000),
    (4.853, 4.614, -0.000), (5.408, 4.316, 0.890), (5.661, 4.221, -1.232), (5.123, 4.521, -2.131),
    (6.630, 4.719, -1.206), (5.809, 3.141, -1.241), (4.713, 6.129, 0.000), (3.601, 6.653, 0.000),
    (5.846, 6.835, 0.000), (6.737, 6.359, -0.000), (5.846, 8.284, 0.000), (4.819, 8.648, 0.000),
    (6.360, 8.648, 0.890), (6.360, 8.648, -0.890)]


@dataclass
class Molecule:
    atom_type: np.ndarray   # (n,) int64 atomic numbers
    bond_index: np.ndarray  # (2, E_b) int64, directed, sorted by row*n+col
    bond_type: np.ndarray   # (E_b,) int64

    @property
    def num_nodes(self) -> int:
        return int(self.atom_type.shape[0])


def _finish(z: Sequence[int], und_bonds) -> Molecule:
    n = len(z)
    rows, cols, types = [], [], []
    for a, b, t in und_bonds:
        rows += [a, b]
        cols += [b, a]
        types += [t, t]
    rows, cols, types = np.asarray(rows, np.int64), np.asarray(cols, np.int64), np.asarray(types, np.int64)
    perm = np.argsort(rows * n + cols, kind="stable")
    return Molecule(np.asarray(z, np.int64), np.stack([rows[perm], cols[perm]]), types[perm])


def alanine_dipeptide() -> Molecule:
    return _finish(ALANINE_Z, ALANINE_BONDS)


def _random_molecule(rng: np.random.Generator, n: int, n_heavy: int, heavy_types: Sequence[int],
                     n_rings: int, bond_types=(1, 1, 1, 1, 2, 12, 12, 3)) -> Molecule:
    n_heavy = max(1, min(n_heavy, n))
    z = [int(rng.choice(heavy_types)) for _ in range(n_heavy)]
    deg = [0] * n_heavy
    bonds = []
    have = set()
    for a in range(1, n_heavy):
        cand = [b for b in range(a) if deg[b] < 3]
        b = int(rng.choice(cand)) if cand else int(np.argmin(deg[:a]))
        t = int(rng.choice(bond_types))
        bonds.append((a, b, t))
        have.add((min(a, b), max(a, b)))
        deg[a] += 1
        deg[b] += 1
    for _ in range(n_rings):
        cand = [a for a in range(n_heavy) if deg[a] < 4]
        if len(cand) < 2:
            break
        a, b = (int(v) for v in rng.choice(cand, size=2, replace=False))
        if (min(a, b), max(a, b)) in have:
            continue
        bonds.append((a, b, int(rng.choice((1, 12)))))
        have.add((min(a, b), max(a, b)))
        deg[a] += 1
        deg[b] += 1
    # hydrogens onto free valences (round-robin over the heavy atoms with the lowest degree)
    for h in range(n_heavy, n):
        free = [a for a in range(n_heavy) if deg[a] < 4]
        a = int(rng.choice(free)) if free else int(rng.integers(n_heavy))
        z.append(1)
        bonds.append((h, a, 1))
        deg[a] += 1
    return _finish(z, bonds)


def qm9_like(num_mols: int, seed: int = 2021) -> List[Molecule]:
    """QM9 shape: n = clip(round(N(18,3)), 5, 29), at most 9 heavy atoms from {C,N,O,F}."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(num_mols):
        n = int(np.clip(np.rint(rng.normal(18.0, 3.0)), 5, 29))
        n_heavy = int(np.clip(np.rint(n * 0.5), 1, 9))
        out.append(_random_molecule(rng, n, n_heavy, (6, 6, 6, 6, 7, 8, 8, 9), int(rng.integers(0, 2))))
    return out


def drugs_like(num_mols: int, seed: int = 2021, force_max: bool = True) -> List[Molecule]:
    """GEOM-Drugs shape: n = clip(round(Gamma(9, 4.9)), 8, 181) (mean ~44), heavy:H ~ 55:45,
    ring closures ~ n_heavy / 8; one 181-atom molecule is forced when ``force_max``."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(num_mols):
        n = int(np.clip(np.rint(rng.gamma(9.0, 4.9)), 8, 181))
        if force_max and i == num_mols // 2:
            n = 181
        n_heavy = max(2, int(np.rint(n * 0.55)))
        out.append(_random_molecule(rng, n, n_heavy, (6, 6, 6, 6, 6, 7, 7, 8, 8, 9, 15, 16, 17, 35, 53),
                                    max(0, n_heavy // 8)))
    return out
