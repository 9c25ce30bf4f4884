"""CPU oracle of the COV/MAT path (test infrastructure: only tests/ may import it).

Restates utils/evaluation/covmat.py of the reference in numpy:
  * get_rmsd_confusion_matrix (covmat.py:16-34): RMSD of every generated conformer onto every reference conformer after
    optimal superposition.  The reference calls RDKit's GetBestRMS on hydrogen-stripped molecules (utils/chem.py get_best_rmsd),
    which also minimises over the molecule's symmetry permutations; RDKit is absent here, so this oracle - like the CUDA
    kernel it checks - keeps the given atom order (SVD Kabsch with the reflection correction).  Parity unpinned against RDKit:
    the value is an upper bound of GetBestRMS and equals it for molecules without non-trivial heavy-atom automorphisms.
  * CovMatEvaluator.__call__ (covmat.py:97-165): COV-R / MAT-R / COV-P / MAT-P from the confusion matrices.
"""
import numpy as np


def kabsch_rmsd(a: np.ndarray, b: np.ndarray) -> float:
    """RMSD of b onto a (n x 3 each) after optimal translation + proper rotation (float64)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    a0 = a - a.mean(0)
    b0 = b - b.mean(0)
    u, s, vt = np.linalg.svd(a0.T @ b0)
    if np.linalg.det(u) * np.linalg.det(vt) < 0:   # reflection: flip the smallest singular direction
        s = s.copy()
        s[-1] = -s[-1]
    r2 = ((a0 ** 2).sum() + (b0 ** 2).sum() - 2.0 * s.sum()) / a.shape[0]
    return float(np.sqrt(max(r2, 0.0)))


def rmsd_confusion_matrix(pos_ref: np.ndarray, pos_gen: np.ndarray, sel=None) -> np.ndarray:
    """[num_ref, num_gen], covmat.py:16-34 (rows = reference conformers)."""
    pos_ref = np.asarray(pos_ref, np.float64)
    pos_gen = np.asarray(pos_gen, np.float64)
    if sel is not None:
        pos_ref, pos_gen = pos_ref[:, sel], pos_gen[:, sel]
    out = np.empty((pos_ref.shape[0], pos_gen.shape[0]))
    for i in range(pos_ref.shape[0]):
        for j in range(pos_gen.shape[0]):
            out[i, j] = kabsch_rmsd(pos_ref[i], pos_gen[j])
    return out


def covmat_scores(confusion: np.ndarray, thresholds: np.ndarray):
    """covmat.py:131-150 for one molecule: (COV-R [T], MAT-R, COV-P [T], MAT-P)."""
    ref_min = confusion.min(-1)
    gen_min = confusion.min(0)
    cov_r = (ref_min.reshape(-1, 1) <= thresholds.reshape(1, -1)).mean(0)
    cov_p = (gen_min.reshape(-1, 1) <= thresholds.reshape(1, -1)).mean(0)
    return cov_r, ref_min.mean(), cov_p, gen_min.mean()
