"""COV / MAT evaluation of generated conformers on the GPU - the interface of the reference's
``utils/evaluation/covmat.py`` (``get_rmsd_confusion_matrix``, ``evaluate_conf``, ``CovMatEvaluator``,
``print_covmat_results``) on top of one CUDA kernel, ``agd_op_kabsch_rmsd``.

Difference to the reference, stated once: RDKit's ``GetBestRMS`` (utils/chem.py ``get_best_rmsd``) minimises the aligned
RMSD over the molecule's symmetry permutations as well; this implementation keeps the atom order it is given (RDKit is
not a dependency), so its RMSD is an upper bound of the reference's - equal for molecules without non-trivial
heavy-atom automorphisms - and COV scores are lower bounds, MAT scores upper bounds.  Like the reference it compares
heavy atoms only (``RemoveHs``), which needs the atomic numbers: ``data["atom_type"]`` or ``data["rdmol"]``.  The
force-field option (``useFF``) is RDKit's MMFF and is not available.
"""
import ctypes as C
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib


def _atomic_numbers(data):
    if "atom_type" in data and data["atom_type"] is not None:
        return torch.as_tensor(data["atom_type"]).view(-1).cpu()
    mol = data["rdmol"] if "rdmol" in data else None
    if mol is not None:
        return torch.tensor([a.GetAtomicNum() for a in mol.GetAtoms()])
    raise KeyError("COV/MAT needs the atomic numbers: data['atom_type'] (or an RDKit molecule in data['rdmol'])")


def rmsd_matrix(pos_ref, pos_gen, sel=None, device="cuda:0") -> torch.Tensor:
    """[num_ref, num_gen] aligned RMSD (device tensor); pos_* are [num, n_atoms, 3], sel an index tensor of the atoms compared."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("agdiff_b200 has no CPU path: rmsd_matrix needs a CUDA device")
    ref = torch.as_tensor(pos_ref, dtype=torch.float32).to(dev).contiguous()
    gen = torch.as_tensor(pos_gen, dtype=torch.float32).to(dev).contiguous()
    if ref.dim() != 3 or gen.dim() != 3 or ref.size(1) != gen.size(1) or ref.size(2) != 3 or gen.size(2) != 3:
        raise ValueError("expected [num_ref, n, 3] and [num_gen, n, 3]")
    n = ref.size(1)
    out = torch.empty(ref.size(0), gen.size(0), device=dev)
    sel_t = None if sel is None else torch.as_tensor(sel, dtype=torch.int32).to(dev).contiguous()
    lib = _lib.load()
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.agd_op_kabsch_rmsd(C.c_void_p(ref.data_ptr()), C.c_void_p(gen.data_ptr()),
                                          C.c_void_p(sel_t.data_ptr()) if sel_t is not None else None,
                                          0 if sel_t is None else sel_t.numel(), n, ref.size(0), gen.size(0),
                                          C.c_void_p(out.data_ptr()), st))
    return out


def get_rmsd_confusion_matrix(data, useFF=False, device="cuda:0") -> np.ndarray:
    """covmat.py:16-34: rows = reference conformers, columns = generated conformers, heavy atoms only."""
    if useFF:
        raise NotImplementedError("useFF=True relies on RDKit's MMFF optimiser")
    z = _atomic_numbers(data)
    n = z.numel()
    pos_ref = torch.as_tensor(data["pos_ref"]).reshape(-1, n, 3)
    pos_gen = torch.as_tensor(data["pos_gen"]).reshape(-1, n, 3)
    heavy = torch.nonzero(z > 1).view(-1)
    return rmsd_matrix(pos_ref, pos_gen, heavy if heavy.numel() < n else None, device).double().cpu().numpy()


def evaluate_conf(data, useFF=False, threshold=0.5, device="cuda:0"):
    """covmat.py:37-40: (COV-R at `threshold`, MAT-R) of one molecule."""
    m = get_rmsd_confusion_matrix(data, useFF=useFF, device=device)
    ref_min = m.min(-1)
    return (ref_min <= threshold).mean(), ref_min.mean()


def print_covmat_results(results, print_fn=print):
    """covmat.py:43-73 without the pandas dependency: the same table as text."""
    rows = ["%9s %10s %12s %9s %10s %12s %9s" % ("threshold", "COV-R_mean", "COV-R_median", "COV-R_std", "COV-P_mean", "COV-P_median", "COV-P_std")]
    for k, t in enumerate(results.thresholds):
        r, p = results.CoverageR[:, k], results.CoverageP[:, k]
        rows.append("%9.2f %10.4f %12.4f %9.4f %10.4f %12.4f %9.4f" % (t, r.mean(), np.median(r), r.std(), p.mean(), np.median(p), p.std()))
    print_fn("\n" + "\n".join(rows))
    print_fn("MAT-R_mean: %.4f | MAT-R_median: %.4f | MAT-R_std %.4f" % (np.mean(results.MatchingR), np.median(results.MatchingR), np.std(results.MatchingR)))
    print_fn("MAT-P_mean: %.4f | MAT-P_median: %.4f | MAT-P_std %.4f" % (np.mean(results.MatchingP), np.median(results.MatchingP), np.std(results.MatchingP)))
    return rows


class CovMatEvaluator(object):
    """covmat.py:76-171.  `num_workers` is accepted and ignored (the reference fans RDKit calls out over a process pool; here
    every molecule is one kernel launch)."""

    def __init__(self, num_workers=8, use_force_field=False, thresholds=np.arange(0.05, 3.05, 0.05), ratio=2,
                 filter_disconnected=True, print_fn=print, device="cuda:0"):
        if use_force_field:
            raise NotImplementedError("use_force_field=True relies on RDKit's MMFF optimiser")
        self.num_workers = num_workers
        self.use_force_field = use_force_field
        self.thresholds = np.array(thresholds).flatten()
        self.ratio = ratio
        self.filter_disconnected = filter_disconnected
        self.print_fn = print_fn
        self.device = device

    def __call__(self, packed_data_list, start_idx=0):
        kept = []
        for data in packed_data_list:
            if "pos_gen" not in data or "pos_ref" not in data:
                continue
            if self.filter_disconnected and ("." in data.get("smiles", "")):
                continue
            n = _atomic_numbers(data).numel()
            pos_ref = torch.as_tensor(data["pos_ref"]).reshape(-1, n, 3)
            pos_gen = torch.as_tensor(data["pos_gen"]).reshape(-1, n, 3)
            num_gen = pos_ref.shape[0] * self.ratio
            if pos_gen.shape[0] < num_gen:
                continue
            kept.append(dict(data, pos_ref=pos_ref, pos_gen=pos_gen[:num_gen]))
        kept = kept[start_idx:]
        self.print_fn("Filtered: %d / %d" % (len(kept), len(packed_data_list)))
        covr, matr, covp, matp = [], [], [], []
        th = self.thresholds.reshape(1, -1)
        for data in kept:
            m = get_rmsd_confusion_matrix(data, device=self.device)
            ref_min, gen_min = m.min(-1), m.min(0)
            matr.append(ref_min.mean())
            covr.append((ref_min.reshape(-1, 1) <= th).mean(0, keepdims=True))
            matp.append(gen_min.mean())
            covp.append((gen_min.reshape(-1, 1) <= th).mean(0, keepdims=True))
        nt = self.thresholds.size
        return SimpleNamespace(CoverageR=np.vstack(covr) if covr else np.zeros((0, nt)), MatchingR=np.array(matr),
                               thresholds=self.thresholds, CoverageP=np.vstack(covp) if covp else np.zeros((0, nt)),
                               MatchingP=np.array(matp))

    def close(self):
        pass
