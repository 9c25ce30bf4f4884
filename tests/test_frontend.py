"""Host logic of the batched sampling front-end (agdiff_b200/frontend.py) against a mocked sampler: grouping, independence
of the grouping, the reference's per-molecule FloatingPointError retry (scripts/test.py:144-181), resume and save_traj."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from agdiff_b200 import frontend, synth


class MockModel:
    """stands in for DualEncoderEpsNetwork: a deterministic 'sampler' whose result for an atom depends only on that atom's
    pos_init, its conformer id and clip_local - like the real one, it is independent of what else is in the batch"""

    def __init__(self, bad_sizes=(), hopeless_sizes=()):
        self.config = SimpleNamespace(edge_order=3)
        self.bad_sizes, self.hopeless_sizes = set(bad_sizes), set(hopeless_sizes)
        self.calls = []

    def parameters(self):
        yield torch.zeros(1)

    def langevin_dynamics_sample_diffusion(self, atom_type, pos_init, bond_index, bond_type, batch, num_graphs, extend_order,
                                           n_steps=5000, clip_local=None, **kw):
        assert extend_order is False and bond_index.size(1) == bond_type.numel()
        counts = torch.bincount(batch, minlength=num_graphs)
        assert num_graphs == kw["mol_gid"].numel() and int(batch.max()) + 1 == num_graphs
        assert int(bond_index.max()) < atom_type.numel()
        self.calls.append((counts.tolist(), clip_local))
        sizes = set(counts.tolist())
        if sizes & self.hopeless_sizes or (clip_local is None and sizes & self.bad_sizes):
            raise FloatingPointError()
        gid = kw["mol_gid"][batch].to(torch.float32).unsqueeze(-1)
        pos = pos_init * 2.0 + gid + (0.5 if clip_local else 0.0)
        traj = [pos.cpu() * (s + 1) for s in range(n_steps)] if kw.get("return_traj", True) else []
        return pos, traj


def _mols():
    return synth.drugs_like(7, seed=3, force_max=False) + synth.qm9_like(5, seed=4)


def test_plan_batches():
    assert frontend.plan_batches([10, 20, 30, 5], [2, 2, 2, 2], 70) == [[0, 1], [2, 3]]
    assert frontend.plan_batches([100, 5], [2, 2], 50) == [[0], [1]]          # oversize molecule gets its own call
    assert frontend.plan_batches([], [], 10) == []


@pytest.mark.parametrize("budget", [10 ** 9, 150, 1])
def test_results_do_not_depend_on_the_grouping(budget):
    mols = _mols()
    ref = frontend.sample_conformers(MockModel(), mols, 2, n_steps=3, max_atoms_per_call=10 ** 9)
    m = MockModel()
    out = frontend.sample_conformers(m, mols, 2, n_steps=3, max_atoms_per_call=budget)
    assert len(out) == len(mols)
    for i, (a, b) in enumerate(zip(out, ref)):
        assert a.index == i and a.num_samples == 2 and a.clip_local is None and not a.failed
        assert a.pos_gen.shape == (2 * mols[i].num_nodes, 3) and torch.equal(a.pos_gen, b.pos_gen)
    if budget == 1:
        assert len(m.calls) == len(mols)


def test_nan_retry_is_per_molecule():
    """one molecule produces NaN without local clipping: only IT is re-sampled with clip_local=20, the others keep their
    unclipped result (what the reference's one-molecule-at-a-time loop gives them)"""
    mols = _mols()
    sizes = [m.num_nodes for m in mols]
    bad = sizes[3]
    assert sizes.count(bad) == 1
    clean = frontend.sample_conformers(MockModel(), mols, 2, n_steps=2)
    m = MockModel(bad_sizes=[bad])
    out = frontend.sample_conformers(m, mols, 2, n_steps=2)
    for i, r in enumerate(out):
        if i == 3:
            assert r.clip_local == 20.0 and not r.failed
            assert torch.equal(r.pos_gen, clean[i].pos_gen + 0.5)
        else:
            assert r.clip_local is None and torch.equal(r.pos_gen, clean[i].pos_gen)
    clipped_calls = [c for c in m.calls if c[1] is not None]
    assert clipped_calls == [([bad, bad], 20.0)]
    assert len(m.calls) <= 2 * int(np.ceil(np.log2(len(mols)))) + 3            # bisection, not one call per molecule


def test_nan_after_retry_marks_failure_and_continues():
    mols = _mols()
    sizes = [m.num_nodes for m in mols]
    out = frontend.sample_conformers(MockModel(hopeless_sizes=[sizes[0]]), mols, 1, n_steps=2, max_atoms_per_call=10 ** 9)
    assert out[0].failed and out[0].pos_gen.numel() == 0
    assert all(not r.failed for r in out[1:])


def test_resume_num_samples_and_traj():
    mols = _mols()
    seen = []
    out = frontend.sample_conformers(MockModel(), mols, lambda i: 1 + i % 3, n_steps=4, save_traj=True, done=[1, 4],
                                     on_result=lambda r: seen.append(r.index))
    assert out[1] is None and out[4] is None and sorted(seen) == [i for i in range(len(mols)) if i not in (1, 4)]
    for i, r in enumerate(out):
        if r is not None:
            assert r.num_samples == 1 + i % 3 and r.pos_gen.shape == (4, r.num_samples * mols[i].num_nodes, 3)


def test_repeat_data_and_batch_round_trip():
    """repeat_data (utils/misc.py:88-90) on a duck-typed Data, and PyG-style Batch -> Molecule records -> the same collate"""
    from agdiff_b200 import graph
    mols = [graph.extend_bond_order_host(m) for m in synth.qm9_like(3, seed=8)]
    z, bi, bt, b, G = graph.collate([mols[1]], 3)
    data = SimpleNamespace(atom_type=torch.as_tensor(mols[1].atom_type), edge_index=torch.as_tensor(mols[1].bond_index),
                           edge_type=torch.as_tensor(mols[1].bond_type))
    rep = frontend.repeat_data(data, 3)
    assert rep.num_graphs == 3 and rep.num_nodes == z.numel()
    assert torch.equal(rep.atom_type, z) and torch.equal(rep.edge_index, bi) and torch.equal(rep.edge_type, bt) and torch.equal(rep.batch, b)
    z, bi, bt, b, G = graph.collate(mols, 1)
    perm = torch.randperm(bi.size(1), generator=torch.Generator().manual_seed(0))       # a Batch need not be sorted
    back = frontend.molecules_from_batch(SimpleNamespace(atom_type=z, edge_index=bi[:, perm], edge_type=bt[perm], batch=b))
    assert len(back) == len(mols)
    for a, m in zip(back, mols):
        assert np.array_equal(a.atom_type, m.atom_type) and np.array_equal(a.bond_index, m.bond_index) and np.array_equal(a.bond_type, m.bond_type)


def test_nan_report_avoids_the_bisection():
    """when the sampler says WHICH conformers went NaN (FloatingPointError.bad_graphs), the good molecules are repeated in one
    call and only the offenders are retried with clip_local - no bisection"""
    mols = _mols()
    bad_size = mols[2].num_nodes

    class Reporting(MockModel):
        def langevin_dynamics_sample_diffusion(self, atom_type, pos_init, bond_index, bond_type, batch, num_graphs, extend_order,
                                               n_steps=5000, clip_local=None, **kw):
            counts = torch.bincount(batch, minlength=num_graphs).tolist()
            if clip_local is None and bad_size in counts:
                self.calls.append((counts, clip_local))
                err = FloatingPointError()
                err.bad_graphs = [g for g, c in enumerate(counts) if c == bad_size]
                raise err
            return super().langevin_dynamics_sample_diffusion(atom_type, pos_init, bond_index, bond_type, batch, num_graphs,
                                                              extend_order, n_steps=n_steps, clip_local=clip_local, **kw)
    m = Reporting()
    out = frontend.sample_conformers(m, mols, 2, n_steps=2, max_atoms_per_call=10 ** 9)
    n_bad = sum(1 for x in mols if x.num_nodes == bad_size)
    # 1 failed call with everything, 1 call with the good ones, then per offender: 1 failing unclipped call + 1 clipped
    assert len(m.calls) == 2 + 2 * n_bad
    for i, r in enumerate(out):
        assert r.clip_local == (20.0 if mols[i].num_nodes == bad_size else None) and not r.failed
