"""Module-level parity tier (tier 2) and the corners the composite tests only implied: every encoder stage
against the oracle's ``collect`` on the golden inputs of the unmodified reference, per-block node states through reduced-depth
twins, a 100-step Drugs trajectory, one oracle forward at BASELINE size, the two position/score clamps actually clamping,
and the sampling front-end against the oracle per molecule.  Bars as in test_gpu_parity.py, plus the plain statement of what is
met: max|err| <= 5e-5 * max|ref| on every compared tensor."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import agdiff_b200
from agdiff_b200 import frontend, graph, synth
from oracle import agdiff_oracle as O
from util import CONFIGS, assert_close, fp32_noise, golden, kabsch_free_rmsd, make_model, rel_err, state_dict_cpu

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
PLAIN = 5e-5          # max|err| / max|ref| met by every stage and output (north_star asks rtol 1e-4)

FWD = ["fwd_alanine2_qm9", "fwd_alanine2_far_qm9", "fwd_qm9x6_perturbed", "fwd_drugs_mixed_smooth_perturbed",
       "fwd_drugs_mixed_far_smooth"]


def _stages(m, sd, cfg, z, pos, bi, bt, b):
    """(name, cuda tensor, oracle tensor) of every stage the library can export, in the oracle's row order"""
    col = {}
    with torch.no_grad():
        ref = O.forward(sd, cfg, z, pos, bi, bt, b, extend_order=False, collect=col)
    m._renorm_embedding(z.to(DEV))
    m._sync_weights()
    nb = m._prepare(z.to(DEV), bi.to(DEV), bt.to(DEV), b.to(DEV), False)
    out = []
    try:
        res = m._forward_native(nb, pos.to(DEV))
        N, E = z.numel(), res[2].shape[1]
        out.append(("edge_inv_global", res[0].cpu(), ref[0]))
        out.append(("edge_inv_local", res[1].cpu(), ref[1]))
        out.append(("h_global (SchNetEncoder)", nb.fetch("h_global", N * 128).view(N, 128).cpu(), col["node_global"]))
        out.append(("h_local (GINEncoder)", nb.fetch("h_local", N * 128).view(N, 128).cpu(), col["node_local"]))
        lmask = res[3].cpu() > 0
        ea = nb.fetch("ea_local", max(nb.n_local, 1) * 128).view(-1, 128).cpu()
        out.append(("edge_attr local (MLPEdgeEncoder)", ea, col["edge_attr"][lmask][nb.lc_canon.long().cpu()]))
        # CFConv of the LAST block: x = LeakyReLU(BN(lin1 h)) and the aggregated messages (CSC order == atom order)
        k = cfg["num_convs"] - 1
        with torch.no_grad():
            agg_ref = _oracle_agg(sd, cfg, z, pos, bi, bt, b, k)
        agg = nb.fetch("agg", N * 192).view(N, 192).cpu()
        has_in = torch.bincount(res[2][1].cpu(), minlength=N) > 0        # rows of atoms without in-edges are not written (read as zero)
        out.append(("cfconv agg (block %d)" % k, agg[has_in], agg_ref[has_in]))
    finally:
        nb.close()
    return out


def _oracle_agg(sd, cfg, z, pos, bi, bt, b, k):
    """agg_i = sum_e x[src_e] * W_e of block k's two convs, [N, 192] (conv1 | conv2): restated from schnet.py:136-162 with the
    oracle's own pieces"""
    import torch.nn.functional as F
    ei, et = O.build_edges(pos, bi, bt, b, cfg, extend_order=False)
    elen = O.edge_lengths(pos, ei).unsqueeze(-1)
    ea = O.edge_encoder(sd, "edge_encoder_global.", elen, et)
    col = {}
    O.schnet_encoder(sd, "encoder_global.", z, ei, elen, ea, dict(cfg, num_convs=k), col) if k > 0 else None
    w = sd["encoder_global.embedding.weight"]
    h = w[z]
    nrm = h.norm(dim=1, keepdim=True)
    h = torch.where(nrm > 10.0, h * (10.0 / (nrm + 1e-7)), h)
    if k > 0:
        h = col["schnet_h%d" % (k - 1)]
    outs = []
    for conv in ("conv1.", "conv2."):
        pre = "encoder_global.interactions.%d.%s" % (k, conv)
        d = elen
        lw = torch.sigmoid(O._lin(sd, pre + "distance_weighting.layer2", F.relu(O._lin(sd, pre + "distance_weighting.layer1", d.unsqueeze(-1))))).squeeze(-1)
        if cfg["smooth_conv"]:
            C = 0.5 * (torch.cos(d * torch.pi / cfg["cutoff"]) + 1.0)
            C = C * (d <= cfg["cutoff"])
        else:
            C = torch.exp(-((d - cfg["cutoff"]) ** 2) / (2 * cfg["cutoff"] ** 2))
        C = C * (d <= cfg["cutoff"]) * (d >= 0.0)
        Wf = O._lin(sd, pre + "nn.2", O._ssp(O._lin(sd, pre + "nn.0", ea), sd[pre + "nn.1.beta"])) * (lw * C.view(-1, 1))
        x = F.leaky_relu(O._bn_eval(sd, pre + "norm1", O._lin(sd, pre + "lin1", h)), 0.2)
        outs.append(torch.zeros(z.numel(), Wf.size(1), dtype=x.dtype).index_add_(0, ei[1], x[ei[0]] * Wf))
    return torch.cat(outs, 1)


@pytest.mark.parametrize("name", FWD)
def test_every_stage_matches_the_oracle_on_reference_goldens(name):
    g = golden(name)
    cfg = CONFIGS[g["cfg_name"]]
    m = make_model(g["cfg_name"], g["seed"], g["perturb"])
    sd = state_dict_cpu(m)
    m = m.to(DEV)
    args = (g["atom_type"], g["pos"], g["bond_index"], g["bond_type"], g["batch"])
    stages = _stages(m, sd, cfg, *args)
    # the composite outputs also against the golden tensors of the reference itself
    assert rel_err(stages[0][1], g["edge_inv_global"]) < PLAIN and rel_err(stages[1][1], g["edge_inv_local"]) < PLAIN
    for what, a, r in stages:
        assert a.shape == r.shape, "%s: %s vs %s" % (what, tuple(a.shape), tuple(r.shape))
        e = rel_err(a, r)
        assert e < PLAIN, "%s: max|err| / max|ref| = %.2e" % (what, e)


@pytest.mark.parametrize("cfg_name,kind", [("drugs", "drugs"), ("qm9", "qm9")])
def test_per_block_node_states_through_reduced_depth_twins(cfg_name, kind):
    """schnet_h{k} / gin_h{k} of the oracle's ``collect`` for every k: the library keeps h in place, so block k's state is read
    from a twin with num_convs = k + 1 (num_convs_local = k + 1) that shares the first k + 1 blocks' weights."""
    full = make_model(cfg_name, 2021, 7)
    sd_full = state_dict_cpu(full)
    cfg = CONFIGS[cfg_name]
    mols = [graph.extend_bond_order_host(x) for x in (synth.drugs_like(6, seed=5, force_max=False) if kind == "drugs" else synth.qm9_like(12, seed=5))]
    z, bi, bt, b, G = graph.collate(mols, 1)
    pos = torch.randn(z.numel(), 3, generator=torch.Generator().manual_seed(3)) * 2.0
    N = z.numel()
    for k in range(max(cfg["num_convs"], cfg["num_convs_local"])):
        nc, nl = min(k + 1, cfg["num_convs"]), min(k + 1, cfg["num_convs_local"])
        sub = dict(cfg, num_convs=nc, num_convs_local=nl)
        torch.manual_seed(1)
        tw = agdiff_b200.get_model(SimpleNamespace(**sub)).eval()
        tw.load_state_dict({kk: v for kk, v in sd_full.items() if kk in tw.state_dict()}, strict=False)
        sd = state_dict_cpu(tw)
        tw = tw.to(DEV)
        col = {}
        with torch.no_grad():
            O.forward(sd, sub, z, pos, bi, bt, b, extend_order=False, collect=col)
        tw._renorm_embedding(z.to(DEV))
        tw._sync_weights()
        nb = tw._prepare(z.to(DEV), bi.to(DEV), bt.to(DEV), b.to(DEV), False)
        try:
            tw._forward_native(nb, pos.to(DEV))
            hg = nb.fetch("h_global", N * 128).view(N, 128).cpu()
            hl = nb.fetch("h_local", N * 128).view(N, 128).cpu()
        finally:
            nb.close()
        if k < cfg["num_convs"]:
            e = rel_err(hg, col["schnet_h%d" % (nc - 1)])
            assert e < PLAIN, "schnet_h%d: %.2e" % (nc - 1, e)
        if k < cfg["num_convs_local"]:
            e = rel_err(hl, col["gin_h%d" % (nl - 1)])
            assert e < PLAIN, "gin_h%d: %.2e" % (nl - 1, e)


def test_trajectory_drugs_100_steps_vs_oracle():
    """north_star: 100-step trajectories under identical injected noise within 1e-3 A RMSD - Drugs shape, including the
    181-atom molecule, across the global-start boundary (sigma crosses 0.5 at i = 2012)"""
    m = make_model("drugs", 2021, 0)
    sd = state_dict_cpu(m)
    m = m.to(DEV)
    mols = [graph.extend_bond_order_host(x) for x in synth.drugs_like(4, seed=3, force_max=True)]
    z, bi, bt, b, G = graph.collate(mols, 1)
    n_steps, t_start = 100, 2062
    gen = torch.Generator().manual_seed(8)
    pos0 = O.center_pos(torch.randn(z.numel(), 3, generator=gen) * 1.5, b)
    noise = torch.randn(n_steps, z.numel(), 3, generator=gen)
    kw = dict(extend_order=False, n_steps=n_steps, step_lr=1e-6, clip=1000.0, clip_local=20.0, global_start_sigma=0.5,
              w_global=1.0, noise=noise, t_start=t_start, scale_init=False)
    with torch.no_grad():
        ref, _ = O.sample(sd, CONFIGS["drugs"], z, pos0, bi, bt, b, G, keep_traj=False, **kw)
    pos, traj = m.langevin_dynamics_sample_diffusion(z.to(DEV), pos0.to(DEV), bi.to(DEV), bt.to(DEV), b.to(DEV), G, **kw)
    r = kabsch_free_rmsd(pos, ref, b)
    assert len(traj) == n_steps and float(r.max()) <= 1e-3, "RMSD %.3e A" % float(r.max())


def test_forward_at_baseline_size_vs_oracle():
    """BASELINE-sized Drugs batch (416 molecules x 2, incl. the 181-atom one; ~37 k atoms, > 1 M edges) in ONE native forward
    against the CPU oracle evaluated molecule-chunk by molecule-chunk (molecules do not interact)"""
    m = make_model("drugs", 2021, 3)
    sd = state_dict_cpu(m)
    m = m.to(DEV)
    cfg = CONFIGS["drugs"]
    mols = [graph.extend_bond_order_host(x) for x in synth.drugs_like(416, seed=2021)]
    z, bi, bt, b, G = graph.collate(mols, 2)
    pos = torch.randn(z.numel(), 3, generator=torch.Generator().manual_seed(1)) * 1.8
    out = m(z.to(DEV), pos.to(DEV), bi.to(DEV), bt.to(DEV), b.to(DEV), None, return_edges=True, extend_order=False)
    eg, el, ei, et = out[0].cpu(), out[1].cpu(), out[2].cpu(), out[3].cpu()
    assert ei.size(1) > 1_000_000
    counts = torch.bincount(b)
    starts = torch.cumsum(counts, 0) - counts
    ref_g, ref_l, ref_i, ref_t = [], [], [], []
    chunk = 64
    with torch.no_grad():
        for g0 in range(0, G, chunk):
            g1 = min(g0 + chunk, G)
            a0, a1 = int(starts[g0]), int(starts[g1 - 1] + counts[g1 - 1])
            msk = (bi[0] >= a0) & (bi[0] < a1)
            r = O.forward(sd, cfg, z[a0:a1], pos[a0:a1], bi[:, msk] - a0, bt[msk], b[a0:a1] - g0, extend_order=False)
            ref_g.append(r[0]); ref_l.append(r[1]); ref_i.append(r[2] + a0); ref_t.append(r[3])
    ref_g, ref_l, ref_i, ref_t = torch.cat(ref_g), torch.cat(ref_l), torch.cat(ref_i, 1), torch.cat(ref_t)
    assert torch.equal(ei, ref_i) and torch.equal(et, ref_t)          # 1.2 M edges, bit-exact incl. order
    assert rel_err(eg, ref_g) < PLAIN and rel_err(el, ref_l) < PLAIN
    assert_close(eg, ref_g, what="edge_inv_global @ BASELINE size", extra_atol=1e-5)
    assert_close(el, ref_l, what="edge_inv_local @ BASELINE size", extra_atol=1e-4)


def test_global_clip_and_clip_pos_really_clamp():
    """clip_norm of the GLOBAL score (limit well below its norm) and clip_pos (positions beyond the box) both bite, and the
    trajectory still follows the oracle (dualenc.py:506-545,586-589)"""
    m = make_model("qm9", 2021, 0)
    sd = state_dict_cpu(m)
    m = m.to(DEV)
    cfg = CONFIGS["qm9"]
    mols = [graph.extend_bond_order_host(x) for x in synth.qm9_like(6, seed=17)]
    z, bi, bt, b, G = graph.collate(mols, 2)
    n_steps = 12
    gen = torch.Generator().manual_seed(5)
    pos0 = O.center_pos(torch.randn(z.numel(), 3, generator=gen) * 1.6, b)
    noise = torch.randn(n_steps, z.numel(), 3, generator=gen)
    kw = dict(extend_order=False, n_steps=n_steps, step_lr=1e-6, clip=0.05, clip_local=20.0, clip_pos=2.0, global_start_sigma=0.5,
              w_global=1.0, noise=noise, t_start=1800, scale_init=False)
    with torch.no_grad():
        # the clamps are active: the unclipped global score is larger than the limit, and atoms sit outside the box
        eg, el, ei, et, elen, mask = O.forward(sd, cfg, z, pos0, bi, bt, b, extend_order=False)
        ng = O.eq_transform(eg * (1 - mask.view(-1, 1).float()), pos0, ei, elen)
        assert float(ng.norm(dim=-1).max()) > 0.05 and float(pos0.abs().max()) > 2.0
        ref, ref_traj = O.sample(sd, cfg, z, pos0, bi, bt, b, G, keep_traj=True, **kw)
    assert float(ref.abs().max()) == 2.0                               # ... atoms end up ON the clamp
    pos, traj = m.langevin_dynamics_sample_diffusion(z.to(DEV), pos0.to(DEV), bi.to(DEV), bt.to(DEV), b.to(DEV), G, **kw)
    assert float(pos.abs().max()) == 2.0
    for k in (0, n_steps - 1):
        assert float(kabsch_free_rmsd(traj[k], ref_traj[k], b).max()) <= 1e-4
    assert float((pos.cpu() - ref).abs().max()) <= 1e-4


def test_frontend_matches_the_oracle_per_molecule():
    """frontend.sample_conformers (batched) against the reference's loop restated with the oracle: ONE molecule x its samples
    per O.sample call (scripts/test.py:130-164), same pos_init, same injected noise stream per conformer"""
    m = make_model("drugs", 2021, 0)
    with torch.no_grad():   # bounded dynamics without clip_local (see test_frontend_batched_equals_one_molecule_per_call)
        m.grad_local_dist_mlp.layers[2].weight.zero_()
        m.grad_local_dist_mlp.layers[2].bias.fill_(-1.0)
    sd = state_dict_cpu(m)
    m = m.to(DEV)
    cfg = CONFIGS["drugs"]
    mols = synth.drugs_like(4, seed=41, force_max=False)
    n_steps = 6
    noise_of = {}

    def noise_fn(index, n_atoms_total):        # one fixed noise tensor per molecule, shared by both sides
        g = torch.Generator().manual_seed(900 + index)
        noise_of[index] = torch.randn(n_steps, n_atoms_total, 3, generator=g)
        return noise_of[index]
    kw = dict(n_steps=n_steps, global_start_sigma=float("inf"), w_global=0.5)
    res = frontend.sample_conformers(m, mols, 2, noise_fn=noise_fn, max_atoms_per_call=10 ** 9, **kw)
    for i, (r, mol) in enumerate(zip(res, mols)):
        ext = graph.extend_bond_order_host(mol)
        z, bi, bt, b, G = graph.collate([ext], 2)
        pos_init = frontend.initial_positions(i, 2 * mol.num_nodes)
        with torch.no_grad():
            ref, _ = O.sample(sd, cfg, z, pos_init, bi, bt, b, G, False, n_steps=n_steps, step_lr=1e-6, clip=1000.0,
                              global_start_sigma=float("inf"), w_global=0.5, noise=noise_of[i], keep_traj=False)
        assert float(kabsch_free_rmsd(r.pos_gen, ref, b).max()) <= 1e-3, "molecule %d" % i


def test_sampler_without_radius_graph():
    """extend_radius=False (a legal argument of dualenc.py:441-461): bond / 2-hop / 3-hop edges only, no global contribution"""
    m = make_model("qm9", 2021, 0)
    sd = state_dict_cpu(m)
    m = m.to(DEV)
    cfg = CONFIGS["qm9"]
    mols = [graph.extend_bond_order_host(x) for x in synth.qm9_like(5, seed=23)]
    z, bi, bt, b, G = graph.collate(mols, 2)
    n_steps = 10
    gen = torch.Generator().manual_seed(6)
    pos0 = O.center_pos(torch.randn(z.numel(), 3, generator=gen) * 1.5, b)
    noise = torch.randn(n_steps, z.numel(), 3, generator=gen)
    kw = dict(extend_order=False, extend_radius=False, n_steps=n_steps, step_lr=1e-6, clip=1000.0, clip_local=20.0,
              global_start_sigma=0.5, w_global=1.0, noise=noise, t_start=1900, scale_init=False)
    with torch.no_grad():
        ref, _ = O.sample(sd, cfg, z, pos0, bi, bt, b, G, keep_traj=False, **kw)
    pos, traj = m.langevin_dynamics_sample_diffusion(z.to(DEV), pos0.to(DEV), bi.to(DEV), bt.to(DEV), b.to(DEV), G, **kw)
    assert len(traj) == n_steps and float((pos.cpu() - ref).abs().max()) <= 1e-4


def test_nan_report_names_the_molecule_and_stops_early():
    """one molecule of a batch is made to diverge (huge coordinates -> inf distances -> NaN): FloatingPointError carries the
    conformers that went NaN, the others are untouched by it, and the call stops long before its 4000 steps"""
    import time
    m = make_model("qm9", 2021, 0).to(DEV)
    mols = [graph.extend_bond_order_host(x) for x in synth.qm9_like(6, seed=29)]
    z, bi, bt, b, G = graph.collate(mols, 1)
    pos0 = torch.randn(z.numel(), 3, generator=torch.Generator().manual_seed(1))
    pos0[b == 3] *= 3.0e19            # squared distances overflow fp32 in molecule 3 only
    kw = dict(extend_order=False, step_lr=1e-6, clip=1000.0, clip_local=20.0, global_start_sigma=0.5, w_global=1.0, seed=3,
              scale_init=False, return_traj=False)
    t0 = time.perf_counter()
    with pytest.raises(FloatingPointError) as ei:
        m.langevin_dynamics_sample_diffusion(z.to(DEV), pos0.to(DEV), bi.to(DEV), bt.to(DEV), b.to(DEV), G, n_steps=4000, **kw)
    dt = time.perf_counter() - t0
    assert ei.value.bad_graphs == [3] and ei.value.first_nan_step == 0
    t0 = time.perf_counter()
    keep = b != 3
    zz, bb = z[keep], b[keep]
    bb = torch.unique_consecutive(bb, return_inverse=True)[1]
    emask = keep[bi[0]]
    remap = torch.cumsum(keep.long(), 0) - 1
    m.langevin_dynamics_sample_diffusion(zz.to(DEV), pos0[keep].to(DEV), remap[bi[:, emask]].to(DEV), bt[emask].to(DEV), bb.to(DEV), G - 1,
                                         n_steps=4000, **kw)
    full = time.perf_counter() - t0
    # (the diverging call is tried twice - fp16-split kernels, then the 3xTF32 ones after the range flag - and still ends sooner)
    assert dt < 0.7 * full, "no early exit: %.2f s for a NaN at step 0 vs %.2f s for 4000 good steps" % (dt, full)


def test_default_seed_is_fresh_per_call_and_reproducible():
    m = make_model("qm9", 2021, 0).to(DEV)
    z, bi, bt, b, G = graph.collate([graph.extend_bond_order_host(synth.alanine_dipeptide())], 2)
    pos0 = torch.randn(z.numel(), 3, generator=torch.Generator().manual_seed(0))
    args = (z.to(DEV), pos0.to(DEV), bi.to(DEV), bt.to(DEV), b.to(DEV), G)
    kw = dict(extend_order=False, n_steps=5, step_lr=1e-6, clip=1000.0, clip_local=20.0, global_start_sigma=0.5, w_global=1.0,
              return_traj=False)
    torch.manual_seed(77)
    p1, _ = m.langevin_dynamics_sample_diffusion(*args, **kw)
    p2, _ = m.langevin_dynamics_sample_diffusion(*args, **kw)
    torch.manual_seed(77)
    p3, _ = m.langevin_dynamics_sample_diffusion(*args, **kw)
    assert not torch.equal(p1, p2) and torch.equal(p1, p3)


@pytest.mark.parametrize("act", ["gelu", "silu", "tanh", "leaky_relu", "softplus"])
def test_pair_mlp_activations_other_than_relu(act):
    """config field mlp_act (common.py:44-84 takes any torch.nn.functional name): forward against the oracle"""
    cfg = dict(CONFIGS["drugs"], mlp_act=act)
    torch.manual_seed(2021)
    m = agdiff_b200.get_model(SimpleNamespace(**cfg)).eval()
    m.load_state_dict(O.perturb_state_dict(m.state_dict(), seed=4), strict=False)
    sd = state_dict_cpu(m)
    m = m.to(DEV)
    mols = [graph.extend_bond_order_host(x) for x in synth.drugs_like(5, seed=9, force_max=False)]
    z, bi, bt, b, G = graph.collate(mols, 1)
    pos = torch.randn(z.numel(), 3, generator=torch.Generator().manual_seed(2)) * 2.0
    with torch.no_grad():
        ref = O.forward(sd, cfg, z, pos, bi, bt, b, extend_order=False)
    out = m(z.to(DEV), pos.to(DEV), bi.to(DEV), bt.to(DEV), b.to(DEV), None, return_edges=True, extend_order=False)
    assert torch.equal(out[2].cpu(), ref[2])
    ng, nl = fp32_noise(sd, cfg, z, pos, bi, bt, b, ref)
    assert_close(out[0], ref[0], what="edge_inv_global (%s)" % act, extra_atol=4 * ng)
    assert_close(out[1], ref[1], what="edge_inv_local (%s)" % act, extra_atol=4 * nl)


# ------------------------------------------------------------------------------- COV / MAT
def test_kabsch_rmsd_matrix_matches_numpy_oracle():
    """aligned RMSD of every (reference, generated) pair vs the SVD oracle: rotated + translated copies (0), noisy copies,
    a mirror image (must NOT align to 0: proper rotations only), heavy-atom selection, a planar and a 3-atom molecule"""
    from agdiff_b200 import evaluation
    from oracle import kabsch_oracle as K
    rng = np.random.default_rng(0)

    def rot():
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        return q * np.sign(np.linalg.det(q))

    for n, planar in ((3, False), (17, False), (60, False), (181, False), (24, True)):
        base = rng.normal(size=(n, 3)) * 2.0
        if planar:
            base[:, 2] = 0.0
        ref = np.stack([base, base @ rot().T + rng.normal(size=3), base + 0.3 * rng.normal(size=(n, 3))])
        gen = np.stack([base @ rot().T + 5.0, base * np.array([1.0, 1.0, -1.0]), base + 0.05 * rng.normal(size=(n, 3)),
                        rng.normal(size=(n, 3)) * 2.0, base @ rot().T])
        sel = np.sort(rng.choice(n, size=max(3, n // 2), replace=False)) if n > 3 else None
        for s_ in (None, sel):
            want = K.rmsd_confusion_matrix(ref, gen, s_)
            got = evaluation.rmsd_matrix(ref, gen, s_).double().cpu().numpy()
            assert got.shape == want.shape
            assert np.abs(got - want).max() <= 2e-5 * max(1.0, want.max()), (n, planar, np.abs(got - want).max())
        assert got[0, 0] < 1e-4                                 # the rotated + translated copy aligns
        if not planar and n > 3:                                # (three points or a planar molecule ARE their mirror image)
            assert want[0, 1] > 1e-2 and got[0, 1] > 1e-2       # a mirror image does not: proper rotations only


def test_covmat_evaluator_matches_oracle():
    """CovMatEvaluator (covmat.py:76-171): filtering (disconnected smiles, too few generated conformers, missing keys), the
    ratio cut, and COV-R / MAT-R / COV-P / MAT-P against the numpy restatement; evaluate_conf on one molecule"""
    from agdiff_b200 import evaluation
    from oracle import kabsch_oracle as K
    rng = np.random.default_rng(1)
    mols = synth.drugs_like(4, seed=12, force_max=False)
    packed = []
    for k, mol in enumerate(mols):
        n = mol.num_nodes
        n_ref = 2 + k
        ref = rng.normal(size=(n_ref, n, 3)) * 1.5
        gen = np.concatenate([ref + 0.1 * (j + 1) * rng.normal(size=ref.shape) for j in range(3)])   # 3 * n_ref >= ratio * n_ref
        packed.append({"smiles": "C" * (k + 1), "atom_type": torch.as_tensor(mol.atom_type), "pos_ref": torch.tensor(ref).reshape(-1, 3).float(),
                       "pos_gen": torch.tensor(gen).reshape(-1, 3).float()})
    packed.append(dict(packed[0], smiles="CC.O"))                                              # disconnected: dropped
    packed.append(dict(packed[1], pos_gen=packed[1]["pos_gen"][: mols[1].num_nodes]))          # 1 generated < ratio * n_ref: dropped
    packed.append({"smiles": "C", "atom_type": packed[0]["atom_type"], "pos_ref": packed[0]["pos_ref"]})   # no pos_gen: dropped
    log = []
    ev = evaluation.CovMatEvaluator(thresholds=np.arange(0.05, 1.55, 0.05), ratio=2, print_fn=log.append)
    res = ev(packed)
    ev.close()
    assert log[0] == "Filtered: 4 / 7" and res.CoverageR.shape == (4, 30) and res.MatchingP.shape == (4,)
    for k, mol in enumerate(mols):
        n = mol.num_nodes
        heavy = np.nonzero(np.asarray(mol.atom_type) > 1)[0]
        ref = packed[k]["pos_ref"].reshape(-1, n, 3).numpy()
        gen = packed[k]["pos_gen"].reshape(-1, n, 3).numpy()[: 2 * ref.shape[0]]
        conf = K.rmsd_confusion_matrix(ref, gen, heavy)
        # thresholds sit on a 0.05 grid: compare the coverage only where no RMSD is within 1e-4 of a threshold
        cov_r, mat_r, cov_p, mat_p = K.covmat_scores(conf, ev.thresholds)
        safe_r = np.array([np.abs(conf.min(-1) - t).min() > 1e-4 for t in ev.thresholds])
        safe_p = np.array([np.abs(conf.min(0) - t).min() > 1e-4 for t in ev.thresholds])
        assert np.array_equal(res.CoverageR[k][safe_r], cov_r[safe_r]) and np.array_equal(res.CoverageP[k][safe_p], cov_p[safe_p])
        assert abs(res.MatchingR[k] - mat_r) <= 2e-5 and abs(res.MatchingP[k] - mat_p) <= 2e-5
    c, m_ = evaluation.evaluate_conf(packed[2], threshold=0.5)
    conf = K.rmsd_confusion_matrix(packed[2]["pos_ref"].reshape(-1, mols[2].num_nodes, 3).numpy(),
                                   packed[2]["pos_gen"].reshape(-1, mols[2].num_nodes, 3).numpy(), np.nonzero(np.asarray(mols[2].atom_type) > 1)[0])
    assert abs(m_ - conf.min(-1).mean()) <= 2e-5 and abs(c - (conf.min(-1) <= 0.5).mean()) <= 1e-12
    rows = evaluation.print_covmat_results(res, print_fn=lambda s_: None)
    assert len(rows) == 31
