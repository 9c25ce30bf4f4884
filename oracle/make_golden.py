"""TEST INFRASTRUCTURE: generates tests/golden/*.pt by running the UNMODIFIED reference
(/root/reference/src, third-party wheels shimmed by oracle/ref_shims.py) on seeded inputs.
Only runs in the build container; the committed fixtures travel to the GPU box.

    python -m oracle.make_golden

Weights are not stored (5.4 MB): they are regenerated from ``torch.manual_seed(seed)`` through
``get_model`` -- the product twin reproduces the reference's random init bit-for-bit
(tests/test_host.py::test_oracle_matches_reference_forward) and every fixture carries a float64 checksum of the
state_dict so a mismatch is detected on the GPU box.
"""
from __future__ import annotations

import os
import sys

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from agdiff_b200 import graph, synth  # noqa: E402
from oracle import agdiff_oracle as O  # noqa: E402
from oracle import ref_shims  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
CONFIGS = {
    "qm9": dict(type="diffusion", network="dualenc", hidden_dim=128, num_convs=6, num_convs_local=4, cutoff=10.0,
                mlp_act="relu", beta_schedule="sigmoid", beta_start=1.e-7, beta_end=2.e-3,
                num_diffusion_timesteps=5000, edge_order=3, edge_encoder="mlp", smooth_conv=False),
}
CONFIGS["drugs"] = dict(CONFIGS["qm9"], smooth_conv=True)


def checksum(sd):
    return float(sum(v.double().abs().sum() for k, v in sorted(sd.items()) if v.is_floating_point()))


def ref_model(ref, cfg_name, seed, perturb):
    cfg = ref_shims.AttrDict(CONFIGS[cfg_name])
    torch.manual_seed(seed)
    m = ref.get_model(cfg).eval()
    if perturb:
        sd = O.perturb_state_dict(m.state_dict(), seed=perturb)
        m.load_state_dict(sd, strict=False)
    return m


def case_inputs(kind, pos_seed, pos_scale):
    if kind == "alanine2":
        mols, rep = [graph.extend_bond_order_host(synth.alanine_dipeptide())], 2
    elif kind == "qm9x6":
        mols, rep = [graph.extend_bond_order_host(m) for m in synth.qm9_like(6, seed=11)], 1
    elif kind == "drugs_mixed":
        ms = synth.drugs_like(4, seed=5, force_max=False)
        big = synth._random_molecule(__import__("numpy").random.default_rng(3), 70, 38, (6, 6, 7, 8, 16), 4)
        mols, rep = [graph.extend_bond_order_host(m) for m in ms + [big]], 1
    else:
        raise KeyError(kind)
    z, bi, bt, b, G = graph.collate(mols, rep)
    g = torch.Generator().manual_seed(pos_seed)
    pos = torch.randn(z.numel(), 3, generator=g) * pos_scale
    return z, bi, bt, b, G, pos


def gen_forward(ref, name, cfg_name, kind, seed, perturb, pos_seed, pos_scale):
    m = ref_model(ref, cfg_name, seed, perturb)
    z, bi, bt, b, G, pos = case_inputs(kind, pos_seed, pos_scale)
    csum = checksum(m.state_dict())     # before forward: the embedding renorm mutates weights in place
    with torch.no_grad():
        eg, el, ei, et, elen, mask = m(z, pos, bi, bt, b, None, return_edges=True, extend_order=False)
    out = dict(cfg_name=cfg_name, kind=kind, seed=seed, perturb=perturb, pos_seed=pos_seed, pos_scale=pos_scale,
               checksum=csum, atom_type=z, bond_index=bi, bond_type=bt, batch=b, pos=pos,
               edge_inv_global=eg, edge_inv_local=el, edge_index=ei, edge_type=et, edge_length=elen)
    torch.save(out, os.path.join(GOLDEN, name + ".pt"))
    print(name, "N", z.numel(), "E", ei.size(1), "E_loc", int(mask.sum()))


def gen_order(ref, name):
    """_extend_graph_order on raw bond graphs (batched), common.py:135-205."""
    mols = synth.qm9_like(3, seed=4) + synth.drugs_like(2, seed=8, force_max=False) + [synth.alanine_dipeptide()]
    z, bi, bt, b, G = graph.collate(mols, 1)
    ei, et = ref.common._extend_graph_order(z.numel(), bi, bt, 3)
    torch.save(dict(atom_type=z, bond_index=bi, bond_type=bt, batch=b, ext_index=ei, ext_type=et),
               os.path.join(GOLDEN, name + ".pt"))
    print(name, "bonds", bi.size(1), "->", ei.size(1))


def gen_traj(ref, name, cfg_name, kind, seed, n_steps, t_start, clip_local, gss, w_global, pos_scale):
    """n_steps of the reference loop body (dualenc.py:479-545).  t_start == T uses the reference's
    own sampler with torch.randn_like patched; other windows re-run the same statements around the
    reference's forward/eq_transform/clip_norm/center_pos."""
    m = ref_model(ref, cfg_name, seed, 0)
    csum = checksum(m.state_dict())
    z, bi, bt, b, G, pos0 = case_inputs(kind, 77, 1.0)
    g = torch.Generator().manual_seed(1234)
    noise = torch.randn(n_steps, z.numel(), 3, generator=g)
    T = m.num_timesteps
    D = ref.dualenc
    if t_start == T:
        it = iter(noise)
        orig = torch.randn_like
        torch.randn_like = lambda x, *a, **k: next(it)
        try:
            pos, traj = m.langevin_dynamics_sample_diffusion(
                z, pos0, bi, bt, b, G, extend_order=False, n_steps=n_steps, step_lr=1e-6, w_global=w_global,
                global_start_sigma=gss, clip=1000.0, clip_local=clip_local)
        finally:
            torch.randn_like = orig
        pos_init = pos0
        scale_init = True
    else:
        sigmas = (1.0 - m.alphas).sqrt() / m.alphas.sqrt()
        pos = pos0 * pos_scale
        pos = D.center_pos(pos, b)
        pos_init = pos.clone()
        scale_init = False
        traj = []
        with torch.no_grad():
            for s, i in enumerate(range(t_start - 1, t_start - n_steps - 1, -1)):
                eg, el, ei, et, elen, mask = m(z, pos, bi, bt, b, None, return_edges=True, extend_order=False)
                nl = D.eq_transform(el, pos, ei[:, mask], elen[mask])
                if clip_local is not None:
                    nl = D.clip_norm(nl, limit=clip_local)
                if sigmas[i] < gss:
                    eg = eg * (1 - mask.view(-1, 1).float())
                    ng = D.clip_norm(D.eq_transform(eg, pos, ei, elen), limit=1000.0)
                else:
                    ng = 0
                eps_pos = nl + ng * w_global
                step_size = 1e-6 * (sigmas[i] / 0.01) ** 2
                pos = pos + step_size * eps_pos / sigmas[i] + noise[s] * torch.sqrt(step_size * 2)
                assert not torch.isnan(pos).any()
                pos = D.center_pos(pos, b)
                traj.append(pos.clone())
    keep = sorted(set([0, 1, 4, 9, n_steps // 2, n_steps - 1]))
    torch.save(dict(cfg_name=cfg_name, kind=kind, seed=seed, n_steps=n_steps, t_start=t_start, clip_local=clip_local,
                    global_start_sigma=gss, w_global=w_global, checksum=csum, noise_seed=1234,
                    atom_type=z, bond_index=bi, bond_type=bt, batch=b, pos_init=pos_init, scale_init=scale_init,
                    pos_final=pos, traj_steps=torch.tensor(keep), traj=torch.stack([traj[k] for k in keep])),
               os.path.join(GOLDEN, name + ".pt"))
    print(name, "steps", n_steps, "t_start", t_start, "|pos|max", float(pos.abs().max()))


def gen_loss(ref, name, cfg_name, kind, seed, perturb, pos_seed, pos_scale, rng_seed):
    """forward value of get_loss_diffusion (dualenc.py:284-395); its two random draws (noise levels, position noise) come from the
    global CPU generator seeded with ``rng_seed`` - the test re-creates them with the same two calls in the same order"""
    m = ref_model(ref, cfg_name, seed, perturb)
    z, bi, bt, b, G, pos = case_inputs(kind, pos_seed, pos_scale)
    csum = checksum(m.state_dict())
    torch.manual_seed(rng_seed)
    with torch.no_grad():
        loss, lg, ll = m.get_loss(z, pos, bi, bt, b, None, G, return_unreduced_loss=True, extend_order=False)
    torch.save(dict(cfg_name=cfg_name, kind=kind, seed=seed, perturb=perturb, pos_seed=pos_seed, pos_scale=pos_scale, checksum=csum,
                    rng_seed=rng_seed, atom_type=z, bond_index=bi, bond_type=bt, batch=b, pos=pos, num_graphs=G,
                    loss=loss, loss_global=lg, loss_local=ll), os.path.join(GOLDEN, name + ".pt"))
    print(name, "N", z.numel(), "loss mean %.4e global %.4e local %.4e" % (float(loss.mean()), float(lg.mean()), float(ll.mean())))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    ref = ref_shims.load_reference()
    gen_forward(ref, "fwd_alanine2_qm9", "qm9", "alanine2", 2021, 0, 1, 3.0)
    gen_forward(ref, "fwd_alanine2_far_qm9", "qm9", "alanine2", 2021, 0, 2, 9.0)
    gen_forward(ref, "fwd_qm9x6_perturbed", "qm9", "qm9x6", 2021, 7, 3, 2.0)
    gen_forward(ref, "fwd_drugs_mixed_smooth_perturbed", "drugs", "drugs_mixed", 2021, 9, 4, 2.5)
    gen_forward(ref, "fwd_drugs_mixed_far_smooth", "drugs", "drugs_mixed", 2021, 0, 5, 7.0)
    gen_order(ref, "bond_order_ext")
    gen_traj(ref, "traj_alanine2_high", "qm9", "alanine2", 2021, 100, 5000, 20.0, 0.5, 1.0, 1.0)
    gen_traj(ref, "traj_alanine2_low", "qm9", "alanine2", 2021, 100, 100, 20.0, 0.5, 1.0, 1.2)
    gen_traj(ref, "traj_qm9x6_low_smooth", "drugs", "qm9x6", 2021, 40, 1500, 20.0, 0.5, 1.0, 1.5)
    gen_loss(ref, "loss_drugs_mixed_smooth", "drugs", "drugs_mixed", 2021, 9, 6, 1.5, 77)
    gen_loss(ref, "loss_qm9x6", "qm9", "qm9x6", 2021, 0, 7, 1.2, 78)


if __name__ == "__main__":
    main()
