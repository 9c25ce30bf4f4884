"""Shared helpers of the test-suite (test infrastructure)."""
from __future__ import annotations

import os
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

BASE_CFG = dict(type="diffusion", network="dualenc", hidden_dim=128, num_convs=6, num_convs_local=4, cutoff=10.0,
                mlp_act="relu", beta_schedule="sigmoid", beta_start=1.e-7, beta_end=2.e-3,
                num_diffusion_timesteps=5000, edge_order=3, edge_encoder="mlp", smooth_conv=False)
CONFIGS = {"qm9": dict(BASE_CFG), "drugs": dict(BASE_CFG, smooth_conv=True)}


def golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def checksum(sd):
    return float(sum(v.double().abs().sum() for k, v in sorted(sd.items()) if v.is_floating_point()))


def make_model(cfg_name, seed=2021, perturb=0):
    """Product twin with the reference's seeded random init (+ the oracle's deterministic
    perturbation of BN stats / betas / eps when ``perturb``)."""
    import agdiff_b200
    from oracle import agdiff_oracle as O
    torch.manual_seed(seed)
    m = agdiff_b200.get_model(SimpleNamespace(**CONFIGS[cfg_name])).eval()
    if perturb:
        m.load_state_dict(O.perturb_state_dict(m.state_dict(), seed=perturb), strict=False)
    return m


def state_dict_cpu(m):
    return {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}


def rel_err(a, b):
    a, b = a.double().cpu().reshape(-1), b.double().cpu().reshape(-1)
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def fp32_noise(sd, cfg, z, pos, bi, bt, b, ref32, extend_order=False):
    """max |fp32 reference - fp64 oracle| for (edge_inv_global, edge_inv_local): the reference's OWN rounding
    noise on this input.  Ill-conditioned inputs (far geometry: pre-activations ~1e3 for outputs ~1) make it
    exceed 1e-5*max|ref|; a parity bar tighter than the reference's own noise is not meaningful, so the forward
    tests allow 4x this on top of rtol 1e-4."""
    from oracle import agdiff_oracle as O
    with torch.no_grad():
        o64 = O.forward(O.to_dtype(sd, torch.float64), cfg, z, pos.double(), bi, bt, b, extend_order=extend_order)
    return (float((ref32[0].double().cpu() - o64[0]).abs().max()), float((ref32[1].double().cpu() - o64[1]).abs().max()))


def assert_close(a, b, rtol=1e-4, atol_scale=1e-5, what="", extra_atol=0.0):
    """|a-b| <= rtol*|b| + atol, atol = atol_scale * max|b| (outputs span 0.1..100) + extra_atol."""
    a, b = a.double().cpu().reshape(-1), b.double().cpu().reshape(-1)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, tuple(a.shape), tuple(b.shape))
    atol = atol_scale * float(b.abs().max().clamp(min=1e-30)) + extra_atol
    bad = (a - b).abs() > (atol + rtol * b.abs())
    assert not bool(bad.any()), "%s: %d/%d elements off, max abs err %.3e (max|ref| %.3e)" % (
        what, int(bad.sum()), a.numel(), float((a - b).abs().max()), float(b.abs().max()))


def kabsch_free_rmsd(a, b, batch):
    """per-molecule RMSD between two trajectories' end points (no alignment: same frame)."""
    d2 = ((a.double().cpu() - b.double().cpu()) ** 2).sum(-1)
    g = int(batch.max()) + 1
    tot = torch.zeros(g, dtype=torch.float64).index_add_(0, batch.cpu(), d2)
    cnt = torch.bincount(batch.cpu(), minlength=g).double()
    return (tot / cnt).sqrt()
