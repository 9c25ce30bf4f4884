"""The portable oracle restatement against golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py).  This is what pins the oracle on machines without /root/reference."""
import pytest
import torch

from oracle import agdiff_oracle as O
from util import CONFIGS, assert_close, checksum, golden, make_model, state_dict_cpu

FWD = ["fwd_alanine2_qm9", "fwd_alanine2_far_qm9", "fwd_qm9x6_perturbed", "fwd_drugs_mixed_smooth_perturbed",
       "fwd_drugs_mixed_far_smooth"]


@pytest.mark.parametrize("name", FWD)
def test_forward_matches_reference(name):
    g = golden(name)
    m = make_model(g["cfg_name"], g["seed"], g["perturb"])
    sd = state_dict_cpu(m)
    assert abs(checksum(sd) - g["checksum"]) <= 1e-9 * g["checksum"], "random init differs from the reference's"
    with torch.no_grad():
        eg, el, ei, et, elen, mask = O.forward(sd, CONFIGS[g["cfg_name"]], g["atom_type"], g["pos"], g["bond_index"],
                                               g["bond_type"], g["batch"], extend_order=False)
    assert torch.equal(ei, g["edge_index"]) and torch.equal(et, g["edge_type"])      # integer work: bit-exact
    assert_close(elen, g["edge_length"], rtol=1e-6, what="edge_length")
    assert_close(eg, g["edge_inv_global"], rtol=2e-6, atol_scale=2e-6, what="edge_inv_global")
    assert_close(el, g["edge_inv_local"], rtol=2e-6, atol_scale=2e-6, what="edge_inv_local")


def test_bond_order_extension_matches_reference():
    g = golden("bond_order_ext")
    ei, et = O.bond_order_extension(g["atom_type"].numel(), g["bond_index"], g["bond_type"], 3)
    assert torch.equal(ei, g["ext_index"]) and torch.equal(et, g["ext_type"])


@pytest.mark.parametrize("name", ["traj_alanine2_high", "traj_alanine2_low", "traj_qm9x6_low_smooth"])
def test_trajectory_matches_reference(name):
    g = golden(name)
    m = make_model(g["cfg_name"], g["seed"], 0)
    sd = state_dict_cpu(m)
    assert abs(checksum(sd) - g["checksum"]) <= 1e-9 * g["checksum"]
    noise = torch.randn(g["n_steps"], g["atom_type"].numel(), 3, generator=torch.Generator().manual_seed(g["noise_seed"]))
    with torch.no_grad():
        pos, traj = O.sample(sd, CONFIGS[g["cfg_name"]], g["atom_type"], g["pos_init"], g["bond_index"], g["bond_type"],
                             g["batch"], int(g["batch"].max()) + 1, extend_order=False, n_steps=g["n_steps"],
                             step_lr=1e-6, clip=1000.0, clip_local=g["clip_local"],
                             global_start_sigma=g["global_start_sigma"], w_global=g["w_global"], noise=noise,
                             t_start=g["t_start"], scale_init=g["scale_init"])
    for k, ref in zip(g["traj_steps"].tolist(), g["traj"]):
        assert float((traj[k] - ref).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max())), "step %d" % k
    assert float((pos - g["pos_final"]).abs().max()) <= 1e-4 * max(1.0, float(g["pos_final"].abs().max()))


@pytest.mark.parametrize("name", ["loss_drugs_mixed_smooth", "loss_qm9x6"])
def test_oracle_loss_matches_reference_golden(name):
    """forward value of get_loss_diffusion (dualenc.py:284-395): the restatement with the reference's two random draws re-created
    from the recorded seed reproduces the unmodified reference's per-atom losses"""
    g = golden(name)
    m = make_model(g["cfg_name"], g["seed"], g["perturb"])
    sd = state_dict_cpu(m)
    ts, noise = O.loss_draws(g["num_graphs"], g["atom_type"].numel(), 5000, g["rng_seed"])
    with torch.no_grad():
        loss, lg, ll = O.loss_diffusion(sd, CONFIGS[g["cfg_name"]], g["atom_type"], g["pos"], g["bond_index"], g["bond_type"],
                                        g["batch"], g["num_graphs"], ts, noise, extend_order=False)
    for a, b, what in ((loss, g["loss"], "loss"), (lg, g["loss_global"], "global"), (ll, g["loss_local"], "local")):
        assert a.shape == b.shape
        assert float((a - b).abs().max()) <= 2e-5 * float(b.abs().max()), what
