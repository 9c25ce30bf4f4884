"""TEST INFRASTRUCTURE (lives under tests/ because it uses the oracle as checker): error of the CUDA forward vs the fp32 reference
(golden) and vs the fp64 oracle, for every arithmetic mode (fp16-split, 3xTF32, FFMA).  Evidence for DESIGN.md section 5 /
profiles/r01_accuracy_tc_vs_ffma.md; run on a GPU box:  python tests/accuracy_report.py"""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))   # util.py
from oracle import agdiff_oracle as O
from util import CONFIGS, golden, make_model, state_dict_cpu

CASES = ["fwd_alanine2_qm9", "fwd_alanine2_far_qm9", "fwd_qm9x6_perturbed", "fwd_drugs_mixed_smooth_perturbed",
         "fwd_drugs_mixed_far_smooth"]

if len(sys.argv) > 1 and sys.argv[1] == "child":
    for name in CASES:
        g = golden(name)
        m = make_model(g["cfg_name"], g["seed"], g["perturb"])
        sd = state_dict_cpu(m)
        m = m.to("cuda:0")
        d = "cuda:0"
        out = m(g["atom_type"].to(d), g["pos"].to(d), g["bond_index"].to(d), g["bond_type"].to(d), g["batch"].to(d), None,
                return_edges=True, extend_order=False)
        with torch.no_grad():
            o64 = O.forward(O.to_dtype(sd, torch.float64), CONFIGS[g["cfg_name"]], g["atom_type"], g["pos"].double(),
                            g["bond_index"], g["bond_type"], g["batch"], extend_order=False)
        row = [name]
        for k, key in ((0, "edge_inv_global"), (1, "edge_inv_local")):
            ours, r32, r64 = out[k].double().cpu(), g[key].double(), o64[k]
            s = float(r64.abs().max())
            row.append("%s: ours-ref32 %.1e  ours-fp64 %.1e  ref32-fp64 %.1e (max|ref| %.2f)" % (
                key.split("_")[-1], float((ours - r32).abs().max()) / s, float((ours - r64).abs().max()) / s,
                float((r32 - r64).abs().max()) / s, s))
        print(" | ".join(row))
    from util import kabsch_free_rmsd
    for name in ["traj_alanine2_high", "traj_alanine2_low", "traj_qm9x6_low_smooth"]:
        g = golden(name)
        m = make_model(g["cfg_name"], g["seed"], 0).to("cuda:0")
        d = "cuda:0"
        n = g["n_steps"]
        noise = torch.randn(n, g["atom_type"].numel(), 3, generator=torch.Generator().manual_seed(g["noise_seed"]))
        pos, traj = m.langevin_dynamics_sample_diffusion(
            g["atom_type"].to(d), g["pos_init"].to(d), g["bond_index"].to(d), g["bond_type"].to(d), g["batch"].to(d),
            int(g["batch"].max()) + 1, extend_order=False, n_steps=n, step_lr=1e-6, clip=1000.0, clip_local=g["clip_local"],
            global_start_sigma=g["global_start_sigma"], w_global=g["w_global"], noise=noise, t_start=g["t_start"],
            scale_init=g["scale_init"])
        r = kabsch_free_rmsd(pos, g["pos_final"], g["batch"])
        print("%s: %d steps, max per-molecule RMSD vs the reference's trajectory %.2e A (|pos|max %.1f A)" % (
            name, n, float(r.max()), float(g["pos_final"].abs().max())))
else:
    names = {"2": "tcgen05, fp16-split two-slot filter kernels", "1": "tcgen05 3xTF32", "0": "fp32 FFMA"}
    for tc, extra in (("2", {}), ("2", {"AGD_F16_LOSHIFT": "0"}), ("1", {}), ("0", {})):
        print("== AGD_TC_FILTERS=%s (%s) %s" % (tc, names[tc], extra or ""))
        env = dict(os.environ, AGD_TC_FILTERS=tc, **extra)
        subprocess.run([sys.executable, __file__, "child"], env=env, check=True)
