"""TEST INFRASTRUCTURE ONLY -- the parity oracle.  Never imported by ``agdiff_b200`` itself;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may use it.

A portable, pure-torch CPU restatement of AGDIFF's sampling hot path
(``DualEncoderEpsNetwork.forward`` + ``langevin_dynamics_sample_diffusion``), written
functionally over a plain ``state_dict`` so it runs on the GPU box where neither
``/root/reference`` nor PyG exist.  Every function cites the reference lines it follows
(paths relative to ``/root/reference/``).

PINNING.  The reference has no tests or golden vectors for this path, so
the oracle is pinned against outputs of the reference itself: ``oracle/make_golden.py`` imports
the UNMODIFIED reference modules in the build container (third-party wheels replaced by the
stand-ins in ``oracle/ref_shims.py``) and commits its outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this restatement against them (bit-exact for edge
lists, <= 2e-6 relative for fp32 tensors) and ``tests/test_host.py::test_oracle_matches_reference_forward`` re-checks
against the live reference whenever ``/root/reference`` is present.

Third-party arithmetic restated here because it is absent from ``/root/reference`` and
un-pinned upstream (README.md:50-59): torch_cluster ``radius`` (CUDA flavour: index-order
scan, first 33 hits), torch_sparse ``coalesce``, torch_scatter ``scatter_add/mean``,
PyG ``MessagePassing(aggr="add")``.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

NUM_BOND_TYPES = 22          # src/agdiff/utils/chem.py:17
MAX_RADIUS_NEIGHBORS = 32    # torch_cluster default used by common.py:217


# ----------------------------------------------------------------------------- schedule
def beta_schedule(cfg) -> np.ndarray:
    """src/agdiff/models/epsnet/dualenc.py:21-51 (float64 numpy)."""
    T = cfg["num_diffusion_timesteps"]
    b0, b1 = cfg["beta_start"], cfg["beta_end"]
    kind = cfg["beta_schedule"]
    if kind == "sigmoid":
        x = np.linspace(-6, 6, T)
        return 1.0 / (np.exp(-x) + 1.0) * (b1 - b0) + b0
    if kind == "linear":
        return np.linspace(b0, b1, T, dtype=np.float64)
    if kind == "quad":
        return np.linspace(b0 ** 0.5, b1 ** 0.5, T, dtype=np.float64) ** 2
    if kind == "const":
        return b1 * np.ones(T, dtype=np.float64)
    if kind == "jsd":
        return 1.0 / np.linspace(T, 1, T, dtype=np.float64)
    raise NotImplementedError(kind)


def alphas_from_cfg(cfg) -> torch.Tensor:
    """dualenc.py:121-125: betas -> fp32, alphas = cumprod(1 - betas) in fp32."""
    betas = torch.from_numpy(beta_schedule(cfg)).float()
    return (1.0 - betas).cumprod(dim=0)


# ----------------------------------------------------------------------------- edges
def bond_order_extension(num_nodes, edge_index, edge_type, order=3):
    """common.py:135-205 restated with boolean reachability instead of int64 matmuls."""
    N = num_nodes
    M = int(edge_index.max().item()) + 1 if edge_index.numel() else 0
    adj = torch.zeros(M, M, dtype=torch.bool)
    adj[edge_index[0], edge_index[1]] = True
    tmat = torch.zeros(M * M, dtype=torch.long)
    tmat.index_add_(0, edge_index[0] * M + edge_index[1], edge_type.long())
    tmat = tmat.view(M, M)
    eye = torch.eye(M, dtype=torch.bool)
    step = (adj | eye).float()
    reach_prev, reach = eye, adj | eye
    order_mat = torch.zeros(M, M, dtype=torch.long)
    order_mat += (reach & ~reach_prev).long()
    for k in range(2, order + 1):
        nxt = (reach.float() @ step) > 0
        order_mat += (nxt & ~reach).long() * k
        reach_prev, reach = reach, nxt
    high = torch.where(order_mat > 1, NUM_BOND_TYPES + order_mat - 1, torch.zeros_like(order_mat))
    tnew = tmat + high
    idx = tnew.nonzero(as_tuple=False).t().contiguous()       # row-major == sorted by row*N+col
    return idx, tnew[idx[0], idx[1]]


def radius_pairs(pos, batch, cutoff, max_num_neighbors=MAX_RADIUS_NEIGHBORS):
    """torch_cluster radius (CUDA rule) behind common.py:217; see SURVEY.md appendix A.
    d2 = (dx*dx + dy*dy) + dz*dz in pos.dtype, strict '<' against r*r, first 33 hits per
    query in ascending candidate index (self included), self then removed."""
    n = pos.size(0)
    r2 = torch.tensor(float(cutoff), dtype=pos.dtype) * torch.tensor(float(cutoff), dtype=pos.dtype)
    rows, cols = [], []
    if n == 0:
        return torch.zeros(2, 0, dtype=torch.long)
    counts = torch.bincount(batch).tolist()
    s = 0
    for c in counts:
        if c == 0:
            continue
        p = pos[s:s + c]
        dx = p[None, :, 0] - p[:, None, 0]
        dy = p[None, :, 1] - p[:, None, 1]
        dz = p[None, :, 2] - p[:, None, 2]
        hit = ((dx * dx + dy * dy) + dz * dz) < r2              # [query, candidate]
        keep = hit & (torch.cumsum(hit.long(), 1) <= max_num_neighbors + 1)
        keep &= ~torch.eye(c, dtype=torch.bool)
        q, j = torch.nonzero(keep, as_tuple=True)
        rows.append(j + s)
        cols.append(q + s)
        s += c
    return torch.stack([torch.cat(rows), torch.cat(cols)])


def union_sorted(n, idx_a, val_a, idx_b, val_b):
    """sparse add + coalesce (common.py:215-231): sort by row*N+col, sum duplicate values."""
    key = torch.cat([idx_a[0] * n + idx_a[1], idx_b[0] * n + idx_b[1]])
    val = torch.cat([val_a, val_b])
    uniq, inv = torch.unique(key, sorted=True, return_inverse=True)
    out = torch.zeros(uniq.numel(), dtype=torch.long)
    out.index_add_(0, inv, val)
    return torch.stack([uniq // n, uniq % n]), out


def build_edges(pos, bond_index, bond_type, batch, cfg, extend_order=True, extend_radius=True):
    """common.py:236-264."""
    n = pos.size(0)
    ei, et = bond_index, bond_type
    if extend_order:
        ei, et = bond_order_extension(n, ei, et, order=cfg["edge_order"])
    if extend_radius:
        rp = radius_pairs(pos, batch, cfg["cutoff"])
        ei, et = union_sorted(n, ei, et.long(), rp, torch.zeros(rp.size(1), dtype=torch.long))
    return ei, et


def edge_lengths(pos, edge_index):
    """geometry.py:5-6."""
    return (pos[edge_index[0]] - pos[edge_index[1]]).norm(dim=-1)


# ----------------------------------------------------------------------------- network pieces
def _lin(sd, key, x):
    b = sd.get(key + ".bias")
    return F.linear(x, sd[key + ".weight"], b)


def _ssp(x, beta):
    """schnet.py:77-80: softplus(beta*x) - ln 2 (shift is an fp32 constant)."""
    return F.softplus(beta * x) - torch.log(torch.tensor(2.0)).to(x.dtype)


def _bn_eval(sd, key, x, eps=1e-5):
    """nn.BatchNorm1d in eval mode (schnet.py:154,158; gin.py:132)."""
    return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"],
                        sd[key + ".weight"], sd[key + ".bias"], False, 0.0, eps)


def edge_encoder(sd, pre, edge_length, edge_type):
    """MLPEdgeEncoder.forward, edge.py:84-103. The softmax is over a size-1 dim."""
    x = F.gelu(_lin(sd, pre + "feature_expansion", edge_length))
    b = sd[pre + "bond_emb.weight"][edge_type]
    p = _lin(sd, pre + "edge_feature_mlp.2", F.gelu(_lin(sd, pre + "edge_feature_mlp.0", torch.cat([x, b], 1))))
    a = _lin(sd, pre + "combination_mlp.2", F.gelu(_lin(sd, pre + "combination_mlp.0", torch.cat([p, b], 1))))
    att = torch.softmax(_lin(sd, pre + "attention.2", torch.tanh(_lin(sd, pre + "attention.0", a))), dim=1)
    return a * att.expand_as(a)


def cfconv(sd, pre, x, edge_index, edge_length, edge_attr, cutoff, smooth):
    """CFConv.forward/message schnet.py:136-162 (+ DistanceWeightingNetwork :90-100)."""
    d = edge_length                                                    # (E,1)
    lw = torch.sigmoid(_lin(sd, pre + "distance_weighting.layer2",
                            F.relu(_lin(sd, pre + "distance_weighting.layer1", d.unsqueeze(-1))))).squeeze(-1)
    if smooth:
        C = 0.5 * (torch.cos(d * torch.pi / cutoff) + 1.0)
        C = C * (d <= cutoff)
    else:
        C = torch.exp(-((d - cutoff) ** 2) / (2 * cutoff ** 2))
    C = C * (d <= cutoff) * (d >= 0.0)
    comb = lw * C.view(-1, 1)
    W = _lin(sd, pre + "nn.2", _ssp(_lin(sd, pre + "nn.0", edge_attr), sd[pre + "nn.1.beta"])) * comb
    h = F.leaky_relu(_bn_eval(sd, pre + "norm1", _lin(sd, pre + "lin1", x)), 0.2)
    msg = h[edge_index[0]] * W
    agg = torch.zeros(x.size(0), msg.size(1), dtype=x.dtype).index_add_(0, edge_index[1], msg)
    return _bn_eval(sd, pre + "norm2", _lin(sd, pre + "lin2", agg))


def schnet_encoder(sd, pre, z, edge_index, edge_length, edge_attr, cfg, collect=None):
    """SchNetEncoder.forward schnet.py:268-282, InteractionBlock :201-216,
    AdaptiveScalingModule :230-234 (avg-pool over a size-1 dim is the identity).
    Embedding has max_norm=10 (:254): looked-up rows with norm > 10 are rescaled by
    10/(norm+1e-7) -- done functionally here, the weights are not mutated."""
    w = sd[pre + "embedding.weight"]
    h = w[z]
    nrm = h.norm(dim=1, keepdim=True)
    h = torch.where(nrm > 10.0, h * (10.0 / (nrm + 1e-7)), h)
    nblk = cfg["num_convs"]
    for k in range(nblk):
        ip = "%sinteractions.%d." % (pre, k)
        p1 = cfconv(sd, ip + "conv1.", h, edge_index, edge_length, edge_attr, cfg["cutoff"], cfg["smooth_conv"])
        p2 = cfconv(sd, ip + "conv2.", h, edge_index, edge_length, edge_attr, cfg["cutoff"], cfg["smooth_conv"])
        xc = _lin(sd, ip + "lin", _ssp(torch.cat([p1, p2], -1), sd[ip + "act.beta"]))
        att = torch.sigmoid(_lin(sd, ip + "attention.2", F.relu(_lin(sd, ip + "attention.0", xc))))
        y = xc * att
        sp = "%sscaling_modules.%d." % (pre, k)
        s = torch.sigmoid(F.linear(F.relu(F.linear(y, sd[sp + "fc.0.weight"])), sd[sp + "fc.2.weight"]))
        h = h + y * s
        if collect is not None:
            collect["schnet_h%d" % k] = h
    return h


def gin_encoder(sd, pre, z, edge_index, edge_attr, cfg, collect=None):
    """GINEncoder.forward gin.py:112-148 with GINEConv :38-69 (activation relu)."""
    x = sd[pre + "node_emb.weight"][z]
    L = cfg["num_convs_local"]
    for k in range(L):
        msg = F.relu(x[edge_index[0]] + edge_attr)
        out = torch.zeros_like(x).index_add_(0, edge_index[1], msg)
        out = out + (1 + sd["%sconvs.%d.eps" % (pre, k)]) * x
        hcur = _lin(sd, "%sconvs.%d.nn.layers.1" % (pre, k), F.relu(_lin(sd, "%sconvs.%d.nn.layers.0" % (pre, k), out)))
        hcur = _bn_eval(sd, "%sbatch_norms.%d" % (pre, k), hcur)
        if k < L - 1:
            hcur = F.relu(hcur)
        x = hcur + x
        if collect is not None:
            collect["gin_h%d" % k] = x
    return x


def pair_mlp(sd, pre, node_attr, edge_index, edge_attr, act="relu"):
    """assemble_atom_pair_feature common.py:106-109 + MultiLayerPerceptron :86-103."""
    f = getattr(F, act)
    hp = torch.cat([node_attr[edge_index[0]] * node_attr[edge_index[1]], edge_attr], -1)
    x = f(_lin(sd, pre + "layers.0", hp))
    x = f(_lin(sd, pre + "layers.1", x))
    return _lin(sd, pre + "layers.2", x)


def forward(sd: Dict[str, torch.Tensor], cfg, atom_type, pos, bond_index, bond_type, batch,
            extend_order=True, extend_radius=True, edges=None, collect: Optional[dict] = None):
    """DualEncoderEpsNetwork.forward dualenc.py:142-251 (return_edges=True form).
    NB (:214) the local branch re-uses edge_encoder_GLOBAL; edge_encoder_local is dead."""
    if edges is None:
        edge_index, edge_type = build_edges(pos, bond_index, bond_type, batch, cfg, extend_order, extend_radius)
        edge_length = edge_lengths(pos, edge_index).unsqueeze(-1)
    else:
        edge_index, edge_type, edge_length = edges
    mask = edge_type > 0
    ea = edge_encoder(sd, "edge_encoder_global.", edge_length, edge_type)
    hg = schnet_encoder(sd, "encoder_global.", atom_type, edge_index, edge_length, ea, cfg, collect)
    eg = pair_mlp(sd, "grad_global_dist_mlp.", hg, edge_index, ea, cfg["mlp_act"])
    li = edge_index[:, mask]
    hl = gin_encoder(sd, "encoder_local.", atom_type, li, ea[mask], cfg, collect)
    el = pair_mlp(sd, "grad_local_dist_mlp.", hl, li, ea[mask], cfg["mlp_act"])
    if collect is not None:
        collect.update(edge_attr=ea, node_global=hg, node_local=hl)
    return eg, el, edge_index, edge_type, edge_length, mask


# ----------------------------------------------------------------------------- sampler pieces
def eq_transform(score_d, pos, edge_index, edge_length):
    """geometry.py:9-17."""
    dd = (1.0 / edge_length) * (pos[edge_index[0]] - pos[edge_index[1]])
    out = torch.zeros_like(pos)
    out.index_add_(0, edge_index[0], dd * score_d)
    out.index_add_(0, edge_index[1], -dd * score_d)
    return out


def clip_norm(vec, limit):
    """dualenc.py:586-589."""
    norm = torch.norm(vec, dim=-1, p=2, keepdim=True)
    return vec * torch.where(norm > limit, limit / norm, torch.ones_like(norm))


def center_pos(pos, batch):
    """dualenc.py:581-583 (scatter_mean)."""
    g = int(batch.max().item()) + 1
    tot = torch.zeros(g, 3, dtype=pos.dtype).index_add_(0, batch, pos)
    cnt = torch.bincount(batch, minlength=g).clamp(min=1).to(pos.dtype)
    return pos - (tot / cnt[:, None])[batch]


def step_scalars(alphas: torch.Tensor, i: int, step_lr: float):
    """dualenc.py:468,532-533: sigma_i, step_size, sqrt(2*step_size) as fp32 0-dim tensors."""
    sigmas = (1.0 - alphas).sqrt() / alphas.sqrt()
    step_size = step_lr * (sigmas[i] / 0.01) ** 2
    return sigmas[i], step_size, torch.sqrt(step_size * 2)


def sample(sd, cfg, atom_type, pos_init, bond_index, bond_type, batch, num_graphs, extend_order,
           extend_radius=True, n_steps=5000, step_lr=1e-6, clip=1000, clip_local=None, clip_pos=None,
           global_start_sigma=float("inf"), w_global=0.2, noise=None, t_start=None, scale_init=True,
           keep_traj=True):
    """langevin_dynamics_sample_diffusion dualenc.py:441-547.  ``noise`` (n_steps,N,3) replaces
    randn_like (:529) so trajectories can be compared; ``t_start`` (default T) lets a window
    i = t_start-1 ... t_start-n_steps be run (the reference always uses t_start = T)."""
    alphas = sd["alphas"]
    T = alphas.numel()
    t_start = T if t_start is None else t_start
    sigmas = (1.0 - alphas).sqrt() / alphas.sqrt()
    pos = pos_init * sigmas[-1] if scale_init else pos_init.clone()
    traj = []
    for s, i in enumerate(range(t_start - 1, t_start - n_steps - 1, -1)):
        eg, el, ei, et, elen, mask = forward(sd, cfg, atom_type, pos, bond_index, bond_type, batch,
                                             extend_order, extend_radius)
        nl = eq_transform(el, pos, ei[:, mask], elen[mask])
        if clip_local is not None:
            nl = clip_norm(nl, clip_local)
        if sigmas[i] < global_start_sigma:
            eg = eg * (1 - mask.view(-1, 1).to(pos.dtype))
            ng = clip_norm(eq_transform(eg, pos, ei, elen), clip)
        else:
            ng = 0
        eps_pos = nl + ng * w_global
        z = torch.randn_like(pos) if noise is None else noise[s]
        step_size = step_lr * (sigmas[i] / 0.01) ** 2
        pos = pos + step_size * eps_pos / sigmas[i] + z * torch.sqrt(step_size * 2)
        if torch.isnan(pos).any():
            raise FloatingPointError()
        pos = center_pos(pos, batch)
        if clip_pos is not None:
            pos = torch.clamp(pos, min=-clip_pos, max=clip_pos)
        if keep_traj:
            traj.append(pos.clone())
    return pos, traj


# ----------------------------------------------------------------------------- helpers for tests
def loss_draws(num_graphs, n_atoms, num_timesteps, rng_seed):
    """the two random draws of get_loss_diffusion in the reference's order (dualenc.py:302-313), from the global CPU generator"""
    torch.manual_seed(rng_seed)
    ts = torch.randint(0, num_timesteps, size=(num_graphs // 2 + 1,))
    ts = torch.cat([ts, num_timesteps - ts - 1], dim=0)[:num_graphs]
    noise = torch.zeros(n_atoms, 3)
    noise.normal_()
    return ts, noise


def loss_diffusion(sd, cfg, atom_type, pos, bond_index, bond_type, batch, num_graphs, time_step, pos_noise, extend_order=True,
                   extend_radius=True):
    """get_loss_diffusion dualenc.py:284-395 with the two random draws injected -> (loss, loss_global, loss_local), each (N, 1)."""
    alphas = alphas_from_cfg(cfg).to(pos.dtype)
    a = alphas.index_select(0, time_step)
    a_pos = a.index_select(0, batch).unsqueeze(-1)
    pos_p = pos + pos_noise.to(pos.dtype) * (1.0 - a_pos).sqrt() / a_pos.sqrt()
    eg, el, ei, et, elen, lmask = forward(sd, cfg, atom_type, pos_p, bond_index, bond_type, batch, extend_order, extend_radius)
    a_edge = a.index_select(0, batch.index_select(0, ei[0])).unsqueeze(-1)
    d_gt = edge_lengths(pos, ei).unsqueeze(-1)
    d_target = (d_gt - elen) / (1.0 - a_edge).sqrt() * a_edge.sqrt()          # is_train_edge == all True (:570-572)
    lm = lmask.unsqueeze(-1)
    gmask = torch.logical_and(torch.logical_or(elen <= cfg["cutoff"], lm), ~lm)
    tgt_g = torch.where(gmask, d_target, torch.zeros_like(d_target))
    eg = torch.where(gmask, eg, torch.zeros_like(eg))
    loss_g = 2 * ((eq_transform(eg, pos_p, ei, elen) - eq_transform(tgt_g, pos_p, ei, elen)) ** 2).sum(-1, keepdim=True)
    li, ll_len = ei[:, lmask], elen[lmask]
    loss_l = 5 * ((eq_transform(el, pos_p, li, ll_len) - eq_transform(d_target[lmask], pos_p, li, ll_len)) ** 2).sum(-1, keepdim=True)
    return loss_g + loss_l, loss_g, loss_l


def perturb_state_dict(sd: Dict[str, torch.Tensor], seed: int = 7) -> Dict[str, torch.Tensor]:
    """Deterministically perturbs the constants that are trivial at init (BN running stats and
    affine, ShiftedSoftplus beta, GIN eps, zero biases) so that folded constants are exercised
    (tier 2).  Keys are visited in sorted order."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(sd):
        v = sd[k].clone()
        if k.startswith("model_global.") or k.startswith("model_local."):
            continue
        if k.endswith("running_mean"):
            v = 0.2 * torch.randn(v.shape, generator=g)
        elif k.endswith("running_var"):
            v = 0.5 + torch.rand(v.shape, generator=g)
        elif (".norm" in k or "batch_norms" in k) and k.endswith(".weight"):
            v = 0.8 + 0.4 * torch.rand(v.shape, generator=g)
        elif (".norm" in k or "batch_norms" in k) and k.endswith(".bias"):
            v = 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith(".beta"):
            v = (0.7 + 0.6 * torch.rand((), generator=g)).to(v.dtype)
        elif k.endswith(".eps"):
            v = 0.3 * torch.rand(v.shape, generator=g)
        elif (k.endswith("lin1.bias") or k.endswith("lin2.bias")) and "conv" in k:
            v = 0.05 * torch.randn(v.shape, generator=g)
        out[k] = v
    return out


def to_dtype(sd, dtype):
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
