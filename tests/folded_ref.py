"""TEST INFRASTRUCTURE: evaluates the network from the *folded/packed* weights with the same
algebra the CUDA kernels use (per-type tables, merged Linears, BN folded, CSC aggregation order
irrelevant in fp64).  It separates two failure classes on the GPU: if this agrees with the oracle
but the kernels do not, the bug is in a kernel; if this disagrees, the bug is in pack.py."""
from __future__ import annotations

import math

import numpy as np
import torch

LN2 = math.log(2.0)


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).double()


def _gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def _ssp(x, beta):
    y = beta * x
    return torch.where(y > 20.0, y, torch.log1p(torch.exp(torch.clamp(y, max=20.0)))) - LN2


def encoder_g2(f, d, typ):
    x = _gelu(d[:, None] * _t(f["enc.fe_w"])[None, :] + _t(f["enc.fe_b"])[None, :])
    g1 = _gelu(x @ _t(f["enc.W1"]).reshape(128, 128) + _t(f["enc.T1"]).reshape(100, 128)[typ])
    return _gelu(g1 @ _t(f["enc.M2"]).reshape(128, 128) + _t(f["enc.T2"]).reshape(100, 128)[typ])


def edge_weight(f, key, d, cutoff, smooth):
    dw = _t(f[key])
    hdn = torch.relu(d[:, None] * dw[0:32][None, :] + dw[32:64][None, :])
    lw = torch.sigmoid(hdn @ dw[64:96] + dw[96])
    if smooth:
        C = 0.5 * (torch.cos(d * math.pi / cutoff) + 1.0)
    else:
        C = torch.exp(-((d - cutoff) ** 2) / (2 * cutoff ** 2))
    C = C * (d <= cutoff) * (d >= 0)
    return lw * C


def forward(f, cfg, atom_type, edge_index, edge_type, edge_length, collect=None):
    """edge list given (canonical order) -> (edge_inv_global, edge_inv_local)."""
    d = edge_length.double().view(-1)
    src, dst = edge_index[0], edge_index[1]
    N = atom_type.numel()
    g2 = encoder_g2(f, d, edge_type)
    h = _t(f["sch.emb"]).reshape(100, 128)[atom_type]
    leaky = torch.nn.functional.leaky_relu

    def lin1(k, h):
        p = "blk%d." % k
        xa = leaky(h @ _t(f[p + "L1a"]).reshape(128, 128) + _t(f[p + "l1ab"]), 0.2)
        xb = leaky(h @ _t(f[p + "L1b"]).reshape(128, 64) + _t(f[p + "l1bb"]), 0.2)
        return torch.cat([xa, xb], 1)

    xcat = lin1(0, h)
    for k in range(cfg["num_convs"]):
        p = "blk%d." % k
        sc = f[p + "sc"]
        wa = (_ssp(g2 @ _t(f[p + "F1a"]).reshape(128, 128) + _t(f[p + "f1ab"]), sc[0]) @ _t(f[p + "F2a"]).reshape(128, 128)
              + _t(f[p + "f2ab"])) * edge_weight(f, p + "dw1", d, cfg["cutoff"], cfg["smooth_conv"])[:, None]
        wb = (_ssp(g2 @ _t(f[p + "F1b"]).reshape(128, 64) + _t(f[p + "f1bb"]), sc[1]) @ _t(f[p + "F2b"]).reshape(64, 64)
              + _t(f[p + "f2bb"])) * edge_weight(f, p + "dw2", d, cfg["cutoff"], cfg["smooth_conv"])[:, None]
        filt = torch.cat([wa, wb], 1)
        agg = torch.zeros(N, 192, dtype=torch.float64).index_add_(0, dst, xcat[src] * filt)
        v1 = _ssp(agg[:, :128] @ _t(f[p + "L2a"]).reshape(128, 128) + _t(f[p + "l2ab"]), sc[2])
        v2 = _ssp(agg[:, 128:] @ _t(f[p + "L2b"]).reshape(64, 128) + _t(f[p + "l2bb"]), sc[2])
        xc = torch.cat([v1, v2], 1) @ _t(f[p + "LIN"]).reshape(256, 128) + _t(f[p + "linb"])
        gate = torch.sigmoid(torch.relu(xc @ _t(f[p + "A1"]).reshape(128, 64) + _t(f[p + "a1b"])) @ _t(f[p + "a2w"]) + sc[3])
        y = xc * gate[:, None]
        s = torch.sigmoid(torch.relu(y @ _t(f[p + "S1"]).reshape(128, 8)) @ _t(f[p + "S2"]).reshape(8, 128))
        h = h + y * s
        if collect is not None:
            collect["schnet_h%d" % k] = h
        if k + 1 < cfg["num_convs"]:
            xcat = lin1(k + 1, h)

    def pair(p, hn, feat, s_, d_):
        a = torch.relu((hn[s_] * hn[d_]) @ _t(f[p + "P1h"]).reshape(128, 128) + feat @ _t(f[p + "P1e"]).reshape(128, 128)
                       + _t(f[p + "p1b"]))
        b = torch.relu(a @ _t(f[p + "P2"]).reshape(128, 64) + _t(f[p + "p2b"]))
        return b @ _t(f[p + "p3w"]) + _t(f[p + "p3b"])

    eg = pair("pg.", h, g2, src, dst)
    m = edge_type > 0
    ls, ld = src[m], dst[m]
    ea = g2[m] @ _t(f["enc.C2"]).reshape(128, 128) + _t(f["enc.c2b"])
    x = _t(f["gin.emb"]).reshape(100, 128)[atom_type]
    L = cfg["num_convs_local"]
    for k in range(L):
        p = "gin%d." % k
        msg = torch.relu(x[ls] + ea)
        o = torch.zeros_like(x).index_add_(0, ld, msg) + f[p + "sc"][0] * x
        o = torch.relu(o @ _t(f[p + "G1"]).reshape(128, 128) + _t(f[p + "g1b"])) @ _t(f[p + "G2"]).reshape(128, 128) + _t(f[p + "g2b"])
        if k < L - 1:
            o = torch.relu(o)
        x = o + x
    el = pair("pl.", x, ea, ls, ld)
    if collect is not None:
        collect.update(g2=g2, node_global=h, node_local=x, ea_local=ea)
    return eg, el
