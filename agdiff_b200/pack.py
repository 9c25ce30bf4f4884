"""state_dict -> packed device weights.

Done once per model (fp64 on the host), so that no per-edge FLOP is spent on it:
  * BatchNorm1d in eval mode (running stats, eps=1e-5) is folded into the preceding Linear
    (reference schnet.py:153-158, gin.py:131-132);
  * every Linear is stored transposed ([K][N]) for the [K][N]-streaming tile GEMM;
  * the bond-embedding halves of the edge encoder's two 256->128 Linears become per-type tables,
    and edge_feature_mlp.2 -> combination_mlp.0 collapses into one 128x128 matrix (edge.py:84-101);
  * on the global branch ``edge_attr`` is only consumed by Linears (CFConv filter nets
    schnet.py:151, pair MLP common.py:106-109), so combination_mlp.2 is merged into those;
  * the SchNet embedding table has its max_norm=10 renormalisation pre-applied (schnet.py:254).
The attention branch of MLPEdgeEncoder multiplies by a softmax over a size-1 dimension (== 1.0),
``edge_encoder_local`` and ``CFConv.attention`` are never evaluated by the reference forward
(dualenc.py:214, schnet.py:126) -- none of them is packed.

The slot list (names, sizes, order) is owned by the native library and queried at run time.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

BN_EPS = 1e-5


def _f64(t: torch.Tensor) -> np.ndarray:
    return t.detach().to("cpu", torch.float64).numpy()


def _fold_bn(sd, lin_key, bn_key):
    w, b = _f64(sd[lin_key + ".weight"]), _f64(sd[lin_key + ".bias"])
    g, beta = _f64(sd[bn_key + ".weight"]), _f64(sd[bn_key + ".bias"])
    mean, var = _f64(sd[bn_key + ".running_mean"]), _f64(sd[bn_key + ".running_var"])
    s = g / np.sqrt(var + BN_EPS)
    return (w * s[:, None]).T.copy(), (b - mean) * s + beta      # [K][N], [N]


def umma_image(W: np.ndarray) -> np.ndarray:
    """tcgen05 B-operand image of a Linear weight W[N][K] (N = out, K = in, i.e. K-major): fp32 split into
    TF32 values hi = rna(w) and lo = rna(w - hi) (round-to-nearest, like cvt.rna.tf32.f32), each laid out as the canonical
    K-major SWIZZLE_128B shared-memory layout the UMMA descriptor in csrc/tc_filter.cu describes: K in atoms of
    32 floats (128 B rows), per atom N rows of 128 B, 16-byte chunk c of row n stored at chunk c ^ (n % 8).
    Returns [hi image | lo image] as float32 bit patterns."""
    w32 = np.ascontiguousarray(W, dtype=np.float64).astype(np.float32)
    N, K = w32.shape
    assert K % 32 == 0 and N % 8 == 0
    def rna(x):   # cvt.rna.tf32.f32: round the magnitude to 10 mantissa bits, ties away from zero
        return ((x.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    hi = rna(w32)
    lo = rna((w32 - hi).astype(np.float32))
    n = np.arange(N)[:, None]
    k = np.arange(K)[None, :]
    off = (k // 32) * (N * 32) + n * 32 + ((((k % 32) // 4) ^ (n % 8)) * 4) + (k % 4)
    out = np.zeros(2 * N * K, dtype=np.float32)
    out[off.reshape(-1)] = hi.reshape(-1)
    out[N * K + off.reshape(-1)] = lo.reshape(-1)
    return out


def f16_scale_exp(*mats) -> int:
    """power-of-two exponent e with max|w| * 2^e in [2^13, 2^14) over all the given matrices (0 for all-zero input)"""
    m = max(float(np.abs(np.asarray(W, dtype=np.float64).astype(np.float32)).max()) for W in mats)
    e = 0 if not np.isfinite(m) or m == 0.0 else 13 - int(np.floor(np.log2(m)))
    return max(-100, min(100, e))


def umma_image_f16(W: np.ndarray, lo_shift: int = 11, scale_exp=None):
    """tcgen05 kind::f16 B-operand image of a Linear weight W[N][K] for the fp16-split kernels (csrc/tc_filter16.cu).
    The fp32 weight is scaled by a power of two s (max|w|*s in [2^13, 2^14): every part stays a NORMAL fp16 number for
    weights down to 2^-17 of the largest) and split into hi = rn_f16(w*s), lo' = rn_f16((w*s - hi) * 2^lo_shift); each
    part is laid out K-major SWIZZLE_128B: K in atoms of 64 halves (128 B rows), per atom N rows of 128 B, 16-byte
    chunk c of row n stored at chunk c ^ (n % 8).  Returns ([hi image | lo' image] as float32 bit patterns, 1/s)."""
    w32 = np.ascontiguousarray(W, dtype=np.float64).astype(np.float32)
    N, K = w32.shape
    assert K % 64 == 0 and N % 8 == 0
    e = f16_scale_exp(w32) if scale_exp is None else int(scale_exp)
    ws = np.ldexp(w32, e).astype(np.float32)                 # exact (power of two)
    hi = ws.astype(np.float16)
    lo = np.ldexp(ws - hi.astype(np.float32), lo_shift).astype(np.float16)
    n = np.arange(N)[:, None]
    k = np.arange(K)[None, :]
    off = (k // 64) * (N * 64) + n * 64 + ((((k % 64) // 8) ^ (n % 8)) * 8) + (k % 8)
    img = np.zeros(2 * N * K, dtype=np.float16)
    img[off.reshape(-1)] = hi.reshape(-1)
    img[N * K + off.reshape(-1)] = lo.reshape(-1)
    return img.view(np.float32).copy(), float(np.ldexp(1.0, -e))


def fold_state_dict(sd: Dict[str, torch.Tensor], num_convs: int, num_convs_local: int, lo_shift: int = 11) -> Dict[str, np.ndarray]:
    """-> {slot name: float64 array} (operand images: float32 bit patterns)."""
    out: Dict[str, np.ndarray] = {}
    e = "edge_encoder_global."
    bond = _f64(sd[e + "bond_emb.weight"])
    L1, b1 = _f64(sd[e + "edge_feature_mlp.0.weight"]), _f64(sd[e + "edge_feature_mlp.0.bias"])
    L2, b2 = _f64(sd[e + "edge_feature_mlp.2.weight"]), _f64(sd[e + "edge_feature_mlp.2.bias"])
    C1, cb1 = _f64(sd[e + "combination_mlp.0.weight"]), _f64(sd[e + "combination_mlp.0.bias"])
    C2, cb2 = _f64(sd[e + "combination_mlp.2.weight"]), _f64(sd[e + "combination_mlp.2.bias"])
    H = L2.shape[0]
    out["enc.fe_w"] = _f64(sd[e + "feature_expansion.weight"])[:, 0]
    out["enc.fe_b"] = _f64(sd[e + "feature_expansion.bias"])
    out["enc.W1"] = L1[:, :H].T
    out["enc.T1"] = bond @ L1[:, H:].T + b1
    out["enc.M2"] = (C1[:, :H] @ L2).T
    out["enc.T2"] = bond @ C1[:, H:].T + (C1[:, :H] @ b2 + cb1)
    out["enc.C2"] = C2.T
    out["tenc.W1"] = umma_image(L1[:, :H])
    out["tenc.M2"] = umma_image(C1[:, :H] @ L2)
    out["tenc.C2"] = umma_image(C2)
    out["enc.c2b"] = cb2
    hsc = np.zeros(4)
    out["henc.W1"], hsc[0] = umma_image_f16(L1[:, :H], lo_shift)
    out["henc.M2"], hsc[1] = umma_image_f16(C1[:, :H] @ L2, lo_shift)
    out["henc.C2"], hsc[2] = umma_image_f16(C2, lo_shift)
    out["henc.sc"] = hsc

    g = "encoder_global."
    emb = _f64(sd[g + "embedding.weight"]).copy()
    nrm = np.linalg.norm(emb, axis=1, keepdims=True)
    emb = np.where(nrm > 10.0, emb * (10.0 / (nrm + 1e-7)), emb)
    out["sch.emb"] = emb
    for k in range(num_convs):
        ip = "%sinteractions.%d." % (g, k)
        p = "blk%d." % k
        for c, tag in ((1, "a"), (2, "b")):
            cp = "%sconv%d." % (ip, c)
            W, bb = _fold_bn(sd, cp + "lin1", cp + "norm1")
            out[p + "L1" + tag], out[p + "l1%sb" % tag] = W, bb
            out[p + "tL1" + tag] = umma_image(W.T)                           # images take W[N=out][K=in]
            w0, b0 = _f64(sd[cp + "nn.0.weight"]), _f64(sd[cp + "nn.0.bias"])
            out[p + "F1" + tag] = (w0 @ C2).T
            out[p + "f1%sb" % tag] = w0 @ cb2 + b0
            out[p + "F2" + tag] = _f64(sd[cp + "nn.2.weight"]).T
            out[p + "tF1" + tag] = umma_image(w0 @ C2)                       # W[N=out][K=in]
            out[p + "tF2" + tag] = umma_image(_f64(sd[cp + "nn.2.weight"]))
            out[p + "f2%sb" % tag] = _f64(sd[cp + "nn.2.bias"])
            dw = np.zeros(128)
            dw[0:32] = _f64(sd[cp + "distance_weighting.layer1.weight"])[:, 0]
            dw[32:64] = _f64(sd[cp + "distance_weighting.layer1.bias"])
            dw[64:96] = _f64(sd[cp + "distance_weighting.layer2.weight"])[0]
            dw[96] = _f64(sd[cp + "distance_weighting.layer2.bias"])[0]
            out[p + "dw%d" % c] = dw
            W, bb = _fold_bn(sd, cp + "lin2", cp + "norm2")
            out[p + "L2" + tag], out[p + "l2%sb" % tag] = W, bb
            out[p + "tL2" + tag] = umma_image(W.T)
        hsc = np.zeros(4)
        for j, tag in enumerate(("F1a", "F2a", "F1b", "F2b")):
            cp = "%sconv%d." % (ip, 1 if tag.endswith("a") else 2)
            Wm = (_f64(sd[cp + "nn.0.weight"]) @ C2) if tag.startswith("F1") else _f64(sd[cp + "nn.2.weight"])
            out[p + "h" + tag], hsc[j] = umma_image_f16(Wm, lo_shift)
        out[p + "hsc"] = hsc
        nsc = np.zeros(8)
        for c, tag in ((1, "a"), (2, "b")):
            cp = "%sconv%d." % (ip, c)
            W2, _ = _fold_bn(sd, cp + "lin2", cp + "norm2")                  # [K][N]
            out[p + "hL2" + tag], nsc[0 if c == 1 else 2] = umma_image_f16(W2.T, lo_shift)
            W1, _ = _fold_bn(sd, cp + "lin1", cp + "norm1")
            out[p + "hL1" + tag], nsc[5 if c == 1 else 6] = umma_image_f16(W1.T, lo_shift)
        out[p + "hLINa"], nsc[1] = umma_image_f16(_f64(sd[ip + "lin.weight"])[:, :128], lo_shift)
        out[p + "hLINb"], nsc[3] = umma_image_f16(_f64(sd[ip + "lin.weight"])[:, 128:], lo_shift)
        out[p + "hA1"], nsc[4] = umma_image_f16(_f64(sd[ip + "attention.0.weight"]), lo_shift)
        out[p + "hnsc"] = nsc
        out[p + "LIN"] = _f64(sd[ip + "lin.weight"]).T
        out[p + "tLINa"] = umma_image(_f64(sd[ip + "lin.weight"])[:, :128])
        out[p + "tLINb"] = umma_image(_f64(sd[ip + "lin.weight"])[:, 128:])
        out[p + "tA1"] = umma_image(_f64(sd[ip + "attention.0.weight"]))
        out[p + "linb"] = _f64(sd[ip + "lin.bias"])
        out[p + "A1"] = _f64(sd[ip + "attention.0.weight"]).T
        out[p + "a1b"] = _f64(sd[ip + "attention.0.bias"])
        out[p + "a2w"] = _f64(sd[ip + "attention.2.weight"])[0]
        sp = "%sscaling_modules.%d." % (g, k)
        out[p + "S1"] = _f64(sd[sp + "fc.0.weight"]).T
        out[p + "S2"] = _f64(sd[sp + "fc.2.weight"]).T
        out[p + "sc"] = np.array([float(sd[ip + "conv1.nn.1.beta"]), float(sd[ip + "conv2.nn.1.beta"]),
                                  float(sd[ip + "act.beta"]), float(_f64(sd[ip + "attention.2.bias"])[0])])

    for pre, p, merged in (("grad_global_dist_mlp.", "pg.", True), ("grad_local_dist_mlp.", "pl.", False)):
        W0, b0 = _f64(sd[pre + "layers.0.weight"]), _f64(sd[pre + "layers.0.bias"])
        out[p + "P1h"] = W0[:, :H].T
        if merged:
            out[p + "P1e"] = (W0[:, H:] @ C2).T
            out[p + "p1b"] = b0 + W0[:, H:] @ cb2
        else:
            out[p + "P1e"] = W0[:, H:].T
            out[p + "p1b"] = b0
        out[p + "P2"] = _f64(sd[pre + "layers.1.weight"]).T
        tp = "t" + p
        out[tp + "P1h"] = umma_image(W0[:, :H])
        out[tp + "P1e"] = umma_image(W0[:, H:] @ C2 if merged else W0[:, H:])
        out[tp + "P2"] = umma_image(_f64(sd[pre + "layers.1.weight"]))
        hp, hsc = "h" + p, np.zeros(4)
        # both halves of layers.0 accumulate into ONE tensor-core accumulator: common weight scale.  The edge-feature half is
        # added onto the finished h half, where the scale-input-d fold is not available: unscaled lo parts (operand and weight)
        P1e = W0[:, H:] @ C2 if merged else W0[:, H:]
        e1 = f16_scale_exp(W0[:, :H], P1e)
        out[hp + "P1h"], hsc[0] = umma_image_f16(W0[:, :H], lo_shift, e1)
        out[hp + "P1e"], hsc[1] = umma_image_f16(P1e, 0, e1)
        out[hp + "P2"], hsc[2] = umma_image_f16(_f64(sd[pre + "layers.1.weight"]), lo_shift)
        out[hp + "sc"] = hsc
        out[p + "p2b"] = _f64(sd[pre + "layers.1.bias"])
        out[p + "p3w"] = _f64(sd[pre + "layers.2.weight"])[0]
        out[p + "p3b"] = _f64(sd[pre + "layers.2.bias"])

    l = "encoder_local."
    out["gin.emb"] = _f64(sd[l + "node_emb.weight"])
    for k in range(num_convs_local):
        p = "gin%d." % k
        out[p + "G1"] = _f64(sd["%sconvs.%d.nn.layers.0.weight" % (l, k)]).T
        out[p + "g1b"] = _f64(sd["%sconvs.%d.nn.layers.0.bias" % (l, k)])
        W, bb = _fold_bn(sd, "%sconvs.%d.nn.layers.1" % (l, k), "%sbatch_norms.%d" % (l, k))
        out[p + "G2"], out[p + "g2b"] = W, bb
        out[p + "tG1"] = umma_image(_f64(sd["%sconvs.%d.nn.layers.0.weight" % (l, k)]))
        out[p + "tG2"] = umma_image(W.T)
        out[p + "sc"] = np.array([1.0 + float(_f64(sd["%sconvs.%d.eps" % (l, k)])[0])])
    return out


def pack(folded: Dict[str, np.ndarray], slot_names, slot_sizes, align: int = 32):
    """Lay the folded arrays out in the library's slot order -> (float32 buffer, int64 offsets)."""
    offsets = []
    total = 0
    for name, size in zip(slot_names, slot_sizes):
        if name not in folded:
            raise KeyError("packer has no tensor for native weight slot %r" % name)
        arr = folded[name]
        if arr.size != size:
            raise ValueError("slot %s: expected %d floats, packer produced %d" % (name, size, arr.size))
        offsets.append(total)
        total += (size + align - 1) // align * align
    buf = np.zeros(total, dtype=np.float32)
    for name, size, off in zip(slot_names, slot_sizes, offsets):
        arr = folded[name]
        if arr.dtype == np.float32:      # operand images are bit patterns: copy them untouched
            buf[off:off + size] = arr.reshape(-1)
        else:
            buf[off:off + size] = np.ascontiguousarray(arr, dtype=np.float64).reshape(-1).astype(np.float32)
    return buf, np.asarray(offsets, dtype=np.int64)
