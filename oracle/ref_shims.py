"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Pure-torch stand-ins for the third-party wheels the AGDIFF reference imports but
which are not installed offline (torch_geometric, torch_scatter, torch_sparse,
torch_cluster, rdkit via agdiff.utils.chem).  With these registered in
``sys.modules`` the reference's model files import and run UNMODIFIED from
``/root/reference/src``.  This only works in the build
container; ``/root/reference`` does not exist on the GPU box, where the portable
restatement ``oracle/agdiff_oracle.py`` plus ``tests/golden/`` take over.

Each stand-in restates the published semantics of the symbol the reference calls
(call sites cited per function).  Third-party packages shimmed (un-pinned upstream,
README.md:50-59): torch_geometric, torch_scatter, torch_sparse, torch_cluster.
"""
from __future__ import annotations

import inspect
import os
import sys
import types

import torch

REFERENCE_SRC = os.environ.get("AGDIFF_REFERENCE_SRC", "/root/reference/src")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "agdiff", "models"))


# --------------------------------------------------------------------------- torch_scatter
def _scatter_add(src, index, dim=0, out=None, dim_size=None):
    """torch_scatter.scatter_add (geometry.py:12-16): segment sum along ``dim``."""
    assert dim in (0, -src.dim()), "shim only supports dim=0"
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() else 0
    res = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    res.index_add_(0, index, src)
    return res


def _scatter_mean(src, index, dim=0, out=None, dim_size=None):
    """torch_scatter.scatter_mean (dualenc.py:582): sum / max(count, 1)."""
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() else 0
    tot = _scatter_add(src, index, dim=0, dim_size=dim_size)
    cnt = torch.zeros(dim_size, dtype=src.dtype, device=src.device)
    cnt.index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
    cnt = cnt.clamp(min=1)
    return tot / cnt.view((-1,) + (1,) * (src.dim() - 1))


def _scatter_max(src, index, dim=0, out=None, dim_size=None):
    if dim_size is None:
        dim_size = int(index.max().item()) + 1
    res = torch.full((dim_size,) + tuple(src.shape[1:]), torch.iinfo(src.dtype).min
                     if not src.is_floating_point() else float("-inf"), dtype=src.dtype)
    res = res.scatter_reduce(0, index, src, reduce="amax", include_self=True)
    return res, None


def _scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    if reduce in ("sum", "add"):
        return _scatter_add(src, index, dim, out, dim_size)
    if reduce == "mean":
        return _scatter_mean(src, index, dim, out, dim_size)
    raise NotImplementedError(reduce)


# --------------------------------------------------------------------------- torch_sparse
def _coalesce(index, value, m, n, op="add"):
    """torch_sparse.coalesce (common.py:193, transforms.py:62): sort entries by
    row*n+col and sum the values of duplicates."""
    key = index[0] * n + index[1]
    uniq, inv = torch.unique(key, sorted=True, return_inverse=True)
    new_index = torch.stack([uniq // n, uniq % n], dim=0)
    if value is None:
        return new_index, None
    new_val = torch.zeros((uniq.numel(),) + tuple(value.shape[1:]), dtype=value.dtype)
    new_val.index_add_(0, inv, value)
    return new_index, new_val


# --------------------------------------------------------------------------- torch_cluster / PyG nn
def _radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32,
                  flow="source_to_target", num_workers=1):
    """torch_geometric.nn.radius_graph as called at common.py:217 -- CUDA flavour of
    torch_cluster.radius: for every query atom i scan the atoms j of the same molecule
    in ascending index order (j == i included), keep the first
    ``max_num_neighbors + 1`` with ``|x_j - x_i|^2 < r^2`` (strict), then drop j == i.
    Returns [neighbour j ; query i].  The squared distance is evaluated in the
    storage dtype as (dx*dx + dy*dy) + dz*dz without fused multiply-add."""
    assert flow == "source_to_target"
    n = x.size(0)
    if batch is None:
        batch = torch.zeros(n, dtype=torch.long)
    limit = max_num_neighbors if loop else max_num_neighbors + 1
    r2 = torch.tensor(float(r), dtype=x.dtype) * torch.tensor(float(r), dtype=x.dtype)
    rows, cols = [], []
    # per-molecule dense evaluation; molecules are contiguous because batch is sorted
    counts = torch.bincount(batch, minlength=int(batch.max().item()) + 1 if n else 0)
    start = 0
    for c in counts.tolist():
        if c == 0:
            continue
        p = x[start:start + c]
        d = p.unsqueeze(0) - p.unsqueeze(1)          # d[i, j] = x_j - x_i
        dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]
        d2 = (dx * dx + dy * dy) + dz * dz
        hit = d2 < r2                                 # [query i, candidate j]
        rank = torch.cumsum(hit.to(torch.long), dim=1)
        keep = hit & (rank <= limit)
        if not loop:
            keep = keep & ~torch.eye(c, dtype=torch.bool)
        qi, cj = torch.nonzero(keep, as_tuple=True)   # sorted by query, then candidate
        rows.append(cj + start)
        cols.append(qi + start)
        start += c
    if rows:
        return torch.stack([torch.cat(rows), torch.cat(cols)], dim=0)
    return torch.zeros((2, 0), dtype=torch.long)


class _MessagePassing(torch.nn.Module):
    """torch_geometric.nn.MessagePassing subset used at schnet.py:113,156 and
    gin.py:14,57: x_j = x[edge_index[0]]; message(...); sum into edge_index[1]."""

    def __init__(self, aggr="add", **kwargs):
        super().__init__()
        assert aggr == "add"
        self.aggr = aggr

    def propagate(self, edge_index, size=None, **kwargs):
        x = kwargs.get("x")
        x_src, x_dst = (x if isinstance(x, (tuple, list)) else (x, x))
        src, dst = edge_index[0], edge_index[1]
        params = inspect.signature(self.message).parameters
        args = {}
        for name in params:
            if name == "x_j":
                args[name] = x_src.index_select(0, src)
            elif name == "x_i":
                args[name] = x_dst.index_select(0, dst)
            else:
                args[name] = kwargs[name]
        msg = self.message(**args)
        out = torch.zeros((x_dst.size(0),) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device)
        out.index_add_(0, dst, msg)
        return out


# --------------------------------------------------------------------------- PyG utils
def _to_dense_adj(edge_index, batch=None, edge_attr=None, max_num_nodes=None):
    """torch_geometric.utils.to_dense_adj (common.py:179,182): (1, M, M) dense matrix,
    M = edge_index.max()+1, duplicates summed."""
    m = int(edge_index.max().item()) + 1 if edge_index.numel() else 0
    if max_num_nodes is not None:
        m = max_num_nodes
    if edge_attr is None:
        edge_attr = torch.ones(edge_index.size(1), dtype=torch.float)
    adj = torch.zeros((m * m,) + tuple(edge_attr.shape[1:]), dtype=edge_attr.dtype)
    adj.index_add_(0, edge_index[0] * m + edge_index[1], edge_attr)
    return adj.view((1, m, m) + tuple(edge_attr.shape[1:]))


def _dense_to_sparse(adj):
    """torch_geometric.utils.dense_to_sparse (common.py:189-190): row-major nonzeros."""
    assert adj.dim() == 2
    idx = adj.nonzero(as_tuple=False).t().contiguous()
    return idx, adj[idx[0], idx[1]]


class _Data:  # import-only
    pass


class _Batch(_Data):  # import-only
    pass


def install() -> None:
    """Register the stand-in modules and put the reference's src/ on sys.path."""
    if "torch_scatter" in sys.modules and getattr(sys.modules["torch_scatter"], "_agd_shim", False):
        return
    if not reference_available():
        raise RuntimeError("reference sources not found at %s" % REFERENCE_SRC)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m._agd_shim = True
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    mod("torch_scatter", scatter=_scatter, scatter_add=_scatter_add,
        scatter_mean=_scatter_mean, scatter_max=_scatter_max)
    mod("torch_sparse", coalesce=_coalesce, SparseTensor=type("SparseTensor", (), {}),
        matmul=lambda *a, **k: (_ for _ in ()).throw(NotImplementedError()))
    tg = mod("torch_geometric")
    tg.nn = mod("torch_geometric.nn", MessagePassing=_MessagePassing, radius_graph=_radius_graph,
                radius=lambda *a, **k: (_ for _ in ()).throw(NotImplementedError()))
    tg.nn.conv = mod("torch_geometric.nn.conv", MessagePassing=_MessagePassing)
    tg.utils = mod("torch_geometric.utils", to_dense_adj=_to_dense_adj, dense_to_sparse=_dense_to_sparse)
    tg.data = mod("torch_geometric.data", Data=_Data, Batch=_Batch)
    from typing import Optional, Tuple
    tg.typing = mod("torch_geometric.typing", Adj=torch.Tensor, OptTensor=Optional[torch.Tensor],
                    OptPairTensor=Tuple[torch.Tensor, Optional[torch.Tensor]],
                    Size=Optional[Tuple[int, int]])
    # agdiff.utils.chem needs rdkit; only len(BOND_TYPES) == 22 matters (chem.py:17).
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import agdiff  # noqa: F401  (the real package, from /root/reference/src)
    utils_pkg = types.ModuleType("agdiff.utils")
    utils_pkg.__path__ = []  # prevent importing the real utils (rdkit, torchvision, ...)
    sys.modules["agdiff.utils"] = utils_pkg
    chem = mod("agdiff.utils.chem", BOND_TYPES={i: i for i in range(22)})
    utils_pkg.chem = chem


class AttrDict(dict):
    """EasyDict stand-in: attribute access, AttributeError on missing keys (needed for deepcopy)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def load_reference():
    """Returns the reference's own modules (unmodified code)."""
    install()
    import warnings
    warnings.filterwarnings("ignore", category=UserWarning)
    from agdiff.models import common, geometry
    from agdiff.models.epsnet import dualenc, get_model
    return types.SimpleNamespace(common=common, geometry=geometry, dualenc=dualenc, get_model=get_model)
