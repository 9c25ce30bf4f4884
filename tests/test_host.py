"""CPU-side tests: host logic, weight packing algebra, the C-ABI library's exported symbols and the
world_size-2 (gloo) sharding path.  No GPU compute."""
import os
import re
import sys

import numpy as np
import pytest
import torch

import folded_ref
from agdiff_b200 import graph, pack, synth
from oracle import agdiff_oracle as O
from util import CONFIGS, ROOT, golden, make_model, state_dict_cpu


def test_library_exports_every_declared_symbol():
    from agdiff_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "agdiff_b200.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|void|const char\*)\s+(agd_[a-z0-9_]+)\s*\(", header, flags=re.M))
    assert len(declared) >= 19
    for name in sorted(declared):
        assert hasattr(lib, name), "libagdiff_b200.so does not export %s" % name
    assert set(_lib.EXPORTS) == declared
    assert lib.agd_abi_version() == 1


def test_no_cpu_path():
    """the product fails loudly without a CUDA device instead of falling back"""
    m = make_model("qm9")
    mol = graph.extend_bond_order_host(synth.alanine_dipeptide())
    z, bi, bt, b, G = graph.collate([mol], 1)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(z, torch.randn(z.numel(), 3), bi, bt, b, None, extend_order=False)
    src = open(os.path.join(ROOT, "agdiff_b200", "epsnet.py")).read() + open(os.path.join(ROOT, "agdiff_b200", "_lib.py")).read()
    assert "oracle" not in src.replace("oracle's", ""), "product code must not import the oracle"


def test_module_contract():
    m = make_model("drugs")
    sd = m.state_dict()
    assert len(sd) == 854
    assert list(sd)[:2] == ["betas", "alphas"]
    assert sd["model_global.1.embedding.weight"].data_ptr() == sd["encoder_global.embedding.weight"].data_ptr()
    assert sum(p.numel() for p in m.parameters()) == 1345720
    assert m.num_timesteps == 5000
    sig = (1.0 - m.alphas).sqrt() / m.alphas.sqrt()
    assert abs(float(sig[4999]) - 12.168) < 2e-3 and abs(float(sig[0]) - 0.00225) < 1e-5
    from types import SimpleNamespace
    import agdiff_b200
    with pytest.raises(NotImplementedError):
        agdiff_b200.get_model(SimpleNamespace(**dict(CONFIGS["qm9"], network="other")))
    with pytest.raises(NotImplementedError):
        agdiff_b200.get_model(SimpleNamespace(**dict(CONFIGS["qm9"], edge_encoder="foo")))
    with pytest.raises(NotImplementedError):
        agdiff_b200.get_model(SimpleNamespace(**dict(CONFIGS["qm9"], beta_schedule="foo")))


def test_step_schedule_matches_reference_arithmetic():
    m = make_model("qm9")
    sigmas, sig, stp, nsc, glb = m.step_schedule(5000, 1e-6, 0.5)
    assert int(glb.sum()) == 2012 and glb[-2012:].all() and not glb[:-2012].any()
    for s in (0, 17, 2500, 4999):
        a, b, c = O.step_scalars(m.alphas.detach(), 4999 - s, 1e-6)
        assert sig[s] == np.float32(float(a)) and stp[s] == np.float32(float(b)) and nsc[s] == np.float32(float(c))


@pytest.mark.parametrize("cfg_name", ["qm9", "drugs"])
def test_packed_algebra_equals_oracle_fp64(cfg_name):
    """BN folding, per-type tables and merged Linears reproduce the network in exact arithmetic"""
    cfg = CONFIGS[cfg_name]
    m = make_model(cfg_name, 2021, perturb=7)
    sd = state_dict_cpu(m)
    mols = [graph.extend_bond_order_host(x) for x in synth.drugs_like(3, seed=5, force_max=False)]
    z, bi, bt, b, G = graph.collate(mols, 1)
    pos = torch.randn(z.numel(), 3, dtype=torch.float64, generator=torch.Generator().manual_seed(1)) * 2.5
    with torch.no_grad():
        eg, el, ei, et, elen, mask = O.forward(O.to_dtype(sd, torch.float64), cfg, z, pos, bi, bt, b, extend_order=False)
    f = pack.fold_state_dict(sd, cfg["num_convs"], cfg["num_convs_local"])
    eg2, el2 = folded_ref.forward(f, cfg, z, ei, et, elen)
    assert float((eg.view(-1) - eg2).abs().max()) < 1e-7 * float(eg.abs().max())
    assert float((el.view(-1) - el2).abs().max()) < 1e-7 * float(el.abs().max())


def test_pack_layout_follows_library_slots():
    from agdiff_b200 import _lib
    m = make_model("qm9")
    f = pack.fold_state_dict(state_dict_cpu(m), 6, 4)
    names = sorted(f)
    sizes = [f[n].size for n in names]
    buf, offs = pack.pack(f, names, sizes)
    assert buf.dtype == np.float32 and all(o % 32 == 0 for o in offs)
    for n, s, o in zip(names, sizes, offs):
        assert np.array_equal(buf[o:o + s], np.asarray(f[n], np.float64).reshape(-1).astype(np.float32))
    with pytest.raises(ValueError):
        pack.pack(f, names, [s + 1 for s in sizes])


@pytest.mark.parametrize("N,K,scale", [(128, 128, 0.09), (64, 128, 3.0e-4), (64, 64, 40.0)])
def test_fp16_split_weight_image(N, K, scale):
    """pack.umma_image_f16: hi + lo' * 2^-S reproduces the power-of-two-scaled fp32 weight to 2^-22, every part is a NORMAL fp16
    number for weights down to 2^-17 of the largest, and element (n, k) sits where the K-major SWIZZLE_128B descriptor of
    csrc/tc16_common.cuh expects it (64-half atoms, 16-byte chunk c of row n at chunk c ^ (n % 8))."""
    rng = np.random.default_rng(3)
    W = (rng.standard_normal((N, K)) * scale).astype(np.float64)
    W[0, 0] = 0.0
    for S in (11, 0):
        img, inv = pack.umma_image_f16(W, S)
        assert img.dtype == np.float32 and img.size == N * K
        halves = img.view(np.float16)
        s = 1.0 / inv
        assert 2.0 ** 13 <= np.abs(W.astype(np.float32)).max() * s < 2.0 ** 14
        n, k = np.arange(N)[:, None], np.arange(K)[None, :]
        off = (k // 64) * (N * 64) + n * 64 + ((((k % 64) // 8) ^ (n % 8)) * 8) + (k % 8)
        hi = halves[off].astype(np.float64)
        lo = halves[N * K + off].astype(np.float64)
        ws = W.astype(np.float32).astype(np.float64) * s
        err = np.abs(hi + lo * 2.0 ** -S - ws)
        assert float((err / np.maximum(np.abs(ws), 2.0 ** -3)).max()) <= 2.0 ** -21
        big = np.abs(ws) >= 2.0 ** -3                      # = 2^-17 of the largest weight or more
        assert np.all(np.abs(hi[big]) >= 2.0 ** -14)       # hi normal
        if S == 11:
            nz = big & (lo != 0)
            assert np.all(np.abs(lo[nz]) >= 2.0 ** -14)    # lo' normal thanks to the 2^S pre-scale
    # a common exponent makes two matrices share one accumulator scale (pair MLP, layers.0)
    e = pack.f16_scale_exp(W, 8 * W)
    _, inv_a = pack.umma_image_f16(W, 11, e)
    _, inv_b = pack.umma_image_f16(8 * W, 0, e)
    assert inv_a == inv_b == 2.0 ** -e


def test_bond_order_host_matches_reference_golden():
    g = golden("bond_order_ext")
    mols = synth.qm9_like(3, seed=4) + synth.drugs_like(2, seed=8, force_max=False) + [synth.alanine_dipeptide()]
    z, bi, bt, b, G = graph.collate([graph.extend_bond_order_host(m) for m in mols], 1)
    assert torch.equal(bi, g["ext_index"]) and torch.equal(bt, g["ext_type"])
    ala = graph.extend_bond_order_host(synth.alanine_dipeptide())
    counts = {t: int((ala.bond_type == t).sum()) for t in (1, 2, 23, 24)}
    assert ala.bond_index.shape[1] == 196 and counts == {1: 38, 2: 4, 23: 72, 24: 82}


def test_synthetic_shapes():
    d = synth.drugs_like(400, seed=2021)
    n = np.array([m.num_nodes for m in d])
    assert n.min() >= 8 and n.max() == 181 and 38 < n.mean() < 50
    q = synth.qm9_like(300, seed=2021)
    nq = np.array([m.num_nodes for m in q])
    assert nq.min() >= 5 and nq.max() <= 29 and 16 < nq.mean() < 20
    for m in d[:20] + q[:20]:
        r, c = m.bond_index
        key = r * m.num_nodes + c
        assert (np.diff(key) > 0).all()                                   # sorted, no duplicates
        und = {(min(a, b), max(a, b)) for a, b in zip(r.tolist(), c.tolist())}
        assert len(und) * 2 == r.size                                     # both directions present


def test_shard_molecules_balanced_and_complete():
    sizes = [m.num_nodes for m in synth.drugs_like(200, seed=1)]
    for w in (1, 2, 4, 8):
        parts = graph.shard_molecules(sizes, w)
        assert sorted(i for p in parts for i in p) == list(range(200))
        load = [sum(graph.molecule_cost(sizes[i]) for i in p) for p in parts]
        assert max(load) <= 1.05 * (sum(load) / w) + graph.molecule_cost(181)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    from agdiff_b200.distributed import sample_sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mols = [graph.extend_bond_order_host(m) for m in synth.qm9_like(7, seed=3)]
    n_tot = sum(m.num_nodes for m in mols) * 2
    pos0 = torch.arange(n_tot * 3, dtype=torch.float32).view(n_tot, 3)

    torch.manual_seed(1000 + rank)            # the usual per-rank seeding: the noise seed must NOT follow it
    seen = {}

    def fake_sampler(z, pos, bi, bt, b, G, mol_gid=None, return_traj=False, **kw):
        # stands in for the CUDA sampler: depends on the rows it is given and on the global conformer id
        assert mol_gid.numel() == G and int(b.max()) + 1 == G
        seen["seed"] = kw.get("seed")
        return pos * 2.0 + mol_gid[b].view(-1, 1).float(), []

    out = sample_sharded(fake_sampler, mols, 2, pos0, "cpu")
    seeds = [None] * world
    dist.all_gather_object(seeds, seen.get("seed"))
    if rank == 0:
        q.put((out, seeds))
    dist.destroy_process_group()


def test_sharded_sampling_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out, seeds = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert seeds[0] is not None and seeds[0] == seeds[1]      # rank 0's seed was broadcast: results do not depend on the rank count
    mols = [graph.extend_bond_order_host(m) for m in synth.qm9_like(7, seed=3)]
    n_tot = sum(m.num_nodes for m in mols) * 2
    pos0 = torch.arange(n_tot * 3, dtype=torch.float32).view(n_tot, 3)
    conf = torch.cat([torch.full((m.num_nodes,), i * 2 + s) for i, m in enumerate(mols) for s in range(2)])
    assert torch.equal(out, pos0 * 2.0 + conf.view(-1, 1).float())           # == the unsharded result, global order


def test_oracle_against_live_reference():
    """when /root/reference is present (build container), re-check the restatement and the twin's
    random init against the unmodified reference code"""
    from oracle import ref_shims
    if not ref_shims.reference_available():
        pytest.skip("reference sources not present on this machine")
    ref = ref_shims.load_reference()
    cfg = ref_shims.AttrDict(CONFIGS["drugs"])
    torch.manual_seed(5)
    rm = ref.get_model(cfg).eval()
    m = make_model("drugs", seed=5)
    a, b_ = rm.state_dict(), m.state_dict()
    assert list(a) == list(b_) and all(torch.equal(a[k], b_[k]) for k in a)
    mols = synth.drugs_like(2, seed=12, force_max=False)
    z, bi, bt, b, G = graph.collate(mols, 1)
    pos = torch.randn(z.numel(), 3, generator=torch.Generator().manual_seed(2)) * 3
    sd = state_dict_cpu(m)
    with torch.no_grad():
        r = rm(z, pos, bi, bt, b, None, return_edges=True)                 # extend_order=True default
        o = O.forward(sd, CONFIGS["drugs"], z, pos, bi, bt, b, extend_order=True)
    assert torch.equal(r[2], o[2]) and torch.equal(r[3], o[3])
    assert float((r[0] - o[0]).abs().max()) <= 2e-6 * float(r[0].abs().max())
    assert float((r[1] - o[1]).abs().max()) <= 2e-6 * float(r[1].abs().max())


def test_reference_checkpoint_loads_without_easydict(tmp_path):
    """scripts/train.py:218-229 pickles an easydict.EasyDict config next to the state_dict; the loader supplies a stand-in for the
    absent package, the config keeps attribute access and the state_dict goes into the product model unchanged"""
    import sys
    import types
    from agdiff_b200 import checkpoint
    assert "easydict" not in sys.modules
    fake = types.ModuleType("easydict")

    class EasyDict(dict):                       # what the real package pickles: a dict subclass living in module `easydict`
        def __init__(self, d=None):
            super().__init__()
            for k, v in (d or {}).items():
                self[k] = EasyDict(v) if isinstance(v, dict) else v

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)
    EasyDict.__module__ = "easydict"
    EasyDict.__qualname__ = "EasyDict"
    fake.EasyDict = EasyDict
    m = make_model("drugs", 2021, perturb=3)
    cfg = EasyDict({"model": dict(CONFIGS["drugs"]), "train": {"seed": 2021, "batch_size": 32}})
    sys.modules["easydict"] = fake
    try:
        torch.save({"config": cfg, "model": m.state_dict(), "iteration": 5000, "avg_val_loss": 0.1}, tmp_path / "5000.pt")
    finally:
        del sys.modules["easydict"]
    ckpt = checkpoint.load_checkpoint(str(tmp_path / "5000.pt"))
    assert "easydict" not in sys.modules
    assert isinstance(ckpt["config"], checkpoint.AttrDict) and ckpt["config"].model.hidden_dim == 128
    assert ckpt["config"].model.smooth_conv is True and ckpt["config"].train.seed == 2021 and ckpt["iteration"] == 5000
    import agdiff_b200
    m2 = agdiff_b200.get_model(ckpt["config"].model)
    m2.load_state_dict(ckpt["model"])
    from util import checksum
    assert checksum(state_dict_cpu(m2)) == checksum(state_dict_cpu(m))
    cfg_path = tmp_path / "drugs.yml"
    cfg_path.write_text("model:\n  network: dualenc\n  hidden_dim: 128\n  beta_start: 1.e-7\ntrain:\n  seed: 2021\n")
    c = checkpoint.load_config(str(cfg_path))
    assert c.model.network == "dualenc" and c.model.beta_start == 1e-7 and c.train.seed == 2021


@pytest.mark.parametrize("ascale", [1.0, 1.0e-3, 3.0e3])
def test_fp16_split_arithmetic_model(ascale):
    """Numerical model of the "3xFP16" scheme of csrc/tc16_common.cuh, emulated in numpy (products exact, fp64 accumulation):
    x = hi + lo' * 2^-11 with fp16 parts, D = (sum a_hi.w_lo' + a_lo'.w_hi) * 2^-11 + sum a_hi.w_hi reproduces the fp32 product to
    ~2^-22 at ANY activation magnitude inside the fp16 range (the 2^11 pre-scale keeps the lo parts normal), whereas unscaled lo
    parts lose precision once they go subnormal - which is why the cross terms are folded with scale-input-d."""
    rng = np.random.default_rng(5)
    M, K, N, S = 256, 128, 64, 11
    a = (rng.standard_normal((M, K)) * ascale).astype(np.float32)
    w = rng.uniform(-0.15, 0.15, (N, K)).astype(np.float32)
    ref = a.astype(np.float64) @ w.astype(np.float64).T
    img, inv = pack.umma_image_f16(w, S)
    e = int(round(-np.log2(inv)))
    ws = np.ldexp(w, e).astype(np.float32)
    w_hi = ws.astype(np.float16).astype(np.float64)
    w_lo = np.ldexp(ws - w_hi.astype(np.float32), S).astype(np.float16).astype(np.float64)

    def split(x, shift):
        hi = x.astype(np.float16).astype(np.float32)
        lo = np.ldexp(x - hi, shift).astype(np.float16)
        return hi.astype(np.float64), lo.astype(np.float64)

    a_hi, a_lo = split(a, S)
    cross = a_hi @ w_lo.T + a_lo @ w_hi.T
    d = (cross * 2.0 ** -S + a_hi @ w_hi.T) * inv
    err = float(np.abs(d - ref).max() / np.abs(ref).max())
    assert err < 3e-7, err                                     # fp32 FFMA itself: ~4e-7 on this problem
    # A/B: unscaled lo parts of the ACTIVATION (weights keep their power-of-two pre-scale)
    a_hi0, a_lo0 = split(a, 0)
    w_lo0 = (ws - w_hi.astype(np.float32)).astype(np.float16).astype(np.float64)
    d0 = (a_hi0 @ w_lo0.T + a_lo0 @ w_hi.T + a_hi0 @ w_hi.T) * inv
    err0 = float(np.abs(d0 - ref).max() / np.abs(ref).max())
    if ascale < 0.01:
        assert err0 > 10 * err                                 # subnormal lo parts: visibly worse
    else:
        assert err0 < 1e-6


def test_kabsch_oracle_properties():
    """the COV/MAT oracle itself: invariance under proper rigid motion, no alignment of mirror images, the closed form for a
    pure scaling, symmetry of the matrix for identical sets, and the COV/MAT bookkeeping of covmat.py:131-150"""
    from oracle import kabsch_oracle as K
    rng = np.random.default_rng(3)
    a = rng.normal(size=(30, 3))
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    q *= np.sign(np.linalg.det(q))
    assert K.kabsch_rmsd(a, a @ q.T + 3.0) < 1e-7
    assert K.kabsch_rmsd(a, a * np.array([1.0, 1.0, -1.0])) > 0.1
    a0 = a - a.mean(0)
    assert abs(K.kabsch_rmsd(a, 1.5 * a) - 0.5 * np.sqrt((a0 ** 2).sum() / 30)) < 1e-9
    confs = rng.normal(size=(4, 30, 3))
    m = K.rmsd_confusion_matrix(confs, confs)
    assert np.allclose(m, m.T, atol=1e-9) and np.abs(np.diag(m)).max() < 1e-7
    conf = np.array([[0.1, 0.9], [0.6, 0.4], [2.0, 3.0]])
    cov_r, mat_r, cov_p, mat_p = K.covmat_scores(conf, np.array([0.5, 1.0]))
    assert np.allclose(cov_r, [2 / 3, 2 / 3]) and abs(mat_r - (0.1 + 0.4 + 2.0) / 3) < 1e-12
    assert np.allclose(cov_p, [1.0, 1.0]) and abs(mat_p - 0.25) < 1e-12


def test_evaluation_has_no_cpu_path():
    from agdiff_b200 import evaluation
    with pytest.raises(RuntimeError):
        evaluation.rmsd_matrix(np.zeros((1, 4, 3)), np.zeros((1, 4, 3)), device="cpu")


def test_local_pair_map_properties():
    """agd_host_local_pairs (the host routine behind the pair mode of the local branch, no CUDA involved): on symmetric bond
    graphs every pair holds exactly the two directions of one bond; with dropped directions, changed types, self loops and an
    unsorted segment the map stays valid - two edges share a pair only if they are each other's reverse with equal type"""
    import ctypes as C
    from agdiff_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(11)

    def run(src, dst, typ, n):
        order = np.lexsort((src, dst))                      # CSC: grouped by destination, sources ascending
        src, dst, typ = (np.ascontiguousarray(a[order], dtype=np.int32) for a in (src, dst, typ))
        ptr = np.zeros(n + 1, np.int32)
        ptr[1:] = np.cumsum(np.bincount(dst, minlength=n))
        of = np.full(src.size, -7, np.int32)
        P = C.c_int32(-1)
        _lib.check(lib.agd_host_local_pairs(src.ctypes.data, dst.ctypes.data, typ.ctypes.data, ptr.ctypes.data, src.size, n,
                                            of.ctypes.data, C.byref(P)))
        return src, dst, typ, of, P.value

    def check(src, dst, typ, of, P):
        assert of.min() >= 0 and of.max() == P - 1 and np.unique(of).size == P
        members = {}
        for e, p in enumerate(of):
            members.setdefault(int(p), []).append(e)
        have = {(int(s), int(d)): int(t) for s, d, t in zip(src, dst, typ)}
        for p, es in members.items():
            assert len(es) <= 2
            if len(es) == 2:
                a, b = es
                assert src[a] == dst[b] and dst[a] == src[b] and typ[a] == typ[b] and src[a] != dst[a]
        for e in range(src.size):                            # completeness: a same-type twin always shares the pair
            tw = have.get((int(dst[e]), int(src[e])))
            if tw is not None and tw == int(typ[e]) and src[e] != dst[e]:
                assert len(members[int(of[e])]) == 2
        return members

    for mol in synth.drugs_like(6, seed=5, force_max=False):
        ext = graph.extend_bond_order_host(mol)
        src, dst, typ = ext.bond_index[0], ext.bond_index[1], ext.bond_type
        s, d, t, of, P = run(src, dst, typ, ext.num_nodes)
        check(s, d, t, of, P)
        assert P * 2 == s.size                                # symmetric graph: exactly half
        keep = rng.random(s.size) > 0.2                       # drop directions, flip some types, add self loops
        t2 = t.copy()
        t2[rng.random(s.size) < 0.1] += 1
        s2 = np.concatenate([s[keep], np.array([0, 3], np.int32)])
        d2 = np.concatenate([d[keep], np.array([0, 3], np.int32)])
        t3 = np.concatenate([t2[keep], np.array([1, 2], np.int32)])
        check(*run(s2, d2, t3, ext.num_nodes)[0:3], *run(s2, d2, t3, ext.num_nodes)[3:5])
    # an empty list is fine
    P = C.c_int32(-1)
    _lib.check(lib.agd_host_local_pairs(None, None, None, None, 0, 4, None, C.byref(P)))
    assert P.value == 0
