"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes), against the CPU oracle and
the committed golden vectors.  Bar: bit-exact list assignments, neighbour ids AND distances (the GPU
evaluates the oracle's canonical fp32 order); blended features within 1e-5 relative (north star) --
and in practice bit-exact, which is asserted too.
"""
import io
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FLT_MAX = np.float32(3.4028234663852886e38)
REL_TOL = 1e-5   # BASELINE.json north_star: "blended features must match within 1e-5 relative"


def write_oracle_index(oracle, tmp_path, x, cent, name="o.index"):
    oidx = oracle.OracleIndex.from_vectors(x, cent)
    p = str(tmp_path / name)
    oracle.write_index(oidx.to_data(), p)
    return oidx, p


def assert_blend_close(got, want):
    nan_w = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan_w)
    denom = np.maximum(np.abs(want[~nan_w]), 1e-30)
    rel = np.abs(got[~nan_w] - want[~nan_w]) / denom
    assert rel.max(initial=0.0) <= REL_TOL, rel.max()


@pytest.mark.parametrize("name", ["retrieve_v2_768", "retrieve_v1_256"])
def test_golden_search_and_blend(rvcr, oracle, golden, tmp_path, name):
    import torch
    g = golden(name)
    oidx, path = write_oracle_index(oracle, tmp_path, g["x"], g["centroids"])
    index = rvcr.read_index(path)
    assert bool(index) and index.ntotal == len(g["x"]) and index.d == g["x"].shape[1] and index.nprobe == 1
    # big_npy = index.reconstruct_n(0, index.ntotal)   pipeline.py:556
    assert np.array_equal(index.reconstruct_n(0, index.ntotal), g["x"])
    # score, ix = index.search(npy, k=8)               pipeline.py:499 (keyword k)
    D, I = index.search(g["q"], k=8)
    assert D.dtype == np.float32 and I.dtype == np.int64 and D.shape == (len(g["q"]), 8)
    assert np.array_equal(I, g["ix"])
    assert np.array_equal(D, g["score"])
    # quantizer assignment + canonical coarse distance
    lst, dist = index.assign(torch.from_numpy(g["q"]).cuda())
    assert np.array_equal(lst.cpu().numpy(), g["assign"]) and np.array_equal(dist.cpu().numpy(), g["coarse_dist"])
    # fused _retrieve_speaker_embeddings (pipeline.py:497-507): like the reference, a new tensor by default ...
    feats = torch.from_numpy(g["q"].copy()).cuda().unsqueeze(0)
    out = rvcr.retrieve_speaker_embeddings(feats, index, None, float(g["index_rate"]))
    assert out.data_ptr() != feats.data_ptr() and out.shape == feats.shape and np.array_equal(feats[0].cpu().numpy(), g["q"])
    assert np.array_equal(out[0].cpu().numpy(), g["out"], equal_nan=True)
    # ... in place on request (what patch_pipeline installs), refusing a view it could not blend in place
    out = rvcr.retrieve_speaker_embeddings(feats, index, None, float(g["index_rate"]), inplace=True)
    assert out.data_ptr() == feats.data_ptr() and out.shape == feats.shape
    with pytest.raises(ValueError):
        rvcr.retrieve_speaker_embeddings(torch.zeros((1, 8, 2 * index.d), device="cuda")[:, :, ::2], index, None, 0.5, inplace=True)
    got = out[0].cpu().numpy()
    assert_blend_close(got, g["out"])
    assert np.array_equal(got, g["out"], equal_nan=True)          # in fact bit-exact
    # host-buffer variant (the e2e path of bench.py): pageable buffer, then pinned buffer (rows are written
    # straight into the mapped host buffer by the blend kernel, wave by wave)
    h = g["q"].copy()
    index.retrieve_blend_host_(h, float(g["index_rate"]))
    assert np.array_equal(h, g["out"], equal_nan=True)
    hp = torch.from_numpy(g["q"].copy()).pin_memory()
    index.retrieve_blend_host_(hp.numpy(), float(g["index_rate"]))
    assert np.array_equal(hp.numpy(), g["out"], equal_nan=True)
    # the unfused, faiss-shaped call sequence the unmodified reference would run gives the same numbers
    with np.errstate(all="ignore"):
        ref = oracle.tail_numpy_verbatim(D, I, index.reconstruct_n(0, index.ntotal), float(g["index_rate"]), g["q"])
    assert np.array_equal(ref, g["out"], equal_nan=True)


@pytest.mark.parametrize("name", ["retrieve_v2_768", "retrieve_v1_256"])
def test_add_and_write_match_oracle_bytes(rvcr, oracle, golden, tmp_path, name):
    """extract_index.py:62-71 with given centroids: factory, nprobe=1, add in batches, write_index."""
    g = golden(name)
    x, cent = g["x"], g["centroids"]
    oidx, opath = write_oracle_index(oracle, tmp_path, x, cent)
    index = rvcr.index_factory(x.shape[1], f"IVF{len(cent)},Flat")
    rvcr.extract_index_ivf(index).nprobe = 1
    assert not index.is_trained
    with pytest.raises(rvcr.RvcrError):
        index.add(x[:4])
    index.set_centroids(cent)
    for i in range(0, len(x), 97):                       # ragged batches
        index.add(x[i:i + 97])
    assert index.ntotal == len(x)
    assert np.array_equal(index.list_sizes(), np.diff(oidx.list_off))
    p = str(tmp_path / "g.index")
    rvcr.write_index(index, p)
    assert open(p, "rb").read() == open(opath, "rb").read()
    # searching the freshly built index == searching the file-loaded one
    D, I = index.search(g["q"], k=8)
    assert np.array_equal(I, g["ix"]) and np.array_equal(D, g["score"])
    # adding more after a search (lists grow, order stays ascending id)
    index.add(x[:50])
    o2 = oracle.OracleIndex.from_vectors(np.concatenate([x, x[:50]]), cent)
    D2, I2 = index.search(g["q"], k=8)
    Do, Io = o2.search(g["q"], k=8)
    assert np.array_equal(I2, Io) and np.array_equal(D2, Do)


def _problem(rvcr, n, d, nq, nlist, **kw):
    from importlib import import_module
    synth = import_module("codename-rvc-fork-3_b200.synth")
    return synth.make_host_problem(n, d, nq, nlist, **kw)


@pytest.mark.parametrize("cfg", [
    dict(n=50_000, d=768, nq=500, nlist=1282),                  # C1 (BASELINE.json configs[0])
    dict(n=50_000, d=768, nq=500, nlist=1282, iid=True),        # C1, skewed lists: hundreds of lists < 8
    dict(n=200_000, d=256, nq=3000, nlist=5128),                # C2 (configs[1], RVC v1 256-d)
])
def test_config_parity_with_oracle(rvcr, oracle, tmp_path, cfg):
    import torch
    pr = _problem(rvcr, **cfg)
    a_x, _ = oracle.coarse_assign(pr.x, pr.centroids)
    off, vecs, ids = oracle.build_lists(pr.x, a_x, cfg["nlist"])
    oidx = oracle.OracleIndex(cfg["d"], pr.centroids, off, vecs, ids, cfg["n"])
    index = rvcr.index_factory(cfg["d"], f"IVF{cfg['nlist']},Flat")
    index.nprobe = 1
    index.set_centroids(pr.centroids)
    for i in range(0, cfg["n"], 8192):                   # extract_index.py:67-69
        index.add(pr.x[i:i + 8192])
    assert np.array_equal(index.list_sizes(), np.diff(off))       # list assignment of every vector
    vg, ig, og, _ = index.device_views()
    assert np.array_equal(ig.cpu().numpy(), ids)                  # per-list ascending-id order
    assert np.array_equal(vg.cpu().numpy(), vecs)
    q = pr.q
    a_q, cd = oracle.coarse_assign(q, pr.centroids)
    lst, dist = index.assign(torch.from_numpy(q).cuda())
    assert np.array_equal(lst.cpu().numpy(), a_q) and np.array_equal(dist.cpu().numpy(), cd)
    Do, Io = oracle.scan_topk(q, a_q, off, vecs, ids)
    D, I = index.search(q, k=8)
    assert np.array_equal(I, Io) and np.array_equal(D, Do)
    if cfg.get("iid"):
        assert (Io < 0).any()
    for rate in (0.75, 0.3, 1.0):
        with np.errstate(all="ignore"):
            want = oracle.tail(Do, Io, pr.x, rate, q)
        f = torch.from_numpy(q.copy()).cuda()
        index.retrieve_blend_(f, rate)
        got = f.cpu().numpy()
        assert_blend_close(got, want)
        assert np.array_equal(got, want, equal_nan=True)
        hp = torch.from_numpy(q.copy()).pin_memory()                  # pinned host path: multi-wave write-back
        index.retrieve_blend_host_(hp.numpy(), rate)
        assert np.array_equal(hp.numpy(), want, equal_nan=True)
    assert index.get_option("scan_impl") == 0                         # default: auto (exact FFMA scan for C1's 500 queries, tcgen05 for C2's 3 000)
    for impl in (1, 2, 3, 4, 5):                                      # every scan kernel, tcgen05 included, is bit-identical
        index.set_option("scan_impl", impl)
        Dk, Ik = index.search(q, k=8)
        assert np.array_equal(Ik, Io) and np.array_equal(Dk, Do)
        f = torch.from_numpy(q.copy()).cuda()
        index.retrieve_blend_(f, rate)
        assert np.array_equal(f.cpu().numpy(), want, equal_nan=True)
    # and the coarse step: forced tcgen05 vs forced exact FFMA vs auto
    la, da = index.assign(torch.from_numpy(q).cuda())
    for mode in (1, 2):
        index.set_option("coarse_mode", mode)
        lm, dm = index.assign(torch.from_numpy(q).cuda())
        assert torch.equal(lm, la) and torch.equal(dm, da)
    index.set_option("coarse_mode", 0)
    index.set_option("scan_impl", 0)


def test_coarse_splits_and_ties(rvcr, oracle):
    """few queries x many centroids (centroid range split across CTAs, atomicMin merge); duplicate
    centroids -> lowest id; a NaN query -> list -1 and an all-pad result row."""
    import torch
    rng = np.random.default_rng(11)
    d, nlist = 256, 20_000
    cent = (rng.standard_normal((nlist, d)) * 0.33).astype(np.float32)
    cent[15_000] = cent[123]
    cent[19_999] = cent[0]
    q = np.concatenate([cent[[123, 15_000, 0, 19_999]] + np.float32(1e-3),
                        (rng.standard_normal((37, d)) * 0.33).astype(np.float32)]).astype(np.float32)
    q[7, 5] = np.nan
    index = rvcr.index_factory(d, f"IVF{nlist},Flat")
    index.set_centroids(cent)
    index.set_option("coarse_mode", 1)
    lst, dist = index.assign(torch.from_numpy(q).cuda())
    a, cd = oracle.coarse_assign(q, cent)
    assert a[0] == 123 and a[1] == 123 and a[2] == 0 and a[3] == 0 and a[7] == -1
    assert np.array_equal(lst.cpu().numpy(), a)
    ok = a >= 0
    assert np.array_equal(dist.cpu().numpy()[ok], cd[ok])
    x = (rng.standard_normal((1000, d)) * 0.33).astype(np.float32)
    index.add(x)
    D, I = index.search(q, k=8)
    assert (I[7] == -1).all() and (D[7] == FLT_MAX).all()


def test_list_length_edges_and_exact_match(rvcr, oracle):
    """lists of length 0, 1, 7, 8, 9 and > 32 (two row blocks); a query equal to a stored vector gives
    score 0 -> NaN row (reference behaviour); duplicates keep the earlier position."""
    import torch
    rng = np.random.default_rng(5)
    d = 128
    sizes = [0, 1, 7, 8, 9, 33, 64, 65, 100, 2]
    cent = (rng.standard_normal((len(sizes), d)) * 3).astype(np.float32)
    rows, assign = [], []
    for l, n in enumerate(sizes):
        rows.append(cent[l] + (rng.standard_normal((n, d)) * 0.1).astype(np.float32))
        assign += [l] * n
    x = np.concatenate(rows).astype(np.float32)
    perm = rng.permutation(len(x))
    x, assign = x[perm], np.array(assign, np.int32)[perm]
    x[np.where(assign == 8)[0][5]] = x[np.where(assign == 8)[0][2]]       # duplicate inside list 8
    oidx = oracle.OracleIndex.from_vectors(x, cent)
    assert np.array_equal(np.diff(oidx.list_off), sizes)
    index = rvcr.index_factory(d, f"IVF{len(sizes)},Flat")
    index.set_centroids(cent)
    index.add(x)
    q = np.concatenate([cent + np.float32(0.01), x[np.where(assign == 8)[0][[2, 5]]], x[:3]]).astype(np.float32)
    Do, Io = oidx.search(q)
    D, I = index.search(q)
    assert np.array_equal(I, Io) and np.array_equal(D, Do)
    assert (I[0] == -1).all()                              # empty list: all padding
    assert (I[1, 1:] == -1).all() and I[1, 0] >= 0
    assert (I[2, 7] == -1) and (I[2, :7] >= 0).all()
    assert (I[3] >= 0).all()
    with np.errstate(all="ignore"):
        want = oracle.tail(Do, Io, x, 0.75, q)
    f = torch.from_numpy(q.copy()).cuda()
    index.retrieve_blend_(f, 0.75)
    got = f.cpu().numpy()
    assert np.array_equal(got, want, equal_nan=True)
    assert np.isnan(got[0]).all()                          # all-pad row -> 0/0
    assert np.isnan(got[len(sizes)]).all()                 # exact match -> inf/inf
    assert np.isfinite(got[1]).all()


def test_api_errors_and_empty(rvcr, oracle, tmp_path):
    import torch
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((300, 128)) * 0.33).astype(np.float32)
    index = rvcr.index_factory(128, "IVF8,Flat")
    with pytest.raises(rvcr.RvcrError) as e:
        index.search(x[:2], k=8)                           # untrained
    assert e.value.code == -6
    index.set_centroids(x[:8])
    index.add(x)
    D, I = index.search(np.zeros((0, 128), np.float32), k=8)
    assert D.shape == (0, 8) and I.shape == (0, 8)
    with pytest.raises(rvcr.RvcrError) as e:
        index.search(x[:2], k=4)
    assert e.value.code == -5
    index.nprobe = 1000                                     # clamped to nlist like IndexIVF::search: every list is scanned
    D, I = index.search(x[:2], k=8)
    assert I[0, 0] == 0 and D[0, 0] == 0 and I[1, 0] == 1
    with pytest.raises(rvcr.RvcrError):
        index.set_option("nosuchoption", 1)
    index.nprobe = 1
    with pytest.raises(AssertionError):
        index.search(x[:2, :64], k=8)
    bad = tmp_path / "bad.index"
    bad.write_bytes(b"IxF2" + b"\0" * 64)
    with pytest.raises(rvcr.RvcrError) as e:
        rvcr.read_index(str(bad))
    assert e.value.code == -3
    with pytest.raises(rvcr.RvcrError) as e:
        rvcr.read_index(str(tmp_path / "missing.index"))
    assert e.value.code == -2
    # pipeline.py:554-559: any exception -> index = None -> features pass through unchanged
    st = index.stats()
    assert st["kernels_launched"] > 0 and st["queries"] >= 0
    # a batch-of-one and a non-multiple-of-16 batch agree with the batched result
    D1, I1 = index.search(x[5:6], k=8)
    D37, I37 = index.search(x[:37], k=8)
    assert np.array_equal(I1[0], I37[5]) and np.array_equal(D1[0], D37[5])
    t = torch.from_numpy(x[:37].copy()).cuda()
    Dt, It = index.search(t, k=8)
    assert np.array_equal(It.cpu().numpy(), I37) and np.array_equal(Dt.cpu().numpy(), D37)


def test_sparse_file_and_shard_reader(rvcr, oracle, tmp_path):
    """'sprs' list-size encoding; a 2-shard read covers the index exactly once and every shard answers
    for the lists it owns with global ids."""
    import torch
    rng = np.random.default_rng(2)
    d, nlist = 128, 64
    cent = (rng.standard_normal((nlist, d)) * 2).astype(np.float32)
    x = (cent[rng.integers(0, 20, 900)] + rng.standard_normal((900, d)).astype(np.float32) * 0.1).astype(np.float32)
    oidx, path = write_oracle_index(oracle, tmp_path, x, cent, "sp.index")
    assert (np.diff(oidx.list_off) > 0).sum() <= nlist // 2      # writer chose 'sprs'
    assert b"sprs" in open(path, "rb").read()[:64 * 1024 * 1024]
    full = rvcr.read_index(path)
    assert np.array_equal(full.reconstruct_n(0, 900), x)
    q = (x[:100] + np.float32(0.01)).astype(np.float32)
    Df, If = full.search(q)
    Do, Io = oidx.search(q)
    assert np.array_equal(If, Io) and np.array_equal(Df, Do)
    shards = [rvcr.read_index(path, shard=(r, 2)) for r in range(2)]
    infos = [s.info for s in shards]
    assert infos[0]["owned_begin"] == 0 and infos[0]["owned_end"] == infos[1]["owned_begin"]
    assert infos[1]["owned_end"] == nlist and infos[0]["nlist"] == infos[1]["nlist"] == nlist
    assert infos[0]["nstored"] + infos[1]["nstored"] == 900
    assert abs(infos[0]["nstored"] - 450) <= np.diff(oidx.list_off).max()
    from importlib import import_module
    synth = import_module("codename-rvc-fork-3_b200.synth")
    assert [(i["owned_begin"], i["owned_end"]) for i in infos] == synth.shard_ranges(np.diff(oidx.list_off), 2)
    # the quantizer is replicated: every shard gives the global assignment; each answers for its lists
    qd = torch.from_numpy(q).cuda()
    ao, _ = oracle.coarse_assign(q, cent)
    D2 = np.full_like(Do, np.float32(3.4028234663852886e38)); I2 = np.full_like(Io, -1)
    for s in shards:
        la, _ = s.assign(qd)
        assert np.array_equal(la.cpu().numpy(), ao)
        Ds, Is = s.search(qd)                         # pads for queries whose list another shard stores
        Ds, Is = Ds.cpu().numpy(), Is.cpu().numpy()
        mine = (ao >= s.info["owned_begin"]) & (ao < s.info["owned_end"])
        assert (Is[~mine] == -1).all()
        D2[mine], I2[mine] = Ds[mine], Is[mine]
    assert np.array_equal(I2, Io) and np.array_equal(D2, Do)
    with pytest.raises(rvcr.RvcrError):
        rvcr.write_index(shards[0], str(tmp_path / "no.index"))       # a shard cannot be serialised
