"""tcgen05 tf32 list scan (scan_impl = 5, csrc/scan_tc.cu) against the oracle and against the exact FFMA
scan (scan_impl = 4).  The tensor-core pass only nominates candidates; the re-check evaluates the canonical
fp32 distance, so ids AND distances must be bit-identical -- including ties, duplicates, short lists and
the overflow path that falls back to the brute-force exact kernel."""
from importlib import import_module

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TC_IMPL = 5


def _build(rvcr, x, cent):
    index = rvcr.index_factory(x.shape[1], f"IVF{len(cent)},Flat")
    index.nprobe = 1
    index.set_centroids(cent)
    for i in range(0, len(x), 65536):
        index.add(x[i:i + 65536])
    assert index.get_option("scan_impl") == 0            # library default: auto (exact FFMA scan below 1 Ki queries)
    index.set_option("scan_impl", TC_IMPL)               # this file tests the tensor-core scan at every size
    return index


def test_scan_tc_matches_oracle_and_exact_scan(rvcr, oracle):
    """long lists (~390 rows, 4 row tiles, ragged tails), up to > 16 queries per list (several work items)"""
    synth = import_module("codename-rvc-fork-3_b200.synth")
    n, d, nq, nlist = 200_000, 768, 12_000, 512
    pr = synth.make_host_problem(n, d, nq, nlist)
    index = _build(rvcr, pr.x, pr.centroids)
    a_x, _ = oracle.coarse_assign(pr.x, pr.centroids)
    off, vecs, ids = oracle.build_lists(pr.x, a_x, nlist)
    a_q, _ = oracle.coarse_assign(pr.q, pr.centroids)
    Do, Io = oracle.scan_topk(pr.q, a_q, off, vecs, ids)
    s0 = index.stats()
    assert index.get_option("scan_impl") == TC_IMPL
    D5, I5 = index.search(pr.q, k=8)
    s1 = index.stats()
    assert np.array_equal(I5, Io) and np.array_equal(D5, Do)
    index.set_option("scan_impl", 4)
    D4, I4 = index.search(pr.q, k=8)
    assert np.array_equal(I4, I5) and np.array_equal(D4, D5)
    # the tensor-core pass did the work: candidates were re-checked, and only a few queries overflowed
    cand = s1["scan_tc_candidates"] - s0["scan_tc_candidates"]
    fb = s1["scan_tc_fallback"] - s0["scan_tc_fallback"]
    assert cand >= 8 * (nq - fb) * 0.9
    assert fb <= nq // 20, (fb, cand)
    print(f"scan_tc: {cand / max(1, nq - fb):.1f} candidates per query, {fb} of {nq} queries fell back")


def test_scan_tc_v1_256_and_small_lists(rvcr, oracle):
    """d = 256 (8 k-blocks), i.i.d. data: hundreds of lists shorter than 8 (padding), many empty probes"""
    synth = import_module("codename-rvc-fork-3_b200.synth")
    n, d, nq, nlist = 30_000, 256, 2_000, 1500
    pr = synth.make_host_problem(n, d, nq, nlist, iid=True)
    index = _build(rvcr, pr.x, pr.centroids)
    oidx = oracle.OracleIndex.from_vectors(pr.x, pr.centroids)
    Do, Io = oidx.search(pr.q, k=8)
    D, I = index.search(pr.q, k=8)
    assert (Io < 0).any()
    assert np.array_equal(I, Io) and np.array_equal(D, Do)


def test_scan_tc_overflow_falls_back_to_exact(rvcr, oracle):
    """a list of 300 identical rows: every row is inside the margin, the quarter lists overflow, and the
    brute-force exact kernel must answer with the first 8 positions (faiss keeps the earlier entry on ties);
    a second list with 40 near-duplicates (> 32 candidates in one quarter is impossible here, but > 8 ties)"""
    rng = np.random.default_rng(3)
    d = 768
    cent = (rng.standard_normal((3, d)) * 2).astype(np.float32)
    a = np.repeat(cent[0:1] + np.float32(0.05), 300, axis=0)
    b = cent[1] + (rng.standard_normal((40, d)) * 1e-4).astype(np.float32)
    c = cent[2] + (rng.standard_normal((500, d)) * 0.14).astype(np.float32)
    x = np.concatenate([a, b, c]).astype(np.float32)
    index = _build(rvcr, x, cent)
    oidx = oracle.OracleIndex.from_vectors(x, cent)
    q = np.concatenate([cent[0:1] + np.float32(0.01), cent[1:2] + np.float32(0.01), cent[2:3],
                        c[:20] + np.float32(0.003)]).astype(np.float32)
    s0 = index.stats()
    D, I = index.search(q, k=8)
    s1 = index.stats()
    Do, Io = oidx.search(q, k=8)
    assert np.array_equal(I, Io) and np.array_equal(D, Do)
    assert list(I[0]) == list(range(8))
    assert s1["scan_tc_fallback"] - s0["scan_tc_fallback"] >= 1


def test_scan_tc_blend_and_host_waves(rvcr, oracle):
    """the fused retrieve+blend and the pinned-host wave pipeline run on top of the tensor-core scan"""
    import torch
    synth = import_module("codename-rvc-fork-3_b200.synth")
    n, d, nq, nlist = 120_000, 768, 6_000, 1200
    pr = synth.make_host_problem(n, d, nq, nlist)
    index = _build(rvcr, pr.x, pr.centroids)
    oidx = oracle.OracleIndex.from_vectors(pr.x, pr.centroids)
    Do, Io = oidx.search(pr.q, k=8)
    with np.errstate(all="ignore"):
        want = oracle.tail(Do, Io, pr.x, 0.75, pr.q)
    f = torch.from_numpy(pr.q.copy()).cuda()
    index.retrieve_blend_(f, 0.75)
    assert np.array_equal(f.cpu().numpy(), want, equal_nan=True)
    hp = torch.from_numpy(pr.q.copy()).pin_memory()
    index.retrieve_blend_host_(hp.numpy(), 0.75)
    assert np.array_equal(hp.numpy(), want, equal_nan=True)


def test_fused_tail_equals_separate_blend_kernel(rvcr, oracle):
    """option fused_blend: the re-check / fallback kernels also do pipeline.py:500-506; both settings must be
    bit-identical to the separate gather_blend kernel and to the oracle -- on ordinary lists, on the overflow
    fallback (identical rows), on 40 near-duplicates (3 candidate passes), on short lists (padding -> big_npy[-1], weight 0), on an empty probe (NaN
    row), an exact-match query (NaN row) and a NaN query (assignment -1)."""
    import torch
    rng = np.random.default_rng(11)
    d = 768
    cent = (rng.standard_normal((6, d)) * 2).astype(np.float32)
    a = np.repeat(cent[0:1] + np.float32(0.05), 300, axis=0)                       # overflow -> exact fallback
    b = cent[1] + (rng.standard_normal((5, d)) * 0.14).astype(np.float32)         # short list: 3 pads
    c = cent[2] + (rng.standard_normal((700, d)) * 0.14).astype(np.float32)       # two chunks
    e = cent[4] + (rng.standard_normal((64, d)) * 0.14).astype(np.float32)        # list 3 stays empty
    g = cent[5] + (rng.standard_normal((40, d)) * 1e-4).astype(np.float32)        # 40 candidates inside the margin
    x = np.concatenate([a, b, c, e, g]).astype(np.float32)
    index = _build(rvcr, x, cent)
    oidx = oracle.OracleIndex.from_vectors(x, cent)
    q = np.concatenate([cent[0:1] + np.float32(0.01), cent[1:2] + np.float32(0.01), c[:40] + np.float32(0.003),
                        cent[3:4], c[100:101], e[:30] + np.float32(0.002), cent[5:6] + np.float32(0.01),
                        cent[2:3]]).astype(np.float32)
    q[-1, 7] = np.nan
    Do, Io = oidx.search(q, k=8)
    with np.errstate(all="ignore"):
        want = oracle.tail(Do, Io, x, 0.3, q)
    assert np.isnan(want[42]).all() and np.isnan(want[43]).all()           # empty probe, exact match
    for fused in (1, 0):
        index.set_option("fused_blend", fused)
        f = torch.from_numpy(q.copy()).cuda()
        _, D, I = index.retrieve_blend_(f, 0.3, want_neighbours=True)
        assert np.array_equal(I.cpu().numpy(), Io) and np.array_equal(D.cpu().numpy(), Do), fused
        assert np.array_equal(f.cpu().numpy(), want, equal_nan=True), fused
        D2, I2 = index.search(q, k=8)
        assert np.array_equal(I2, Io) and np.array_equal(D2, Do), fused


def test_fused_tail_v1_256(rvcr, oracle):
    """d = 256, i.i.d. data with many short / empty lists, index_rate 1.0 through the fused tail"""
    import torch
    synth = import_module("codename-rvc-fork-3_b200.synth")
    n, d, nq, nlist = 30_000, 256, 3_000, 1500
    pr = synth.make_host_problem(n, d, nq, nlist, iid=True)
    index = _build(rvcr, pr.x, pr.centroids)
    oidx = oracle.OracleIndex.from_vectors(pr.x, pr.centroids)
    Do, Io = oidx.search(pr.q, k=8)
    with np.errstate(all="ignore"):
        want = oracle.tail(Do, Io, pr.x, 1.0, pr.q)
    f = torch.from_numpy(pr.q.copy()).cuda()
    index.retrieve_blend_(f, 1.0)
    assert np.array_equal(f.cpu().numpy(), want, equal_nan=True)
