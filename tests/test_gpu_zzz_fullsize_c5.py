"""GPU: BASELINE.json's headline configuration at full size on ONE B200 (C5: 32 M x 768 index = 98.3 GB of list
vectors, nlist 90 509, 1 M query frames -- the workload bench.py times), checked through size-independent properties
exactly like tests/test_gpu_fullsize.py does for C3:

  * the default tensor-core path (tcgen05 coarse + tcgen05 scan + exact re-checks) gives bit-identical list ids,
    neighbour ids and distances to the exact FP32 FFMA kernels over all 1 M queries x 90 509 centroids / 32 M rows;
  * structure: distances ascending, ids inside the assigned list, ties ordered by id, planted exact copies found at 0;
  * on a sample, the CPU oracle (canonical coarse argmin over ALL centroids; scan over the staged lists) gives the same
    ids and distances, and the fused tail equals the oracle's tail (NaN rows for exact copies included).

(Named zzz: added after the last GPU session of round 1, it runs after the established suites.  Needs ~110 GB of HBM.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c5(rvcr):
    import gc

    import torch
    from importlib import import_module
    synth = import_module("codename-rvc-fork-3_b200.synth")
    gc.collect()
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info(0)
    if free < 125 * (1 << 30):
        pytest.skip(f"C5 needs ~110 GB of free HBM, {free >> 30} GB available")
    N, d, nq, nlist = synth.CONFIGS["C5"]
    dev = torch.device("cuda:0")
    sizes = synth.list_sizes(N, nlist)
    cent = synth.all_centroids(nlist, d, dev)
    index = rvcr.index_factory(d, f"IVF{nlist},Flat", device=0)
    index.nprobe = 1
    synth.fill_index_by_construction(index, cent, sizes, N)
    q = synth.device_queries(cent, nq)
    vecs, ids, offs, _ = index.device_views()
    g = torch.Generator(device=dev).manual_seed(5)
    rows = torch.randint(0, N, (nq // 50,), generator=g, device=dev)
    q[: nq // 50] = vecs[rows]                       # 2 % exact copies of stored rows: score 0 -> NaN row in the reference
    yield dict(index=index, q=q, N=N, d=d, nq=nq, nlist=nlist, cent=cent, sizes=sizes)
    index.close()
    torch.cuda.empty_cache()


def _search(index, q, coarse_mode, scan_impl):
    index.set_option("coarse_mode", coarse_mode)
    index.set_option("scan_impl", scan_impl)
    lst, dist = index.assign(q)
    D, I = index.search(q, k=8)
    return lst, dist, D, I


def test_tensor_core_path_equals_exact_kernels_at_c5(c5):
    import torch
    index, q = c5["index"], c5["q"]
    lst_tc, dist_tc, D_tc, I_tc = _search(index, q, 2, 5)          # tcgen05 coarse + tcgen05 scan (what bench.py times)
    st = index.stats()
    assert st["coarse_tc_queries"] > 0 and st["scan_tc_candidates"] > 0
    lst_ex, dist_ex, D_ex, I_ex = _search(index, q, 1, 4)          # exact FFMA coarse (139 TFLOP: ~7 s) + exact FFMA scan
    index.set_option("coarse_mode", 0)
    index.set_option("scan_impl", 0)
    assert torch.equal(lst_tc, lst_ex) and torch.equal(dist_tc, dist_ex)
    assert torch.equal(I_tc, I_ex) and torch.equal(D_tc, D_ex)
    assert bool((D_tc[:, 1:] >= D_tc[:, :-1]).all())
    assert int(I_tc.min()) >= 0 and int(I_tc.max()) < c5["N"]      # every list holds >= 8 rows at C5: no padding
    offs = torch.as_tensor(np.concatenate([[0], np.cumsum(c5["sizes"])]), device=q.device)
    lo, hi = offs[lst_tc.long()], offs[lst_tc.long() + 1]           # ids == global row positions by construction
    assert bool(((I_tc >= lo[:, None]) & (I_tc < hi[:, None])).all())
    tie = D_tc[:, 1:] == D_tc[:, :-1]
    assert bool((I_tc[:, 1:][tie] > I_tc[:, :-1][tie]).all())
    assert int((D_tc[:, 0] == 0).sum()) >= int(0.99 * (c5["nq"] // 50))


def test_oracle_sample_and_fused_tail_at_c5(c5, oracle):
    import torch
    index, q = c5["index"], c5["q"]
    D, I = index.search(q, k=8)
    lst, _ = index.assign(q)
    feats = q.clone()
    index.retrieve_blend_(feats, 0.75)
    g = torch.Generator(device=q.device).manual_seed(3)
    pick = torch.cat([torch.arange(0, 64, device=q.device),                     # exact-copy queries (NaN rows)
                      torch.randint(0, q.shape[0], (700,), generator=g, device=q.device)])
    vecs, ids, offs, cent = index.device_views()
    lists = torch.unique(lst[pick].long()).cpu().numpy()
    off_h = offs.cpu().numpy()
    rows = np.concatenate([np.arange(off_h[l], off_h[l + 1]) for l in lists])
    rows_d = torch.as_tensor(rows, device=q.device)
    sub_vecs = vecs[rows_d].cpu().numpy()
    sub_ids = ids[rows_d].cpu().numpy()
    sub_off = np.concatenate([[0], np.cumsum([off_h[l + 1] - off_h[l] for l in lists])]).astype(np.int64)
    remap = {int(l): i for i, l in enumerate(lists)}
    qs = q[pick].cpu().numpy()
    a_local = np.array([remap[int(l)] for l in lst[pick].cpu().numpy()], np.int32)
    Do, Io = oracle.scan_topk(qs, a_local, sub_off, sub_vecs, sub_ids, k=8)
    assert np.array_equal(I[pick].cpu().numpy(), Io) and np.array_equal(D[pick].cpu().numpy(), Do)
    small = pick[64:164]                              # canonical argmin over all 90 509 centroids on the CPU
    a_o, _ = oracle.coarse_assign(q[small].cpu().numpy(), cent.cpu().numpy())
    assert np.array_equal(lst[small].cpu().numpy(), a_o)
    ix = I[pick]
    gathered = vecs[ix.reshape(-1)].reshape(len(pick), 8, -1).cpu().numpy()
    big = gathered.reshape(-1, gathered.shape[-1])
    ix_local = np.arange(len(pick) * 8, dtype=np.int64).reshape(len(pick), 8)
    with np.errstate(all="ignore"):
        want = oracle.tail(D[pick].cpu().numpy(), ix_local, big, 0.75, qs)
    got = feats[pick].cpu().numpy()
    assert np.array_equal(got, want, equal_nan=True)
    z = D[pick, 0].cpu().numpy() == 0
    assert z[:64].sum() >= 60
    assert np.array_equal(np.isnan(got).all(axis=1), z) and np.array_equal(np.isnan(got).any(axis=1), z)
