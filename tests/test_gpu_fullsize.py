"""GPU: BASELINE.json's full single-GPU batch configuration (C3: 2 M x 768 index, nlist 22 627, 180 k query
frames -- one hour of audio) checked through size-independent properties, since the CPU oracle cannot finish
that size in seconds:

  * the tensor-core kernels (tcgen05 coarse + tcgen05 scan, the default path) give bit-identical list ids,
    neighbour ids and distances to the exact FP32 FFMA kernels (the brute-force canonical argmin / scan);
  * structure: distances ascending, ids inside the assigned list, ties ordered by id;
  * batch independence: a permuted batch and a split batch give the same per-query answers;
  * the fused tail equals the oracle's tail on a sample of rows fed with the GPU's own (score, ix);
  * on the same sample the CPU oracle itself (coarse + scan over the staged lists) gives the same ids.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c3(rvcr):
    import torch
    from importlib import import_module
    synth = import_module("codename-rvc-fork-3_b200.synth")
    N, d, nq, nlist = synth.CONFIGS["C3"]
    dev = torch.device("cuda:0")
    sizes = synth.list_sizes(N, nlist)
    cent = synth.all_centroids(nlist, d, dev)
    index = rvcr.index_factory(d, f"IVF{nlist},Flat", device=0)
    index.nprobe = 1
    synth.fill_index_by_construction(index, cent, sizes, N)
    q = synth.device_queries(cent, nq)
    # make it harder than the generator does: 2 % of the queries are exact copies of stored rows (score 0 -> NaN
    # row in the reference)
    vecs, ids, offs, _ = index.device_views()
    g = torch.Generator(device=dev).manual_seed(5)
    rows = torch.randint(0, N, (nq // 50,), generator=g, device=dev)
    q[: nq // 50] = vecs[rows]
    return dict(index=index, q=q, N=N, d=d, nq=nq, nlist=nlist, cent=cent, sizes=sizes)


def _search(index, q, coarse_mode, scan_impl):
    index.set_option("coarse_mode", coarse_mode)
    index.set_option("scan_impl", scan_impl)
    lst, dist = index.assign(q)
    D, I = index.search(q, k=8)
    return lst, dist, D, I


def test_tensor_core_path_equals_exact_kernels_at_c3(c3):
    import torch
    index, q = c3["index"], c3["q"]
    lst_tc, dist_tc, D_tc, I_tc = _search(index, q, 2, 5)          # tcgen05 coarse + tcgen05 scan (defaults)
    st = index.stats()
    assert st["coarse_tc_queries"] > 0 and st["scan_tc_candidates"] > 0   # the tensor-core kernels really ran
    lst_ex, dist_ex, D_ex, I_ex = _search(index, q, 1, 4)          # exact FFMA coarse + exact FFMA scan
    assert torch.equal(lst_tc, lst_ex) and torch.equal(dist_tc, dist_ex)
    assert torch.equal(I_tc, I_ex) and torch.equal(D_tc, D_ex)
    index.set_option("coarse_mode", 0)
    index.set_option("scan_impl", 5)
    # structure
    assert bool((D_tc[:, 1:] >= D_tc[:, :-1]).all())
    assert int(I_tc.min()) >= 0 and int(I_tc.max()) < c3["N"]      # lists hold >= 8 rows at C3: no padding
    offs = torch.as_tensor(np.concatenate([[0], np.cumsum(c3["sizes"])]), device=q.device)
    lo, hi = offs[lst_tc.long()], offs[lst_tc.long() + 1]           # ids == global row positions by construction
    assert bool(((I_tc >= lo[:, None]) & (I_tc < hi[:, None])).all())
    tie = D_tc[:, 1:] == D_tc[:, :-1]
    assert bool((I_tc[:, 1:][tie] > I_tc[:, :-1][tie]).all())       # equal distances: ascending id
    zero = (D_tc[:, 0] == 0)
    # the planted exact copies are found at distance 0 (all but the rare row whose nearest centroid is a sibling's)
    assert int(zero.sum()) >= int(0.99 * (c3["nq"] // 50))


def test_batch_independence_at_c3(c3):
    import torch
    index, q = c3["index"], c3["q"]
    D, I = index.search(q, k=8)
    g = torch.Generator(device=q.device).manual_seed(9)
    perm = torch.randperm(q.shape[0], generator=g, device=q.device)
    Dp, Ip = index.search(q[perm].contiguous(), k=8)
    assert torch.equal(Dp, D[perm]) and torch.equal(Ip, I[perm])
    cut = 70_001                                                    # ragged split
    D1, I1 = index.search(q[:cut].contiguous(), k=8)
    D2, I2 = index.search(q[cut:].contiguous(), k=8)
    assert torch.equal(torch.cat([D1, D2]), D) and torch.equal(torch.cat([I1, I2]), I)


def test_fused_tail_and_oracle_on_a_sample_at_c3(c3, oracle):
    import torch
    index, q = c3["index"], c3["q"]
    D, I = index.search(q, k=8)
    lst, _ = index.assign(q)
    feats = q.clone()
    index.retrieve_blend_(feats, 0.75)
    g = torch.Generator(device=q.device).manual_seed(3)
    pick = torch.cat([torch.arange(0, 64, device=q.device),                     # exact-copy queries (NaN rows)
                      torch.randint(0, q.shape[0], (1500,), generator=g, device=q.device)])
    vecs, ids, offs, cent = index.device_views()
    # stage the probed lists of the sample on the host, as a small index with the same global ids
    lists = torch.unique(lst[pick].long())
    off_h = offs.cpu().numpy()
    rows = np.concatenate([np.arange(off_h[l], off_h[l + 1]) for l in lists.cpu().numpy()])
    sub_vecs = vecs[torch.as_tensor(rows, device=q.device)].cpu().numpy()
    sub_ids = ids[torch.as_tensor(rows, device=q.device)].cpu().numpy()
    sub_off = np.concatenate([[0], np.cumsum([off_h[l + 1] - off_h[l] for l in lists.cpu().numpy()])]).astype(np.int64)
    remap = {int(l): i for i, l in enumerate(lists.cpu().numpy())}
    qs = q[pick].cpu().numpy()
    a_local = np.array([remap[int(l)] for l in lst[pick].cpu().numpy()], np.int32)
    Do, Io = oracle.scan_topk(qs, a_local, sub_off, sub_vecs, sub_ids, k=8)
    assert np.array_equal(I[pick].cpu().numpy(), Io) and np.array_equal(D[pick].cpu().numpy(), Do)
    # coarse: the oracle's canonical argmin over ALL centroids for a smaller sample
    small = pick[:200]
    a_o, d_o = oracle.coarse_assign(q[small].cpu().numpy(), cent.cpu().numpy())
    assert np.array_equal(lst[small].cpu().numpy(), a_o)
    # tail: big_npy[ix] gathered from the device storage (ids == row positions by construction)
    ix = I[pick]
    gathered = vecs[ix.reshape(-1)].reshape(len(pick), 8, -1).cpu().numpy()
    big = gathered.reshape(-1, gathered.shape[-1])
    ix_local = np.arange(len(pick) * 8, dtype=np.int64).reshape(len(pick), 8)
    with np.errstate(all="ignore"):
        want = oracle.tail(D[pick].cpu().numpy(), ix_local, big, 0.75, qs)
    got = feats[pick].cpu().numpy()
    assert np.array_equal(got, want, equal_nan=True)
    z = D[pick, 0].cpu().numpy() == 0                                # exact copies: inf / inf, like the reference
    assert z[:64].sum() >= 60
    assert np.array_equal(np.isnan(got).all(axis=1), z) and np.array_equal(np.isnan(got).any(axis=1), z)
