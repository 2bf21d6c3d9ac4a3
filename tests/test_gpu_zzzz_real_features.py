"""GPU: parity on REAL ContentVec features (tests/real_feats.py; the CPU side of the same problems is tests/test_real_features.py).

Real frames have a few dominant dimensions and, at jitter 0.25, 1024 centroids squeezed around 318 distinct frames: dozens of
centroids inside the e4m3 margin and several inside the bf16 one, i.e. the candidate lists of the tensor-core epilogue
overflow and queries take the exact fallback.  Whatever path a query takes, list ids, distances, neighbours and blended
rows must equal the oracle's bit for bit.  (Named zzzz: written after the round's last GPU session, it runs last.)"""
import numpy as np
import pytest

from real_feats import real_problem

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("jitter", [0.25, 1.0])
def test_real_feature_parity_with_oracle(rvcr, oracle, jitter):
    import torch
    n, nq, nlist, d = 40_000, 2048, 1024, 768
    x, cent, q = real_problem(n, nq, nlist, jitter=jitter)
    a_x, _ = oracle.coarse_assign(x, cent)
    off, vecs, ids = oracle.build_lists(x, a_x, nlist)
    index = rvcr.index_factory(d, f"IVF{nlist},Flat")
    index.nprobe = 1
    index.set_centroids(cent)
    for i in range(0, n, 8192):                          # extract_index.py:67-69
        index.add(x[i:i + 8192])
    assert np.array_equal(index.list_sizes(), np.diff(off))
    a_q, cd = oracle.coarse_assign(q, cent)
    qd = torch.from_numpy(q).cuda()
    for mode in (0, 1, 2, 3, 4):                         # auto, exact FFMA, tcgen05 bf16 / 3-way split / e4m3
        index.set_option("coarse_mode", mode)
        lst, dist = index.assign(qd)
        assert np.array_equal(lst.cpu().numpy(), a_q), mode
        assert np.array_equal(dist.cpu().numpy(), cd), mode
    index.set_option("coarse_mode", 0)
    Do, Io = oracle.scan_topk(q, a_q, off, vecs, ids)
    with np.errstate(all="ignore"):
        want = oracle.tail(Do, Io, x, 0.75, q)
    for impl in (0, 4, 5):                               # auto, exact FFMA scan, tcgen05 tf32 scan + exact re-check
        index.set_option("scan_impl", impl)
        D, I = index.search(q, k=8)
        assert np.array_equal(I, Io) and np.array_equal(D, Do), impl
        f = torch.from_numpy(q.copy()).cuda()
        index.retrieve_blend_(f, 0.75)
        assert np.array_equal(f.cpu().numpy(), want, equal_nan=True), impl
    index.set_option("scan_impl", 0)
