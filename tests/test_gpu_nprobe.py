"""GPU: index.search / the fused retrieve + blend on an index whose header says nprobe > 1 (files made by other
forks arrive through tabs/download/download.py:41-45 and are searched as they are, pipeline.py:499, 555; faiss honours the
stored nprobe).  Reference semantics: the nprobe nearest centroids, their lists scanned in coarse-rank
order into one top-8 with the strict replacement rule.  Everything bit-identical to the oracle restatement."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def build(rvcr, oracle, n, d, nq, nlist, seed=5, iid=False):
    synth = importlib.import_module("codename-rvc-fork-3_b200.synth")
    pr = synth.make_host_problem(n, d, nq, nlist, seed=seed, iid=iid)
    oidx = oracle.OracleIndex.from_vectors(pr.x, pr.centroids)
    index = rvcr.index_factory(d, f"IVF{nlist},Flat")
    index.set_centroids(pr.centroids)
    for i in range(0, n, 8192):
        index.add(pr.x[i:i + 8192])
    return pr, oidx, index


@pytest.mark.parametrize("d,nlist,nprobe", [(768, 64, 2), (768, 64, 5), (256, 200, 16), (256, 40, 40), (256, 40, 1000)])
def test_nprobe_search_and_blend_match_oracle(rvcr, oracle, d, nlist, nprobe):
    import torch
    pr, oidx, index = build(rvcr, oracle, 6000, d, 700, nlist, seed=nprobe)
    q = pr.q.copy()
    q[3] = pr.x[17]                      # exact hit: distance 0 -> NaN row in the blend (reference behaviour)
    q[5, 7] = np.nan                     # NaN query: no list, all pads
    index.nprobe = nprobe
    assert index.nprobe == nprobe
    Do, Io = oidx.search(q, k=8, nprobe=nprobe)
    D, I = index.search(q, k=8)
    assert np.array_equal(I, Io) and np.array_equal(D, Do)
    assert (I[5] == -1).all()
    # more probes find nearer neighbours than one probe on this data (otherwise the test would prove nothing)
    D1, _ = oidx.search(q, k=8, nprobe=1)
    ok = np.isfinite(D1).all(1) & np.isfinite(Do).all(1)
    assert (Do[ok] <= D1[ok]).all() and (Do[ok] < D1[ok]).any()
    # fused retrieve + blend in place, device and host entry points (pipeline.py:497-507)
    big = oidx.reconstruct_n(0, oidx.ntotal)
    with np.errstate(all="ignore"):
        want = oracle.tail(Do, Io, big, 0.75, q)
    f = torch.from_numpy(q.copy()).cuda()
    _, Dd, Id = index.retrieve_blend_(f, 0.75, want_neighbours=True)
    assert np.array_equal(Id.cpu().numpy(), Io) and np.array_equal(Dd.cpu().numpy(), Do)
    assert np.array_equal(f.cpu().numpy(), want, equal_nan=True)
    assert np.isnan(want[3]).all() and np.isnan(want[5]).all()
    h = q.copy()
    index.retrieve_blend_host_(h, 0.75)
    assert np.array_equal(h, want, equal_nan=True)
    index.nprobe = 1
    D, I = index.search(q, k=8)
    assert np.array_equal(I, oidx.search(q, k=8)[1])


def test_nprobe_short_lists_ties_and_file_roundtrip(rvcr, oracle, tmp_path):
    """i.i.d. rows give many lists shorter than 8 (pads filled from the next probes), duplicated rows in DIFFERENT lists give
    equal distances across probes (the earlier-scanned list must win), and the value survives write_index / read_index."""
    pr, oidx, index = build(rvcr, oracle, 1500, 256, 300, 400, seed=11, iid=True)
    x = pr.x.copy()
    # rebuild with duplicates forced into different lists: copy row a onto row b and keep the original assignment of b
    a = oracle.coarse_assign(x, pr.centroids)[0]
    rng = np.random.default_rng(0)
    src, dst = rng.choice(len(x), 60, replace=False).reshape(2, 30)
    x[dst] = x[src]
    off, vecs, ids = oracle.build_lists(x, a, 400)           # assignment of the ORIGINAL rows: duplicates sit in other lists
    oidx = oracle.OracleIndex(256, pr.centroids, off, vecs, ids, len(x), nprobe=7)
    path = str(tmp_path / "np7.index")
    oracle.write_index(oidx.to_data(), path)
    index = rvcr.read_index(path)
    assert index.nprobe == 7
    q = np.concatenate([pr.q, x[src[:20]]]).astype(np.float32)
    Do, Io = oidx.search(q, k=8)
    D, I = index.search(q, k=8)
    assert np.array_equal(I, Io) and np.array_equal(D, Do)
    assert (np.diff(np.sort(Do[-20:], axis=1), axis=1) == 0).any()          # equal distances really occur
    out = str(tmp_path / "np7_out.index")
    rvcr.write_index(index, out)
    assert rvcr.read_index(out).nprobe == 7 and open(out, "rb").read() == open(path, "rb").read()
