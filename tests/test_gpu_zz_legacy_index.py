"""GPU: the legacy IndexIVFFlat file layouts faiss 1.7.3 still reads ("IvFL" / "IvFl") load into the same
device-resident index as the current layout -- same reconstruct_n, same neighbours, same blended features -- whole and
as shards.  (Named zz so that it runs after the established suites: the legacy reader was added after the last GPU
session of round 1 and is verified at the parser level on CPU, tests/test_index_probe.py.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def build(oracle, d=256, nlist=16, n=900, seed=21):
    rng = np.random.default_rng(seed)
    cent = (rng.standard_normal((nlist, d)) * 0.5).astype(np.float32)
    x = (cent[rng.integers(0, nlist, n)] + rng.standard_normal((n, d)).astype(np.float32) * 0.1).astype(np.float32)
    oidx = oracle.OracleIndex.from_vectors(x, cent)
    q = (x[:96] + np.float32(0.02)).astype(np.float32)
    return oidx, x, q


@pytest.mark.parametrize("fourcc", [b"IvFL", b"IvFl"])
def test_legacy_file_gives_the_same_index(rvcr, oracle, tmp_path, fourcc):
    import torch
    oidx, x, q = build(oracle)
    new, old = str(tmp_path / "new.index"), str(tmp_path / "old.index")
    oracle.write_index(oidx.to_data(), new)
    oracle.write_index_legacy(oidx.to_data(), old, fourcc)
    a, b = rvcr.read_index(new), rvcr.read_index(old)
    assert b.ntotal == a.ntotal == len(x) and b.d == a.d and b.nlist == a.nlist and b.nprobe == 1
    assert np.array_equal(b.list_sizes(), a.list_sizes())
    assert np.array_equal(b.reconstruct_n(0, b.ntotal), x)
    Do, Io = oidx.search(q, k=8)
    D, I = b.search(q, k=8)
    assert np.array_equal(I, Io) and np.array_equal(D, Do)
    fa, fb = torch.from_numpy(q.copy()).cuda(), torch.from_numpy(q.copy()).cuda()
    a.retrieve_blend_(fa, 0.75)
    b.retrieve_blend_(fb, 0.75)
    assert torch.equal(fa, fb)
    # written back, it is the current layout, byte for byte what the current-layout file was
    out = str(tmp_path / "rewritten.index")
    rvcr.write_index(b, out)
    assert open(out, "rb").read() == open(new, "rb").read()


def test_legacy_file_loads_as_shards(rvcr, oracle, tmp_path):
    import torch
    oidx, x, q = build(oracle, seed=22)
    old = str(tmp_path / "old.index")
    oracle.write_index_legacy(oidx.to_data(), old, b"IvFL")
    sizes = rvcr.probe_index(old, list_sizes=True)["list_sizes"]
    Do, Io = oidx.search(q, k=8)
    got_D = np.full_like(Do, np.inf)
    got_I = np.full_like(Io, -1)
    seen = 0
    for r in range(3):
        sh = rvcr.read_index(old, shard=(r, 3))
        lo, hi = rvcr.shard_range(sizes, r, 3)
        inf = sh.info
        assert (inf["owned_begin"], inf["owned_end"]) == (lo, hi)
        assert inf["nstored"] == int(sizes[lo:hi].sum())
        seen += inf["nstored"]
        assign, _ = sh.assign(torch.from_numpy(q).cuda())
        a = assign.cpu().numpy()
        mine = (a >= lo) & (a < hi)
        D, I = sh.search(q, k=8)                       # rows whose list lives on another shard come back as pads
        got_D[mine], got_I[mine] = D[mine], I[mine]
    assert seen == len(x)
    assert np.array_equal(got_I, Io) and np.array_equal(got_D, Do)
