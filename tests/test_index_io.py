"""CPU: `.index` byte layout -- oracle writer/reader against a hand-assembled fixture."""
import io
import struct

import numpy as np
import pytest


def hand_assembled(d, centroids, lists, ntotal, nprobe=1, sparse=False):
    """Third, independent restatement of Appendix B: struct.pack field by field."""
    nlist = len(lists)
    b = b"IwFl" + struct.pack("<i", d) + struct.pack("<q", ntotal) + struct.pack("<qq", 1 << 20, 1 << 20)
    b += struct.pack("<B", 1) + struct.pack("<i", 1) + struct.pack("<Q", nlist) + struct.pack("<Q", nprobe)
    b += b"IxF2" + struct.pack("<i", d) + struct.pack("<q", nlist) + struct.pack("<qq", 1 << 20, 1 << 20)
    b += struct.pack("<B", 1) + struct.pack("<i", 1) + struct.pack("<Q", nlist * d)
    b += np.asarray(centroids, "<f4").tobytes()
    b += struct.pack("<B", 0) + struct.pack("<Q", 0)
    b += b"ilar" + struct.pack("<QQ", nlist, 4 * d)
    if sparse:
        nz = [(i, len(ids)) for i, (v, ids) in enumerate(lists) if len(ids)]
        b += b"sprs" + struct.pack("<Q", 2 * len(nz))
        for i, n in nz:
            b += struct.pack("<QQ", i, n)
    else:
        b += b"full" + struct.pack("<Q", nlist)
        for v, ids in lists:
            b += struct.pack("<Q", len(ids))
    for v, ids in lists:
        if len(ids):
            b += np.asarray(v, "<f4").tobytes() + np.asarray(ids, "<i8").tobytes()
    return b


@pytest.mark.parametrize("d,sparse", [(768, False), (256, False), (16, True)])
def test_writer_matches_hand_assembled_bytes(oracle, d, sparse):
    rng = np.random.default_rng(d)
    nlist = 10
    cent = rng.standard_normal((nlist, d)).astype(np.float32)
    sizes = [0, 0, 0, 0, 0, 0, 3, 0, 5, 0] if sparse else [3, 0, 5, 1, 2, 0, 4, 9, 1, 1]
    lists, nid = [], 0
    for n in sizes:
        v = rng.standard_normal((n, d)).astype(np.float32)
        ids = np.arange(nid, nid + n, dtype=np.int64)[::1]
        nid += n
        lists.append((v, ids))
    want = hand_assembled(d, cent, lists, nid, sparse=sparse)
    data = oracle.IVFFlatData(d=d, ntotal=nid, nprobe=1, centroids=cent, list_sizes=np.array(sizes),
                              vecs=np.concatenate([v for v, _ in lists]), ids=np.concatenate([i for _, i in lists]))
    buf = io.BytesIO()
    oracle.write_index(data, buf)
    assert buf.getvalue() == want
    # header offsets of Appendix B
    assert want[:4] == b"IwFl" and struct.unpack_from("<i", want, 4)[0] == d
    assert want[32] == 1 and struct.unpack_from("<i", want, 33)[0] == 1
    assert struct.unpack_from("<QQ", want, 37) == (nlist, 1) and want[53:57] == b"IxF2"
    assert struct.unpack_from("<Q", want, 90)[0] == nlist * d
    back = oracle.read_index(want)
    assert back.d == d and back.ntotal == nid and back.nprobe == 1
    assert np.array_equal(back.centroids, cent) and back.list_sizes.tolist() == sizes
    assert np.array_equal(back.vecs, data.vecs) and np.array_equal(back.ids, data.ids)


def test_reader_rejects_garbage(oracle):
    with pytest.raises(ValueError):
        oracle.read_index(b"IxF2" + b"\0" * 100)
    good = hand_assembled(8, np.zeros((2, 8), np.float32), [(np.zeros((1, 8), np.float32), [0]), (np.zeros((0, 8)), [])], 1)
    with pytest.raises(ValueError):
        oracle.read_index(good[:-5])


def test_oracle_index_roundtrip_through_file(oracle, tmp_path):
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((400, 32)) * 0.33).astype(np.float32)
    idx = oracle.OracleIndex.from_vectors(x, x[:12].copy())
    p = tmp_path / "a.index"
    oracle.write_index(idx.to_data(), str(p))
    idx2 = oracle.OracleIndex.from_data(oracle.read_index(str(p)))
    assert np.array_equal(idx2.reconstruct_n(0, 400), x)
    q = x[:20] + 0.01
    assert all(np.array_equal(a, b) for a, b in zip(idx.search(q), idx2.search(q)))
