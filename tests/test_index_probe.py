"""CPU: the library's `.index` parser through the host-only `rvcr_index_probe` (no GPU involved) -- the same parse
`rvcr_index_read` performs before it streams the list payload to the device (csrc/index_io.cu).

What the reference relies on: `faiss.read_index(file_index)` (rvc/infer/pipeline.py:555) either returns an index or
raises, and the caller turns the exception into "no retrieval" (:557-559).  Index files are downloaded from the
internet by the reference's own UI (tabs/download/download.py), so the parser must survive corrupt and hostile
files: every count is checked against the file length before it is trusted.  Covered here: the current layout
("IwFl", full / sprs list tables, direct maps, IxFl quantizer), the legacy layouts faiss 1.7.3 still reads ("IvFL",
"IvFl") against the oracle's writer, refusals with the documented codes, truncation at every byte,
hostile counts, random mutations, and the shard-range rule against its numpy restatement."""
import os
import struct

import numpy as np
import pytest

HDR = 4 + 8 + 8 + 8 + 1 + 4        # d, ntotal, dummy, dummy, is_trained, metric_type


def make_data(oracle, d=128, nlist=12, n=300, seed=3, empty_every=0):
    rng = np.random.default_rng(seed)
    cent = (rng.standard_normal((nlist, d)) * 0.5).astype(np.float32)
    assign = rng.integers(0, nlist, n)
    if empty_every:
        assign = (assign // empty_every) * empty_every          # most lists stay empty -> "sprs" table
    x = (cent[assign] + rng.standard_normal((n, d)).astype(np.float32) * 0.1).astype(np.float32)
    oidx = oracle.OracleIndex.from_vectors(x, cent)
    return oidx.to_data(), x


def write(tmp_path, name, raw):
    p = tmp_path / name
    p.write_bytes(bytes(raw))
    return str(p)


def new_format_bytes(oracle, tmp_path, **kw):
    data, x = make_data(oracle, **kw)
    p = str(tmp_path / "base.index")
    oracle.write_index(data, p)
    return bytearray(open(p, "rb").read()), data


def offsets(data):
    off_nprobe = 4 + HDR + 8
    off_qfourcc = 4 + HDR + 16
    off_count = off_qfourcc + 4 + HDR
    off_dm = off_count + 8 + data.nlist * data.d * 4
    off_il = off_dm + 9
    return dict(nlist=4 + HDR, nprobe=off_nprobe, qfourcc=off_qfourcc, count=off_count, dm=off_dm, il=off_il,
                lt=off_il + 4 + 16, cnt=off_il + 4 + 16 + 4, table=off_il + 4 + 16 + 4 + 8)


def test_probe_current_layout_full_and_sparse(rvcr, oracle, tmp_path):
    for name, kw, enc in (("full", dict(), 0), ("sprs", dict(nlist=40, empty_every=8), 1)):
        raw, data = new_format_bytes(oracle, tmp_path, **kw)
        path = write(tmp_path, name + ".index", raw)
        inf = rvcr.probe_index(path, list_sizes=True)
        assert inf["fourcc"] == "IwFl" and inf["quantizer_fourcc"] == "IxF2" and inf["layout"] == 0
        assert (inf["d"], inf["ntotal"], inf["nlist"], inf["nprobe"], inf["metric"]) == (data.d, data.ntotal, data.nlist, 1, 1)
        assert inf["is_trained"] == 1 and inf["direct_map_type"] == 0 and inf["list_encoding"] == enc
        assert np.array_equal(inf["list_sizes"], data.list_sizes)
        assert inf["nvectors"] == int(data.list_sizes.sum()) == data.ntotal
        assert inf["max_list_size"] == int(data.list_sizes.max())
        assert inf["nonempty_lists"] == int((data.list_sizes > 0).sum())
        assert inf["file_bytes"] == len(raw)
        assert inf["payload_offset"] == len(raw) - data.ntotal * (4 * data.d + 8)
        assert inf["searchable"] == 1


@pytest.mark.parametrize("fourcc,layout", [(b"IvFL", 1), (b"IvFl", 2)])
def test_probe_legacy_layouts_match_oracle_writer(rvcr, oracle, tmp_path, fourcc, layout):
    data, x = make_data(oracle, d=256, nlist=9, n=200, seed=5)
    path = str(tmp_path / "legacy.index")
    oracle.write_index_legacy(data, path, fourcc)
    inf = rvcr.probe_index(path, list_sizes=True)
    assert inf["fourcc"] == fourcc.decode() and inf["layout"] == layout and inf["list_encoding"] == 3
    assert (inf["d"], inf["ntotal"], inf["nlist"], inf["nprobe"]) == (256, 200, 9, 1)
    assert np.array_equal(inf["list_sizes"], data.list_sizes) and inf["nvectors"] == 200
    assert inf["payload_offset"] == -1 and inf["searchable"] == 1
    # the oracle's reader returns the same index from either layout
    back = oracle.read_index(path)
    assert np.array_equal(back.vecs, data.vecs) and np.array_equal(back.ids, data.ids)
    assert np.array_equal(back.centroids, data.centroids) and np.array_equal(back.list_sizes, data.list_sizes)
    # a legacy list whose id and code vectors disagree is refused
    raw = bytearray(open(path, "rb").read())
    first_ids = 4 + HDR + 16 + 4 + HDR + 8 + 9 * 256 * 4
    n0 = struct.unpack_from("<Q", raw, first_ids)[0]
    assert n0 == data.list_sizes[0]
    bad = raw[:first_ids] + struct.pack("<Q", n0 - 1) + raw[first_ids + 8 + 8:]        # drop one id of list 0
    with pytest.raises(rvcr.RvcrError) as e:
        rvcr.probe_index(write(tmp_path, "legacy_bad.index", bad))
    assert e.value.code == -3


def test_probe_other_writers_and_refusals(rvcr, oracle, tmp_path):
    """The same edited files tests/test_gpu_interop.py loads on the GPU, checked at the parser."""
    raw, data = new_format_bytes(oracle, tmp_path, d=256, nlist=16, n=700, seed=8)
    o, n = offsets(data), data.ntotal
    f1 = bytearray(raw)
    f1[o["qfourcc"]:o["qfourcc"] + 4] = b"IxFl"
    assert rvcr.probe_index(write(tmp_path, "ixfl.index", f1))["quantizer_fourcc"] == "IxFl"
    dm = o["dm"]
    f2 = bytearray(raw[:dm]) + bytes([1]) + struct.pack("<Q", n) + np.arange(n, dtype=np.int64).tobytes() + raw[dm + 9:]
    i2 = rvcr.probe_index(write(tmp_path, "dm_array.index", f2), list_sizes=True)
    assert i2["direct_map_type"] == 1 and np.array_equal(i2["list_sizes"], data.list_sizes)
    pairs = np.stack([np.arange(n, dtype=np.int64)] * 2, axis=1).tobytes()
    f3 = bytearray(raw[:dm]) + bytes([2]) + struct.pack("<Q", 0) + struct.pack("<Q", n) + pairs + raw[dm + 9:]
    i3 = rvcr.probe_index(write(tmp_path, "dm_hash.index", f3), list_sizes=True)
    assert i3["direct_map_type"] == 2 and np.array_equal(i3["list_sizes"], data.list_sizes)
    f4 = bytearray(raw)
    struct.pack_into("<Q", f4, o["nprobe"], 4)
    i4 = rvcr.probe_index(write(tmp_path, "np4.index", f4))
    assert i4["nprobe"] == 4 and i4["searchable"] == 1             # honoured by the multi-probe search path
    # trailing bytes after the payload are ignored, like faiss
    assert rvcr.probe_index(write(tmp_path, "trail.index", raw + b"xyz" * 11))["nvectors"] == n
    # an empty-lists object: faiss write_InvertedLists(nullptr) emits the 4-byte fourcc "il00" and NOTHING else, so a genuine
    # file ends right there (round-1 parsers wanted 16 more bytes and refused it)
    f7 = bytearray(raw[:o["il"]]) + b"il00"
    i7 = rvcr.probe_index(write(tmp_path, "il00.index", f7))
    assert i7["list_encoding"] == 2 and i7["nvectors"] == 0 and i7["file_bytes"] == o["il"] + 4
    d7 = oracle.read_index(bytes(f7))
    assert d7.nlist == data.nlist and int(d7.list_sizes.sum()) == 0 and np.array_equal(d7.centroids, data.centroids)
    # refused: the legacy fourcc on a current-layout body, inner product, truncation, unknown fourccs
    f5 = bytearray(raw)
    f5[0:4] = b"IvFl"
    f6 = bytearray(raw)
    struct.pack_into("<i", f6, 4 + HDR - 4, 0)                     # METRIC_INNER_PRODUCT
    f8 = bytearray(raw)
    f8[0:4] = b"IxF2"
    f9 = bytearray(raw)
    f9[o["il"]:o["il"] + 4] = b"ilvx"
    f10 = bytearray(raw)
    f10[o["lt"]:o["lt"] + 4] = b"xxxx"
    for name, blob, code in (("ivfl", f5, -3), ("ip", f6, -5), ("trunc", raw[:len(raw) - 1000], -3), ("flat", f8, -3),
                             ("ilvx", f9, -3), ("lt", f10, -3), ("empty", b"", -3)):
        with pytest.raises(rvcr.RvcrError) as e:
            rvcr.probe_index(write(tmp_path, name + ".index", blob))
        assert e.value.code == code, name
    with pytest.raises(rvcr.RvcrError) as e:
        rvcr.probe_index(str(tmp_path / "missing.index"))
    assert e.value.code == -2


def test_probe_hostile_counts_fail_fast(rvcr, oracle, tmp_path):
    """Counts far beyond the file length must be refused before anything is allocated or skipped over."""
    raw, data = new_format_bytes(oracle, tmp_path, d=128, nlist=12, n=300)
    o = offsets(data)
    huge = (1 << 62) + 12345

    def patched(off, fmt, val, cut=None):
        b = bytearray(raw)
        struct.pack_into(fmt, b, off, val)
        return b[:cut] if cut else b

    cases = {
        "nlist_2^31": patched(o["nlist"], "<Q", 1 << 31),
        "nlist_huge": patched(o["nlist"], "<Q", huge),
        "nlist_zero": patched(o["nlist"], "<Q", 0),
        "d_negative": patched(4, "<i", -128),
        "d_huge": patched(4, "<i", 0x7fffffff),
        "ntotal_negative": patched(8, "<q", -1),
        "count_huge": patched(o["count"], "<Q", huge),
        "dm_array_huge": patched(o["dm"] + 1, "<Q", huge),
        "dm_type_bad": patched(o["dm"], "<B", 9),
        "table_cnt_huge": patched(o["cnt"], "<Q", huge),
        "size_huge": patched(o["table"], "<Q", huge),
        "size_negative": patched(o["table"] + 8, "<q", -5),
        "size_sum_over": patched(o["table"] + 16, "<Q", 10 ** 6),
    }
    # consistent but absurd: nlist and count agree on 2^31 lists of d = 128 while the file is 150 KB
    b = bytearray(raw)
    struct.pack_into("<Q", b, o["nlist"], 1 << 31)
    struct.pack_into("<q", b, o["qfourcc"] + 4 + 4, 1 << 31)
    struct.pack_into("<Q", b, o["count"], (1 << 31) * 128)
    cases["consistent_but_absurd"] = b
    # hash direct map with a huge table
    dm = o["dm"]
    cases["dm_hash_huge"] = bytearray(raw[:dm]) + bytes([2]) + struct.pack("<Q", 0) + struct.pack("<Q", huge) + raw[dm + 9:]
    for name, blob in cases.items():
        with pytest.raises(rvcr.RvcrError) as e:
            rvcr.probe_index(write(tmp_path, name + ".index", blob))
        assert e.value.code == -3, name
    # sprs table naming a list that does not exist / with an odd entry count
    raws, datas = new_format_bytes(oracle, tmp_path, nlist=40, empty_every=8)
    os_ = offsets(datas)
    assert bytes(raws[os_["lt"]:os_["lt"] + 4]) == b"sprs"
    bad = bytearray(raws)
    struct.pack_into("<Q", bad, os_["table"], 40)
    with pytest.raises(rvcr.RvcrError):
        rvcr.probe_index(write(tmp_path, "sprs_badlist.index", bad))
    bad = bytearray(raws)
    struct.pack_into("<Q", bad, os_["cnt"], struct.unpack_from("<Q", raws, os_["cnt"])[0] - 1)
    with pytest.raises(rvcr.RvcrError):
        rvcr.probe_index(write(tmp_path, "sprs_odd.index", bad))


def test_probe_every_truncation_and_random_mutations(rvcr, oracle, tmp_path):
    data, _ = make_data(oracle, d=128, nlist=3, n=10, seed=1)
    blobs = {}
    p = str(tmp_path / "small.index")
    oracle.write_index(data, p)
    blobs["IwFl"] = open(p, "rb").read()
    for four in (b"IvFL", b"IvFl"):
        oracle.write_index_legacy(data, p, four)
        blobs[four.decode()] = open(p, "rb").read()
    lib = rvcr.lib()
    path = str(tmp_path / "mut.index").encode()

    def probe(blob):
        with open(path, "wb") as f:
            f.write(blob)
        from importlib import import_module
        ffi = import_module("codename-rvc-fork-3_b200._ffi")
        inf = ffi.FileInfo()
        return lib.rvcr_index_probe(path, inf, None, 0), inf

    rng = np.random.default_rng(11)
    for name, blob in blobs.items():
        rc, inf = probe(blob)
        assert rc == 0 and inf.nvectors == 10, name
        # header and tables: every prefix; payload: a sample of prefixes -- none is a complete file
        head = len(blob) - 10 * (128 * 4 + 8)
        cuts = list(range(0, head + 16)) + sorted(rng.integers(head + 16, len(blob), 60).tolist())
        for cut in cuts:
            rc, _ = probe(blob[:cut])
            assert rc == -3, (name, cut)
        # random byte flips anywhere outside the float payload: any verdict is fine, crashing or lying is not
        for _ in range(400):
            b = bytearray(blob)
            for pos in rng.integers(0, 4 + HDR + 16 + 4 + HDR + 8, rng.integers(1, 4)):
                b[pos] = rng.integers(0, 256)
            tail = rng.integers(len(b) - 400, len(b))
            b[tail] ^= 0xFF
            rc, inf = probe(bytes(b))
            assert rc in (0, -3, -5), (name, rc)
            if rc == 0:
                assert 0 <= inf.nvectors * (4 * inf.d + 8) <= len(b) and 0 < inf.nlist * inf.d * 4 <= len(b)


def test_shard_range_matches_numpy_restatement(rvcr):
    synth = rvcr.synth
    rng = np.random.default_rng(2)
    for trial in range(40):
        nlist = int(rng.integers(1, 200))
        sizes = rng.integers(0, 50, nlist).astype(np.int64)
        if trial % 7 == 0:
            sizes[:] = 0
        if trial % 5 == 0:
            sizes[rng.integers(0, nlist)] = 5000                 # one dominant list
        for world in (1, 2, 3, 4, 8):
            want = synth.shard_ranges(sizes, world)
            got = [rvcr.shard_range(sizes, r, world) for r in range(world)]
            assert got == [tuple(int(v) for v in w) for w in want], (trial, world)
            assert got[0][0] == 0 and got[-1][1] == nlist and all(a[1] == b[0] for a, b in zip(got[:-1], got[1:]))
    with pytest.raises(rvcr.RvcrError):
        rvcr.shard_range(np.array([1, -2, 3]), 0, 2)
    with pytest.raises(rvcr.RvcrError):
        rvcr.shard_range(np.array([1, 2, 3]), 2, 2)


def test_probe_agrees_with_oracle_writer_on_random_valid_files(rvcr, oracle, tmp_path):
    """300 random shapes (d from 4 to 768, 1-40 lists, 0-400 vectors: empty indices, single lists, sprs and full tables)
    in all three layouts: the library's parse and the oracle's reader agree with what was written."""
    rng = np.random.default_rng(0)
    p = str(tmp_path / "r.index")
    for _ in range(300):
        d = int(rng.choice([4, 8, 64, 128, 256, 768]))
        nlist = int(rng.integers(1, 40))
        n = int(rng.choice([0, 1, 5, 50, 400]))
        sizes = np.bincount(rng.integers(0, nlist, n), minlength=nlist).astype(np.int64) if n else np.zeros(nlist, np.int64)
        data = oracle.IVFFlatData(d=d, ntotal=n, nprobe=1, centroids=rng.standard_normal((nlist, d)).astype(np.float32),
                                  list_sizes=sizes, vecs=rng.standard_normal((n, d)).astype(np.float32),
                                  ids=np.arange(n, dtype=np.int64))
        for kind in ("IwFl", "IvFL", "IvFl"):
            if kind == "IwFl":
                oracle.write_index(data, p)
            else:
                oracle.write_index_legacy(data, p, kind.encode())
            inf = rvcr.probe_index(p, list_sizes=True)
            assert (inf["fourcc"], inf["d"], inf["ntotal"], inf["nlist"], inf["nvectors"]) == (kind, d, n, nlist, n)
            assert np.array_equal(inf["list_sizes"], sizes) and inf["file_bytes"] == os.path.getsize(p)
            back = oracle.read_index(p)
            assert np.array_equal(back.list_sizes, sizes) and np.array_equal(back.vecs, data.vecs)
            assert np.array_equal(back.ids, data.ids) and np.array_equal(back.centroids, data.centroids)
