"""GPU: `.index` files as other writers produce them and concurrent use of one handle (SURVEY 8b
"Threading": one VoiceConverter is shared process-wide and Gradio calls it from worker threads).

Files are assembled from the oracle writer's bytes by editing the header fields a different faiss build / fork
would write: the legacy flat-quantizer fourcc `IxFl`, a populated array direct map, a populated hash direct map,
nprobe != 1 in the header.  A non-L2 metric and a legacy index fourcc `IvFl` stamped on a current-layout body (it does not
parse as the legacy layout) must be refused loudly; genuine legacy files are covered by tests/test_gpu_zz_legacy_index.py
and, at the parser, by tests/test_index_probe.py."""
import struct
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HDR = 4 + 8 + 8 + 8 + 1 + 4        # d, ntotal, dummy, dummy, is_trained, metric_type


def base_file(oracle, tmp_path, d=256):
    rng = np.random.default_rng(8)
    nlist = 16
    cent = (rng.standard_normal((nlist, d)) * 0.5).astype(np.float32)
    x = (cent[rng.integers(0, nlist, 700)] + rng.standard_normal((700, d)).astype(np.float32) * 0.1).astype(np.float32)
    oidx = oracle.OracleIndex.from_vectors(x, cent)
    p = str(tmp_path / "base.index")
    oracle.write_index(oidx.to_data(), p)
    raw = bytearray(open(p, "rb").read())
    q = (x[:64] + np.float32(0.02)).astype(np.float32)
    # offsets of the fields that get edited
    off_nprobe = 4 + HDR + 8
    off_qfourcc = 4 + HDR + 16
    off_dm = off_qfourcc + 4 + HDR + 8 + nlist * d * 4
    assert raw[off_qfourcc:off_qfourcc + 4] == b"IxF2" and raw[off_dm] == 0
    assert struct.unpack_from("<Q", raw, off_dm + 1)[0] == 0 and raw[off_dm + 9:off_dm + 13] == b"ilar"
    return dict(raw=raw, x=x, q=q, oidx=oidx, off_nprobe=off_nprobe, off_qfourcc=off_qfourcc, off_dm=off_dm)


def check_same_answers(rvcr, path, b):
    index = rvcr.read_index(path)
    assert index.ntotal == len(b["x"]) and np.array_equal(index.reconstruct_n(0, index.ntotal), b["x"])
    D, I = index.search(b["q"], k=8)
    Do, Io = b["oidx"].search(b["q"], k=8)
    assert np.array_equal(I, Io) and np.array_equal(D, Do)
    return index


def test_files_from_other_writers(rvcr, oracle, tmp_path):
    b = base_file(oracle, tmp_path)
    raw, n = b["raw"], len(b["x"])
    # legacy flat quantizer fourcc
    f1 = bytearray(raw)
    f1[b["off_qfourcc"]:b["off_qfourcc"] + 4] = b"IxFl"
    (tmp_path / "ixfl.index").write_bytes(f1)
    check_same_answers(rvcr, str(tmp_path / "ixfl.index"), b)
    # array direct map (type 1): ntotal x int64 list-offset codes after the type byte
    dm = b["off_dm"]
    f2 = bytearray(raw[:dm]) + bytes([1]) + struct.pack("<Q", n) + np.arange(n, dtype=np.int64).tobytes() + raw[dm + 9:]
    (tmp_path / "dm_array.index").write_bytes(f2)
    check_same_answers(rvcr, str(tmp_path / "dm_array.index"), b)
    # hash direct map (type 2): empty array, then a vector of (id, lo) pairs
    pairs = np.stack([np.arange(n, dtype=np.int64), np.arange(n, dtype=np.int64)], axis=1).tobytes()
    f3 = bytearray(raw[:dm]) + bytes([2]) + struct.pack("<Q", 0) + struct.pack("<Q", n) + pairs + raw[dm + 9:]
    (tmp_path / "dm_hash.index").write_bytes(f3)
    check_same_answers(rvcr, str(tmp_path / "dm_hash.index"), b)
    # nprobe taken from the header and honoured (faiss does: IndexIVF::search reads the stored field)
    f4 = bytearray(raw)
    struct.pack_into("<Q", f4, b["off_nprobe"], 4)
    (tmp_path / "np4.index").write_bytes(f4)
    idx = rvcr.read_index(str(tmp_path / "np4.index"))
    assert idx.nprobe == 4
    D, I = idx.search(b["q"], k=8)
    Do, Io = b["oidx"].search(b["q"], k=8, nprobe=4)
    assert np.array_equal(I, Io) and np.array_equal(D, Do)
    idx.nprobe = 1
    D, I = idx.search(b["q"], k=8)
    assert np.array_equal(I, b["oidx"].search(b["q"], k=8)[1])
    # refused loudly: legacy IndexIVFFlat fourcc on a current-layout body, inner-product metric, truncated payload
    f5 = bytearray(raw)
    f5[0:4] = b"IvFl"
    (tmp_path / "ivfl.index").write_bytes(f5)
    f6 = bytearray(raw)
    struct.pack_into("<i", f6, 4 + HDR - 4, 0)               # METRIC_INNER_PRODUCT
    (tmp_path / "ip.index").write_bytes(f6)
    (tmp_path / "trunc.index").write_bytes(raw[:len(raw) - 1000])
    for name, code in (("ivfl.index", -3), ("ip.index", -5), ("trunc.index", -3)):
        with pytest.raises(rvcr.RvcrError) as e:
            rvcr.read_index(str(tmp_path / name))
        assert e.value.code == code, name


def test_concurrent_calls_on_one_handle(rvcr, oracle, tmp_path):
    """Worker threads share one index object: every thread searches / blends its own clip on its own CUDA stream;
    answers equal the single-threaded ones (calls on a handle are serialised by its mutex and ordered across
    streams by an event)."""
    import torch
    b = base_file(oracle, tmp_path, d=768)
    index = check_same_answers(rvcr, str(tmp_path / "base.index"), b)
    rng = np.random.default_rng(4)
    clips = [(b["x"][rng.integers(0, len(b["x"]), 50 + 37 * t)] + np.float32(0.01 * (t + 1))).astype(np.float32)
             for t in range(6)]
    want = []
    for c in clips:
        with np.errstate(all="ignore"):
            out, score, ix = oracle.retrieve(c, b["oidx"], b["x"], 0.75)
        want.append((out, score, ix))
    got = [None] * len(clips)
    errs = []

    def worker(t):
        try:
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                for _ in range(5):
                    feats = torch.from_numpy(clips[t].copy()).cuda()
                    _, D, I = index.retrieve_blend_(feats, 0.75, want_neighbours=True)
                    stream.synchronize()
                    got[t] = (feats.cpu().numpy(), D.cpu().numpy(), I.cpu().numpy())
        except Exception as ex:   # surfaced below
            errs.append(ex)

    th = [threading.Thread(target=worker, args=(t,)) for t in range(len(clips))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for (o, s, i), (go, gs, gi) in zip(want, got):
        assert np.array_equal(gi, i) and np.array_equal(gs, s) and np.array_equal(go, o, equal_nan=True)
