"""numpy restatement of faiss 1.7.3 `write_index` / `read_index` for IndexIVFFlat over IndexFlat.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows faiss impl/index_write.cpp /
index_read.cpp as summarised in SURVEY.md Appendix B; call sites in the reference:
`faiss.read_index(file_index)` rvc/infer/pipeline.py:555 and `faiss.write_index(index, path)`
rvc/train/process/extract_index.py:71.  Little-endian, no padding (the 1-byte `is_trained` at
offset 32 leaves everything after it unaligned).

The product has its own, independent C++ reader/writer (csrc/index_io.cpp); tests compare the
bytes the two produce and a third, hand-assembled struct.pack fixture.
"""
from __future__ import annotations

import io
import struct
from dataclasses import dataclass

import numpy as np

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1
DUMMY = 1 << 20


@dataclass
class IVFFlatData:
    d: int
    ntotal: int
    nprobe: int
    centroids: np.ndarray      # (nlist, d) float32
    list_sizes: np.ndarray     # (nlist,) int64
    vecs: np.ndarray           # (sum(list_sizes), d) float32, lists concatenated in list order
    ids: np.ndarray            # (sum(list_sizes),) int64
    metric_type: int = METRIC_L2
    is_trained: bool = True

    @property
    def nlist(self) -> int:
        return int(self.centroids.shape[0])


def _write_header(f, d, ntotal, is_trained, metric_type):
    f.write(struct.pack("<iqqqBi", d, ntotal, DUMMY, DUMMY, 1 if is_trained else 0, metric_type))


def write_index(data: IVFFlatData, path_or_file) -> None:
    own = isinstance(path_or_file, (str, bytes))
    f = open(path_or_file, "wb") if own else path_or_file
    try:
        nlist, d = data.nlist, data.d
        f.write(b"IwFl")
        _write_header(f, d, data.ntotal, data.is_trained, data.metric_type)
        f.write(struct.pack("<QQ", nlist, data.nprobe))
        # quantizer: IndexFlatL2 / IndexFlatIP
        f.write(b"IxF2" if data.metric_type == METRIC_L2 else
                b"IxFI" if data.metric_type == METRIC_INNER_PRODUCT else b"IxFl")
        _write_header(f, d, nlist, True, data.metric_type)
        f.write(struct.pack("<Q", nlist * d))
        f.write(np.ascontiguousarray(data.centroids, "<f4").tobytes())
        # direct map: NoMap, empty array
        f.write(struct.pack("<BQ", 0, 0))
        # ArrayInvertedLists
        f.write(b"ilar")
        f.write(struct.pack("<QQ", nlist, 4 * d))
        sizes = np.asarray(data.list_sizes, np.int64)
        n_non0 = int((sizes > 0).sum())
        if n_non0 > nlist // 2:
            f.write(b"full")
            f.write(struct.pack("<Q", nlist))
            f.write(sizes.astype("<u8").tobytes())
        else:
            f.write(b"sprs")
            nz = np.nonzero(sizes)[0]
            pairs = np.stack([nz, sizes[nz]], axis=1).astype("<u8")
            f.write(struct.pack("<Q", pairs.size))
            f.write(pairs.tobytes())
        off = 0
        vecs = np.ascontiguousarray(data.vecs, "<f4")
        ids = np.ascontiguousarray(data.ids, "<i8")
        for n in sizes:
            n = int(n)
            if n > 0:
                f.write(vecs[off:off + n].tobytes())
                f.write(ids[off:off + n].tobytes())
                off += n
    finally:
        if own:
            f.close()


def write_index_legacy(data: IVFFlatData, path_or_file, fourcc: bytes = b"IvFL") -> None:
    """The pre-1.5 IndexIVFFlat layouts faiss 1.7.3 still reads (impl/index_read.cpp, `h == fourcc("IvFl") ||
    h == fourcc("IvFL")` branch): ivf header, quantizer, then `READVECTOR(ids[i])` for every list
    (read_ivf_header with `ids`), the direct map, then one vector of codes per list -- bytes for b"IvFL",
    floats for the oldest b"IvFl".  faiss no longer WRITES these; this writer exists to test readers."""
    assert fourcc in (b"IvFL", b"IvFl")
    own = isinstance(path_or_file, (str, bytes))
    f = open(path_or_file, "wb") if own else path_or_file
    try:
        nlist, d = data.nlist, data.d
        f.write(fourcc)
        _write_header(f, d, data.ntotal, data.is_trained, data.metric_type)
        f.write(struct.pack("<QQ", nlist, data.nprobe))
        f.write(b"IxF2" if data.metric_type == METRIC_L2 else b"IxFI")
        _write_header(f, d, nlist, True, data.metric_type)
        f.write(struct.pack("<Q", nlist * d))
        f.write(np.ascontiguousarray(data.centroids, "<f4").tobytes())
        sizes = [int(n) for n in np.asarray(data.list_sizes, np.int64)]
        ids = np.ascontiguousarray(data.ids, "<i8")
        vecs = np.ascontiguousarray(data.vecs, "<f4")
        off = 0
        for n in sizes:
            f.write(struct.pack("<Q", n))
            f.write(ids[off:off + n].tobytes())
            off += n
        f.write(struct.pack("<BQ", 0, 0))                      # maintain_direct_map = false, empty array
        off = 0
        for n in sizes:
            f.write(struct.pack("<Q", n * d * (4 if fourcc == b"IvFL" else 1)))   # byte count / float count
            f.write(vecs[off:off + n].tobytes())
            off += n
    finally:
        if own:
            f.close()


def _rd(f, fmt):
    n = struct.calcsize(fmt)
    b = f.read(n)
    if len(b) != n:
        raise ValueError("truncated index file")
    return struct.unpack(fmt, b)


def _read_header(f):
    d, ntotal, _, _, trained, metric = _rd(f, "<iqqqBi")
    if metric > 1:
        _rd(f, "<f")  # metric_arg
    return d, ntotal, bool(trained), metric


def _read_legacy_lists(f, four, d, ntotal, nprobe, cent, metric, trained) -> IVFFlatData:
    nlist = cent.shape[0]
    id_parts = []
    for _ in range(nlist):
        (n,) = _rd(f, "<Q")
        b = f.read(8 * n)
        if len(b) != 8 * n:
            raise ValueError("truncated index file")
        id_parts.append(np.frombuffer(b, "<i8"))
    (dm_type,) = _rd(f, "<B")
    (dm_n,) = _rd(f, "<Q")
    f.read(8 * dm_n)
    if dm_type == 2:
        (hn,) = _rd(f, "<Q")
        f.read(16 * hn)
    unit = 1 if four == b"IvFL" else 4
    vec_parts = []
    for l in range(nlist):
        (m,) = _rd(f, "<Q")
        if m * unit != len(id_parts[l]) * 4 * d:
            raise ValueError("legacy list: id / code counts disagree")
        b = f.read(m * unit)
        if len(b) != m * unit:
            raise ValueError("truncated index file")
        vec_parts.append(np.frombuffer(b, "<f4").reshape(-1, d))
    sizes = np.array([len(i) for i in id_parts], np.int64)
    return IVFFlatData(d=d, ntotal=ntotal, nprobe=nprobe, centroids=cent, list_sizes=sizes,
                       vecs=np.concatenate(vec_parts).astype(np.float32) if nlist else np.empty((0, d), np.float32),
                       ids=np.concatenate(id_parts).astype(np.int64) if nlist else np.empty(0, np.int64),
                       metric_type=metric, is_trained=trained)


def read_index(path_or_bytes) -> IVFFlatData:
    if isinstance(path_or_bytes, (bytes, bytearray)):
        f = io.BytesIO(path_or_bytes)
        own = False
    else:
        f = open(path_or_bytes, "rb")
        own = True
    try:
        four = f.read(4)
        if four not in (b"IwFl", b"IvFL", b"IvFl"):
            raise ValueError(f"unsupported index fourcc {four!r} (expected b'IwFl' or legacy b'IvFL' / b'IvFl')")
        d, ntotal, trained, metric = _read_header(f)
        nlist, nprobe = _rd(f, "<QQ")
        qf = f.read(4)
        if qf not in (b"IxF2", b"IxFI", b"IxFl"):
            raise ValueError(f"unsupported quantizer fourcc {qf!r}")
        qd, qn, _, _ = _read_header(f)
        (count,) = _rd(f, "<Q")
        if qd != d or qn != nlist or count != nlist * d:
            raise ValueError("inconsistent quantizer header")
        cent = np.frombuffer(f.read(4 * count), "<f4").reshape(nlist, d).copy()
        if four != b"IwFl":
            return _read_legacy_lists(f, four, d, ntotal, nprobe, cent, metric, trained)
        (dm_type,) = _rd(f, "<B")
        (dm_n,) = _rd(f, "<Q")
        f.read(8 * dm_n)
        if dm_type == 2:
            (hn,) = _rd(f, "<Q")
            f.read(16 * hn)
        il = f.read(4)
        if il == b"il00":                      # faiss write_InvertedLists(nullptr): the fourcc and nothing else
            sizes = np.zeros(nlist, np.int64)
        elif il == b"ilar":
            il_nlist, code_size = _rd(f, "<QQ")
            if il_nlist != nlist or code_size != 4 * d:
                raise ValueError("inconsistent inverted-list header")
            lt = f.read(4)
            (cnt,) = _rd(f, "<Q")
            raw = np.frombuffer(f.read(8 * cnt), "<u8").astype(np.int64)
            if lt == b"full":
                if cnt != nlist:
                    raise ValueError("bad 'full' size vector")
                sizes = raw.copy()
            elif lt == b"sprs":
                sizes = np.zeros(nlist, np.int64)
                pairs = raw.reshape(-1, 2)
                sizes[pairs[:, 0]] = pairs[:, 1]
            else:
                raise ValueError(f"unsupported list type {lt!r}")
        else:
            raise ValueError(f"unsupported inverted lists fourcc {il!r}")
        tot = int(sizes.sum())
        vecs = np.empty((tot, d), np.float32)
        ids = np.empty(tot, np.int64)
        off = 0
        for n in sizes:
            n = int(n)
            if n > 0:
                b = f.read(4 * d * n)
                if len(b) != 4 * d * n:
                    raise ValueError("truncated index file")
                vecs[off:off + n] = np.frombuffer(b, "<f4").reshape(n, d)
                b = f.read(8 * n)
                if len(b) != 8 * n:
                    raise ValueError("truncated index file")
                ids[off:off + n] = np.frombuffer(b, "<i8")
                off += n
        return IVFFlatData(d=d, ntotal=ntotal, nprobe=nprobe, centroids=cent, list_sizes=sizes,
                           vecs=vecs, ids=ids, metric_type=metric, is_trained=trained)
    finally:
        if own:
            f.close()
