"""CPU oracle for the RVC feature-retrieval path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package, and only as the checker (or as the timed CPU baseline).  The product
package never imports it.

PARITY UNPINNED for the faiss part (coarse assignment, list scan, top-k, add): faiss-cpu==1.7.3
is neither vendored in the reference nor installable here, and the reference holds no golden
vectors for it.  The numpy/torch tail (pipeline.py:500-506) IS pinned: tests/golden holds outputs
of the reference's own `_retrieve_speaker_embeddings` body executed in this container.

Functions mirror the reference call sites:

    coarse_assign      index.search -> IndexFlatL2::search(k=1)      pipeline.py:499 / extract_index.py:69
    scan_topk          index.search -> IVFFlatScanner + heap          pipeline.py:499
    tail               weights, gather-sum, blend                     pipeline.py:500-506
    retrieve           the whole of _retrieve_speaker_embeddings      pipeline.py:497-507
    build_lists        IndexIVFFlat::add_core                         extract_index.py:67-69
    kmeans_train       index.train -> faiss Clustering::train (A9)    extract_index.py:65   (PARITY UNPINNED)
    final_proj         v1 query producer, nn.Linear(768, 256)         pipeline.py:451-453  (PINNED within 1e-5)
    coarse_topn / scan_topk_multi   index.search with nprobe > 1      SURVEY A3 / 8f-4     (PARITY UNPINNED)
    upsample_protect   2x nearest upsample + pitch-protect mix        pipeline.py:463-481  (PINNED:
                       tests/golden/post_*.npz are outputs of the reference's own statements, make_golden_post.py)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import build as _build
from .index_io import IVFFlatData, read_index, write_index, write_index_legacy  # noqa: F401

FLT_MAX = np.float32(3.4028234663852886e38)

_lib = None


def lib():
    global _lib
    if _lib is None:
        so = _build.build()
        L = C.CDLL(so)
        fp, ip, lp, dp = (C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                          C.POINTER(C.c_double))
        L.rvco_l2sqr.restype = C.c_float
        L.rvco_l2sqr.argtypes = [fp, fp, C.c_int]
        L.rvco_l2sqr_nofma.restype = C.c_float
        L.rvco_l2sqr_nofma.argtypes = [fp, fp, C.c_int]
        L.rvco_l2sqr_f64.restype = C.c_double
        L.rvco_l2sqr_f64.argtypes = [fp, fp, C.c_int]
        L.rvco_num_threads.restype = C.c_int
        L.rvco_set_num_threads.restype = None
        L.rvco_set_num_threads.argtypes = [C.c_int]
        L.rvco_coarse_assign.restype = None
        L.rvco_coarse_assign.argtypes = [fp, C.c_int64, fp, C.c_int64, C.c_int, C.c_int, ip, fp, dp]
        L.rvco_coarse_topn.restype = None
        L.rvco_coarse_topn.argtypes = [fp, C.c_int64, fp, C.c_int64, C.c_int, C.c_int, ip, fp]
        L.rvco_scan_topk.restype = None
        L.rvco_scan_topk.argtypes = [fp, C.c_int64, ip, lp, fp, lp, C.c_int, C.c_int, C.c_int, fp, lp, lp]
        L.rvco_tail.restype = None
        L.rvco_tail.argtypes = [fp, lp, C.c_int64, C.c_int, fp, C.c_int64, C.c_int, C.c_double, fp, fp]
        L.rvco_build_lists.restype = None
        L.rvco_build_lists.argtypes = [fp, C.c_int64, ip, C.c_int64, C.c_int, lp, fp, lp]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, ty):
    return a.ctypes.data_as(C.POINTER(ty))


def num_threads() -> int:
    return int(lib().rvco_num_threads())


def set_num_threads(n: int) -> int:
    """OpenMP threads of the C restatement (scan, canonical coarse, tail); returns the count now in force."""
    lib().rvco_set_num_threads(int(n))
    return num_threads()


def l2sqr(x, y, fma: bool = True) -> np.float32:
    """Canonical fp32 squared L2 of two vectors (faiss fvec_L2sqr AVX order)."""
    x, y = _f32(x), _f32(y)
    assert x.ndim == 1 and x.shape == y.shape and x.shape[0] % 8 == 0
    f = lib().rvco_l2sqr if fma else lib().rvco_l2sqr_nofma
    return np.float32(f(_p(x, C.c_float), _p(y, C.c_float), x.shape[0]))


def coarse_assign(q, centroids, fma: bool = True, want_gap: bool = False):
    """Nearest centroid per query, lowest id on ties.  Brute force, canonical distance.

    Returns (list_id int32[nq], dist float32[nq][, gap float64[nq]])."""
    q, cent = _f32(q), _f32(centroids)
    nq, d = q.shape
    nlist = cent.shape[0]
    assert cent.shape[1] == d and d % 8 == 0
    out = np.empty(nq, np.int32)
    dist = np.empty(nq, np.float32)
    gap = np.empty(nq, np.float64) if want_gap else None
    lib().rvco_coarse_assign(_p(q, C.c_float), nq, _p(cent, C.c_float), nlist, d, 0 if fma else 1,
                             _p(out, C.c_int32), _p(dist, C.c_float),
                             _p(gap, C.c_double) if want_gap else None)
    return (out, dist, gap) if want_gap else (out, dist)


def coarse_topn(q, centroids, nprobe: int):
    """The nprobe nearest centroids per query (canonical distance, ascending; lower id first among equal distances):
    quantizer->search(n, x, nprobe) of IndexIVF::search.  Returns (lists int32[nq, nprobe], dist f32[nq, nprobe]);
    missing entries (NaN queries, nprobe > nlist) are -1 / FLT_MAX."""
    q, cent = _f32(q), _f32(centroids)
    nq, d = q.shape
    nprobe = int(nprobe)
    assert nprobe >= 1 and cent.shape[1] == d and d % 8 == 0
    out = np.empty((nq, nprobe), np.int32)
    dist = np.empty((nq, nprobe), np.float32)
    lib().rvco_coarse_topn(_p(q, C.c_float), nq, _p(cent, C.c_float), cent.shape[0], d, nprobe, _p(out, C.c_int32),
                           _p(dist, C.c_float))
    return out, dist


def scan_topk_multi(q, lists, list_off, vecs, ids, k: int = 8, want_pos: bool = False):
    """IndexIVF::search_preassigned with nprobe > 1: the probed lists are scanned in coarse-rank order
    into ONE k-heap with the strict replacement rule, so among equal distances the entry scanned first (lower probe
    rank, then lower list position) is the one kept / listed first.  Restated as: top-k of each list (same scan as
    nprobe = 1), concatenated in probe order, stable sort by distance."""
    lists = np.ascontiguousarray(lists, np.int32)
    nq, nprobe = lists.shape
    Ds, Is, Ps = [], [], []
    for p in range(nprobe):
        D, I, P = scan_topk(q, lists[:, p], list_off, vecs, ids, k=k, want_pos=True)
        Ds.append(D), Is.append(I), Ps.append(P)
    D, I, P = np.concatenate(Ds, 1), np.concatenate(Is, 1), np.concatenate(Ps, 1)
    order = np.argsort(D, axis=1, kind="stable")[:, :k]
    rows = np.arange(nq)[:, None]
    D, I, P = D[rows, order], I[rows, order], P[rows, order]
    return (D, I, P) if want_pos else (D, I)


def coarse_assign_blas(q, centroids, block: int = 4096):
    """faiss's BLAS path for >= 20 queries: ||x||^2 + ||y||^2 - 2 x.y via sgemm,
    clamped at 0, argmin with lowest id on ties.  Used as the timed CPU baseline for the coarse
    step and for tie-zone classification; not bit-comparable to the canonical distance."""
    q, cent = _f32(q), _f32(centroids)
    yn = np.einsum("ij,ij->i", cent, cent).astype(np.float32)
    out = np.empty(q.shape[0], np.int32)
    dist = np.empty(q.shape[0], np.float32)
    for s in range(0, q.shape[0], block):
        x = q[s:s + block]
        xn = np.einsum("ij,ij->i", x, x).astype(np.float32)
        ip = x @ cent.T
        dis = xn[:, None] + yn[None, :] - np.float32(2.0) * ip
        np.maximum(dis, 0, out=dis)
        j = dis.argmin(axis=1)  # first occurrence == lowest id
        out[s:s + block] = j
        dist[s:s + block] = dis[np.arange(len(j)), j]
    return out, dist


def scan_topk(q, assign, list_off, vecs, ids, k: int = 8, fma: bool = True, want_pos: bool = False):
    """Single-list scan + top-k per query.  Returns (D f32[nq,k], I i64[nq,k][, P])."""
    q, vecs = _f32(q), _f32(vecs)
    assign = np.ascontiguousarray(assign, np.int32)
    list_off = np.ascontiguousarray(list_off, np.int64)
    ids = np.ascontiguousarray(ids, np.int64)
    nq, d = q.shape
    D = np.empty((nq, k), np.float32)
    I = np.empty((nq, k), np.int64)
    P = np.empty((nq, k), np.int64) if want_pos else None
    vp = _p(vecs, C.c_float) if vecs.size else None
    lib().rvco_scan_topk(_p(q, C.c_float), nq, _p(assign, C.c_int32), _p(list_off, C.c_int64), vp,
                         _p(ids, C.c_int64) if ids.size else None, d, k, 0 if fma else 1,
                         _p(D, C.c_float), _p(I, C.c_int64), _p(P, C.c_int64) if want_pos else None)
    return (D, I, P) if want_pos else (D, I)


def tail(score, ix, big_npy, index_rate: float, feats):
    """pipeline.py:500-506 in exact fp32 operation order."""
    score, big_npy, feats = _f32(score), _f32(big_npy), _f32(feats)
    ix = np.ascontiguousarray(ix, np.int64)
    nq, k = score.shape
    ntotal, d = big_npy.shape
    assert feats.shape == (nq, d)
    out = np.empty((nq, d), np.float32)
    lib().rvco_tail(_p(score, C.c_float), _p(ix, C.c_int64), nq, k, _p(big_npy, C.c_float), ntotal, d,
                    float(index_rate), _p(feats, C.c_float), _p(out, C.c_float))
    return out


def tail_numpy_verbatim(score, ix, big_npy, index_rate, feats):
    """The reference's own statements (pipeline.py:500-506) with torch replaced by numpy fp32 for
    the blend line; used to cross-check `tail` and as the timed CPU baseline's tail."""
    with np.errstate(all="ignore"):
        weight = np.square(1 / score)
        weight /= weight.sum(axis=1, keepdims=True)
        npy = np.sum(big_npy[ix] * np.expand_dims(weight, axis=2), axis=1)
    r32 = np.float32(index_rate)
    omr32 = np.float32(1 - index_rate)
    return npy * r32 + omr32 * feats


def rand_perm(n: int, seed: int) -> np.ndarray:
    """faiss utils/random.cpp rand_perm [EXT, SURVEY A9]: identity, then for i in 0..n-2 swap(perm[i],
    perm[i + mt() % (n - i)]) with std::mt19937(seed).  numpy's legacy RandomState(seed) seeds MT19937 with the
    same init_genrand(seed); random_raw yields its raw 32-bit outputs."""
    raw = np.random.RandomState(seed)._bit_generator.random_raw(max(n - 1, 0)).astype(np.uint64)
    perm = list(range(n))
    for i in range(n - 1):
        j = i + int(raw[i] % np.uint64(n - i))
        perm[i], perm[j] = perm[j], perm[i]
    return np.asarray(perm, np.int64)


def split_clusters(hass, cent, n: int) -> int:
    """faiss Clustering.cpp split_clusters [EXT, SURVEY A9] in place on hass (k float32 counts) and cent (k, d) float32:
    every empty cluster ci takes a copy of a donor cj drawn with probability (hass[cj] - 1) / (n - k) by walking
    cj = 0, 1, ... with one rand_float() per step; the copy and the donor are perturbed by 1 +- 1/1024 with alternating
    sign per dimension; the donor's count is halved.  `RandomGenerator rng(1234)` is a local of the function, i.e.
    re-seeded on every call; rand_float() = mt() / float(mt.max()) in fp32."""
    k, d = cent.shape
    eps = np.float32(1.0 / 1024.0)
    sign = np.where(np.arange(d) % 2 == 0, eps, -eps).astype(np.float32)
    rs = np.random.RandomState(1234)
    nsplit = 0
    for ci in np.nonzero(hass == 0)[0]:
        cj = 0
        while True:
            p = np.float32((np.float64(hass[cj]) - 1.0) / np.float64(np.float32(n - k)))
            r = np.float32(rs._bit_generator.random_raw(1)[0]) / np.float32(4294967295.0)   # mt() / float(mt.max())
            if r < p:
                break
            cj = (cj + 1) % k
        cent[ci] = cent[cj]
        cent[ci] *= (np.float32(1) + sign)
        cent[cj] *= (np.float32(1) - sign)
        hass[ci] = hass[cj] / np.float32(2)
        hass[cj] -= hass[ci]
        nsplit += 1
    return nsplit


def kmeans_train(x, k: int, niter: int = 10, seed: int = 1234, max_points_per_centroid: int = 256):
    """faiss 1.7.3 Clustering::train as IndexIVF::train runs it [EXT -- PARITY UNPINNED: restated
    from the published algorithm, faiss is not installable here].  Returns (centroids (k, d) f32, assign of the
    LAST iteration, objective per iteration).  Means accumulate in fp32 in point order like faiss's
    compute_centroids; empty clusters are re-seeded by split_clusters (EPS = 1/1024, RandomGenerator(1234))."""
    x = _f32(x)
    n, d = x.shape
    if n < k:
        raise RuntimeError(f"Number of training points ({n}) should be at least as large as number of clusters ({k})")
    if n > k * max_points_per_centroid:
        x = np.ascontiguousarray(x[rand_perm(n, seed)[: k * max_points_per_centroid]])
        n = x.shape[0]
    cent = np.ascontiguousarray(x[rand_perm(n, seed + 1)[:k]])
    obj = []
    a = None
    for _ in range(niter):
        a, dist = coarse_assign(x, cent)
        obj.append(float(dist.astype(np.float64).sum()))
        sums = np.zeros((k, d), np.float32)
        np.add.at(sums, a, x)                               # fp32, point order (faiss compute_centroids)
        hass = np.bincount(a, minlength=k).astype(np.float32)          # faiss keeps hassign as float
        nz = hass > 0
        cent = cent.copy()
        cent[nz] = sums[nz] * (np.float32(1.0) / hass[nz])[:, None]       # faiss: norm = 1 / hassign; c *= norm
        split_clusters(hass, cent, n)
    return cent.astype(np.float32), a, obj


def upsample_protect(feats, feats0=None, pitchf=None, protect: float = 0.5):
    """pipeline.py:463-481 on (n, d) arrays -> (2n, d): F.interpolate(scale_factor=2) along time is nearest
    (out[t] = in[t // 2]); with a pitch track, pitchff = pitchf; pitchff[pitchf > 0] = 1; pitchff[pitchf < 1] =
    protect, and out = feats * pitchff + feats0 * (1 - pitchff): three separately rounded fp32 operations.
    The CALLER applies the reference's conditions (pitch guidance and protect < 0.5) by passing pitchf or None."""
    up = np.repeat(_f32(feats), 2, axis=0)
    if pitchf is None or feats0 is None:
        return up
    up0 = np.repeat(_f32(feats0), 2, axis=0)
    pitchf = _f32(pitchf).reshape(-1)
    assert pitchf.shape[0] == up.shape[0], "pitchf must have 2n entries (the reference broadcasts it over 2n frames)"
    pf = pitchf.copy()
    pf[pitchf > 0] = np.float32(1)
    pf[pitchf < 1] = np.float32(protect)
    om = (np.float32(1) - pf).astype(np.float32)
    return (up * pf[:, None]).astype(np.float32) + (up0 * om[:, None]).astype(np.float32)


def final_proj(x, weight, bias=None):
    """`model.final_proj(feats[0])` for v1 models (pipeline.py:451-453): nn.Linear, out = x W^T + b in float32.  Floating
    point: numpy's BLAS sums in its own order, like torch's -- comparisons use a 1e-5 relative tolerance (PINNED within that
    on tests/golden/final_proj_v1.npz, produced by the reference statement itself)."""
    out = _f32(x) @ _f32(weight).T
    if bias is not None:
        out = out + _f32(bias)[None, :]
    return out.astype(np.float32)


def build_lists(x, assign, nlist: int):
    """IndexIVFFlat::add_core over the whole data set: returns (list_off, vecs, ids)."""
    x = _f32(x)
    assign = np.ascontiguousarray(assign, np.int32)
    n, d = x.shape
    off = np.empty(nlist + 1, np.int64)
    kept = int((assign >= 0).sum())
    vecs = np.empty((kept, d), np.float32)
    ids = np.empty(kept, np.int64)
    lib().rvco_build_lists(_p(x, C.c_float), n, _p(assign, C.c_int32), nlist, d, _p(off, C.c_int64),
                           _p(vecs, C.c_float) if kept else None, _p(ids, C.c_int64) if kept else None)
    return off, vecs, ids


@dataclass
class OracleIndex:
    """In-memory IndexIVFFlat restatement: quantizer centroids + list-ordered storage."""
    d: int
    centroids: np.ndarray        # (nlist, d) f32
    list_off: np.ndarray         # (nlist+1,) i64
    vecs: np.ndarray             # (ntotal_stored, d) f32, list order
    ids: np.ndarray              # (ntotal_stored,) i64
    ntotal: int
    nprobe: int = 1

    @property
    def nlist(self) -> int:
        return int(self.centroids.shape[0])

    @classmethod
    def from_vectors(cls, x, centroids) -> "OracleIndex":
        """extract_index.py:62-69 given trained centroids: assign every vector, append in id order."""
        x, centroids = _f32(x), _f32(centroids)
        a, _ = coarse_assign(x, centroids)
        off, vecs, ids = build_lists(x, a, centroids.shape[0])
        return cls(x.shape[1], centroids, off, vecs, ids, x.shape[0])

    @classmethod
    def from_data(cls, data: IVFFlatData) -> "OracleIndex":
        off = np.zeros(data.nlist + 1, np.int64)
        np.cumsum(data.list_sizes, out=off[1:])
        return cls(data.d, data.centroids, off, data.vecs, data.ids, data.ntotal, data.nprobe)

    def to_data(self) -> IVFFlatData:
        return IVFFlatData(d=self.d, ntotal=self.ntotal, nprobe=self.nprobe, centroids=self.centroids,
                           list_sizes=np.diff(self.list_off).astype(np.int64), vecs=self.vecs, ids=self.ids)

    def reconstruct_n(self, i0: int = 0, ni: int | None = None) -> np.ndarray:
        """IndexIVF::reconstruct_n: row (id - i0) = stored vector with that id."""
        ni = self.ntotal - i0 if ni is None else ni
        out = np.zeros((ni, self.d), np.float32)
        sel = (self.ids >= i0) & (self.ids < i0 + ni)
        out[self.ids[sel] - i0] = self.vecs[sel]
        return out

    def search(self, q, k: int = 8, fma: bool = True, nprobe: int | None = None):
        nprobe = min(self.nlist, self.nprobe if nprobe is None else int(nprobe))     # IndexIVF::search clamps to nlist
        if nprobe > 1:
            lists, _ = coarse_topn(q, self.centroids, nprobe)
            return scan_topk_multi(q, lists, self.list_off, self.vecs, self.ids, k=k)
        a, _ = coarse_assign(q, self.centroids, fma=fma)
        return scan_topk(q, a, self.list_off, self.vecs, self.ids, k=k, fma=fma)


def retrieve(feats, index: OracleIndex, big_npy, index_rate: float, k: int = 8):
    """_retrieve_speaker_embeddings (pipeline.py:497-507) on a (nq, d) array."""
    feats = _f32(feats)
    score, ix = index.search(feats, k=k)
    with np.errstate(all="ignore"):
        return tail(score, ix, big_npy, index_rate, feats), score, ix
