"""faiss-shaped host objects over the C ABI, so the reference call sites run unmodified.

Mirrors exactly the surface the reference uses:

    rvc/infer/pipeline.py:555-556      index = faiss.read_index(path); big_npy = index.reconstruct_n(0, index.ntotal)
    rvc/infer/pipeline.py:456-457      if index:                      (object must be truthy)
    rvc/infer/pipeline.py:499          score, ix = index.search(npy, k=8)
    rvc/train/process/extract_index.py:62-71
                                       index_factory(768, "IVF{n},Flat"); extract_index_ivf(i).nprobe = 1;
                                       i.train(x); i.add(x[a:b]); write_index(i, path)

Same names, argument meaning and error behaviour as faiss's Python wrapper: `search`
wants an (n, d) array, casts to C-contiguous float32 and returns new (n, k) float32 / int64 arrays;
failures raise (RvcrError is a RuntimeError) and the reference's own try/except handles them.
All arithmetic runs in librvcr.so on the GPU; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import threading
import weakref
from collections import OrderedDict

import numpy as np

from . import _ffi
from ._ffi import Info, RvcrError, Stats, check, lib

METRIC_L2 = 1


def _default_device() -> int:
    try:
        import torch
        if torch.cuda.is_available():
            return int(torch.cuda.current_device())
    except Exception:
        pass
    return 0


def _as_f32_2d(x, d: int, what: str) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim != 2 or x.shape[1] != d:
        raise AssertionError(f"{what}: expected shape (n, {d}), got {x.shape}")
    return x


class IndexIVFFlat:
    """Device-resident IVF-Flat index (faiss.IndexIVFFlat stand-in)."""

    def __init__(self, handle: int, owner: bool = True):
        self._h = C.c_void_p(handle)
        self._owner = owner

    # -- lifecycle -------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value and self._owner:
            lib().rvcr_index_free(self._h)
        self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __bool__(self) -> bool:  # `if index:` at pipeline.py:456-457
        return True

    @property
    def handle(self) -> C.c_void_p:
        if not self._h.value:
            raise RvcrError(-6, "index handle is closed")
        return self._h

    def _info(self) -> Info:
        inf = Info()
        check(lib().rvcr_index_info(self.handle, C.byref(inf)))
        return inf

    # -- faiss attributes ------------------------------------------------------------------
    @property
    def d(self) -> int:
        return int(self._info().d)

    @property
    def ntotal(self) -> int:
        return int(self._info().ntotal)

    @property
    def nlist(self) -> int:
        return int(self._info().nlist)

    @property
    def is_trained(self) -> bool:
        return bool(self._info().is_trained)

    @property
    def metric_type(self) -> int:
        return int(self._info().metric)

    @property
    def nprobe(self) -> int:
        return int(self._info().nprobe)

    @nprobe.setter
    def nprobe(self, v: int) -> None:  # extract_index.py:64
        check(lib().rvcr_index_set_nprobe(self.handle, int(v)))

    @property
    def device(self) -> int:
        return int(self._info().device)

    @property
    def info(self) -> dict:
        inf = self._info()
        return {k: int(getattr(inf, k)) for k, _ in Info._fields_}

    # -- quantizer -------------------------------------------------------------------------
    def set_centroids(self, centroids) -> None:
        inf = self._info()
        c = _as_f32_2d(centroids, inf.d, "set_centroids")
        check(lib().rvcr_index_set_centroids(self.handle, c.ctypes.data_as(C.c_void_p), c.shape[0]))

    def set_shard(self, owned_begin: int, owned_end: int) -> None:
        """This handle stores only lists [owned_begin, owned_end); the quantizer stays complete."""
        check(lib().rvcr_index_set_shard(self.handle, int(owned_begin), int(owned_end)))

    def get_centroids(self) -> np.ndarray:
        inf = self._info()
        out = np.empty((inf.nlist, inf.d), np.float32)
        check(lib().rvcr_index_get_centroids(self.handle, out.ctypes.data_as(C.c_void_p)))
        return out

    def list_sizes(self) -> np.ndarray:
        out = np.empty(self._info().nlist, np.int64)
        check(lib().rvcr_index_list_sizes(self.handle, out.ctypes.data_as(C.c_void_p)))
        return out

    # -- build (extract_index.py:65-69) ----------------------------------------------------
    def train(self, x) -> None:
        """index.train(x) (extract_index.py:65): faiss Clustering k-means, natively in librvcr
        (rvcr_index_train: tensor-core assignment, deterministic fp32 means, split_clusters)."""
        inf = self._info()
        x = _as_f32_2d(x, inf.d, "train")
        check(lib().rvcr_index_train(self.handle, x.ctypes.data_as(C.c_void_p), x.shape[0]))

    def add(self, x) -> None:
        """index.add(x): host (numpy) or device (torch.cuda) rows."""
        inf = self._info()
        if _is_cuda_tensor(x):
            t = _check_cuda_tensor(x, inf, "add")
            check(lib().rvcr_index_add(self.handle, C.c_void_p(t.data_ptr()), t.shape[0], 1, _stream_of(t)))
            return
        x = _as_f32_2d(x, inf.d, "add")
        check(lib().rvcr_index_add(self.handle, x.ctypes.data_as(C.c_void_p), x.shape[0], 0, None))

    # -- query (pipeline.py:499, 556) ------------------------------------------------------
    def search(self, x, k: int = 8):
        """index.search(x, k): (D (n,k) float32 squared-L2 ascending, I (n,k) int64), padded with
        (FLT_MAX, -1).  numpy in -> numpy out (H2D + D2H inside); torch.cuda in -> torch.cuda out."""
        inf = self._info()
        if _is_cuda_tensor(x):
            import torch
            t = _check_cuda_tensor(x, inf, "search")
            D = torch.empty((t.shape[0], k), dtype=torch.float32, device=t.device)
            I = torch.empty((t.shape[0], k), dtype=torch.int64, device=t.device)
            check(lib().rvcr_search(self.handle, C.c_void_p(t.data_ptr()), t.shape[0], k,
                                    C.c_void_p(D.data_ptr()), C.c_void_p(I.data_ptr()), _stream_of(t)))
            return D, I
        x = _as_f32_2d(x, inf.d, "search")
        n = x.shape[0]
        D = np.empty((n, k), np.float32)
        I = np.empty((n, k), np.int64)
        check(lib().rvcr_search_host(self.handle, x.ctypes.data_as(C.c_void_p), n, k,
                                     D.ctypes.data_as(C.c_void_p), I.ctypes.data_as(C.c_void_p)))
        return D, I

    def assign(self, x):
        """quantizer.assign: nearest-centroid list id (int32) and canonical distance per row."""
        import torch
        inf = self._info()
        t = _check_cuda_tensor(x, inf, "assign")
        lst = torch.empty(t.shape[0], dtype=torch.int32, device=t.device)
        dist = torch.empty(t.shape[0], dtype=torch.float32, device=t.device)
        check(lib().rvcr_assign(self.handle, C.c_void_p(t.data_ptr()), t.shape[0], C.c_void_p(lst.data_ptr()),
                                C.c_void_p(dist.data_ptr()), _stream_of(t)))
        return lst, dist

    lazy_reconstruct = False      # patch_pipeline sets this on the class: reconstruct_n -> LazyBigNpy

    def reconstruct_n(self, n0: int = 0, ni: int | None = None):
        """index.reconstruct_n(n0, ni) -> new (ni, d) float32 array (`big_npy`, pipeline.py:556).
        With `lazy_reconstruct` the (ntotal, d) host copy is only built if somebody actually indexes
        it (the fused path never does: the gather reads the index's own list-ordered storage)."""
        inf = self._info()
        ni = inf.ntotal - n0 if ni is None or ni < 0 else ni
        if n0 < 0 or ni < 0 or n0 + ni > inf.ntotal:
            raise RvcrError(-1, f"reconstruct_n: range [{n0}, {n0 + ni}) outside [0, {inf.ntotal})")
        if self.lazy_reconstruct:
            # one placeholder per (index, range): with the index cache on, every pipeline() call asks again
            # (pipeline.py:556) and must not trigger a second host copy if somebody does index it
            key = (int(n0), int(ni), int(inf.ntotal))
            memo = self.__dict__.setdefault("_lazy_memo", {})
            lazy = memo.get(key)
            if lazy is None:
                memo.clear()                                   # ntotal changed (index.add) or another range: drop the old copy
                lazy = memo.setdefault(key, LazyBigNpy(self, int(n0), int(ni), int(inf.d)))
            return lazy
        return self._reconstruct_now(int(n0), int(ni), int(inf.d))

    def _reconstruct_now(self, n0: int, ni: int, d: int) -> np.ndarray:
        out = np.empty((ni, d), np.float32)
        check(lib().rvcr_reconstruct_n(self.handle, n0, ni, out.ctypes.data_as(C.c_void_p), 0))
        return out

    def retrieve_blend_(self, feats, index_rate: float, want_neighbours: bool = False):
        """Fused replacement of the body of `_retrieve_speaker_embeddings` (pipeline.py:498-506):
        search(k=8) + 1/score^2 weights + gather + blend, IN PLACE on a (n, d) or (1, n, d) CUDA
        float32 tensor.  Returns feats (and (D, I) when asked)."""
        import torch
        inf = self._info()
        t = feats
        if t.dim() == 3:
            if t.shape[0] != 1:
                raise AssertionError("retrieve_blend_: batch dimension must be 1 (pipeline.py:498 uses feats[0])")
            t = t[0]
        t = _check_cuda_tensor(t, inf, "retrieve_blend_")
        n = t.shape[0]
        D = I = None
        if want_neighbours:
            D = torch.empty((n, 8), dtype=torch.float32, device=t.device)
            I = torch.empty((n, 8), dtype=torch.int64, device=t.device)
        check(lib().rvcr_retrieve_blend(self.handle, C.c_void_p(t.data_ptr()), n, float(index_rate),
                                        C.c_void_p(D.data_ptr()) if D is not None else None,
                                        C.c_void_p(I.data_ptr()) if I is not None else None, _stream_of(t)))
        return (feats, D, I) if want_neighbours else feats

    def retrieve_blend_upsample(self, feats, index_rate: float, pitchf=None, protect: float = 0.5,
                                want_neighbours: bool = False):
        """The path plus the step after it in one call (pipeline.py:455-481): retrieval + blend, 2x
        nearest upsample along time, and -- when `pitchf` (2n values, CUDA float32) is given and protect < 0.5 --
        the pitch-protect mix against the UN-blended features.  `feats` ((n, d) or (1, n, d)) is left untouched
        (it plays the reference's `feats0` clone); returns a NEW (2n, d) / (1, 2n, d) tensor."""
        import torch
        inf = self._info()
        t = feats[0] if feats.dim() == 3 else feats
        if feats.dim() == 3 and feats.shape[0] != 1:
            raise AssertionError("retrieve_blend_upsample: batch dimension must be 1")
        t = _check_cuda_tensor(t, inf, "retrieve_blend_upsample")
        n = t.shape[0]
        pf = _check_pitchf(pitchf, n, t.device) if (pitchf is not None and protect < 0.5) else None
        out = torch.empty((2 * n, inf.d), dtype=torch.float32, device=t.device)
        D = I = None
        if want_neighbours:
            D = torch.empty((n, 8), dtype=torch.float32, device=t.device)
            I = torch.empty((n, 8), dtype=torch.int64, device=t.device)
        check(lib().rvcr_retrieve_blend_upsample(
            self.handle, C.c_void_p(t.data_ptr()), n, float(index_rate),
            C.c_void_p(pf.data_ptr()) if pf is not None else None, float(protect), C.c_void_p(out.data_ptr()),
            C.c_void_p(D.data_ptr()) if D is not None else None, C.c_void_p(I.data_ptr()) if I is not None else None,
            _stream_of(t)))
        out = out.unsqueeze(0) if feats.dim() == 3 else out
        return (out, D, I) if want_neighbours else out

    def final_proj_retrieve_blend(self, x, weight, bias, index_rate: float, want_neighbours: bool = False):
        """The step before the path as the query producer (pipeline.py:451-453 + 497-507): for a v1 model,
        `model.final_proj(feats[0])` -- out = x W^T + b, nn.Linear(768, 256) -- then retrieval + blend on the projected rows,
        one call, nothing returns to the host in between.  x: (n, d_in) or (1, n, d_in) CUDA float32 encoder output;
        weight (d, d_in) / bias (d,) = `model.final_proj.weight` / `.bias` on the same device.  Returns a NEW (n, d) /
        (1, n, d) tensor holding the blended projected features."""
        import torch
        inf = self._info()
        t = x[0] if x.dim() == 3 else x
        if x.dim() == 3 and x.shape[0] != 1:
            raise AssertionError("final_proj_retrieve_blend: batch dimension must be 1")
        w, b = _check_linear(t, weight, bias, int(inf.d))
        n = t.shape[0]
        out = torch.empty((n, inf.d), dtype=torch.float32, device=t.device)
        D = I = None
        if want_neighbours:
            D = torch.empty((n, 8), dtype=torch.float32, device=t.device)
            I = torch.empty((n, 8), dtype=torch.int64, device=t.device)
        check(lib().rvcr_final_proj_retrieve_blend(
            self.handle, C.c_void_p(t.data_ptr()), n, int(t.shape[1]), C.c_void_p(w.data_ptr()),
            C.c_void_p(b.data_ptr()) if b is not None else None, C.c_void_p(out.data_ptr()), float(index_rate),
            C.c_void_p(D.data_ptr()) if D is not None else None, C.c_void_p(I.data_ptr()) if I is not None else None,
            _stream_of(t)))
        out = out.unsqueeze(0) if x.dim() == 3 else out
        return (out, D, I) if want_neighbours else out

    def retrieve_blend_host_(self, feats: np.ndarray, index_rate: float) -> np.ndarray:
        """Same, on a host (n, d) float32 C-contiguous array, in place (H2D / D2H inside)."""
        inf = self._info()
        if not (isinstance(feats, np.ndarray) and feats.dtype == np.float32 and feats.flags.c_contiguous
                and feats.ndim == 2 and feats.shape[1] == inf.d):
            raise AssertionError(f"retrieve_blend_host_: need a C-contiguous float32 (n, {inf.d}) array")
        check(lib().rvcr_retrieve_blend_host(self.handle, feats.ctypes.data_as(C.c_void_p), feats.shape[0],
                                             float(index_rate)))
        return feats

    # -- options / stats -------------------------------------------------------------------
    def set_option(self, key: str, value: int) -> None:
        check(lib().rvcr_set_option(self.handle, key.encode(), int(value)))

    def get_option(self, key: str) -> int:
        v = C.c_int64(0)
        check(lib().rvcr_get_option(self.handle, key.encode(), C.byref(v)))
        return int(v.value)

    def stats(self) -> dict:
        s = Stats()
        check(lib().rvcr_get_stats(self.handle, C.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in Stats._fields_ if k != "reserved"}

    def debug_coarse_tc(self, q, split: int = 0, cluster: int = 2, cap_pairs: int = 4096):
        """Raw output of the tcgen05 stage: (v (n, P) f32, j (n, P) i32, ||q||^2 (n,), max||c||, re-check margin (n,),
        dropped (n, P / 8) f32): P / 8 candidate lists per query, each with up to 8 centroids within the margin of the best it
        saw (free slots +inf / -1) and the smallest v it let go of."""
        import torch
        inf = self._info()
        t = _check_cuda_tensor(q, inf, "debug_coarse_tc")
        n = t.shape[0]
        v = np.empty((n, cap_pairs), np.float32)
        j = np.empty((n, cap_pairs), np.int32)
        qn = np.empty(n, np.float32)
        margin = np.empty(n, np.float32)
        dropped = np.empty((n, cap_pairs // 8), np.float32)
        pairs = C.c_int32(0)
        cmax = C.c_float(0)
        check(lib().rvcr_debug_coarse_tc(self.handle, C.c_void_p(t.data_ptr()), n, int(split), int(cluster), cap_pairs,
                                         v.ctypes.data_as(C.c_void_p), j.ctypes.data_as(C.c_void_p), C.byref(pairs),
                                         qn.ctypes.data_as(C.c_void_p), C.byref(cmax), margin.ctypes.data_as(C.c_void_p),
                                         dropped.ctypes.data_as(C.c_void_p), _stream_of(t)))
        P = pairs.value
        return (v.reshape(-1)[: n * P].reshape(n, P).copy(), j.reshape(-1)[: n * P].reshape(n, P).copy(), qn,
                float(cmax.value), margin, dropped.reshape(-1)[: n * (P // 8)].reshape(n, P // 8).copy())

    PHASES = ("coarse_prep", "coarse_main", "coarse_recheck", "group", "scan", "blend", "scan_recheck", "scan_fallback")

    def phase_ms(self) -> dict:
        """Per-phase device time (ms) of the last search-type call; needs set_option("profile", 1)."""
        arr = (C.c_float * 8)()
        check(lib().rvcr_get_phase_ms(self.handle, arr))
        return {name: float(arr[i]) for i, name in enumerate(self.PHASES) if arr[i] >= 0}

    # -- zero-copy construction ------------------------------------------------------------
    def alloc_lists(self, sizes, ntotal: int) -> None:
        sizes = np.ascontiguousarray(sizes, np.int64)
        assert sizes.shape == (self._info().nlist,)
        check(lib().rvcr_index_alloc_lists(self.handle, sizes.ctypes.data_as(C.c_void_p), int(ntotal)))

    def device_views(self):
        """torch views (no copy) of the list-ordered storage: (vecs (nstored,d) f32, ids (nstored,)
        i64, offsets (nlist+1,) i64, centroids (nlist,d) f32).  Call commit() after writing."""
        inf = self._info()
        pv, pi, po, pc = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(lib().rvcr_index_device_ptrs(self.handle, C.byref(pv), C.byref(pi), C.byref(po), C.byref(pc)))
        inf = self._info()
        dev = int(inf.device)
        return (_view(pv.value, (int(inf.nstored), int(inf.d)), "<f4", dev),
                _view(pi.value, (int(inf.nstored),), "<i8", dev),
                _view(po.value, (int(inf.nlist) + 1,), "<i8", dev),
                _view(pc.value, (int(inf.nlist), int(inf.d)), "<f4", dev) if inf.is_trained else None)

    def commit(self) -> None:
        check(lib().rvcr_index_commit(self.handle, None))


class LazyBigNpy:
    """Stand-in for `big_npy = index.reconstruct_n(0, index.ntotal)` (pipeline.py:556) that defers the
    D2H scatter until the array is really consumed.  Quacks like the (ni, d) float32 ndarray for what the
    reference does with it -- `big_npy[ix]` (pipeline.py:502), np.asarray, shape / dtype / len -- so the
    UNPATCHED `_retrieve_speaker_embeddings` body still works; the fused path ignores it entirely."""

    def __init__(self, index: "IndexIVFFlat", n0: int, ni: int, d: int):
        # weak: the index memoises its placeholder (reconstruct_n), and a strong back reference would be a cycle that
        # keeps up to 100 GB of HBM alive until the cyclic GC runs instead of until the last user drops the index
        self._index_ref, self._n0, self._ni, self._d = weakref.ref(index), n0, ni, d
        self._arr = None
        self._lock = threading.Lock()     # one materialisation even when worker threads index it at the same time

    shape = property(lambda self: (self._ni, self._d))
    dtype = np.dtype(np.float32)
    ndim = 2

    @property
    def materialized(self) -> bool:
        return self._arr is not None

    def _get(self) -> np.ndarray:
        if self._arr is None:
            with self._lock:
                if self._arr is None:
                    index = self._index_ref()
                    if index is None:
                        raise RvcrError(-6, "big_npy placeholder used after its index was freed")
                    self._arr = index._reconstruct_now(self._n0, self._ni, self._d)
        return self._arr

    def __array__(self, dtype=None, copy=None):
        a = self._get()
        return a if dtype is None or np.dtype(dtype) == a.dtype else a.astype(dtype)

    def __getitem__(self, key):
        return self._get()[key]

    def __len__(self) -> int:
        return self._ni


class _CudaArray:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"data": (ptr, False), "shape": tuple(shape), "typestr": typestr,
                                         "version": 2, "strides": None}


def _view(ptr, shape, typestr, device):
    import torch
    if not ptr or int(np.prod(shape)) == 0:
        dt = torch.float32 if typestr == "<f4" else torch.int64
        return torch.empty(shape, dtype=dt, device=f"cuda:{device}")
    with torch.cuda.device(device):
        return torch.as_tensor(_CudaArray(ptr, shape, typestr), device=f"cuda:{device}")


def _is_cuda_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


def _check_cuda_tensor(t, inf: Info, what: str):
    import torch
    if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.shape[1] == inf.d and t.is_contiguous()):
        raise AssertionError(f"{what}: need a contiguous float32 CUDA tensor of shape (n, {inf.d}), "
                             f"got {tuple(t.shape)} {t.dtype} {t.device}")
    if t.device.index != inf.device:
        raise AssertionError(f"{what}: tensor is on {t.device}, index lives on cuda:{inf.device}")
    return t


def _check_pitchf(pitchf, n: int, device):
    import torch
    pf = pitchf.reshape(-1)
    if not (pf.is_cuda and pf.dtype == torch.float32 and pf.is_contiguous() and pf.device == device):
        raise AssertionError("pitchf: need a contiguous float32 CUDA tensor on the features' device")
    if pf.shape[0] != 2 * n:
        raise AssertionError(f"pitchf: expected {2 * n} values (2 per feature frame, pipeline.py:466-474), got {pf.shape[0]}")
    return pf


def upsample_protect(feats, feats0=None, pitchf=None, protect: float = 0.5):
    """pipeline.py:463-481 alone (for calls that skipped retrieval): `feats` / `feats0` (n, d) or (1, n, d) CUDA
    float32 -> new (2n, d) / (1, 2n, d) tensor.  The protect mix is applied iff feats0 and pitchf are given and
    protect < 0.5 (the reference's `if pitch_guidance` / `if protect < 0.5`)."""
    import torch
    t = feats[0] if feats.dim() == 3 else feats
    if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.is_contiguous()):
        raise AssertionError("upsample_protect: need a contiguous float32 CUDA tensor (n, d) or (1, n, d)")
    n, d = t.shape
    mix = feats0 is not None and pitchf is not None and protect < 0.5
    t0 = pf = None
    if mix:
        t0 = feats0[0] if feats0.dim() == 3 else feats0
        if not (t0.is_cuda and t0.dtype == torch.float32 and t0.shape == t.shape and t0.is_contiguous()
                and t0.device == t.device):
            raise AssertionError("upsample_protect: feats0 must match feats (shape, dtype, device)")
        pf = _check_pitchf(pitchf, n, t.device)
    out = torch.empty((2 * n, d), dtype=torch.float32, device=t.device)
    check(lib().rvcr_upsample_protect(C.c_void_p(t.data_ptr()), C.c_void_p(t0.data_ptr()) if mix else None,
                                      C.c_void_p(pf.data_ptr()) if mix else None, n, d, float(protect),
                                      C.c_void_p(out.data_ptr()), int(t.device.index), _stream_of(t)))
    return out.unsqueeze(0) if feats.dim() == 3 else out


def _check_linear(t, weight, bias, d_out):
    import torch
    if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.is_contiguous()):
        raise AssertionError("final_proj: need a contiguous float32 CUDA tensor (n, d_in) or (1, n, d_in)")
    w = weight.detach()
    if not (w.is_cuda and w.device == t.device and w.dtype == torch.float32 and w.is_contiguous()
            and tuple(w.shape) == (d_out, t.shape[1])):
        raise AssertionError(f"final_proj: weight must be a contiguous float32 ({d_out}, {t.shape[1]}) tensor on {t.device}")
    b = None
    if bias is not None:
        b = bias.detach()
        if not (b.is_cuda and b.device == t.device and b.dtype == torch.float32 and b.is_contiguous() and tuple(b.shape) == (d_out,)):
            raise AssertionError(f"final_proj: bias must be a contiguous float32 ({d_out},) tensor on {t.device}")
    return w, b


def final_proj(x, weight, bias=None):
    """`model.final_proj(feats[0])` (pipeline.py:451-453) alone: x (n, d_in) or (1, n, d_in) CUDA float32, weight
    (d_out, d_in), bias (d_out,) -> new (n, d_out) / (1, n, d_out) tensor.  fp32 FFMA, ascending k (csrc/final_proj.cu)."""
    import torch
    t = x[0] if x.dim() == 3 else x
    w, b = _check_linear(t, weight, bias, int(weight.shape[0]))
    out = torch.empty((t.shape[0], w.shape[0]), dtype=torch.float32, device=t.device)
    check(lib().rvcr_final_proj(C.c_void_p(t.data_ptr()), t.shape[0], int(t.shape[1]), C.c_void_p(w.data_ptr()),
                                C.c_void_p(b.data_ptr()) if b is not None else None, int(w.shape[0]),
                                C.c_void_p(out.data_ptr()), int(t.device.index), _stream_of(t)))
    return out.unsqueeze(0) if x.dim() == 3 else out


class MiniBatchKMeans:
    """Drop-in for the one use the reference makes of sklearn.cluster.MiniBatchKMeans (rvc/train/process/extract_index.py:
    43-56, the default "Auto" build path above 200 000 rows):

        big_npy = MiniBatchKMeans(n_clusters=10000, verbose=True, batch_size=256 * cpu_count(), compute_labels=False,
                                  init="random").fit(big_npy).cluster_centers_

    Same algorithm (scikit-learn's mini-batch k-means: 3 random inits scored on a validation subset, per-centre learning
    rates, random reassignment of starved centres, early stop on the smoothed inertia), with the nearest-centre step of
    every mini-batch on the GPU coarse kernels (csrc/minibatch_kmeans.cu).  Like the reference's call it is unseeded unless
    `random_state` is given; runs agree in distribution, not bit for bit.  Only what that call uses is implemented:
    init="random", compute_labels=False, dense float32 rows."""

    def __init__(self, n_clusters=8, *, init="random", max_iter=100, batch_size=1024, verbose=0, compute_labels=False,
                 random_state=None, tol=0.0, max_no_improvement=10, init_size=None, n_init="auto", reassignment_ratio=0.01,
                 device: int = 0):
        if init != "random":
            raise NotImplementedError('only init="random" (the reference call, extract_index.py:52) is implemented')
        if compute_labels:
            raise NotImplementedError("compute_labels=True is not implemented (the reference passes False)")
        if tol != 0.0 or init_size is not None:
            raise NotImplementedError("tol / init_size: only the estimator's defaults are implemented")
        self.n_clusters, self.max_iter, self.batch_size, self.verbose = int(n_clusters), int(max_iter), int(batch_size), verbose
        self.random_state, self.max_no_improvement = random_state, max_no_improvement
        self.n_init = 3 if n_init == "auto" else int(n_init)
        self.reassignment_ratio, self.device = float(reassignment_ratio), int(device)
        self.cluster_centers_ = None

    def fit(self, X, y=None, sample_weight=None):
        from ._ffi import MbkStats
        if sample_weight is not None:
            raise NotImplementedError("sample_weight is not implemented")
        X = np.ascontiguousarray(X, dtype=np.float32)
        if X.ndim != 2:
            raise ValueError("Expected 2D array")
        n, d = X.shape
        seed = 0 if self.random_state is None else (int(self.random_state) % (2 ** 63)) + 1
        out = np.empty((self.n_clusters, d), np.float32)
        st = MbkStats()
        check(lib().rvcr_minibatch_kmeans(X.ctypes.data_as(C.c_void_p), n, d, self.n_clusters, self.batch_size, self.max_iter,
                                          0 if self.max_no_improvement is None else int(self.max_no_improvement),
                                          self.reassignment_ratio, self.n_init, seed, self.device, 1 if self.verbose else 0,
                                          out.ctypes.data_as(C.c_void_p), C.byref(st)))
        self.cluster_centers_ = out
        self.n_steps_ = int(st.n_steps)
        self.n_iter_ = int(np.ceil(st.n_steps * st.batch_size / n))
        self.n_features_in_ = d
        self.stats_ = {k: getattr(st, k) for k, _ in MbkStats._fields_ if k != "reserved"}
        return self


def _stream_of(t):
    import torch
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


# ---- module-level faiss functions ------------------------------------------------------------

def _open_index(path: str, dev: int, shard: tuple[int, int] | None) -> IndexIVFFlat:
    """Stream the `.index` file into list-ordered HBM (rvcr_index_read / rvcr_index_read_shard)."""
    h = C.c_void_p()
    if shard is None:
        check(lib().rvcr_index_read(str(path).encode(), dev, C.byref(h)))
    else:
        check(lib().rvcr_index_read_shard(str(path).encode(), dev, int(shard[0]), int(shard[1]), C.byref(h)))
    return IndexIVFFlat(h.value)


class IndexCache:
    """Device-resident index cache.

    The reference calls `faiss.read_index(file_index)` + `index.reconstruct_n(0, ntotal)` at the top of
    EVERY `Pipeline.pipeline()` call (pipeline.py:553-561), i.e. once per converted file / chunk
    (infer.py:291-311): the whole file is re-read and a second host copy rebuilt each time.  With the
    cache enabled `read_index` returns the SAME device-resident object while the file is unchanged --
    keyed by (realpath, mtime_ns, size, device, shard) -- so only the first call touches the disk.
    LRU over `max_entries` handles and `max_bytes` of list storage; an evicted handle is freed when its
    last user drops it (plain Python reference counting, no handle is ever closed under a user).

    Off by default: faiss's `read_index` returns independent objects, and a caller that `add`s to a
    loaded index must not mutate a shared one.  `patch_pipeline` turns it on (the inference path is
    read-only)."""

    def __init__(self, max_entries: int = 4, max_bytes: int = 120 << 30):
        self.max_entries = int(max_entries)
        self.max_bytes = int(max_bytes)
        self.enabled = False
        self.hits = 0
        self.misses = 0
        self._lock = threading.Lock()
        self._entries: "OrderedDict[tuple, tuple[IndexIVFFlat, int]]" = OrderedDict()
        self._loading: "dict[tuple, threading.Event]" = {}    # keys being read right now (one reader per key)

    @staticmethod
    def key(path: str, dev: int, shard) -> tuple:
        rp = os.path.realpath(path)
        st = os.stat(rp)                                    # FileNotFoundError propagates like faiss's RuntimeError
        return (rp, st.st_mtime_ns, st.st_size, int(dev), tuple(shard) if shard is not None else None)

    def get(self, path: str, dev: int, shard, opener) -> IndexIVFFlat:
        k = self.key(path, dev, shard)
        while True:
            with self._lock:
                hit = self._entries.get(k)
                if hit is not None:
                    self._entries.move_to_end(k)
                    self.hits += 1
                    return hit[0]
                pending = self._loading.get(k)
                if pending is None:
                    # a rewritten file (same path, new mtime / size) supersedes its stale entry
                    for old in [e for e in self._entries if e[0] == k[0] and e[3:] == k[3:]]:
                        del self._entries[old]
                    pending = self._loading[k] = threading.Event()
                    break
            # another thread (Gradio worker, core.py:38-42 shares one converter) is reading this very file: wait for it
            # instead of loading a second copy of up to 100 GB; if its read failed, take over
            pending.wait()
        try:
            index = opener(path, dev, shard)               # outside the lock: may take seconds for a 100 GB file
        except BaseException:
            with self._lock:
                del self._loading[k]
            pending.set()
            raise
        nbytes = int(k[2])
        with self._lock:
            self.misses += 1
            self._entries[k] = (index, nbytes)
            self._entries.move_to_end(k)
            while len(self._entries) > max(self.max_entries, 1) or (
                    len(self._entries) > 1 and sum(b for _, b in self._entries.values()) > self.max_bytes):
                self._entries.popitem(last=False)
            del self._loading[k]
        pending.set()
        return index

    def clear(self) -> None:
        with self._lock:
            self._entries.clear()

    def __len__(self) -> int:
        return len(self._entries)


index_cache = IndexCache()


def enable_index_cache(enabled: bool = True, max_entries: int | None = None, max_bytes: int | None = None) -> IndexCache:
    index_cache.enabled = bool(enabled)
    if max_entries is not None:
        index_cache.max_entries = int(max_entries)
    if max_bytes is not None:
        index_cache.max_bytes = int(max_bytes)
    if not enabled:
        index_cache.clear()
    return index_cache


def read_index(path: str, device: int | None = None, shard: tuple[int, int] | None = None,
               cache: bool | None = None) -> IndexIVFFlat:
    """faiss.read_index(path) (pipeline.py:555).  `shard=(rank, world)` keeps one shard's lists.
    `cache` (default: the module switch set by enable_index_cache / patch_pipeline) returns the
    already-resident index while the file is unchanged."""
    dev = _default_device() if device is None else int(device)
    use_cache = index_cache.enabled if cache is None else bool(cache)
    if use_cache:
        try:
            return index_cache.get(str(path), dev, shard, _open_index)
        except FileNotFoundError:
            pass                                            # fall through: the C reader reports RVCR_ERR_IO
    return _open_index(path, dev, shard)


def probe_index(path: str, list_sizes: bool = False) -> dict:
    """Describe an `.index` file on the host, no GPU involved (rvcr_index_probe): d, ntotal, nlist, nprobe, layout
    ('IwFl' or the legacy 'IvFL' / 'IvFl'), list-size statistics, `searchable`.  Raises RvcrError exactly where
    `read_index` would refuse the file.  With `list_sizes=True` the dict also holds the (nlist,) int64 vector counts."""
    from ._ffi import FileInfo
    inf = FileInfo()
    p = os.fsencode(path)
    check(lib().rvcr_index_probe(p, C.byref(inf), None, 0))
    out = {k: getattr(inf, k) for k, _ in FileInfo._fields_ if not k.startswith("reserved")}
    out["fourcc"] = inf.fourcc.decode("ascii", "replace")
    out["quantizer_fourcc"] = inf.quantizer_fourcc.decode("ascii", "replace")
    if list_sizes:
        sizes = np.zeros(inf.nlist, np.int64)
        check(lib().rvcr_index_probe(p, None, C.c_void_p(sizes.ctypes.data), sizes.size))
        out["list_sizes"] = sizes
    return out


def shard_range(list_sizes, rank: int, world: int) -> tuple[int, int]:
    """List range [lo, hi) that `read_index(path, shard=(rank, world))` loads: contiguous, balanced by vector count."""
    sizes = np.ascontiguousarray(list_sizes, np.int64)
    lo, hi = C.c_int64(0), C.c_int64(0)
    check(lib().rvcr_shard_range(C.c_void_p(sizes.ctypes.data), sizes.size, rank, world, C.byref(lo), C.byref(hi)))
    return lo.value, hi.value


def write_index(index: IndexIVFFlat, path: str) -> None:
    """faiss.write_index(index, path) (extract_index.py:71)."""
    check(lib().rvcr_index_write(index.handle, str(path).encode()))


def index_factory(d: int, description: str, metric: int = METRIC_L2, device: int | None = None) -> IndexIVFFlat:
    """faiss.index_factory(d, "IVF{nlist},Flat") (extract_index.py:62); other strings are out of scope."""
    m = re.fullmatch(r"IVF(\d+),Flat", description.strip())
    if not m:
        raise RvcrError(-5, f"index_factory: only 'IVF<nlist>,Flat' is supported, got {description!r}")
    if metric != METRIC_L2:
        raise RvcrError(-5, "index_factory: only METRIC_L2 is supported")
    h = C.c_void_p()
    dev = _default_device() if device is None else int(device)
    check(lib().rvcr_index_create(int(d), int(m.group(1)), dev, C.byref(h)))
    return IndexIVFFlat(h.value)


def extract_index_ivf(index: IndexIVFFlat) -> IndexIVFFlat:
    """faiss.extract_index_ivf(index) (extract_index.py:63): the IVF index is the object itself."""
    return index
