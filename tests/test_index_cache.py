"""CPU: host logic of the device-resident index cache and the lazy big_npy placeholder.

The reference re-reads the `.index` file and rebuilds big_npy on every Pipeline.pipeline() call
(rvc/infer/pipeline.py:553-561); the cache must hand back the same resident object while the file is
unchanged, notice a rewritten file, and bound what it keeps.  The opener is mocked: no GPU here."""
import importlib
import os
import time

import numpy as np
import pytest

fc = importlib.import_module("codename-rvc-fork-3_b200.faiss_compat")
retrieve = importlib.import_module("codename-rvc-fork-3_b200.retrieve")


class FakeIndex:
    opened = 0

    def __init__(self, path, dev, shard):
        FakeIndex.opened += 1
        self.path, self.dev, self.shard = path, dev, shard


def opener(path, dev, shard):
    return FakeIndex(path, dev, shard)


def touch(path, payload=b"x" * 64, bump_ns=0):
    with open(path, "wb") as f:
        f.write(payload)
    if bump_ns:
        st = os.stat(path)
        os.utime(path, ns=(st.st_atime_ns, st.st_mtime_ns + bump_ns))


def test_same_file_is_opened_once(tmp_path):
    c = fc.IndexCache(max_entries=4)
    p = str(tmp_path / "a.index")
    touch(p)
    FakeIndex.opened = 0
    a = c.get(p, 0, None, opener)
    b = c.get(p, 0, None, opener)
    assert a is b and FakeIndex.opened == 1 and (c.hits, c.misses) == (1, 1)
    # a symlink / relative spelling of the same file is the same entry (keyed by realpath)
    link = str(tmp_path / "link.index")
    os.symlink(p, link)
    assert c.get(link, 0, None, opener) is a
    # another device or another shard is another resident copy
    assert c.get(p, 1, None, opener) is not a
    assert c.get(p, 0, (0, 2), opener) is not a
    assert FakeIndex.opened == 3


def test_rewritten_file_is_reloaded_and_stale_entry_dropped(tmp_path):
    c = fc.IndexCache(max_entries=4)
    p = str(tmp_path / "a.index")
    touch(p)
    a = c.get(p, 0, None, opener)
    touch(p, b"y" * 64, bump_ns=5_000_000)            # same size, newer mtime
    b = c.get(p, 0, None, opener)
    assert b is not a and len(c) == 1
    touch(p, b"z" * 65, bump_ns=0)                     # size change alone is enough
    assert c.get(p, 0, None, opener) is not b and len(c) == 1


def test_lru_bounds(tmp_path):
    c = fc.IndexCache(max_entries=2)
    ps = []
    for i in range(3):
        p = str(tmp_path / f"{i}.index")
        touch(p)
        ps.append(p)
    a0 = c.get(ps[0], 0, None, opener)
    c.get(ps[1], 0, None, opener)
    assert c.get(ps[0], 0, None, opener) is a0          # refresh 0: 1 becomes the LRU entry
    c.get(ps[2], 0, None, opener)
    assert len(c) == 2
    assert c.get(ps[0], 0, None, opener) is a0          # still resident
    n = FakeIndex.opened
    c.get(ps[1], 0, None, opener)                       # was evicted: opened again
    assert FakeIndex.opened == n + 1
    # byte budget: two 64-byte files do not fit 100 bytes, the newest always stays
    c2 = fc.IndexCache(max_entries=8, max_bytes=100)
    c2.get(ps[0], 0, None, opener)
    c2.get(ps[1], 0, None, opener)
    assert len(c2) == 1


def test_read_index_switch_and_missing_file(tmp_path, monkeypatch):
    calls = []
    monkeypatch.setattr(fc, "_open_index", lambda path, dev, shard: calls.append(path) or FakeIndex(path, dev, shard))
    p = str(tmp_path / "a.index")
    touch(p)
    try:
        fc.enable_index_cache(False)
        assert fc.read_index(p, device=0) is not fc.read_index(p, device=0)          # faiss semantics by default
        assert fc.read_index(p, device=0, cache=True) is fc.read_index(p, device=0, cache=True)
        fc.enable_index_cache(True, max_entries=3)
        assert fc.read_index(p, device=0) is fc.read_index(p, device=0)
        assert fc.read_index(p, device=0, cache=False) is not fc.read_index(p, device=0)
        # a missing file must reach the C reader (which raises RVCR_ERR_IO), not die in os.stat
        n = len(calls)
        fc.read_index(str(tmp_path / "missing.index"), device=0)
        assert len(calls) == n + 1
    finally:
        fc.enable_index_cache(False)
        assert len(fc.index_cache) == 0


class CountingIndex:
    """Only what LazyBigNpy touches."""

    def __init__(self, arr):
        self.arr, self.calls = arr, 0

    def _reconstruct_now(self, n0, ni, d):
        self.calls += 1
        return self.arr[n0:n0 + ni].copy()


def test_lazy_big_npy_defers_and_quacks_like_the_array():
    rng = np.random.default_rng(0)
    big = rng.standard_normal((50, 16)).astype(np.float32)
    idx = CountingIndex(big)
    lazy = fc.LazyBigNpy(idx, 0, 50, 16)
    assert lazy.shape == (50, 16) and lazy.dtype == np.float32 and len(lazy) == 50 and lazy.ndim == 2
    assert idx.calls == 0 and not lazy.materialized              # nothing copied yet
    ix = rng.integers(-1, 50, size=(7, 8))
    w = rng.random((7, 8)).astype(np.float32)
    got = np.sum(lazy[ix] * np.expand_dims(w, axis=2), axis=1)     # pipeline.py:502 verbatim on the placeholder
    want = np.sum(big[ix] * np.expand_dims(w, axis=2), axis=1)
    assert np.array_equal(got, want) and idx.calls == 1
    assert np.array_equal(np.asarray(lazy), big) and idx.calls == 1   # materialised once


def test_patch_pipeline_installs_everything():
    class Pipeline:
        def _retrieve_speaker_embeddings(self, feats, index, big_npy, index_rate):
            raise AssertionError("reference body should have been replaced")

    class Mod:
        pass

    mod = Mod()
    mod.Pipeline = Pipeline
    mod.faiss = None
    try:
        retrieve.patch_pipeline(mod)
        assert mod.faiss.read_index is fc.read_index
        assert Pipeline._retrieve_speaker_embeddings is retrieve._method
        assert fc.index_cache.enabled and fc.IndexIVFFlat.lazy_reconstruct
        with pytest.raises(TypeError):                                  # a foreign index object is refused loudly
            Pipeline()._retrieve_speaker_embeddings(None, object(), None, 0.75)
    finally:
        fc.enable_index_cache(False)
        fc.IndexIVFFlat.lazy_reconstruct = False


def test_concurrent_readers_of_one_file_load_it_once(tmp_path):
    """Gradio worker threads share one converter (core.py:38-42): several `pipeline()` calls hitting the same cold file at
    once must produce ONE read (a second copy of a 100 GB index would not fit), and a failed read must not wedge the
    waiters."""
    import threading
    import time
    faiss_compat = importlib.import_module("codename-rvc-fork-3_b200.faiss_compat")
    p = str(tmp_path / "a.index")
    open(p, "wb").write(b"x" * 10)
    c = faiss_compat.IndexCache()
    calls = []

    def slow_opener(path, dev, shard):
        calls.append(threading.get_ident())
        time.sleep(0.2)
        return FakeIndex(path, dev, shard)

    got = [None] * 6

    def worker(i):
        got[i] = c.get(p, 0, None, slow_opener)

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(6)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert len(calls) == 1 and all(g is got[0] for g in got)
    assert c.misses == 1 and c.hits == 5 and not c._loading

    # the first reader fails: exactly one waiter takes over, the others get its result
    q = str(tmp_path / "b.index")
    open(q, "wb").write(b"y" * 10)
    attempts = []
    lock = threading.Lock()

    def flaky_opener(path, dev, shard):
        with lock:
            attempts.append(1)
            first = len(attempts) == 1
        time.sleep(0.1)
        if first:
            raise RuntimeError("disk hiccup")
        return FakeIndex(path, dev, shard)

    res, errs = [], []

    def worker2():
        try:
            res.append(c.get(q, 0, None, flaky_opener))
        except RuntimeError as e:
            errs.append(e)

    ts = [threading.Thread(target=worker2) for _ in range(5)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert len(errs) == 1 and len(res) == 4 and all(r is res[0] for r in res)
    assert len(attempts) == 2 and not c._loading


def test_lazy_big_npy_materialises_once_under_threads_and_is_shared_per_index():
    import threading
    import time

    class SlowIdx:
        calls = 0

        def _reconstruct_now(self, n0, ni, d):
            SlowIdx.calls += 1
            time.sleep(0.1)
            return np.arange(ni * d, dtype=np.float32).reshape(ni, d)

    slow = SlowIdx()
    lazy = fc.LazyBigNpy(slow, 0, 20, 8)                         # the placeholder holds its index weakly
    outs = []
    ts = [threading.Thread(target=lambda: outs.append(lazy[3])) for _ in range(5)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert SlowIdx.calls == 1 and all(np.array_equal(o, outs[0]) for o in outs)

    # reconstruct_n on the same index and range hands back the same placeholder (pipeline() asks on every call)
    class Inf:
        ntotal, d = 20, 8

    class Idx(fc.IndexIVFFlat):
        lazy_reconstruct = True

        def __init__(self):
            pass

        def _info(self):
            return Inf

        def __del__(self):
            pass

    idx = Idx()
    a, b = idx.reconstruct_n(0, 20), idx.reconstruct_n(0, Inf.ntotal)
    assert a is b and isinstance(a, fc.LazyBigNpy)
    Inf.ntotal = 30                                             # index.add happened: a new placeholder, the old copy is dropped
    c = idx.reconstruct_n(0, 30)
    assert c is not a and c.shape == (30, 8) and len(idx._lazy_memo) == 1
    # no reference cycle: dropping the index frees it at once (its HBM with it); an unmaterialised orphan says so
    import gc
    import weakref
    gc.disable()
    try:
        w = weakref.ref(idx)
        del idx, a, b
        assert w() is None
        with pytest.raises(fc.RvcrError):
            c[0]
    finally:
        gc.enable()
