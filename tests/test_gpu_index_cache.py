"""GPU: the device-resident index cache and the lazy big_npy.  The reference re-reads the `.index` file and
rebuilds a second host copy at the top of every `Pipeline.pipeline()` call (rvc/infer/pipeline.py:553-561, once per
converted file / chunk: rvc/infer/infer.py:291-311); with the cache the same file is read once and stays in HBM."""
import gc
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_cache_hits_reload_and_eviction_on_the_device(rvcr, oracle, golden, tmp_path):
    import torch
    g = golden("retrieve_v2_768")
    oidx = oracle.OracleIndex.from_vectors(g["x"], g["centroids"])
    path = str(tmp_path / "a.index")
    oracle.write_index(oidx.to_data(), path)
    cache = rvcr.enable_index_cache(True, max_entries=2)
    try:
        cache.hits = cache.misses = 0
        a1 = rvcr.read_index(path)
        vec_ptr = a1.device_views()[0].data_ptr()
        a2 = rvcr.read_index(path)
        assert a2 is a1 and (cache.hits, cache.misses) == (1, 0 + 1) and a2.device_views()[0].data_ptr() == vec_ptr
        D1, I1 = a1.search(g["q"], k=8)
        # the file is not read again: make it unreadable -- a hit never opens it (the key is stat() only)
        os.chmod(path, 0o000)
        try:
            a3 = rvcr.read_index(path)
            assert a3 is a1
        finally:
            os.chmod(path, 0o644)
        D3, I3 = a3.search(g["q"], k=8)
        assert np.array_equal(I1, I3) and np.array_equal(D1, D3) and np.array_equal(I1, g["ix"])
        # cache=False gives an independent object (what faiss.read_index does)
        b = rvcr.read_index(path, cache=False)
        assert b is not a1 and b.device_views()[0].data_ptr() != vec_ptr
        del b
        # a rewritten file (other rows, same path) supersedes the resident one
        x2 = g["x"][::-1].copy()
        o2 = oracle.OracleIndex.from_vectors(x2, g["centroids"])
        oracle.write_index(o2.to_data(), path)
        os.utime(path, ns=(os.stat(path).st_atime_ns, os.stat(path).st_mtime_ns + 5_000_000))
        c = rvcr.read_index(path)
        assert c is not a1 and len(cache) == 1
        assert np.array_equal(c.search(g["q"], k=8)[1], o2.search(g["q"], k=8)[1])
        # LRU eviction frees HBM once the last user drops the handle
        big_rows = np.random.default_rng(0).standard_normal((400_000, 768)).astype(np.float32)     # 1.2 GB of list storage
        cent = big_rows[:64].copy()
        paths = []
        for i in range(3):
            idx = rvcr.index_factory(768, "IVF64,Flat")
            idx.set_centroids(cent)
            idx.add(big_rows)
            p = str(tmp_path / f"big{i}.index")
            rvcr.write_index(idx, p)
            paths.append(p)
            del idx
        del a1, a2, a3, c
        cache.clear()
        gc.collect()
        torch.cuda.synchronize()
        free0 = torch.cuda.mem_get_info()[0]
        h0 = rvcr.read_index(paths[0])
        h1 = rvcr.read_index(paths[1])
        used2 = free0 - torch.cuda.mem_get_info()[0]
        assert used2 >= 2 * 1.2e9
        del h0, h1
        h2 = rvcr.read_index(paths[2])                       # third entry: the least recently used one is evicted and freed
        gc.collect()
        torch.cuda.synchronize()
        assert len(cache) == 2
        used3 = free0 - torch.cuda.mem_get_info()[0]
        assert used3 < used2 + 0.4e9, (used2, used3)         # not three resident copies
        del h2
    finally:
        rvcr.enable_index_cache(False)
        gc.collect()


def test_lazy_big_npy_is_only_built_when_indexed(rvcr, oracle, golden, tmp_path):
    g = golden("retrieve_v1_256")
    oidx = oracle.OracleIndex.from_vectors(g["x"], g["centroids"])
    path = str(tmp_path / "v1.index")
    oracle.write_index(oidx.to_data(), path)
    index = rvcr.read_index(path)
    rvcr.IndexIVFFlat.lazy_reconstruct = True
    try:
        big = index.reconstruct_n(0, index.ntotal)            # pipeline.py:556
        assert isinstance(big, rvcr.LazyBigNpy) and big.shape == g["x"].shape and not big.materialized
        assert np.array_equal(big[g["ix"][0]], g["x"][g["ix"][0]]) and big.materialized       # fancy indexing like pipeline.py:502
        assert np.array_equal(np.asarray(big), g["x"])
    finally:
        rvcr.IndexIVFFlat.lazy_reconstruct = False
