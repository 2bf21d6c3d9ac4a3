"""GPU: the index-build call sequence of rvc/train/process/extract_index.py:59-71, statement for statement,
on the faiss-shaped objects -- factory, nprobe, train (k-means whose assignment runs on the GPU), add in
8192-row batches, write_index -- then the inference-side load (pipeline.py:555-556) and search.

k-means (rvcr_index_train, csrc/kmeans.cu) is compared with the oracle's restatement of faiss Clustering::train
(PARITY UNPINNED with respect to the faiss wheel, see oracle.kmeans_train).  The native path is deterministic --
exact assignments, means accumulated in fp32 in point order like faiss's compute_centroids -- so the centroids
after all 10 iterations equal the oracle's BIT FOR BIT, as does everything downstream of them (list assignment,
list order, file bytes, search)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def clustered(rng, n, d, nb, sigma=0.08):
    centres = (rng.standard_normal((nb, d)) * 0.3).astype(np.float32)
    return (centres[rng.integers(0, nb, n)] + rng.standard_normal((n, d)).astype(np.float32) * sigma).astype(np.float32)


def test_extract_index_sequence(rvcr, oracle, tmp_path):
    rng = np.random.default_rng(99)
    big_npy = clustered(rng, 20_000, 768, 400)
    faiss = rvcr                                                    # `import rvcr as faiss`
    # ---- extract_index.py:59-71 ----
    n_ivf = min(int(16 * np.sqrt(big_npy.shape[0])), big_npy.shape[0] // 39)
    assert n_ivf == 512
    index_added = faiss.index_factory(768, f"IVF{n_ivf},Flat")
    index_ivf_added = faiss.extract_index_ivf(index_added)
    index_ivf_added.nprobe = 1
    index_added.train(big_npy)
    batch_size_add = 8192
    for i in range(0, big_npy.shape[0], batch_size_add):
        index_added.add(big_npy[i: i + batch_size_add])
    path = str(tmp_path / "model.index")
    faiss.write_index(index_added, path)
    # ---- centroids vs the oracle's Clustering::train ----
    assert index_added.is_trained and index_added.ntotal == len(big_npy)
    cent = index_added.get_centroids()
    want, a_last, obj = oracle.kmeans_train(big_npy, n_ivf)
    assert all(b <= a * (1 + 1e-6) for a, b in zip(obj, obj[1:]))   # Lloyd: the objective never rises
    assert obj[-1] < 0.5 * obj[0]                                   # and it did converge from the random init
    assert np.array_equal(cent, want)                               # whole 10-iteration trajectory, bit for bit
    # niter = 0 is the init alone: x[rand_perm(n, seed + 1)[:k]]; niter = 1 is one Lloyd step from it
    from importlib import import_module
    km = import_module("codename-rvc-fork-3_b200.kmeans")
    probe = faiss.index_factory(768, f"IVF{n_ivf},Flat")
    c0 = km.train_ivf_centroids(probe, big_npy, n_ivf, niter=0)
    assert np.array_equal(c0, big_npy[oracle.rand_perm(len(big_npy), 1235)[:n_ivf]])
    c1 = km.train_ivf_centroids(probe, big_npy, n_ivf, niter=1)
    w1, _, _ = oracle.kmeans_train(big_npy, n_ivf, niter=1)
    assert np.array_equal(c1, w1)
    # ---- everything after the centroids is exact: lists, order, file bytes ----
    a_x, _ = oracle.coarse_assign(big_npy, cent)
    off, vecs, ids = oracle.build_lists(big_npy, a_x, n_ivf)
    assert np.array_equal(index_added.list_sizes(), np.diff(off))
    opath = str(tmp_path / "oracle.index")
    oracle.write_index(oracle.OracleIndex(768, cent, off, vecs, ids, len(big_npy)).to_data(), opath)
    assert open(path, "rb").read() == open(opath, "rb").read()
    # ---- pipeline.py:555-556,499 on the written file ----
    index = faiss.read_index(path)
    assert index.nprobe == 1 and index.ntotal == len(big_npy)
    assert np.array_equal(index.reconstruct_n(0, index.ntotal), big_npy)
    q = big_npy[rng.integers(0, len(big_npy), 300)] + rng.standard_normal((300, 768)).astype(np.float32) * 0.03
    score, ix = index.search(q.astype(np.float32), k=8)
    Do, Io = oracle.OracleIndex.from_data(oracle.read_index(path)).search(q.astype(np.float32), k=8)
    assert np.array_equal(ix, Io) and np.array_equal(score, Do)


def test_train_subsamples_and_reseeds_empty_clusters(rvcr, oracle):
    """n > 256 * nlist triggers faiss's rand_perm subsample; duplicated points leave clusters empty so
    split_clusters runs.  Too few points raise like faiss."""
    rng = np.random.default_rng(3)
    k, d = 8, 128                                        # d must be a multiple of 128 (PD_KC tile of the exact kernels)
    base = clustered(rng, 300, d, 3, sigma=0.02)
    x = np.concatenate([base, np.repeat(base[:1], 2400, axis=0)]).astype(np.float32)     # 2700 > 256 * 8
    from importlib import import_module
    km = import_module("codename-rvc-fork-3_b200.kmeans")
    index = rvcr.index_factory(d, f"IVF{k},Flat")
    for niter in (0, 1, 2, 10):                          # 1 = mean update + 7 split_clusters re-seeds
        cent = km.train_ivf_centroids(index, x, k, niter=niter)
        want, _, _ = oracle.kmeans_train(x, k, niter=niter)
        assert np.isfinite(cent).all()
        assert np.array_equal(cent, want), niter
    index.train(x)
    assert index.is_trained and np.array_equal(index.get_centroids(), want)
    index.add(x[:100])
    with pytest.raises(RuntimeError):
        index.train(x)                                    # training an index that already holds vectors
    with pytest.raises(RuntimeError):
        rvcr.index_factory(d, "IVF64,Flat").train(x[:10])
