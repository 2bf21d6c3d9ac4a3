"""index.train(x) for IndexIVFFlat: faiss `Clustering::train` as a host-side Lloyd loop whose
assignment step -- the 10 x (n x nlist x d) contractions that are all of the cost -- runs on the GPU
through rvcr_assign (tcgen05 coarse quantizer + exact re-check).

Follows faiss 1.7.3 defaults for Level1Quantizer: niter = 10, nredo = 1, seed = 1234,
max_points_per_centroid = 256 (subsample with rand_perm(n, seed)), initial centroids =
x[rand_perm(n, seed + 1)[:k]], mean update, empty clusters re-seeded by split_clusters (EPS = 1/1024,
RandomGenerator(1234)).  SURVEY 8f-2 ranks this row "next": the structure matches faiss, but centroid
parity with the faiss wheel is NOT claimed (rand_perm / mt19937 draw order restated from memory, and the
reference's own build is not reproducible: extract_index.py:40 shuffles unseeded).

Reference call site: `index.train(big_npy)`, rvc/train/process/extract_index.py:65.
"""
from __future__ import annotations

import numpy as np

EPS = 1.0 / 1024.0


def faiss_rand_perm(n: int, seed: int) -> np.ndarray:
    """faiss utils/random.cpp rand_perm: Fisher-Yates driven by std::mt19937(seed) % (n - i)."""
    rs = np.random.RandomState(seed)                       # init_genrand(seed), same as std::mt19937(seed)
    draws = rs.randint(0, 2 ** 32, size=max(n - 1, 0), dtype=np.uint64)
    perm = np.arange(n, dtype=np.int64)
    span = np.arange(n, 1, -1, dtype=np.uint64)            # n - i for i = 0 .. n-2
    j = (np.arange(n - 1, dtype=np.int64) + (draws % span).astype(np.int64)).tolist()
    p = perm.tolist()
    for i, i2 in enumerate(j):
        p[i], p[i2] = p[i2], p[i]
    return np.asarray(p, dtype=np.int64)


def train_ivf_centroids(index, x: np.ndarray, k: int, niter: int = 10, seed: int = 1234,
                        max_points_per_centroid: int = 256) -> np.ndarray:
    import torch
    n, d = x.shape
    if n < k:
        raise RuntimeError(f"Number of training points ({n}) should be at least as large as number of clusters ({k})")
    if n > k * max_points_per_centroid:
        x = x[faiss_rand_perm(n, seed)[: k * max_points_per_centroid]]
        n = x.shape[0]
    dev = torch.device(f"cuda:{index.device}")
    xd = torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(dev)
    cent = xd[torch.from_numpy(faiss_rand_perm(n, seed + 1)[:k]).to(dev)].clone()
    rng = np.random.RandomState(1234)
    for _ in range(niter):
        index.set_centroids(cent.cpu().numpy())
        assign, _ = index.assign(xd)                                     # GPU: n x k x d contraction
        a = assign.to(torch.int64)
        ok = a >= 0
        cnt = torch.bincount(a[ok], minlength=k)
        sums = torch.zeros((k, d), dtype=torch.float32, device=dev)
        sums.index_add_(0, a[ok], xd[ok])
        nz = cnt > 0
        cent = torch.where(nz[:, None], sums * (1.0 / cnt.clamp(min=1).to(torch.float32))[:, None], cent)   # faiss: c *= 1 / hassign
        empty = torch.nonzero(~nz).flatten().tolist()
        if empty:                                                          # faiss split_clusters
            h = cnt.cpu().numpy().astype(np.float64)
            c = cent.cpu().numpy()
            sign = np.where(np.arange(d) % 2 == 0, 1.0, -1.0).astype(np.float32)
            for ci in empty:
                cj = 0
                while True:
                    p = (h[cj] - 1.0) / float(n - k)
                    r = float(np.float32(rng.randint(0, 2 ** 32, dtype=np.uint64)) / np.float32(4294967295.0))   # mt() / float(mt.max())
                    if r < p:
                        break
                    cj = (cj + 1) % k
                c[ci] = c[cj]
                c[ci] *= (1 + sign * EPS)
                c[cj] *= (1 - sign * EPS)
                h[ci] = h[cj] // 2
                h[cj] -= h[ci]
            cent = torch.from_numpy(c).to(dev)
    return cent.cpu().numpy()
