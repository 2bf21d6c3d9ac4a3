"""index.train(x) for IndexIVFFlat: faiss `Clustering::train` -- a thin host wrapper over
`rvcr_index_train` (csrc/kmeans.cu), where the subsample, the initial centroids, the 10 Lloyd iterations (assignment
on the tcgen05 coarse quantizer + exact re-check, deterministic point-order fp32 means) and faiss's split_clusters
for empty clusters all run natively.

Follows faiss 1.7.3 defaults for Level1Quantizer: niter = 10, nredo = 1, seed = 1234, max_points_per_centroid = 256
(subsample with rand_perm(n, seed)), initial centroids = x[rand_perm(n, seed + 1)[:k]], mean update, empty clusters
re-seeded by split_clusters (EPS = 1/1024, RandomGenerator(1234)).  Centroid parity with the faiss WHEEL is not
claimed (faiss is not installable here and the reference's own build is not reproducible: extract_index.py:40 shuffles
unseeded); parity with the oracle's restatement of the same algorithm is bit for bit (tests/test_gpu_build_flow.py).

Reference call site: `index.train(big_npy)`, rvc/train/process/extract_index.py:65.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._ffi import check, lib


def train_ivf_centroids(index, x: np.ndarray, k: int, niter: int = 10) -> np.ndarray:
    """Train `index` (an untrained, empty IndexIVFFlat with nlist == k) on host rows x; returns the centroids."""
    x = np.ascontiguousarray(x, np.float32)
    if x.ndim != 2 or x.shape[1] != index.d or k != index.nlist:
        raise AssertionError(f"train: expected (n, {index.d}) rows and k == nlist == {index.nlist}")
    index.set_option("kmeans_niter", int(niter))
    try:
        check(lib().rvcr_index_train(index.handle, x.ctypes.data_as(C.c_void_p), x.shape[0]))
    finally:
        index.set_option("kmeans_niter", 10)
    return index.get_centroids()
