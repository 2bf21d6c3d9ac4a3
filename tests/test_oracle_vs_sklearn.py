"""CPU: the oracle's IVF-Flat search semantics against an implementation nobody here wrote -- scikit-learn's brute-force
nearest neighbours (float64 distances) -- on a tie-free problem.

faiss itself cannot be installed (parity of the faiss arithmetic is unpinned); this pins the SEMANTICS of the
restatement on independent code: nearest centroid by squared L2 (IndexFlatL2 k = 1), then the 8 nearest rows of that ONE
list in ascending distance (IVFFlatScanner + heap, nprobe = 1), ids = positions in add order, pads (FLT_MAX, -1) for short
lists.  Bits are not compared (sklearn's arithmetic order is its own): every decision the two disagree on must be a
documented near-tie, i.e. the float64 gap at that rank is within fp32 rounding of the distance."""
from importlib import import_module

import numpy as np
import pytest

sk_pairwise = pytest.importorskip("sklearn.metrics.pairwise")
sk_neighbors = pytest.importorskip("sklearn.neighbors")


@pytest.mark.parametrize("d,iid,n,nlist", [(256, False, 12_000, 96), (128, True, 3_000, 600)])
def test_oracle_search_semantics_match_sklearn_brute_force(oracle, d, iid, n, nlist):
    synth = import_module("codename-rvc-fork-3_b200.synth")
    pr = synth.make_host_problem(n, d, 600, nlist, seed=7, dup_frac=0.0, iid=iid)   # iid: skewed lists, many shorter than 8
    x, cent, q = pr.x, pr.centroids, pr.q
    oidx = oracle.OracleIndex.from_vectors(x, cent)
    D, I = oidx.search(q, k=8)
    a_oracle, _ = oracle.coarse_assign(q, cent)

    x64, c64, q64 = x.astype(np.float64), cent.astype(np.float64), q.astype(np.float64)
    # 1. coarse: sklearn's argmin over centroids; disagreements only where the float64 gap is a rounding-size tie
    a_sk = sk_pairwise.pairwise_distances_argmin(q64, c64, metric="sqeuclidean")
    dq = ((q64[:, None, :] - c64[None, :, :]) ** 2).sum(-1)
    differ = np.nonzero(a_sk != a_oracle)[0]
    for i in differ:
        assert abs(dq[i, a_sk[i]] - dq[i, a_oracle[i]]) <= 1e-5 * dq[i].min(), i
    assert len(differ) <= 2
    # 2. list membership: the oracle's lists are exactly sklearn's assignment of the stored rows (same tie caveat)
    ax_sk = sk_pairwise.pairwise_distances_argmin(x64, c64, metric="sqeuclidean")
    ax_oracle, _ = oracle.coarse_assign(x, cent)
    assert (ax_sk != ax_oracle).sum() <= 3
    # 3. scan: per probed list, sklearn brute-force 8-NN over the rows of that list (ids = global row numbers)
    checked = padded = 0
    for lst in np.unique(a_oracle):
        rows = np.nonzero(ax_oracle == lst)[0]                     # ascending ids == add order
        qs = np.nonzero(a_oracle == lst)[0]
        k = min(8, len(rows))
        if k == 0:
            assert (I[qs] == -1).all()
            continue
        nn = sk_neighbors.NearestNeighbors(n_neighbors=k, algorithm="brute", metric="sqeuclidean").fit(x64[rows])
        dist, idx = nn.kneighbors(q64[qs])
        for r, qi in enumerate(qs):
            got_ids, want_ids = I[qi, :k], rows[idx[r]]
            if not np.array_equal(got_ids, want_ids):
                # any disagreement must be between neighbours whose float64 distances are a rounding-size tie
                bad = np.nonzero(got_ids != want_ids)[0]
                for b in bad:
                    d_got = ((q64[qi] - x64[got_ids[b]]) ** 2).sum()
                    assert abs(d_got - dist[r, b]) <= 2e-6 * max(dist[r, b], 1e-12), (qi, b)
            assert np.allclose(D[qi, :k], dist[r], rtol=2e-5, atol=1e-6)      # fp32 canonical order vs float64
            assert (np.diff(D[qi, :k]) >= 0).all()
            if k < 8:
                assert (I[qi, k:] == -1).all() and (D[qi, k:] == np.float32(3.4028234663852886e38)).all()
                padded += 1
            checked += 1
    assert checked == len(q) and (not iid or padded > 0)


def test_oracle_kmeans_is_lloyd_from_the_faiss_init_like_sklearn(oracle):
    """oracle.kmeans_train (faiss Clustering::train restated) against scikit-learn's Lloyd iterations started from
    the SAME initial centroids (the first k rows of faiss's rand_perm(seed + 1)), on data without empty clusters and below
    the 256-points-per-centroid subsampling threshold: ten plain Lloyd steps, so the centroids agree up to the few points
    whose two nearest centroids are a rounding-size tie, and the objective decreases monotonically."""
    sk_cluster = pytest.importorskip("sklearn.cluster")
    synth = import_module("codename-rvc-fork-3_b200.synth")
    pr = synth.make_host_problem(6_000, 64, 0, 40, seed=11, dup_frac=0.0)
    x, k = pr.x, 40
    cent, a_last, obj = oracle.kmeans_train(x, k, niter=10)
    assert all(b <= a * (1 + 1e-6) for a, b in zip(obj[:-1], obj[1:]))
    init = x[oracle.rand_perm(len(x), 1235)[:k]]
    # 10 assignments + 10 mean updates in faiss == max_iter = 10 in sklearn (tol = 0: it only stops early once the labels
    # no longer change, after which further faiss iterations change nothing either)
    km = sk_cluster.KMeans(n_clusters=k, init=init.astype(np.float64), n_init=1, max_iter=10, tol=0.0, algorithm="lloyd")
    km.fit(x.astype(np.float64))
    assert np.bincount(a_last, minlength=k).min() > 0                       # no split_clusters involved
    diff = np.abs(cent.astype(np.float64) - km.cluster_centers_).max(axis=1)
    assert np.median(diff) <= 1e-5 and (diff <= 1e-3).mean() >= 0.9, (np.median(diff), diff.max())
    obj_sk = ((x.astype(np.float64) - km.cluster_centers_[km.predict(x.astype(np.float64))]) ** 2).sum()
    a_fin, d_fin = oracle.coarse_assign(x, cent)
    assert abs(float(d_fin.astype(np.float64).sum()) - obj_sk) <= 1e-3 * obj_sk
