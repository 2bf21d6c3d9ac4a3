"""CPU prototype (numpy) of an EXACT two-level bound for the coarse step (round-1 review, item 5c): ~300 super-centroids,
triangle-inequality pruning of 256-centroid tiles, survivors through the usual exact re-check.

For a query q with an upper bound ub on its nearest-centroid distance (here: the true nearest distance, i.e. the best
bound any scheme could have), a tile T of centroids with super-centroid s_T and radius R_T = max_{c in T} |c - s_T| cannot
hold the nearest centroid if |q - s_T| - R_T > sqrt(ub).  A tensor-core tile of the coarse kernel serves a whole m-block of
256 queries, so a tile can only be skipped if EVERY query of the block prunes it; queries are therefore sorted by their
nearest super-centroid first.  Centroids are grouped into tiles by k-means over the centroids themselves (tile = the 256
centroids of a group, the best case for the radius).

Reports the fraction of (m-block, tile) pairs that can be skipped on (a) the bench's synthetic construction (Gaussian
blobs in 768-d) and (b) the real ContentVec rows the reference ships (tests/golden: 350 + 169 rows, clustered but tiny).
    python tools/two_level_probe.py > profiles/r02_two_level_probe.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def kmeans(x, k, iters, rng):
    c = x[rng.choice(len(x), k, replace=False)].copy()
    for _ in range(iters):
        d = (x ** 2).sum(1)[:, None] + (c ** 2).sum(1)[None, :] - 2 * x @ c.T
        a = d.argmin(1)
        for j in range(k):
            m = a == j
            if m.any():
                c[j] = x[m].mean(0)
    d = (x ** 2).sum(1)[:, None] + (c ** 2).sum(1)[None, :] - 2 * x @ c.T
    return c, d.argmin(1)


def probe(cent, q, tile=256, mblock=256, seed=0):
    rng = np.random.default_rng(seed)
    nlist = len(cent)
    nsuper = max(2, nlist // tile)
    sup, grp = kmeans(cent.astype(np.float64), nsuper, 8, rng)
    radius = np.array([np.sqrt(((cent[grp == g] - sup[g]) ** 2).sum(1)).max() if (grp == g).any() else 0.0 for g in range(nsuper)])
    d2 = (q.astype(np.float64) ** 2).sum(1)[:, None] + (cent.astype(np.float64) ** 2).sum(1)[None, :] - 2 * q.astype(np.float64) @ cent.astype(np.float64).T
    ub = np.sqrt(np.maximum(d2.min(1), 0))                               # the best possible bound: the true nearest distance
    ds = np.sqrt(np.maximum((q.astype(np.float64) ** 2).sum(1)[:, None] + (sup ** 2).sum(1)[None, :] - 2 * q.astype(np.float64) @ sup.T, 0))
    prune_q = ds - radius[None, :] > ub[:, None]                         # query x group: group cannot hold the nearest centroid
    order = np.argsort(ds.argmin(1), kind="stable")                      # queries sorted by nearest super-centroid
    pq = prune_q[order]
    nb = (len(q) + mblock - 1) // mblock
    skipped = sum(int(pq[b * mblock:(b + 1) * mblock].all(0).sum()) for b in range(nb))
    return {"nlist": int(nlist), "queries": int(len(q)), "groups": int(nsuper), "median_group_radius": float(np.median(radius)),
            "median_query_to_nearest_centroid": float(np.median(ub)), "median_query_to_other_super": float(np.median(np.sort(ds, 1)[:, 1])),
            "per_query_groups_prunable_frac": float(prune_q.mean()), "mblock_tiles_skippable_frac": skipped / float(nb * nsuper)}


def main():
    import importlib
    synth = importlib.import_module("codename-rvc-fork-3_b200.synth")
    import torch
    out = {}
    nlist, d = 16384, 768
    cent = synth.all_centroids(nlist, d, torch.device("cpu")).numpy()
    q = synth.device_queries(torch.from_numpy(cent), 8192).numpy()
    out["bench_construction_768d"] = probe(cent, q)
    z = np.load(os.path.join(ROOT, "tests", "golden", "retrieve_v2_768.npz"))
    g = np.load(os.path.join(ROOT, "tests", "golden", "final_proj_v1.npz"))
    real = np.concatenate([z["x"], g["x"]]).astype(np.float32)          # real ContentVec rows (519 x 768)
    rng = np.random.default_rng(1)
    # a "real-like" centroid set: the real rows + small jitter, 4096 centroids (the rows have low intrinsic dimension)
    cent_r = (real[rng.integers(0, len(real), 4096)] + 0.05 * rng.standard_normal((4096, 768))).astype(np.float32)
    q_r = (real[rng.integers(0, len(real), 2048)] + 0.05 * rng.standard_normal((2048, 768))).astype(np.float32)
    out["contentvec_rows_jittered_768d"] = probe(cent_r, q_r, tile=256, mblock=256)
    out["note"] = ("exact triangle-inequality pruning with the TRUE nearest distance as the bound (an upper limit on what any two-level "
                   "scheme can skip); a tile is skippable only if all 256 queries of an m-block prune it")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
