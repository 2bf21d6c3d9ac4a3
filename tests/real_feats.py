"""Test helper: problems built from REAL ContentVec frame features (tests/golden/contentvec_real_318.npz, made by
tests/golden/make_real_feats_fixture.py from the data files the reference ships).  Real rows have a few dominant dimensions
(max |x| 3.6 where the median per-dimension maximum is 0.53) -- the case a per-row power-of-two e4m3 scale and a bf16 margin
have to survive; the Gaussian blobs of synth.py do not have it."""
import os

import numpy as np

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "contentvec_real_318.npz")


def real_rows() -> np.ndarray:
    return np.load(FIXTURE)["rows"].astype(np.float32)


def real_problem(n: int, nq: int, nlist: int, seed: int = 7, jitter: float = 0.25, lloyd_iters: int = 2):
    """(x, centroids, q): x = n training rows = a real frame + per-dimension noise of `jitter` x that dimension's standard
    deviation (what more audio of the same speaker looks like to first order); centroids = nlist rows of x + Lloyd iterations;
    q = the real frames themselves, tiled with a small jitter (a clip converted against an index of its own speaker)."""
    rng = np.random.default_rng(seed)
    rows = real_rows()
    sd = rows.std(axis=0).astype(np.float32)
    x = rows[rng.integers(0, len(rows), n)] + rng.standard_normal((n, rows.shape[1])).astype(np.float32) * (jitter * sd)
    x = np.ascontiguousarray(x.astype(np.float32))
    cent = x[rng.choice(n, size=nlist, replace=False)].copy()
    for _ in range(lloyd_iters):
        yn = np.einsum("ij,ij->i", cent, cent)
        a = np.concatenate([(yn[None, :] - 2.0 * (x[s:s + 8192] @ cent.T)).argmin(axis=1) for s in range(0, n, 8192)])
        sums = np.zeros(cent.shape, np.float64)
        np.add.at(sums, a, x)
        cnt = np.bincount(a, minlength=nlist)
        nz = cnt > 0
        cent[nz] = (sums[nz] / cnt[nz, None]).astype(np.float32)
    q = rows[rng.integers(0, len(rows), nq)] + rng.standard_normal((nq, rows.shape[1])).astype(np.float32) * (0.05 * sd)
    q[:len(rows)] = rows[:nq]                      # the real frames themselves come first
    return x, np.ascontiguousarray(cent.astype(np.float32)), np.ascontiguousarray(q.astype(np.float32))
