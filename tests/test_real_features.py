"""CPU: the tensor-core coarse step's error margin and the oracle on REAL ContentVec features (tests/real_feats.py).

The Gaussian blobs used elsewhere have no dominant dimensions; real frames do, which is the hard case for a per-row e4m3
scale.  Checked here, on the numpy restatement of the kernel's margin (oracle.f8_model.rigorous_margin == tc_margin /
make_margin in csrc/coarse_tc.cu): |v_lowprec - v_true| <= margin / 2 for every (query, centroid) pair, for bf16 and e4m3
operands, so the candidate set {v <= v_min + margin} always holds the true nearest centroid.  Also: the oracle's canonical
coarse assignment and list scan agree with float64 brute force on these rows outside exact fp32 ties."""
import numpy as np
import pytest

from real_feats import real_problem, real_rows


def test_fixture_is_the_real_distribution():
    r = real_rows()
    assert r.shape == (318, 768) and r.dtype == np.float32 and np.isfinite(r).all()
    per_dim_max = np.abs(r).max(axis=0)
    assert per_dim_max.max() > 5 * np.median(per_dim_max)          # a few dominant dimensions


@pytest.mark.parametrize("jitter", [0.25, 1.0])
@pytest.mark.parametrize("kind", ["bf16", "e4m3"])
def test_rigorous_margin_bounds_the_error_on_real_features(kind, jitter):
    """jitter 0.25: 1024 centroids squeezed around 318 distinct frames -- dense near-ties, many centroids inside the margin (the
    kernel then falls back to the exact path for the queries whose 8-slot lists overflow; the result does not change);
    jitter 1.0: training rows as far from a frame as the speaker's own frame-to-frame spread."""
    from oracle import f8_model
    x, cent, q = real_problem(20_000, 600, 1024, jitter=jitter)
    v, vtrue, margin = f8_model.rigorous_margin(q, cent, kind)
    err = np.abs(v - vtrue)
    ratio = (err / margin[:, None]).max()
    assert ratio <= 0.5, ratio
    cand = v <= v.min(axis=1, keepdims=True) + margin[:, None]
    assert cand[np.arange(len(q)), vtrue.argmin(axis=1)].all()
    per_query = cand.sum(axis=1)
    print(f"{kind}, jitter {jitter}: worst |v - v_true| / margin = {ratio:.3f}; candidates per query mean {per_query.mean():.1f}, "
          f"median {np.median(per_query):.0f}, max {per_query.max()}")


def test_oracle_on_real_features_matches_float64_brute_force(oracle=None):
    import oracle as o
    o.lib()
    x, cent, q = real_problem(20_000, 600, 512)
    a, dist = o.coarse_assign(q, cent)
    d64 = ((q.astype(np.float64)[:, None, :] - cent.astype(np.float64)[None, :, :]) ** 2).sum(-1)
    best = d64.argmin(axis=1)
    # the canonical fp32 distance decides; it may differ from float64 only where the two smallest float64 distances are within
    # fp32 rounding of each other
    srt = np.sort(d64, axis=1)
    near_tie = (srt[:, 1] - srt[:, 0]) <= 1e-5 * srt[:, 0]
    assert (a == best)[~near_tie].all() and near_tie.mean() < 0.01
    assert np.allclose(dist, d64[np.arange(len(q)), a], rtol=2e-6)
    a_x, _ = o.coarse_assign(x, cent)
    off, vecs, ids = o.build_lists(x, a_x, len(cent))
    D, I = o.scan_topk(q, a, off, vecs, ids)
    for i in range(0, len(q), 37):
        lo, hi = off[a[i]], off[a[i] + 1]
        dd = ((vecs[lo:hi].astype(np.float64) - q[i].astype(np.float64)) ** 2).sum(1)
        k = min(8, hi - lo)
        want = ids[lo:hi][np.argsort(dd, kind="stable")[:k]]
        got = I[i][:k]
        # same neighbours (order may swap inside an fp32 near-tie)
        assert set(got.tolist()) == set(want.tolist()) or np.allclose(np.sort(dd)[:k], D[i][:k], rtol=2e-6)
        assert (I[i][k:] == -1).all()


def test_tie_zone_against_blas_order_on_real_features():
    """Real faiss (>= 20 queries) decides the coarse step by the BLAS expansion |x|^2 + |y|^2 - 2 x.y, the oracle and the GPU by
    the canonical fvec_L2sqr order ("tie zone").  On real-distribution rows the two argmins differ for at
    most a few rows in 10^4 -- measured here on the dense-tie problem (1024 centroids around 318 frames) and the spread one."""
    import oracle as o
    o.lib()
    for jitter in (0.25, 1.0):
        x, cent, q = real_problem(40_000, 2048, 1024, jitter=jitter)
        a_q, _ = o.coarse_assign(q, cent)
        a_x, _ = o.coarse_assign(x, cent)
        b_q, b_x = o.coarse_assign_blas(q, cent), o.coarse_assign_blas(x, cent)
        b_q = b_q[0] if isinstance(b_q, tuple) else b_q
        b_x = b_x[0] if isinstance(b_x, tuple) else b_x
        dq, dx = int((a_q != b_q).sum()), int((a_x != b_x).sum())
        print(f"jitter {jitter}: canonical vs BLAS-order argmin differ for {dq} of {len(q)} queries, {dx} of {len(x)} stored rows")
        assert dq <= 2 and dx <= 20


def test_scan_margin_bounds_the_tf32_error_on_real_features():
    """The list scan's validated split precision (csrc/scan_tc.cu) on real rows: two approximate values v further apart than
    margin(q) must be ordered like the true distances, i.e. |v - vtrue| <= margin / 4 leaves the documented factor 2 for the
    tensor core's own fp32 accumulation (not modelled here: the products are summed exactly)."""
    from oracle import f8_model
    x, cent, q = real_problem(6_000, 400, 64)
    v, vtrue, margin = f8_model.scan_margin(q, x)
    ratio = (np.abs(v - vtrue) / margin[:, None]).max()
    assert ratio <= 0.25, ratio
    # the candidate set {v <= 8th smallest v + margin} holds the true top-8 of every query
    kth = np.sort(v, axis=1)[:, 7]
    cand = v <= (kth + margin)[:, None]
    top8 = np.argsort(vtrue, axis=1, kind="stable")[:, :8]
    assert np.take_along_axis(cand, top8, axis=1).all()
    print(f"tf32 scan: worst |v - v_true| / margin = {ratio:.3f}; candidates per query mean {cand.sum(1).mean():.1f} (8 compulsory)")
