"""GPU: the tcgen05 coarse path.  (1) the raw tensor-core output against a float64 restatement of
v = ||c||^2 - 2 <bf16(q), bf16(c)>, which validates descriptors / swizzle / TMEM plumbing and the error
margin the exact re-check relies on; (2) the end result (tensor cores + exact re-check) bit-identical to
the brute-force exact kernel and to the oracle, including near-tie-rich centroid sets."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bf16_round(x):
    """round-to-nearest-even float32 -> bfloat16 -> float32 (numpy)."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).reshape(np.shape(x))


def clustered_problem(rng, nq, nlist, d, sib=0.03):
    blobs = (rng.standard_normal((max(4, nlist // 4), d)) * 0.30).astype(np.float32)
    cent = (blobs[np.arange(nlist) // 4 % len(blobs)] + rng.standard_normal((nlist, d)).astype(np.float32) * sib).astype(np.float32)
    q = (cent[rng.integers(0, nlist, nq)] + rng.standard_normal((nq, d)).astype(np.float32) * 0.14).astype(np.float32)
    return q, cent


def check_candidate_lists(v, j, dropped, margin, vref, nlist, tol):
    """What the tensor-core stage hands to the exact re-check, per query: lists (one per centroid split and epilogue column
    group) of up to 8 (v, centroid) pairs within margin(q) of the smallest v the list saw, plus the smallest v each list let
    go of.  Every kept value is the reference value of its centroid; the overall best is always kept; and every centroid
    within the margin of the best is present unless some list reports having dropped something that near."""
    nq = v.shape[0]
    for i in range(nq):
        valid = j[i] >= 0
        assert valid.any() and (j[i][valid] < nlist).all()
        assert len(np.unique(j[i][valid])) == valid.sum()                                   # no centroid twice
        assert np.isinf(v[i][~valid]).all()
        got = v[i][valid]
        assert np.abs(got - vref[i, j[i][valid]]).max() <= tol, np.abs(got - vref[i, j[i][valid]]).max()   # fp32 accumulation only
        best = vref[i].min()
        assert abs(got.min() - best) <= tol
        if dropped[i].min() > best + margin[i] + 2 * tol:
            need = np.flatnonzero(vref[i] <= best + margin[i] - 2 * tol)
            assert np.isin(need, j[i][valid]).all(), (i, need, j[i][valid])


@pytest.mark.parametrize("cluster", [1, 2])
@pytest.mark.parametrize("split", [0, 1])
@pytest.mark.parametrize("shape", [(300, 1000, 768), (200, 700, 256)])
def test_tc_raw_values_match_float64(rvcr, cluster, split, shape):
    import torch
    nq, nlist, d = shape
    rng = np.random.default_rng(nq + d + split)
    q, cent = clustered_problem(rng, nq, nlist, d)
    index = rvcr.index_factory(d, f"IVF{nlist},Flat")
    index.set_centroids(cent)
    v, j, qn2, cmax, margin, dropped = index.debug_coarse_tc(torch.from_numpy(q).cuda(), split=split, cluster=cluster)
    assert v.shape[1] % 8 == 0 and v.shape[1] >= 8 and dropped.shape == (nq, v.shape[1] // 8)
    q64, c64 = q.astype(np.float64), cent.astype(np.float64)
    cn = (c64 ** 2).sum(1)
    assert np.allclose(qn2, (q64 ** 2).sum(1), rtol=1e-5)
    assert abs(cmax - np.sqrt(cn.max())) <= 1e-4 * cmax
    if split:
        qh, ch = bf16_round(q), bf16_round(cent)
        ql, cl = bf16_round(q - qh), bf16_round(cent - ch)
        dot = qh.astype(np.float64) @ ch.astype(np.float64).T + ql.astype(np.float64) @ ch.astype(np.float64).T \
            + qh.astype(np.float64) @ cl.astype(np.float64).T
    else:
        dot = bf16_round(q).astype(np.float64) @ bf16_round(cent).astype(np.float64).T
    vref = cn[None, :] - 2.0 * dot                     # what the tensor cores should produce
    vtrue = cn[None, :] - 2.0 * (q64 @ c64.T)          # exact-arithmetic value
    check_candidate_lists(v, j, dropped, margin, vref, nlist, tol=2e-3)
    # the margin the re-check uses (built from the exact residual norms of both operands, coarse_tc.cu tc_margin) really
    # bounds the tensor-core error, and on ordinary rows it is not larger than the old worst-case-looking constant
    kdim = 3 * d if split else d
    qn = np.sqrt(qn2.astype(np.float64))
    old = ((12 * 1.01 / 262144 if split else 1.01 / 64) + kdim / 2097152) * qn * cmax + (qn2 + cmax * cmax) / 1048576
    assert (margin > 0).all() and (margin <= 1.25 * old + 0.02).all(), (margin / old).max()
    for i in range(0, nq, 7):
        valid = j[i] >= 0
        err = np.abs(v[i][valid].astype(np.float64) - vtrue[i, j[i][valid]])
        assert err.max() <= 0.5 * margin[i], (err.max(), margin[i])


def midpoint_rows(rng, n, d, down_cols, scale=1.0, steps=128):
    """Rows whose every element sits 2^-20 (relative) off a bf16 rounding midpoint: |x - bf16(x)| is the format's worst
    case, 2^-8 relative (8 significant bits, round to nearest) -- twice what a typical row shows.  Columns in `down_cols`
    round down (towards the grid value a), the others up (to a + ulp); a is a random bf16 grid value in [1, 1 + steps/128):
    steps = 2 keeps the significand next to 1, where the relative error is largest."""
    a = 1.0 + rng.integers(0, steps, (n, d)) / 128.0
    tiny = 2.0 ** -20
    x = a + 2.0 ** -8 + np.where(down_cols[None, :], -tiny, tiny)
    return (x * scale).astype(np.float32), (a * scale).astype(np.float32)


def adversarial_midpoint_problem(rng, d, nlist, scale, nq=40, k=9):
    """The round-1 review's counterexample, randomised: q rounds down on the first half of the columns and up on the
    second; c_a = q on the first half (0 elsewhere) rounds down with it, c_b lives on the second half k bf16 ulps above q
    and rounds up.  True distances: c_a nearer by sum (k ulp)^2 (1.9 at d = 768, k = 9, scale 1); the bf16 products
    under-estimate <q, c_a> and over-estimate <q, c_b> by the full 2^-8 on every element, so v_b - v_a is off by ~12 --
    beyond a margin derived from a 2^-9 rounding assumption, inside the one built from the exact residual norms."""
    half = d // 2
    down = np.arange(d) < half
    q, a = midpoint_rows(rng, nq, d, down, scale, steps=2)
    # second half of every query uses the same grid values as its first half: the two halves have (almost) equal norms
    tiny = 2.0 ** -20
    q[:, half:] = ((a[:, :half].astype(np.float64) / scale + 2.0 ** -8 + tiny) * scale).astype(np.float32)
    cent = -(rng.random((nlist, d)) * 0.4 + 0.2).astype(np.float32) * np.float32(scale)     # far away, small norms
    want = np.empty(nq, np.int64)
    for i in range(nq):
        ja, jb = 2 * i, 2 * i + 1
        ca = np.zeros(d, np.float32)
        ca[:half] = q[i, :half]
        cb = np.zeros(d, np.float32)
        cb[half:] = ((a[i, :half].astype(np.float64) / scale + k * 2.0 ** -7 + 2.0 ** -8 + tiny) * scale).astype(np.float32)
        cent[ja], cent[jb] = ca, cb
        want[i] = ja
    return q, cent, want


@pytest.mark.parametrize("mode", [2, 3])
@pytest.mark.parametrize("d,scale", [(768, 1.0), (768, 0.25), (256, 1.0), (256, 2.0 ** -5)])
def test_tc_adversarial_bf16_midpoints(rvcr, oracle, mode, d, scale):
    """bf16 rounds to 2^-8 relative in the worst case, not 2^-9: rows built on rounding midpoints with the errors of q, c_a
    and c_b aligned must still get the exact kernel's (= the oracle's) centroid.  Fails on a margin that assumes 2^-9."""
    import torch
    rng = np.random.default_rng(int(d + 1000 * scale) + mode)
    nlist = 320
    q, cent, want = adversarial_midpoint_problem(rng, d, nlist, scale)
    q64, c64 = q.astype(np.float64), cent.astype(np.float64)
    dist = ((q64[:, None, :] - c64[None, :, :]) ** 2).sum(-1)
    assert np.array_equal(dist.argmin(1), want)                      # c_a is the true nearest centroid ...
    gap = dist[np.arange(len(q)), want + 1] - dist[np.arange(len(q)), want]
    assert (gap > 0).all() and (gap < 2.5 * scale * scale * d / 768 + 1e-3).all()      # ... by a little
    if mode == 2:
        # the bf16 products alone get it wrong: that is what makes the case adversarial
        vb = (c64 ** 2).sum(1)[None, :] - 2.0 * bf16_round(q).astype(np.float64) @ bf16_round(cent).astype(np.float64).T
        assert (vb.argmin(1) != want).all()
        assert ((vb[np.arange(len(q)), want] - vb[np.arange(len(q)), want + 1]) >
                (1.01 / 64 + d / 2097152) * np.sqrt((q64 ** 2).sum(1)) * np.sqrt((c64 ** 2).sum(1).max())).all(), \
            "not beyond the round-1 margin: the case lost its teeth"
    index = rvcr.index_factory(d, f"IVF{nlist},Flat")
    index.set_centroids(cent)
    qd = torch.from_numpy(q).cuda()
    index.set_option("coarse_mode", 1)
    l_ex, d_ex = [t.cpu().numpy() for t in index.assign(qd)]
    assert np.array_equal(l_ex, want)
    a, cd = oracle.coarse_assign(q, cent)
    assert np.array_equal(a, want) and np.array_equal(cd, d_ex)
    for cluster in (2, 1):
        index.set_option("coarse_mode", mode)
        index.set_option("tc_cluster", cluster)
        l_tc, d_tc = [t.cpu().numpy() for t in index.assign(qd)]
        assert np.array_equal(l_tc, want), (mode, cluster, np.flatnonzero(l_tc != want)[:5])
        assert np.array_equal(d_tc, d_ex)
        # raw values: the margin bounds the error on these rows too
        v, j, qn2, cmax, margin, _ = index.debug_coarse_tc(qd, split=1 if mode == 3 else 0, cluster=cluster)
        vtrue = (c64 ** 2).sum(1)[None, :] - 2.0 * (q64 @ c64.T)
        for i in range(len(q)):
            valid = j[i] >= 0
            err = np.abs(v[i][valid].astype(np.float64) - vtrue[i, j[i][valid]])
            assert err.max() <= 0.5 * margin[i], (i, err.max(), margin[i])


@pytest.mark.parametrize("d", [256, 768])
def test_tc_random_midpoint_rows_tie_rich(rvcr, d):
    """Every element of every row within 2^-20 of a bf16 midpoint, random rounding directions, centroids crowded around the
    queries: the residual norms are twice the typical ones and near-ties abound; tensor cores + re-check == exact kernel."""
    import torch
    rng = np.random.default_rng(d)
    nlist, nq = 600, 500
    base, _ = midpoint_rows(rng, 8, d, rng.random(d) < 0.5, 0.25)
    down = rng.random(d) < 0.5
    cent, _ = midpoint_rows(rng, nlist, d, down, 0.25)
    # crowd them: keep only 24 random columns different between centroids of a group of 8
    for g in range(nlist):
        keep = rng.choice(d, 24, replace=False)
        row = base[g % 8].copy()
        row[keep] = cent[g, keep]
        cent[g] = row
    q = cent[rng.integers(0, nlist, nq)].copy()
    cols = rng.integers(0, d, (nq, 6))
    q[np.arange(nq)[:, None], cols] *= np.float32(1 + 2.0 ** -7)
    index = rvcr.index_factory(d, f"IVF{nlist},Flat")
    index.set_centroids(cent)
    qd = torch.from_numpy(q).cuda()
    index.set_option("coarse_mode", 1)
    l_ex, d_ex = [t.cpu().numpy() for t in index.assign(qd)]
    for mode in (2, 3):
        index.set_option("coarse_mode", mode)
        l_tc, d_tc = [t.cpu().numpy() for t in index.assign(qd)]
        assert np.array_equal(l_tc, l_ex) and np.array_equal(d_tc, d_ex), mode


@pytest.mark.parametrize("mode,cluster", [(2, 1), (2, 2), (3, 2)])
def test_tc_assign_bit_identical_to_exact_kernel(rvcr, oracle, mode, cluster):
    import torch
    rng = np.random.default_rng(99)
    nq, nlist, d = 6000, 6000, 768
    q, cent = clustered_problem(rng, nq, nlist, d, sib=0.02)      # sibling centroids inside the bf16 margin
    cent[4001] = cent[17]                                         # exact duplicate centroid: lowest id must win
    q[5] = cent[17] + np.float32(0.01)
    q[6, 3] = np.nan
    index = rvcr.index_factory(d, f"IVF{nlist},Flat")
    index.set_centroids(cent)
    qd = torch.from_numpy(q).cuda()
    index.set_option("coarse_mode", 1)
    l_ex, d_ex = [t.cpu().numpy() for t in index.assign(qd)]
    index.set_option("coarse_mode", mode)
    index.set_option("tc_cluster", cluster)
    l_tc, d_tc = [t.cpu().numpy() for t in index.assign(qd)]
    assert np.array_equal(l_tc, l_ex)
    ok = l_ex >= 0
    assert np.array_equal(d_tc[ok], d_ex[ok])
    assert l_tc[5] == 17 and l_tc[6] == -1
    st = index.stats()
    assert st["coarse_tc_queries"] == nq and st["coarse_candidates"] >= nq - st["coarse_fallback"] - 1
    # oracle on a sample (brute force on the CPU is slow)
    sel = rng.choice(nq, 200, replace=False)
    a, cd = oracle.coarse_assign(q[sel], cent)
    assert np.array_equal(l_tc[sel], a)
    assert np.array_equal(d_tc[sel][a >= 0], cd[a >= 0])
    print("tc stats", mode, cluster, st)


def test_tc_small_and_ragged_shapes(rvcr):
    """n not a multiple of 128/256, nlist not a multiple of 256, tiny n (many centroid splits)."""
    import torch
    rng = np.random.default_rng(3)
    for nq, nlist, d in [(1, 1000, 256), (37, 300, 768), (129, 257, 256), (700, 5000, 768)]:
        q, cent = clustered_problem(rng, nq, nlist, d)
        index = rvcr.index_factory(d, f"IVF{nlist},Flat")
        index.set_centroids(cent)
        qd = torch.from_numpy(q).cuda()
        index.set_option("coarse_mode", 1)
        l_ex, d_ex = [t.cpu().numpy() for t in index.assign(qd)]
        for cluster in (1, 2):
            index.set_option("coarse_mode", 2)
            index.set_option("tc_cluster", cluster)
            l_tc, d_tc = [t.cpu().numpy() for t in index.assign(qd)]
            assert np.array_equal(l_tc, l_ex) and np.array_equal(d_tc, d_ex), (nq, nlist, d, cluster)


@pytest.mark.parametrize("d", [768, 256, 1024])
def test_tc_resident_query_blocks(rvcr, d):
    """option tc_resident: 0 streams both operands; 8 / 9 / 10 pin that many 64-element k-blocks of the query
    operand in shared memory (d = 256: everything resident; d = 1024: 16 k-blocks, 6-8 still streamed).  Several
    work items per CTA pair (resident block reloaded), identical assignments and distances every way."""
    import torch
    rng = np.random.default_rng(5)
    nq, nlist = 128 * 2 * 74 * 2 + 77, 700          # > 2 m-blocks per CTA pair on 148 SMs
    q, cent = clustered_problem(rng, nq, nlist, d)
    index = rvcr.index_factory(d, f"IVF{nlist},Flat")
    index.set_centroids(cent)
    qd = torch.from_numpy(q).cuda()
    index.set_option("coarse_mode", 1)
    l_ex, d_ex = [t.cpu().numpy() for t in index.assign(qd)]
    index.set_option("coarse_mode", 2)
    for cluster in (2, 1):
        index.set_option("tc_cluster", cluster)
        for res in (0, 6, 8, 9, 10):
            index.set_option("tc_resident", res)
            l_tc, d_tc = [t.cpu().numpy() for t in index.assign(qd)]
            assert np.array_equal(l_tc, l_ex) and np.array_equal(d_tc, d_ex), (d, cluster, res)
