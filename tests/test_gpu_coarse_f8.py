"""GPU: coarse_mode 4 -- the tcgen05 coarse step with e4m3 operands (kind::f8f6f4, csrc/coarse_tc.cu), first run on a B200 in
round 2: raw values against a float64 restatement of the scaled e4m3 arithmetic, the error margin (built from the exact
residual norms of both operands), the end result bit-identical to the exact kernel (more candidates / fallbacks than bf16 on
tie-rich centroids; the answer must not change), and the auto mode's per-centroid-set decision."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


from oracle.f8_model import e4m3_round, f8_scale  # noqa: E402  (numpy model of the e4m3 arithmetic; CPU-checked in tests/test_f8_margin_model.py)


def clustered_problem(rng, nq, nlist, d, sib=0.03):
    blobs = (rng.standard_normal((max(4, nlist // 4), d)) * 0.30).astype(np.float32)
    cent = (blobs[np.arange(nlist) // 4 % len(blobs)] + rng.standard_normal((nlist, d)).astype(np.float32) * sib).astype(np.float32)
    q = (cent[rng.integers(0, nlist, nq)] + rng.standard_normal((nq, d)).astype(np.float32) * 0.14).astype(np.float32)
    return q, cent


@pytest.mark.parametrize("cluster", [1, 2])
@pytest.mark.parametrize("shape", [(300, 1000, 768), (200, 700, 256)])
def test_f8_raw_values_match_float64(rvcr, cluster, shape):
    import torch
    nq, nlist, d = shape
    rng = np.random.default_rng(nq + d)
    q, cent = clustered_problem(rng, nq, nlist, d)
    q[3] *= np.float32(1e-3)                          # a row far below the others: its own scale keeps it accurate
    index = rvcr.index_factory(d, f"IVF{nlist},Flat")
    index.set_centroids(cent)
    v, j, qn2, cmax, margin, dropped = index.debug_coarse_tc(torch.from_numpy(q).cuda(), split=2, cluster=cluster)
    q64, c64 = q.astype(np.float64), cent.astype(np.float64)
    cn = (c64 ** 2).sum(1)
    sc = f8_scale(np.abs(c64).max())
    sq = f8_scale(np.abs(q64).max(axis=1))
    q8 = e4m3_round(q64 * sq[:, None])
    c8 = e4m3_round(c64 * sc)
    assert np.abs(q8).max() <= 256 and np.abs(c8).max() <= 256      # scaled maxima are < 256 and may round up to 256 (<= 448)
    vref = cn[None, :] - 2.0 * (q8 @ c8.T) / (sq[:, None] * sc)
    vtrue = cn[None, :] - 2.0 * (q64 @ c64.T)
    from test_gpu_coarse_tc import check_candidate_lists
    check_candidate_lists(v, j, dropped, margin, vref, nlist, tol=4e-3)      # fp32 accumulation + the fp32 ||c||^2 only
    qn = np.sqrt(qn2.astype(np.float64))
    assert (margin > 0).all() and (margin <= (0.53 + d / 2097152) * qn * cmax + 0.05).all()   # never above the worst-case form
    worst = 0.0
    for i in range(nq):
        valid = j[i] >= 0
        err = np.abs(v[i][valid].astype(np.float64) - vtrue[i, j[i][valid]])
        assert err.max() <= 0.5 * margin[i], (i, err.max(), margin[i])
        worst = max(worst, err.max() / margin[i])
    print(f"e4m3 coarse: worst |v - v_true| / margin = {worst:.3f} (must stay below 0.5); "
          f"margin / (|q| c_max) median {np.median(margin / (qn * cmax)):.3f} (worst-case form 0.53)")


@pytest.mark.parametrize("cluster", [1, 2])
def test_f8_assign_bit_identical_to_exact_kernel(rvcr, oracle, cluster):
    import torch
    rng = np.random.default_rng(98)
    nq, nlist, d = 6000, 6000, 768
    q, cent = clustered_problem(rng, nq, nlist, d, sib=0.03)
    cent[4001] = cent[17]                                         # exact duplicate centroid: lowest id must win
    q[5] = cent[17] + np.float32(0.01)
    q[6, 3] = np.nan
    q[7] = 0                                                      # all-zero row: no scale -> exact kernel
    q[8, 11] = np.inf
    index = rvcr.index_factory(d, f"IVF{nlist},Flat")
    index.set_centroids(cent)
    qd = torch.from_numpy(q).cuda()
    index.set_option("coarse_mode", 1)
    l_ex, d_ex = [t.cpu().numpy() for t in index.assign(qd)]
    index.set_option("coarse_mode", 4)
    index.set_option("tc_cluster", cluster)
    s0 = index.stats()
    l_tc, d_tc = [t.cpu().numpy() for t in index.assign(qd)]
    s1 = index.stats()
    assert np.array_equal(l_tc, l_ex)
    ok = l_ex >= 0
    assert np.array_equal(d_tc[ok], d_ex[ok])
    assert l_tc[5] == 17 and l_tc[6] == -1
    fb = s1["coarse_fallback"] - s0["coarse_fallback"]
    cand = s1["coarse_candidates"] - s0["coarse_candidates"]
    print(f"e4m3 coarse, cluster {cluster}: {cand / max(1, nq - fb):.1f} candidates per query, {fb} of {nq} fell back")
    assert fb <= nq // 10                                         # 4 siblings inside the margin, the rest ~100 away


def test_f8_auto_probe_decides_per_centroid_set(rvcr, oracle):
    """option coarse_f8_auto: e4m3 operands are kept where the probe shows (almost) no fallbacks -- the bench's blob
    construction -- and dropped on centroids that crowd inside the margin; results are the exact kernel's either way."""
    import torch
    rng = np.random.default_rng(4)
    d, nlist, nq = 256, 4096, 70_000
    for crowded, want in ((False, 1), (True, -1)):
        if crowded:      # all centroids within a few units of each other: hundreds inside the e4m3 margin
            cent = (rng.standard_normal((1, d)) * 0.3 + rng.standard_normal((nlist, d)) * 0.02).astype(np.float32)
            q = (cent[rng.integers(0, nlist, nq)] + rng.standard_normal((nq, d)).astype(np.float32) * 0.02).astype(np.float32)
        else:
            q, cent = clustered_problem(rng, nq, nlist, d)
        index = rvcr.index_factory(d, f"IVF{nlist},Flat")
        index.set_centroids(cent)
        qd = torch.from_numpy(q).cuda()
        index.set_option("coarse_mode", 0)
        index.set_option("coarse_f8_auto", 1)
        assert index.get_option("coarse_f8_decision") == 0
        l_auto, d_auto = [t.cpu().numpy() for t in index.assign(qd)]
        assert index.get_option("coarse_f8_decision") == want, crowded
        sel = rng.choice(nq, 4000, replace=False)
        index.set_option("coarse_mode", 1)
        l_ex, d_ex = [t.cpu().numpy() for t in index.assign(torch.from_numpy(q[sel]).cuda())]
        assert np.array_equal(l_auto[sel], l_ex) and np.array_equal(d_auto[sel], d_ex)
