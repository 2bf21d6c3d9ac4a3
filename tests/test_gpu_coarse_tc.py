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
    v, j, qn2, cmax = index.debug_coarse_tc(torch.from_numpy(q).cuda(), split=split, cluster=cluster)
    assert v.shape[1] % 8 == 0 and v.shape[1] >= 8
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
    # merged top-8 over splits == the 8 smallest reference values
    for i in range(nq):
        valid = j[i] >= 0
        assert valid.sum() >= 8 and (j[i][valid] < nlist).all()
        got = v[i][valid]
        ref_at = vref[i, j[i][valid]]
        assert np.abs(got - ref_at).max() <= 2e-3, np.abs(got - ref_at).max()      # fp32 accumulation only
        kth = np.sort(vref[i])[7]
        best8 = np.sort(got)[:8]
        assert np.abs(best8 - np.sort(vref[i])[:8]).max() <= 4e-3
        assert (np.sort(vref[i, j[i][valid]])[:8] <= kth + 4e-3).all()
    # the margin the re-check uses really bounds the tensor-core error (DESIGN.md "Numerics")
    kdim = 3 * d if split else d
    rel = (12 * 1.01 / 262144 if split else 1.01 / 64) + kdim / 2097152
    qn = np.sqrt(qn2.astype(np.float64))
    margin = rel * qn * cmax + (qn2 + cmax * cmax) / 1048576
    for i in range(0, nq, 7):
        valid = j[i] >= 0
        err = np.abs(v[i][valid].astype(np.float64) - vtrue[i, j[i][valid]])
        assert err.max() <= 0.5 * margin[i], (err.max(), margin[i])


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
