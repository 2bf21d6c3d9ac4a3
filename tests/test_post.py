"""The step after the path: 2x upsample + pitch-protect mix, rvc/infer/pipeline.py:463-481.

CPU: the oracle restatement against goldens produced by the reference's own statements
(tests/golden/make_golden_post.py).  GPU: the stand-alone kernel and the variant fused into the blend epilogue
(rvcr_retrieve_blend_upsample), through the C ABI, bit-exact against goldens and oracle on every blend route
(tcgen05 scan re-check, exact-scan + gather_blend, exact fallback)."""
import numpy as np
import pytest

STANDALONE = ["post_standalone_768", "post_standalone_30"]
CHAINED = [("retrieve_v2_768", "post_v2_768"), ("retrieve_v1_256", "post_v1_256")]


# ------------------------------------------------------------------------------ CPU: oracle pinned ----

@pytest.mark.parametrize("name", STANDALONE)
def test_oracle_post_matches_reference_golden(oracle, golden, name):
    g = golden(name)
    got = oracle.upsample_protect(g["feats"], g["feats0"], g["pitchf"], float(g["protect"]))
    assert got.dtype == np.float32 and got.shape == (2 * len(g["feats"]), g["feats"].shape[1])
    assert np.array_equal(got, g["out_mix"])
    up = oracle.upsample_protect(g["feats"])
    assert np.array_equal(up[0::2], g["feats"]) and np.array_equal(up[1::2], g["feats"])


@pytest.mark.parametrize("rname,pname", CHAINED)
def test_oracle_chain_matches_reference_golden(oracle, golden, rname, pname):
    r, p = golden(rname), golden(pname)
    with np.errstate(all="ignore"):
        got = oracle.upsample_protect(r["out"], r["q"], p["pitchf"], float(p["protect"]))
    assert np.array_equal(got, p["out_mix"], equal_nan=True)
    assert np.array_equal(oracle.upsample_protect(r["out"]), p["out_up"], equal_nan=True)


def test_protect_factor_rules(oracle):
    """pitchff[pitchf > 0] = 1 THEN pitchff[pitchf < 1] = protect: (0, 1) ends at protect, exactly 1 stays 1."""
    f = np.ones((4, 4), np.float32)
    f0 = np.zeros((4, 4), np.float32)
    pitchf = np.array([0.0, 0.5, 1.0, 220.0, -3.0, 0.999, 1.0001, 0.0], np.float32)
    got = oracle.upsample_protect(f, f0, pitchf, 0.25)[:, 0]
    assert np.array_equal(got, np.array([0.25, 0.25, 1, 1, 0.25, 0.25, 1, 0.25], np.float32))


# ------------------------------------------------------------------------------------- GPU parity ----

@pytest.mark.gpu
@pytest.mark.parametrize("name", STANDALONE)
def test_gpu_upsample_protect_standalone(rvcr, oracle, golden, name):
    import torch
    g = golden(name)
    feats = torch.from_numpy(g["feats"]).cuda()
    feats0 = torch.from_numpy(g["feats0"]).cuda()
    pitchf = torch.from_numpy(g["pitchf"]).cuda()
    protect = float(g["protect"])
    out = rvcr.upsample_protect(feats.unsqueeze(0), feats0.unsqueeze(0), pitchf.unsqueeze(0), protect)
    assert out.shape == (1, 2 * feats.shape[0], feats.shape[1])
    assert np.array_equal(out[0].cpu().numpy(), g["out_mix"])
    # protect >= 0.5, or no pitch track: upsample only (pipeline.py:467,475)
    want_up = oracle.upsample_protect(g["feats"])
    assert np.array_equal(rvcr.upsample_protect(feats, feats0, pitchf, 0.5).cpu().numpy(), want_up)
    assert np.array_equal(rvcr.upsample_protect(feats).cpu().numpy(), want_up)
    assert np.array_equal(feats.cpu().numpy(), g["feats"])             # inputs untouched
    with pytest.raises(AssertionError):
        rvcr.upsample_protect(feats, feats0, pitchf[:-1], protect)     # p_len must be 2n
    # larger ragged case vs the oracle (grid-stride path, more rows than CTAs)
    rng = np.random.default_rng(5)
    n, d = 40_003, 256
    a = rng.standard_normal((n, d)).astype(np.float32)
    b = rng.standard_normal((n, d)).astype(np.float32)
    pf = np.where(rng.random(2 * n) < 0.4, 0.0, rng.uniform(0.2, 300.0, 2 * n)).astype(np.float32)
    got = rvcr.upsample_protect(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(pf).cuda(), 0.1)
    assert np.array_equal(got.cpu().numpy(), oracle.upsample_protect(a, b, pf, 0.1))


@pytest.mark.gpu
@pytest.mark.parametrize("rname,pname", CHAINED)
@pytest.mark.parametrize("route", ["default", "scan_tc", "exact_scan_gather", "scan_tc_unfused"])
def test_gpu_retrieve_blend_upsample_fused(rvcr, oracle, golden, tmp_path, rname, pname, route):
    """Retrieval + blend + upsample + protect in one call == the reference statements chained (goldens)."""
    import torch
    r, p = golden(rname), golden(pname)
    path = str(tmp_path / "o.index")
    oracle.write_index(oracle.OracleIndex.from_vectors(r["x"], r["centroids"]).to_data(), path)
    index = rvcr.read_index(path)
    if route == "scan_tc":
        index.set_option("scan_impl", 5)
        index.set_option("fused_blend", 1)
    elif route == "scan_tc_unfused":
        index.set_option("scan_impl", 5)
        index.set_option("fused_blend", 0)
    elif route == "exact_scan_gather":
        index.set_option("scan_impl", 2)
    q = torch.from_numpy(r["q"].copy()).cuda().unsqueeze(0)
    pitchf = torch.from_numpy(p["pitchf"]).cuda().unsqueeze(0)
    rate, protect = float(r["index_rate"]), float(p["protect"])
    out, D, I = index.retrieve_blend_upsample(q, rate, pitchf=pitchf, protect=protect, want_neighbours=True)
    assert out.shape == (1, 2 * q.shape[1], q.shape[2]) and out.data_ptr() != q.data_ptr()
    assert np.array_equal(q[0].cpu().numpy(), r["q"])                                   # feats0 role: untouched
    assert np.array_equal(I.cpu().numpy(), r["ix"]) and np.array_equal(D.cpu().numpy(), r["score"])
    assert np.array_equal(out[0].cpu().numpy(), p["out_mix"], equal_nan=True)
    # no pitch guidance / protect >= 0.5: upsample of the blended rows only
    for kw in (dict(pitchf=None, protect=protect), dict(pitchf=pitchf, protect=0.5)):
        up = index.retrieve_blend_upsample(q, rate, **kw)
        assert np.array_equal(up[0].cpu().numpy(), p["out_up"], equal_nan=True)
    # the one-call host helper mirrors pipeline.py:455-481; without an index only the post step runs
    got = rvcr.convert_features(q, index, rate, pitchf=pitchf, protect=protect)
    assert np.array_equal(got[0].cpu().numpy(), p["out_mix"], equal_nan=True)
    noidx = rvcr.convert_features(q, None, rate, pitchf=pitchf, protect=protect)
    assert np.array_equal(noidx[0].cpu().numpy(), oracle.upsample_protect(r["q"], r["q"], p["pitchf"], protect))
    # and the plain in-place call on the same handle still behaves (post state does not leak)
    f = q.clone()
    index.retrieve_blend_(f, rate)
    assert np.array_equal(f[0].cpu().numpy(), r["out"], equal_nan=True)
    with pytest.raises(AssertionError):
        index.retrieve_blend_upsample(q, rate, pitchf=pitchf[:, :-2], protect=protect)


@pytest.mark.gpu
def test_gpu_fused_post_on_c1(rvcr, oracle):
    """C1-sized synthetic problem: fused call vs oracle retrieve + oracle post (ids exact, features bit-exact)."""
    import torch
    from importlib import import_module
    synth = import_module("codename-rvc-fork-3_b200.synth")
    prob = synth.make_host_problem(50_000, 768, 500, 1282)
    x, cent, q = prob.x, prob.centroids, prob.q
    oidx = oracle.OracleIndex.from_vectors(x, cent)
    index = rvcr.index_factory(768, f"IVF{len(cent)},Flat")
    index.set_centroids(cent)
    for i in range(0, len(x), 8192):                      # extract_index.py:67-69
        index.add(x[i:i + 8192])
    rng = np.random.default_rng(11)
    pf = np.where(rng.random(2 * len(q)) < 0.3, 0.0, rng.uniform(60.0, 500.0, 2 * len(q))).astype(np.float32)
    with np.errstate(all="ignore"):
        blended, score, ix = oracle.retrieve(q, oidx, x, 0.75)
        want = oracle.upsample_protect(blended, q, pf, 0.33)
    out, D, I = index.retrieve_blend_upsample(torch.from_numpy(q).cuda(), 0.75, pitchf=torch.from_numpy(pf).cuda(),
                                              protect=0.33, want_neighbours=True)
    assert np.array_equal(I.cpu().numpy(), ix) and np.array_equal(D.cpu().numpy(), score)
    assert np.array_equal(out.cpu().numpy(), want, equal_nan=True)
