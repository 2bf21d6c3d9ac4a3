"""GPU: `model.final_proj(feats[0])` as the query producer (rvc/infer/pipeline.py:451-453) -- the fp32 FFMA kernel
against the golden produced by the reference's own statement (within 1e-5 relative: a floating-point step whose summation
order a GEMM library does not specify), against torch's fp32 linear on the same GPU, and chained with retrieval + blend in
one call (everything after the projection bit-exact given the projected rows)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_final_proj_matches_reference_statement(rvcr, golden):
    import torch
    g = golden("final_proj_v1")
    x, w, b = (torch.from_numpy(g[k]).cuda() for k in ("x", "weight", "bias"))
    got = rvcr.final_proj(x, w, b).cpu().numpy()
    scale = np.abs(g["proj"]).max(axis=1, keepdims=True)
    assert got.shape == g["proj"].shape and (np.abs(got - g["proj"]) <= 1e-5 * scale).all()
    assert np.array_equal(rvcr.final_proj(x.unsqueeze(0), w, b)[0].cpu().numpy(), got)          # (1, n, d) accepted
    # plain fp32 torch reference of the same op on the same GPU (TF32 off)
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = torch.nn.functional.linear(x, w, b).cpu().numpy()
    assert (np.abs(got - ref) <= 1e-5 * scale).all()
    # shapes: ragged n, no bias, larger n (several CTAs per column block)
    rng = np.random.default_rng(3)
    for n in (1, 63, 65, 3000):
        xx = torch.from_numpy(rng.standard_normal((n, 768)).astype(np.float32)).cuda()
        r = torch.nn.functional.linear(xx, w).cpu().numpy()
        o = rvcr.final_proj(xx, w).cpu().numpy()
        assert (np.abs(o - r) <= 1e-5 * np.abs(r).max(axis=1, keepdims=True)).all(), n
    with pytest.raises(AssertionError):
        rvcr.final_proj(x, w[:, :700].contiguous(), b)


def test_final_proj_retrieve_blend_chain(rvcr, oracle, golden, tmp_path):
    import torch
    g, z = golden("final_proj_v1"), golden("retrieve_v1_256")
    oidx = oracle.OracleIndex.from_vectors(z["x"], z["centroids"])
    path = str(tmp_path / "v1.index")
    oracle.write_index(oidx.to_data(), path)
    index = rvcr.read_index(path)
    x, w, b = (torch.from_numpy(g[k]).cuda() for k in ("x", "weight", "bias"))
    rate = float(g["index_rate"])
    out, D, I = index.final_proj_retrieve_blend(x.unsqueeze(0), w, b, rate, want_neighbours=True)
    torch.cuda.synchronize()
    assert out.shape == (1, len(g["x"]), 256)
    # given OUR projected rows the rest of the chain is the bit-exact path
    proj = rvcr.final_proj(x, w, b).cpu().numpy()
    with np.errstate(all="ignore"):
        want, Do, Io = oracle.retrieve(proj, oidx, z["x"], rate)
    assert np.array_equal(I.cpu().numpy(), Io) and np.array_equal(D.cpu().numpy(), Do)
    assert np.array_equal(out[0].cpu().numpy(), want, equal_nan=True)
    # and against the all-reference chain (projection by the reference statement): same neighbours wherever the two
    # projections give the same ids, blended rows within the 1e-5 relative budget of the north star there
    _, Ig = oracle.retrieve(g["proj"], oidx, z["x"], rate)[1:]
    same = (Io == Ig).all(axis=1)
    assert same.mean() > 0.95
    ref = g["out"][same]
    got = out[0].cpu().numpy()[same]
    fin = np.isfinite(ref).all(axis=1)
    assert (np.abs(got[fin] - ref[fin]) <= 1e-4 * np.abs(ref[fin]).max(axis=1, keepdims=True)).all()
    assert np.array_equal(x.cpu().numpy(), g["x"])                       # the encoder rows are left untouched
