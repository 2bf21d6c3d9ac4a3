"""CPU: the oracle's restatement of the v1 query producer (pipeline.py:451-453) against the golden produced by the reference's
own statement (tests/golden/make_golden_final_proj.py)."""
import numpy as np


def test_oracle_final_proj_matches_reference_statement(oracle, golden):
    g = golden("final_proj_v1")
    got = oracle.final_proj(g["x"], g["weight"], g["bias"])
    assert got.dtype == np.float32 and got.shape == g["proj"].shape == (len(g["x"]), 256)
    # tolerance of this floating-point step: 1e-5 relative to the row scale (BLAS summation orders differ)
    scale = np.abs(g["proj"]).max(axis=1, keepdims=True)
    assert (np.abs(got - g["proj"]) <= 1e-5 * scale).all()
    assert np.array_equal(oracle.final_proj(g["x"], g["weight"]), (g["x"] @ g["weight"].T).astype(np.float32))


def test_reference_chain_on_projected_rows(oracle, golden):
    """projection -> retrieval, all by reference statements: the golden's `out` is what the oracle retrieval gives on `proj`."""
    g, z = golden("final_proj_v1"), golden("retrieve_v1_256")
    oidx = oracle.OracleIndex.from_vectors(z["x"], z["centroids"])
    with np.errstate(all="ignore"):
        out, _, _ = oracle.retrieve(g["proj"], oidx, z["x"], float(g["index_rate"]))
    assert np.array_equal(out, g["out"], equal_nan=True)
