"""CPU: the numerics of the experimental e4m3 coarse step (coarse_mode 4, csrc/coarse_tc.cu) on a numpy model.

The kernel itself has not run on a B200 yet; what CAN be checked here is the claim its exact re-check relies on:
|v_e4m3 - v_true| <= margin / 2 for every (query, centroid) pair -- on the bench's data construction, on badly scaled
rows, and on adversarial rows that sit on the e4m3 rounding boundaries -- and that the candidate set {v <= v_min + margin}
therefore always contains the true nearest centroid while staying small on the bench's data."""
import numpy as np
import pytest
import torch


@pytest.fixture(scope="module")
def model():
    from oracle import f8_model
    return f8_model


def test_e4m3_model_equals_torch_conversion(model):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(200_000) * 100, rng.standard_normal(100_000) * 0.01,
                        [0, 448, 449, 460, 500, -500, 2.0 ** -9, 2.0 ** -10, 1.5 * 2.0 ** -9, 2.0 ** -6, 0.0175, 255.9, 17.0]])
    x = x.astype(np.float32)
    want = torch.from_numpy(np.clip(x, -448, 448)).to(torch.float8_e4m3fn).to(torch.float64).numpy()
    assert np.array_equal(model.e4m3_round(x), want)


def bench_like(rng, nq, nlist, d, sib=0.03):
    blobs = (rng.standard_normal((max(4, nlist // 4), d)) * 0.30).astype(np.float32)
    cent = (blobs[np.arange(nlist) // 4 % len(blobs)] + rng.standard_normal((nlist, d)).astype(np.float32) * sib).astype(np.float32)
    q = (cent[rng.integers(0, nlist, nq)] + rng.standard_normal((nq, d)).astype(np.float32) * 0.14).astype(np.float32)
    return q, cent


@pytest.mark.parametrize("d", [768, 256])
def test_margin_bounds_the_e4m3_error_on_bench_like_data(model, d):
    rng = np.random.default_rng(d)
    q, cent = bench_like(rng, 400, 2000, d)
    v8, vtrue, margin = model.coarse_values(q, cent)
    err = np.abs(v8 - vtrue)
    assert (err <= 0.5 * margin[:, None]).all(), (err / margin[:, None]).max()
    # the candidate set of the re-check: contains the true argmin, and only the sibling centroids on this data
    cand = v8 <= v8.min(axis=1, keepdims=True) + margin[:, None]
    best = vtrue.argmin(axis=1)
    assert cand[np.arange(len(q)), best].all()
    per_query = cand.sum(axis=1)
    assert per_query.max() <= 8 and per_query.mean() <= 4.5, (per_query.max(), per_query.mean())   # top-8 per split never truncates
    # measured slack: the rigorous bound is ~6x the observed error (Cauchy-Schwarz, worst-case rounding)
    assert (err / margin[:, None]).max() < 0.2


def test_margin_holds_for_badly_scaled_and_adversarial_rows(model):
    rng = np.random.default_rng(5)
    d = 768
    q, cent = bench_like(rng, 64, 512, d)
    q[0] *= np.float32(1e-6)                           # tiny row: own power-of-two scale
    q[1] *= np.float32(1e4)                            # huge row
    q[2, 1:] *= np.float32(1e-4)                       # one dominant element: the rest falls into the subnormal range
    cent[3] *= np.float32(1e-3)                        # a centroid far below the global scale: mostly subnormal / zero in e4m3
    cent[4, 5] = np.float32(np.abs(cent).max() * 3)    # one element sets the global scale, pushing all others down
    # rows sitting exactly on e4m3 rounding midpoints (worst-case relative error 2^-4 on every element)
    mid = np.float32(1.0625)                           # halfway between 1.0 and 1.125 in the binade [1, 2)
    q[5] = mid * np.sign(rng.standard_normal(d)).astype(np.float32) * np.float32(0.25)
    cent[6] = mid * np.sign(rng.standard_normal(d)).astype(np.float32) * np.float32(0.25)
    v8, vtrue, margin = model.coarse_values(q, cent)
    err = np.abs(v8 - vtrue)
    assert (err <= 0.5 * margin[:, None]).all(), np.unravel_index((err / margin[:, None]).argmax(), err.shape)
    cand = v8 <= v8.min(axis=1, keepdims=True) + margin[:, None]
    assert cand[np.arange(len(q)), vtrue.argmin(axis=1)].all()
