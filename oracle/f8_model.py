"""numpy model of the e4m3 operand arithmetic of coarse_mode 4 (csrc/coarse_tc.cu, experimental).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used by tests to check the kernel's raw values and, on the CPU, that the
error margin the exact re-check relies on really bounds the e4m3 error.  `e4m3_round` is verified against torch's
float8_e4m3fn conversion in tests/test_f8_margin_model.py."""
import numpy as np


def e4m3_round(x):
    """float64 value of round-to-nearest-even e4m3 (bias 7, 3 mantissa bits, subnormal step 2^-9, max 448, saturating)."""
    x = np.asarray(x, np.float64)
    a = np.abs(x)
    e = np.floor(np.log2(np.maximum(a, 2.0 ** -20)))
    e = np.maximum(e, -6.0)                         # below 2^-6: the subnormal grid of the lowest binade
    step = 2.0 ** (e - 3)
    r = np.round(a / step) * step                   # np.round: half to even, like the hardware conversion
    r = np.minimum(r, 448.0)
    return np.sign(x) * r


def f8_scale(mx):
    """power-of-two scale s with 128 <= mx * s < 256 (f8_scale_for in coarse_tc.cu)"""
    m, e = np.frexp(mx)
    return np.ldexp(1.0, 8 - e)


def coarse_values(q, cent):
    """(v8, vtrue, margin): v8 = what the e4m3 tensor-core pass computes (exact dot of the scaled e4m3 operands, float64),
    vtrue = ||c||^2 - 2 <q, c> in float64, margin = the per-query re-check margin of coarse_tc() for coarse_mode 4."""
    q64, c64 = np.asarray(q, np.float64), np.asarray(cent, np.float64)
    d = q64.shape[1]
    cn = (c64 ** 2).sum(1)
    sc = f8_scale(np.abs(c64).max())
    sq = f8_scale(np.abs(q64).max(axis=1))
    q8 = e4m3_round(q64 * sq[:, None])
    c8 = e4m3_round(c64 * sc)
    v8 = cn[None, :] - 2.0 * (q8 @ c8.T) / (sq[:, None] * sc)
    vtrue = cn[None, :] - 2.0 * (q64 @ c64.T)
    qn2 = (q64 ** 2).sum(1)
    cmax = np.sqrt(cn.max())
    rel = 0.53 + d / 2097152
    margin = rel * np.sqrt(qn2) * cmax + (qn2 + cmax * cmax) / 1048576
    return v8, vtrue, margin


def bf16_round(x):
    """float64 value of round-to-nearest-even bfloat16 (torch's conversion)."""
    import torch
    return torch.from_numpy(np.asarray(x, np.float32)).to(torch.bfloat16).to(torch.float64).numpy()


def rigorous_margin(q, cent, kind):
    """numpy restatement of tc_margin / make_margin (csrc/coarse_tc.cu) for kind = "bf16" (coarse_mode 2) or "e4m3" (coarse_mode
    4): (v_lp, vtrue, margin) with v_lp = ||c||^2 - 2 <q', c'> from the low-precision operands q', c' (products and sums exact,
    float64), vtrue the float64 value, and margin(q) built from the EXACT residual norms r = ||x - x'||:
        E = r_q c_max + (|q| + r_q) max_j r_c ;  margin = 4.008 E + acc_rel (|q| + r_q)(c_max + r_c) + abs_rel (|q|^2 + c_max^2).
    The kernel's contract is |v_lp - vtrue| <= margin / 2 for every (query, centroid) pair."""
    q64, c64 = np.asarray(q, np.float64), np.asarray(cent, np.float64)
    d = q64.shape[1]
    if kind == "bf16":
        ql, cl = bf16_round(q64), bf16_round(c64)
        acc_rel = 1.02 * d / 2097152
    elif kind == "e4m3":
        sc = f8_scale(np.abs(c64).max())
        sq = f8_scale(np.abs(q64).max(axis=1))
        ql = e4m3_round(q64 * sq[:, None]) / sq[:, None]
        cl = e4m3_round(c64 * sc) / sc
        acc_rel = max(1.02 * d / 2097152, 1.0 / 512)
    else:
        raise ValueError(kind)
    abs_rel = 1.01 * (4 * (d // 8 + 5) + 2 * (d // 32 + 5) + 8) / 16777216
    cn = (c64 ** 2).sum(1)
    qn = np.sqrt((q64 ** 2).sum(1))
    cmax = np.sqrt(cn.max())
    rq = np.sqrt(((q64 - ql) ** 2).sum(1))
    rc = np.sqrt(((c64 - cl) ** 2).sum(1)).max()
    e = rq * cmax + (qn + rq) * rc
    margin = 4.008 * e + acc_rel * (qn + rq) * (cmax + rc) + abs_rel * (qn * qn + cmax * cmax)
    v_lp = cn[None, :] - 2.0 * (ql @ cl.T)
    vtrue = cn[None, :] - 2.0 * (q64 @ c64.T)
    return v_lp, vtrue, margin


def _f32_bits(x):
    return np.asarray(x, np.float32).view(np.uint32)


def tf32_rna(x):
    """cvt.rna.tf32.f32: round to nearest, ties away from zero, to 10 mantissa bits (low 13 bits cleared); float64 value."""
    b = _f32_bits(x).astype(np.uint64)
    r = ((b + 0x1000) & 0xFFFFE000).astype(np.uint32)       # add half an ulp of the kept part to the magnitude bits, truncate
    return r.view(np.float32).astype(np.float64)


def tf32_trunc(x):
    """what the tensor core keeps of a raw fp32 operand: the top 19 bits (sign, exponent, 10 mantissa bits); float64 value."""
    return (_f32_bits(x) & np.uint32(0xFFFFE000)).view(np.float32).astype(np.float64)


def scan_margin(q, rows):
    """numpy restatement of the tensor-core list scan's approximate values and margin (csrc/scan_tc.cu): the stored rows enter
    as y_hi (top 19 bits) + y_lo (the exact remainder, of which the tensor core again keeps the top 19 bits), the queries as
    tf32(q) rounded to nearest; v = ||y||^2 - 2 <tf32(q), y_hi + y_lo'>.  margin(q) = 4.04 y_max (||q - tf32(q)|| + acc_rel ||q||)
    + abs_margin (||q||^2 + y_max^2) with acc_rel = 2^-20 + d 2^-21, abs_margin = 2^-19 (scan_tc_range).  A row with
    v <= U + margin is a candidate, U >= the 8th smallest v: the contract is 2 |v - vtrue| <= margin / 2 ... i.e. two
    approximate values further apart than margin are ordered like the true ones.  Returns (v, vtrue, margin)."""
    q64, y64 = np.asarray(q, np.float64), np.asarray(rows, np.float64)
    d = q64.shape[1]
    qt = tf32_rna(q)
    yhi = tf32_trunc(rows)
    ylo = tf32_trunc((y64 - yhi).astype(np.float32))      # y - y_hi is exact in fp32
    yn = (y64 ** 2).sum(1)
    v = yn[None, :] - 2.0 * (qt @ (yhi + ylo).T)
    vtrue = yn[None, :] - 2.0 * (q64 @ y64.T)
    qn2 = (q64 ** 2).sum(1)
    ql = np.sqrt(((q64 - qt) ** 2).sum(1))
    ymax = np.sqrt(yn.max())
    acc_rel = 1.0 / 1048576 + d / 2097152
    margin = 4.04 * ymax * (ql + acc_rel * np.sqrt(qn2)) + (qn2 + ymax * ymax) / 524288
    return v, vtrue, margin
