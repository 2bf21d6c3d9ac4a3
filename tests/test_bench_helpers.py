"""CPU: the pure helpers of bench.py -- roofline assembly from phase times (checked on the committed round-1 C5 / C1
lines under profiles/), peak selection (sustained vs burst cuBLAS figure), clock-sample parsing."""
import importlib.util
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


PK = {"hbm_gbs": 6547.5, "bf16_tflops": 1658.4, "bf16_tflops_sustained": 1390.6, "source": "measured"}


def test_roofline_from_committed_c5_line(bench):
    j = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_c5_n1_v5.json")))
    roof, dom = bench.build_roofline(j["phase_ms"], True, PK, nst=32_000_000, n_home=1_000_000, owned_q=1e6, nlist=90509,
                                     d=768, scan_impl=0, workload="C5", world=1)
    assert dom == "coarse" and set(roof) == {"coarse", "scan", "scan_stream", "scan_tail"}
    c = roof["coarse"]
    assert c["bound"] == "tensor" and c["flops_per_launch"] == 2.0 * 1e6 * 90509 * 768
    assert abs(c["achieved"] - j["roofline"]["achieved"]) < 1e-6          # same arithmetic as the committed line
    # a 105 ms launch inside back-to-back steps is compared with the sustained cuBLAS figure; both fractions reported
    assert c["peak"] == PK["bf16_tflops_sustained"] and "sustained" in c["peak_source"]
    assert abs(c["frac"] - c["achieved"] / c["peak"]) < 1e-12
    assert abs(c["frac_vs_burst"] - j["roofline"]["frac"]) < 1e-9 and c["frac_vs_sustained"] == c["frac"]
    s, ss, st = roof["scan"], roof["scan_stream"], roof["scan_tail"]
    assert abs(s["achieved"] - j["roofline_all"]["scan"]["achieved"]) < 1e-6
    assert s["bytes_per_launch"] == ss["bytes_per_launch"] + st["bytes_per_launch"]
    assert abs(ss["ms"] + st["ms"] - s["ms"]) < 1e-9
    assert ss["bytes_per_launch"] == 32_000_000 * (4 * 768 + 8) + 1e6 * 4 * 768 + 1e6 * 8 * 12       # B_s, SURVEY 8d
    assert st["bytes_per_launch"] == 1e6 * (8 * 4 * 768 + 8 * 768) + 1e6 * 8 * 12                    # B_g
    for r in roof.values():
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["unit"] in ("GB/s", "TFLOP/s")
    # ncu traffic of the committed capture is attached per entry at N = 1 only
    tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["C5"]
    assert roof["coarse"]["traffic"] == tr["coarse"] and roof["scan"]["traffic"] == tr["scan"]
    assert roof["scan_stream"]["traffic"] + roof["scan_tail"]["traffic"] == pytest.approx(tr["scan"], rel=1e-6)
    assert 100e9 < tr["scan_stream"] < 120e9 and 55e9 < tr["scan_tail"] < 75e9 and tr["coarse"] < 20e9
    roof2, _ = bench.build_roofline(j["phase_ms"], True, PK, nst=16_000_000, n_home=500_000, owned_q=5e5, nlist=90509,
                                    d=768, scan_impl=0, workload="C5", world=2)
    assert all(r["traffic"] is None for r in roof2.values())


def test_roofline_small_and_exact_paths(bench):
    # C1-sized launch (tens of microseconds): burst peak; exact FFMA coarse: FP32 micro-peak; separate blend kernel
    ph = {"coarse_main": 0.0413, "scan": 0.0681, "blend": 0.0087}
    roof, dom = bench.build_roofline(ph, True, PK, nst=50_000, n_home=500, owned_q=500.0, nlist=1282, d=768, scan_impl=4,
                                     workload="C1", world=1)
    assert dom == "scan" and set(roof) == {"coarse", "scan", "blend"}
    assert roof["coarse"]["peak"] == PK["bf16_tflops"] and "burst" in roof["coarse"]["peak_source"]
    assert roof["scan"]["kernel"] == "scan_topk_kernel_v4" and roof["scan"]["bytes_note"].startswith("B_s (")
    roof, _ = bench.build_roofline(ph, False, PK, nst=50_000, n_home=500, owned_q=500.0, nlist=1282, d=768, scan_impl=4,
                                   workload="C1", world=1)
    assert roof["coarse"]["bound"] == "fp32-ffma" and roof["coarse"]["peak"] == 72.5 and "frac_vs_burst" not in roof["coarse"]
    assert bench.build_roofline({}, True, PK, nst=1, n_home=1, owned_q=1.0, nlist=1, d=128, scan_impl=0, workload="C1",
                                world=1) == ({}, None)


def test_clock_sampler_parsing(bench):
    s = bench.ClockSampler(0)
    s.proc = type("P", (), {"terminate": lambda self: None})()
    s.lines = [(10.00, "2026/10/17 12:00:00.000, 0, 1965, 1965, 300.1, 0x0, Not Active, Not Active, Not Active, Not Active"),
               (10.10, "2026/10/17 12:00:00.100, 0, 1740, 1965, 990.0, 0x4, Not Active, Not Active, Not Active, Active"),
               (10.12, "2026/10/17 12:00:00.120, 0, 1710, 1965, 995.0, 0x4, Not Active, Not Active, Not Active, Active"),
               (10.14, "garbage line"),
               (10.30, "2026/10/17 12:00:00.300, 0, 1965, 1965, 310.0, 0x0, Not Active, Not Active, Not Active, Not Active")]
    s.t0, s.t1 = 10.05, 10.2
    out = s.stop()
    assert out["sm_mhz"] == 1725.0 and out["sm_max_mhz"] == 1965.0 and out["samples"] == 2
    assert out["reasons"] == ["sw_power_cap"] and out["window"] == "timed region"
    s.t0, s.t1 = 10.11, 10.115                       # region shorter than the polling: warm-up samples are used, and said so
    out = s.stop()
    assert out["samples"] == 3 and out["window"].startswith("warm-up")


def test_peaks_reads_driver_file_or_falls_back(bench, tmp_path, monkeypatch):
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    p = bench.peaks()
    assert p["source"] == "fallback" and p["hbm_gbs"] > 0 and p["bf16_tflops_sustained"] <= p["bf16_tflops"]
    (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps({"hbm_gbs": 6500, "bf16_tflops": 1600}))
    p = bench.peaks()
    assert p == {"hbm_gbs": 6500.0, "bf16_tflops": 1600.0, "bf16_tflops_sustained": 1600.0, "source": "measured"}


def test_roofline_e4m3_coarse_uses_twice_the_bf16_figure(bench):
    ph = {"coarse_main": 55.0, "scan": 32.0, "scan_recheck": 12.0}
    roof, dom = bench.build_roofline(ph, True, PK, nst=32_000_000, n_home=1_000_000, owned_q=1e6, nlist=90509, d=768,
                                     scan_impl=0, workload="C5", world=2, coarse_e4m3=True)
    c = roof["coarse"]
    assert dom == "coarse" and "e4m3" in c["kernel"] and "x 2" in c["peak_source"]
    assert c["peak"] == 2 * PK["bf16_tflops_sustained"] and abs(c["frac"] - c["achieved"] / c["peak"]) < 1e-12
    assert c["frac"] < 1.0 and abs(c["frac_vs_burst"] - c["achieved"] / (2 * PK["bf16_tflops"])) < 1e-12


def test_parity_sample_counts_and_never_raises(bench):
    import numpy as np
    rng = np.random.default_rng(0)
    c = rng.standard_normal((50, 16)).astype(np.float32)
    g = c.copy()
    c[3] = np.nan
    g[3] = np.nan                                    # NaN rows (exact-match queries) count as identical
    g[7] += np.float32(0.5)                          # a near-tie row: different centroid, different neighbours
    out = bench.parity_sample(g, c)
    assert out["rows"] == 50 and out["rows_bit_identical"] == 49 and out["rows_differing"] == 1
    assert out["max_rel_diff_in_differing_rows"] > 0
    assert bench.parity_sample(None, c) is None
    assert "error" in bench.parity_sample("nonsense", c) or bench.parity_sample("nonsense", c) is not None
    assert bench.parity_sample(g[:10], c)["rows"] == 10
