"""CPU: the C-ABI library builds, loads, exports every symbol include/rvcr.h declares, and fails loudly
(no CPU fallback) when no CUDA device is present.  No compute calls here."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rvcr.h")).read()
    return sorted(set(re.findall(r"RVCR_API\s+[\w\s\*]+?\b(rvcr_\w+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound(rvcr):
    from importlib import import_module
    ffi = import_module("codename-rvc-fork-3_b200._ffi")
    L = rvcr.lib()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/rvcr.h but not exported by librvcr.so"
        assert s in ffi.SIGNATURES, f"{s} has no ctypes signature in _ffi.py"
    assert sorted(ffi.SIGNATURES) == syms
    out = subprocess.run(["nm", "-D", "--defined-only", ffi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (rvcr_\w+)", out))
    assert exported == set(syms)        # nothing else leaks out of the library (hidden visibility)
    assert L.rvcr_version() == 100


def test_header_is_plain_c(tmp_path):
    c = tmp_path / "t.c"
    c.write_text('#include "rvcr.h"\nint main(void){rvcr_info i; rvcr_stats s; (void)i; (void)s; return RVCR_OK;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(c),
                    "-o", str(tmp_path / "t.o")], check=True)


def test_sass_is_sm100a_only():
    from importlib import import_module
    ffi = import_module("codename-rvc-fork-3_b200._ffi")
    out = subprocess.run(["cuobjdump", "-lelf", ffi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu(rvcr):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rvcr.RvcrError) as e:
        rvcr.index_factory(768, "IVF16,Flat")
    assert e.value.code == -4
    with pytest.raises(rvcr.RvcrError):
        rvcr.read_index("/nonexistent.index")
    with pytest.raises(rvcr.RvcrError):
        rvcr.index_factory(768, "PQ16")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "codename-rvc-fork-3_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
                assert "rvc_oracle" not in txt, f


def test_native_kmeans_host_pieces_match_oracle(rvcr, oracle):
    """rvcr_index_train's host-only parts (csrc/kmeans.cu) against the oracle's independent restatement -- no GPU
    needed: faiss rand_perm on std::mt19937 vs numpy's legacy MT19937 stream, and split_clusters bit for bit."""
    import ctypes as C

    import numpy as np
    L = rvcr.lib()
    for n, seed in ((0, 1), (1, 1), (2, 7), (17, 1234), (1000, 1235), (65537, 99)):
        out = np.empty(n, np.int64)
        assert L.rvcr_debug_rand_perm(n, seed, out.ctypes.data_as(C.c_void_p)) == 0
        assert np.array_equal(out, oracle.rand_perm(n, seed)), (n, seed)
    rng = np.random.default_rng(3)
    for k, d, n_empty in ((8, 6, 3), (64, 33, 20), (5, 4, 4)):
        hass = rng.integers(2, 60, k).astype(np.float32)
        hass[rng.choice(k, n_empty, replace=False)] = 0
        n = int(hass.sum())
        cent = rng.standard_normal((k, d)).astype(np.float32)
        h_o, c_o = hass.copy(), cent.copy()
        ns_o = oracle.split_clusters(h_o, c_o, n)
        h_n, c_n = hass.copy(), cent.copy()
        ns = C.c_int32(0)
        rc = L.rvcr_debug_split_clusters(d, k, n, h_n.ctypes.data_as(C.c_void_p), c_n.ctypes.data_as(C.c_void_p), C.byref(ns))
        assert rc == 0 and ns.value == ns_o == n_empty
        assert np.array_equal(h_n, h_o) and np.array_equal(c_n, c_o)
        assert (h_n > 0).all() and abs(float(h_n.sum()) - n) < 1e-3
    bad = np.zeros(4, np.float32)
    assert L.rvcr_debug_split_clusters(2, 4, 10, bad.ctypes.data_as(C.c_void_p), np.zeros((4, 2), np.float32).ctypes.data_as(C.c_void_p), None) == -1
