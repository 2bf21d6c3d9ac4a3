"""CPU: the `.index` parser under AddressSanitizer + UndefinedBehaviorSanitizer.

csrc/index_io.cu is host-only C++, so the very source that ships in librvcr.so is compiled with g++ -fsanitize=address,
undefined (tools/asan_probe_main.cpp stubs the device half, which the probe never reaches) and fed a corpus of valid files
(current and legacy layouts), every truncation of their headers and tables, hostile counts, and thousands of random
mutations.  Any out-of-bounds access, overflow the sanitizers catch, leak, or unexpected return code fails the test --
index files are downloaded from the internet by the reference's own UI (tabs/download/download.py)."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"
HDR = 4 + 8 + 8 + 8 + 1 + 4


@pytest.fixture(scope="module")
def asan_probe(tmp_path_factory):
    gxx = shutil.which("g++")
    if not gxx or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    out = str(tmp_path_factory.mktemp("asan") / "asan_probe")
    cmd = [gxx, "-x", "c++", "-std=c++17", "-g", "-O1", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
           "-I", os.path.join(ROOT, "include"), "-I", CUDA_INC,
           os.path.join(ROOT, "codename-rvc-fork-3_b200", "csrc", "index_io.cu"), os.path.join(ROOT, "tools", "asan_probe_main.cpp"),
           "-L", "/usr/local/cuda/lib64", "-lcudart", "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("sanitizer build unavailable here: " + r.stderr[-300:])
    return out


def run(binary, paths):
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1",
               LD_LIBRARY_PATH="/usr/local/cuda/lib64:" + os.environ.get("LD_LIBRARY_PATH", ""))
    for i in range(0, len(paths), 1500):
        r = subprocess.run([binary] + paths[i:i + 1500], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, (r.returncode, r.stdout[-500:], r.stderr[-3000:])
    return r.stdout


def test_parser_is_clean_under_asan_and_ubsan(asan_probe, oracle, tmp_path):
    rng = np.random.default_rng(17)
    d, nlist, n = 128, 5, 24
    cent = rng.standard_normal((nlist, d)).astype(np.float32)
    x = (cent[rng.integers(0, nlist, n)] + rng.standard_normal((n, d)).astype(np.float32) * 0.1).astype(np.float32)
    data = oracle.OracleIndex.from_vectors(x, cent).to_data()
    blobs = {}
    p = str(tmp_path / "seed.index")
    oracle.write_index(data, p)
    blobs["IwFl"] = open(p, "rb").read()
    for four in (b"IvFL", b"IvFl"):
        oracle.write_index_legacy(data, p, four)
        blobs[four.decode()] = open(p, "rb").read()
    sparse = oracle.OracleIndex.from_vectors(x[:6], np.concatenate([cent, cent + 5, cent + 9])).to_data()   # 'sprs' table
    oracle.write_index(sparse, p)
    blobs["sprs"] = open(p, "rb").read()
    assert b"sprs" in blobs["sprs"]

    paths = []

    def emit(name, blob):
        q = str(tmp_path / name)
        with open(q, "wb") as f:
            f.write(blob)
        paths.append(q)

    for name, blob in blobs.items():
        emit(f"{name}_ok.index", blob)
        emit(f"{name}_trailing.index", blob + b"\0" * 77)
        head = len(blob) - int(np.asarray(data.list_sizes if name != "sprs" else sparse.list_sizes).sum()) * (4 * d + 8)
        for cut in list(range(0, min(head + 64, len(blob)))) + sorted(rng.integers(0, len(blob), 40).tolist()):
            emit(f"{name}_cut{cut}.index", blob[:cut])
        hdr_len = 4 + HDR + 16 + 4 + HDR + 8
        for k in range(700):
            b = bytearray(blob)
            for pos in rng.integers(0, hdr_len, rng.integers(1, 5)):
                b[pos] = rng.integers(0, 256)
            for pos in rng.integers(hdr_len + nlist * d * 4 if name != "sprs" else hdr_len + 15 * d * 4, len(b), rng.integers(0, 6)):
                b[pos] = rng.integers(0, 256)
            emit(f"{name}_mut{k}.index", bytes(b))
        # counts far beyond the file, in every 8-byte field of the header / tables region
        for off in range(4, min(len(blob) - 8, hdr_len + 64), 4):
            for val in (0xFFFFFFFFFFFFFFFF, 1 << 62, 1 << 31, (1 << 63) + 5):
                b = bytearray(blob)
                struct.pack_into("<Q", b, off, val)
                emit(f"{name}_big{off}_{val & 0xFFFF}.index", bytes(b))
    emit("empty.index", b"")
    out = run(asan_probe, paths)
    assert "parsed" in out
    # the valid seeds did parse (sanity: the harness is not refusing everything)
    ok = run(asan_probe, [q for q in paths if q.endswith("_ok.index") or q.endswith("_trailing.index")])
    assert ok.startswith("8 parsed, 0 refused"), ok
