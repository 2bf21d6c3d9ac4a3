"""GPU: the reference's OWN statements, unmodified, running against the shim (SURVEY 8b "reference files run unmodified").

tests/golden/reference_snippets/*.txt hold, verbatim, the index-load block of `Pipeline.pipeline` (rvc/infer/pipeline.py:
553-561), `Pipeline._retrieve_speaker_embeddings` (:497-507) and the body of the build script (rvc/train/process/
extract_index.py:14-72) -- cut out of the reference with `ast` by tests/golden/make_dropin_fixture.py, because the
reference tree does not exist on the GPU box.  They are executed here with `faiss` bound to this package (and, for the
"Auto" build, `MiniBatchKMeans` bound to the GPU one): nothing is restated."""
import importlib
import os
import sys
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SNIP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_snippets")


def snippet(name):
    text = open(os.path.join(SNIP, name)).read()
    assert text.startswith("# REFERENCE TEXT")
    return text.split("\n", 1)[1]


class FakeSelf:
    device = "cuda:0"


def test_pipeline_statements_run_unmodified_on_the_shim(rvcr, oracle, golden, tmp_path, capsys):
    import torch
    g = golden("retrieve_v2_768")
    oidx = oracle.OracleIndex.from_vectors(g["x"], g["centroids"])
    path = str(tmp_path / "model.index")
    oracle.write_index(oidx.to_data(), path)
    load = compile(snippet("pipeline_load_block.txt"), "reference:pipeline.py:553-561", "exec")
    fn_ns = {"np": np, "torch": torch}
    exec(compile(snippet("retrieve_speaker_embeddings.txt"), "reference:pipeline.py:497-507", "exec"), fn_ns)
    ref_fn = fn_ns["_retrieve_speaker_embeddings"]
    rate = float(g["index_rate"])

    # (1) plain faiss-shaped objects: read_index, reconstruct_n, truthiness, search(npy, k=8) on a host array
    rvcr.enable_index_cache(False)
    rvcr.IndexIVFFlat.lazy_reconstruct = False
    ns = {"faiss": rvcr, "os": os, "file_index": path, "index_rate": rate}
    exec(load, ns)
    index, big_npy = ns["index"], ns["big_npy"]
    assert index and isinstance(big_npy, np.ndarray) and np.array_equal(big_npy, g["x"])
    feats = torch.from_numpy(g["q"].copy()).cuda().unsqueeze(0)
    with np.errstate(all="ignore"):
        out = ref_fn(FakeSelf(), feats, index, big_npy, rate)               # the reference's own numpy / torch tail
    assert out.shape == feats.shape and out.device == feats.device
    assert np.array_equal(out[0].cpu().numpy(), g["out"], equal_nan=True)
    assert np.array_equal(feats[0].cpu().numpy(), g["q"])                   # reference semantics: the input is not modified

    # (2) the guards of the block: missing file, index_rate 0, unreadable file -> index = big_npy = None ("no retrieval")
    for fi, r in ((str(tmp_path / "nope.index"), rate), (path, 0.0), ("", rate)):
        ns = {"faiss": rvcr, "os": os, "file_index": fi, "index_rate": r}
        exec(load, ns)
        assert ns["index"] is None and ns["big_npy"] is None
    bad = str(tmp_path / "bad.index")
    open(bad, "wb").write(b"IwFl" + b"\0" * 40)
    ns = {"faiss": rvcr, "os": os, "file_index": bad, "index_rate": rate}
    exec(load, ns)
    assert ns["index"] is None and "An error occurred reading the FAISS index" in capsys.readouterr().out

    # (3) patch_pipeline on a module that holds the reference's method: cached device-resident index, lazy big_npy, fused call
    mod = types.ModuleType("fake_rvc_infer_pipeline")
    mod.faiss = None
    mod.Pipeline = type("Pipeline", (), {"_retrieve_speaker_embeddings": ref_fn, "device": "cuda:0"})
    rvcr.patch_pipeline(mod)
    try:
        assert mod.faiss is rvcr and mod.Pipeline._retrieve_speaker_embeddings is not ref_fn
        reads = []
        for _ in range(3):                                                  # pipeline() is called per file / chunk (infer.py:291-311)
            ns = {"faiss": mod.faiss, "os": os, "file_index": path, "index_rate": rate}
            exec(load, ns)
            reads.append(ns["index"])
            feats = torch.from_numpy(g["q"].copy()).cuda().unsqueeze(0)
            out = mod.Pipeline()._retrieve_speaker_embeddings(feats, ns["index"], ns["big_npy"], rate)
            assert np.array_equal(out[0].cpu().numpy(), g["out"], equal_nan=True)
        assert reads[0] is reads[1] is reads[2]                             # one device-resident index, not three file reads
        assert not isinstance(ns["big_npy"], np.ndarray)                    # never materialised on this path
        assert np.array_equal(np.asarray(ns["big_npy"][:5]), g["x"][:5])    # ... but there if somebody indexes it
    finally:
        rvcr.enable_index_cache(False)
        rvcr.IndexIVFFlat.lazy_reconstruct = False
        rvcr.index_cache.clear()


def run_extract_index(rvcr, exp_dir, algo, extra_modules=None):
    body = compile(snippet("extract_index_body.txt"), "reference:extract_index.py:14-72", "exec")
    from multiprocessing import cpu_count
    from sklearn.cluster import MiniBatchKMeans as SkMBK
    ns = {"os": os, "sys": sys, "np": np, "faiss": rvcr, "cpu_count": cpu_count, "MiniBatchKMeans": SkMBK,
          "exp_dir": exp_dir, "index_algorithm": algo}
    ns.update(extra_modules or {})
    exec(body, ns)
    return ns


def test_extract_index_script_runs_unmodified_on_the_shim(rvcr, oracle, tmp_path, capsys):
    """extract_index.py with index_algorithm = "Faiss": concatenate, shuffle, n_ivf, factory, nprobe, train, add in 8192-row
    batches, write -- every faiss call lands in the shim; the file equals what the oracle builds from the same rows."""
    synth = importlib.import_module("codename-rvc-fork-3_b200.synth")
    exp_dir = str(tmp_path / "mymodel")
    os.makedirs(os.path.join(exp_dir, "extracted"))
    rng = np.random.default_rng(3)
    rows, _, _ = synth.mixture_points(rng, 20_000, 768, 60)
    for i, part in enumerate(np.array_split(rows, 5)):
        np.save(os.path.join(exp_dir, "extracted", f"{i:03d}.npy"), part)
    np.random.seed(1234)                                    # the script shuffles with the global numpy generator (:40)
    run_extract_index(rvcr, exp_dir, "Faiss")
    out = os.path.join(exp_dir, "mymodel.index")
    assert os.path.exists(out) and "Saved index file" in capsys.readouterr().out
    # the same rows in the same shuffled order, through the oracle
    np.random.seed(1234)
    idx = np.arange(len(rows))
    np.random.shuffle(idx)
    big_npy = rows[idx]
    n_ivf = min(int(16 * np.sqrt(len(big_npy))), len(big_npy) // 39)
    cent = oracle.kmeans_train(big_npy, n_ivf)[0]
    oidx = oracle.OracleIndex.from_vectors(big_npy, cent)
    ref = str(tmp_path / "ref.index")
    oracle.write_index(oidx.to_data(), ref)
    assert open(out, "rb").read() == open(ref, "rb").read()
    # a second run finds the file and does nothing (the script's own guard)
    before = os.path.getmtime(out)
    run_extract_index(rvcr, exp_dir, "Faiss")
    assert os.path.getmtime(out) == before


def test_extract_index_auto_path_with_gpu_minibatch_kmeans(rvcr, tmp_path, capsys):
    """index_algorithm = "Auto" above 200 000 rows: the script's MiniBatchKMeans(n_clusters=10000, ...) call, bound to the GPU
    estimator, then IVF256 over the 10 000 centres.  (Unseeded in the reference: checked for shape, health and searchability.)"""
    synth = importlib.import_module("codename-rvc-fork-3_b200.synth")
    exp_dir = str(tmp_path / "big")
    os.makedirs(os.path.join(exp_dir, "extracted"))
    rng = np.random.default_rng(4)
    rows, _, _ = synth.mixture_points(rng, 200_500, 768, 400)
    for i, part in enumerate(np.array_split(rows, 4)):
        np.save(os.path.join(exp_dir, "extracted", f"{i}.npy"), part)
    run_extract_index(rvcr, exp_dir, "Auto", {"MiniBatchKMeans": rvcr.MiniBatchKMeans})
    out = os.path.join(exp_dir, "big.index")
    assert os.path.exists(out)
    info = rvcr.probe_index(out)
    assert (info["d"], info["ntotal"], info["nlist"], info["nprobe"]) == (768, 10000, 256, 1)
    index = rvcr.read_index(out)
    big = index.reconstruct_n(0, index.ntotal)
    assert np.isfinite(big).all()
    # the 10 000 centres cover the data: a row's nearest centre is closer than the typical inter-blob distance.  The estimator
    # and the script's shuffle are unseeded (as in the reference), so nothing about individual lists is fixed: a probed list may
    # hold fewer than 8 centres, and then -- and only then -- faiss semantics pad the tail of the row with -1
    D, I = index.search(rows[:2000], k=8)
    sizes = np.asarray(index.list_sizes())
    import torch
    probed = index.assign(torch.from_numpy(np.ascontiguousarray(rows[:2000])).cuda())[0].cpu().numpy().astype(np.int64)
    have = np.minimum(sizes[probed], 8)
    assert np.array_equal((I >= 0).sum(axis=1), have)
    assert ((I >= 0) == (np.arange(8)[None, :] < have[:, None])).all()
    # blob centres are 2 * 768 * 0.30^2 = 138 apart (squared), a row is 768 * 0.14^2 = 15 from a k-means centre of its own
    # blob: the median nearest centre must be a same-blob one (half the inter-blob distance as the line)
    assert (have >= 1).mean() > 0.99 and np.median(D[have >= 1, 0]) < 0.5 * 2 * 768 * 0.30 ** 2
