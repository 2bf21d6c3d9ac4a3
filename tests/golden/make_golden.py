"""Generate tests/golden/*.npz by executing the REFERENCE's own function body.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_golden.py

What is pinned: `Pipeline._retrieve_speaker_embeddings` (rvc/infer/pipeline.py:497-507) is cut out of
the reference source with `ast`, compiled unchanged, and executed with numpy + torch (CPU) on real
ContentVec fixtures from the reference's logs/ directory.  faiss itself cannot be imported here
(un-vendored faiss-cpu==1.7.3, no wheel, no network), so the `index` object handed to the reference
function is a stand-in whose `.search(npy, k=8)` returns the oracle's (score, ix): the numpy/torch
tail of the path is therefore pinned by the reference's own code, the faiss part is not
("parity unpinned").
"""
from __future__ import annotations

import ast
import os
import sys
import textwrap

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def reference_function():
    src = open(os.path.join(REF, "rvc/infer/pipeline.py")).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "_retrieve_speaker_embeddings":
            fn_src = textwrap.dedent(ast.get_source_segment(src, node))
            ns = {"np": np, "torch": torch}
            exec(compile(fn_src, "reference:rvc/infer/pipeline.py", "exec"), ns)
            return ns["_retrieve_speaker_embeddings"], fn_src
    raise RuntimeError("reference function not found")


class FakeSelf:
    device = torch.device("cpu")


class OracleBackedIndex:
    """Stand-in for the faiss index: search() answers with the oracle (keyword k as at pipeline.py:499)."""

    def __init__(self, oidx):
        self.oidx = oidx
        self.last = None

    def search(self, npy, k=8):
        self.last = self.oidx.search(npy, k=k)
        return self.last


def make_case(name, x, centroids, q, index_rate):
    fn, fn_src = reference_function()
    oidx = oracle.OracleIndex.from_vectors(x, centroids)
    big_npy = oidx.reconstruct_n(0, oidx.ntotal)
    assert np.array_equal(big_npy, x)
    fake = OracleBackedIndex(oidx)
    feats = torch.from_numpy(q.copy()).unsqueeze(0)
    with np.errstate(all="ignore"):
        out = fn(FakeSelf(), feats, fake, big_npy, index_rate)
    score, ix = fake.last
    assign, cdist = oracle.coarse_assign(q, centroids)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), x=x, centroids=centroids, q=q,
                        index_rate=np.float64(index_rate), score=score, ix=ix, assign=assign, coarse_dist=cdist,
                        out=out[0].numpy(), list_off=oidx.list_off, list_ids=oidx.ids)
    nan_rows = int(np.isnan(out[0].numpy()).any(axis=1).sum())
    pad_rows = int((ix < 0).any(axis=1).sum())
    print(f"{name}: N={x.shape[0]} d={x.shape[1]} nq={q.shape[0]} nlist={centroids.shape[0]} "
          f"nan_rows={nan_rows} pad_rows={pad_rows}")
    return fn_src


def main():
    rng = np.random.default_rng(20261017)
    mute = np.load(os.path.join(REF, "logs/mute/extracted/mute.npy")).astype(np.float32)
    spin = np.load(os.path.join(REF, "logs/mute_spin/extracted/mute.npy")).astype(np.float32)
    ref = np.load(os.path.join(REF, "logs/reference/ref_feats.npy")).astype(np.float32)

    # case 1: real ContentVec rows (with their exact duplicates) + synthetic rows, d = 768
    synth = (ref[rng.integers(0, len(ref), 90)] + rng.standard_normal((90, 768)).astype(np.float32) * 0.1)
    x = np.concatenate([mute[:100], spin[:100], ref[:60], synth]).astype(np.float32)
    x = x[rng.permutation(len(x))]
    nlist = 24
    cent = x[rng.choice(len(x), nlist, replace=False)].copy()
    a, _ = oracle.coarse_assign(x, cent)
    for l in range(nlist):
        if (a == l).any():
            cent[l] = x[a == l].mean(axis=0)
    # queries: the reference preview features, some rows identical to indexed rows (score 0 -> NaN row)
    q = np.concatenate([ref[60:], x[:3], mute[100:120]]).astype(np.float32)
    src = make_case("retrieve_v2_768", x, cent.astype(np.float32), q, 0.75)

    # case 2: v1-style 256-d index with many tiny lists (padding with -1 / FLT_MAX), index_rate 0.3
    x2 = (rng.standard_normal((300, 256)) * 0.33).astype(np.float32)
    x2[10] = x2[11]  # exact duplicate rows -> equal distances
    cent2 = x2[rng.choice(300, 60, replace=False)].copy()
    q2 = (x2[rng.integers(0, 300, 64)] + rng.standard_normal((64, 256)).astype(np.float32) * 0.05).astype(np.float32)
    make_case("retrieve_v1_256", x2, cent2, q2, 0.3)

    with open(os.path.join(OUT, "reference_function.txt"), "w") as f:
        f.write("# extracted verbatim from /root/reference/rvc/infer/pipeline.py by make_golden.py\n")
        f.write("# (kept only as a provenance record of what was executed; not imported by anything)\n")
        f.write("# sha256 of the function source: ")
        import hashlib
        f.write(hashlib.sha256(src.encode()).hexdigest() + "\n")


if __name__ == "__main__":
    main()
