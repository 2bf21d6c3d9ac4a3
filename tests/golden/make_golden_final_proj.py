"""Generate tests/golden/final_proj_v1.npz by executing the REFERENCE's own statement for the step before the retrieval
path of a v1 model, rvc/infer/pipeline.py:451-453:

    feats = (model.final_proj(feats[0]).unsqueeze(0) if version == "v1" else feats)

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_golden_final_proj.py

The assignment is cut out of `Pipeline.voice_conversion` with `ast`, compiled unchanged and executed with torch (CPU,
fp32) on the real ContentVec rows the reference ships (logs/reference/ref_feats.npy, 169 x 768) with a seeded
nn.Linear(768, 256) standing in for the encoder's projection (the checkpoint is not in the repository).  The projected
rows are then the queries of the committed v1 retrieval case: the golden also holds the output of the reference's
`_retrieve_speaker_embeddings` body (cut out the same way, see make_golden.py) on them against that case's index."""
from __future__ import annotations

import ast
import hashlib
import os
import sys
import textwrap

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))


def reference_statement() -> str:
    src = open(os.path.join(REF, "rvc/infer/pipeline.py")).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "voice_conversion":
            body = next(n for n in node.body if isinstance(n, ast.With)).body
            st = next(st for st in body if isinstance(st, ast.Assign) and "final_proj" in ast.get_source_segment(src, st))
            return textwrap.dedent("\n".join(src.splitlines()[st.lineno - 1: st.end_lineno]))
    raise RuntimeError("voice_conversion not found")


def reference_retrieve():
    src = open(os.path.join(REF, "rvc/infer/pipeline.py")).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "_retrieve_speaker_embeddings":
            return textwrap.dedent("\n".join(src.splitlines()[node.lineno - 1: node.end_lineno]))
    raise RuntimeError("_retrieve_speaker_embeddings not found")


class Model:
    pass


def main():
    import oracle
    stmt = reference_statement()
    x = np.load(os.path.join(REF, "logs/reference/ref_feats.npy")).astype(np.float32)
    torch.manual_seed(20261017)
    lin = torch.nn.Linear(768, 256)
    model = Model()
    model.final_proj = lin
    ns = {"model": model, "version": "v1", "feats": torch.from_numpy(x).unsqueeze(0)}
    with torch.no_grad():
        exec(compile(stmt, "reference:rvc/infer/pipeline.py:451-453", "exec"), ns)
    proj = ns["feats"]
    assert proj.shape == (1, x.shape[0], 256) and proj.dtype == torch.float32
    ns2 = dict(ns, version="v2", feats=torch.from_numpy(x).unsqueeze(0))
    exec(compile(stmt, "reference", "exec"), ns2)
    assert ns2["feats"].shape == (1, x.shape[0], 768)            # v2: untouched
    proj = proj[0].numpy().copy()
    # retrieval of the projected rows against the committed v1 case's index, through the reference's own function body
    z = np.load(os.path.join(OUT, "retrieve_v1_256.npz"))
    oidx = oracle.OracleIndex.from_vectors(z["x"], z["centroids"])

    class Idx:
        def search(self, npy, k=8):
            return oidx.search(npy, k=k)

    glb = {"np": np, "torch": torch}
    exec(compile(reference_retrieve(), "reference:_retrieve_speaker_embeddings", "exec"), glb)

    class Self:
        device = "cpu"

    rate = 0.75
    with np.errstate(all="ignore"):
        out = glb["_retrieve_speaker_embeddings"](Self(), torch.from_numpy(proj).unsqueeze(0), Idx(), z["x"], rate)
    np.savez_compressed(os.path.join(OUT, "final_proj_v1.npz"), x=x, weight=lin.weight.detach().numpy(), bias=lin.bias.detach().numpy(),
                        proj=proj, index_rate=np.float64(rate), out=out[0].numpy())
    with open(os.path.join(OUT, "reference_final_proj_statement.txt"), "w") as f:
        f.write("sha256 of the statement executed (rvc/infer/pipeline.py:451-453): " + hashlib.sha256(stmt.encode()).hexdigest() + "\n")
    print("final_proj golden:", proj.shape, "blend", out.shape, "finite", bool(np.isfinite(out.numpy()).all()))


if __name__ == "__main__":
    main()
