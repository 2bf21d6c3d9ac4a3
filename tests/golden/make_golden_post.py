"""Generate tests/golden/post_*.npz by executing the REFERENCE's own statements for the step after the
retrieval path: 2x upsample + pitch-protect mix, rvc/infer/pipeline.py:463-481.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_golden_post.py

The statements live inline in `Pipeline.voice_conversion`; they are cut out with `ast` -- the
`feats = F.interpolate(...)` assignment, the `p_len = min(...)` line and the whole `if pitch_guidance:` block
-- compiled unchanged and executed with torch (CPU).  Inputs: (a) the blended features / queries of the
committed retrieval goldens (so retrieval + post chains end to end through reference code only), (b) a small
stand-alone case.  pitchf mixes 0 (unvoiced), values in (0, 1), exactly 1 and ordinary F0 values in Hz.
"""
from __future__ import annotations

import ast
import hashlib
import os
import textwrap

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def reference_statements() -> str:
    src = open(os.path.join(REF, "rvc/infer/pipeline.py")).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "voice_conversion":
            with_node = next(n for n in node.body if isinstance(n, ast.With))
            body = with_node.body
            start = next(i for i, st in enumerate(body)
                         if isinstance(st, ast.Assign) and "F.interpolate(feats.permute" in ast.get_source_segment(src, st))
            end = next(i for i, st in enumerate(body) if isinstance(st, ast.If) and
                       ast.get_source_segment(src, st.test) == "pitch_guidance")
            assert end == start + 2, "reference layout changed"
            lines = src.splitlines()
            seg = "\n".join(lines[body[start].lineno - 1: body[end].end_lineno])
            return textwrap.dedent(seg)
    raise RuntimeError("voice_conversion not found")


class FakeSelf:
    window = 160


def run_reference(code, feats, feats0, pitchf, protect, pitch_guidance):
    n = feats.shape[0]
    ns = {"F": F, "torch": torch, "self": FakeSelf(), "protect": protect, "pitch_guidance": pitch_guidance,
          "feats": torch.from_numpy(feats.copy()).unsqueeze(0),
          "feats0": torch.from_numpy(feats0.copy()).unsqueeze(0) if pitch_guidance else None,
          "audio0": np.zeros(2 * n * FakeSelf.window, np.float32),
          "pitch": torch.zeros((1, 2 * n + 5), dtype=torch.long) if pitch_guidance else None,
          "pitchf": torch.from_numpy(np.concatenate([pitchf, np.zeros(5, np.float32)])).unsqueeze(0)
          if pitch_guidance else None}
    exec(code, ns)
    out = ns["feats"]
    assert out.shape == (1, 2 * n, feats.shape[1]) and out.dtype == torch.float32
    return out[0].numpy().copy()


def make_pitchf(rng, m):
    p = rng.uniform(80.0, 400.0, m).astype(np.float32)
    kind = rng.integers(0, 6, m)
    p[kind == 0] = 0.0                                            # unvoiced frames
    p[kind == 1] = rng.uniform(0.0, 1.0, int((kind == 1).sum())).astype(np.float32)   # 0 < f0 < 1 -> protect
    p[kind == 2] = 1.0                                            # exactly 1 -> 1
    return p


def main():
    stmts = reference_statements()
    code = compile(stmts, "reference:rvc/infer/pipeline.py:463-481", "exec")
    rng = np.random.default_rng(20261018)
    for name, protect in (("retrieve_v1_256", 0.33), ("retrieve_v2_768", 0.2)):
        z = np.load(os.path.join(OUT, name + ".npz"))
        q, blended = z["q"], z["out"]
        pitchf = make_pitchf(rng, 2 * q.shape[0])
        out_mix = run_reference(code, blended, q, pitchf, protect, True)
        out_up = run_reference(code, blended, q, pitchf, 0.5, True)          # protect >= 0.5: upsample only
        out_nopitch = run_reference(code, blended, q, pitchf, protect, False)  # no pitch guidance: upsample only
        assert np.array_equal(out_up, out_nopitch, equal_nan=True)
        np.savez_compressed(os.path.join(OUT, name.replace("retrieve", "post") + ".npz"), pitchf=pitchf,
                            protect=np.float64(protect), out_mix=out_mix, out_up=out_up)
        print(name, "->", out_mix.shape, "nan rows", int(np.isnan(out_mix).any(axis=1).sum()))
    # stand-alone: d not a multiple of 4 exercises the scalar kernel, d = 768 the float4 one
    for d in (768, 30):
        n = 21
        feats = (rng.standard_normal((n, d)) * 0.33).astype(np.float32)
        feats0 = (rng.standard_normal((n, d)) * 0.33).astype(np.float32)
        pitchf = make_pitchf(rng, 2 * n)
        out_mix = run_reference(code, feats, feats0, pitchf, 0.33, True)
        np.savez_compressed(os.path.join(OUT, f"post_standalone_{d}.npz"), feats=feats, feats0=feats0, pitchf=pitchf,
                            protect=np.float64(0.33), out_mix=out_mix)
    with open(os.path.join(OUT, "reference_post_statements.txt"), "w") as f:
        f.write("# statements cut from /root/reference/rvc/infer/pipeline.py (voice_conversion) by make_golden_post.py\n")
        f.write("# provenance record only; not imported by anything.  sha256 of the executed source: ")
        f.write(hashlib.sha256(stmts.encode()).hexdigest() + "\n")


if __name__ == "__main__":
    main()
