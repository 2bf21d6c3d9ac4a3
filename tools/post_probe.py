"""On-GPU probe (not the bench) of the post step, SURVEY 8f-3 (rvc/infer/pipeline.py:463-481):

  1. rvcr_upsample_protect alone vs the reference's own torch statements on the same GPU (1 M x 768 frames);
  2. on the C3 workload (2 M x 768 index, 180 k frames): retrieve_blend_ in place vs the one-call
     retrieve_blend_upsample (blend epilogue also upsamples + mixes) vs retrieve_blend_ + torch post statements.

Prints one JSON object; CUDA events on the launching stream, inputs >> L2, warm-up first."""
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rvcr = importlib.import_module("codename-rvc-fork-3_b200")
synth = importlib.import_module("codename-rvc-fork-3_b200.synth")
dev = torch.device("cuda:0")
PEAK = 6540.2
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def torch_post(feats, feats0, pitchf, protect):
    """pipeline.py:463-481 as written (feats (1, n, d), pitchf (1, 2n))."""
    feats = F.interpolate(feats.permute(0, 2, 1), scale_factor=2).permute(0, 2, 1)
    feats0 = F.interpolate(feats0.permute(0, 2, 1), scale_factor=2).permute(0, 2, 1)
    pitchff = pitchf.clone()
    pitchff[pitchf > 0] = 1
    pitchff[pitchf < 1] = protect
    feats = feats * pitchff.unsqueeze(-1) + feats0 * (1 - pitchff.unsqueeze(-1))
    return feats


res = {}
# ---- 1. the stand-alone kernel -------------------------------------------------------------------
n, d = 1_000_000, 768
g = torch.Generator(device=dev).manual_seed(7)
feats = torch.randn((1, n, d), generator=g, device=dev)
feats0 = torch.randn((1, n, d), generator=g, device=dev)
pitchf = torch.where(torch.rand((1, 2 * n), generator=g, device=dev) < 0.3, torch.zeros((), device=dev),
                     torch.rand((1, 2 * n), generator=g, device=dev) * 400 + 50)
ms_ours = timeit(lambda: rvcr.upsample_protect(feats, feats0, pitchf, 0.33))
ms_torch = timeit(lambda: torch_post(feats, feats0, pitchf, 0.33))
same = bool(torch.equal(rvcr.upsample_protect(feats, feats0, pitchf, 0.33), torch_post(feats, feats0, pitchf, 0.33)))
bytes_alg = n * d * 4 * 4 + 2 * n * 4            # read feats + feats0, write 2 rows each; pitchf
res["standalone_1Mx768"] = {"ours_ms": ms_ours, "torch_reference_ms": ms_torch, "bit_identical_to_torch": same,
                            "algorithmic_bytes": bytes_alg, "ours_GBps": bytes_alg / ms_ours / 1e6,
                            "frac_of_hbm_peak": bytes_alg / ms_ours / 1e6 / PEAK, "peak_GBps": PEAK}
ms_up = timeit(lambda: rvcr.upsample_protect(feats))
res["standalone_1Mx768"]["upsample_only_ms"] = ms_up
res["standalone_1Mx768"]["upsample_only_GBps"] = n * d * 4 * 3 / ms_up / 1e6
del feats, feats0, pitchf
torch.cuda.empty_cache()

# ---- 2. fused into the blend epilogue, C3 ---------------------------------------------------------
N, d, nq, nlist = synth.CONFIGS["C3"]
sizes = synth.list_sizes(N, nlist)
cent = synth.all_centroids(nlist, d, dev)
index = rvcr.index_factory(d, f"IVF{nlist},Flat", device=0)
index.nprobe = 1
synth.fill_index_by_construction(index, cent, sizes, N)
q = synth.device_queries(cent, nq)
pitchf = torch.where(torch.rand((1, 2 * nq), generator=g, device=dev) < 0.3, torch.zeros((), device=dev),
                     torch.rand((1, 2 * nq), generator=g, device=dev) * 400 + 50)
bufs = [q.clone() for _ in range(8)]
it = iter(range(10 ** 9))


def inplace():
    index.retrieve_blend_(bufs[next(it) % 8].copy_(q), 0.75)


def fused():
    index.retrieve_blend_upsample(q, 0.75, pitchf=pitchf, protect=0.33)


def unfused_torch():
    b = bufs[next(it) % 8].copy_(q)
    f0 = b.clone()                                          # pipeline.py:455
    index.retrieve_blend_(b, 0.75)
    torch_post(b.unsqueeze(0), f0.unsqueeze(0), pitchf, 0.33)


def unfused_ours():
    b = bufs[next(it) % 8].copy_(q)
    f0 = b.clone()
    index.retrieve_blend_(b, 0.75)
    rvcr.upsample_protect(b, f0, pitchf, 0.33)


ms_copy = timeit(lambda: bufs[next(it) % 8].copy_(q))
a = index.retrieve_blend_upsample(q, 0.75, pitchf=pitchf, protect=0.33)
b = q.clone()
index.retrieve_blend_(b, 0.75)
want = torch_post(b.unsqueeze(0), q.unsqueeze(0), pitchf, 0.33)[0]
res["c3_fused"] = {"retrieve_blend_inplace_ms": timeit(inplace) - ms_copy,
                   "retrieve_blend_upsample_fused_ms": timeit(fused),
                   "retrieve_blend_then_rvcr_upsample_protect_ms": timeit(unfused_ours) - ms_copy,
                   "retrieve_blend_then_torch_reference_post_ms": timeit(unfused_torch) - ms_copy,
                   "fused_bit_identical_to_unfused_torch": bool(torch.equal(a, want)),
                   "query_frames": nq, "index_vectors": N}
print(json.dumps(res))
