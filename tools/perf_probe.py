"""Quick on-GPU probe (not the bench): build a C3-like index through the add path and time the phases."""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rvcr = importlib.import_module("codename-rvc-fork-3_b200")

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=2_000_000)
ap.add_argument("--d", type=int, default=768)
ap.add_argument("--nq", type=int, default=180_000)
ap.add_argument("--nlist", type=int, default=22_627)
ap.add_argument("--mode", type=int, default=1)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--cluster", type=int, default=2)
ap.add_argument("--scan", type=int, default=5)
ap.add_argument("--addmode", type=int, default=2)
a = ap.parse_args()

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1234)
nb = max(64, a.nlist // 4)
centres = torch.randn((nb, a.d), generator=g, device=dev) * 0.30
which = torch.randint(0, nb, (a.n,), generator=g, device=dev)
x = centres[which] + torch.randn((a.n, a.d), generator=g, device=dev) * 0.14
cent = x[torch.randperm(a.n, generator=g, device=dev)[: a.nlist]].contiguous()
index = rvcr.index_factory(a.d, f"IVF{a.nlist},Flat")
index.set_centroids(cent.cpu().numpy())
index.set_option("coarse_mode", a.addmode)
index.set_option("tc_cluster", a.cluster)
index.set_option("scan_impl", a.scan)
index.set_option("profile", 1)
torch.cuda.synchronize()
t0 = time.time()
B = 262144
for i in range(0, a.n, B):
    index.add(x[i:i + B].contiguous())
sizes = index.list_sizes()
torch.cuda.synchronize()
t_add = time.time() - t0
print(json.dumps({"add_s": t_add, "add_vec_per_s": a.n / t_add, "list_min": int(sizes.min()), "list_max": int(sizes.max()),
                  "list_mean": float(sizes.mean())}))
index.set_option("coarse_mode", a.mode)
rows = torch.randint(0, a.n, (a.nq,), generator=g, device=dev)
q0 = (x[rows] + torch.randn((a.nq, a.d), generator=g, device=dev) * 0.05).contiguous()
del x
for r in range(a.reps):
    q = q0.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    index.retrieve_blend_(q, 0.75)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ph = index.phase_ms()
    nst = int(sizes.sum())
    scan_bytes = nst * (4 * a.d + 8) + a.nq * 4 * a.d + a.nq * 8 * 12      # upper bound: every list probed
    blend_bytes = a.nq * (8 * 4 * a.d + 8 * a.d) + a.nq * 8 * 12
    out = {"rep": r, "ms": ms, "frames_per_s": a.nq / ms * 1e3, "phase_ms": ph,
           "coarse_tflops": 2.0 * a.nq * a.nlist * a.d / (ph.get("coarse_main", 1e9) * 1e-3) / 1e12,
           "scan_GBs_alllists": scan_bytes / (ph.get("scan", 1e9) * 1e-3) / 1e9,
           "blend_GBs": blend_bytes / (ph.get("blend", 1e9) * 1e-3) / 1e9}
    print(json.dumps(out))
print(json.dumps(index.stats()))
