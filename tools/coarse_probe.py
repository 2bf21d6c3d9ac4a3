"""Coarse-quantizer probe at the C5 centroid count: assign() time and per-phase times for several batch sizes,
back to back and after an idle gap (clock / power state), for the tc_resident settings."""
import importlib, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
rvcr = importlib.import_module("codename-rvc-fork-3_b200"); synth = rvcr.synth
dev = torch.device("cuda:0")
nlist, d = 90509, 768
cent = synth.all_centroids(nlist, d, dev)
index = rvcr.index_factory(d, f"IVF{nlist},Flat"); index.nprobe = 1
index.set_centroids(cent.cpu().numpy())
index.set_option("profile", 1)
q = synth.device_queries(cent, 1_000_000)
def timed(n, reps=3, chunks=1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    index.assign(q[:n]); torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        for c in range(chunks):
            index.assign(q[c * n:(c + 1) * n])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for n in (32768, 65536, 125000, 131072, 250000, 1_000_000):
    for ns in (0, 1, 2, 3, 4, 6, 8, 13):
        if ns:
            os.environ["RVCR_TC_NSPLIT"] = str(ns)
        else:
            os.environ.pop("RVCR_TC_NSPLIT", None)
        ms = timed(n)
        print(json.dumps({"n": n, "nsplit_env": ns, "ms": round(ms, 3), "Mq_per_s": round(n / ms / 1e3, 3),
                          "phase": {k: round(v, 3) for k, v in index.phase_ms().items()}}), flush=True)
