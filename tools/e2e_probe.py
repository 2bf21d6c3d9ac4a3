"""PCIe / e2e probe: plain pinned D2H + H2D bandwidth, and the host-buffer API with 1 / 8 / 16 write-back waves."""
import importlib, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
rvcr = importlib.import_module("codename-rvc-fork-3_b200"); synth = rvcr.synth
dev = torch.device("cuda:0")
N, d, nq, nlist = 8_000_000, 768, 250_000, 22_627
sizes = synth.list_sizes(N, nlist)
cent = synth.all_centroids(nlist, d, dev)
index = rvcr.index_factory(d, f"IVF{nlist},Flat"); index.nprobe = 1
synth.fill_index_by_construction(index, cent, sizes, N)
q = synth.device_queries(cent, nq)
host = torch.empty((nq, d), dtype=torch.float32).pin_memory()
for name, fn in (("d2h", lambda: host.copy_(q, non_blocking=True)), ("h2d", lambda: q.copy_(host, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 3
    print(json.dumps({"probe": name, "GBs": nq * d * 4 / dt / 1e9, "ms": dt * 1e3}))
src = q.cpu()
for waves in (1, 4, 8, 16):
    os.environ["RVCR_HOST_WAVES"] = str(waves)
    ts = []
    for it in range(3):
        host.copy_(src); torch.cuda.synchronize(); t = time.perf_counter()
        index.retrieve_blend_host_(host.numpy(), 0.75)
        ts.append(time.perf_counter() - t)
    print(json.dumps({"waves": waves, "ms": min(ts) * 1e3}))
b = q.clone(); index.set_option("profile", 1); index.retrieve_blend_(b, 0.75); torch.cuda.synchronize(); print(json.dumps(index.phase_ms()))
