"""On-GPU probe (not the bench) of BASELINE config C4 -- the index build of rvc/train/process/extract_index.py:62-71
with given centroids: nearest-centroid assignment + list add for 4 M synthetic 768-d training frames into
nlist = 32 000, single B200.  Reports vectors/s for

  a. the reference's own loop shape: host numpy rows, `index.add(big_npy[i:i+8192])` (H2D inside each call);
  b. the same 8192-row batches already on the device;
  c. device batches of 262 144 rows (what a device-resident feature extractor would hand over);

each followed by the fold into the list layout (first consumer: list_sizes), plus the CPU port's add rate on a
65 536-row sample (OpenBLAS coarse + list append) for the same centroids.  Properties checked at full size: every
vector lands in exactly one list, ids ascend inside every list, and the list of a sample of vectors equals the
oracle's canonical nearest centroid."""
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
synth = importlib.import_module("codename-rvc-fork-3_b200.synth")
import oracle  # noqa: E402  (checker + CPU baseline only)

dev = torch.device("cuda:0")
N, d, _, nlist = synth.CONFIGS["C4"]
cent = synth.all_centroids(nlist, d, dev)
cent_h = cent.cpu().numpy()


def make_rows(a, b):
    """rows a..b of the synthetic training set: centroid[list] + N(0, 0.14^2), list uniform (device)."""
    g = torch.Generator(device=dev).manual_seed(777 + a)
    lists = torch.randint(0, nlist, (b - a,), generator=g, device=dev)
    return cent[lists] + torch.randn((b - a, d), generator=g, device=dev) * 0.14


def run(batch, host):
    index = rvcr.index_factory(d, f"IVF{nlist},Flat", device=0)
    index.nprobe = 1
    index.set_centroids(cent_h)
    chunk = 1 << 19                                       # rows generated per outer step (1.6 GB)
    t_total = 0.0
    for c0 in range(0, N, chunk):
        x = make_rows(c0, min(N, c0 + chunk))
        xh = x.cpu().numpy() if host else None
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(0, x.shape[0], batch):
            index.add(xh[i:i + batch] if host else x[i:i + batch])
        torch.cuda.synchronize()
        t_total += time.perf_counter() - t0
        del x, xh
    t0 = time.perf_counter()
    sizes = index.list_sizes()                            # folds the staged rows into the list layout
    torch.cuda.synchronize()
    t_fold = time.perf_counter() - t0
    assert int(sizes.sum()) == N == index.ntotal
    return index, sizes, t_total, t_fold


res = {"config": {"workload": "C4", "vectors": N, "d": d, "nlist": nlist}}
for name, batch, host in (("host_batches_8192", 8192, True), ("device_batches_8192", 8192, False),
                          ("device_batches_262144", 262144, False)):
    index, sizes, t_add, t_fold = run(batch, host)
    res[name] = {"add_s": t_add, "fold_s": t_fold, "vectors_per_s": N / (t_add + t_fold),
                 "assign_tflops": 2.0 * N * nlist * d / t_add / 1e12}
    if name != "device_batches_262144":
        del index
        torch.cuda.empty_cache()

# ---- properties at full size (last index) ----------------------------------------------------------
vecs, ids, offs, _ = index.device_views()
seg = torch.repeat_interleave(torch.arange(nlist, device=dev), torch.as_tensor(sizes, device=dev))
same = seg[1:] == seg[:-1]
res["ids_ascending_inside_every_list"] = bool((ids[1:][same] > ids[:-1][same]).all())
res["every_id_exactly_once"] = bool(torch.equal(torch.sort(ids).values, torch.arange(N, device=dev)))
g = torch.Generator(device=dev).manual_seed(1)
pos = torch.randint(0, N, (512,), generator=g, device=dev)
a_o, _ = oracle.coarse_assign(vecs[pos].cpu().numpy(), cent_h)
res["sample_lists_equal_oracle_argmin"] = bool(np.array_equal(seg[pos].cpu().numpy(), a_o))

# ---- CPU port on a bounded sample (faiss-style: sgemm coarse for >= 20 rows, then append) -----------
xs = make_rows(0, 65536).cpu().numpy()
t0 = time.perf_counter()
a = oracle.coarse_assign_blas(xs, cent_h)
a = a[0] if isinstance(a, tuple) else a
oracle.build_lists(xs, np.asarray(a, np.int32), nlist)
t_cpu = time.perf_counter() - t0
res["cpu_port"] = {"sample_rows": 65536, "seconds": t_cpu, "vectors_per_s": 65536 / t_cpu,
                   "threads": oracle.num_threads(), "kind": "port (extrapolates linearly to 4 M rows)"}
# ---- index.train at C4 scale: rvcr_index_train on 4 M host rows (extract_index.py:65) ----------------
del index, vecs, ids, offs, seg
torch.cuda.empty_cache()
if "--train" in sys.argv:
    xh = np.empty((N, d), np.float32)
    for c0 in range(0, N, 1 << 19):
        xh[c0:c0 + (1 << 19)] = make_rows(c0, min(N, c0 + (1 << 19))).cpu().numpy()
    tindex = rvcr.index_factory(d, f"IVF{nlist},Flat", device=0)
    t0 = time.perf_counter()
    tindex.train(xh)
    torch.cuda.synchronize()
    t_train = time.perf_counter() - t0
    c = tindex.get_centroids()
    # 10 assignments of N rows against nlist centroids (n <= 256 * nlist: no subsampling at C4)
    res["train"] = {"seconds": t_train, "niter": 10, "assign_pflop": 10 * 2.0 * N * nlist * d / 1e15,
                    "pflops": 10 * 2.0 * N * nlist * d / 1e15 / t_train, "finite": bool(np.isfinite(c).all()),
                    "cpu_port_estimate_s": 10 * N / res["cpu_port"]["vectors_per_s"],
                    "cpu_port_note": "10 x the measured CPU assignment rate above (sgemm coarse), extrapolated"}
print(json.dumps(res))
