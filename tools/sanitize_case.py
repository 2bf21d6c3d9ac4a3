"""The workload tools/sanitize.sh runs under compute-sanitizer: every kernel family once on small inputs, tensor-core
paths forced (the size rules would pick the exact FFMA kernels for inputs this small), results checked against the oracle."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (checker only)

rvcr = importlib.import_module("codename-rvc-fork-3_b200")
synth = importlib.import_module("codename-rvc-fork-3_b200.synth")

pr = synth.make_host_problem(6000, 768, 700, 160, seed=21)
oidx = oracle.OracleIndex.from_vectors(pr.x, pr.centroids)
index = rvcr.index_factory(768, "IVF160,Flat")
index.nprobe = 1
index.set_centroids(pr.centroids)
for i in range(0, len(pr.x), 2048):
    index.add(pr.x[i:i + 2048])                                  # coarse + stage + fold (lists.cu)
Do, Io = oidx.search(pr.q, k=8)
big = oidx.reconstruct_n(0, oidx.ntotal)
with np.errstate(all="ignore"):
    want = oracle.tail(Do, Io, big, 0.75, pr.q)
for cm, si in ((1, 4), (2, 5), (3, 5), (4, 5)):                  # exact kernels; tcgen05 bf16 / 3-way split / e4m3 + tcgen05 scan
    index.set_option("coarse_mode", cm)
    index.set_option("scan_impl", si)
    for cl in (2, 1):
        index.set_option("tc_cluster", cl)
        D, I = index.search(pr.q, k=8)
        assert np.array_equal(I, Io) and np.array_equal(D, Do), (cm, si, cl)
        f = torch.from_numpy(pr.q.copy()).cuda()
        index.retrieve_blend_(f, 0.75)
        assert np.array_equal(f.cpu().numpy(), want, equal_nan=True), (cm, si, cl)
pf = torch.rand(2 * len(pr.q), device="cuda") * 300
out = index.retrieve_blend_upsample(torch.from_numpy(pr.q.copy()).cuda(), 0.75, pitchf=pf, protect=0.33)   # fused post step
assert out.shape == (2 * len(pr.q), 768)
h = pr.q.copy()
index.retrieve_blend_host_(h, 0.75)                              # host variant (waves into a pageable buffer)
assert np.array_equal(h, want, equal_nan=True)
index.nprobe = 3                                                 # multiprobe.cu
D3, I3 = index.search(pr.q, k=8)
assert np.array_equal(I3, oidx.search(pr.q, k=8, nprobe=3)[1])
index.nprobe = 1
assert np.array_equal(index.reconstruct_n(0, index.ntotal), pr.x)  # lists.cu reconstruct
t = rvcr.index_factory(768, "IVF40,Flat")
t.train(pr.x)                                                    # kmeans.cu
assert np.array_equal(t.get_centroids(), oracle.kmeans_train(pr.x, 40)[0])
w = torch.randn(256, 768, device="cuda")
rvcr.final_proj(torch.from_numpy(pr.q).cuda(), w, None)          # final_proj.cu
rvcr.MiniBatchKMeans(n_clusters=64, batch_size=512, random_state=1).fit(pr.x)   # minibatch_kmeans.cu
torch.cuda.synchronize()
print("sanitize case ok")
