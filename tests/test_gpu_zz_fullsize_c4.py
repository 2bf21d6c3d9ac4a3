"""GPU: BASELINE.json config 4 at full size on one B200 -- the index build of rvc/train/process/extract_index.py:62-71 with
given centroids: nearest-centroid assignment + list add of 4 M synthetic 768-d training frames into nlist = 32 000 (the
workload `bench.py --workload C4` times).  Checked through properties that do not need a 4 M-row CPU run:

  * every id 0 .. N-1 is stored exactly once, ids ascend inside every list (faiss's sequential append), the stored row of
    an id is the row that was added, ntotal / list sizes add up;
  * the tensor-core assignment equals the exact FFMA kernel on every row (4 M x 32 000 pairs), and the CPU oracle's
    canonical argmin on a sample;
  * the file written for a 200 k-row slice of the same build is byte-identical to the oracle's writer.

Needs ~30 GB of HBM."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_c4_build_properties(rvcr, oracle, tmp_path):
    import torch
    synth = importlib.import_module("codename-rvc-fork-3_b200.synth")
    free, _ = torch.cuda.mem_get_info(0)
    if free < 40 * (1 << 30):
        pytest.skip(f"C4 needs ~30 GB of free HBM, {free >> 30} GB available")
    N, d, _, nlist = synth.CONFIGS["C4"]
    dev = torch.device("cuda:0")
    cent = synth.all_centroids(nlist, d, dev)
    cent_h = cent.cpu().numpy()
    x = torch.empty((N, d), dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev)
    for c0 in range(0, N, 1 << 19):
        g.manual_seed(777 + c0)
        n = min(N, c0 + (1 << 19)) - c0
        lists = torch.randint(0, nlist, (n,), generator=g, device=dev)
        x[c0:c0 + n] = cent[lists] + torch.randn((n, d), generator=g, device=dev) * 0.14
    x[12345] = x[777]                                            # an exact duplicate row: same list, id order kept
    index = rvcr.index_factory(d, f"IVF{nlist},Flat", device=0)  # extract_index.py:62
    rvcr.extract_index_ivf(index).nprobe = 1                     # :63-64
    index.set_centroids(cent_h)
    for i in range(0, N, 8192 * 16):                             # :67-69 (device batches; the 8192-row host loop is bench.py's e2e leg)
        index.add(x[i:i + 8192 * 16])
    sizes = index.list_sizes()
    assert index.ntotal == N and int(sizes.sum()) == N
    vecs, ids, offs, _ = index.device_views()
    assert torch.equal(torch.sort(ids).values, torch.arange(N, device=dev))
    seg = torch.repeat_interleave(torch.arange(nlist, device=dev), torch.as_tensor(sizes, device=dev))
    same = seg[1:] == seg[:-1]
    assert bool((ids[1:][same] > ids[:-1][same]).all())
    step = 1 << 18
    for s in range(0, N, step):                                  # stored row == added row, every row
        assert torch.equal(vecs[s:s + step], x[ids[s:s + step]])
    # list of every row: tensor-core assignment == exact FFMA kernel; == the list the row was stored in
    index.set_option("coarse_mode", 2)
    a_tc = torch.cat([index.assign(x[i:i + (1 << 20)])[0] for i in range(0, N, 1 << 20)])
    index.set_option("coarse_mode", 1)
    sel = torch.randint(0, N, (65536,), generator=g, device=dev)
    a_ex = index.assign(x[sel])[0]
    assert torch.equal(a_tc[sel], a_ex)
    stored_list = torch.empty(N, dtype=torch.int64, device=dev)
    stored_list[ids] = seg
    assert torch.equal(stored_list, a_tc.to(torch.int64))
    assert stored_list[12345] == stored_list[777]
    sub = sel[:512]
    a_o, _ = oracle.coarse_assign(x[sub].cpu().numpy(), cent_h)
    assert np.array_equal(a_ex[:512].cpu().numpy(), a_o)
    # a 200 k-row slice of the same build, written and compared byte for byte with the oracle's writer
    xs = x[:200_000].cpu().numpy()
    small = rvcr.index_factory(d, f"IVF{nlist},Flat", device=0)
    small.nprobe = 1
    small.set_centroids(cent_h)
    for i in range(0, len(xs), 8192):
        small.add(xs[i:i + 8192])
    path = str(tmp_path / "slice.index")
    rvcr.write_index(small, path)
    a_s = a_tc[:200_000].cpu().numpy()                            # (assignments already checked above)
    off, ov, oi = oracle.build_lists(xs, a_s.astype(np.int32), nlist)
    data = oracle.IVFFlatData(d=d, ntotal=len(xs), nprobe=1, centroids=cent_h, list_sizes=np.diff(off).astype(np.int64), vecs=ov, ids=oi)
    ref = str(tmp_path / "slice_ref.index")
    oracle.write_index(data, ref)
    assert open(path, "rb").read() == open(ref, "rb").read()
