"""GPU, min(8, device count) ranks on one box (NVLink peer memory + NCCL): the sharded path through the C ABI equals the
single-GPU path bit for bit -- assignments, neighbour ids, distances and blended rows -- over both transports (the
device-driven inbox step of csrc/peer.cu and the NCCL all-to-all route), with ragged and empty home batches and with a
query mass that lands almost entirely on one shard.  Skipped with < 2 GPUs."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    import importlib

    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    rvcr = importlib.import_module("codename-rvc-fork-3_b200")
    sharded = importlib.import_module("codename-rvc-fork-3_b200.sharded")
    synth = importlib.import_module("codename-rvc-fork-3_b200.synth")
    path = os.path.join(tmp, "full.index")
    z = np.load(os.path.join(tmp, "problem.npz"))
    q_all = z["q"]
    shard = rvcr.read_index(path, device=rank, shard=(rank, world))
    full = rvcr.read_index(path, device=rank)
    sizes = full.list_sizes()
    assert (shard.info["owned_begin"], shard.info["owned_end"]) == synth.shard_ranges(sizes, world)[rank]
    sh = sharded.ShardedIndex(sharded.CudaShardBackend(shard))
    n = len(q_all)
    lo = n * rank // world + (3 if rank else 0)                              # ragged home batches
    hi = n * (rank + 1) // world + (3 if rank + 1 < world else 0)
    q = q_all[lo:hi]
    for mode in (1, 2):
        shard.set_option("coarse_mode", mode)
        full.set_option("coarse_mode", mode)
        ref = torch.from_numpy(q.copy()).cuda()
        _, Dr, Ir = full.retrieve_blend_(ref, 0.75, want_neighbours=True)
        got = torch.from_numpy(q.copy()).cuda()
        _, Dg, Ig = sh.retrieve_blend_(got, 0.75, want_neighbours=True)
        torch.cuda.synchronize()
        assert torch.equal(Ig, Ir) and torch.equal(Dg, Dr)
        assert np.array_equal(got.cpu().numpy(), ref.cpu().numpy(), equal_nan=True)
        Ds, Is = sh.search(torch.from_numpy(q.copy()).cuda())
        assert torch.equal(Is, Ir) and torch.equal(Ds, Dr)
        # host rows: chunked H2D overlapped with the coarse step (3 chunks here), one D2H back
        hp = torch.from_numpy(q.copy()).pin_memory()
        sh.retrieve_blend_host_(hp, 0.75, chunk_rows=400)
        assert np.array_equal(hp.numpy(), ref.cpu().numpy(), equal_nan=True)
    # ---- peer memory transport: owner pulls rows over NVLink, blend epilogue pushes them back (sharded.py) ----
    peer = sh.enable_peer(4096)
    for mode, scan in ((2, 5), (1, 2), (0, 5)):                  # tcgen05 scan re-check epilogue, gather_blend epilogue
        for hnd in (shard, full):
            hnd.set_option("coarse_mode", mode)
            hnd.set_option("scan_impl", scan)
        ref = torch.from_numpy(q.copy()).cuda()
        full.retrieve_blend_(ref, 0.75)
        feats = peer.rows(len(q), 5 + rank)
        feats.copy_(torch.from_numpy(q))
        launched = shard.stats()["kernels_launched"]
        out = sh.retrieve_blend_(feats, 0.75)
        torch.cuda.synchronize()
        assert out.data_ptr() == feats.data_ptr() and peer.row_offset(feats) == 5 + rank
        assert np.array_equal(feats.cpu().numpy(), ref.cpu().numpy(), equal_nan=True), (mode, scan)
        assert shard.stats()["kernels_launched"] > launched
        # host rows staged in the peer buffer
        hp = torch.from_numpy(q.copy()).pin_memory()
        sh.retrieve_blend_host_(hp, 0.75, chunk_rows=400, device_rows=peer.rows(len(q), 2000))
        assert np.array_equal(hp.numpy(), ref.cpu().numpy(), equal_nan=True)
    # a tensor outside the peer buffer is refused locally (no rank starts a collective, nothing to agree on) ...
    got = torch.from_numpy(q.copy()).cuda()
    with pytest.raises(ValueError):
        sh.retrieve_blend_(got, 0.75)
    # ... and asking for the neighbour lists takes the NCCL route on every rank, peer buffer or not
    _, Dn, In = sh.retrieve_blend_(got, 0.75, want_neighbours=True)
    assert np.array_equal(got.cpu().numpy(), ref.cpu().numpy(), equal_nan=True)
    # skewed query mass: (almost) every query of every rank probes a list of shard 0 -- one inbox takes the whole step
    own0 = shard.info["owned_end"] if rank == 0 else None
    t = torch.tensor([own0 or 0], device="cuda")
    dist.broadcast(t, src=0)
    a_all = full.assign(torch.from_numpy(q_all).cuda())[0].cpu().numpy()
    heavy = q_all[(a_all >= 0) & (a_all < int(t.item()))]
    qs = heavy[rank::world][:1500]
    qs = np.concatenate([qs, q[:5]]).astype(np.float32)          # plus a few queries for the other shards
    ref = torch.from_numpy(qs.copy()).cuda()
    full.retrieve_blend_(ref, 0.75)
    fs = peer.rows(len(qs), 64)
    fs.copy_(torch.from_numpy(qs))
    for it in range(4):                                          # back-to-back steps reuse the inbox (count reset, epochs)
        peer.pull_rows_first = it == 3                           # last step: the variant that pulls the rows into a scratch first
        fs.copy_(torch.from_numpy(qs))
        sh.retrieve_blend_(fs, 0.75)
        torch.cuda.synchronize()
        assert np.array_equal(fs.cpu().numpy(), ref.cpu().numpy(), equal_nan=True), it
    peer.pull_rows_first = False
    # an empty home batch on one rank, over both transports
    q2 = q[:0] if rank == world - 1 else q[:7]
    fp = peer.rows(len(q2), 100)
    fp.copy_(torch.from_numpy(q2))
    sh.retrieve_blend_(fp, 0.3)
    f2 = torch.from_numpy(q2.copy()).cuda()
    sh.peer, keep = None, sh.peer                                # the NCCL all-to-all transport
    sh.retrieve_blend_(f2, 0.3)
    sh.peer = keep
    torch.cuda.synchronize()
    assert np.array_equal(fp.cpu().numpy(), f2.cpu().numpy(), equal_nan=True)
    if len(q2):
        r2 = torch.from_numpy(q2.copy()).cuda()
        full.retrieve_blend_(r2, 0.3)
        assert np.array_equal(f2.cpu().numpy(), r2.cpu().numpy(), equal_nan=True)
    dist.barrier()
    sh.peer.close()
    dist.destroy_process_group()
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")


def test_shards_equal_one_gpu(rvcr, oracle, tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = min(8, torch.cuda.device_count())
    from importlib import import_module
    synth = import_module("codename-rvc-fork-3_b200.synth")
    pr = synth.make_host_problem(40_000, 768, 2000, 512, seed=9, dup_frac=0.005)
    pr.q[17, 5] = np.nan
    oidx = oracle.OracleIndex.from_vectors(pr.x, pr.centroids)
    oracle.write_index(oidx.to_data(), str(tmp_path / "full.index"))
    np.savez(tmp_path / "problem.npz", q=pr.q)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
