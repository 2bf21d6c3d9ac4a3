"""CPU, world_size 2 and 3, gloo: the host-side logic of the sharded path (partitioning, the small
all-gather + merge, ownership, the all-to-all that returns rows to their home rank).  The compute
backend injected here is the oracle (test infrastructure) -- on the GPU the same ShardedIndex code
drives the C-ABI backend (tests/test_gpu_sharded.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleShardBackend:
    """The oracle standing in for the C-ABI backend: full quantizer, lists [lo, hi) only."""
    device = "cpu"

    def __init__(self, oracle, x, cent, lo, hi):
        self.o = oracle
        self.x = x
        self.cent = cent
        full = oracle.OracleIndex.from_vectors(x, cent)
        self.nlist = len(cent)
        self.owned = (lo, hi)
        sizes = np.diff(full.list_off)
        local = np.zeros_like(sizes)
        local[lo:hi] = sizes[lo:hi]
        self.off = np.zeros(self.nlist + 1, np.int64)
        np.cumsum(local, out=self.off[1:])
        b, e = full.list_off[lo], full.list_off[hi]
        self.vecs, self.ids = full.vecs[b:e], full.ids[b:e]

    def assign(self, q):
        a, d = self.o.coarse_assign(q.numpy(), self.cent)
        return torch.from_numpy(a), torch.from_numpy(d)

    def scan_assigned(self, feats, assign, index_rate, blend=True):
        q = feats.numpy()
        a = assign.numpy().astype(np.int64)
        lo, hi = self.owned
        own = (a < 0) | ((a >= lo) & (a < hi))          # invalid assignments are answered with pads / NaN
        n = q.shape[0]
        D = np.full((n, 8), np.float32(3.4028234663852886e38), np.float32)
        I = np.full((n, 8), -1, np.int64)
        if own.any():
            Do, Io = self.o.scan_topk(q[own], a[own].astype(np.int32), self.off, self.vecs, self.ids)
            D[own], I[own] = Do, Io
            if blend:
                with np.errstate(all="ignore"):
                    q[own] = self.o.tail(Do, Io, self.x, index_rate, q[own])
        return torch.from_numpy(D), torch.from_numpy(I)

    def scan_assigned_push(self, rows, assign, index_rate, peer, dst_rank, dst_row):
        """what rvcr_retrieve_blend_assigned_push does in its epilogue: blended row i -> buffer[dst_rank[i]][dst_row[i]]"""
        self.scan_assigned(rows, assign, index_rate, blend=True)
        peer.push_rows(dst_rank, dst_row, rows)

    def merge_topk(self, D_all, I_all):
        D, I = D_all.numpy(), I_all.numpy()
        G, n, k = D.shape
        Dc = np.transpose(D, (1, 0, 2)).reshape(n, G * k)
        Ic = np.transpose(I, (1, 0, 2)).reshape(n, G * k)
        key_i = np.where(Ic < 0, np.iinfo(np.int64).max, Ic)
        order = np.lexsort((key_i, Dc), axis=1)[:, :k]
        return torch.from_numpy(np.take_along_axis(Dc, order, 1)), torch.from_numpy(np.take_along_axis(Ic, order, 1))


class SharedMemoryPeer:
    """Stand-in for sharded.CudaPeer on CPU: the per-rank feature buffers are torch shared-memory tensors that every
    process maps (what cudaIpc does for device memory), so the pull / push bookkeeping of the peer step runs for real."""

    def __init__(self, buffers, rank):
        self.buffers, self.rank = buffers, rank
        self.d = buffers[0].shape[1]
        self.pulled = self.pushed = 0

    def rows(self, n, row0=0):
        return self.buffers[self.rank][row0:row0 + n]

    def row_offset(self, t):
        own = self.buffers[self.rank]
        off = t.data_ptr() - own.data_ptr()
        if t.dim() != 2 or off < 0 or off % (self.d * 4) or off + t.numel() * 4 > own.numel() * 4:
            return None
        return off // (self.d * 4)

    def pull_rows(self, src_rank, src_row):
        self.pulled += len(src_rank)
        if len(src_rank) == 0:
            return torch.empty((0, self.d), dtype=torch.float32)
        return torch.stack([self.buffers[int(r)][int(i)].clone() for r, i in zip(src_rank, src_row)])

    def push_rows(self, dst_rank, dst_row, rows):
        self.pushed += len(dst_rank)
        for k, (r, i) in enumerate(zip(dst_rank, dst_row)):
            self.buffers[int(r)][int(i)] = rows[k]

    def fence(self, d):
        d.barrier()


def _peer_worker(rank, world, port, tmp, buffers):
    sys.path.insert(0, ROOT)
    import importlib
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sharded = importlib.import_module("codename-rvc-fork-3_b200.sharded")
    synth = importlib.import_module("codename-rvc-fork-3_b200.synth")
    pr = synth.make_host_problem(3000, 64, 0, 48, seed=5, dup_frac=0.01)
    x, cent = pr.x, pr.centroids
    full = oracle.OracleIndex.from_vectors(x, cent)
    lo, hi = synth.shard_ranges(np.diff(full.list_off), world)[rank]
    sh = sharded.ShardedIndex(OracleShardBackend(oracle, x, cent, lo, hi))
    peer = sh.enable_peer(0, transport=SharedMemoryPeer(buffers, rank))
    rng = np.random.default_rng(100 + rank)
    for step, row0 in enumerate((0, 7 + rank)):                   # ragged batches at different offsets of the buffers
        n_r = 0 if (step == 1 and rank == world - 1) else 40 + 13 * rank
        q = (x[rng.integers(0, len(x), n_r)] + rng.standard_normal((n_r, 64)).astype(np.float32) * 0.05).astype(np.float32)
        if rank == 0:
            q[3, 0] = np.nan                                        # invalid assignment (-1): answered by shard 0
        feats = peer.rows(n_r, row0)
        feats.copy_(torch.from_numpy(q))
        assert n_r == 0 or peer.row_offset(feats) == row0
        dist.barrier()
        before = (peer.pulled, peer.pushed)
        out = sh.retrieve_blend_(feats, 0.75)
        assert out.data_ptr() == feats.data_ptr()
        with np.errstate(all="ignore"):
            want = oracle.tail(*full.search(q), x, 0.75, q) if n_r else q
        assert np.array_equal(feats.numpy(), want, equal_nan=True)
        assert peer.pulled - before[0] == peer.pushed - before[1]  # every pulled row was pushed back exactly once
        tot = torch.tensor([peer.pulled - before[0], n_r])
        dist.all_reduce(tot)
        assert int(tot[0]) == int(tot[1])                           # and every query of every rank was served once
    # a tensor outside the peer buffer on ONE rank sends every rank down the all-to-all route (collective choice)
    q = (x[rng.integers(0, len(x), 20)]).astype(np.float32) + np.float32(0.01)
    feats = torch.from_numpy(q.copy()) if rank == 0 else peer.rows(20, 0).copy_(torch.from_numpy(q))
    before = peer.pulled
    sh.retrieve_blend_(feats, 0.75)
    with np.errstate(all="ignore"):
        assert np.array_equal(feats.numpy(), oracle.tail(*full.search(q), x, 0.75, q), equal_nan=True)
    assert peer.pulled == before
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmp, f"peer_ok{rank}"), "w").write("ok")


@pytest.mark.parametrize("world", [2, 3, 8])
def test_peer_pull_push_step_on_gloo(world, tmp_path, oracle):
    """The peer-memory step (owner pulls rows from the home buffers, blends, pushes them back) == single index.
    world 8 = the ranks of one B200 box (48 lists over 8 owners: some shards own very few lists)."""
    buffers = [torch.zeros((256, 64), dtype=torch.float32).share_memory_() for _ in range(world)]
    mp.spawn(_peer_worker, args=(world, _free_port(), str(tmp_path), buffers), nprocs=world, join=True)
    assert all((tmp_path / f"peer_ok{r}").exists() for r in range(world))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    import importlib
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sharded = importlib.import_module("codename-rvc-fork-3_b200.sharded")
    synth = importlib.import_module("codename-rvc-fork-3_b200.synth")
    pr = synth.make_host_problem(3000, 64, 0, 48, seed=5, dup_frac=0.01)
    x, cent = pr.x, pr.centroids
    full = oracle.OracleIndex.from_vectors(x, cent)
    ranges = synth.shard_ranges(np.diff(full.list_off), world)
    lo, hi = ranges[rank]
    be = OracleShardBackend(oracle, x, cent, lo, hi)
    sh = sharded.ShardedIndex(be)
    assert sh.ranges == [tuple(r) for r in ranges]
    # ragged home batches: rank r brings 40 + 13 r queries (one of them NaN on rank 0)
    rng = np.random.default_rng(100 + rank)
    n_r = 40 + 13 * rank
    q = (x[rng.integers(0, len(x), n_r)] + rng.standard_normal((n_r, 64)).astype(np.float32) * 0.05).astype(np.float32)
    if rank == 0:
        q[3, 0] = np.nan
    Dw, Iw = full.search(q)
    with np.errstate(all="ignore"):
        want = oracle.tail(Dw, Iw, x, 0.75, q)
    feats = torch.from_numpy(q.copy())
    out, D, I = sh.retrieve_blend_(feats, 0.75, want_neighbours=True)
    assert out.data_ptr() == feats.data_ptr()
    np.testing.assert_array_equal(D.numpy(), Dw)
    np.testing.assert_array_equal(I.numpy(), Iw)
    assert np.array_equal(out.numpy(), want, equal_nan=True)
    Ds, Is = sh.search(torch.from_numpy(q.copy()))
    np.testing.assert_array_equal(Ds.numpy(), Dw)
    np.testing.assert_array_equal(Is.numpy(), Iw)
    # host-buffer entry point (on CPU: same step on a copy, written back in place)
    fh = torch.from_numpy(q.copy())
    assert sh.retrieve_blend_host_(fh, 0.75) is fh
    assert np.array_equal(fh.numpy(), want, equal_nan=True)
    # an empty home batch on one rank
    q2 = q[:0] if rank == world - 1 else q[:5]
    f2 = torch.from_numpy(q2.copy())
    sh.retrieve_blend_(f2, 0.3)
    if len(q2):
        with np.errstate(all="ignore"):
            w2 = oracle.tail(*full.search(q2), x, 0.3, q2)
        assert np.array_equal(f2.numpy(), w2, equal_nan=True)
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_equals_single_on_gloo(world, tmp_path, oracle):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_shard_ranges_cover_and_balance():
    import importlib
    sys.path.insert(0, ROOT)
    synth = importlib.import_module("codename-rvc-fork-3_b200.synth")
    rng = np.random.default_rng(0)
    sizes = rng.integers(0, 500, 1000)
    for world in (1, 2, 4, 8):
        rs = synth.shard_ranges(sizes, world)
        assert rs[0][0] == 0 and rs[-1][1] == 1000 and all(a[1] == b[0] for a, b in zip(rs[:-1], rs[1:]))
        tot = [int(sizes[a:b].sum()) for a, b in rs]
        assert max(tot) - min(tot) <= 2 * sizes.max()
