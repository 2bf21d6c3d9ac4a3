"""IVF lists sharded across the GPUs of one box: one process per GPU, torch.distributed (NCCL over
NVLink / NVSwitch) for the plumbing, librvcr kernels for every number.

The reference has no multi-GPU retrieval (the path is single-process CPU faiss); this is
the B200-native widening BASELINE.json's north_star names.  Partition: contiguous centroid-id ranges
balanced by vector count (rvcr_index_read_shard / synth.shard_ranges); ids and list ids stay global.

One retrieve_blend step, queries sharded by "home" rank (each rank brings its own clips' frames):

  1. all-gather the query rows                     (n x d fp32; the input exchange, NVLink)
  2. every shard: nearest LOCAL centroid per query (tcgen05 coarse over nlist/G centroids: the dominant
                                                    cost, so it divides by G)
  3. all-gather (distance, list id)                (8 B / query / shard -- the "small all-gather")
  4. merge: global argmin by (canonical distance, list id)   -> exactly the single-GPU assignment,
                                                    because the canonical distance is bit-identical on
                                                    every shard
  5. the owner shard scans its list + blends       (scan and gather traffic divide by G too)
  6. blended rows (and the top-8) return to the query's home rank with one all-to-all; `search()`
     instead all-gathers every shard's local top-8 and merges the candidates (north star wording)

No floating-point reduction crosses ranks, so G-GPU results equal 1-GPU results bit for bit.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

FLT_MAX = 3.4028234663852886e38


class CudaShardBackend:
    """Compute backend over the C ABI (the product path)."""

    def __init__(self, index):
        from ._ffi import check, lib
        self.index = index
        self._check, self._lib = check, lib
        inf = index.info
        self.list_begin, self.nlist, self.d, self.device = inf["list_begin"], inf["nlist"], inf["d"], inf["device"]

    def _stream(self, t):
        import torch
        return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)

    def assign(self, q):
        return self.index.assign(q)

    def merge_coarse(self, dist_all, list_all):
        import torch
        G, n = dist_all.shape
        out = torch.empty(n, dtype=torch.int32, device=dist_all.device)
        self._check(self._lib().rvcr_merge_coarse(C.c_void_p(dist_all.data_ptr()), C.c_void_p(list_all.data_ptr()), G, n,
                                                  C.c_void_p(out.data_ptr()), self.device, self._stream(out)))
        return out

    def retrieve_blend_assigned(self, feats, assign, index_rate, blend=True):
        import torch
        n = feats.shape[0]
        D = torch.empty((n, 8), dtype=torch.float32, device=feats.device)
        I = torch.empty((n, 8), dtype=torch.int64, device=feats.device)
        if blend:
            self._check(self._lib().rvcr_retrieve_blend_assigned(
                self.index.handle, C.c_void_p(feats.data_ptr()), n, C.c_void_p(assign.data_ptr()), float(index_rate),
                C.c_void_p(D.data_ptr()), C.c_void_p(I.data_ptr()), self._stream(feats)))
        else:
            self._check(self._lib().rvcr_search_assigned(
                self.index.handle, C.c_void_p(feats.data_ptr()), n, C.c_void_p(assign.data_ptr()),
                C.c_void_p(D.data_ptr()), C.c_void_p(I.data_ptr()), self._stream(feats)))
        return D, I

    def merge_topk(self, D_all, I_all):
        import torch
        G, n, k = D_all.shape
        D = torch.empty((n, k), dtype=torch.float32, device=D_all.device)
        I = torch.empty((n, k), dtype=torch.int64, device=D_all.device)
        self._check(self._lib().rvcr_merge_topk(C.c_void_p(D_all.data_ptr()), C.c_void_p(I_all.data_ptr()), G, n, k,
                                                C.c_void_p(D.data_ptr()), C.c_void_p(I.data_ptr()), self.device,
                                                self._stream(D)))
        return D, I


class ShardedIndex:
    """A shard of an IVF-Flat index plus the collectives that make G shards answer as one index."""

    def __init__(self, backend, group=None):
        import torch.distributed as dist
        self.be = backend
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        import torch
        dev = self._device()
        mine = torch.tensor([backend.list_begin, backend.nlist], dtype=torch.int64, device=dev)
        allr = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allr, mine, group=group)
        self.ranges = [(int(t[0]), int(t[0] + t[1])) for t in allr]
        for a, b in zip(self.ranges[:-1], self.ranges[1:]):
            if a[1] != b[0]:
                raise ValueError(f"shard list ranges are not contiguous: {self.ranges}")
        self.bounds = torch.tensor([r[0] for r in self.ranges[1:]], dtype=torch.int32, device=dev)
        self.last_counts = None

    def _device(self):
        import torch
        d = getattr(self.be, "device", None)
        return torch.device("cpu") if d is None or d == "cpu" else torch.device(f"cuda:{d}")

    # -- collectives -------------------------------------------------------------------------
    def _gather_rows(self, x, nmax):
        """all-gather of (n_r, ...) tensors padded to nmax rows -> (G, nmax, ...)."""
        import torch
        pad = torch.zeros((nmax,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        pad[: x.shape[0]] = x
        out = torch.empty((self.world, nmax) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        self.dist.all_gather_into_tensor(out.view(self.world * nmax, *x.shape[1:]) if x.dim() > 1 else out.view(-1), pad,
                                         group=self.group)
        return out

    def _counts(self, n_local, device):
        import torch
        t = torch.tensor([n_local], dtype=torch.int64, device=device)
        allc = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(allc, t, group=self.group)
        return [int(c) for c in allc]

    def _owner(self, assign):
        """owner shard of each global list id (-1 for an invalid assignment)."""
        import torch
        own = torch.bucketize(assign, self.bounds, right=True).to(torch.int32)
        return torch.where(assign < 0, torch.full_like(own, -1), own)

    # -- the step ----------------------------------------------------------------------------
    def _coarse_global(self, q_all):
        """steps 2-4: (N, d) replicated queries -> replicated global assignment (N,) int32."""
        import torch
        N = q_all.shape[0]
        lst, dis = self.be.assign(q_all)
        G = self.world
        dist_all = torch.empty((G, N), dtype=torch.float32, device=q_all.device)
        list_all = torch.empty((G, N), dtype=torch.int32, device=q_all.device)
        self.dist.all_gather_into_tensor(dist_all.view(-1), dis.contiguous(), group=self.group)
        self.dist.all_gather_into_tensor(list_all.view(-1), lst.contiguous(), group=self.group)
        return self.be.merge_coarse(dist_all, list_all)

    def retrieve_blend_(self, feats, index_rate: float, counts=None, want_neighbours: bool = False):
        """In place on this rank's (n_r, d) float32 rows; every rank of the group must call it.
        Returns feats (and this rank's (D, I) rows when asked)."""
        import torch
        G, r = self.world, self.rank
        n_local, d = feats.shape
        counts = self._counts(n_local, feats.device) if counts is None else counts
        nmax = max(counts)
        if nmax == 0:
            return (feats, None, None) if want_neighbours else feats
        q_all = self._gather_rows(feats, nmax).view(G * nmax, d)                    # 1
        assign = self._coarse_global(q_all)                                         # 2-4
        valid = torch.zeros(G * nmax, dtype=torch.bool, device=feats.device)
        for h in range(G):
            valid[h * nmax: h * nmax + counts[h]] = True
        assign = torch.where(valid, assign, torch.full_like(assign, -1))             # pad rows: no owner
        D, I = self.be.retrieve_blend_assigned(q_all, assign, index_rate)            # 5 (owned rows only)
        owner = self._owner(assign)
        home = torch.arange(G * nmax, device=feats.device, dtype=torch.int32) // nmax
        # 6: rows go home.  Owned rows in ascending query order are already grouped by home rank.
        mat = torch.bincount((owner.to(torch.int64) * G + home.to(torch.int64))[owner >= 0], minlength=G * G)
        mat = mat.view(G, G).cpu()                                                   # [owner, home]
        send_idx = torch.nonzero(owner == r, as_tuple=False).flatten()
        send_splits = [int(mat[r, h]) for h in range(G)]
        recv_splits = [int(mat[o, r]) for o in range(G)]
        send_rows = q_all.index_select(0, send_idx)
        recv_rows = torch.empty((sum(recv_splits), d), dtype=feats.dtype, device=feats.device)
        self.dist.all_to_all_single(recv_rows, send_rows, recv_splits, send_splits, group=self.group)
        owner_home = owner[r * nmax: r * nmax + n_local]
        has_owner = owner_home >= 0
        pos = torch.argsort(torch.where(has_owner, owner_home, torch.full_like(owner_home, G)), stable=True)
        pos = pos[: int(has_owner.sum())]
        feats.index_copy_(0, pos, recv_rows)
        # rows without a valid list (NaN query): the reference yields an all-pad search row -> NaN blend
        if (~has_owner).any():
            feats[~has_owner] = float("nan")
        if not want_neighbours:
            return feats
        sendD, sendI = D.index_select(0, send_idx), I.index_select(0, send_idx)
        recvD = torch.empty((sum(recv_splits), 8), dtype=D.dtype, device=D.device)
        recvI = torch.empty((sum(recv_splits), 8), dtype=I.dtype, device=I.device)
        self.dist.all_to_all_single(recvD, sendD, recv_splits, send_splits, group=self.group)
        self.dist.all_to_all_single(recvI, sendI, recv_splits, send_splits, group=self.group)
        Dh = torch.full((n_local, 8), FLT_MAX, dtype=D.dtype, device=D.device)
        Ih = torch.full((n_local, 8), -1, dtype=I.dtype, device=I.device)
        Dh.index_copy_(0, pos, recvD)
        Ih.index_copy_(0, pos, recvI)
        return feats, Dh, Ih

    def search(self, q, counts=None):
        """index.search(q, k=8) over all shards: every shard returns its local top-8 for the lists it
        owns, the candidates are all-gathered (96 B / query / shard) and merged by (distance, id).
        Returns this rank's (D, I) rows."""
        import torch
        G, r = self.world, self.rank
        n_local, d = q.shape
        counts = self._counts(n_local, q.device) if counts is None else counts
        nmax = max(counts)
        q_all = self._gather_rows(q, nmax).view(G * nmax, d)
        assign = self._coarse_global(q_all)
        D, I = self.be.retrieve_blend_assigned(q_all, assign, 0.0, blend=False)
        D_all = torch.empty((G,) + tuple(D.shape), dtype=D.dtype, device=D.device)
        I_all = torch.empty((G,) + tuple(I.shape), dtype=I.dtype, device=I.device)
        self.dist.all_gather_into_tensor(D_all.view(-1), D.view(-1), group=self.group)
        self.dist.all_gather_into_tensor(I_all.view(-1), I.view(-1), group=self.group)
        Dm, Im = self.be.merge_topk(D_all, I_all)
        return Dm[r * nmax: r * nmax + n_local], Im[r * nmax: r * nmax + n_local]
