"""IVF lists sharded across the GPUs of one box: one process per GPU, torch.distributed (NCCL over
NVLink / NVSwitch) for the plumbing, librvcr kernels for every number.

The reference has no multi-GPU retrieval (the path is single-process CPU faiss); this is
the B200-native widening BASELINE.json's north_star names.  Partition: the inverted lists (98 GB at
32M x 768) are split into contiguous list-id ranges balanced by vector count (rvcr_index_read_shard /
synth.shard_ranges); the coarse quantizer (nlist x d centroids, 278 MB at most) is replicated, so any
rank can assign any query and list ids are global.

One retrieve_blend step, queries sharded by "home" rank (each rank brings its own clips' frames):

  1. home rank: nearest centroid of its own queries   tcgen05 coarse over nq/G queries x ALL centroids:
                                                      the dominant cost divides by G, and the result IS
                                                      the single-GPU assignment (same kernel, same data)
  2. dispatch: all-to-all of the query rows (+ list ids) to the shard that owns the assigned list
                                                      3 KB / query, once, over NVLink
  3. owner: list scan + top-8 + gather + blend        scan and gather traffic divide by G too
  4. combine: all-to-all of the blended rows (and optionally the top-8) back to the home rank

`search()` is the variant the north star words: every shard returns its local top-8 for all queries
(real candidates only for the lists it owns), the candidates are all-gathered (96 B / query / shard)
and merged by (distance, id).

With peer memory enabled (`ShardedIndex.enable_peer`: every rank's feature buffer and an INBOX are cudaMalloc
allocations mapped into all other ranks with cudaIpc over NVLink / NVSwitch) steps 2 and 4 need no collective and no
host synchronisation: the home rank's dispatch kernel files (list id, home rank, home row) into the owner's inbox
(remote atomics reserve the slots, NVLink stores fill them), a flag barrier kernel says "everyone has dispatched", the
owner PULLS its rows out of their home ranks' buffers and the blend epilogue of step 3 PUSHES each blended row
straight into its home rank's buffer -- compute and combine in one kernel, no send / recv staging, no scatter pass --
and a second flag barrier says "every row is home" (csrc/peer.cu).  The NCCL all-to-all transport stays for tensors
outside the peer buffer and for callers that want the neighbour lists back.

No floating-point reduction crosses ranks, so G-GPU results equal 1-GPU results bit for bit.
"""
from __future__ import annotations

import ctypes as C
import os

FLT_MAX = 3.4028234663852886e38


class CudaShardBackend:
    """Compute backend over the C ABI (the product path)."""

    def __init__(self, index):
        from ._ffi import check, lib
        self.index = index
        self._check, self._lib = check, lib
        inf = index.info
        self.owned = (inf["owned_begin"], inf["owned_end"])
        self.nlist, self.d, self.device = inf["nlist"], inf["d"], inf["device"]

    def _stream(self, t):
        import torch
        return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)

    def assign(self, q):
        return self.index.assign(q)

    def scan_assigned(self, feats, assign, index_rate, blend=True):
        """top-8 (and, with blend, the in-place weighted blend) for the rows whose list this shard owns."""
        import torch
        n = feats.shape[0]
        D = torch.empty((n, 8), dtype=torch.float32, device=feats.device)
        I = torch.empty((n, 8), dtype=torch.int64, device=feats.device)
        if n == 0:
            return D, I
        if blend:
            self._check(self._lib().rvcr_retrieve_blend_assigned(
                self.index.handle, C.c_void_p(feats.data_ptr()), n, C.c_void_p(assign.data_ptr()), float(index_rate),
                C.c_void_p(D.data_ptr()), C.c_void_p(I.data_ptr()), self._stream(feats)))
        else:
            self._check(self._lib().rvcr_search_assigned(
                self.index.handle, C.c_void_p(feats.data_ptr()), n, C.c_void_p(assign.data_ptr()),
                C.c_void_p(D.data_ptr()), C.c_void_p(I.data_ptr()), self._stream(feats)))
        return D, I

    def scan_assigned_push(self, rows, assign, index_rate, peer, dst_rank, dst_row):
        """Step 3 + 4 fused: scan + top-8 + gather + blend of the pulled rows; the blend epilogue writes row i to
        peer buffer dst_rank[i], row dst_row[i] (NVLink stores) instead of `rows`."""
        m = rows.shape[0]
        if m == 0:
            return
        self._check(self._lib().rvcr_retrieve_blend_assigned_push(
            self.index.handle, C.c_void_p(rows.data_ptr()), m, C.c_void_p(assign.data_ptr()), float(index_rate),
            C.c_void_p(peer.bases_dev.data_ptr()), C.c_void_p(dst_rank.data_ptr()), C.c_void_p(dst_row.data_ptr()),
            self._stream(rows)))

    def scan_inbox_push(self, peer, index_rate):
        """Steps 3 + 4 of the device-driven step: seal the inbox, pull the filed rows, scan + top-8 + gather + blend with
        the epilogue writing every row to its home buffer; counts stay on the device."""
        import torch
        st = C.c_void_p(torch.cuda.current_stream(torch.device(f"cuda:{peer.device}")).cuda_stream)
        scratch = peer.rows_scratch_tensor() if peer.pull_rows_first else None
        self._check(self._lib().rvcr_retrieve_blend_inbox_push(
            self.index.handle, C.c_void_p(peer.inbox_ptrs[peer.rank]), peer.inbox_cap,
            C.c_void_p(scratch.data_ptr()) if scratch is not None else None,
            float(index_rate), C.c_void_p(peer.bases_dev.data_ptr()), st))

    def merge_topk(self, D_all, I_all):
        import torch
        G, n, k = D_all.shape
        D = torch.empty((n, k), dtype=torch.float32, device=D_all.device)
        I = torch.empty((n, k), dtype=torch.int64, device=D_all.device)
        if n:
            self._check(self._lib().rvcr_merge_topk(C.c_void_p(D_all.data_ptr()), C.c_void_p(I_all.data_ptr()), G, n, k,
                                                    C.c_void_p(D.data_ptr()), C.c_void_p(I.data_ptr()), self.device,
                                                    self._stream(D)))
        return D, I


class CudaPeer:
    """One feature buffer per rank (rvcr_peer_alloc), IPC-mapped into every other rank of the group.

    `rows(n, slot)` hands out (n, d) float32 tensors inside the own buffer; `pull_rows` gathers rows out of any
    rank's buffer through the peer mapping.  Construction is collective; it raises on every rank if any rank cannot
    map its peers (no NVLink / P2P between the GPUs, IPC disabled), so the caller can stay on the NCCL transport."""

    INBOX_HEADER = 1024          # RVCR_PEER_INBOX_HEADER (include/rvcr.h)

    def __init__(self, capacity_rows: int, d: int, device: int, group=None, max_batch_rows: int | None = None):
        """capacity_rows: rows of this rank's feature buffer.  max_batch_rows: the largest batch ANY rank passes to one
        retrieve_blend_ step (default: capacity_rows); the inbox and the owner's row scratch hold world x that many
        entries, so no step can overflow them whatever the query mass looks like."""
        import torch
        import torch.distributed as dist
        from ._ffi import check, lib
        from .faiss_compat import _view
        self._lib, self._check, self._view = lib, check, _view
        self.d, self.device, self.group = int(d), int(device), group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.capacity_rows = int(capacity_rows)
        self.nbytes = self.capacity_rows * self.d * 4
        self.max_batch_rows = int(max_batch_rows if max_batch_rows is not None else capacity_rows)
        self.inbox_cap = self.world * self.max_batch_rows
        self.inbox_bytes = self.INBOX_HEADER + 12 * self.inbox_cap
        self._own = C.c_void_p()
        self._own_inbox = C.c_void_p()
        self._opened = []
        handle, handle_in = (C.c_ubyte * 64)(), (C.c_ubyte * 64)()
        err = None
        try:
            check(lib().rvcr_peer_alloc(self.device, self.nbytes, C.byref(self._own), handle))
            check(lib().rvcr_peer_alloc(self.device, self.inbox_bytes, C.byref(self._own_inbox), handle_in))
        except Exception as ex:                                    # keep the collective below in step on every rank
            err = ex
        mine = (bytes(handle), self.nbytes, err is None, bytes(handle_in), self.inbox_cap)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        self.ptrs = [0] * self.world
        self.inbox_ptrs = [0] * self.world
        if err is None and all(e[2] for e in everyone) and any(e[4] != self.inbox_cap for e in everyone):
            err = RuntimeError("the ranks disagree on max_batch_rows")
        if err is None and all(e[2] for e in everyone):
            for r, (hd, nb, _, hd_in, _) in enumerate(everyone):
                if r == self.rank:
                    self.ptrs[r] = self._own.value
                    self.inbox_ptrs[r] = self._own_inbox.value
                    continue
                try:
                    for hdl, dst in ((hd, self.ptrs), (hd_in, self.inbox_ptrs)):
                        p = C.c_void_p()
                        check(lib().rvcr_peer_open(self.device, (C.c_ubyte * 64).from_buffer_copy(hdl), C.byref(p)))
                        self._opened.append(p)
                        dst[r] = p.value
                except Exception as ex:
                    err = ex
                    break
        elif err is None:
            err = RuntimeError("a peer rank could not allocate its buffer")
        ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=f"cuda:{self.device}")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            self.close()
            raise RuntimeError(f"peer memory unavailable: {err if err else 'another rank failed to map its peers'}")
        self.sizes = [e[1] for e in everyone]
        self.bases_dev = torch.tensor(self.ptrs, dtype=torch.int64, device=f"cuda:{self.device}")
        self.inbox_bases_dev = torch.tensor(self.inbox_ptrs, dtype=torch.int64, device=f"cuda:{self.device}")
        self.pull_rows_first = os.environ.get("RVCR_PEER_PULL") == "1"   # True: the owner copies the filed rows into a local scratch before scanning (A/B knob)
        self._rows_scratch = None
        self._flag = torch.zeros(1, dtype=torch.int32, device=f"cuda:{self.device}")
        self._epoch = 0
        dist.barrier(group=group)                                 # every inbox is zeroed and mapped before the first step

    def rows_scratch_tensor(self):
        import torch
        if self._rows_scratch is None:
            self._rows_scratch = torch.empty((self.inbox_cap, self.d), dtype=torch.float32, device=f"cuda:{self.device}")
        return self._rows_scratch

    def dispatch(self, assign, row0: int, bounds):
        """File this rank's queries (list ids `assign`, rows row0.. of the own feature buffer) into their owners' inboxes."""
        import torch
        n = assign.shape[0]
        if n > self.max_batch_rows:
            raise ValueError(f"batch of {n} rows exceeds max_batch_rows={self.max_batch_rows} given to enable_peer")
        if n == 0:
            return
        st = C.c_void_p(torch.cuda.current_stream(assign.device).cuda_stream)
        self._check(self._lib().rvcr_peer_dispatch(C.c_void_p(assign.data_ptr()), n, int(row0), self.rank, self.world,
                                                   C.c_void_p(bounds.data_ptr()) if self.world > 1 else None,
                                                   C.c_void_p(self.inbox_bases_dev.data_ptr()), self.inbox_cap, self.device, st))

    def barrier(self):
        """Stream-ordered barrier over the ranks (a kernel that stores / spins on flags in peer memory; no host sync)."""
        import torch
        self._epoch += 1
        st = C.c_void_p(torch.cuda.current_stream(torch.device(f"cuda:{self.device}")).cuda_stream)
        self._check(self._lib().rvcr_peer_barrier(C.c_void_p(self.inbox_bases_dev.data_ptr()), self.rank, self.world,
                                                  self._epoch, self.device, st))

    def rows(self, n: int, row0: int = 0):
        """(n, d) float32 tensor over rows [row0, row0 + n) of this rank's buffer (no copy)."""
        if row0 < 0 or n < 0 or row0 + n > self.capacity_rows:
            raise ValueError(f"rows [{row0}, {row0 + n}) outside the peer buffer of {self.capacity_rows} rows")
        return self._view(self.ptrs[self.rank] + row0 * self.d * 4, (n, self.d), "<f4", self.device)

    def row_offset(self, t):
        """Row index of tensor t inside this rank's buffer, or None if t lives elsewhere."""
        if not (t.is_cuda and t.dim() == 2 and t.shape[1] == self.d and t.is_contiguous()):
            return None
        off = t.data_ptr() - self.ptrs[self.rank]
        if off < 0 or off % (self.d * 4) or off + t.shape[0] * self.d * 4 > self.nbytes:
            return None
        return off // (self.d * 4)

    def pull_rows(self, src_rank, src_row):
        """rows[i] = buffer[src_rank[i]][src_row[i]] (int32 device tensors) -> new local (m, d) tensor."""
        import torch
        m = src_rank.shape[0]
        out = torch.empty((m, self.d), dtype=torch.float32, device=src_rank.device)
        if m:
            st = C.c_void_p(torch.cuda.current_stream(out.device).cuda_stream)
            self._check(self._lib().rvcr_peer_pull_rows(C.c_void_p(self.bases_dev.data_ptr()), C.c_void_p(src_rank.data_ptr()),
                                                        C.c_void_p(src_row.data_ptr()), m, self.d, C.c_void_p(out.data_ptr()),
                                                        self.device, st))
        return out

    def fence(self, dist):
        """Stream-ordered cross-rank barrier: when it completes, every rank's pushes issued before its own fence
        have landed (kernel completion makes peer stores visible)."""
        dist.all_reduce(self._flag, group=self.group)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def close(self):
        for p in self._opened:
            try:
                self._lib().rvcr_peer_close(self.device, p)
            except Exception:
                pass
        self._opened = []
        if self._own.value:
            self._lib().rvcr_peer_free(self.device, self._own)
            self._own = C.c_void_p()
        if getattr(self, "_own_inbox", None) is not None and self._own_inbox.value:
            self._lib().rvcr_peer_free(self.device, self._own_inbox)
            self._own_inbox = C.c_void_p()


class ShardedIndex:
    """A shard of an IVF-Flat index plus the collectives that make G shards answer as one index."""
    peer = None

    def __init__(self, backend, group=None):
        import torch
        import torch.distributed as dist
        self.be = backend
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        dev = self._device()
        mine = torch.tensor(list(backend.owned), dtype=torch.int64, device=dev)
        allr = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allr, mine, group=group)
        self.ranges = [(int(t[0]), int(t[1])) for t in allr]
        if self.ranges[0][0] != 0 or self.ranges[-1][1] != backend.nlist or \
                any(a[1] != b[0] for a, b in zip(self.ranges[:-1], self.ranges[1:])):
            raise ValueError(f"shard list ranges do not tile [0, {backend.nlist}): {self.ranges}")
        self.bounds = torch.tensor([r[0] for r in self.ranges[1:]], dtype=torch.int32, device=dev)

    def _device(self):
        import torch
        d = getattr(self.be, "device", None)
        return torch.device("cpu") if d is None or d == "cpu" else torch.device(f"cuda:{d}")

    def _owner(self, assign):
        """owner shard of each list id; invalid assignments (-1, a NaN query) go to shard 0, whose scan
        answers them with the all-pad row the reference produces."""
        import torch
        own = torch.bucketize(assign, self.bounds, right=True).to(torch.int64)
        return torch.where(assign < 0, torch.zeros_like(own), own)

    def _exchange_counts(self, send_counts):
        """send_counts: (G,) int64 device tensor -> (send list, recv list) on the host."""
        import torch
        recv = torch.empty_like(send_counts)
        self.dist.all_to_all_single(recv, send_counts, group=self.group)
        both = torch.stack([send_counts, recv]).cpu()
        return [int(x) for x in both[0]], [int(x) for x in both[1]]

    # -- the step ----------------------------------------------------------------------------
    def retrieve_blend_(self, feats, index_rate: float, want_neighbours: bool = False):
        """In place on this rank's (n_r, d) float32 rows; every rank of the group must call it.
        Returns feats (and this rank's (D, I) rows when asked)."""
        import torch
        n = feats.shape[0]
        if n:
            assign, _ = self.be.assign(feats)                                            # 1
        else:
            assign = torch.empty(0, dtype=torch.int32, device=feats.device)
        if self.peer is not None and hasattr(self.peer, "dispatch"):
            # device-driven step.  The transport choice must be the same on every rank and nothing is exchanged to agree on
            # it: neighbours wanted -> everyone takes the NCCL route (the flag is part of the call); rows outside the peer
            # buffer are an error, not a silent fallback.
            if want_neighbours:
                return self._route_scan_combine(feats, assign, index_rate, True)
            row0 = 0 if n == 0 else self.peer.row_offset(feats)
            if row0 is None:
                raise ValueError("peer transport: the rows must live in this rank's peer buffer (ShardedIndex.peer.rows(n, row0))")
            return self._peer_inbox_step(feats, assign, row0, index_rate)
        if self.peer is not None:                                  # host-driven variant (stand-in transports of the CPU tests)
            row0 = None if want_neighbours else (0 if n == 0 else self.peer.row_offset(feats))   # an empty batch fits anywhere
            return self._peer_scan_push(feats, assign, -1 if row0 is None else row0, index_rate, want_neighbours)
        return self._route_scan_combine(feats, assign, index_rate, want_neighbours)

    def _peer_inbox_step(self, feats, assign, row0, index_rate):
        """Steps 2-4 with no host synchronisation and no collective (csrc/peer.cu): dispatch kernel -> flag barrier -> owner
        seals its inbox, pulls, scans, blends and pushes -> flag barrier.  feats = rows [row0, row0 + n) of this rank's peer
        buffer, blended in place by whichever ranks own the lists."""
        p = self.peer
        p.dispatch(assign, row0, self.bounds)                      # 2: (list, home rank, home row) -> the owner's inbox
        p.barrier()                                                # everyone's coarse step and dispatch are done
        self.be.scan_inbox_push(p, index_rate)                     # 3 + 4 fused
        p.barrier()                                                # every blended row has landed in its home buffer
        return feats

    def enable_peer(self, capacity_rows: int, d: int | None = None, transport=None, max_batch_rows: int | None = None):
        """Collective.  Allocate a peer-mapped feature buffer of `capacity_rows` rows (and an inbox for world x
        `max_batch_rows` entries) on every rank; tensors handed out by `self.peer.rows(n, row0)` then take the pull / push
        path in `retrieve_blend_`.  Raises (on every rank) when the GPUs cannot map each other's memory."""
        self.peer = transport if transport is not None else CudaPeer(capacity_rows, d or self.be.d, self.be.device, self.group,
                                                                     max_batch_rows=max_batch_rows)
        return self.peer

    def transport_name(self) -> str:
        if self.peer is None:
            return "nccl all-to-all (dispatch + combine)"
        if hasattr(self.peer, "dispatch"):
            return ("peer memory over NVLink (cudaIpc), device-driven: dispatch kernel files (list, home row) into the owner's inbox, "
                    "flag barrier, owner's scan kernels read the rows over NVLink, blend epilogue pushes them home, flag barrier; "
                    "no host sync, no collective")
        return "peer memory over NVLink (cudaIpc): owner pulls rows, blend epilogue pushes them home"

    def _peer_scan_push(self, feats, assign, row0, index_rate, want_neighbours=False):
        """Steps 2-4 over peer memory.  feats = rows [row0, row0 + n) of this rank's peer buffer, blended in place by
        whichever ranks own the lists.  row0 = -1 (tensor outside the peer buffer, or neighbours wanted) on ANY rank
        sends every rank down the NCCL all-to-all route: the transport choice is collective."""
        import torch
        G, me, dev = self.world, self.rank, feats.device
        n = feats.shape[0]
        meta = torch.tensor([n, row0], dtype=torch.int64, device=dev)
        metas = torch.empty((G, 2), dtype=torch.int64, device=dev)
        self.dist.all_gather_into_tensor(metas.view(-1), meta, group=self.group)
        metas_h = metas.tolist()                                   # host waits here for the coarse step (nothing else to do)
        if any(int(x[1]) < 0 for x in metas_h):
            return self._route_scan_combine(feats, assign, index_rate, want_neighbours)
        nmax = max(int(x[0]) for x in metas_h)
        if nmax == 0:
            return feats
        apad = torch.full((nmax,), -2, dtype=torch.int32, device=dev)              # -2 = padding, -1 = invalid (NaN) query
        apad[:n] = assign
        a_all = torch.empty((G, nmax), dtype=torch.int32, device=dev)
        # 4 bytes per query; also the barrier: nobody scans (and pushes into a home buffer) before everyone's coarse is done
        self.dist.all_gather_into_tensor(a_all.view(-1), apad, group=self.group)
        mine = (self._owner(a_all.view(-1)).view(G, nmax) == me) & (a_all != -2)
        sel = mine.nonzero()                                        # (m, 2): (home rank, row in the home batch), home-major
        src_rank = sel[:, 0].to(torch.int32)
        base = torch.tensor([int(x[1]) for x in metas_h], dtype=torch.int64, device=dev)
        src_row = (sel[:, 1] + base[sel[:, 0]]).to(torch.int32)
        rows = self.peer.pull_rows(src_rank, src_row)                                # 2: pull over NVLink
        rows_assign = a_all[sel[:, 0], sel[:, 1]].contiguous()
        self.be.scan_assigned_push(rows, rows_assign, index_rate, self.peer, src_rank, src_row)   # 3 + 4 fused
        self.peer.fence(self.dist)                                  # all pushes into my buffer have landed
        return feats

    def retrieve_blend_host_(self, host, index_rate: float, chunk_rows: int = 131072, device_rows=None):
        """Same step on this rank's HOST rows (a pinned (n_r, d) float32 torch tensor), in place: the rows go up in
        chunks on a copy stream while the tensor cores already assign the previous chunk (the coarse step is
        per-query independent, rvcr_retrieve_blend_host does the same on one GPU); dispatch / scan / combine run on
        the whole batch; one device -> host copy brings the blended rows back.  Every rank must call it.
        `device_rows`: optional (n, d) device staging tensor -- rows of the peer buffer (`self.peer.rows`) make the
        step take the pull / push transport."""
        import torch
        n, d = host.shape
        dev = self._device()
        if dev.type != "cuda":
            feats = host.clone()
            self.retrieve_blend_(feats, index_rate)
            host.copy_(feats)
            return host
        feats = device_rows if device_rows is not None else torch.empty((n, d), dtype=torch.float32, device=dev)
        assert tuple(feats.shape) == (n, d)
        main = torch.cuda.current_stream(dev)
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(dev)
        offs, off, c = [], 0, min(32768, chunk_rows)
        while off < n:                       # chunk sizes ramp up: only a small first copy is exposed
            offs.append((off, min(n, off + c)))
            off += c
            c = min(chunk_rows, c * 2)
        events = []
        self._copy_stream.wait_stream(main)
        with torch.cuda.stream(self._copy_stream):
            for a, b in offs:
                feats[a:b].copy_(host[a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
                events.append(ev)
        parts = []
        for (a, b), ev in zip(offs, events):
            main.wait_event(ev)
            parts.append(self.be.assign(feats[a:b])[0])                                  # 1, chunk by chunk
        assign = torch.cat(parts) if parts else torch.empty(0, dtype=torch.int32, device=dev)
        if self.peer is not None and hasattr(self.peer, "dispatch"):
            row0 = 0 if n == 0 else self.peer.row_offset(feats)
            if row0 is None:
                raise ValueError("peer transport: device_rows must be rows of this rank's peer buffer")
            self._peer_inbox_step(feats, assign, row0, index_rate)
        elif self.peer is not None:
            row0 = 0 if n == 0 else self.peer.row_offset(feats)
            self._peer_scan_push(feats, assign, -1 if row0 is None else row0, index_rate)
        else:
            self._route_scan_combine(feats, assign, index_rate, False)
        host.copy_(feats, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return host

    def _route_scan_combine(self, feats, assign, index_rate, want_neighbours):
        import torch
        G = self.world
        n, d = feats.shape
        owner = self._owner(assign)
        order = torch.argsort(owner, stable=True)                                        # rows grouped by owner
        send_counts = torch.bincount(owner, minlength=G)
        send_splits, recv_splits = self._exchange_counts(send_counts)
        m = sum(recv_splits)
        send_rows = feats.index_select(0, order)
        send_assign = assign.index_select(0, order)
        recv_rows = torch.empty((m, d), dtype=feats.dtype, device=feats.device)
        recv_assign = torch.empty(m, dtype=assign.dtype, device=feats.device)
        self.dist.all_to_all_single(recv_rows, send_rows, recv_splits, send_splits, group=self.group)      # 2
        self.dist.all_to_all_single(recv_assign, send_assign, recv_splits, send_splits, group=self.group)
        D, I = self.be.scan_assigned(recv_rows, recv_assign, index_rate, blend=True)    # 3 (in place on recv_rows)
        self.dist.all_to_all_single(send_rows, recv_rows, send_splits, recv_splits, group=self.group)      # 4
        feats.index_copy_(0, order, send_rows)
        if not want_neighbours:
            return feats
        Dh = torch.empty((n, 8), dtype=D.dtype, device=D.device)
        Ih = torch.empty((n, 8), dtype=I.dtype, device=I.device)
        self.dist.all_to_all_single(Dh, D, send_splits, recv_splits, group=self.group)
        self.dist.all_to_all_single(Ih, I, send_splits, recv_splits, group=self.group)
        Do, Io = torch.empty_like(Dh), torch.empty_like(Ih)
        Do.index_copy_(0, order, Dh)
        Io.index_copy_(0, order, Ih)
        return feats, Do, Io

    def search(self, q):
        """index.search(q, k=8) over all shards, the way the north star words it: queries and their
        list ids are all-gathered, each shard returns its local top-8 (pads for lists it does not own),
        the candidates are all-gathered and merged by (distance, id).  Returns this rank's (D, I) rows."""
        import torch
        G, r = self.world, self.rank
        n, d = q.shape
        cnt = torch.tensor([n], dtype=torch.int64, device=q.device)
        allc = [torch.empty_like(cnt) for _ in range(G)]
        self.dist.all_gather(allc, cnt, group=self.group)
        counts = [int(c) for c in allc]
        nmax = max(counts)
        if nmax == 0:
            return (torch.empty((0, 8), dtype=torch.float32, device=q.device),
                    torch.empty((0, 8), dtype=torch.int64, device=q.device))
        qpad = torch.zeros((nmax, d), dtype=q.dtype, device=q.device)
        apad = torch.full((nmax,), -1, dtype=torch.int32, device=q.device)
        if n:
            qpad[:n] = q
            apad[:n] = self.be.assign(q)[0]
        q_all = torch.empty((G * nmax, d), dtype=q.dtype, device=q.device)
        a_all = torch.empty(G * nmax, dtype=torch.int32, device=q.device)
        self.dist.all_gather_into_tensor(q_all, qpad, group=self.group)
        self.dist.all_gather_into_tensor(a_all, apad, group=self.group)
        D, I = self.be.scan_assigned(q_all, a_all, 0.0, blend=False)
        D_all = torch.empty((G,) + tuple(D.shape), dtype=D.dtype, device=D.device)
        I_all = torch.empty((G,) + tuple(I.shape), dtype=I.dtype, device=I.device)
        self.dist.all_gather_into_tensor(D_all.view(-1), D.view(-1), group=self.group)
        self.dist.all_gather_into_tensor(I_all.view(-1), I.view(-1), group=self.group)
        Dm, Im = self.be.merge_topk(D_all, I_all)
        return Dm[r * nmax: r * nmax + n], Im[r * nmax: r * nmax + n]
