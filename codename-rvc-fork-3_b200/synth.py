"""Seeded synthetic inputs of the shapes BASELINE.json names.  Host (numpy) generators
for the parity tests; an on-device (torch) by-construction generator for the large bench configs.

Index vectors follow a Gaussian mixture so that IVF is meaningful: blob centres ~ N(0, 0.30^2 I),
points = centre + N(0, 0.14^2 I)  (overall std ~ 0.33, ||x|| ~ 9 like the ContentVec fixtures in
the reference's logs/), plus a fraction of exact duplicate rows (true distance ties).
Queries: 80 % = an index row + N(0, 0.05^2 I) noise, 20 % fresh mixture samples.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

CONFIGS = {
    # name: (N, d, nq, nlist)   nlist = min(int(16*sqrt(N)), N//39)   extract_index.py:59
    "C1": (50_000, 768, 500, 1282),
    "C2": (200_000, 256, 3_000, 5128),
    "C3": (2_000_000, 768, 180_000, 22_627),
    "C4": (4_000_000, 768, 0, 32_000),
    "C5": (32_000_000, 768, 1_000_000, 90_509),
}


def nlist_rule(n: int) -> int:
    """extract_index.py:59"""
    return min(int(16 * math.sqrt(n)), n // 39)


@dataclass
class HostProblem:
    x: np.ndarray          # (N, d) index vectors in id order
    centroids: np.ndarray  # (nlist, d)
    q: np.ndarray          # (nq, d)


def mixture_points(rng, n: int, d: int, n_blobs: int, sigma_c: float = 0.30, sigma_p: float = 0.14):
    centres = (rng.standard_normal((n_blobs, d)) * sigma_c).astype(np.float32)
    which = rng.integers(0, n_blobs, size=n)
    x = centres[which] + (rng.standard_normal((n, d)) * sigma_p).astype(np.float32)
    return x.astype(np.float32), centres, which


def make_host_problem(n: int, d: int, nq: int, nlist: int, seed: int = 1234, dup_frac: float = 0.001,
                      iid: bool = False, lloyd_iters: int = 1) -> HostProblem:
    """Small / medium problems for parity tests (everything on the host)."""
    rng = np.random.default_rng(seed)
    if iid:
        x = (rng.standard_normal((n, d)) * 0.33).astype(np.float32)
    else:
        x, _, _ = mixture_points(rng, n, d, max(8, nlist // 4))
    ndup = int(n * dup_frac)
    if ndup:
        src = rng.integers(0, n, size=ndup)
        dst = rng.integers(0, n, size=ndup)
        x[dst] = x[src]
    # centroids: rows drawn without replacement + Lloyd iterations (numpy, BLAS distances)
    cent = x[rng.choice(n, size=nlist, replace=False)].copy()
    for _ in range(lloyd_iters):
        a = _assign_blas(x, cent)
        sums = np.zeros((nlist, d), np.float64)
        np.add.at(sums, a, x)
        cnt = np.bincount(a, minlength=nlist)
        nz = cnt > 0
        cent[nz] = (sums[nz] / cnt[nz, None]).astype(np.float32)
    rng_q = np.random.default_rng(seed + 3087)  # 4321 for the default seed
    n_near = int(nq * 0.8)
    rows = rng_q.integers(0, n, size=n_near)
    q_near = x[rows] + (rng_q.standard_normal((n_near, d)) * 0.05).astype(np.float32)
    if iid:
        q_far = (rng_q.standard_normal((nq - n_near, d)) * 0.33).astype(np.float32)
    else:
        q_far = x[rng_q.integers(0, n, size=nq - n_near)] + (rng_q.standard_normal((nq - n_near, d)) * 0.14).astype(np.float32)
    q = np.concatenate([q_near, q_far]).astype(np.float32)
    q = q[rng_q.permutation(nq)]
    return HostProblem(x=x, centroids=cent.astype(np.float32), q=np.ascontiguousarray(q))


def _assign_blas(x, cent, block: int = 8192):
    yn = np.einsum("ij,ij->i", cent, cent)
    out = np.empty(x.shape[0], np.int64)
    for s in range(0, x.shape[0], block):
        xb = x[s:s + block]
        out[s:s + block] = (yn[None, :] - 2.0 * (xb @ cent.T)).argmin(axis=1)
    return out


# ---- on-device by-construction generator (bench, large configs) --------------------------------

def device_index_by_construction(index, n_total: int, nlist: int, d: int, seed: int = 1234, sigma_c: float = 0.30,
                                 sigma_s: float = 0.05, sigma_p: float = 0.14, rows_per_chunk: int = 1 << 20,
                                 list_range: tuple[int, int] | None = None, skew: float = 0.35):
    """Fill `index` (created with index_factory(d, f"IVF{nlist_local},Flat")) directly on its GPU.

    Structure: nlist // 4 blob centres ~ N(0, sigma_c^2); every centroid = its blob centre +
    N(0, sigma_s^2) (4 sibling centroids per blob, a sibling gap comparable to the bf16 margin of the
    tensor-core coarse path, so that path's exact re-check is exercised); list l holds size[l] points
    = centroid_l + N(0, sigma_p^2).  List sizes are log-normal around n_total / nlist (sigma = `skew`).
    ids are a fixed pseudo-random permutation-like map (id = row position: ids ascend within a list,
    which is the only order property faiss guarantees).  Every quantity is derived from (seed, global
    list id), so a shard built with list_range=(lo, hi) is bit-identical to the same lists of the
    full index.  Returns (sizes_all int64[nlist] on host).
    """
    import torch
    dev = torch.device(f"cuda:{index.device}")
    g = torch.Generator(device="cpu").manual_seed(seed)
    # list sizes (host, deterministic): log-normal, rescaled to sum exactly n_total
    z = torch.randn(nlist, generator=g, dtype=torch.float64)
    w = torch.exp(skew * z)
    sizes = torch.floor(w / w.sum() * n_total).to(torch.int64)
    sizes[: int(n_total - int(sizes.sum()))] += 1
    sizes_np = sizes.numpy().copy()
    lo, hi = (0, nlist) if list_range is None else list_range
    offs_all = np.zeros(nlist + 1, np.int64)
    np.cumsum(sizes_np, out=offs_all[1:])
    local_sizes = sizes_np[lo:hi]
    index.alloc_lists(local_sizes, int(n_total))

    def centroids_for(l0: int, l1: int):
        out = torch.empty((l1 - l0, d), dtype=torch.float32, device=dev)
        gl = torch.Generator(device=dev)
        for l in range(l0, l1, 4096):
            m = min(4096, l1 - l)
            ids = torch.arange(l, l + m, device=dev)
            blob = ids // 4
            # blob centre: seeded by blob id block so that it is independent of the chunking
            b0, b1 = int(blob[0]), int(blob[-1]) + 1
            gl.manual_seed(seed * 1000003 + 17 + b0 // 1024)
            base_blob = (b0 // 1024) * 1024
            nb = ((b1 - base_blob + 1023) // 1024) * 1024
            centres = torch.randn((nb, d), generator=gl, device=dev, dtype=torch.float32) * sigma_c
            if nb > 1024:  # crossing a 1024-blob block: regenerate per block for determinism
                parts = []
                for bb in range(base_blob, base_blob + nb, 1024):
                    gl.manual_seed(seed * 1000003 + 17 + bb // 1024)
                    parts.append(torch.randn((1024, d), generator=gl, device=dev, dtype=torch.float32) * sigma_c)
                centres = torch.cat(parts)
            gl.manual_seed(seed * 1000003 + 7_000_003 + l // 4096)
            sib = torch.randn((4096, d), generator=gl, device=dev, dtype=torch.float32)[: m] * sigma_s
            out[l - l0:l - l0 + m] = centres[blob - base_blob] + sib
        return out

    cent = centroids_for(lo, hi)
    index.set_centroids(cent.cpu().numpy())
    vecs, ids, offsets, _ = index.device_views()
    gl = torch.Generator(device=dev)
    local_off = np.zeros(hi - lo + 1, np.int64)
    np.cumsum(local_sizes, out=local_off[1:])
    # rows are generated per list block of 256 lists: deterministic regardless of sharding
    LB = 256
    for l in range((lo // LB) * LB, hi, LB):
        l0, l1 = max(l, lo), min(l + LB, hi)
        if l0 >= l1:
            continue
        blk_sizes = sizes_np[l:min(l + LB, nlist)]
        blk_rows = int(blk_sizes.sum())
        gl.manual_seed(seed * 1000003 + 9_000_011 + l // LB)
        noise = torch.randn((blk_rows, d), generator=gl, device=dev, dtype=torch.float32) * sigma_p
        blk_off = np.zeros(len(blk_sizes) + 1, np.int64)
        np.cumsum(blk_sizes, out=blk_off[1:])
        r0, r1 = int(blk_off[l0 - l]), int(blk_off[l1 - l])
        rep = torch.repeat_interleave(torch.arange(l0 - lo, l1 - lo, device=dev),
                                      torch.as_tensor(sizes_np[l0:l1], device=dev))
        dst0 = int(local_off[l0 - lo])
        vecs[dst0:dst0 + (r1 - r0)] = cent[rep] + noise[r0:r1]
        ids[dst0:dst0 + (r1 - r0)] = torch.arange(int(offs_all[l0]), int(offs_all[l0]) + (r1 - r0), device=dev)
        del noise, rep
    index.commit()
    return sizes_np


def device_queries(index_vecs, nstored: int, nq: int, d: int, seed: int = 4321, sigma_q: float = 0.05,
                   device=None):
    """nq queries = random stored rows + N(0, sigma_q^2) noise, generated on the GPU."""
    import torch
    dev = index_vecs.device if device is None else device
    gl = torch.Generator(device=dev).manual_seed(seed)
    rows = torch.randint(0, nstored, (nq,), generator=gl, device=dev)
    q = index_vecs[rows] + torch.randn((nq, d), generator=gl, device=dev, dtype=torch.float32) * sigma_q
    return q.contiguous()
