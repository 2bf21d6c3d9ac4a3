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
#
# Every quantity is a pure function of (seed, global list id / query block id), so a shard built for
# lists [lo, hi) is bit-identical to the same lists of the full index and the query set does not
# depend on the number of GPUs.

LIST_BLOCK = 256        # lists whose rows share one generator stream
QUERY_BLOCK = 65536     # queries per generator stream


def list_sizes(n_total: int, nlist: int, seed: int = 1234, skew: float = 0.35) -> np.ndarray:
    """Log-normal list sizes (sigma = skew) rescaled to sum exactly to n_total; host, deterministic."""
    rng = np.random.default_rng(seed + 77)
    w = np.exp(skew * rng.standard_normal(nlist))
    sizes = np.floor(w / w.sum() * n_total).astype(np.int64)
    sizes[: int(n_total - sizes.sum())] += 1
    return sizes


def shard_ranges(sizes: np.ndarray, world: int):
    """Contiguous list ranges balanced by vector count -- same rule as csrc/index_io.cu shard_range."""
    nlist = len(sizes)
    total = int(sizes.sum())
    pre = np.concatenate([[0], np.cumsum(sizes)])
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        bounds.append(int(np.searchsorted(pre[:-1], target, side="left")) if total else nlist * r // world)
    bounds.append(nlist)
    return [(bounds[r], max(bounds[r], bounds[r + 1])) for r in range(world)]


def all_centroids(nlist: int, d: int, device, seed: int = 1234, sigma_c: float = 0.30, sigma_s: float = 0.03):
    """nlist // 4 blob centres ~ N(0, sigma_c^2); centroid l = centre[l // 4] + N(0, sigma_s^2): four
    sibling centroids per blob whose mutual distance (~2 d sigma_s^2 = 1.4 at d = 768) is comparable to
    the bf16 error margin of the tensor-core coarse path, so its exact re-check is exercised."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed * 1000003 + 17)
    nb = (nlist + 3) // 4
    centres = torch.randn((nb, d), generator=g, device=device, dtype=torch.float32) * sigma_c
    sib = torch.randn((nlist, d), generator=g, device=device, dtype=torch.float32) * sigma_s
    return centres[torch.arange(nlist, device=device) // 4] + sib


def fill_index_by_construction(index, cent_all, sizes: np.ndarray, n_total: int, list_range=None, seed: int = 1234,
                               sigma_p: float = 0.14, dup_every: int = 1000):
    """`index` = index_factory(d, f"IVF{nlist},Flat") on its GPU.  List l holds sizes[l] points
    = centroid_l + N(0, sigma_p^2); ids = global row position (ascending inside a list, the only order
    property faiss guarantees).  One row in every `dup_every` (0.1 %) is an exact copy of another row of
    the same list -- true distance ties inside a list, which the top-8 must break by list position.  With
    list_range=(lo, hi) only those lists are stored (a shard); the quantizer is always complete."""
    import torch
    dev = cent_all.device
    nlist, d = cent_all.shape
    lo, hi = (0, nlist) if list_range is None else list_range
    offs_all = np.zeros(nlist + 1, np.int64)
    np.cumsum(sizes, out=offs_all[1:])
    local = np.zeros(nlist, np.int64)
    local[lo:hi] = sizes[lo:hi]
    index.alloc_lists(local, int(n_total))
    index.set_centroids(cent_all.cpu().numpy())
    if (lo, hi) != (0, nlist):
        index.set_shard(lo, hi)
    vecs, ids, _, _ = index.device_views()
    g = torch.Generator(device=dev)
    base = int(offs_all[lo])
    for b0 in range((lo // LIST_BLOCK) * LIST_BLOCK, hi, LIST_BLOCK):
        b1 = min(b0 + LIST_BLOCK, nlist)
        rows = int(offs_all[b1] - offs_all[b0])
        if rows == 0:
            continue
        g.manual_seed(seed * 1000003 + 9_000_011 + b0 // LIST_BLOCK)
        noise = torch.randn((rows, d), generator=g, device=dev, dtype=torch.float32) * sigma_p
        if dup_every and rows >= dup_every:
            # exact duplicates: destination row k * dup_every + jitter (distinct by construction), source = a random row of
            # the same list; drawn from the block's own stream over the WHOLE block, so a shard reproduces them
            nd = rows // dup_every
            dst = torch.arange(nd, device=dev) * dup_every + torch.randint(0, dup_every, (nd,), generator=g, device=dev)
            boff = torch.as_tensor(offs_all[b0:b1 + 1] - offs_all[b0], device=dev)
            lst = torch.searchsorted(boff, dst, right=True) - 1
            u = torch.rand((nd,), generator=g, device=dev, dtype=torch.float64)
            src = boff[lst] + (u * (boff[lst + 1] - boff[lst]).to(torch.float64)).to(torch.int64)
            noise[dst] = noise[src]                    # the right-hand side is gathered before the scatter
            del dst, lst, u, src, boff
        l0, l1 = max(b0, lo), min(b1, hi)
        r0, r1 = int(offs_all[l0] - offs_all[b0]), int(offs_all[l1] - offs_all[b0])
        if r1 > r0:
            rep = torch.repeat_interleave(torch.arange(l0, l1, device=dev), torch.as_tensor(sizes[l0:l1], device=dev))
            dst = int(offs_all[l0]) - base
            vecs[dst:dst + (r1 - r0)] = cent_all[rep] + noise[r0:r1]
            ids[dst:dst + (r1 - r0)] = torch.arange(int(offs_all[l0]), int(offs_all[l1]), device=dev)
            del rep
        del noise
    index.commit()


def zipf_list_cdf(nlist: int, s: float, device, seed: int = 4321):
    """Cumulative distribution over list ids with mass proportional to (rank + 1)^-s under a fixed random ranking of
    the lists (heavy lists land anywhere in the id range, hence on any shard).  s = 0: uniform."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed * 1000003 + 555)
    rank_of = torch.randperm(nlist, generator=g, device=device)
    w = (rank_of.to(torch.float64) + 1.0) ** (-float(s))
    cdf = torch.cumsum(w / w.sum(), 0)
    cdf[-1] = 1.0
    return cdf


def device_queries(cent_all, nq: int, q_range=None, seed: int = 4321, sigma_p: float = 0.14, zipf: float = 0.0,
                   near_frac: float = 0.0, sizes=None, sigma_near: float = 0.05):
    """Queries of the bench, generated per block of QUERY_BLOCK so any contiguous range reproduces the same rows
    whatever the number of GPUs.  Query i draws a list (uniform, or Zipf with exponent `zipf` over a fixed random ranking
    of the lists: the skewed variant).  A fresh query is centroid[list] + N(0, sigma_p^2).  With near_frac > 0 (and
    `sizes` = the list sizes of the index) that fraction of the queries is marked "near": the caller replaces them by a
    stored row of that list + N(0, sigma_near^2) (`apply_near_rows`); returned as (q, near) with near = dict(mask, gpos,
    noise) or None."""
    import torch
    dev = cent_all.device
    nlist, d = cent_all.shape
    q0, q1 = (0, nq) if q_range is None else q_range
    out = torch.empty((q1 - q0, d), dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev)
    cdf = zipf_list_cdf(nlist, zipf, dev, seed) if zipf > 0 else None
    want_near = near_frac > 0 and sizes is not None
    if want_near:
        offs = torch.as_tensor(np.concatenate([[0], np.cumsum(sizes)]), device=dev)
        mask = torch.zeros(q1 - q0, dtype=torch.bool, device=dev)
        gpos = torch.zeros(q1 - q0, dtype=torch.int64, device=dev)
    for b0 in range((q0 // QUERY_BLOCK) * QUERY_BLOCK, q1, QUERY_BLOCK):
        g.manual_seed(seed * 1000003 + b0 // QUERY_BLOCK)
        if cdf is None:
            lists = torch.randint(0, nlist, (QUERY_BLOCK,), generator=g, device=dev)
        else:
            u = torch.rand((QUERY_BLOCK,), generator=g, device=dev, dtype=torch.float64)
            lists = torch.searchsorted(cdf, u).clamp_(max=nlist - 1)
        qb = cent_all[lists] + torch.randn((QUERY_BLOCK, d), generator=g, device=dev, dtype=torch.float32) * sigma_p
        a, b = max(b0, q0), min(b0 + QUERY_BLOCK, q1)
        if want_near:
            u = torch.rand((2, QUERY_BLOCK), generator=g, device=dev, dtype=torch.float64)
            ln = (offs[lists + 1] - offs[lists])
            nb = (u[0] < near_frac) & (ln > 0)
            pos = offs[lists] + torch.minimum((u[1] * ln.to(torch.float64)).to(torch.int64), (ln - 1).clamp_(min=0))
            # near queries: keep only the sigma_near noise here (rescaled from the fresh noise), the row is added later
            qb = torch.where(nb[:, None], (qb - cent_all[lists]) * (sigma_near / sigma_p), qb)
            mask[a - q0:b - q0] = nb[a - b0:b - b0]
            gpos[a - q0:b - q0] = pos[a - b0:b - b0]
        out[a - q0:b - q0] = qb[a - b0:b - b0]
    if want_near:
        return out, {"mask": mask, "gpos": gpos}
    return out


def apply_near_rows(q, near, vecs, row_base: int = 0, fetch=None):
    """q[i] += stored row at global list-ordered position near['gpos'][i] for the marked queries.  Single GPU: the rows
    are read from `vecs` (position - row_base).  Sharded: `fetch(gpos) -> rows` gets them from their owners."""
    idx = near["mask"].nonzero().squeeze(1)
    if idx.numel() == 0:
        return q
    gp = near["gpos"][idx]
    step = 1 << 18
    for s in range(0, idx.numel(), step):
        sl = slice(s, s + step)
        rows = fetch(gp[sl]) if fetch is not None else vecs[gp[sl] - row_base]
        q[idx[sl]] += rows
    return q
