// scan_topk.cu -- inverted-list scan + top-8 selection (nprobe = 1).
//
// Replaces, for rvc/infer/pipeline.py:499 `index.search(npy, k=8)`, the faiss stage
// IndexIVF::search_preassigned -> IVFFlatScanner<METRIC_L2>::scan_codes -> heap.
//
// faiss re-reads a list once per query that probes it.  Here queries are first grouped by their
// probed list (stable radix sort of (list, query) pairs), then one CTA takes a work item
// = (list, up to 16 of its queries): the list streams HBM -> shared memory once per item and all
// 16 queries are scored against each staged chunk, so a list probed by <= 16 queries is read from
// HBM exactly once.  Distances are bit-identical to the canonical fvec_L2sqr order (pairdist.cuh);
// the running top-8 of a query lives in lanes 0..7 of one warp, ordered by (distance, list
// position) -- the strict-">" heap of faiss keeps the earlier-scanned entry on ties, which is the
// same thing.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include "pairdist.cuh"

namespace rvcr {

constexpr int SC_TQ = 4;
constexpr int SC_QT = 4 * SC_TQ;  // 16 queries per item
constexpr int SC_NSTAGE = 8;
constexpr int SC_DSTRIDE = 33;

// ------------------------------------------------------------------ grouping by list --------

__global__ void make_keys_kernel(const int32_t* __restrict__ assign, int64_t n, int64_t own_lo, int64_t own_hi,
                                 int64_t nlist, uint32_t* __restrict__ keys, int32_t* __restrict__ vals) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t a = (int64_t)assign[i];
    // nlist = invalid assignment (NaN query): answered with an all-pad row; nlist + 1 = another shard's list
    keys[i] = a < 0 ? (uint32_t)nlist : (a < own_lo || a >= own_hi) ? (uint32_t)nlist + 1u : (uint32_t)a;
    vals[i] = (int32_t)i;
}

// flags[p] = 1 iff sorted position p starts a work item: first of its list segment, or a multiple
// of SC_QT past the segment start.  Positions whose key == nlist + 1 (a list another shard owns) form
// no item.
__global__ void flag_items_kernel(const uint32_t* __restrict__ keys, int64_t n, uint32_t nlist, int skip_foreign,
                                  uint8_t* __restrict__ flags) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint32_t k = keys[p];
    if (skip_foreign && k == nlist + 1u) {
        flags[p] = 0;
        return;
    }
    // lower_bound of k in keys[0..p]
    int64_t lo = 0, hi = p;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < k) lo = mid + 1; else hi = mid;
    }
    flags[p] = ((p - lo) % SC_QT) == 0;
}

__global__ void reset_counters_kernel(int32_t* c, int nc) {
    if (threadIdx.x < nc) c[threadIdx.x] = 0;
}

// ------------------------------------------------------------------------ scan kernel --------

__global__ void __launch_bounds__(PD_THREADS, 1)
scan_topk_kernel_v1(const float* __restrict__ q, const int32_t* __restrict__ qorder, const uint32_t* __restrict__ keys,
                 const int32_t* __restrict__ items, const int32_t* __restrict__ nitems_ptr, int32_t* __restrict__ ticket,
                 const float* __restrict__ vecs, const int64_t* __restrict__ ids, const int64_t* __restrict__ offsets,
                 int64_t nstored, uint32_t nlist, int64_t n, int d, float* __restrict__ D, int32_t* __restrict__ P,
                 int64_t* __restrict__ I) {
    extern __shared__ __align__(16) float smem[];
    const int QS = d + 8;
    float* sQ = smem;                            // SC_QT x QS
    float* sR = sQ + SC_QT * QS;                 // SC_NSTAGE x 32 x 128
    float* sD = sR + SC_NSTAGE * PD_RB * PD_KC;  // SC_QT x 33
    __shared__ int s_item;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int l = tid & 7, qg = (tid >> 3) & 3;
    const int nkc = d / PD_KC;
    const int nitems = *nitems_ptr;

    for (;;) {
        __syncthreads();  // previous item fully consumed (sQ, sR, sD, s_item)
        if (tid == 0) s_item = atomicAdd(ticket, 1);
        __syncthreads();
        const int it_idx = s_item;
        if (it_idx >= nitems) break;

        const int64_t p0 = items[it_idx];
        const uint32_t key = keys[p0];
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < SC_QT; ++i) cnt += (p0 + i < n && keys[p0 + i] == key) ? 1 : 0;

        int64_t lb = 0, len = 0;
        if (key < nlist) {
            lb = offsets[key];
            len = offsets[key + 1] - lb;
        }

        // running top-8 of the two queries this warp owns (slot u -> query w + 8u); lanes 0..7
        float td[2] = {RVCR_FLT_MAX, RVCR_FLT_MAX};
        int tp[2] = {0x7fffffff, 0x7fffffff};

        if (len > 0) {
            const int nblk = (int)((len + PD_RB - 1) / PD_RB);
            const int total = nblk * nkc;
            const float* lbase = vecs + lb * d;
            const int64_t rmax = len - 1;

            // prologue: queries + first NSTAGE-1 stages
            {
                const int nvec = d >> 2;
                for (int idx = tid; idx < SC_QT * nvec; idx += PD_THREADS) {
                    int r = idx / nvec, c = idx - r * nvec;
                    int rr = r < cnt ? r : cnt - 1;
                    int64_t qi = qorder[p0 + rr];
                    cp_async16(sQ + r * QS + c * 4, q + qi * d + c * 4);
                }
            }
#pragma unroll
            for (int s = 0; s < SC_NSTAGE - 1; ++s) {
                if (s < total)
                    pd_load_stage(sR + s * PD_RB * PD_KC, lbase, (int64_t)(s / nkc) * PD_RB, rmax, d, s % nkc, tid);
                cp_async_commit();
            }

            PairTile<SC_TQ> tile;
            tile.clear();
            int kc = 0, rb = 0;
            for (int it = 0; it < total; ++it) {
                cp_async_wait<SC_NSTAGE - 2>();
                __syncthreads();
                {
                    int nx = it + SC_NSTAGE - 1;
                    if (nx < total)
                        pd_load_stage(sR + (nx % SC_NSTAGE) * PD_RB * PD_KC, lbase, (int64_t)(nx / nkc) * PD_RB, rmax,
                                      d, nx % nkc, tid);
                    cp_async_commit();
                }
                tile.chunk(sQ + qg * QS + kc * PD_KC + l, QS, sR + (it % SC_NSTAGE) * PD_RB * PD_KC + (4 * w) * PD_KC + l);
                if (++kc == nkc) {
                    kc = 0;
                    tile.reduce();
#pragma unroll
                    for (int a = 0; a < SC_TQ; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b)
                            if (((a * 4 + b) & 7) == l) sD[(qg + 4 * a) * SC_DSTRIDE + 4 * w + b] = tile.acc[a][b];
                    tile.clear();
                    __syncthreads();
                    const int64_t rowi = (int64_t)rb * PD_RB + lane;  // row within list
                    const bool rvalid = rowi < len;
                    const int cpos = (int)(lb + rowi);
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int qq = w + 8 * u;
                        if (qq < cnt) {  // warp-uniform
                            const float cd = sD[qq * SC_DSTRIDE + lane];
                            const float thr_d = __shfl_sync(0xffffffffu, td[u], 7);
                            const int thr_p = __shfl_sync(0xffffffffu, tp[u], 7);
                            unsigned m = __ballot_sync(0xffffffffu, rvalid && cand_less(cd, cpos, thr_d, thr_p));
                            while (m) {
                                const int src = __ffs(m) - 1;
                                m &= m - 1;
                                const float xd = __shfl_sync(0xffffffffu, cd, src);
                                const int xp = __shfl_sync(0xffffffffu, cpos, src);
                                const bool lt = cand_less(td[u], tp[u], xd, xp);
                                const int rank = __popc(__ballot_sync(0xffffffffu, lt) & 0xffu);
                                const float ud = __shfl_up_sync(0xffffffffu, td[u], 1);
                                const int up = __shfl_up_sync(0xffffffffu, tp[u], 1);
                                if (lane < 8) {
                                    if (lane == rank) {
                                        td[u] = xd;
                                        tp[u] = xp;
                                    } else if (lane > rank) {
                                        td[u] = ud;
                                        tp[u] = up;
                                    }
                                }
                            }
                        }
                    }
                    ++rb;
                }
            }
            cp_async_wait<0>();
        }

        // write-out: (FLT_MAX, -1) padding for short / empty / foreign lists
        if (lane < RVCR_K) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int qq = w + 8 * u;
                if (qq < cnt) {
                    const int64_t qi = qorder[p0 + qq];
                    const bool pad = tp[u] == 0x7fffffff;
                    D[qi * RVCR_K + lane] = pad ? RVCR_FLT_MAX : td[u];
                    P[qi * RVCR_K + lane] = pad ? -1 : tp[u];
                    if (I) I[qi * RVCR_K + lane] = pad ? (int64_t)-1 : ids[tp[u]];
                }
            }
        }
    }
}

// ------------------------------------------------------------------ scan kernel, v2 ---------
//
// Same work items and the same canonical distances as v1, re-mapped for the Blackwell FP32 pipe:
//   * a thread owns TWO adjacent accumulator lanes (2l', 2l'+1) of a pair, so queries and rows are
//     read with LDS.64 and the (x - y), fma(t, t, acc) steps are packed sub.rn.f32x2 / fma.rn.f32x2
//     (FADD2 / FFMA2): half the issue slots for the same 128 fp32 lane-ops/clk/SM;
//   * 4 threads per pair, 64 pair-groups per CTA; a warp = 4 query groups x 2 row groups, so every
//     LDS.64 is one shared-memory wavefront (queries: 4 x 32 B distinct; rows: 2 x 32 B distinct,
//     row stride padded to 136 floats so the two row groups hit different banks);
//   * the CTA tile is (4*TQ queries) x 64 rows with TQ in {1,2,3,4} chosen per work item from its
//     query count, so a list probed by 5 queries does 8-query work, not 16-query work.
// Per (query,row) pair-element: 1 packed instruction + 0.5 LDS.64 at TQ = 4.

constexpr int S2_RB = 64;             // rows per block
constexpr int S2_RS = PD_KC + 8;      // padded stage row stride (floats)
constexpr int S2_NSTAGE = 4;
constexpr int S2_DSTRIDE = 65;

typedef unsigned long long u64;
__device__ __forceinline__ u64 f2_sub(u64 a, u64 b) {
    u64 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 f2_fma(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 f2_add(u64 a, u64 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

__device__ __forceinline__ void s2_load_stage(float* sStage, const float* __restrict__ lbase, int64_t row0, int64_t rmax,
                                              int d, int kc, int tid) {
    const int c = tid & 31, r0 = tid >> 5;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int r = r0 + 8 * j;
        int64_t gr = row0 + r;
        if (gr > rmax) gr = rmax;
        cp_async16(sStage + r * S2_RS + c * 4, lbase + gr * d + kc * PD_KC + c * 4);
    }
}

template <int TQ>
__device__ __forceinline__ void s2_run_item(const float* __restrict__ q, const int32_t* __restrict__ qorder, int64_t p0,
                                            int cnt, const float* __restrict__ lbase, int64_t lb, int64_t len, int d,
                                            float* sQ, float* sR, float* sD, float (&td)[2], int (&tp)[2]) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int lp = lane & 3, qg = (lane >> 2) & 3, rg = lane >> 4;
    const int QS = d + 8;
    const int nkc = d / PD_KC;
    const int nblk = (int)((len + S2_RB - 1) / S2_RB);
    const int total = nblk * nkc;
    const int64_t rmax = len - 1;

    {   // queries of this item -> sQ (rows beyond cnt replicate the last query; never reported)
        const int nvec = d >> 2;
        for (int r = w; r < 4 * TQ; r += 8) {
            const int rr = r < cnt ? r : cnt - 1;
            const float* src = q + (int64_t)qorder[p0 + rr] * d;
            for (int c = lane; c < nvec; c += 32) cp_async16(sQ + r * QS + c * 4, src + c * 4);
        }
    }
#pragma unroll
    for (int s = 0; s < S2_NSTAGE - 1; ++s) {
        if (s < total) s2_load_stage(sR + s * S2_RB * S2_RS, lbase, (int64_t)(s / nkc) * S2_RB, rmax, d, s % nkc, tid);
        cp_async_commit();
    }

    u64 acc[TQ][4];
#pragma unroll
    for (int a = 0; a < TQ; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0ull;

    const float* myQ = sQ + qg * QS + 2 * lp;
    const int myRowOff = (8 * w + rg) * S2_RS + 2 * lp;
    int kc = 0, rb = 0, ld_kc = (S2_NSTAGE - 1) % nkc, ld_rb = (S2_NSTAGE - 1) / nkc;
    for (int it = 0; it < total; ++it) {
        cp_async_wait<S2_NSTAGE - 2>();
        __syncthreads();
        {
            const int nx = it + S2_NSTAGE - 1;
            if (nx < total)
                s2_load_stage(sR + (nx % S2_NSTAGE) * S2_RB * S2_RS, lbase, (int64_t)ld_rb * S2_RB, rmax, d, ld_kc, tid);
            cp_async_commit();
            if (++ld_kc == nkc) { ld_kc = 0; ++ld_rb; }
        }
        const float* sq = myQ + kc * PD_KC;
        const float* sr = sR + (it % S2_NSTAGE) * S2_RB * S2_RS + myRowOff;
#pragma unroll 4
        for (int s = 0; s < PD_KC / 8; ++s) {
            u64 qv[TQ], rv[4];
#pragma unroll
            for (int a = 0; a < TQ; ++a) qv[a] = *reinterpret_cast<const u64*>(sq + (4 * a) * QS + 8 * s);
#pragma unroll
            for (int b = 0; b < 4; ++b) rv[b] = *reinterpret_cast<const u64*>(sr + (2 * b) * S2_RS + 8 * s);
#pragma unroll
            for (int a = 0; a < TQ; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const u64 t = f2_sub(qv[a], rv[b]);
                    acc[a][b] = f2_fma(t, t, acc[a][b]);
                }
        }
        if (++kc == nkc) {
            kc = 0;
            // canonical reduction: lanes (2l', 2l'+1) per thread -> (a0+a4, a1+a5) | (a2+a6, a3+a7) -> m0+m1 | m2+m3 -> sum
#pragma unroll
            for (int a = 0; a < TQ; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    u64 v = acc[a][b];
                    v = f2_add(v, __shfl_xor_sync(0xffffffffu, v, 2));
                    float x = __uint_as_float((unsigned)(v & 0xffffffffull));
                    float y = __uint_as_float((unsigned)(v >> 32));
                    float sdist = __fadd_rn(x, y);
                    sdist = __fadd_rn(sdist, __shfl_xor_sync(0xffffffffu, sdist, 1));
                    if (((a * 4 + b) & 3) == lp) sD[(qg + 4 * a) * S2_DSTRIDE + 8 * w + 2 * b + rg] = sdist;
                    acc[a][b] = 0ull;
                }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int qq = w + 8 * u;
                if (qq < cnt) {  // warp-uniform
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int64_t rowi = (int64_t)rb * S2_RB + hh * 32 + lane;
                        top8_update(td[u], tp[u], sD[qq * S2_DSTRIDE + hh * 32 + lane], (int)(lb + rowi), rowi < len, lane);
                    }
                }
            }
            ++rb;
        }
    }
    cp_async_wait<0>();
}

__global__ void __launch_bounds__(PD_THREADS, 1)
scan_topk_kernel_v2(const float* __restrict__ q, const int32_t* __restrict__ qorder, const uint32_t* __restrict__ keys,
                    const int32_t* __restrict__ items, const int32_t* __restrict__ nitems_ptr, int32_t* __restrict__ ticket,
                    const float* __restrict__ vecs, const int64_t* __restrict__ ids, const int64_t* __restrict__ offsets,
                    int64_t nstored, uint32_t nlist, int64_t n, int d, float* __restrict__ D, int32_t* __restrict__ P,
                    int64_t* __restrict__ I) {
    extern __shared__ __align__(16) float smem[];
    const int QS = d + 8;
    float* sQ = smem;                              // 16 x QS
    float* sR = sQ + SC_QT * QS;                   // S2_NSTAGE x 64 x 136
    float* sD = sR + S2_NSTAGE * S2_RB * S2_RS;    // 16 x 65
    __shared__ int s_item;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int nitems = *nitems_ptr;

    for (;;) {
        __syncthreads();  // previous item fully consumed (sQ, sR, sD, s_item)
        if (tid == 0) s_item = atomicAdd(ticket, 1);
        __syncthreads();
        const int it_idx = s_item;
        if (it_idx >= nitems) break;

        const int64_t p0 = items[it_idx];
        const uint32_t key = keys[p0];
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < SC_QT; ++i) cnt += (p0 + i < n && keys[p0 + i] == key) ? 1 : 0;
        int64_t lb = 0, len = 0;
        if (key < nlist) {
            lb = offsets[key];
            len = offsets[key + 1] - lb;
        }
        float td[2] = {RVCR_FLT_MAX, RVCR_FLT_MAX};
        int tp[2] = {0x7fffffff, 0x7fffffff};
        if (len > 0) {
            const float* lbase = vecs + lb * d;
            if (cnt <= 4) s2_run_item<1>(q, qorder, p0, cnt, lbase, lb, len, d, sQ, sR, sD, td, tp);
            else if (cnt <= 8) s2_run_item<2>(q, qorder, p0, cnt, lbase, lb, len, d, sQ, sR, sD, td, tp);
            else if (cnt <= 12) s2_run_item<3>(q, qorder, p0, cnt, lbase, lb, len, d, sQ, sR, sD, td, tp);
            else s2_run_item<4>(q, qorder, p0, cnt, lbase, lb, len, d, sQ, sR, sD, td, tp);
        }
        if (lane < RVCR_K) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int qq = w + 8 * u;
                if (qq < cnt) {
                    const int64_t qi = qorder[p0 + qq];
                    const bool pad = tp[u] == 0x7fffffff;
                    D[qi * RVCR_K + lane] = pad ? RVCR_FLT_MAX : td[u];
                    P[qi * RVCR_K + lane] = pad ? -1 : tp[u];
                    if (I) I[qi * RVCR_K + lane] = pad ? (int64_t)-1 : ids[tp[u]];
                }
            }
        }
    }
}

// ------------------------------------------------------------------ scan kernel, v3 ---------
//
// v2's packed mapping with 16 warps per CTA (4 per scheduler instead of 2) so that the FADD2 -> FFMA2
// dependency and the LDS latency of one warp are covered by the other three.  Warps 0-7 own queries
// 0..7 of the item, warps 8-15 queries 8..15; every warp covers 8 of the 64 rows of a block; a thread
// holds TQ in {0,1,2} query slots (4*TQ queries per warp half), chosen per warp half from the item's
// query count.  Register budget <= 128 (512 threads x 128 = the whole register file).

constexpr int S3_THREADS = 512;

// rows [row0, row0 + 64) of the list, k-chunk kc -> one stage.  `gbase` = lbase + kc*128 + c*4 (this
// thread's column), computed by the caller; the clamp against the list end is only evaluated for the
// last row block (block-uniform `tail`).
__device__ __forceinline__ void s3_load_stage(float* sdst, const float* __restrict__ gbase, int64_t row0, int64_t rmax,
                                              int d, bool tail, int r0) {
    const float* g = gbase + row0 * d;
    if (!tail) {
#pragma unroll
        for (int j = 0; j < 4; ++j) cp_async16(sdst + (16 * j) * S2_RS, g + (r0 + 16 * j) * d);
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int64_t r = row0 + r0 + 16 * j;
            if (r > rmax) r = rmax;
            cp_async16(sdst + (16 * j) * S2_RS, gbase + r * d);
        }
    }
}

template <int TQ>
__device__ __forceinline__ void s3_run_item(const float* __restrict__ q, const int32_t* __restrict__ qorder, int64_t p0,
                                            int cnt, const float* __restrict__ lbase, int64_t lb, int64_t len, int d,
                                            float* sQ, float* sR, float* sD, float& td, int& tp) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int lp = lane & 3, qg = (lane >> 2) & 3, rg = lane >> 4;
    const int hq = w >> 3, wr = w & 7;
    const int QS = d + 8;
    const int nkc = d / PD_KC;
    const int nblk = (int)((len + S2_RB - 1) / S2_RB);
    const int total = nblk * nkc;
    const int64_t rmax = len - 1;

    {   // queries of this item -> sQ (rows beyond cnt replicate the last query; never reported)
        const int nvec = d >> 2;
        const int qt = cnt <= 8 ? 8 : 16;
        for (int r = w; r < qt; r += 16) {
            const int rr = r < cnt ? r : cnt - 1;
            const float* src = q + (int64_t)qorder[p0 + rr] * d;
            for (int c = lane; c < nvec; c += 32) cp_async16(sQ + r * QS + c * 4, src + c * 4);
        }
    }
    // this thread's copy slots: rows w, w+16, w+32, w+48 of a block, float4 column `lane`
    const float* gcol = lbase + lane * 4;
    float* scol = sR + w * S2_RS + lane * 4;
    int ld_kc = 0, ld_rb = 0;
#pragma unroll
    for (int s = 0; s < S2_NSTAGE - 1; ++s) {
        if (s < total) {
            s3_load_stage(scol + s * S2_RB * S2_RS, gcol + ld_kc * PD_KC, (int64_t)ld_rb * S2_RB, rmax, d, ld_rb == nblk - 1, w);
            if (++ld_kc == nkc) { ld_kc = 0; ++ld_rb; }
        }
        cp_async_commit();
    }

    constexpr int TQA = TQ > 0 ? TQ : 1;
    float2 acc[TQA][4];
#pragma unroll
    for (int a = 0; a < TQA; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = make_float2(0.f, 0.f);

    const float* myQ = sQ + (8 * hq + qg) * QS + 2 * lp;
    const int myRowOff = (8 * wr + rg) * S2_RS + 2 * lp;
    int kc = 0, rb = 0, slot = 0, ld_slot = S2_NSTAGE - 1;
    for (int it = 0; it < total; ++it) {
        cp_async_wait<S2_NSTAGE - 2>();
        __syncthreads();
        if (it + S2_NSTAGE - 1 < total) {
            s3_load_stage(scol + ld_slot * S2_RB * S2_RS, gcol + ld_kc * PD_KC, (int64_t)ld_rb * S2_RB, rmax, d,
                          ld_rb == nblk - 1, w);
            if (++ld_kc == nkc) { ld_kc = 0; ++ld_rb; }
        }
        cp_async_commit();
        if (++ld_slot == S2_NSTAGE) ld_slot = 0;
        if (TQ > 0) {
            const float2* sq = reinterpret_cast<const float2*>(myQ + kc * PD_KC);
            const float2* sr = reinterpret_cast<const float2*>(sR + slot * S2_RB * S2_RS + myRowOff);
#pragma unroll 8
            for (int s = 0; s < PD_KC / 8; ++s) {
                float2 qv[TQA], rv[4];
#pragma unroll
                for (int a = 0; a < TQA; ++a) qv[a] = sq[(4 * a) * (QS / 2) + 4 * s];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const float2 r = sr[(2 * b) * (S2_RS / 2) + 4 * s];
                    rv[b] = make_float2(-r.x, -r.y);
                }
#pragma unroll
                for (int a = 0; a < TQA; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const float2 t = __fadd2_rn(qv[a], rv[b]);          // x - y, both lanes, round-to-nearest
                        acc[a][b] = __ffma2_rn(t, t, acc[a][b]);            // fused, both lanes
                    }
            }
        }
        if (++slot == S2_NSTAGE) slot = 0;
        if (++kc == nkc) {
            kc = 0;
            if (TQ > 0) {
                // canonical reduction: (a0,a1)|(a2,a3)|(a4,a5)|(a6,a7) over 4 threads
                //   xor 2: (a0+a4, a1+a5) | (a2+a6, a3+a7);  in-thread: m0+m1 | m2+m3;  xor 1: (m0+m1)+(m2+m3)
#pragma unroll
                for (int a = 0; a < TQA; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        float2 v = acc[a][b];
                        float2 o;
                        o.x = __shfl_xor_sync(0xffffffffu, v.x, 2);
                        o.y = __shfl_xor_sync(0xffffffffu, v.y, 2);
                        v = __fadd2_rn(v, o);
                        float sdist = __fadd_rn(v.x, v.y);
                        sdist = __fadd_rn(sdist, __shfl_xor_sync(0xffffffffu, sdist, 1));
                        if (((a * 4 + b) & 3) == lp) sD[(8 * hq + qg + 4 * a) * S2_DSTRIDE + 8 * wr + 2 * b + rg] = sdist;
                        acc[a][b] = make_float2(0.f, 0.f);
                    }
            }
            __syncthreads();
            if (w < cnt) {  // warp-uniform: warp w keeps the top-8 of query w
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int64_t rowi = (int64_t)rb * S2_RB + hh * 32 + lane;
                    top8_update(td, tp, sD[w * S2_DSTRIDE + hh * 32 + lane], (int)(lb + rowi), rowi < len, lane);
                }
            }
            ++rb;
        }
    }
    cp_async_wait<0>();
}

__global__ void __launch_bounds__(S3_THREADS, 1)
scan_topk_kernel_v3(const float* __restrict__ q, const int32_t* __restrict__ qorder, const uint32_t* __restrict__ keys,
                    const int32_t* __restrict__ items, const int32_t* __restrict__ nitems_ptr, int32_t* __restrict__ ticket,
                    const float* __restrict__ vecs, const int64_t* __restrict__ ids, const int64_t* __restrict__ offsets,
                    int64_t nstored, uint32_t nlist, int64_t n, int d, float* __restrict__ D, int32_t* __restrict__ P,
                    int64_t* __restrict__ I) {
    extern __shared__ __align__(16) float smem[];
    const int QS = d + 8;
    float* sQ = smem;                              // 16 x QS
    float* sR = sQ + SC_QT * QS;                   // S2_NSTAGE x 64 x 136
    float* sD = sR + S2_NSTAGE * S2_RB * S2_RS;    // 16 x 65
    __shared__ int s_item;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int nitems = *nitems_ptr;

    for (;;) {
        __syncthreads();  // previous item fully consumed (sQ, sR, sD, s_item)
        if (tid == 0) s_item = atomicAdd(ticket, 1);
        __syncthreads();
        const int it_idx = s_item;
        if (it_idx >= nitems) break;

        const int64_t p0 = items[it_idx];
        const uint32_t key = keys[p0];
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < SC_QT; ++i) cnt += (p0 + i < n && keys[p0 + i] == key) ? 1 : 0;
        int64_t lb = 0, len = 0;
        if (key < nlist) {
            lb = offsets[key];
            len = offsets[key + 1] - lb;
        }
        float td = RVCR_FLT_MAX;
        int tp = 0x7fffffff;
        if (len > 0) {
            const float* lbase = vecs + lb * d;
            const int nloc = cnt - 8 * (w >> 3);   // queries in this warp half (warp-uniform)
            if (nloc <= 0) s3_run_item<0>(q, qorder, p0, cnt, lbase, lb, len, d, sQ, sR, sD, td, tp);
            else if (nloc <= 4) s3_run_item<1>(q, qorder, p0, cnt, lbase, lb, len, d, sQ, sR, sD, td, tp);
            else s3_run_item<2>(q, qorder, p0, cnt, lbase, lb, len, d, sQ, sR, sD, td, tp);
        }
        if (lane < RVCR_K && w < cnt) {
            const int64_t qi = qorder[p0 + w];
            const bool pad = tp == 0x7fffffff;
            D[qi * RVCR_K + lane] = pad ? RVCR_FLT_MAX : td;
            P[qi * RVCR_K + lane] = pad ? -1 : tp;
            if (I) I[qi * RVCR_K + lane] = pad ? (int64_t)-1 : ids[tp];
        }
    }
}

// ------------------------------------------------------------------ scan kernel, v4 ---------
//
// ncu on v2 / v3 showed the scan bound by the shared-memory pipe, not by HBM or the FP32 pipe: an LDS.64
// always costs two wavefronts (256 B requested), so a (TQ x 4) register tile moves 2-3 B of shared memory
// per pair-element against a 128 B/clk pipe.  v4 doubles the row dimension of the register tile:
//   thread tile  TQ x 8  (TQ in 1..4 query slots, 8 rows), two accumulator lanes per thread (float2,
//                FADD2/FFMA2) -> 1.5 B of LDS per pair-element at TQ = 4: the FP32 pipe becomes the limit;
//   CTA tile     16 queries x 128 rows, 8 warps (a warp = 4 query groups x 2 row groups x 4 lane pairs
//                covers 16 rows); k is streamed in 64-float chunks (256 B row segments, 36 KB stages,
//                4-deep cp.async ring);
//   work items   as before, plus the next item's metadata is fetched while the current one computes.

constexpr int S4_THREADS = 256;
constexpr int S4_RB = 128;            // rows per block
constexpr int S4_KC = 64;             // floats per k-chunk
constexpr int S4_RS = S4_KC + 8;      // padded stage row stride (floats): 72 = 8 (mod 32) banks
constexpr int S4_NSTAGE = 4;
constexpr int S4_DSTRIDE = 129;

// rows [row0, row0+128) x k-chunk -> one stage; 128 rows x 16 float4 = 2048 copies, 8 per thread
__device__ __forceinline__ void s4_load_stage(float* sdst, const float* __restrict__ gbase, int64_t row0, int64_t rmax,
                                              int d, bool tail, int r0) {
    if (!tail) {
        const float* g = gbase + (row0 + r0) * d;
#pragma unroll
        for (int j = 0; j < 8; ++j) cp_async16(sdst + (16 * j) * S4_RS, g + (16 * j) * d);
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int64_t r = row0 + r0 + 16 * j;
            if (r > rmax) r = rmax;
            cp_async16(sdst + (16 * j) * S4_RS, gbase + r * d);
        }
    }
}

template <int TQ>
__device__ __forceinline__ void s4_run_item(const float* __restrict__ q, const int32_t* __restrict__ qorder, int64_t p0,
                                            int cnt, const float* __restrict__ lbase, int64_t lb, int64_t len, int d,
                                            float* sQ, float* sR, float* sD, float (&td)[2], int (&tp)[2]) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int lp = lane & 3, qg = (lane >> 2) & 3, rg = lane >> 4;
    const int QS = d + 8;
    const int nkc = d / S4_KC;
    const int nblk = (int)((len + S4_RB - 1) / S4_RB);
    const int total = nblk * nkc;
    const int64_t rmax = len - 1;

    {   // queries of this item -> sQ (slots beyond cnt replicate the last query; never reported)
        const int nvec = d >> 2;
        for (int r = w; r < 4 * TQ; r += 8) {
            const int rr = r < cnt ? r : cnt - 1;
            const float* src = q + (int64_t)qorder[p0 + rr] * d;
            for (int c = lane; c < nvec; c += 32) cp_async16(sQ + r * QS + c * 4, src + c * 4);
        }
    }
    // this thread's copy slots: rows (tid >> 4) + 16 j of a block, float4 column (tid & 15)
    const int cr0 = tid >> 4;
    const float* gcol = lbase + (tid & 15) * 4;
    float* scol = sR + cr0 * S4_RS + (tid & 15) * 4;
    int ld_kc = 0, ld_rb = 0;
#pragma unroll
    for (int s = 0; s < S4_NSTAGE - 1; ++s) {
        if (s < total) {
            s4_load_stage(scol + s * S4_RB * S4_RS, gcol + ld_kc * S4_KC, (int64_t)ld_rb * S4_RB, rmax, d, ld_rb == nblk - 1, cr0);
            if (++ld_kc == nkc) { ld_kc = 0; ++ld_rb; }
        }
        cp_async_commit();
    }

    float2 acc[TQ][8];
#pragma unroll
    for (int a = 0; a < TQ; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = make_float2(0.f, 0.f);

    // thread (w, rg, qg, lp): queries qg + 4a, rows 16 w + 2 b + rg (b < 8), accumulator lanes 2lp, 2lp+1
    const float* myQ = sQ + qg * QS + 2 * lp;
    const int myRowOff = (16 * w + rg) * S4_RS + 2 * lp;
    int kc = 0, rb = 0, slot = 0, ld_slot = S4_NSTAGE - 1;
    for (int it = 0; it < total; ++it) {
        cp_async_wait<S4_NSTAGE - 2>();
        __syncthreads();
        if (it + S4_NSTAGE - 1 < total) {
            s4_load_stage(scol + ld_slot * S4_RB * S4_RS, gcol + ld_kc * S4_KC, (int64_t)ld_rb * S4_RB, rmax, d,
                          ld_rb == nblk - 1, cr0);
            if (++ld_kc == nkc) { ld_kc = 0; ++ld_rb; }
        }
        cp_async_commit();
        if (++ld_slot == S4_NSTAGE) ld_slot = 0;
        {
            const float2* sq = reinterpret_cast<const float2*>(myQ + kc * S4_KC);
            const float2* sr = reinterpret_cast<const float2*>(sR + slot * S4_RB * S4_RS + myRowOff);
#pragma unroll 2
            for (int s = 0; s < S4_KC / 8; ++s) {
                float2 qv[TQ], rv[8];
#pragma unroll
                for (int a = 0; a < TQ; ++a) qv[a] = sq[(4 * a) * (QS / 2) + 4 * s];
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const float2 r = sr[(2 * b) * (S4_RS / 2) + 4 * s];
                    rv[b] = make_float2(-r.x, -r.y);
                }
#pragma unroll
                for (int a = 0; a < TQ; ++a)
#pragma unroll
                    for (int b = 0; b < 8; ++b) {
                        const float2 t = __fadd2_rn(qv[a], rv[b]);          // x - y, both lanes, round-to-nearest
                        acc[a][b] = __ffma2_rn(t, t, acc[a][b]);            // fused, both lanes
                    }
            }
        }
        if (++slot == S4_NSTAGE) slot = 0;
        if (++kc == nkc) {
            kc = 0;
#pragma unroll
            for (int a = 0; a < TQ; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    float2 v = acc[a][b];
                    float2 o;
                    o.x = __shfl_xor_sync(0xffffffffu, v.x, 2);
                    o.y = __shfl_xor_sync(0xffffffffu, v.y, 2);
                    v = __fadd2_rn(v, o);
                    float sdist = __fadd_rn(v.x, v.y);
                    sdist = __fadd_rn(sdist, __shfl_xor_sync(0xffffffffu, sdist, 1));
                    if (((a * 8 + b) & 3) == lp) sD[(qg + 4 * a) * S4_DSTRIDE + 16 * w + 2 * b + rg] = sdist;
                    acc[a][b] = make_float2(0.f, 0.f);
                }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int qq = w + 8 * u;
                if (qq < cnt) {  // warp-uniform
#pragma unroll
                    for (int hh = 0; hh < 4; ++hh) {
                        const int64_t rowi = (int64_t)rb * S4_RB + hh * 32 + lane;
                        top8_update(td[u], tp[u], sD[qq * S4_DSTRIDE + hh * 32 + lane], (int)(lb + rowi), rowi < len, lane);
                    }
                }
            }
            ++rb;
        }
    }
    cp_async_wait<0>();
}

struct S4Item {
    int idx;        // ticket (>= nitems: stop)
    int cnt;
    long long p0, lb, len;
};

__device__ __forceinline__ void s4_fetch_item(S4Item* out, int32_t* ticket, int item_begin, int nitems,
                                              const int32_t* __restrict__ items, const uint32_t* __restrict__ keys,
                                              const int64_t* __restrict__ offsets, uint32_t nlist, int64_t n) {
    const int t = item_begin + atomicAdd(ticket, 1);
    S4Item it;
    it.idx = t;
    it.cnt = 0;
    it.p0 = 0;
    it.lb = 0;
    it.len = 0;
    if (t < nitems) {
        it.p0 = items[t];
        const uint32_t key = keys[it.p0];
        int c = 0;
#pragma unroll
        for (int i = 0; i < SC_QT; ++i) c += (it.p0 + i < n && keys[it.p0 + i] == key) ? 1 : 0;
        it.cnt = c;
        if (key < nlist) {
            it.lb = offsets[key];
            it.len = offsets[key + 1] - it.lb;
        }
    }
    *out = it;
}

__global__ void __launch_bounds__(S4_THREADS, 1)
scan_topk_kernel_v4(const float* __restrict__ q, const int32_t* __restrict__ qorder, const uint32_t* __restrict__ keys,
                    const int32_t* __restrict__ items, const int32_t* __restrict__ nitems_ptr, int32_t* __restrict__ ticket,
                    const float* __restrict__ vecs, const int64_t* __restrict__ ids, const int64_t* __restrict__ offsets,
                    int64_t nstored, uint32_t nlist, int64_t n, int d, float* __restrict__ D, int32_t* __restrict__ P,
                    int64_t* __restrict__ I, int item_begin, int item_end) {
    extern __shared__ __align__(16) float smem[];
    const int QS = d + 8;
    float* sQ = smem;                              // 16 x QS
    float* sR = sQ + SC_QT * QS;                   // S4_NSTAGE x 128 x 72
    float* sD = sR + S4_NSTAGE * S4_RB * S4_RS;    // 16 x 129
    __shared__ S4Item s_items[2];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int nitems = min(*nitems_ptr, item_end);   // this launch covers work items [item_begin, item_end)

    if (tid == 0) s4_fetch_item(&s_items[0], ticket, item_begin, nitems, items, keys, offsets, nlist, n);
    int cur = 0;
    for (;;) {
        __syncthreads();  // s_items[cur] visible; previous item fully consumed (sQ, sR, sD)
        const S4Item item = s_items[cur];
        if (item.idx >= nitems) break;
        // fetch the next item's metadata while this one computes (one thread, latency hidden)
        if (tid == 0) s4_fetch_item(&s_items[cur ^ 1], ticket, item_begin, nitems, items, keys, offsets, nlist, n);
        cur ^= 1;

        const int64_t p0 = item.p0, lb = item.lb, len = item.len;
        const int cnt = item.cnt;
        float td[2] = {RVCR_FLT_MAX, RVCR_FLT_MAX};
        int tp[2] = {0x7fffffff, 0x7fffffff};
        if (len > 0) {
            const float* lbase = vecs + lb * d;
            if (cnt <= 4) s4_run_item<1>(q, qorder, p0, cnt, lbase, lb, len, d, sQ, sR, sD, td, tp);
            else if (cnt <= 8) s4_run_item<2>(q, qorder, p0, cnt, lbase, lb, len, d, sQ, sR, sD, td, tp);
            else if (cnt <= 12) s4_run_item<3>(q, qorder, p0, cnt, lbase, lb, len, d, sQ, sR, sD, td, tp);
            else s4_run_item<4>(q, qorder, p0, cnt, lbase, lb, len, d, sQ, sR, sD, td, tp);
        }
        if (lane < RVCR_K) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int qq = w + 8 * u;
                if (qq < cnt) {
                    const int64_t qi = qorder[p0 + qq];
                    const bool pad = tp[u] == 0x7fffffff;
                    D[qi * RVCR_K + lane] = pad ? RVCR_FLT_MAX : td[u];
                    P[qi * RVCR_K + lane] = pad ? -1 : tp[u];
                    if (I) I[qi * RVCR_K + lane] = pad ? (int64_t)-1 : ids[tp[u]];
                }
            }
        }
    }
}

// Queries that formed no item (foreign lists on a shard) still need defined outputs.
__global__ void fill_pad_kernel(float* __restrict__ D, int32_t* __restrict__ P, int64_t* __restrict__ I, int64_t nk) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nk) return;
    D[i] = RVCR_FLT_MAX;
    P[i] = -1;
    if (I) I[i] = -1;
}

// ----------------------------------------------------------------------------- host side ----

int fill_pad(float* D, int32_t* P, int64_t* I, int64_t nk, cudaStream_t s) {
    if (nk <= 0) return RVCR_OK;
    fill_pad_kernel<<<(unsigned)((nk + 255) / 256), 256, 0, s>>>(D, P, I, nk);
    RVCR_CHECK_LAUNCH();
    return RVCR_OK;
}

static int ceil_log2(uint64_t v) {
    int b = 0;
    while ((1ull << b) < v) ++b;
    return b;
}

// Sort queries by probed list and build the work-item table.
//   ws_keys2 : sorted keys (uint32, local list id; nlist = foreign/invalid)
//   ws_vals2 : qorder (sorted position -> query index)
//   ws_items : item head positions (int32), ws_misc[0] = nitems, ws_misc[1] = ticket
int group_queries(rvcr_index* h, const int32_t* assign, int64_t n, bool only_local, cudaStream_t s) {
    if (n >= (1ll << 31)) {
        set_error("query batch too large (%lld)", (long long)n);
        return RVCR_ERR_INVALID;
    }
    RVCR_TRY(h->ws_keys.reserve(n * 4));
    RVCR_TRY(h->ws_vals.reserve(n * 4));
    RVCR_TRY(h->ws_keys2.reserve(n * 4));
    RVCR_TRY(h->ws_vals2.reserve(n * 4));
    RVCR_TRY(h->ws_items.reserve(n * 4));
    RVCR_TRY(h->ws_flags.reserve(n));
    RVCR_TRY(h->ws_misc.reserve(64));
    const int tb = 256;
    const unsigned nb = (unsigned)((n + tb - 1) / tb);
    make_keys_kernel<<<nb, tb, 0, s>>>(assign, n, h->own_lo, h->own_hi, h->nlist, h->ws_keys.as<uint32_t>(),
                                       h->ws_vals.as<int32_t>());
    RVCR_CHECK_LAUNCH();
    const int end_bit = ceil_log2((uint64_t)h->nlist + 2);
    size_t tmp = 0;
    RVCR_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, h->ws_keys.as<uint32_t>(), h->ws_keys2.as<uint32_t>(),
                                              h->ws_vals.as<int32_t>(), h->ws_vals2.as<int32_t>(), (int)n, 0, end_bit, s));
    size_t tmp2 = 0;
    cub::CountingInputIterator<int32_t> cnt_it(0);
    RVCR_CUDA(cub::DeviceSelect::Flagged(nullptr, tmp2, cnt_it, h->ws_flags.as<uint8_t>(), h->ws_items.as<int32_t>(),
                                         h->ws_misc.as<int32_t>(), (int)n, s));
    RVCR_TRY(h->ws_cub.reserve(tmp > tmp2 ? tmp : tmp2));
    RVCR_CUDA(cub::DeviceRadixSort::SortPairs(h->ws_cub.p, tmp, h->ws_keys.as<uint32_t>(), h->ws_keys2.as<uint32_t>(),
                                              h->ws_vals.as<int32_t>(), h->ws_vals2.as<int32_t>(), (int)n, 0, end_bit, s));
    flag_items_kernel<<<nb, tb, 0, s>>>(h->ws_keys2.as<uint32_t>(), n, (uint32_t)h->nlist, only_local ? 1 : 0,
                                        h->ws_flags.as<uint8_t>());
    RVCR_CHECK_LAUNCH();
    reset_counters_kernel<<<1, 32, 0, s>>>(h->ws_misc.as<int32_t>(), 8);
    RVCR_CHECK_LAUNCH();
    RVCR_CUDA(cub::DeviceSelect::Flagged(h->ws_cub.p, tmp2, cnt_it, h->ws_flags.as<uint8_t>(), h->ws_items.as<int32_t>(),
                                         h->ws_misc.as<int32_t>(), (int)n, s));
    count_launch(h, 6);
    return RVCR_OK;
}

static size_t scan_smem_bytes(int d, int impl) {
    if (impl == 1)
        return sizeof(float) * ((size_t)SC_QT * (d + 8) + (size_t)SC_NSTAGE * PD_RB * PD_KC + (size_t)SC_QT * SC_DSTRIDE);
    if (impl == 4)
        return sizeof(float) * ((size_t)SC_QT * (d + 8) + (size_t)S4_NSTAGE * S4_RB * S4_RS + (size_t)SC_QT * S4_DSTRIDE);
    return sizeof(float) * ((size_t)SC_QT * (d + 8) + (size_t)S2_NSTAGE * S2_RB * S2_RS + (size_t)SC_QT * S2_DSTRIDE);
}

int scan_topk_range(rvcr_index* h, const float* q, int64_t n, float* D, int32_t* P, int64_t* I, int item_begin, int item_end,
                    int ticket_slot, bool fill, cudaStream_t s);

// Requires group_queries() to have run on the same stream for the same batch.
int scan_topk(rvcr_index* h, const float* q, int64_t n, float* D, int32_t* P, int64_t* I, cudaStream_t s) {
    return scan_topk_range(h, q, n, D, P, I, 0, 0x7fffffff, 1, true, s);
}

// Work items [item_begin, item_end) only (items are ordered by list id, so a range is a set of lists and
// a contiguous range of sorted query positions).  ticket_slot selects the ws_misc counter used as the
// dynamic-scheduling ticket (slot 1 is zeroed by group_queries; slots 32.. by the caller).
int scan_topk_range(rvcr_index* h, const float* q, int64_t n, float* D, int32_t* P, int64_t* I, int item_begin, int item_end,
                    int ticket_slot, bool fill, cudaStream_t s) {
    // scan_impl 0 (auto): the tensor-core scan pays off once a call carries enough queries per list (Options)
    const bool want_tc = h->opt.scan_impl == 5 || (h->opt.scan_impl == 0 && n >= h->opt.scan_tc_min_queries);
    if (want_tc && scan_tc_supported(h))
        return scan_tc_range(h, q, n, D, P, I, item_begin, item_end, ticket_slot, fill, s);
    if (h->d % PD_KC != 0) {
        set_error("d=%d unsupported (must be a multiple of %d)", h->d, PD_KC);
        return RVCR_ERR_UNSUPPORTED;
    }
    if (h->nstored >= (1ll << 31)) {
        set_error("more than 2^31 stored vectors per shard is unsupported");
        return RVCR_ERR_UNSUPPORTED;
    }
    const int impl = (h->opt.scan_impl >= 1 && h->opt.scan_impl <= 4) ? (int)h->opt.scan_impl : 4;
    const size_t smem = scan_smem_bytes(h->d, impl);
    if (smem > 226 * 1024) {
        set_error("d=%d needs %zu B of shared memory", h->d, smem);
        return RVCR_ERR_UNSUPPORTED;
    }
    static SmemOptIn opt_in[5];   // per kernel version
    if (impl == 1)
        RVCR_TRY(opt_in[1].ensure(scan_topk_kernel_v1, smem, h->device));
    else if (impl == 2)
        RVCR_TRY(opt_in[2].ensure(scan_topk_kernel_v2, smem, h->device));
    else if (impl == 3)
        RVCR_TRY(opt_in[3].ensure(scan_topk_kernel_v3, smem, h->device));
    else
        RVCR_TRY(opt_in[4].ensure(scan_topk_kernel_v4, smem, h->device));
    if (fill) {
        const int64_t nk = n * RVCR_K;
        fill_pad_kernel<<<(unsigned)((nk + 255) / 256), 256, 0, s>>>(D, P, I, nk);
        RVCR_CHECK_LAUNCH();
    }
    if ((item_begin != 0 || item_end != 0x7fffffff) && h->opt.scan_impl != 4 && h->opt.scan_impl != 5 && h->opt.scan_impl != 0) {
        set_error("work-item ranges need scan_impl 4");
        return RVCR_ERR_UNSUPPORTED;
    }
    int grid = h->num_sms;
    if (n < grid) grid = (int)(n > 0 ? n : 1);
    if (impl == 1)
        scan_topk_kernel_v1<<<grid, PD_THREADS, smem, s>>>(q, h->ws_vals2.as<int32_t>(), h->ws_keys2.as<uint32_t>(),
                                                            h->ws_items.as<int32_t>(), h->ws_misc.as<int32_t>(),
                                                            h->ws_misc.as<int32_t>() + 1, h->vecs, h->ids, h->offsets,
                                                            h->nstored, (uint32_t)h->nlist, n, h->d, D, P, I);
    else if (impl == 2)
        scan_topk_kernel_v2<<<grid, PD_THREADS, smem, s>>>(q, h->ws_vals2.as<int32_t>(), h->ws_keys2.as<uint32_t>(),
                                                            h->ws_items.as<int32_t>(), h->ws_misc.as<int32_t>(),
                                                            h->ws_misc.as<int32_t>() + 1, h->vecs, h->ids, h->offsets,
                                                            h->nstored, (uint32_t)h->nlist, n, h->d, D, P, I);
    else if (impl == 3)
        scan_topk_kernel_v3<<<grid, S3_THREADS, smem, s>>>(q, h->ws_vals2.as<int32_t>(), h->ws_keys2.as<uint32_t>(),
                                                            h->ws_items.as<int32_t>(), h->ws_misc.as<int32_t>(),
                                                            h->ws_misc.as<int32_t>() + 1, h->vecs, h->ids, h->offsets,
                                                            h->nstored, (uint32_t)h->nlist, n, h->d, D, P, I);
    else
        scan_topk_kernel_v4<<<grid, S4_THREADS, smem, s>>>(q, h->ws_vals2.as<int32_t>(), h->ws_keys2.as<uint32_t>(),
                                                            h->ws_items.as<int32_t>(), h->ws_misc.as<int32_t>(),
                                                            h->ws_misc.as<int32_t>() + ticket_slot, h->vecs, h->ids, h->offsets,
                                                            h->nstored, (uint32_t)h->nlist, n, h->d, D, P, I, item_begin, item_end);
    RVCR_CHECK_LAUNCH();
    count_launch(h, 2);
    return RVCR_OK;
}

}  // namespace rvcr
