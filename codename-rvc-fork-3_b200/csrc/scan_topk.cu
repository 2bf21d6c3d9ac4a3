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

// ---- packed fp32 pairs (FADD2 / FFMA2 on the Blackwell FP32 pipe) ----

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
__device__ __forceinline__ void s4_run_item(const float* __restrict__ q, const PushArgs& push, const int32_t* __restrict__ qorder, int64_t p0,
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
            const float* src = query_row(q, push, qorder[p0 + rr], d);
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
                    int64_t* __restrict__ I, int item_begin, int item_end, PushArgs push) {
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
            if (cnt <= 4) s4_run_item<1>(q, push, qorder, p0, cnt, lbase, lb, len, d, sQ, sR, sD, td, tp);
            else if (cnt <= 8) s4_run_item<2>(q, push, qorder, p0, cnt, lbase, lb, len, d, sQ, sR, sD, td, tp);
            else if (cnt <= 12) s4_run_item<3>(q, push, qorder, p0, cnt, lbase, lb, len, d, sQ, sR, sD, td, tp);
            else s4_run_item<4>(q, push, qorder, p0, cnt, lbase, lb, len, d, sQ, sR, sD, td, tp);
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

static size_t scan_smem_bytes(int d) {
    return sizeof(float) * ((size_t)SC_QT * (d + 8) + (size_t)S4_NSTAGE * S4_RB * S4_RS + (size_t)SC_QT * S4_DSTRIDE);
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
//
// scan_impl: 0 = auto (tensor-core scan from Options::scan_tc_min_queries, else the exact FFMA scan), 5 = tensor-core scan,
// 1..4 = the exact FFMA scan (the round-1 kernel generations v1-v3 lost to v4 on every shape and were removed in round 2;
// their option values stay accepted and select v4).
int scan_topk_range(rvcr_index* h, const float* q, int64_t n, float* D, int32_t* P, int64_t* I, int item_begin, int item_end,
                    int ticket_slot, bool fill, cudaStream_t s) {
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
    const size_t smem = scan_smem_bytes(h->d);
    if (smem > 226 * 1024) {
        set_error("d=%d needs %zu B of shared memory", h->d, smem);
        return RVCR_ERR_UNSUPPORTED;
    }
    static SmemOptIn opt_in;
    RVCR_TRY(opt_in.ensure(scan_topk_kernel_v4, smem, h->device));
    if (fill) {
        const int64_t nk = n * RVCR_K;
        fill_pad_kernel<<<(unsigned)((nk + 255) / 256), 256, 0, s>>>(D, P, I, nk);
        RVCR_CHECK_LAUNCH();
    }
    int grid = h->num_sms;
    if (n < grid) grid = (int)(n > 0 ? n : 1);
    scan_topk_kernel_v4<<<grid, S4_THREADS, smem, s>>>(q, h->ws_vals2.as<int32_t>(), h->ws_keys2.as<uint32_t>(),
                                                        h->ws_items.as<int32_t>(), h->ws_misc.as<int32_t>(),
                                                        h->ws_misc.as<int32_t>() + ticket_slot, h->vecs, h->ids, h->offsets,
                                                        h->nstored, (uint32_t)h->nlist, n, h->d, D, P, I, item_begin, item_end, h->fuse.push);
    RVCR_CHECK_LAUNCH();
    count_launch(h, 2);
    return RVCR_OK;
}

}  // namespace rvcr
