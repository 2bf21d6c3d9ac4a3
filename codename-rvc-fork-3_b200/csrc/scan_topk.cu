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
scan_topk_kernel(const float* __restrict__ q, const int32_t* __restrict__ qorder, const uint32_t* __restrict__ keys,
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

// Queries that formed no item (foreign lists on a shard) still need defined outputs.
__global__ void fill_pad_kernel(float* __restrict__ D, int32_t* __restrict__ P, int64_t* __restrict__ I, int64_t nk) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nk) return;
    D[i] = RVCR_FLT_MAX;
    P[i] = -1;
    if (I) I[i] = -1;
}

// ----------------------------------------------------------------------------- host side ----

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
    return sizeof(float) * ((size_t)SC_QT * (d + 8) + (size_t)SC_NSTAGE * PD_RB * PD_KC + (size_t)SC_QT * SC_DSTRIDE);
}

// Requires group_queries() to have run on the same stream for the same batch.
int scan_topk(rvcr_index* h, const float* q, int64_t n, float* D, int32_t* P, int64_t* I, cudaStream_t s) {
    if (h->d % PD_KC != 0) {
        set_error("d=%d unsupported (must be a multiple of %d)", h->d, PD_KC);
        return RVCR_ERR_UNSUPPORTED;
    }
    if (h->nstored >= (1ll << 31)) {
        set_error("more than 2^31 stored vectors per shard is unsupported");
        return RVCR_ERR_UNSUPPORTED;
    }
    const size_t smem = scan_smem_bytes(h->d);
    static size_t attr_smem = 0;  // largest dynamic shared memory size opted into so far
    if (smem > attr_smem) {
        RVCR_CUDA(cudaFuncSetAttribute(scan_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    if (smem > 226 * 1024) {
        set_error("d=%d needs %zu B of shared memory", h->d, smem);
        return RVCR_ERR_UNSUPPORTED;
    }
    {
        const int64_t nk = n * RVCR_K;
        fill_pad_kernel<<<(unsigned)((nk + 255) / 256), 256, 0, s>>>(D, P, I, nk);
        RVCR_CHECK_LAUNCH();
    }
    int grid = h->num_sms;
    if (n < grid) grid = (int)(n > 0 ? n : 1);
    scan_topk_kernel<<<grid, PD_THREADS, smem, s>>>(q, h->ws_vals2.as<int32_t>(), h->ws_keys2.as<uint32_t>(),
                                                     h->ws_items.as<int32_t>(), h->ws_misc.as<int32_t>(),
                                                     h->ws_misc.as<int32_t>() + 1, h->vecs, h->ids, h->offsets, h->nstored,
                                                     (uint32_t)h->nlist, n, h->d, D, P, I);
    RVCR_CHECK_LAUNCH();
    count_launch(h, 2);
    return RVCR_OK;
}

}  // namespace rvcr
