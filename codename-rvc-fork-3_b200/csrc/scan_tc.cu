// scan_tc.cu -- inverted-list scan on the 5th-gen tensor cores (tcgen05 kind::tf32 + TMEM + TMA), made
// FP32-exact by an exact re-check.  scan_impl = 5 (default).
//
// Replaces the same reference stage as scan_topk.cu: index.search(npy, k=8) (rvc/infer/pipeline.py:499)
// -> IndexIVF::search_preassigned -> IVFFlatScanner<METRIC_L2>::scan_codes + heap.
//
// Why: the exact FP32 scan needs 2 lane-ops per (query, row, element); with ~11 queries per probed list at
// C5 that is as much FP32-pipe time as the list bytes take at HBM speed, and in practice
// the shared-memory pipe saturates first.  Here the streamed list rows never touch the CUDA cores:
//
//   TMA      list rows (fp32, row-major, 128 rows x 32 floats = 16 KB boxes, 128B swizzle) HBM -> shared memory
//   tcgen05  kind::tf32 MMA  acc[row, query] = <y_row, q>  (A = the staged rows, B = the item's <= 16 queries,
//            both read from shared memory as K-major fp32; the tensor core uses the top 19 bits of each
//            operand), fp32 accumulation in TMEM
//   epilogue v = ||y||^2 - 2 acc  (||y||^2 precomputed per stored row), per query a sorted list of the 32
//            smallest v per TMEM lane quarter, kept in registers across the warp
//   merge    per query: v8 = 8th smallest v overall; every row with v <= v8 + margin(q) is a candidate
//   recheck  canonical fp32 distance (fvec_L2sqr order, pairdist.cuh) of each candidate, top-8 by
//            (distance, position): bit-identical to the exact scan
//
// margin(q) bounds twice the worst-case |v - v_true| (tf32 truncation of both operands, Cauchy-Schwarz;
// fp32 accumulation of d products; norm roundings), so the true top-8 are always among the candidates
// (DESIGN.md "Numerics").  If a quarter's 32-entry list might have been truncated inside the margin, or more
// than 64 rows qualify, the query is re-done by the brute-force exact kernel below.
#include "tc_common.cuh"

namespace rvcr {

namespace stc {

constexpr int TILE_M = 128;                 // list rows per accumulator tile (UMMA M)
constexpr int QT = 16;                      // queries per work item (UMMA N); == SC_QT of scan_topk.cu
constexpr int KB = 32;                      // fp32 per k-block: 128 B = one swizzle-atom row
constexpr int UMMA_K = 8;                   // tf32: 32 B per MMA
constexpr int STAGE_BYTES = TILE_M * 128;   // 16 KB
constexpr int QKB_BYTES = QT * 128;         // 2 KB: one k-block of the query operand
constexpr int NACC = 4;                     // TMEM accumulator stages (QT columns each)
constexpr int TMEM_COLS = 64;
constexpr int IRING = 4;                    // work-item descriptor ring
constexpr int THREADS = 256;                // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 item scheduler, 4-7 epilogue
constexpr int EPI_WARP0 = 4;
constexpr int EPI_THREADS = 128;
constexpr int TOPT = 32;                    // approximate candidates kept per (query, lane quarter)
constexpr int CCAP = 64;                    // candidates handed to the re-check per query
constexpr int MERGE_BYTES = 4 * QT * TOPT * 8;

struct Item {
    int idx;   // ticket; >= nitems: stop
    int cnt;   // queries in the item (1..16)
    long long p0, lb, len;
};

struct Smem {
    uint64_t full[8], empty[8];
    uint64_t tfull[NACC], tempty[NACC];
    uint64_t ifull[IRING], iempty[IRING];
    uint64_t qfull[2];
    Item items[IRING];
    uint32_t tmem_ptr;
};

__device__ __forceinline__ void fetch_item(Item* out, int32_t* ticket, int item_begin, int nitems,
                                           const int32_t* __restrict__ items, const uint32_t* __restrict__ keys,
                                           const int64_t* __restrict__ offsets, uint32_t nlist, int64_t n) {
    const int t = item_begin + atomicAdd(ticket, 1);
    Item it;
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
        for (int i = 0; i < QT; ++i) c += (it.p0 + i < n && keys[it.p0 + i] == key) ? 1 : 0;
        it.cnt = c;
        if (key < nlist) {
            it.lb = offsets[key];
            it.len = offsets[key + 1] - it.lb;
        }
    }
    *out = it;
}

// insert the warp's 32 values (v, pos) into the sorted (ascending across lanes) list (td, tp)
__device__ __forceinline__ void top32_insert(float& td, int& tp, float v, int pos, int lane) {
    const float thr = __shfl_sync(0xffffffffu, td, 31);
    unsigned m = __ballot_sync(0xffffffffu, v < thr);
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const float x = __shfl_sync(0xffffffffu, v, src);
        const int xp = __shfl_sync(0xffffffffu, pos, src);
        const int rank = __popc(__ballot_sync(0xffffffffu, td <= x));
        const float ud = __shfl_up_sync(0xffffffffu, td, 1);
        const int up = __shfl_up_sync(0xffffffffu, tp, 1);
        if (lane == rank) {
            td = x;
            tp = xp;
        } else if (lane > rank) {
            td = ud;
            tp = up;
        }
    }
}

// bitonic sort of one value per lane, ascending across lanes (first fill of an empty list)
__device__ __forceinline__ void warp_sort32(float& v, int& p, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, v, j);
            const int op = __shfl_xor_sync(0xffffffffu, p, j);
            const bool up = ((lane & k) == 0);          // ascending block
            const bool lower = ((lane & j) == 0);       // this lane keeps the smaller of the pair when ascending
            const bool take_min = (up == lower);
            const bool other_smaller = (ov < v) || (ov == v && op < p);
            if (take_min ? other_smaller : !other_smaller) {
                v = ov;
                p = op;
            }
        }
    }
}

// queries of an item -> K-major, 128B-swizzled tf32 operand tile [kb][16 rows][32 floats]; rows >= cnt
// replicate the last query (finite values; their accumulator columns are never read)
__device__ __forceinline__ void load_queries(uint8_t* sQ, const float* __restrict__ q, const int32_t* __restrict__ qorder,
                                             const Item& it, int d, int nkb, int et) {
    const int per_row = nkb * 8;                 // 16-byte chunks per query row
    const int total = QT * per_row;
    for (int c = et; c < total; c += EPI_THREADS) {
        const int j = c / per_row, rem = c - j * per_row;
        const int kb = rem >> 3, c16 = rem & 7;
        const int jj = j < it.cnt ? j : it.cnt - 1;
        const float* src = q + (int64_t)qorder[it.p0 + jj] * d + kb * KB + c16 * 4;
        uint8_t* dst = sQ + kb * QKB_BYTES + (j >> 3) * 1024 + (j & 7) * 128 + ((c16 ^ (j & 7)) << 4);
        cp_async16(dst, src);
    }
}

__global__ void __launch_bounds__(THREADS, 1)
scan_tc_kernel(const __grid_constant__ CUtensorMap tm32, const __grid_constant__ CUtensorMap tm64,
               const __grid_constant__ CUtensorMap tm96, const __grid_constant__ CUtensorMap tm128,
               const float* __restrict__ q, const int32_t* __restrict__ qorder, const uint32_t* __restrict__ keys,
               const int32_t* __restrict__ items, const int32_t* __restrict__ nitems_ptr, int32_t* __restrict__ ticket,
               const int64_t* __restrict__ offsets, const float* __restrict__ ynorm, uint32_t nlist, int64_t n, int d,
               int nstage, float ymax, float rel_margin, float abs_margin, int item_begin, int item_end,
               int32_t* __restrict__ cand, int32_t* __restrict__ candcnt, int32_t* __restrict__ fb_list,
               int32_t* __restrict__ fb_count, unsigned long long* __restrict__ totals /*[0] fallbacks, [1] candidates*/) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int nkb = d / KB;
    uint8_t* sR = smem;                                         // nstage x 16 KB
    uint8_t* sQ = sR + (size_t)nstage * STAGE_BYTES;            // 2 x nkb x 2 KB
    const int qbuf_bytes = nkb * QKB_BYTES;
    float* sMv = reinterpret_cast<float*>(sQ + 2 * (size_t)qbuf_bytes);   // [4][16][32]
    int* sMp = reinterpret_cast<int*>(sMv + 4 * QT * TOPT);               // [4][16][32]
    Smem* S = reinterpret_cast<Smem*>(reinterpret_cast<uint8_t*>(sMp) + 4 * QT * TOPT * 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nitems = min(*nitems_ptr, item_end);

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tm32);
        tc::prefetch_tmap(&tm64);
        tc::prefetch_tmap(&tm96);
        tc::prefetch_tmap(&tm128);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < nstage; ++i) {
            tc::mbar_init(&S->full[i], 1);
            tc::mbar_init(&S->empty[i], 1);
        }
        for (int i = 0; i < NACC; ++i) {
            tc::mbar_init(&S->tfull[i], 1);
            tc::mbar_init(&S->tempty[i], EPI_THREADS);
        }
        for (int i = 0; i < IRING; ++i) {
            tc::mbar_init(&S->ifull[i], 1);
            tc::mbar_init(&S->iempty[i], 2 + 4);     // TMA thread + MMA thread + one lane per epilogue warp
        }
        tc::mbar_init(&S->qfull[0], EPI_THREADS);
        tc::mbar_init(&S->qfull[1], EPI_THREADS);
        tc::fence_mbar_init();
    } else if (warp == 2) {
        tc::tmem_alloc<1>(&S->tmem_ptr, TMEM_COLS);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = S->tmem_ptr;

    if (warp == 3) {
        // ================================ work-item scheduler ===========================================
        // runs up to IRING items ahead of the consumers, so the ticket / metadata latency is never on the
        // TMA producer's critical path
        if (lane == 0) {
            for (uint32_t sn = 0;; ++sn) {
                Item nxt;
                fetch_item(&nxt, ticket, item_begin, nitems, items, keys, offsets, nlist, n);
                const uint32_t slot = sn % IRING, use = sn / IRING;
                tc::mbar_wait(&S->iempty[slot], (use & 1) ^ 1);
                S->items[slot] = nxt;
                tc::mbar_arrive_local(&S->ifull[slot]);
                if (nxt.idx >= nitems) break;
            }
        }
    } else if (warp == 0) {
        // ================================ TMA producer ==================================================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t it = 0;; ++it) {
                const uint32_t slot = it % IRING, use = it / IRING;
                tc::mbar_wait(&S->ifull[slot], use & 1);
                const Item cur = S->items[slot];
                tc::mbar_arrive_local(&S->iempty[slot]);
                if (cur.idx >= nitems) break;
                const int ntiles = (int)((cur.len + TILE_M - 1) / TILE_M);
                for (int t = 0; t < ntiles; ++t) {
                    const int64_t left = cur.len - (int64_t)t * TILE_M;
                    const int rows = left >= 97 ? 128 : left >= 65 ? 96 : left >= 33 ? 64 : 32;
                    const CUtensorMap* tm = rows == 128 ? &tm128 : rows == 96 ? &tm96 : rows == 64 ? &tm64 : &tm32;
                    const int row0 = (int)(cur.lb + (int64_t)t * TILE_M);
                    for (int kb = 0; kb < nkb; ++kb) {
                        tc::mbar_wait(&S->empty[stage], phase ^ 1);
                        tc::mbar_arrive_expect_tx(&S->full[stage], (uint32_t)rows * 128u);
                        tc::tma_load_2d<1>(tm, &S->full[stage], sR + (size_t)stage * STAGE_BYTES, kb * KB, row0);
                        if (++stage == (uint32_t)nstage) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ===================================================
        if (lane == 0) {
            constexpr uint32_t idesc = tc::make_idesc_tf32(TILE_M, QT);
            uint32_t stage = 0, phase = 0, tile_cnt = 0;
            for (uint32_t it = 0;; ++it) {
                const uint32_t slot = it % IRING, use = it / IRING;
                tc::mbar_wait(&S->ifull[slot], use & 1);
                const Item cur = S->items[slot];
                tc::mbar_arrive_local(&S->iempty[slot]);
                if (cur.idx >= nitems) break;
                tc::mbar_wait(&S->qfull[it & 1], (it >> 1) & 1);        // this item's queries are in sQ[it & 1]
                const uint32_t qaddr = tc::smem_u32(sQ + (size_t)(it & 1) * qbuf_bytes);
                const int ntiles = (int)((cur.len + TILE_M - 1) / TILE_M);
                for (int t = 0; t < ntiles; ++t, ++tile_cnt) {
                    const uint32_t as = tile_cnt % NACC, ause = tile_cnt / NACC;
                    tc::mbar_wait(&S->tempty[as], (ause & 1) ^ 1);      // epilogue drained this accumulator
                    tc::tc_fence_after();
                    const uint32_t tmem_d = tmem_base + as * QT;
                    for (int kb = 0; kb < nkb; ++kb) {
                        tc::mbar_wait(&S->full[stage], phase);
                        tc::tc_fence_after();
                        const uint64_t da = tc::make_smem_desc(tc::smem_u32(sR + (size_t)stage * STAGE_BYTES));
                        const uint64_t db = tc::make_smem_desc(qaddr + kb * QKB_BYTES);
#pragma unroll
                        for (int k = 0; k < KB / UMMA_K; ++k)   // +32 B inside the 128 B swizzle row = +2 (16-byte units)
                            tc::umma_tf32(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        tc::umma_commit<1>(&S->empty[stage]);
                        if (kb == nkb - 1) tc::umma_commit<1>(&S->tfull[as]);
                        if (++stage == (uint32_t)nstage) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ================================ epilogue =====================================================
        const int ew = warp - EPI_WARP0;          // == warp % 4: the TMEM lane quarter this warp may read
        const int et = threadIdx.x - EPI_WARP0 * 32;
        uint32_t tile_cnt = 0;
        unsigned long long n_fb = 0, n_cand = 0;    // lane 0: running totals for rvcr_get_stats
        tc::mbar_wait(&S->ifull[0], 0);
        Item cur = S->items[0];
        __syncwarp();
        if (lane == 0) tc::mbar_arrive_local(&S->iempty[0]);
        if (cur.idx < nitems) {
            load_queries(sQ, q, qorder, cur, d, nkb, et);
            cp_async_commit();
            cp_async_wait<0>();
            tc::fence_proxy_async_smem();
            tc::mbar_arrive_local(&S->qfull[0]);
        }
        for (uint32_t it = 0;; ++it) {
            if (cur.idx >= nitems) break;
            Item nxt;
            {
                const uint32_t sn = it + 1, slot = sn % IRING, use = sn / IRING;
                tc::mbar_wait(&S->ifull[slot], use & 1);
                nxt = S->items[slot];
                __syncwarp();
                if (lane == 0) tc::mbar_arrive_local(&S->iempty[slot]);
            }
            const bool have_next = nxt.idx < nitems;
            // buffer (it+1)&1 was last read by item it-1, whose MMAs and merge are complete
            if (have_next) load_queries(sQ + (size_t)((it + 1) & 1) * qbuf_bytes, q, qorder, nxt, d, nkb, et);
            cp_async_commit();
            bool next_published = !have_next;

            float td[QT];
            int tp[QT];
#pragma unroll
            for (int j = 0; j < QT; ++j) {
                td[j] = __int_as_float(0x7f800000);
                tp[j] = -1;
            }
            const int cnt = cur.cnt;
            const int ntiles = (int)((cur.len + TILE_M - 1) / TILE_M);
            for (int t = 0; t < ntiles; ++t, ++tile_cnt) {
                const uint32_t as = tile_cnt % NACC, ause = tile_cnt / NACC;
                const int64_t row = (int64_t)t * TILE_M + ew * 32 + lane;
                const bool valid = row < cur.len;
                const int pos = (int)(cur.lb + row);
                const float yn = valid ? __ldg(ynorm + pos) : 0.f;
                tc::mbar_wait(&S->tfull[as], ause & 1);
                tc::tc_fence_after();
                uint32_t r[16];
                tc::tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + as * QT, r);
                tc::tmem_ld_wait();
                tc::tc_fence_before();
                tc::mbar_arrive_local(&S->tempty[as]);
                const bool first = (t == 0);
#pragma unroll
                for (int j = 0; j < QT; ++j) {
                    if (j < cnt) {   // warp-uniform
                        float v = valid ? fmaf(-2.f, __uint_as_float(r[j]), yn) : __int_as_float(0x7f800000);
                        if (!(v == v)) v = __int_as_float(0x7f800000);   // NaN rows are never candidates
                        if (first) {
                            int p = v < __int_as_float(0x7f800000) ? pos : -1;
                            warp_sort32(v, p, lane);
                            td[j] = v;
                            tp[j] = p;
                        } else {
                            top32_insert(td[j], tp[j], v, pos, lane);
                        }
                    }
                }
                if (!next_published) {
                    cp_async_wait<0>();
                    tc::fence_proxy_async_smem();
                    tc::mbar_arrive_local(&S->qfull[(it + 1) & 1]);
                    next_published = true;
                }
            }
            if (!next_published) {
                cp_async_wait<0>();
                tc::fence_proxy_async_smem();
                tc::mbar_arrive_local(&S->qfull[(it + 1) & 1]);
            }

            // ---- merge the four quarter lists, select the candidates of every query of the item ----
            tc::named_bar_sync(1, EPI_THREADS);      // previous item's merge reads of sM are done
#pragma unroll
            for (int j = 0; j < QT; ++j) {
                if (j < cnt) {
                    sMv[(ew * QT + j) * TOPT + lane] = td[j];
                    sMp[(ew * QT + j) * TOPT + lane] = tp[j];
                }
            }
            tc::named_bar_sync(1, EPI_THREADS);
            const uint8_t* sQc = sQ + (size_t)(it & 1) * qbuf_bytes;
            for (int j = ew; j < cnt; j += 4) {
                const int64_t qi = qorder[cur.p0 + j];
                // 8th smallest v overall: it is among the 8 smallest of each quarter
                const float x = sMv[((lane >> 3) * QT + j) * TOPT + (lane & 7)];
                int rk = 0;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float xi = __shfl_sync(0xffffffffu, x, i);
                    rk += (xi < x || (xi == x && i < lane)) ? 1 : 0;
                }
                const unsigned b7 = __ballot_sync(0xffffffffu, rk == 7);
                const float v8 = __shfl_sync(0xffffffffu, x, __ffs(b7) - 1);
                // ||q||^2 from the staged query row (any order; rounding is covered by the margin slack)
                float qn2 = 0.f;
                for (int c = lane; c < nkb * 8; c += 32) {
                    const int kb = c >> 3, c16 = c & 7;
                    const float4 f = *reinterpret_cast<const float4*>(sQc + kb * QKB_BYTES + (j >> 3) * 1024 + (j & 7) * 128 +
                                                                      ((c16 ^ (j & 7)) << 4));
                    qn2 = fmaf(f.x, f.x, qn2);
                    qn2 = fmaf(f.y, f.y, qn2);
                    qn2 = fmaf(f.z, f.z, qn2);
                    qn2 = fmaf(f.w, f.w, qn2);
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) qn2 += __shfl_xor_sync(0xffffffffu, qn2, off);
                const float margin = rel_margin * sqrtf(qn2) * ymax + abs_margin * (qn2 + ymax * ymax);
                const float cutoff = v8 + margin;
                int total = 0;
                bool ovf = false;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const float v = sMv[(w * QT + j) * TOPT + lane];
                    const int p = sMp[(w * QT + j) * TOPT + lane];
                    const bool pass = p >= 0 && v <= cutoff;
                    const unsigned bm = __ballot_sync(0xffffffffu, pass);
                    if (bm >> 31) ovf = true;     // the quarter's 32nd entry qualifies: its list may be truncated
                    const int idx = total + __popc(bm & ((1u << lane) - 1u));
                    if (pass && idx < CCAP) cand[qi * CCAP + idx] = p;
                    total += __popc(bm);
                }
                if (lane == 0) {
                    if (ovf || total > CCAP) {
                        candcnt[qi] = -1;
                        fb_list[atomicAdd(fb_count, 1)] = (int32_t)(cur.p0 + j);   // sorted position
                        ++n_fb;
                    } else {
                        candcnt[qi] = total;
                        n_cand += (unsigned long long)total;
                    }
                }
            }
            cur = nxt;
        }
        if (lane == 0) {
            if (n_fb) atomicAdd(&totals[0], n_fb);
            if (n_cand) atomicAdd(&totals[1], n_cand);
        }
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc<1>(tmem_base, TMEM_COLS);
}

}  // namespace stc

// ------------------------------------------------------------------------- exact re-check ----

// One warp per query (walked in list-sorted order, so neighbouring warps touch the same list): canonical fp32
// distance of each candidate -- two lanes per candidate, lane h owns accumulators 4h..4h+3 (float4 loads),
// sequential over the row exactly like fvec_L2sqr -- then top-8 by (distance, position).
constexpr int RC_WARPS = 4;

__global__ void __launch_bounds__(RC_WARPS * 32)
scan_recheck_kernel(const float* __restrict__ q, const int32_t* __restrict__ qorder, const int32_t* __restrict__ items,
                    const int32_t* __restrict__ nitems_ptr, int item_begin, int item_end, int64_t n, int d,
                    const float* __restrict__ vecs, const int64_t* __restrict__ ids, const int32_t* __restrict__ cand,
                    const int32_t* __restrict__ candcnt, float* __restrict__ D, int32_t* __restrict__ P, int64_t* __restrict__ I) {
    extern __shared__ __align__(16) float rc_smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float* sq = rc_smem + (size_t)w * d;
    const int nit = *nitems_ptr;
    const int64_t first = item_begin < nit ? (int64_t)items[item_begin] : n;
    const int64_t last = item_end < nit ? (int64_t)items[item_end] : n;
    const int h = lane & 1;
    const int nvec = d >> 2;
    for (int64_t p = first + (int64_t)blockIdx.x * RC_WARPS + w; p < last; p += (int64_t)gridDim.x * RC_WARPS) {
        const int64_t qi = qorder[p];
        const int nc = candcnt[qi];
        if (nc < 0) continue;   // overflowed: the exact fallback kernel answers this query
        float td = RVCR_FLT_MAX;
        int tp = 0x7fffffff;
        if (nc > 0) {
            __syncwarp();
            const float4* q4 = reinterpret_cast<const float4*>(q + qi * d);
            for (int c = lane; c < nvec; c += 32) reinterpret_cast<float4*>(sq)[c] = q4[c];
            __syncwarp();
            for (int base = 0; base < nc; base += 16) {
                const int c = base + (lane >> 1);
                const bool have = c < nc;
                const int pos = cand[qi * stc::CCAP + (have ? c : 0)];
                const float4* y4 = reinterpret_cast<const float4*>(vecs + (int64_t)pos * d) + h;
                const float4* x4 = reinterpret_cast<const float4*>(sq) + h;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
                for (int i = 0; i < (d >> 3); ++i) {
                    const float4 y = __ldg(y4 + 2 * i);
                    const float4 x = x4[2 * i];
                    const float t0 = __fsub_rn(x.x, y.x), t1 = __fsub_rn(x.y, y.y), t2 = __fsub_rn(x.z, y.z), t3 = __fsub_rn(x.w, y.w);
                    a0 = __fmaf_rn(t0, t0, a0);
                    a1 = __fmaf_rn(t1, t1, a1);
                    a2 = __fmaf_rn(t2, t2, a2);
                    a3 = __fmaf_rn(t3, t3, a3);
                }
                // canonical reduction ((a0+a4)+(a1+a5))+((a2+a6)+(a3+a7)); lane h holds a[4h..4h+3]
                const float s0 = __fadd_rn(a0, __shfl_xor_sync(0xffffffffu, a0, 1));
                const float s1 = __fadd_rn(a1, __shfl_xor_sync(0xffffffffu, a1, 1));
                const float s2 = __fadd_rn(a2, __shfl_xor_sync(0xffffffffu, a2, 1));
                const float s3 = __fadd_rn(a3, __shfl_xor_sync(0xffffffffu, a3, 1));
                const float dist = __fadd_rn(__fadd_rn(s0, s1), __fadd_rn(s2, s3));
                top8_update(td, tp, dist, pos, have && h == 0, lane);
            }
        }
        if (lane < RVCR_K) {
            const bool pad = tp == 0x7fffffff;
            D[qi * RVCR_K + lane] = pad ? RVCR_FLT_MAX : td;
            P[qi * RVCR_K + lane] = pad ? -1 : tp;
            if (I) I[qi * RVCR_K + lane] = pad ? (int64_t)-1 : ids[tp];
        }
    }
}

// Brute-force exact scan of one list for the queries the tensor-core pass could not bound (rare): one CTA
// per query, 8 lanes per row (canonical order), 32 rows per step.
__global__ void __launch_bounds__(256)
scan_exact_fallback_kernel(const float* __restrict__ q, const int32_t* __restrict__ qorder, const uint32_t* __restrict__ keys,
                           const int32_t* __restrict__ fb_list, const int32_t* __restrict__ fb_count, uint32_t nlist, int d,
                           const float* __restrict__ vecs, const int64_t* __restrict__ ids, const int64_t* __restrict__ offsets,
                           float* __restrict__ D, int32_t* __restrict__ P, int64_t* __restrict__ I) {
    extern __shared__ __align__(16) float fb_smem[];
    float* sq = fb_smem;          // d
    float* sd = fb_smem + d;      // 32
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int l8 = tid & 7, grp = tid >> 3;   // 32 groups of 8 lanes
    const int nfb = *fb_count;
    for (int f = blockIdx.x; f < nfb; f += gridDim.x) {
        const int64_t p = fb_list[f];
        const int64_t qi = qorder[p];
        const uint32_t key = keys[p];
        int64_t lb = 0, len = 0;
        if (key < nlist) {
            lb = offsets[key];
            len = offsets[key + 1] - lb;
        }
        __syncthreads();
        for (int c = tid; c < d; c += 256) sq[c] = q[qi * d + c];
        __syncthreads();
        float td = RVCR_FLT_MAX;
        int tp = 0x7fffffff;
        for (int64_t r0 = 0; r0 < len; r0 += 32) {
            const int64_t r = r0 + grp;
            float acc = 0.f;
            if (r < len) {
                const float* y = vecs + (lb + r) * d;
                for (int i = 0; i < d; i += 8) {
                    const float t = __fsub_rn(sq[i + l8], y[i + l8]);
                    acc = __fmaf_rn(t, t, acc);
                }
            }
            acc = reduce8_canonical(acc);
            if (l8 == 0) sd[grp] = acc;
            __syncthreads();
            if (w == 0) {
                const int64_t rr = r0 + lane;
                top8_update(td, tp, sd[lane], (int)(lb + rr), rr < len, lane);
            }
            __syncthreads();
        }
        if (w == 0 && lane < RVCR_K) {
            const bool pad = tp == 0x7fffffff;
            D[qi * RVCR_K + lane] = pad ? RVCR_FLT_MAX : td;
            P[qi * RVCR_K + lane] = pad ? -1 : tp;
            if (I) I[qi * RVCR_K + lane] = pad ? (int64_t)-1 : ids[tp];
        }
    }
}

// ---------------------------------------------------------------------------- row norms ------

// one warp per stored row: ||y||^2 (fp32) and the running maximum (as uint bits; non-negative floats order as uints)
__global__ void row_norm_kernel(const float* __restrict__ vecs, int64_t nrows, int d, float* __restrict__ norm2,
                                unsigned int* __restrict__ max_bits) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= nrows) return;
    const float4* y4 = reinterpret_cast<const float4*>(vecs + r * d);
    float s = 0.f;
    for (int c = lane; c < (d >> 2); c += 32) {
        const float4 v = __ldg(y4 + c);
        s = fmaf(v.x, v.x, s);
        s = fmaf(v.y, v.y, s);
        s = fmaf(v.z, v.z, s);
        s = fmaf(v.w, v.w, s);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) {
        norm2[r] = s;
        if (s > 0.f && s < __int_as_float(0x7f800000)) atomicMax(max_bits, __float_as_uint(s));
    }
}

// ------------------------------------------------------------------------------ host side ----

int scan_tc_invalidate(rvcr_index* h) {
    h->row_norm_valid = false;
    h->scan_tmap_valid = false;
    return RVCR_OK;
}

static int ensure_row_norms(rvcr_index* h, cudaStream_t s) {
    if (h->row_norm_valid) return RVCR_OK;
    RVCR_TRY(h->row_norm.reserve((size_t)(h->nstored > 0 ? h->nstored : 1) * 4));
    unsigned int* mx = h->ws_misc.as<unsigned int>() + 13;
    RVCR_CUDA(cudaMemsetAsync(mx, 0, 4, s));
    if (h->nstored > 0) {
        const int64_t thr = h->nstored * 32;
        row_norm_kernel<<<(unsigned)((thr + 255) / 256), 256, 0, s>>>(h->vecs, h->nstored, h->d, h->row_norm.as<float>(), mx);
        RVCR_CHECK_LAUNCH();
        count_launch(h);
    }
    unsigned int bits = 0;
    RVCR_CUDA(cudaMemcpyAsync(&bits, mx, 4, cudaMemcpyDeviceToHost, s));
    RVCR_CUDA(cudaStreamSynchronize(s));
    float m2;
    memcpy(&m2, &bits, 4);
    h->row_norm_max = sqrtf(m2);
    h->row_norm_valid = true;
    return RVCR_OK;
}

static int ensure_scan_tmaps(rvcr_index* h) {
    if (h->scan_tmap_valid && h->scan_tmap_base == (const void*)h->vecs && h->scan_tmap_rows == h->nstored) return RVCR_OK;
    tc::EncodeTiledFn fn = tc::get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return RVCR_ERR_CUDA;
    }
    static_assert(sizeof(CUtensorMap) == 128, "tensor map size");
    for (int i = 0; i < 4; ++i) {
        cuuint64_t dims[2] = {(cuuint64_t)h->d, (cuuint64_t)h->nstored};
        cuuint64_t strides[1] = {(cuuint64_t)h->d * 4};
        cuuint32_t box[2] = {(cuuint32_t)stc::KB, (cuuint32_t)(32 * (i + 1))};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = fn(reinterpret_cast<CUtensorMap*>(h->scan_tmaps[i]), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, h->vecs, dims, strides,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled failed (%d) for the %lld x %d list storage", (int)r, (long long)h->nstored, h->d);
            return RVCR_ERR_CUDA;
        }
    }
    h->scan_tmap_valid = true;
    h->scan_tmap_base = h->vecs;
    h->scan_tmap_rows = h->nstored;
    return RVCR_OK;
}

bool scan_tc_supported(const rvcr_index* h) { return h->d % stc::KB == 0 && h->d >= stc::KB && h->d <= 1024 && h->nstored > 0; }

// Same contract as scan_topk_range (scan_topk.cu); requires group_queries() on the same stream.
int scan_tc_range(rvcr_index* h, const float* q, int64_t n, float* D, int32_t* P, int64_t* I, int item_begin, int item_end,
                  int ticket_slot, bool fill, cudaStream_t s) {
    if (h->nstored >= (1ll << 31)) {
        set_error("more than 2^31 stored vectors per shard is unsupported");
        return RVCR_ERR_UNSUPPORTED;
    }
    RVCR_TRY(ensure_row_norms(h, s));
    RVCR_TRY(ensure_scan_tmaps(h));
    const int d = h->d, nkb = d / stc::KB;
    const size_t fixed = 2 * (size_t)nkb * stc::QKB_BYTES + stc::MERGE_BYTES + sizeof(stc::Smem) + 1024;
    int nstage = (int)((226 * 1024 - fixed) / stc::STAGE_BYTES);
    if (nstage > 8) nstage = 8;
    if (nstage < 2) {
        set_error("d=%d leaves no room for the row ring in shared memory", d);
        return RVCR_ERR_UNSUPPORTED;
    }
    const size_t smem = fixed + (size_t)nstage * stc::STAGE_BYTES;
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        RVCR_CUDA(cudaFuncSetAttribute(stc::scan_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    RVCR_TRY(h->ws_cand.reserve((size_t)n * stc::CCAP * 4));
    RVCR_TRY(h->ws_candcnt.reserve((size_t)n * 4));
    RVCR_TRY(h->ws_fb.reserve((size_t)n * 4 + 64));
    int32_t* fb_count = h->ws_misc.as<int32_t>() + 14;
    RVCR_CUDA(cudaMemsetAsync(fb_count, 0, 4, s));
    if (fill) {
        RVCR_TRY(fill_pad(D, P, I, n * RVCR_K, s));
    }
    // margin between two approximate values: 2 x |v - v_true|, |v - v_true| <= 2 |acc - <q,y>| + norm roundings;
    // |acc - <q,y>| <= (2^-9 (tf32 truncation of both operands) + d 2^-21 (fp32 accumulation)) |q||y|
    const float rel = 4.f * (1.01f / 512.f + (float)d / 2097152.f);
    const float absm = 1.f / 524288.f;
    int grid = h->num_sms;
    const CUtensorMap* tm = reinterpret_cast<const CUtensorMap*>(h->scan_tmaps);
    stc::scan_tc_kernel<<<grid, stc::THREADS, smem, s>>>(
        tm[0], tm[1], tm[2], tm[3], q, h->ws_vals2.as<int32_t>(), h->ws_keys2.as<uint32_t>(), h->ws_items.as<int32_t>(),
        h->ws_misc.as<int32_t>(), h->ws_misc.as<int32_t>() + ticket_slot, h->offsets, h->row_norm.as<float>(), (uint32_t)h->nlist, n, d,
        nstage, h->row_norm_max, rel, absm, item_begin, item_end, h->ws_cand.as<int32_t>(), h->ws_candcnt.as<int32_t>(),
        h->ws_fb.as<int32_t>(), fb_count, reinterpret_cast<unsigned long long*>(h->ws_misc.as<int32_t>() + 48));
    RVCR_CHECK_LAUNCH();
    {
        const size_t rsm = (size_t)RC_WARPS * d * 4;
        int64_t rgrid = (int64_t)h->num_sms * 8;
        if (rgrid * RC_WARPS > n) rgrid = (n + RC_WARPS - 1) / RC_WARPS;
        scan_recheck_kernel<<<(unsigned)rgrid, RC_WARPS * 32, rsm, s>>>(q, h->ws_vals2.as<int32_t>(), h->ws_items.as<int32_t>(),
                                                                        h->ws_misc.as<int32_t>(), item_begin, item_end, n, d, h->vecs,
                                                                        h->ids, h->ws_cand.as<int32_t>(), h->ws_candcnt.as<int32_t>(), D,
                                                                        P, I);
        RVCR_CHECK_LAUNCH();
    }
    {
        const size_t fsm = (size_t)(d + 32) * 4;
        scan_exact_fallback_kernel<<<h->num_sms * 2, 256, fsm, s>>>(q, h->ws_vals2.as<int32_t>(), h->ws_keys2.as<uint32_t>(),
                                                                    h->ws_fb.as<int32_t>(), fb_count, (uint32_t)h->nlist, d, h->vecs,
                                                                    h->ids, h->offsets, D, P, I);
        RVCR_CHECK_LAUNCH();
    }
    // running total of fallback queries (read back lazily by rvcr_get_stats)
    h->scan_fb_dirty = true;
    count_launch(h, fill ? 4 : 3);
    return RVCR_OK;
}

}  // namespace rvcr
