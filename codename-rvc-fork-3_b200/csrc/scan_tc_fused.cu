// scan_tc_fused.cu -- scan_impl = 6 (EXPERIMENTAL, never chosen by the auto mode): the tensor-core list scan of scan_tc.cu
// with the exact re-check + gather / weights / blend folded into the same persistent kernel as a fourth warp role.
//
// Same reference stage as scan_tc.cu: index.search(npy, k=8) + the numpy tail, rvc/infer/pipeline.py:499-506.
//
// Why (profiles/r01_ncu_summary.md): with two kernels the re-check runs after ALL lists have been
// streamed, so its 13 candidate rows per query (and the 8 gathered ones) come from HBM a second time: 64.6 GB of DRAM
// traffic for 30.8 GB of algorithmic work at 0.38 of the HBM roofline (a latency-bound kernel), and the scan phase as a
// whole stops at 0.62.  Here a list's candidates are re-checked by warps of the same CTA right after the list has been
// streamed (a 4-deep hand-off ring behind the epilogue warps), while the TMA / MMA / splitter warps are already on the next
// lists.  What that can buy: (1) the latency-bound re-check overlaps the bandwidth-bound stream instead of following it;
// (2) the 8 gather reads follow the distance pass of the same warp within microseconds and should hit L2; (3) SOME of the
// candidate reads hit L2 -- the chip streams ~6 MB per microsecond through a 126 MB L2, so only rows younger than ~15 us
// survive, i.e. the tail of the list that has just been streamed.  Honest expectation: phase 32 -> 26-28 ms, not the 22 ms
// of the roofline; the second read of the candidates is inherent as long as a list (1 MB) does not fit on chip.
//
// Status: written after the last GPU session of round 1 and NOT yet run on a B200; results must equal scan_impl 5 bit for
// bit (same candidates, same canonical distances, same tail arithmetic).  tests/test_gpu_scan_tc.py covers it only when
// RVCR_TEST_SCAN_IMPL=6.  The 768-thread CTA needs per-role register budgets (setmaxnreg): 65536 / 768 = 85 registers
// per thread would spill the splitter and epilogue warps (113 in the 512-thread kernel).
#include <stddef.h>
#include <stdlib.h>

#include "scan_tc.cuh"

namespace rvcr {

namespace stc {

constexpr int RCW = 8;                          // re-check warps (warpgroups 4 and 5)
constexpr int RC_WARP0 = THREADS / 32;          // 16
constexpr int FTHREADS = THREADS + RCW * 32;    // 768
constexpr int RCRING = 4;                       // items the epilogue may run ahead of the re-check warps
// Register budgets per warpgroup (setmaxnreg; multiples of 8).  The CTA is launched with 80 per thread (61 440); the
// budgets add up to 128 x 32 + 128 x 96 + 256 x 104 + 256 x 56 = 57 344, i.e. 4 096 registers of slack so that an
// increase never has to wait for the last register of the pool.  ptxas reports the same 28-56 bytes of spill (all in the
// common prologue, outside the loops) as with the tightest fit 40 / 104 / 112 / 56.
constexpr int REGS_CTRL = 32;                   // warps 0-3: TMA, MMA, TMEM allocator, scheduler
constexpr int REGS_EPI = 96;                    // warps 4-7
constexpr int REGS_SPLIT = 104;                 // warps 8-15
constexpr int REGS_RC = 56;                     // warps 16-23

struct SmemF : Smem {
    uint64_t rcfull[RCRING], rcempty[RCRING];   // epilogue -> re-check hand-off ring
    int rc_p0[RCRING];                          // sorted position of the item's first query
    int rc_cnt[RCRING];                         // queries in the item; -1 = no more items
};

template <int N>
__device__ __forceinline__ void reg_inc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_dec() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(N));
}

// One query (sorted position p) by one warp: the body of scan_recheck_kernel (scan_tc.cu) without the shared-memory copy
// of the query -- all 16 lane pairs read the SAME query addresses, so those loads are one broadcast sector per warp
// instruction from L1 / L2, and this kernel's shared memory is full.  cand / candcnt were written by the epilogue warps of
// this CTA and published through the hand-off barrier: plain (coherent) loads, not the read-only path.
template <int BLEND>
__device__ __forceinline__ void recheck_query(int64_t p, const float* q, const int32_t* __restrict__ qorder, int d,
                                              const float* __restrict__ vecs, const int64_t* __restrict__ ids,
                                              const int32_t* cand, const int32_t* candcnt, float* D, int32_t* P, int64_t* I,
                                              const BlendArgs& ba, int lane) {
    const int64_t qi = qorder[p];
    const int nc = *reinterpret_cast<const volatile int32_t*>(candcnt + qi);
    if (nc < 0) return;   // overflowed: the exact fallback kernel answers this query
    const int h = lane & 1;
    const int nvec = d >> 2;
    const float4* q4 = reinterpret_cast<const float4*>(q + qi * d);
    float td = RVCR_FLT_MAX;
    int tp = 0x7fffffff;
    for (int base = 0; base < nc; base += 16) {
        const int c = base + (lane >> 1);
        const bool have = c < nc;
        const int pos = *reinterpret_cast<const volatile int32_t*>(cand + qi * CCAP + (have ? c : 0));
        const float4* y4 = reinterpret_cast<const float4*>(vecs + (int64_t)pos * d) + h;
        const float4* x4 = q4 + h;
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
    if (lane < RVCR_K) {
        const bool pad = tp == 0x7fffffff;
        D[qi * RVCR_K + lane] = pad ? RVCR_FLT_MAX : td;
        P[qi * RVCR_K + lane] = pad ? -1 : tp;
        if (I) I[qi * RVCR_K + lane] = pad ? (int64_t)-1 : ids[tp];
    }
    if (BLEND) {
        float wn;
        int spos;
        blend_weights(td, tp, lane, ba.last_row_pos, wn, spos);
        float wk[RVCR_K];
        const float4* rows[RVCR_K];
#pragma unroll
        for (int k = 0; k < RVCR_K; ++k) {
            wk[k] = __shfl_sync(0xffffffffu, wn, k);
            const int pk = __shfl_sync(0xffffffffu, spos, k);
            rows[k] = pk >= 0 ? reinterpret_cast<const float4*>(vecs + (int64_t)pk * d) : nullptr;
        }
        float4* o4 = BLEND == 3 ? push_row_ptr(ba.push, qi, d) : reinterpret_cast<float4*>(ba.out + qi * d);
        for (int c = lane; c < nvec; c += 32) {
            float4 v[RVCR_K];
#pragma unroll
            for (int k = 0; k < RVCR_K; ++k) v[k] = rows[k] ? __ldg(rows[k] + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 f0 = q4[c];      // read before this lane overwrites the same 16 bytes (in-place blend)
            const float4 o = blend_vec(v, wk, f0, ba.r32, ba.omr32);
            if (BLEND == 2) post_store(ba.post, qi, d, c, o, f0);
            else o4[c] = o;
        }
    }
}

template <int BLEND>
__global__ void __launch_bounds__(FTHREADS, 1)
scan_tc_fused_kernel(const __grid_constant__ CUtensorMap tm32, const __grid_constant__ CUtensorMap tm64,
               const __grid_constant__ CUtensorMap tm96, const __grid_constant__ CUtensorMap tm128,
               const float* q /* may alias ba.out: blended in place */, const int32_t* __restrict__ qorder, const uint32_t* __restrict__ keys,
               const int32_t* __restrict__ items, const int32_t* __restrict__ nitems_ptr, int32_t* __restrict__ ticket,
               const int64_t* __restrict__ offsets, const float* __restrict__ ynorm, uint32_t nlist, int64_t n, int d,
               int nstage, float ymax, float acc_rel, float abs_margin, int item_begin, int item_end,
               int32_t* cand, int32_t* candcnt /* written by the epilogue warps, read by the re-check warps */, int32_t* __restrict__ fb_list,
               int32_t* __restrict__ fb_count, unsigned long long* __restrict__ totals /*[0] fallbacks, [1] candidates*/,
               float2* __restrict__ stage_buf /*[grid][16][SCAP]*/, int dbg_flags, const float* __restrict__ vecs,
               const int64_t* __restrict__ ids, float* D, int32_t* P, int64_t* I, BlendArgs ba) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int nkb = d / KB;
    uint8_t* sR = smem;                                         // nstage x 16 KB
    uint8_t* sQ = sR + (size_t)nstage * STAGE_BYTES;            // 2 x nkb x 2 KB
    const int qbuf_bytes = nkb * QKB_BYTES;
    float* sV = reinterpret_cast<float*>(sQ + 2 * (size_t)qbuf_bytes);    // [16][CHUNK_ROWS]
    SmemF* S = reinterpret_cast<SmemF*>(reinterpret_cast<uint8_t*>(sV) + SV_BYTES);

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
            tc::mbar_init(&S->empty[i], CONV_THREADS / 32);
        }
        for (int i = 0; i < NTS; ++i) {
            tc::mbar_init(&S->conv[i], CONV_THREADS / 32);
            tc::mbar_init(&S->aempty[i], 1);
        }
        for (int i = 0; i < NACC; ++i) {
            tc::mbar_init(&S->tfull[i], 1);
            tc::mbar_init(&S->tempty[i], EPI_THREADS / 32);
        }
        for (int i = 0; i < IRING; ++i) {
            tc::mbar_init(&S->ifull[i], 1);
            tc::mbar_init(&S->iempty[i], 2 + 4 + 4 * CONV_GROUPS); // TMA thread + MMA thread + one lane per epilogue / splitter warp
        }
        tc::mbar_init(&S->qfull[0], EPI_THREADS / 32);
        tc::mbar_init(&S->qfull[1], EPI_THREADS / 32);
        for (int i = 0; i < RCRING; ++i) {
            tc::mbar_init(&S->rcfull[i], EPI_THREADS / 32);   // every epilogue warp has published its queries' candidates
            tc::mbar_init(&S->rcempty[i], RCW);               // every re-check warp is done with the slot
        }
        tc::fence_mbar_init();
    } else if (warp == 2) {
        tc::tmem_alloc<1>(&S->tmem_ptr, TMEM_COLS);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = S->tmem_ptr;

    const int wg = warp >> 2;
    if (wg == 0) {
    reg_dec<REGS_CTRL>();
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
        // converged warp, uniformized coordinates (see the MMA issuer), one elected lane issues the copies
        const uint32_t sR_addr = tc::warp_uniform(tc::smem_u32(sR));
        const uint32_t bar_addr = tc::warp_uniform(tc::smem_u32(S));
        uint32_t stage = 0, phase = 0;
        for (uint32_t it = 0;; ++it) {
            const uint32_t slot = it % IRING, use = it / IRING;
            mbar_wait_warp(&S->ifull[slot], use & 1, lane);
            const Item cur = S->items[slot];
            mbar_arrive_warp(&S->iempty[slot], lane);
            if (!tc::warp_uniform(cur.idx < nitems ? 1 : 0)) break;
            const int len = tc::warp_uniform((int)cur.len);       // < 2^31 rows per shard
            const int lb = tc::warp_uniform((int)cur.lb);
            const int ntiles = (len + TILE_M - 1) / TILE_M;
            for (int t = 0; t < ntiles; ++t) {
                const int left = len - t * TILE_M;
                const int rows = left >= 97 ? 128 : left >= 65 ? 96 : left >= 33 ? 64 : 32;
                const CUtensorMap* tm = rows == 128 ? &tm128 : rows == 96 ? &tm96 : rows == 64 ? &tm64 : &tm32;
                const int row0 = lb + t * TILE_M;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait_warp(&S->empty[stage], phase ^ 1, lane);
                    if ((dbg_flags & 8) && (kb & 1) == 0 && kb + 1 < nkb) {
                        // tuning variant (RVCR_STC_FLAGS=8): issue k-blocks in pairs -- hold this box until the NEXT stage
                        // is free too, so that the two 128-byte segments of every row (256 contiguous bytes in HBM) are
                        // requested back to back instead of a stage time apart (DRAM page locality of the 3 KB rows)
                        const uint32_t s2 = stage + 1 == (uint32_t)nstage ? 0u : stage + 1;
                        const uint32_t p2 = s2 == 0u ? phase ^ 1u : phase;
                        mbar_wait_warp(&S->empty[s2], p2 ^ 1, lane);
                    }
                    if (tc::elect_one()) {
                        const uint32_t fb = bar_addr + (uint32_t)offsetof(Smem, full) + stage * 8;
                        tc::mbar_arrive_expect_tx_addr(fb, (uint32_t)rows * 128u);
                        tc::tma_load_2d_addr(tm, fb, sR_addr + stage * STAGE_BYTES, kb * KB, row0);
                    }
                    __syncwarp();
                    if (++stage == (uint32_t)nstage) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ===================================================
        // The whole warp runs this loop converged; loop bounds and addresses are REDUX-uniformized so they live in
        // uniform registers, and one elected lane issues the tcgen05 instructions.
        constexpr uint32_t idesc = tc::make_idesc_tf32(TILE_M, QT);
        const uint32_t tbase = tc::warp_uniform(tmem_base);
        const uint32_t sQ_addr = tc::warp_uniform(tc::smem_u32(sQ));
        const uint32_t bar_addr = tc::warp_uniform(tc::smem_u32(S));
        uint32_t ts = 0, tphase = 0, tile_cnt = 0;
        for (uint32_t it = 0;; ++it) {
            const uint32_t slot = it % IRING, use = it / IRING;
            mbar_wait_warp(&S->ifull[slot], use & 1, lane);
            const Item cur = S->items[slot];
            mbar_arrive_warp(&S->iempty[slot], lane);
            if (!tc::warp_uniform(cur.idx < nitems ? 1 : 0)) break;
            const int ntiles = tc::warp_uniform((int)((cur.len + TILE_M - 1) / TILE_M));
            mbar_wait_warp(&S->qfull[it & 1], (it >> 1) & 1, lane);        // this item's queries are in sQ[it & 1]
            const uint32_t qaddr = sQ_addr + (it & 1) * (uint32_t)qbuf_bytes;
            for (int t = 0; t < ntiles; ++t, ++tile_cnt) {
                const uint32_t as = tile_cnt % NACC, ause = tile_cnt / NACC;
                mbar_wait_warp(&S->tempty[as], (ause & 1) ^ 1, lane);      // epilogue drained this accumulator
                tc::tc_fence_after();
                const uint32_t tmem_d = tbase + as * QT;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait_warp(&S->conv[ts], tphase, lane);            // the split row block is in TMEM stage ts
                    tc::tc_fence_after();
                    const uint64_t db = tc::make_smem_desc(qaddr + kb * QKB_BYTES);
                    const uint32_t ta = tbase + A_COL0 + ts * A_STAGE_COLS;
                    if (tc::elect_one()) {
#pragma unroll
                        for (int k = 0; k < KB / UMMA_K; ++k) {   // B: +32 B inside the 128 B swizzle row = +2 (16-byte units)
                            tc::umma_tf32_ts(tmem_d, ta + UMMA_K * k, db + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
                            tc::umma_tf32_ts(tmem_d, ta + KB + UMMA_K * k, db + (uint64_t)(2 * k), idesc, 1u);
                        }
                        tc::umma_commit_addr(bar_addr + (uint32_t)offsetof(Smem, aempty) + ts * 8);
                        if (kb == nkb - 1) tc::umma_commit_addr(bar_addr + (uint32_t)offsetof(Smem, tfull) + as * 8);
                    }
                    __syncwarp();
                    if (++ts == NTS) { ts = 0; tphase ^= 1; }
                }
            }
        }
    }
    } else if (wg == 2 || wg == 3) {
        reg_inc<REGS_SPLIT>();
        // ================================ row splitters ================================================
        // thread t owns row t of every staged [128 rows][32 floats] block: y = hi + lo with hi = y's top 19 bits
        // (exactly representable in tf32, so the result does not depend on how the tensor core would cut a raw fp32
        // operand) and lo = y - hi (exact).  Both go to TMEM as A operands (columns [0,32) hi, [32,64) lo of the
        // stage); the shared-memory stage is handed back to the TMA producer as soon as it has been read.
        const int cw = (warp - CONV_WARP0) & 3;      // == warp % 4: TMEM lane quarter
        const int cg = (warp - CONV_WARP0) >> 2;     // splitter group: handles stage sequence numbers = cg (mod CONV_GROUPS)
        const int t = cw * 32 + lane;
        uint32_t stage = 0, phase = 0, ts = 0, tphase = 0, seq = 0;
        for (uint32_t it = 0;; ++it) {
            const uint32_t slot = it % IRING, use = it / IRING;
            mbar_wait_warp(&S->ifull[slot], use & 1, lane);
            const Item cur = S->items[slot];
            mbar_arrive_warp(&S->iempty[slot], lane);
            if (cur.idx >= nitems) break;
            const int ntiles = (int)((cur.len + TILE_M - 1) / TILE_M);
            for (int tl = 0; tl < ntiles; ++tl) {
                const int64_t left = cur.len - (int64_t)tl * TILE_M;
                const bool active = (int64_t)cw * 32 < left;     // warp-uniform: this quarter holds list rows
                for (int kb = 0; kb < nkb; ++kb, ++seq) {
                    if ((int)(seq % CONV_GROUPS) == cg) {
                        mbar_wait_warp(&S->full[stage], phase, lane);
                        mbar_wait_warp(&S->aempty[ts], tphase ^ 1, lane);
                        tc::tc_fence_after();
                        const uint32_t taddr = tmem_base + ((uint32_t)(cw * 32) << 16) + A_COL0 + ts * A_STAGE_COLS;
                        if (active && !(dbg_flags & 4)) {
                            const uint8_t* row = sR + (size_t)stage * STAGE_BYTES + t * 128;
#pragma unroll
                            for (int hf = 0; hf < 2; ++hf) {
                                uint32_t hi[16], lo[16];
#pragma unroll
                                for (int c = 0; c < 4; ++c) {
                                    const float4 f = *reinterpret_cast<const float4*>(row + (((4 * hf + c) ^ (t & 7)) << 4));
                                    const float fv[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        const uint32_t hb = __float_as_uint(fv[e]) & 0xffffe000u;
                                        hi[4 * c + e] = hb;
                                        lo[4 * c + e] = __float_as_uint(fv[e] - __uint_as_float(hb));
                                    }
                                }
                                if (hf == 1) mbar_arrive_warp(&S->empty[stage], lane);   // the row block is in registers
                                tc::tmem_st16(taddr + 16 * hf, hi);
                                tc::tmem_st16(taddr + KB + 16 * hf, lo);
                            }
                            tc::tmem_st_wait();
                        } else {
                            mbar_arrive_warp(&S->empty[stage], lane);
                        }
                        tc::tc_fence_before();
                        mbar_arrive_warp(&S->conv[ts], lane);
                    }
                    if (++stage == (uint32_t)nstage) { stage = 0; phase ^= 1; }
                    if (++ts == NTS) { ts = 0; tphase ^= 1; }
                }
            }
        }
    } else if (wg == 1) {
        reg_inc<REGS_EPI>();
        // ================================ epilogue =====================================================
        const int ew = warp - EPI_WARP0;          // == warp % 4: the TMEM lane quarter this warp may read
        const int et = threadIdx.x - EPI_WARP0 * 32;
        uint32_t tile_cnt = 0;
        unsigned long long n_fb = 0, n_cand = 0;    // lane 0: running totals for rvcr_get_stats
        uint32_t nproc = 0;                         // items this CTA has handed to the re-check warps
        mbar_wait_warp(&S->ifull[0], 0, lane);
        Item cur = S->items[0];
        mbar_arrive_warp(&S->iempty[0], lane);
        if (cur.idx < nitems) {
            load_queries(sQ, q, qorder, cur, d, nkb, et);
            cp_async_commit();
            cp_async_wait<0>();
            convert_queries(sQ, S->qnorm[0], nkb, et);
            tc::fence_proxy_async_smem();
            mbar_arrive_warp(&S->qfull[0], lane);
        }
        for (uint32_t it = 0;; ++it) {
            if (cur.idx >= nitems) break;
            Item nxt;
            {
                const uint32_t sn = it + 1, slot = sn % IRING, use = sn / IRING;
                mbar_wait_warp(&S->ifull[slot], use & 1, lane);
                nxt = S->items[slot];
                mbar_arrive_warp(&S->iempty[slot], lane);
            }
            const bool have_next = nxt.idx < nitems;
            // buffer (it+1)&1 was last read by item it-1, whose MMAs and selection are complete
            if (have_next) load_queries(sQ + (size_t)((it + 1) & 1) * qbuf_bytes, q, qorder, nxt, d, nkb, et);
            cp_async_commit();
            bool next_published = !have_next;

            const int cnt = cur.cnt;
            const int ntiles = (int)((cur.len + TILE_M - 1) / TILE_M);
            // this warp selects for queries ew, ew+4, ew+8, ew+12 of the item (slot u <-> query ew + 4u)
            float ub[4], mg[4], lmin[4];
            int ncand[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                ub[u] = __int_as_float(0x7f800000);
                lmin[u] = __int_as_float(0x7f800000);   // this lane's minimum over rows = lane (mod 32), all chunks so far
                mg[u] = 0.f;
                ncand[u] = 0;
            }
            // lists longer than one chunk: candidates are staged with their v (the cutoff only tightens from chunk
            // to chunk) and filtered once more against the final cutoff
            const bool multi = ntiles > CHUNK_TILES;
            float2* stage = stage_buf + ((size_t)blockIdx.x * QT) * SCAP;
            for (int t0 = 0; t0 < ntiles || t0 == 0; t0 += CHUNK_TILES) {
                const int t1 = min(t0 + CHUNK_TILES, ntiles);
                // ---- pass 1: v = ||y||^2 - 2 <y, q> of this chunk's tiles -> sV[query][row in chunk] ----
                for (int t = t0; t < t1; ++t, ++tile_cnt) {
                    const uint32_t as = tile_cnt % NACC, ause = tile_cnt / NACC;
                    const int64_t row = (int64_t)t * TILE_M + ew * 32 + lane;
                    const bool valid = row < cur.len;
                    const float yn = valid ? __ldg(ynorm + cur.lb + row) : 0.f;
                    mbar_wait_warp(&S->tfull[as], ause & 1, lane);
                    tc::tc_fence_after();
                    uint32_t r[16];
                    tc::tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + as * QT, r);
                    tc::tmem_ld_wait();
                    tc::tc_fence_before();
                    mbar_arrive_warp(&S->tempty[as], lane);
                    float* dst = sV + (t - t0) * TILE_M + ew * 32 + lane;
#pragma unroll
                    for (int j = 0; j < QT; ++j)
                        if (j < cnt)   // warp-uniform
                            dst[j * CHUNK_ROWS] = valid ? fmaf(-2.f, __uint_as_float(r[j]), yn) : __int_as_float(0x7f800000);
                    if (!next_published) {
                        cp_async_wait<0>();
                        convert_queries(sQ + (size_t)((it + 1) & 1) * qbuf_bytes, S->qnorm[(it + 1) & 1], nkb, et);
                        tc::fence_proxy_async_smem();
                        mbar_arrive_warp(&S->qfull[(it + 1) & 1], lane);
                        next_published = true;
                    }
                }
                if (!next_published) {
                    cp_async_wait<0>();
                    convert_queries(sQ + (size_t)((it + 1) & 1) * qbuf_bytes, S->qnorm[(it + 1) & 1], nkb, et);
                    tc::fence_proxy_async_smem();
                    mbar_arrive_warp(&S->qfull[(it + 1) & 1], lane);
                    next_published = true;
                }
                tc::named_bar_sync(1, EPI_THREADS);      // the chunk's v are in sV
                // ---- pass 2: per query, every row with v <= U + margin(q) is a candidate, where U >= the 8th
                // smallest v of the list: the 8th smallest of the 32 lane minima (8 distinct rows are <= it) ----
                const int rows = (int)min((long long)CHUNK_ROWS, cur.len - (long long)t0 * TILE_M);
                const int pos0 = (int)(cur.lb + (int64_t)t0 * TILE_M);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = ew + 4 * u;
                    if (j < cnt && rows > 0) {   // warp-uniform
                        const int64_t qi = qorder[cur.p0 + j];
                        const float* vj = sV + j * CHUNK_ROWS;
                        if (t0 == 0) {
                            // |v - v_true| <= 2 |acc - <q, y>| + roundings, |acc - <q, y>| <= ||q_lo|| ||y|| (the part of q
                            // the tf32 operand drops) + acc_rel ||q|| ||y|| (lo-part truncation, fp32 accumulation);
                            // two approximate values are compared, hence the factor 4 (DESIGN.md "Numerics")
                            const float2 nn = S->qnorm[it & 1][j];
                            mg[u] = 4.04f * ymax * (sqrtf(nn.y) + acc_rel * sqrtf(nn.x)) + abs_margin * (nn.x + ymax * ymax);
                        }
                        float lm = lmin[u];
                        for (int r0 = lane; r0 < rows; r0 += 32) lm = fminf(lm, vj[r0]);
                        lmin[u] = lm;
                        int rk = 0;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float xi = __shfl_sync(0xffffffffu, lm, i);
                            rk += (xi < lm || (xi == lm && i < lane)) ? 1 : 0;
                        }
                        const unsigned b7 = __ballot_sync(0xffffffffu, rk == 7);
                        ub[u] = __shfl_sync(0xffffffffu, lm, __ffs(b7) - 1);   // non-increasing from chunk to chunk
                        const float cutoff = ub[u] + mg[u];
                        for (int r0 = 0; r0 < rows; r0 += 32) {
                            const int rr = r0 + lane;
                            const float v = rr < rows ? vj[rr] : __int_as_float(0x7f800000);
                            const bool pass = rr < rows && v <= cutoff;
                            const unsigned bm = __ballot_sync(0xffffffffu, pass);
                            const int idx = ncand[u] + __popc(bm & ((1u << lane) - 1u));
                            if (pass) {
                                if (!multi) {
                                    if (idx < CCAP) cand[qi * CCAP + idx] = pos0 + rr;
                                } else if (idx < SCAP) {
                                    stage[(size_t)j * SCAP + idx] = make_float2(v, __int_as_float(pos0 + rr));
                                }
                            }
                            ncand[u] += __popc(bm);
                        }
                    }
                }
                tc::named_bar_sync(1, EPI_THREADS);      // selection done: sV may be overwritten
            }
            if (multi) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = ew + 4 * u;
                    if (j < cnt) {   // warp-uniform
                        const int64_t qi = qorder[cur.p0 + j];
                        const int staged = ncand[u];
                        if (staged > SCAP) {
                            ncand[u] = CCAP + 1;       // staging overflowed: exact fallback
                        } else {
                            __syncwarp();              // this warp's staged entries are visible to all its lanes
                            const float cutoff = ub[u] + mg[u];
                            int kept = 0;
                            for (int e0 = 0; e0 < staged; e0 += 32) {
                                const int e = e0 + lane;
                                float2 ent = make_float2(0.f, 0.f);
                                if (e < staged) ent = stage[(size_t)j * SCAP + e];
                                const bool pass = e < staged && ent.x <= cutoff;
                                const unsigned bm = __ballot_sync(0xffffffffu, pass);
                                const int idx = kept + __popc(bm & ((1u << lane) - 1u));
                                if (pass && idx < CCAP) cand[qi * CCAP + idx] = __float_as_int(ent.y);
                                kept += __popc(bm);
                            }
                            ncand[u] = kept;
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = ew + 4 * u;
                if (j < cnt && lane == 0) {
                    const int64_t qi = qorder[cur.p0 + j];
                    if (ncand[u] > CCAP) {
                        candcnt[qi] = -1;
                        fb_list[atomicAdd(fb_count, 1)] = (int32_t)(cur.p0 + j);   // sorted position
                        ++n_fb;
                    } else {
                        candcnt[qi] = ncand[u];
                        n_cand += (unsigned long long)ncand[u];
                    }
                }
            }
            // ---- hand the item to the re-check warps: its candidate lists (global memory, written above by this
            // warp's lanes) are published by the arrive (release; the waiters acquire) ----
            {
                const uint32_t slot = it % RCRING, use = it / RCRING;
                mbar_wait_warp(&S->rcempty[slot], (use & 1) ^ 1, lane);     // the re-check warps are done with the slot
                if (ew == 0 && lane == 0) {
                    S->rc_p0[slot] = (int)cur.p0;
                    S->rc_cnt[slot] = cnt;
                }
                mbar_arrive_warp(&S->rcfull[slot], lane);
            }
            ++nproc;
            cur = nxt;
        }
        {   // stop marker after the last item of this CTA
            const uint32_t slot = nproc % RCRING, use = nproc / RCRING;
            mbar_wait_warp(&S->rcempty[slot], (use & 1) ^ 1, lane);
            if (ew == 0 && lane == 0) S->rc_cnt[slot] = -1;
            mbar_arrive_warp(&S->rcfull[slot], lane);
        }
        if (lane == 0) {
            if (n_fb) atomicAdd(&totals[0], n_fb);
            if (n_cand) atomicAdd(&totals[1], n_cand);
        }
    } else {
        // ================================ exact re-check + tail ========================================
        // Warps 16..23: as soon as the epilogue has selected an item's candidates, evaluate their canonical fp32
        // distances, keep the top-8 by (distance, position) and (BLEND) gather / weight / blend -- while the TMA, MMA and
        // splitter warps stream the next lists.  The candidate rows went through L2 on their way to shared memory a few
        // microseconds earlier, so most of these reads are L2 hits instead of a second trip to HBM.
        reg_dec<REGS_RC>();
        const int rw = warp - RC_WARP0;
        for (uint32_t seq = 0;; ++seq) {
            const uint32_t slot = seq % RCRING, use = seq / RCRING;
            mbar_wait_warp(&S->rcfull[slot], use & 1, lane);
            const int p0 = S->rc_p0[slot], cnt = S->rc_cnt[slot];
            if (cnt < 0) break;
            for (int j = rw; j < cnt; j += RCW)
                recheck_query<BLEND>((int64_t)p0 + j, q, qorder, d, vecs, ids, cand, candcnt, D, P, I, ba, lane);
            mbar_arrive_warp(&S->rcempty[slot], lane);
        }
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc<1>(tmem_base, TMEM_COLS);
}

}  // namespace stc

// Launch the fused kernel for work items [item_begin, item_end); called by scan_tc_range (scan_tc.cu) in place of the
// streaming kernel + scan_recheck_kernel when option scan_impl is 6.  Row norms, tensor maps and scratch are the caller's.
int launch_scan_tc_fused(rvcr_index* h, const CUtensorMap* tm, const float* q, int64_t n, float rel, float absm, int item_begin,
                         int item_end, int ticket_slot, int32_t* fb_count, float* D, int32_t* P, int64_t* I,
                         const BlendArgs& ba, int bmode, cudaStream_t s) {
    const int d = h->d, nkb = d / stc::KB;
    const size_t fixed = 2 * (size_t)nkb * stc::QKB_BYTES + stc::SV_BYTES + sizeof(stc::SmemF) + 1024;
    int nstage = (int)((226 * 1024 - fixed) / stc::STAGE_BYTES);
    if (nstage > 8) nstage = 8;
    if (nstage < 2) {
        set_error("d=%d leaves no room for the row ring in shared memory", d);
        return RVCR_ERR_UNSUPPORTED;
    }
    const size_t smem = fixed + (size_t)nstage * stc::STAGE_BYTES;
    static SmemOptIn opt_in[4];
#define RVCR_LAUNCH_FUSED(B)                                                                                                     \
    do {                                                                                                                         \
        RVCR_TRY(opt_in[B].ensure(stc::scan_tc_fused_kernel<B>, smem, h->device));                                               \
        stc::scan_tc_fused_kernel<B><<<h->num_sms, stc::FTHREADS, smem, s>>>(                                                    \
            tm[0], tm[1], tm[2], tm[3], q, h->ws_vals2.as<int32_t>(), h->ws_keys2.as<uint32_t>(), h->ws_items.as<int32_t>(),     \
            h->ws_misc.as<int32_t>(), h->ws_misc.as<int32_t>() + ticket_slot, h->offsets, h->row_norm.as<float>(),               \
            (uint32_t)h->nlist, n, d, nstage, h->row_norm_max, rel, absm, item_begin, item_end, h->ws_cand.as<int32_t>(),        \
            h->ws_candcnt.as<int32_t>(), h->ws_fb.as<int32_t>(), fb_count,                                                       \
            reinterpret_cast<unsigned long long*>(h->ws_misc.as<int32_t>() + 48), h->ws_cstage.as<float2>(),                    \
            getenv("RVCR_STC_FLAGS") ? atoi(getenv("RVCR_STC_FLAGS")) : 0, h->vecs, h->ids, D, P, I, ba);                        \
    } while (0)
    if (bmode == 0) RVCR_LAUNCH_FUSED(0);
    else if (bmode == 1) RVCR_LAUNCH_FUSED(1);
    else if (bmode == 2) RVCR_LAUNCH_FUSED(2);
    else RVCR_LAUNCH_FUSED(3);
#undef RVCR_LAUNCH_FUSED
    RVCR_CHECK_LAUNCH();
    return RVCR_OK;
}

}  // namespace rvcr
