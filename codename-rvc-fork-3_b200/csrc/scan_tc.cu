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
//   epilogue v = ||y||^2 - 2 acc  (||y||^2 precomputed per stored row) -> shared memory, [query][row] per chunk
//            of 512 rows
//   select   per query: U = 8th smallest of the 32 lane minima (an upper bound of the 8th smallest v, and
//            within a row or two of it); every row with v <= U + margin(q) is a candidate (ballot-compacted)
//   recheck  canonical fp32 distance (fvec_L2sqr order, pairdist.cuh) of each candidate, top-8 by
//            (distance, position): bit-identical to the exact scan
//
// margin(q) bounds twice the worst-case |v - v_true| (tf32 truncation of both operands, Cauchy-Schwarz;
// fp32 accumulation of d products; norm roundings), so the true top-8 are always among the candidates
// (DESIGN.md "Numerics").  If more than 64 rows qualify, the query is re-done by the brute-force exact
// kernel below.
#include <stddef.h>
#include <stdlib.h>

#include "scan_tc.cuh"

namespace rvcr {

namespace stc {

__global__ void __launch_bounds__(THREADS, 1)
scan_tc_kernel(const __grid_constant__ CUtensorMap tm32, const __grid_constant__ CUtensorMap tm64,
               const __grid_constant__ CUtensorMap tm96, const __grid_constant__ CUtensorMap tm128,
               const float* __restrict__ q, const int32_t* __restrict__ qorder, const uint32_t* __restrict__ keys,
               const int32_t* __restrict__ items, const int32_t* __restrict__ nitems_ptr, int32_t* __restrict__ ticket,
               const int64_t* __restrict__ offsets, const float* __restrict__ ynorm, uint32_t nlist, int64_t n, int d,
               int nstage, float ymax, float acc_rel, float abs_margin, int item_begin, int item_end,
               int32_t* __restrict__ cand, int32_t* __restrict__ candcnt, int32_t* __restrict__ fb_list,
               int32_t* __restrict__ fb_count, unsigned long long* __restrict__ totals /*[0] fallbacks, [1] candidates*/,
               float2* __restrict__ stage_buf /*[grid][16][SCAP]*/, int dbg_flags, PushArgs push) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int nkb = d / KB;
    uint8_t* sR = smem;                                         // nstage x 16 KB
    uint8_t* sQ = sR + (size_t)nstage * STAGE_BYTES;            // 2 x nkb x 2 KB
    const int qbuf_bytes = nkb * QKB_BYTES;
    float* sV = reinterpret_cast<float*>(sQ + 2 * (size_t)qbuf_bytes);    // [16][CHUNK_ROWS]
    Smem* S = reinterpret_cast<Smem*>(reinterpret_cast<uint8_t*>(sV) + SV_BYTES);

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
    } else if (warp >= CONV_WARP0) {
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
    } else if (warp >= EPI_WARP0) {
        // ================================ epilogue =====================================================
        const int ew = warp - EPI_WARP0;          // == warp % 4: the TMEM lane quarter this warp may read
        const int et = threadIdx.x - EPI_WARP0 * 32;
        uint32_t tile_cnt = 0;
        unsigned long long n_fb = 0, n_cand = 0;    // lane 0: running totals for rvcr_get_stats
        mbar_wait_warp(&S->ifull[0], 0, lane);
        Item cur = S->items[0];
        mbar_arrive_warp(&S->iempty[0], lane);
        if (cur.idx < nitems) {
            load_queries(sQ, q, push, qorder, cur, d, nkb, et);
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
            uint8_t* sQn = sQ + (size_t)((it + 1) & 1) * qbuf_bytes;
            // buffer (it+1)&1 was last read by item it-1, whose MMAs and selection are complete.  (One buffer and three more row
            // stages instead -- the load then waits for the item's last MMA -- measured slower: scan phase 32.4 vs 30.8 ms,
            // round 2 call L.)
            if (have_next) load_queries(sQn, q, push, qorder, nxt, d, nkb, et);
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
                        convert_queries(sQn, S->qnorm[(it + 1) & 1], nkb, et);
                        tc::fence_proxy_async_smem();
                        mbar_arrive_warp(&S->qfull[(it + 1) & 1], lane);
                        next_published = true;
                    }
                }
                if (!next_published) {
                    cp_async_wait<0>();
                    convert_queries(sQn, S->qnorm[(it + 1) & 1], nkb, et);
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
// BLEND: the warp that has just found a query's exact top-8 also gathers those 8 rows (it read them a moment ago:
// L1 / L2 hits instead of a second trip to HBM), applies the 1/score^2 weights and the index_rate blend and writes
// the output row -- rvc/infer/pipeline.py:500-506 without a separate gather kernel.
// BLEND: 0 = search only, 1 = + fused tail in place, 2 = + fused tail + 2x upsample / protect mix into post.out2, 3 = + fused
// tail pushed to the home rank's buffer over NVLink (separate instantiations so the default kernel's code and register
// count stay exactly as profiled).
// UNR: float4 loads in flight per lane in the distance loop (8 as profiled in round 1; 16 is a tuning variant, RVCR_RC_UNROLL).
template <int BLEND, int UNR = 8>
__global__ void __launch_bounds__(RC_WARPS * 32)
scan_recheck_kernel(const float* __restrict__ q, const int32_t* __restrict__ qorder, const uint32_t* __restrict__ keys,
                    uint32_t nlist, const int32_t* __restrict__ items, const int32_t* __restrict__ nitems_ptr, int item_begin,
                    int item_end, int64_t n, int d, const float* __restrict__ vecs, const int64_t* __restrict__ ids,
                    const int32_t* __restrict__ cand, const int32_t* __restrict__ candcnt, float* __restrict__ D,
                    int32_t* __restrict__ P, int64_t* __restrict__ I, BlendArgs ba) {
    extern __shared__ __align__(16) float rc_smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float* sq = rc_smem + (size_t)w * d;
    const int nit = *nitems_ptr;
    const int64_t first = item_begin < nit ? (int64_t)items[item_begin] : n;
    const int64_t last = item_end < nit ? (int64_t)items[item_end] : n;
    const int h = lane & 1;
    const int nvec = d >> 2;
    for (int64_t p = first + (int64_t)blockIdx.x * RC_WARPS + w; p < last; p += (int64_t)gridDim.x * RC_WARPS) {
        if (keys[p] == nlist + 1u) continue;   // another shard's query (sorted last; forms no work item)
        const int64_t qi = qorder[p];
        const int nc = candcnt[qi];
        if (nc < 0) continue;   // overflowed: the exact fallback kernel answers this query
        float td = RVCR_FLT_MAX;
        int tp = 0x7fffffff;
        if (nc > 0 || BLEND) {
            __syncwarp();
            const float4* q4 = reinterpret_cast<const float4*>(query_row(q, ba.push, qi, d));
            for (int c = lane; c < nvec; c += 32) reinterpret_cast<float4*>(sq)[c] = BLEND == 3 ? q4[c] : __ldcs(q4 + c);   // (streaming: see the gather below)
            __syncwarp();
        }
        if (nc > 0) {
            for (int base = 0; base < nc; base += 16) {
                const int c = base + (lane >> 1);
                const bool have = c < nc;
                const int pos = cand[qi * stc::CCAP + (have ? c : 0)];
                const float4* y4 = reinterpret_cast<const float4*>(vecs + (int64_t)pos * d) + h;
                const float4* x4 = reinterpret_cast<const float4*>(sq) + h;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll UNR
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
                // What is touched for the last time -- the query row above, these 8 rows, the output row -- is loaded / stored
                // evict-first (ld.global.cs / st.global.cs): it then does not push the candidate rows that other warps still
                // have to gather out of L2.  Tail kernel 11.3 -> 10.3 ms at C5 (round 2, call Q); loading the candidate rows
                // evict-last on top of that changed nothing (10.5).
                // (The multi-GPU instantiation, BLEND == 3, keeps the plain loads / stores it was validated with on 2 and 8 GPUs:
                // its query and output rows live in peer memory.)
                for (int k = 0; k < RVCR_K; ++k)
                    v[k] = rows[k] ? (BLEND == 3 ? __ldg(rows[k] + c) : __ldcs(rows[k] + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 f0 = reinterpret_cast<const float4*>(sq)[c];
                const float4 o = blend_vec(v, wk, f0, ba.r32, ba.omr32);
                if (BLEND == 2) post_store(ba.post, qi, d, c, o, f0);
                else if (BLEND == 3) o4[c] = o;
                else __stcs(o4 + c, o);
            }
        }
    }
}

// Brute-force exact scan of one list for the queries the tensor-core pass could not bound (rare): one CTA
// per query, 8 lanes per row (canonical order), 32 rows per step.
template <int BLEND>
__global__ void __launch_bounds__(256)
scan_exact_fallback_kernel(const float* __restrict__ q, const int32_t* __restrict__ qorder, const uint32_t* __restrict__ keys,
                           const int32_t* __restrict__ fb_list, const int32_t* __restrict__ fb_count, uint32_t nlist, int d,
                           const float* __restrict__ vecs, const int64_t* __restrict__ ids, const int64_t* __restrict__ offsets,
                           float* __restrict__ D, int32_t* __restrict__ P, int64_t* __restrict__ I, BlendArgs ba) {
    extern __shared__ __align__(16) float fb_smem[];
    float* sq = fb_smem;          // d
    float* sd = fb_smem + d;      // 32
    float* sw = fb_smem + d + 32; // 8 weights
    int* sp = reinterpret_cast<int*>(fb_smem + d + 40);   // 8 gather positions
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
        {
            const float* qr = query_row(q, ba.push, qi, d);
            for (int c = tid; c < d; c += 256) sq[c] = qr[c];
        }
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
        if (BLEND) {
            if (w == 0) {
                float wn;
                int spos;
                blend_weights(td, tp, lane, ba.last_row_pos, wn, spos);
                if (lane < RVCR_K) {
                    sw[lane] = wn;
                    sp[lane] = spos;
                }
            }
            __syncthreads();
            float wk[RVCR_K];
            const float4* rows[RVCR_K];
#pragma unroll
            for (int k = 0; k < RVCR_K; ++k) {
                wk[k] = sw[k];
                rows[k] = sp[k] >= 0 ? reinterpret_cast<const float4*>(vecs + (int64_t)sp[k] * d) : nullptr;
            }
            float4* o4 = BLEND == 3 ? push_row_ptr(ba.push, qi, d) : reinterpret_cast<float4*>(ba.out + qi * d);
            for (int c = tid; c < (d >> 2); c += 256) {
                float4 v[RVCR_K];
#pragma unroll
                for (int k = 0; k < RVCR_K; ++k) v[k] = rows[k] ? __ldg(rows[k] + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 f0 = reinterpret_cast<const float4*>(sq)[c];
                const float4 o = blend_vec(v, wk, f0, ba.r32, ba.omr32);
                if (BLEND == 2) post_store(ba.post, qi, d, c, o, f0);
                else o4[c] = o;
            }
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
    const size_t fixed = 2 * (size_t)nkb * stc::QKB_BYTES + stc::SV_BYTES + sizeof(stc::Smem) + 1024;
    int nstage = (int)((226 * 1024 - fixed) / stc::STAGE_BYTES);
    if (nstage > stc::MAX_STAGES) nstage = stc::MAX_STAGES;
    // Stages vs time of the streaming kernel at C5 (round 2, calls K - M): 3 / 4 / 5 / 6 stages -> 35.7 / 28.8 / 23.8 / 19.7 ms, but 9 stages
    // (one query operand buffer instead of two) were slower (the per-item query load then stalls the MMAs), and a TMA L2
    // prefetch (cp.async.bulk.prefetch.tensor) 3 .. 48 stages ahead of the ring cost +6 .. +11 ms: past 6 stages the kernel is
    // bound by what the DRAM system delivers for 128-byte segments at a 3 KB stride (70 % of its cycles active), not by bytes
    // in flight.
    if (const char* e = getenv("RVCR_STC_NSTAGE")) {   // tuning knob
        const int v = atoi(e);
        if (v >= 2 && v < nstage) nstage = v;
    }
    if (nstage < 2) {
        set_error("d=%d leaves no room for the row ring in shared memory", d);
        return RVCR_ERR_UNSUPPORTED;
    }
    const size_t smem = fixed + (size_t)nstage * stc::STAGE_BYTES;
    static SmemOptIn opt_in;
    RVCR_TRY(opt_in.ensure(stc::scan_tc_kernel, smem, h->device));
    RVCR_TRY(h->ws_cand.reserve((size_t)n * stc::CCAP * 4));
    RVCR_TRY(h->ws_candcnt.reserve((size_t)n * 4));
    RVCR_TRY(h->ws_fb.reserve((size_t)n * 4 + 64));
    RVCR_TRY(h->ws_cstage.reserve((size_t)h->num_sms * stc::QT * stc::SCAP * sizeof(float2)));
    int32_t* fb_count = h->ws_misc.as<int32_t>() + 14;
    RVCR_CUDA(cudaMemsetAsync(fb_count, 0, 4, s));
    if (fill) {
        RVCR_TRY(fill_pad(D, P, I, n * RVCR_K, s));
    }
    // relative error of one accumulated dot product beyond the dropped query low part: truncation of the row
    // low part to tf32 (2^-20) and fp32 accumulation of 2 d products in the tensor core (d 2^-21)
    const float rel = 1.f / 1048576.f + (float)d / 2097152.f;
    const float absm = 1.f / 524288.f;
    int grid = h->num_sms;
    const CUtensorMap* tm = reinterpret_cast<const CUtensorMap*>(h->scan_tmaps);
    stc::scan_tc_kernel<<<grid, stc::THREADS, smem, s>>>(
        tm[0], tm[1], tm[2], tm[3], q, h->ws_vals2.as<int32_t>(), h->ws_keys2.as<uint32_t>(), h->ws_items.as<int32_t>(),
        h->ws_misc.as<int32_t>(), h->ws_misc.as<int32_t>() + ticket_slot, h->offsets, h->row_norm.as<float>(), (uint32_t)h->nlist, n, d,
        nstage, h->row_norm_max, rel, absm, item_begin, item_end, h->ws_cand.as<int32_t>(), h->ws_candcnt.as<int32_t>(),
        h->ws_fb.as<int32_t>(), fb_count, reinterpret_cast<unsigned long long*>(h->ws_misc.as<int32_t>() + 48),
        h->ws_cstage.as<float2>(), getenv("RVCR_STC_FLAGS") ? atoi(getenv("RVCR_STC_FLAGS")) : 0, h->fuse.push);
    RVCR_CHECK_LAUNCH();
    // fused tail (search_device sets h->fuse for rvcr_retrieve_blend*): the re-check / fallback kernels also blend
    const bool blend = h->fuse.active;
    BlendArgs ba{h->fuse.out, h->last_row_pos, h->fuse.r32, h->fuse.omr32, h->fuse.post, h->fuse.push};
    const int bmode = !blend ? 0 : (ba.post.out2 ? 2 : (ba.push.bases ? 3 : 1));
    {
        PhaseTimer pt(h, RVCR_PHASE_SCAN_RECHECK, s);
        // One warp per query, 32 warps per SM.  A variant that folded the re-check into the persistent scan kernel as extra
        // warp roles (round 2, "scan_impl 6": 615 lines, bit-identical results) measured 47.7 ms for the phase against
        // 32.6 ms for the two kernels and was deleted (profiles/r02_summary.md).  A variant that staged the candidate rows in shared memory with
        // cp.async.bulk (so the fused gather never left the SM) was measured slower (r01: 14.2 vs 8.1 ms at C5 -- with
        // 52 KB per query only 4 warps fit per SM, and the per-query chain header -> copies -> distances -> top-8 ->
        // blend is serial inside a warp) and was removed.
        {
            const size_t rsm = (size_t)RC_WARPS * d * 4;
            // CTAs per SM x float4 loads in flight per lane, measured at C5 (ms for the kernel): r01 10 x 8: 12.5; r02 call I
            // (RVCR_RC_CTAS x RVCR_RC_UNROLL): 2 x 16: 25.1, 3 x 16: 18.1, 4 x 16: 15.1, 4 x 8: 17.5, 6 x 16: 12.5, 10 x 8: 11.6,
            // 8 x 16: 11.2 (twice) -> default.  Fewer resident warps do NOT buy L2 hits for the second read of the final 8 rows;
            // the kernel wants bytes in flight.  (prefetch.global.L2 of every candidate row up front: 12.3 ms -- call M.)
            int per_sm = 8, unr = 16;
            if (const char* e = getenv("RVCR_RC_CTAS")) per_sm = atoi(e) > 0 ? atoi(e) : per_sm;   // tuning knobs
            if (const char* e = getenv("RVCR_RC_UNROLL")) unr = atoi(e) == 8 ? 8 : 16;
            int64_t rgrid = (int64_t)h->num_sms * per_sm;
            if (rgrid * RC_WARPS > n) rgrid = (n + RC_WARPS - 1) / RC_WARPS;
            auto kern = bmode == 0 ? scan_recheck_kernel<0, 16> : bmode == 1 ? scan_recheck_kernel<1, 16> : bmode == 2 ? scan_recheck_kernel<2, 16> : scan_recheck_kernel<3, 16>;
            if (unr == 8)
                kern = bmode == 0 ? scan_recheck_kernel<0> : bmode == 1 ? scan_recheck_kernel<1> : bmode == 2 ? scan_recheck_kernel<2> : scan_recheck_kernel<3>;
            kern<<<(unsigned)rgrid, RC_WARPS * 32, rsm, s>>>(q, h->ws_vals2.as<int32_t>(), h->ws_keys2.as<uint32_t>(), (uint32_t)h->nlist,
                                                             h->ws_items.as<int32_t>(), h->ws_misc.as<int32_t>(), item_begin, item_end, n,
                                                             d, h->vecs, h->ids, h->ws_cand.as<int32_t>(), h->ws_candcnt.as<int32_t>(),
                                                             D, P, I, ba);
        }
        RVCR_CHECK_LAUNCH();
    }
    {
        PhaseTimer pt(h, RVCR_PHASE_SCAN_FALLBACK, s);
        const size_t fsm = (size_t)(d + 48) * 4;
        auto kern = bmode == 0 ? scan_exact_fallback_kernel<0> : bmode == 1 ? scan_exact_fallback_kernel<1> : bmode == 2 ? scan_exact_fallback_kernel<2> : scan_exact_fallback_kernel<3>;
        kern<<<h->num_sms * 2, 256, fsm, s>>>(q, h->ws_vals2.as<int32_t>(), h->ws_keys2.as<uint32_t>(), h->ws_fb.as<int32_t>(), fb_count,
                                              (uint32_t)h->nlist, d, h->vecs, h->ids, h->offsets, D, P, I, ba);
        RVCR_CHECK_LAUNCH();
    }
    if (blend) h->fuse.done = true;
    // running total of fallback queries (read back lazily by rvcr_get_stats)
    h->scan_fb_dirty = true;
    count_launch(h, fill ? 4 : 3);
    return RVCR_OK;
}

}  // namespace rvcr
