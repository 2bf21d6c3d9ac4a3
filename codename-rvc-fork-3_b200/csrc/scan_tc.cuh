// scan_tc.cuh -- pieces of the tensor-core list scan (scan_tc.cu: streaming scan + exact re-check): tile constants, the
// shared-memory control block, work-item fetch, query staging, and the pipeline.py:500-506 tail arithmetic.
#pragma once

#include <stddef.h>

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
constexpr int A_COL0 = NACC * QT;           // TMEM columns [64, 512): ring of split row blocks (the MMA A operands),
constexpr int NTS = 7;                      //   7 stages x (32 columns hi | 32 columns lo)
constexpr int A_STAGE_COLS = 2 * KB;
constexpr int TMEM_COLS = 512;
constexpr int IRING = 4;                    // work-item descriptor ring
constexpr int THREADS = 512;                // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 item scheduler, 4-7 epilogue, 8-15 row splitters
constexpr int CONV_GROUPS = 2;              // two groups of 4 splitter warps take alternate stages (two stages in flight)
constexpr int EPI_WARP0 = 4;
constexpr int EPI_THREADS = 128;
constexpr int CONV_WARP0 = 8;
constexpr int CONV_THREADS = 128;
constexpr int CCAP = 64;                    // candidates handed to the re-check per query
constexpr int SCAP = 256;                   // (v, position) pairs staged per query while a multi-chunk list streams
#ifndef RVCR_STC_CHUNK_TILES
#define RVCR_STC_CHUNK_TILES 4
#endif
constexpr int CHUNK_TILES = RVCR_STC_CHUNK_TILES;   // the epilogue selects candidates per chunk of 4 tiles = 512 list rows
constexpr int CHUNK_ROWS = CHUNK_TILES * TILE_M;
constexpr int SV_BYTES = QT * CHUNK_ROWS * 4;   // v of one chunk: [16 queries][512 rows] fp32 = 32 KB
constexpr int MAX_STAGES = 8;

struct Item {
    int idx;   // ticket; >= nitems: stop
    int cnt;   // queries in the item (1..16)
    long long p0, lb, len;
};

struct Smem {
    uint64_t full[MAX_STAGES], empty[MAX_STAGES];   // shared-memory row ring: TMA -> splitters -> TMA
    uint64_t conv[NTS], aempty[NTS];       // TMEM operand ring: splitters -> MMA -> splitters
    uint64_t tfull[NACC], tempty[NACC];
    uint64_t ifull[IRING], iempty[IRING];
    uint64_t qfull[2];
    Item items[IRING];
    float2 qnorm[2][QT];      // (||q||^2, ||q - tf32(q)||^2) of the queries staged in sQ[buf]
    uint32_t tmem_ptr;
};

using tc::mbar_arrive_warp;
using tc::mbar_wait_warp;

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

// queries of an item -> K-major, 128B-swizzled operand tile [kb][16 rows][32 floats]; 8 threads per query, thread
// (j, c) copies 16-byte chunk c of every k-block of query j (a 128 B contiguous read per 8 threads); rows >= cnt
// replicate the last query (finite values; their accumulator columns are never read)
__device__ __forceinline__ void load_queries(uint8_t* sQ, const float* __restrict__ q, const PushArgs& push,
                                             const int32_t* __restrict__ qorder, const Item& it, int d, int nkb, int et) {
    const int j = et >> 3, c16 = et & 7;
    const int jj = j < it.cnt ? j : it.cnt - 1;
    const float* src = query_row(q, push, qorder[it.p0 + jj], d) + c16 * 4;
    uint8_t* dst = sQ + (j >> 3) * 1024 + (j & 7) * 128 + ((c16 ^ (j & 7)) << 4);
    for (int kb = 0; kb < nkb; ++kb) cp_async16(dst + kb * QKB_BYTES, src + kb * KB);
}

// After the copies landed: round the staged queries to tf32 in place (round to nearest, so the tensor core reads
// them exactly) and record ||q||^2 and ||q - tf32(q)||^2 -- the part of <q, y> the MMA cannot see, bounded in the
// margin by Cauchy-Schwarz with its exact norm.
__device__ __forceinline__ void convert_queries(uint8_t* sQ, float2* qnorm, int nkb, int et) {
    const int j = et >> 3, c16 = et & 7;
    uint8_t* p = sQ + (j >> 3) * 1024 + (j & 7) * 128 + ((c16 ^ (j & 7)) << 4);
    float n2 = 0.f, l2 = 0.f;
    for (int kb = 0; kb < nkb; ++kb) {
        float4 f = *reinterpret_cast<float4*>(p + kb * QKB_BYTES);
        float4 h;
        h.x = tc::to_tf32_rna(f.x);
        h.y = tc::to_tf32_rna(f.y);
        h.z = tc::to_tf32_rna(f.z);
        h.w = tc::to_tf32_rna(f.w);
        *reinterpret_cast<float4*>(p + kb * QKB_BYTES) = h;
        n2 = fmaf(f.x, f.x, fmaf(f.y, f.y, fmaf(f.z, f.z, fmaf(f.w, f.w, n2))));
        const float e0 = f.x - h.x, e1 = f.y - h.y, e2 = f.z - h.z, e3 = f.w - h.w;   // exact
        l2 = fmaf(e0, e0, fmaf(e1, e1, fmaf(e2, e2, fmaf(e3, e3, l2))));
    }
#pragma unroll
    for (int off = 1; off < 8; off <<= 1) {
        n2 += __shfl_xor_sync(0xffffffffu, n2, off);
        l2 += __shfl_xor_sync(0xffffffffu, l2, off);
    }
    if (c16 == 0) qnorm[j] = make_float2(n2, l2);
}

}  // namespace stc

// pipeline.py:500-506 for one query whose sorted top-8 sits in lanes 0..7 of the warp (td, tp; tp == INT_MAX = pad):
// weights exactly as blend.cu computes them (divide, square, pairwise row sum, divide); every lane returns the
// normalised weight / gather position of neighbour `lane & 7`.
struct BlendArgs {
    float* out;             // blended rows are written to out + qi * d (may alias the queries: in place)
    int64_t last_row_pos;   // numpy's big_npy[-1] for (FLT_MAX, -1) padding
    float r32, omr32;
    PostArgs post;          // optional fused 2x upsample + protect mix: rows go to post.out2 instead of `out`
    PushArgs push;          // optional fused multi-GPU combine: rows go to their home rank's buffer (peer memory)
};

__device__ __forceinline__ void blend_weights(float td, int tp, int lane, int64_t last_row_pos, float& wn, int& spos) {
    float w = 0.f;
    int pos = -1;
    if (lane < RVCR_K) {
        const bool pad = tp == 0x7fffffff;
        const float score = pad ? RVCR_FLT_MAX : td;
        const float inv = __fdiv_rn(1.0f, score);
        w = __fmul_rn(inv, inv);
        pos = pad ? -1 : tp;
    }
    float sum = __fadd_rn(w, __shfl_xor_sync(0xffffffffu, w, 1));
    sum = __fadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, 2));
    sum = __fadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, 4));
    wn = __fdiv_rn(w, sum);
    spos = pos < 0 ? (int)last_row_pos : pos;
}

// out[c] = fl(fl(acc * r32) + fl(omr32 * f)), acc = products rounded then summed k = 0..7 (blend.cu's contract)
__device__ __forceinline__ float4 blend_vec(const float4 (&v)[RVCR_K], const float (&wk)[RVCR_K], const float4 f, float r32,
                                            float omr32) {
    float4 acc;
    acc.x = __fmul_rn(v[0].x, wk[0]);
    acc.y = __fmul_rn(v[0].y, wk[0]);
    acc.z = __fmul_rn(v[0].z, wk[0]);
    acc.w = __fmul_rn(v[0].w, wk[0]);
#pragma unroll
    for (int k = 1; k < RVCR_K; ++k) {
        acc.x = __fadd_rn(acc.x, __fmul_rn(v[k].x, wk[k]));
        acc.y = __fadd_rn(acc.y, __fmul_rn(v[k].y, wk[k]));
        acc.z = __fadd_rn(acc.z, __fmul_rn(v[k].z, wk[k]));
        acc.w = __fadd_rn(acc.w, __fmul_rn(v[k].w, wk[k]));
    }
    float4 o;
    o.x = __fadd_rn(__fmul_rn(acc.x, r32), __fmul_rn(omr32, f.x));
    o.y = __fadd_rn(__fmul_rn(acc.y, r32), __fmul_rn(omr32, f.y));
    o.z = __fadd_rn(__fmul_rn(acc.z, r32), __fmul_rn(omr32, f.z));
    o.w = __fadd_rn(__fmul_rn(acc.w, r32), __fmul_rn(omr32, f.w));
    return o;
}

}  // namespace rvcr
