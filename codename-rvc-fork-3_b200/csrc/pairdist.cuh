// pairdist.cuh -- register-tiled canonical squared-L2 between a resident query tile and a
// stream of database rows (list vectors or centroids), shared by scan_topk.cu and coarse_exact.cu.
//
// Numerics: every (query, row) distance is evaluated in the exact operation order of faiss
// fvec_L2sqr: eight accumulators, accumulator l sums (x[8i+l]-y[8i+l])^2 for
// i = 0,1,2.. sequentially with a fused multiply-add, then the (a0+a4)+(a1+a5) .. reduction.
// To get that order on a GPU, eight threads cooperate on a pair: thread l owns accumulator l
// (elements = l mod 8), the final reduction is an 8-lane xor butterfly (reduce8_canonical).
//
// CTA = 256 threads = 8 warps = 32 groups of 8 threads.  Warp w works on rows {4w .. 4w+3} of the
// current 32-row block; its 4 groups (qg = 0..3) work on queries {qg + 4a, a < TQ}.  Per 8-float
// step a thread loads TQ query scalars + 4 row scalars from shared memory and issues TQ*4
// (FSUB, FFMA) pairs:
//   - row reads: the 4 groups of a warp read the same address -> broadcast, conflict-free;
//   - query reads: row stride (d + 8) floats = 8 (mod 32) banks -> (qg, l) hit 32 distinct banks.
// Rows are streamed through an NSTAGE-deep cp.async ring of [32 rows][KC = 128 floats] chunks
// (512 B contiguous per row segment, 16 KB per stage); all chunks of a 32-row block are in
// flight together so DRAM sees whole 3 KB rows.
#pragma once

#include "common.cuh"

namespace rvcr {

constexpr int PD_THREADS = 256;
constexpr int PD_RB = 32;   // rows per block
constexpr int PD_KC = 128;  // floats per k-chunk

template <int TQ>
struct PairTile {
    static constexpr int QT = 4 * TQ;  // queries per CTA tile
    float acc[TQ][4];

    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int a = 0; a < TQ; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    }

    // one k-chunk: sQ points at this thread's first query row at column (kc*KC + l),
    // sR at this thread's first block row at column l of the stage.
    __device__ __forceinline__ void chunk(const float* __restrict__ sQ, int qstride,
                                          const float* __restrict__ sR) {
#pragma unroll 4
        for (int s = 0; s < PD_KC / 8; ++s) {
            float qv[TQ], rv[4];
#pragma unroll
            for (int a = 0; a < TQ; ++a) qv[a] = sQ[(4 * a) * qstride + 8 * s];
#pragma unroll
            for (int b = 0; b < 4; ++b) rv[b] = sR[b * PD_KC + 8 * s];
#pragma unroll
            for (int a = 0; a < TQ; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    float t = __fsub_rn(qv[a], rv[b]);
                    acc[a][b] = __fmaf_rn(t, t, acc[a][b]);
                }
        }
    }

    __device__ __forceinline__ void reduce() {
#pragma unroll
        for (int a = 0; a < TQ; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = reduce8_canonical(acc[a][b]);
    }
};

// Issue the cp.async copies of stage `it` (= row-block it / nkc, k-chunk it % nkc) of a row stream
// whose row r lives at base + min(r, rmax) * d.  256 threads x 4 x 16 B = 16 KB.
__device__ __forceinline__ void pd_load_stage(float* sStage, const float* __restrict__ base, int64_t row0,
                                              int64_t rmax, int d, int kc, int tid) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int idx = tid + j * PD_THREADS;  // 0..1023
        int r = idx >> 5;                // row in block
        int c = idx & 31;                // float4 column in chunk
        int64_t gr = row0 + r;
        if (gr > rmax) gr = rmax;
        cp_async16(sStage + r * PD_KC + c * 4, base + gr * d + kc * PD_KC + c * 4);
    }
}

}  // namespace rvcr
