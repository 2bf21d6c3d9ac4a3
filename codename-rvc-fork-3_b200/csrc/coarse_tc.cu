// coarse_tc.cu -- coarse-quantizer assignment on the 5th-gen tensor cores (tcgen05 + TMEM + TMA),
// made FP32-exact by an exact re-check.
//
// Replaces the same reference step as coarse_exact.cu: quantizer->search(n, x, k=1) inside
// index.search (rvc/infer/pipeline.py:499) and index.add (rvc/train/process/extract_index.py:69);
// faiss does this with an OpenBLAS sgemm.  At the BASELINE configs this contraction
// (2*nq*nlist*d flops) is ~256x the scan's flops, so it decides the end-to-end frames/s.
//
// Scheme ("validated split precision"):
//   1. prep        q, c  ->  bf16 operands (optionally a 3-way split [h|l|h] x [h|h|l] along K) and
//                  fp32 norms  (coarse_prep kernels)
//   2. tc kernel   v(q,j) = ||c_j||^2 - 2 * <q_bf16, c_bf16>  on tcgen05.mma (fp32 accumulation in
//                  TMEM); the epilogue keeps, per query, the 8 smallest v and their centroid ids
//   3. recheck     candidates whose v is within a rigorous error margin of the smallest v get
//                  their canonical fp32 distance (same fvec_L2sqr order as everywhere else);
//                  argmin by (distance, id).  If all 8 kept values lie inside the margin the
//                  candidate list may be truncated: the query is re-done by coarse_exact.cu.
//   The result is therefore identical to coarse_exact.cu bit for bit (tests assert this); the
//   margin derivation is in DESIGN.md "Numerics".
//
// tc kernel: persistent, warp-specialised, one CTA (or a CTA pair, cta_group::2) per SM (pair):
//   warp 0   TMA producer: A = 128 queries x 64 k, B = 256 (128 per CTA of a pair) centroids x 64 k,
//            128B-swizzled K-major tiles, 4-6 stage mbarrier ring
//   warp 1   MMA issuer (one elected thread; leader CTA only): 4 x tcgen05.mma (K = 16) per stage
//            into a 128 x 256 fp32 accumulator, double-buffered in TMEM (2 x 256 columns)
//   warp 2   TMEM allocator
//   warps 4-19 epilogue (4 lane quarters x 4 column groups): tcgen05.ld 16 / 32 columns at a time, key = 2 acc - ||c||^2,
//            per row the candidates within margin(q) of the running best (queued while the accumulator stage is held,
//            worked into the row's list after it has been handed back); overlaps the next tile's MMAs
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace rvcr {

namespace tc {

constexpr int BLOCK_M = 128;   // queries per CTA
constexpr int BLOCK_N = 256;   // centroids per accumulator tile
constexpr int ACC = 2;         // accumulator stages in TMEM (ACC x BLOCK_N = all 512 columns).  With S stages a tile takes
                               // max(M, (M + E) / S) (M = its MMAs, E = draining it).  4 x 128 columns were tried for the extra
                               // slack (round 2, call F) and lost badly -- bf16 106 -> 160 ms, e4m3 83 -> 93 ms: at N = 128 every
                               // MMA re-reads the A operand from shared memory for half the columns -- so E has to come down instead
constexpr int BLOCK_K = 64;    // bf16 elements per stage row = 128 B = one swizzle atom
constexpr int UMMA_K = 16;
constexpr int TOPT = 8;        // approximate nearest centroids kept per (query, split, column group)
constexpr int EPI_WARP0 = 4;
constexpr int NCG = 4;         // column groups of the epilogue: 4 lane quarters x NCG warps drain one accumulator tile (64 columns each)
constexpr int THREADS = (EPI_WARP0 + 4 * NCG) * 32;   // 640
// Epilogue granularity, measured at C5 (round 2, call J; coarse_main ms, previous kernel on the same box: e4m3 64.6, bf16 102.3):
//   columns per tcgen05.ld x queue slots     16 x 4         32 x 2         16 x 2
//   e4m3 (epilogue on the critical path)     61.6           65.1           63.6
//   bf16 (tensor pipe / power bound)         103.3          102.2          103.7
template <int F8> struct EpiCfg { static constexpr int SUBC = F8 ? 16 : 32, QCAP = F8 ? 4 : 2; };

// Candidate list of one (query, split, column group): up to TOPT (key, centroid) pairs, unsorted.  An element is looked at
// only if its key reaches thr = max(best - margin, smallest key kept once the list is full): it must be within the margin
// of the running best AND among the TOPT largest keys seen -- the first test alone lets everything through until the
// query's own neighbourhood shows up when the margin is as wide as the spread of the far centroids (e4m3), the second
// alone is round 1's sorted top-8.  A passing element takes a free slot (key -inf) or evicts the smallest key kept;
// whatever was evicted is remembered by its largest key in `dropped`.  At the end of the item a full list whose smallest
// key is itself inside the margin may have turned candidates away: it reports that key as dropped too, and the re-check
// hands the query to the exact kernel.
__device__ __forceinline__ float cand_insert(float (&ck)[TOPT], int (&cj)[TOPT], float& best, float& thr, float& dropped,
                                             float margin, float x, int j) {
    float vk = ck[0];
    int vs = 0;
#pragma unroll
    for (int i = 1; i < TOPT; ++i)
        if (ck[i] < vk) {
            vk = ck[i];
            vs = i;
        }
    if (x > vk) {
        dropped = fmaxf(dropped, vk);
#pragma unroll
        for (int i = 0; i < TOPT; ++i)
            if (i == vs) {
                ck[i] = x;
                cj[i] = j;
            }
    } else {
        dropped = fmaxf(dropped, x);
    }
    best = fmaxf(best, x);
    const float kmin = fminf(fminf(fminf(ck[0], ck[1]), fminf(ck[2], ck[3])), fminf(fminf(ck[4], ck[5]), fminf(ck[6], ck[7])));
    thr = fmaxf(best - margin, kmin);   // kmin = -inf while a slot is free
    return kmin;
}

// Order-preserving map float -> uint32 (so atomicMax on the mapped value is a float max); 0 is below every float and
// un-maps to a NaN, which fmaxf() ignores: "no floor yet".
__device__ __forceinline__ uint32_t key_map(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_unmap(uint32_t m) {
    return __uint_as_float((m & 0x80000000u) ? (m & 0x7fffffffu) : ~m);
}

// Work off a lane's queue of remembered keys (see the epilogue): each one that still reaches the threshold goes through
// cand_insert.  Returns the list's smallest key after the last insertion (kmin_in if there was none).
template <int QCAP>
__device__ __forceinline__ float cand_flush(float (&qx)[QCAP], int (&qj)[QCAP], int& qn, float (&ck)[TOPT], int (&cj)[TOPT],
                                            float& best, float& thr, float& dropped, float margin, float gfl, float& thr_e,
                                            float kmin) {
#pragma unroll 1
    for (int i = 0; i < qn; ++i) {
        float x = qx[0];
        int j = qj[0];
#pragma unroll
        for (int u = 1; u < QCAP; ++u)
            if (i == u) {
                x = qx[u];
                j = qj[u];
            }
        if (x >= thr_e) {
            kmin = cand_insert(ck, cj, best, thr, dropped, margin, x, j);
            thr_e = fmaxf(thr, gfl);
        }
    }
    qn = 0;
    return kmin;
}

// tcgen05.wait::ld with the loaded registers as in/out operands: nothing that reads them can be scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait_regs(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

__device__ __forceinline__ void tmem_ld_wait_regs(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}

__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32(taddr, r); }
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld16(taddr, r); }

// 1-D bulk copy global -> this CTA's shared memory, completing (bytes) on a local mbarrier
__device__ __forceinline__ void bulk_load_1d(uint32_t dst_addr, const void* src, uint32_t bytes, uint32_t bar_addr) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_addr),
                 "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar_addr)
                 : "memory");
}

// NRES = k-blocks of the query operand (A) kept resident in shared memory for a whole work item (one m-block
// against every centroid tile of its split); SB = stages of the centroid (B) ring; SA = slots for the A k-blocks
// that are still streamed.  NRES = 0 streams everything (SA == SB, A and B of a k-block share a stage).  The streamed
// k-blocks are spread evenly -- every P-th one (kb % P == P - 1), plus whatever exceeds the NRES resident slots --
// so that a freed A slot is reused about P k-blocks later and its TMA load never waits for the tensor pipe.
//
// Why: with both operands streamed, each CTA pair pulls 64 KB through the L2 per 512 tensor cycles -- 9.5 KB/clk
// chip-wide at full tensor rate, above the ~6.3 KB/clk the L2 slices deliver (ncu r01: lts 6.39 KB/clk, tensor pipe
// 68 % active).  The A block of an item is reused by all of its ~354 centroid tiles, so pinning most of it in shared
// memory cuts the L2 traffic to (NKB + streamed)/(2 NKB) of that and leaves the tensor pipe as the limiter.
template <int CLUSTER, int NRES, int SB, int SA, int P>
struct Cfg {
    static constexpr int LOAD_N = BLOCK_N / CLUSTER;                      // B rows loaded per CTA
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;                 // 16 KB
    static constexpr int B_BYTES = LOAD_N * BLOCK_K * 2;                  // 32 KB / 16 KB
    static constexpr int RES_BYTES = NRES * A_BYTES;
    static constexpr int RING_BYTES = SB * B_BYTES + SA * A_BYTES;
    static constexpr int NORM_BYTES = ACC * BLOCK_N * 4;                  // ||c||^2 of the tile in each accumulator stage
    static constexpr int FLOOR_BYTES = BLOCK_M * 4;                       // per-row floor shared by the NCG column-group warps
    static constexpr int BAR_BYTES = 512;
    // the dynamic shared memory of a kernel without static shared memory starts 1024-byte aligned (the kernel traps if not)
    static constexpr int SMEM_BYTES = RES_BYTES + RING_BYTES + NORM_BYTES + FLOOR_BYTES + BAR_BYTES;
    static constexpr int UMMA_M = BLOCK_M * CLUSTER;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static_assert(NRES > 0 || SA == SB, "fully streamed: one A slot per B stage");
    static_assert((2 * SB + SA + 2 + 4 * ACC) * 8 + 8 <= BAR_BYTES, "barrier block");
};

// grid = persistent CTAs (multiple of CLUSTER); work item = (m-block of 128*CLUSTER queries, split
// of the centroid tiles); item i is taken by cluster (i mod nclusters).
//
// Epilogue (round 2): the r01 kernel drained an accumulator tile with 4 warps, one 32-column tcgen05.ld at a time, the
// ||c||^2 of the columns fetched with 8 LDG.128 per load and 8 data-dependent branches per 32 columns; ncu (profiles/
// r01_ncu_summary.md) showed the tensor pipe idle 35 % of the time behind it (a third of the epilogue warps' samples on
// the first FFMA after the loads, a quarter on branch resolution) -- and e4m3 operands, which halve the MMA time, gained
// only 10 %.  Now 4 x NCG warps drain a tile (warp w: TMEM lane quarter w % 4, columns [cg 256/NCG, (cg + 1) 256/NCG) with
// cg = (w - 4) / 4), the tile's ||c||^2 arrive in shared memory by one 1 KB bulk copy per tile (producer warp, double-
// buffered with the accumulator stages) and are read as broadcast LDS.128, and the columns of one tcgen05.ld (16 for
// e4m3, 32 for bf16: EpiCfg) are looked at one by one only if their maximum reaches the row's current threshold (one
// branch per load).  A key that does is only remembered while the accumulator stage is held (register queue); the list
// maintenance (cand_insert) runs after the stage has gone back to the MMA warp.  Keys are -v = 2 <q', c'> - ||c||^2
// (larger = nearer).  Each (query, split, column group) writes its own list of up to 8; the re-check merges them.
//
// F8 = 1 (coarse_mode 4): e4m3 operands (kind::f8f6f4).  A 128-byte k-block then holds 128 elements
// instead of 64, every tile / descriptor / byte count stays as it is, the instruction descriptor and the MMA kind
// change, and the epilogue multiplies the accumulator by the per-query factor -qk[row] = 2 / (s_q(row) s_c) that undoes
// the power-of-two operand scales (bf16: the constant 2).
template <int CLUSTER, int NRES, int SB, int SA, int P, int F8 = 0>
__global__ void __launch_bounds__(THREADS, 1)
coarse_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const float* __restrict__ cn, int num_mblocks, int num_ntiles, int nsplit, int tiles_per_split,
                 int num_kblocks, float* __restrict__ topv, int* __restrict__ topj, float* __restrict__ topd,
                 const float* __restrict__ qk, const float* __restrict__ qmargin) {
    using C = Cfg<CLUSTER, NRES, SB, SA, P>;
    constexpr int KE = F8 ? 2 * BLOCK_K : BLOCK_K;   // elements per 128-byte k-block (TMA coordinates are in elements)
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sAres = smem;                               // NRES resident A k-blocks
    uint8_t* sB = smem + C::RES_BYTES;                   // SB stages
    uint8_t* sA = sB + SB * C::B_BYTES;                  // SA slots (streamed A k-blocks)
    float* sNorm = reinterpret_cast<float*>(smem + C::RES_BYTES + C::RING_BYTES);   // [2][BLOCK_N]
    uint32_t* sFloor = reinterpret_cast<uint32_t*>(smem + C::RES_BYTES + C::RING_BYTES + C::NORM_BYTES);   // [BLOCK_M], order-mapped keys
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::RES_BYTES + C::RING_BYTES + C::NORM_BYTES + C::FLOOR_BYTES);
    uint64_t* full_bar = bars;                      // [SB]  TMA -> MMA       (leader CTA's are used)
    uint64_t* empty_bar = full_bar + SB;            // [SB]  MMA -> TMA       (each CTA its own)
    uint64_t* aempty_bar = empty_bar + SB;          // [SA]  MMA -> TMA: streamed A slot free (each CTA its own)
    uint64_t* ares_full_bar = aempty_bar + SA;      // [1]   TMA -> MMA: resident A loaded (leader CTA's is used)
    uint64_t* ares_empty_bar = ares_full_bar + 1;   // [1]   MMA -> TMA: the item's MMAs are complete (each CTA its own)
    uint64_t* tmem_full_bar = ares_empty_bar + 1;     // [ACC] MMA -> epilogue  (each CTA its own)
    uint64_t* tmem_empty_bar = tmem_full_bar + ACC;   // [ACC] epilogue -> MMA  (leader CTA's are used)
    uint64_t* norm_full_bar = tmem_empty_bar + ACC;   // [ACC] bulk copy -> epilogue (each CTA its own)
    uint64_t* norm_empty_bar = norm_full_bar + ACC;   // [ACC] epilogue -> producer (each CTA its own)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(norm_empty_bar + ACC);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = CLUSTER == 1 ? 0u : cluster_ctarank();
    const bool leader = cta_rank == 0;
    const int cluster_id = blockIdx.x / CLUSTER;
    const int nclusters = gridDim.x / CLUSTER;
    const int nitems = num_mblocks * nsplit;
    // k-block kb is streamed when it is every P-th one or the resident slots are used up; else it lives in slot kb - kb / P
    auto is_streamed = [](int kb) { return NRES == 0 || kb % P == P - 1 || kb - kb / P >= NRES; };

    if ((smem_u32(smem) & 1023u) != 0) __trap();   // the swizzled tiles need 1024-byte alignment (see Cfg)
    if constexpr (CLUSTER > 1) cluster_sync_all();
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < SB; ++i) {
            mbar_init(&full_bar[i], CLUSTER);   // one arrive per CTA's producer (+ tx bytes)
            mbar_init(&empty_bar[i], 1);        // one tcgen05.commit
        }
        for (int i = 0; i < SA; ++i) mbar_init(&aempty_bar[i], 1);
        mbar_init(ares_full_bar, CLUSTER);
        mbar_init(ares_empty_bar, 1);
        for (int i = 0; i < ACC; ++i) {
            mbar_init(&tmem_full_bar[i], 1);                  // one tcgen05.commit
            mbar_init(&tmem_empty_bar[i], CLUSTER * 4 * NCG); // every epilogue warp of the pair
            mbar_init(&norm_full_bar[i], 1);                  // the producer's arrive (+ tx bytes)
            mbar_init(&norm_empty_bar[i], 4 * NCG);           // every epilogue warp of this CTA
        }
        fence_mbar_init();
    } else if (warp == 2) {
        tmem_alloc<CLUSTER>(tmem_ptr_smem, 512);
    }
    tc_fence_before();
    if constexpr (CLUSTER > 1) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    // The issuing warps run their loops converged (all loop bounds are kernel parameters / special registers, hence
    // uniform); only the tcgen05 / TMA instructions themselves are predicated on one elected lane.  Issuing from inside
    // an `if (elect)` region instead makes the compiler treat every descriptor as divergent and wrap each instruction in
    // an ELECT + R2UR.BROADCAST waterfall (~15 slow instructions per MMA).
    if (warp == 0) {
        // ================================ TMA producer =================================
        const uint32_t sAres_addr = tc::warp_uniform(smem_u32(sAres));
        const uint32_t sA_addr = tc::warp_uniform(smem_u32(sA)), sB_addr = tc::warp_uniform(smem_u32(sB));
        const uint32_t full_addr = tc::warp_uniform(smem_u32(full_bar));
        const uint32_t ares_full_addr = tc::warp_uniform(smem_u32(ares_full_bar));
        const uint32_t norm_addr = tc::warp_uniform(smem_u32(sNorm)), nfull_addr = tc::warp_uniform(smem_u32(norm_full_bar));
        uint32_t stage = 0, phase = 0, astage = 0, aphase = 0, item_cnt = 0, tile_cnt = 0;
        for (int item = cluster_id; item < nitems; item += nclusters, ++item_cnt) {
            const int mb = item / nsplit, sp = item - mb * nsplit;
            const int m0 = mb * (BLOCK_M * CLUSTER) + (int)cta_rank * BLOCK_M;
            const int t0 = sp * tiles_per_split;
            const int t1 = min(t0 + tiles_per_split, num_ntiles);
            if (NRES > 0) {
                // the previous item's MMAs no longer read the resident block
                mbar_wait_warp(ares_empty_bar, (item_cnt & 1) ^ 1, lane);
                if (elect_one()) {
                    int nres = 0;
                    for (int kb = 0; kb < num_kblocks; ++kb)
                        if (!is_streamed(kb)) {
                            tma_load_2d_a<CLUSTER>(&tmap_a, ares_full_addr, sAres_addr + (kb - kb / P) * C::A_BYTES, kb * KE, m0);
                            ++nres;
                        }
                    if (leader) mbar_arrive_expect_tx_addr(ares_full_addr, (uint32_t)nres * C::A_BYTES * CLUSTER);
                    else mbar_arrive_cluster_addr(ares_full_addr, 0);
                }
                __syncwarp();
            }
            for (int t = t0; t < t1; ++t, ++tile_cnt) {
                const int n0 = t * BLOCK_N + (int)cta_rank * C::LOAD_N;
                for (int kb = 0; kb < num_kblocks; ++kb) {
                    const bool streamed = is_streamed(kb);     // uniform
                    mbar_wait_warp(&empty_bar[stage], phase ^ 1, lane);
                    if (NRES > 0 && streamed) mbar_wait_warp(&aempty_bar[astage], aphase ^ 1, lane);
                    if (elect_one()) {
                        const uint32_t fb = full_addr + stage * 8;
                        const uint32_t aslot = NRES > 0 ? astage : stage;
                        if (streamed) tma_load_2d_a<CLUSTER>(&tmap_a, fb, sA_addr + aslot * C::A_BYTES, kb * KE, m0);
                        tma_load_2d_a<CLUSTER>(&tmap_b, fb, sB_addr + stage * C::B_BYTES, kb * KE, n0);
                        if (leader) mbar_arrive_expect_tx_addr(fb, (uint32_t)(C::B_BYTES + (streamed ? C::A_BYTES : 0)) * CLUSTER);
                        else mbar_arrive_cluster_addr(fb, 0);
                    }
                    __syncwarp();
                    if (++stage == SB) { stage = 0; phase ^= 1; }
                    if (NRES > 0 && streamed) {
                        if (++astage == SA) { astage = 0; aphase ^= 1; }
                    }
                }
                {
                    // this tile's ||c||^2 (all BLOCK_N columns: the epilogue of either CTA of a pair sees the whole tile).  Issued
                    // AFTER the tile's operand loads: the slot frees only when the epilogue of tile t - 2 is done, and waiting for
                    // that in front of the operand loads would expose their latency on every tile (round 2, call D: the e4m3
                    // kernel sat at 33 % tensor activity with idle epilogue warps); the norms are not needed before the tile's
                    // MMAs have finished.
                    const uint32_t ns = tile_cnt % ACC, nphase = (tile_cnt / ACC) & 1;
                    mbar_wait_warp(&norm_empty_bar[ns], nphase ^ 1, lane);
                    if (elect_one()) {
                        mbar_arrive_expect_tx_addr(nfull_addr + ns * 8, BLOCK_N * 4);
                        bulk_load_1d(norm_addr + ns * (BLOCK_N * 4), cn + (int64_t)t * BLOCK_N, BLOCK_N * 4, nfull_addr + ns * 8);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ===================================
        if (leader) {
            constexpr uint32_t idesc = F8 ? make_idesc_f8(C::UMMA_M, BLOCK_N) : make_idesc(C::UMMA_M, BLOCK_N);
            const uint32_t sAres_addr = tc::warp_uniform(smem_u32(sAres));
            const uint32_t sA_addr = tc::warp_uniform(smem_u32(sA)), sB_addr = tc::warp_uniform(smem_u32(sB));
            const uint32_t empty_addr = tc::warp_uniform(smem_u32(empty_bar)), tfull_addr = tc::warp_uniform(smem_u32(tmem_full_bar));
            const uint32_t aempty_addr = tc::warp_uniform(smem_u32(aempty_bar));
            const uint32_t ares_empty_addr = tc::warp_uniform(smem_u32(ares_empty_bar));
            const uint32_t tbase = tc::warp_uniform(tmem_base);
            uint32_t stage = 0, phase = 0, astage = 0, tile_cnt = 0, item_cnt = 0;
            for (int item = cluster_id; item < nitems; item += nclusters, ++item_cnt) {
                const int sp = item % nsplit;
                const int t0 = sp * tiles_per_split;
                const int t1 = min(t0 + tiles_per_split, num_ntiles);
                if (NRES > 0) {
                    mbar_wait_warp(ares_full_bar, item_cnt & 1, lane);     // this item's resident A block landed
                    tc_fence_after();
                }
                for (int t = t0; t < t1; ++t, ++tile_cnt) {
                    const uint32_t as = tile_cnt % ACC, aphase = (tile_cnt / ACC) & 1;
                    mbar_wait_warp(&tmem_empty_bar[as], aphase ^ 1, lane);   // epilogue drained this accumulator
                    tc_fence_after();
                    const uint32_t tmem_d = tbase + as * BLOCK_N;
                    for (int kb = 0; kb < num_kblocks; ++kb) {
                        const bool streamed = is_streamed(kb);     // uniform
                        mbar_wait_warp(&full_bar[stage], phase, lane);
                        tc_fence_after();
                        const uint32_t aslot = NRES > 0 ? astage : stage;
                        const uint64_t da = make_smem_desc(streamed ? sA_addr + aslot * C::A_BYTES : sAres_addr + (kb - kb / P) * C::A_BYTES);
                        const uint64_t db = make_smem_desc(sB_addr + stage * C::B_BYTES);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                                // advance 16 bf16 (32 e4m3) = 32 B inside the 128 B swizzle row: +2 in 16-byte units
                                if constexpr (F8)
                                    umma_f8<CLUSTER>(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                                     (kb > 0 || k > 0) ? 1u : 0u);
                                else
                                umma_bf16<CLUSTER>(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                                   (kb > 0 || k > 0) ? 1u : 0u);
                            }
                            umma_commit_a<CLUSTER>(empty_addr + stage * 8);   // frees the smem stage when the MMAs retire
                            if (NRES > 0 && streamed) umma_commit_a<CLUSTER>(aempty_addr + astage * 8);
                            if (kb == num_kblocks - 1) umma_commit_a<CLUSTER>(tfull_addr + as * 8);
                        }
                        __syncwarp();
                        if (++stage == SB) { stage = 0; phase ^= 1; }
                        if (NRES > 0 && streamed) {
                            if (++astage == SA) astage = 0;
                        }
                    }
                }
                if (NRES > 0) {
                    if (elect_one()) umma_commit_a<CLUSTER>(ares_empty_addr);   // resident block may be overwritten
                    __syncwarp();
                }
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ================================ epilogue =====================================
        constexpr int CGW = BLOCK_N / NCG;             // columns per column group
        constexpr int SUBC = EpiCfg<F8>::SUBC;         // columns per tcgen05.ld
        constexpr int NCH = CGW / SUBC;                // loads per warp and tile
        constexpr int QCAP = EpiCfg<F8>::QCAP;         // passing keys a lane can remember per tile before it has to work them off
        const int ew = (warp - EPI_WARP0) & 3;         // == warp % 4: TMEM lane quarter of this warp
        const int cg = (warp - EPI_WARP0) >> 2;        // column group
        const int row = ew * 32 + lane;                // accumulator row == TMEM lane
        uint32_t tile_cnt = 0;
        for (int item = cluster_id; item < nitems; item += nclusters) {
            const int mb = item / nsplit, sp = item - mb * nsplit;
            const int t0 = sp * tiles_per_split;
            const int t1 = min(t0 + tiles_per_split, num_ntiles);
            const int64_t q = (int64_t)mb * (BLOCK_M * CLUSTER) + (int64_t)cta_rank * BLOCK_M + row;
            float ck[TOPT];
            int cj[TOPT];
#pragma unroll
            for (int i = 0; i < TOPT; ++i) { ck[i] = __int_as_float(0xff800000); cj[i] = -1; }
            // margin(q) of the re-check (tc_margin, written by the prep kernel): only keys within it of the row's running best
            // can still be the nearest centroid.  A margin that is NaN or absurd collects nothing -> exact kernel.
            float margin = __ldg(qmargin + q);
            if (!(margin < 1e30f)) margin = __int_as_float(0x7fc00000);
            float best = __int_as_float(0xff800000), dropped = __int_as_float(0xff800000);
            float thr = margin == margin ? __int_as_float(0xff800000) : margin;   // -inf: the first columns all pass
            float nkr = 2.f;   // key = -v = nkr * acc - ||c||^2
            if constexpr (F8) nkr = -__ldg(qk + q);
            // The NCG warps that share these 32 rows (one per column group) also share a per-row floor: the largest "8th key" any
            // of their lists holds.  A key below it has 8 larger keys in that list, so it is not among the row's 8 best; testing
            // against the shared floor cuts the insertions per row from NCG x 8 ln(n / 8 NCG) to about 8 ln(n / 8).  (A candidate
            // turned away by it implies the list that set the floor is full of candidates, which that list reports itself.)
            named_bar_sync(1 + ew, NCG * 32);              // everybody is done with the previous item's floors
            if (cg == 0) sFloor[row] = 0u;
            named_bar_sync(1 + ew, NCG * 32);
            volatile uint32_t* myfloor = sFloor + row;
            for (int t = t0; t < t1; ++t, ++tile_cnt) {
                const uint32_t as = tile_cnt % ACC, aphase = (tile_cnt / ACC) & 1;
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + as * BLOCK_N + cg * CGW;
                const float4* sn4 = reinterpret_cast<const float4*>(sNorm + as * BLOCK_N + cg * CGW);
                const int j0 = t * BLOCK_N + cg * CGW;
                // the long wait of this warp: parked in hardware (mbarrier.try_wait with a suspend hint), not polled.  For e4m3 a polled
                // wait was tried twice (calls H and N) on the suspicion that a parked warp resumes late: no difference either time
                // (63.8 / 64.4 ms polled, 63.7 / 64.2 parked), and the spin loop was 36 % of the kernel's executed instructions.
                mbar_wait_warp_parked(&tmem_full_bar[as], aphase, lane);
                tc_fence_after();
                mbar_wait_warp(&norm_full_bar[as], aphase, lane);
                // Phase A -- the accumulator stage is held.  Read this warp's columns SUBC at a time; keys that reach the threshold are
                // only REMEMBERED (register queue of QCAP (key, centroid) pairs, a few instructions each; a full queue is worked off
                // on the spot).  The stage goes back to the MMA warp as soon as the last columns sit in registers, i.e. before any
                // list maintenance: with two accumulator stages a tile takes max(M, (M + hold) / 2), and the hold time used to contain
                // the slowest of the pair's 32 epilogue warps' insertions (e4m3, calls H / I: tiles took 6260 cycles against M = 3072
                // with the epilogue warps ~40 % busy and the MMA warp waiting for tmem_empty).
                float qx[QCAP];
                int qj[QCAP];
                int qn = 0;
                const float gfl = key_unmap(*myfloor);
                float thr_e = fmaxf(thr, gfl);   // (a NaN -- no floor yet, or a NaN margin in thr -- is ignored by fmaxf)
                float kmin = __int_as_float(0xff800000);
#pragma unroll 1
                for (int c = 0; c < NCH; ++c) {
                    uint32_t r[SUBC];
                    tmem_ld_cols(taddr + c * SUBC, r);
                    tmem_ld_wait_regs(r);
                    if (c == NCH - 1) {
                        // every accumulator column of this warp has been read: hand the stage back (leader CTA's barrier)
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (CLUSTER == 1 || leader) mbar_arrive_local(&tmem_empty_bar[as]);
                            else mbar_arrive_cluster(&tmem_empty_bar[as], 0);
                        }
                    }
                    float key[SUBC];
#pragma unroll
                    for (int g = 0; g < SUBC / 4; ++g) {
                        const float4 nn = sn4[c * (SUBC / 4) + g];   // same address in every lane: one broadcast LDS.128
                        key[4 * g + 0] = fmaf(nkr, __uint_as_float(r[4 * g + 0]), -nn.x);
                        key[4 * g + 1] = fmaf(nkr, __uint_as_float(r[4 * g + 1]), -nn.y);
                        key[4 * g + 2] = fmaf(nkr, __uint_as_float(r[4 * g + 2]), -nn.z);
                        key[4 * g + 3] = fmaf(nkr, __uint_as_float(r[4 * g + 3]), -nn.w);
                    }
                    float mx;
                    {
                        float a8[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            a8[i] = fmaxf(key[i], key[i + 8]);
                            if constexpr (SUBC == 32) a8[i] = fmaxf(a8[i], fmaxf(key[(i + 16) % SUBC], key[(i + 24) % SUBC]));
                        }
                        mx = fmaxf(fmaxf(fmaxf(a8[0], a8[1]), fmaxf(a8[2], a8[3])), fmaxf(fmaxf(a8[4], a8[5]), fmaxf(a8[6], a8[7])));
                    }
                    if (mx >= thr_e) {             // one branch per SUBC columns; a few dozen times per row and item
                        // Which ones?  Per group of 4 columns one test on the group maximum, and inside a hit group a rolled loop
                        // over its 4 keys (picked with selects: the group index is static), so the queueing code exists SUBC / 4
                        // times, not SUBC.
                        const int jb = j0 + c * SUBC;
#pragma unroll
                        for (int g = 0; g < SUBC / 4; ++g) {
                            const float k0 = key[4 * g], k1 = key[4 * g + 1], k2 = key[4 * g + 2], k3 = key[4 * g + 3];
                            if (fmaxf(fmaxf(k0, k1), fmaxf(k2, k3)) >= thr_e) {
#pragma unroll 1
                                for (int e = 0; e < 4; ++e) {
                                    const float x = e == 0 ? k0 : e == 1 ? k1 : e == 2 ? k2 : k3;
                                    // -inf keys are pad centroids (norm +inf): never candidates (and with no finite best yet,
                                    // thr = -inf, they would pass); a NaN key fails the comparison
                                    if (x >= thr_e && x != __int_as_float(0xff800000)) {
                                        if (qn == QCAP)
                                            kmin = cand_flush(qx, qj, qn, ck, cj, best, thr, dropped, margin, gfl, thr_e, kmin);
#pragma unroll
                                        for (int u = 0; u < QCAP; ++u)
                                            if (qn == u) {
                                                qx[u] = x;
                                                qj[u] = jb + 4 * g + e;
                                            }
                                        ++qn;
                                    }
                                }
                            }
                        }
                    }
                }
                // Phase B -- the MMA warp already owns the stage again: list maintenance for what was queued
                kmin = cand_flush(qx, qj, qn, ck, cj, best, thr, dropped, margin, gfl, thr_e, kmin);
                if (kmin != __int_as_float(0xff800000) && !(kmin <= gfl))
                    atomicMax(sFloor + row, key_map(kmin));   // this list's 8th key raises the row's floor
                // the tile's norms are no longer needed either
                __syncwarp();
                if (lane == 0) mbar_arrive_local(&norm_empty_bar[as]);
            }
            // write this (query, split, column group)'s candidates as v = -key (free slots: +inf, -1) and the largest v-wise
            // nearest thing it had to let go
            {   // a full list whose smallest key is inside the margin may have turned candidates away (see cand_insert)
                const float kmin = fminf(fminf(fminf(ck[0], ck[1]), fminf(ck[2], ck[3])), fminf(fminf(ck[4], ck[5]), fminf(ck[6], ck[7])));
                if (kmin >= best - margin) dropped = fmaxf(dropped, kmin);
            }
            const int64_t ol = (q * nsplit + sp) * NCG + cg;
            float4* ov = reinterpret_cast<float4*>(topv + ol * TOPT);
            int4* oj = reinterpret_cast<int4*>(topj + ol * TOPT);
            ov[0] = make_float4(-ck[0], -ck[1], -ck[2], -ck[3]);
            ov[1] = make_float4(-ck[4], -ck[5], -ck[6], -ck[7]);
            oj[0] = make_int4(cj[0], cj[1], cj[2], cj[3]);
            oj[1] = make_int4(cj[4], cj[5], cj[6], cj[7]);
            topd[ol] = -dropped;
        }
    }

    tc_fence_before();
    if constexpr (CLUSTER > 1) cluster_sync_all(); else __syncthreads();
    if (warp == 2) tmem_dealloc<CLUSTER>(tmem_base, 512);
}

}  // namespace tc

// ------------------------------------------------------------------- operand preparation ----

// The margin between two approximate values v(q, j) = ||c_j||^2 - 2 <q', c_j'> (primes: the low-precision operands):
// with r = ||x - x'|| the EXACT residual norms written by the prep kernels (queries: per row; centroids: the maximum over
// the set) and Cauchy-Schwarz,
//     |<q, c> - <q', c'>|  =  |<q - q', c> + <q', c - c'>|  <=  r_q c_max + (|q| + r_q) r_c                 (bf16, e4m3)
//     3-way bf16 split (h_q h_c + l_q h_c + h_q l_c is what the MMAs sum; rho1 = x - h, rho2 = x - h - l):
//     |<q, c> - sum|       =  |<l_q, rho1_c> + <rho2_q, c> + <h_q, rho2_c>|  <=  (r1_q + r2_q) r1_c + r2_q c_max + (|q| + r1_q) r2_c
// One v is off by twice that, two of them compared by four times.  On top: the accumulation of the products in the tensor
// cores (acc_rel |q'||c'|, host-chosen per operand kind) and abs_rel (|q|^2 + c_max^2) for the fp32 roundings -- the norms
// ||c||^2 ((d/32 + 5) ulps), the fma of the epilogue, the threshold sum, and the gap between a canonical fp32 distance and
// the true one ((d/8 + 5) ulps of a distance <= 2 (|q|^2 + c_max^2), on both candidates).  1.002 covers the fp32 evaluation
// of the residual norms and of this formula.  No ulp-size assumption about the rows is left: rows built to sit on bf16
// rounding midpoints (relative error 2^-8 per element, twice the typical) just get a larger margin.
struct TcMargin {
    float cmax;      // max ||c||
    float rc1, rc2;  // max ||c - c'|| (rc2: second-level residual of the 3-way split)
    float acc_rel, abs_rel;
    int split;       // 1 = 3-way bf16 split
};
__device__ __forceinline__ float tc_margin(const TcMargin& mp, float qn2, float r1, float r2) {
    const float qn = sqrtf(qn2);
    float e;
    if (mp.split == 1) e = (r1 + r2) * mp.rc1 + r2 * mp.cmax + (qn + r1) * mp.rc2;
    else e = r1 * mp.cmax + (qn + r1) * mp.rc1;
    return 4.008f * e + mp.acc_rel * (qn + r1) * (mp.cmax + mp.rc1) + mp.abs_rel * (qn2 + mp.cmax * mp.cmax);
}


// one warp per row: fp32 row -> bf16 operand row(s) of width kdim (= d, or 3d for the split form
// [h | l | h] for queries / [h | h | l] for centroids), the row's squared norm, and the EXACT norms of what the
// operand does not carry: res[r] = ||x - h||, and for the split form res[res_stride + r] = ||x - h - l||
// (h = bf16(x), l = bf16(x - h); both differences are exact in fp32, so only the fp32 sum of squares rounds --
// tc_margin() allows for it).  The re-check's margin is built from these, not from a worst-case ulp assumption:
// bf16 round-to-nearest is off by up to 2^-8 relative per element (8 significant bits), and a row whose elements all
// sit next to rounding midpoints reaches that.
__global__ void tc_prep_rows_kernel(const float* __restrict__ x, int64_t n, int64_t n_pad, int d, int split, int is_centroid,
                                    __nv_bfloat16* __restrict__ out, float* __restrict__ norm2, float pad_norm,
                                    float* __restrict__ res, int64_t res_stride, TcMargin mp, float* __restrict__ margin_out) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= n_pad) return;
    const int kdim = split ? 3 * d : d;
    __nv_bfloat16* o = out + r * kdim;
    if (r >= n) {
        for (int c = lane; c < kdim; c += 32) o[c] = __float2bfloat16(0.f);
        if (lane == 0) {
            if (norm2) norm2[r] = pad_norm;
            if (res) {
                res[r] = 0.f;
                if (split) res[res_stride + r] = 0.f;
            }
            if (margin_out) margin_out[r] = 0.f;
        }
        return;
    }
    const float* xr = x + r * d;
    float s = 0.f, s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < d; c += 32) {
        const float v = xr[c];
        s = fmaf(v, v, s);
        const __nv_bfloat16 h = __float2bfloat16(v);
        o[c] = h;
        const float r1 = v - __bfloat162float(h);
        s1 = fmaf(r1, r1, s1);
        if (split) {
            const __nv_bfloat16 l = __float2bfloat16(r1);
            const float r2 = r1 - __bfloat162float(l);
            s2 = fmaf(r2, r2, s2);
            if (is_centroid) {
                o[d + c] = h;
                o[2 * d + c] = l;
            } else {
                o[d + c] = l;
                o[2 * d + c] = h;
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, off);
        s1 += __shfl_xor_sync(0xffffffffu, s1, off);
        s2 += __shfl_xor_sync(0xffffffffu, s2, off);
    }
    if (lane == 0) {
        if (norm2) norm2[r] = s;
        if (res) {
            res[r] = sqrtf(s1);
            if (split) res[res_stride + r] = sqrtf(s2);
        }
        if (margin_out) margin_out[r] = tc_margin(mp, s, sqrtf(s1), sqrtf(s2));   // queries: the re-check's margin
    }
}

// ---- e4m3 operands (coarse_mode 4) ----
// Power-of-two scale s with 128 <= max|x| * s < 256 (e4m3 tops out at 448; a power of two keeps x * s exact), or 0 when no
// such fp32 scale exists / the row is all zero or not finite (the caller then routes the row to the exact kernel).
__host__ __device__ inline float f8_scale_for(float mx) {
    if (!(mx > 0.f) || !(mx < INFINITY)) return 0.f;
    int e;
    frexpf(mx, &e);   // mx = m * 2^e, 0.5 <= m < 1
    e = 8 - e;
    if (e > 100 || e < -100) return 0.f;
    return ldexpf(1.f, e);
}

// max |x| over an array; bit 31 of flags[1] is raised when a value is NaN / inf
__global__ void tc_maxabs_kernel(const float* __restrict__ x, int64_t n, unsigned int* __restrict__ flags) {
    float m = 0.f;
    bool bad = false;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float a = fabsf(x[i]);
        if (a < INFINITY) m = fmaxf(m, a);
        else bad = true;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    bad = __any_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
        if (m > 0.f) atomicMax(&flags[0], __float_as_uint(m));
        if (bad) atomicOr(&flags[1], 0x80000000u);
    }
}

// one warp per row: fp32 row -> e4m3 row scaled by a power of two (`fixed_scale` for the centroids -- one scale for the
// whole set, so the epilogue needs no per-column factor -- or the row's own scale for a query), the row's squared norm, and
// for queries kfac[r] = -2 / (s_q(r) * other_scale), the factor that turns the scaled accumulator into -2 <q, c>.  A query
// row that cannot be scaled (all zero, NaN / inf, exponent out of range) gets kfac = NaN: every v is NaN, no candidate is
// kept, and the re-check hands the query to the exact kernel.
__global__ void tc_prep_rows_f8_kernel(const float* __restrict__ x, int64_t n, int64_t n_pad, int d, float fixed_scale,
                                       float other_scale, uint8_t* __restrict__ out, float* __restrict__ norm2,
                                       float* __restrict__ kfac, float pad_norm, float* __restrict__ res, TcMargin mp,
                                       float* __restrict__ margin_out) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= n_pad) return;
    uint8_t* o = out + r * d;
    if (r >= n) {
        for (int c = lane; c < d; c += 32) o[c] = 0;
        if (lane == 0) {
            if (norm2) norm2[r] = pad_norm;
            if (kfac) kfac[r] = -2.f;
            if (res) res[r] = 0.f;
            if (margin_out) margin_out[r] = 0.f;
        }
        return;
    }
    const float* xr = x + r * d;
    float s = 0.f, mx = 0.f;
    bool bad = false;
    for (int c = lane; c < d; c += 32) {
        const float v = xr[c];
        s = fmaf(v, v, s);
        const float a = fabsf(v);
        if (a < INFINITY) mx = fmaxf(mx, a);
        else bad = true;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, off);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    bad = __any_sync(0xffffffffu, bad);
    float scale = fixed_scale > 0.f ? fixed_scale : f8_scale_for(mx);
    if (!(scale > 0.f)) {
        bad = true;
        scale = 1.f;
    }
    // exact norm of what the e4m3 operand does not carry, in the row's scaled units: x * scale is exact (power of two),
    // and the difference to its 4-significant-bit rounding is exact in fp32
    float s1 = 0.f;
    for (int c = lane; c < d; c += 32) {
        const float xs = xr[c] * scale;
        const __nv_fp8_storage_t b = __nv_cvt_float_to_fp8(xs, __NV_SATFINITE, __NV_E4M3);
        o[c] = (uint8_t)b;
        const float back = __half2float(__half(__nv_cvt_fp8_to_halfraw(b, __NV_E4M3)));
        const float r1 = xs - back;
        s1 = fmaf(r1, r1, s1);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, off);
    if (lane == 0) {
        if (norm2) norm2[r] = s;
        const float rq = bad ? __int_as_float(0x7fc00000) : sqrtf(s1) / scale;   // division by a power of two: exact
        if (res) res[r] = rq;
        if (margin_out) margin_out[r] = tc_margin(mp, s, rq, 0.f);
        if (kfac) {
            const double p = (double)scale * (double)other_scale;
            float k = (float)(-2.0 / p);
            if (bad || !(fabsf(k) >= 1.17549435e-38f) || !(fabsf(k) < INFINITY)) k = __int_as_float(0x7fc00000);
            kfac[r] = k;
        }
    }
}

__global__ void tc_max_kernel(const float* __restrict__ v, int64_t n, unsigned int* __restrict__ out_bits) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    float m = 0.f;
    if (i < n) m = v[i];
    if (!(m == m)) m = INFINITY;   // a NaN row makes the maximum +inf (fmaxf would drop it): every margin becomes +inf
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out_bits, __float_as_uint(m));  // non-negative floats order as uints
}

// ------------------------------------------------------------------------- exact re-check ----

// One warp per query.  Gathers the nsplit x 8 approximate (v, j) pairs (nsplit = centroid splits x column groups; unsorted,
// free slots hold j = -1), keeps those with v <= vmin + margin(q), evaluates their canonical fp32 distance (8 lanes per
// candidate, 4 candidates per pass) and takes the argmin by (distance, id).  Flags the query for the exact kernel when a
// list dropped something inside the margin (it may be incomplete) or no candidate is valid.
__global__ void tc_recheck_kernel(const float* __restrict__ q, int64_t n, int d, const float* __restrict__ cent, int64_t nlist,
                                  const float* __restrict__ qmargin, const float* __restrict__ topv,
                                  const int* __restrict__ topj, const float* __restrict__ topd, int nsplit, int64_t list_begin,
                                  int32_t* __restrict__ list_out, float* __restrict__ dist_out, int32_t* __restrict__ fb_list,
                                  int32_t* __restrict__ counters /*[0] fallbacks of this call; [8..9], [10..11] u64 running totals*/) {
    const int lane = threadIdx.x & 31;
    const int64_t qi = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (qi >= n) return;
    // |v_approx - v_true| <= margin / 2 for every centroid; the same number the tensor-core epilogue collected with
    const float margin = qmargin[qi];
    const int total = nsplit * tc::TOPT;
    const float* tv = topv + qi * total;
    const int* tj = topj + qi * total;
    float vmin = __int_as_float(0x7f800000);
    for (int e = lane; e < total; e += 32) {
        const int j = tj[e];
        if (j >= 0 && j < nlist) vmin = fminf(vmin, tv[e]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, off));
    const float thr = vmin + margin;
    // no valid candidate at all (e.g. NaN query), or no finite margin (overflowing norms): the exact kernel decides
    bool overflow = !(vmin < __int_as_float(0x7f800000)) || !(margin < __int_as_float(0x7f800000));
    // truncation check: a list (one per split and column group) had to let go of something inside the margin
    for (int s = lane; s < nsplit; s += 32)
        if (topd[qi * nsplit + s] <= thr) overflow = true;
    overflow = __any_sync(0xffffffffu, overflow);
    if (overflow) {
        if (lane == 0) {
            const int slot = atomicAdd(&counters[0], 1);
            fb_list[slot] = (int32_t)qi;
            list_out[qi] = -1;
            if (dist_out) dist_out[qi] = RVCR_FLT_MAX;
            atomicAdd(reinterpret_cast<unsigned long long*>(&counters[8]), 1ull);
        }
        return;
    }
    const int l8 = lane & 7, grp = lane >> 3;
    const float* qrow = q + qi * d;
    float bd = RVCR_FLT_MAX;
    int bj = -1;
    int ncand = 0;
    // 32 kept entries at a time: the ones inside the margin are compacted with a ballot and evaluated four at a time, 8 lanes
    // each (canonical order: lane l8 owns accumulator l8), so a query with 4 candidates costs one pass over the row, not four
    for (int e0 = 0; e0 < total; e0 += 32) {
        const int e = e0 + lane;
        int j = -1;
        float v = 0.f;
        if (e < total) {
            j = tj[e];
            v = tv[e];
        }
        unsigned m = __ballot_sync(0xffffffffu, j >= 0 && j < nlist && v <= thr);
        while (m) {
            const unsigned src = __fns(m, 0, grp + 1);              // the (grp+1)-th candidate of this batch, if any
            const int cj = __shfl_sync(0xffffffffu, j, src & 31);
            const bool take = src != 0xffffffffu;
            float acc = 0.f;
            if (take) {
                const float* crow = cent + (int64_t)cj * d;
#pragma unroll 8
                for (int i = 0; i < d; i += 8) {
                    const float t = __fsub_rn(qrow[i + l8], __ldg(crow + i + l8));
                    acc = __fmaf_rn(t, t, acc);
                }
            }
            acc = reduce8_canonical(acc);
            if (take) {
                ++ncand;
                if (acc < bd || (acc == bd && cj < bj)) {
                    bd = acc;
                    bj = cj;
                }
            }
            // drop the (up to) four candidates just evaluated
#pragma unroll
            for (int k = 0; k < 4; ++k) m &= m - 1;
        }
    }
    // merge the 4 groups (lanes 0, 8, 16, 24 hold distinct candidates) by (distance, id)
#pragma unroll
    for (int off = 8; off < 32; off <<= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, bd, off);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, off);
        const int oc = __shfl_xor_sync(0xffffffffu, ncand, off);
        ncand += oc;
        if (oj >= 0 && (bj < 0 || od < bd || (od == bd && oj < bj))) {
            bd = od;
            bj = oj;
        }
    }
    if (lane == 0) {
        if (bj < 0) {   // every candidate distance was NaN: let the exact kernel decide (-> -1)
            const int slot = atomicAdd(&counters[0], 1);
            fb_list[slot] = (int32_t)qi;
            list_out[qi] = -1;
            if (dist_out) dist_out[qi] = RVCR_FLT_MAX;
            atomicAdd(reinterpret_cast<unsigned long long*>(&counters[8]), 1ull);
        } else {
            list_out[qi] = (int32_t)(bj + list_begin);
            if (dist_out) dist_out[qi] = bd;
            atomicAdd(reinterpret_cast<unsigned long long*>(&counters[10]), (unsigned long long)ncand);
        }
    }
}

// ------------------------------------------------------------------------------ host side ----

namespace tc {
EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
}  // namespace tc


namespace {

// row-major bf16 matrix [rows][kdim] -> tensor map with a [box_rows][64] box, 128B swizzle
int make_tmap(CUtensorMap* m, const void* base, int64_t rows, int kdim, int box_rows) {
    tc::EncodeTiledFn fn = tc::get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return RVCR_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)kdim, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kdim * 2};
    cuuint32_t box[2] = {(cuuint32_t)tc::BLOCK_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for a %lld x %d bf16 matrix", (int)r, (long long)rows, kdim);
        return RVCR_ERR_CUDA;
    }
    return RVCR_OK;
}

// row-major e4m3 matrix [rows][kdim] (one byte per element) -> tensor map with a [box_rows][128] box, 128B swizzle
int make_tmap_u8(CUtensorMap* m, const void* base, int64_t rows, int kdim, int box_rows) {
    tc::EncodeTiledFn fn = tc::get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return RVCR_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)kdim, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kdim};
    cuuint32_t box[2] = {(cuuint32_t)(2 * tc::BLOCK_K), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for a %lld x %d e4m3 matrix", (int)r, (long long)rows, kdim);
        return RVCR_ERR_CUDA;
    }
    return RVCR_OK;
}

inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// the margin parameters of the current centroid operands (split: 0 bf16, 1 3-way bf16 split, 2 e4m3)
TcMargin make_margin(const rvcr_index* h, int split) {
    TcMargin mp;
    const int d = h->d;
    const int kdim = split == 1 ? 3 * d : d;
    mp.cmax = h->cent_norm_max;
    mp.rc1 = split == 2 ? h->cent_f8_res : h->cent_res1;
    mp.rc2 = split == 1 ? h->cent_res2 : 0.f;
    // products of bf16 / e4m3 operands are exact in fp32; their accumulation in TMEM is allowed 2 fp32 ulps of |q'||c'| per
    // product (x 2 for v = |c|^2 - 2 <.,.>, x 2 for two values compared).  e4m3: the same, but never less than 2^-9 --
    // kind::f8f6f4 is not documented to keep full fp32 accumulation, and against a margin of ~0.2 |q||c| the allowance is free
    mp.acc_rel = 1.02f * (float)kdim / 2097152.f;
    if (split == 2 && mp.acc_rel < 1.f / 512.f) mp.acc_rel = 1.f / 512.f;
    // fp32 roundings, in ulps (2^-24) of |q|^2 + c_max^2: canonical-vs-true distance 2 x 2 x (d/8 + 5), norms 2 x (d/32 + 5),
    // epilogue fma + threshold sum 2 x 4
    mp.abs_rel = 1.01f * (float)(4 * (d / 8 + 5) + 2 * (d / 32 + 5) + 8) / 16777216.f;
    mp.split = split;
    return mp;
}

// max over v[0..n) of non-negative floats (NaN counts as +inf), through slot `slot` of ws_misc; synchronises the stream
int device_max(rvcr_index* h, const float* v, int64_t n, int slot, float* out, cudaStream_t s) {
    unsigned int* mx = h->ws_misc.as<unsigned int>() + slot;
    RVCR_CUDA(cudaMemsetAsync(mx, 0, 4, s));
    tc_max_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(v, n, mx);
    RVCR_CHECK_LAUNCH();
    unsigned int bits = 0;
    RVCR_CUDA(cudaMemcpyAsync(&bits, mx, 4, cudaMemcpyDeviceToHost, s));
    RVCR_CUDA(cudaStreamSynchronize(s));
    memcpy(out, &bits, 4);
    count_launch(h);
    return RVCR_OK;
}

// centroid operands are built once per (centroid set, split mode)
int ensure_centroid_operands(rvcr_index* h, int split, cudaStream_t s) {
    if (h->cent_bf16 && h->cent_split == split) return RVCR_OK;
    if (h->cent_bf16) cudaFree(h->cent_bf16);
    if (h->cent_norm) cudaFree(h->cent_norm);
    h->cent_bf16 = nullptr;
    h->cent_norm = nullptr;
    const int kdim = split ? 3 * h->d : h->d;
    const int64_t npad = round_up(h->nlist, tc::BLOCK_N);
    RVCR_CUDA(cudaMalloc(&h->cent_bf16, (size_t)npad * kdim * 2));
    RVCR_CUDA(cudaMalloc(&h->cent_norm, (size_t)npad * 4));
    float* res = nullptr;   // ||c - c'|| per centroid (two planes for the split form); only the maxima are kept
    RVCR_CUDA(cudaMalloc(&res, (size_t)npad * 2 * 4));
    const int64_t thr = npad * 32;
    // pad centroids get ||c||^2 = +inf: v = +inf, never kept
    tc_prep_rows_kernel<<<(unsigned)((thr + 255) / 256), 256, 0, s>>>(h->cent, h->nlist, npad, h->d, split, 1,
                                                                       reinterpret_cast<__nv_bfloat16*>(h->cent_bf16),
                                                                       h->cent_norm, INFINITY, res, npad, TcMargin{}, nullptr);
    int rc = cudaGetLastError() == cudaSuccess ? RVCR_OK : RVCR_ERR_CUDA;
    float m2 = 0.f, r1 = 0.f, r2 = 0.f;
    if (rc == RVCR_OK) rc = device_max(h, h->cent_norm, h->nlist, 12, &m2, s);
    if (rc == RVCR_OK) rc = device_max(h, res, h->nlist, 60, &r1, s);
    if (rc == RVCR_OK && split) rc = device_max(h, res + npad, h->nlist, 61, &r2, s);
    cudaFree(res);
    if (rc != RVCR_OK) {
        set_error("building the tensor-core centroid operands failed: %s", cudaGetErrorString(cudaGetLastError()));
        return rc;
    }
    h->cent_norm_max = sqrtf(m2);
    h->cent_res1 = r1;
    h->cent_res2 = r2;
    h->cent_pad = npad;
    h->cent_kdim = kdim;
    h->cent_split = split;
    count_launch(h);
    return RVCR_OK;
}

// e4m3 centroid operand (coarse_mode 4): one power-of-two scale for the whole set; needs the norms of the bf16 build
int ensure_centroid_operands_f8(rvcr_index* h, cudaStream_t s) {
    RVCR_TRY(ensure_centroid_operands(h, 0, s));   // ||c||^2 (+inf for pad rows) and max ||c||
    if (h->cent_f8) return RVCR_OK;
    unsigned int* fl = h->ws_misc.as<unsigned int>() + 56;   // [56] max |c| bits, [57] non-finite flag
    RVCR_CUDA(cudaMemsetAsync(fl, 0, 8, s));
    tc_maxabs_kernel<<<h->num_sms * 8, 256, 0, s>>>(h->cent, h->nlist * (int64_t)h->d, fl);
    RVCR_CHECK_LAUNCH();
    unsigned int got[2] = {0, 0};
    RVCR_CUDA(cudaMemcpyAsync(got, fl, 8, cudaMemcpyDeviceToHost, s));
    RVCR_CUDA(cudaStreamSynchronize(s));
    float mx;
    memcpy(&mx, &got[0], 4);
    const float sc = f8_scale_for(mx);
    if (got[1] || !(sc > 0.f)) {
        set_error("coarse_mode 4 (e4m3 operands) needs finite, non-zero centroids; use coarse_mode 2");
        return RVCR_ERR_UNSUPPORTED;
    }
    RVCR_CUDA(cudaMalloc(&h->cent_f8, (size_t)h->cent_pad * h->d));
    float* res = nullptr;
    RVCR_CUDA(cudaMalloc(&res, (size_t)h->cent_pad * 4));
    const int64_t thr = h->cent_pad * 32;
    tc_prep_rows_f8_kernel<<<(unsigned)((thr + 255) / 256), 256, 0, s>>>(h->cent, h->nlist, h->cent_pad, h->d, sc, 1.f,
                                                                          reinterpret_cast<uint8_t*>(h->cent_f8), nullptr, nullptr, 0.f, res,
                                                                          TcMargin{}, nullptr);
    int rc = cudaGetLastError() == cudaSuccess ? RVCR_OK : RVCR_ERR_CUDA;
    float r1 = 0.f;
    if (rc == RVCR_OK) rc = device_max(h, res, h->nlist, 62, &r1, s);
    cudaFree(res);
    if (rc != RVCR_OK) {
        set_error("building the e4m3 centroid operand failed: %s", cudaGetErrorString(cudaGetLastError()));
        return rc;
    }
    h->cent_f8_scale = sc;
    h->cent_f8_res = r1;
    count_launch(h, 2);
    return RVCR_OK;
}

template <int CLUSTER, int NRES, int SB, int SA, int P, int F8 = 0>
int launch_tc(rvcr_index* h, const CUtensorMap& ta, const CUtensorMap& tb, int num_mblocks, int num_ntiles, int nsplit,
              int tps, int nkb, float* topv, int* topj, cudaStream_t s, const float* qk = nullptr) {
    float* topd = h->ws_tc7.as<float>();
    const float* qmargin = h->ws_tc6.as<float>();
    using C = tc::Cfg<CLUSTER, NRES, SB, SA, P>;
    static SmemOptIn opt_in;   // one per template instantiation
    RVCR_TRY(opt_in.ensure(tc::coarse_tc_kernel<CLUSTER, NRES, SB, SA, P, F8>, (size_t)C::SMEM_BYTES, h->device));
    const int nitems = num_mblocks * nsplit;
    int nclusters = h->num_sms / CLUSTER;
    if (nclusters > nitems) nclusters = nitems;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(nclusters * CLUSTER));
    cfg.blockDim = dim3(tc::THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CLUSTER;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    RVCR_CUDA(cudaLaunchKernelEx(&cfg, tc::coarse_tc_kernel<CLUSTER, NRES, SB, SA, P, F8>, ta, tb, (const float*)h->cent_norm,
                                 num_mblocks, num_ntiles, nsplit, tps, nkb, topv, topj, topd, qk, qmargin));
    return RVCR_OK;
}

}  // namespace

bool coarse_tc_available() { return true; }

int coarse_tc_invalidate(rvcr_index* h) {
    if (h->cent_bf16) cudaFree(h->cent_bf16);
    if (h->cent_norm) cudaFree(h->cent_norm);
    if (h->cent_f8) cudaFree(h->cent_f8);
    h->cent_bf16 = nullptr;
    h->cent_norm = nullptr;
    h->cent_f8 = nullptr;
    h->cent_f8_scale = 0.f;
    h->f8_decision = 0;
    h->f8_last_n = 0;
    h->cent_split = -1;
    return RVCR_OK;
}

// Approximate top-8 nearest centroids per (query, split) on the tensor cores.
// Outputs (device, owned by the handle scratch): topv/topj [n_pad][nsplit][8]; returns nsplit.
int coarse_tc_top8(rvcr_index* h, const float* q, int64_t n, int split, int cluster, float** topv_out, int** topj_out,
                   float** topd_out, int* nsplit_out, float** qnorm_out, float** qmargin_out, cudaStream_t s) {
    // split: 0 = bf16, 1 = bf16 3-way split along K, 2 = e4m3 operands (coarse_mode 4)
    const bool f8 = split == 2;
    const int kelems = f8 ? 2 * tc::BLOCK_K : tc::BLOCK_K;   // elements per 128-byte k-block
    if (h->d % kelems != 0) {
        set_error("d=%d unsupported by the tcgen05 path (must be a multiple of %d)", h->d, kelems);
        return RVCR_ERR_UNSUPPORTED;
    }
    const int kdim = split == 1 ? 3 * h->d : h->d;
    const int mblk = tc::BLOCK_M * cluster;
    const int64_t n_pad = round_up(n, mblk);
    {
        PhaseTimer pt(h, RVCR_PHASE_COARSE_PREP, s);
        const int64_t thr = n_pad * 32;
        if (f8) {
            if (h->d > 1024) {
                set_error("coarse_mode 4: d=%d above 1024 (the e4m3 error bound is derived for d <= 1024)", h->d);
                return RVCR_ERR_UNSUPPORTED;
            }
            RVCR_TRY(ensure_centroid_operands_f8(h, s));
            RVCR_TRY(h->ws_tc0.reserve((size_t)n_pad * kdim));   // e4m3 queries, each row with its own power-of-two scale
            RVCR_TRY(h->ws_tc1.reserve((size_t)n_pad * 4));      // ||q||^2
            RVCR_TRY(h->ws_tc4.reserve((size_t)n_pad * 4));      // -2 / (s_q s_c) per query
            RVCR_TRY(h->ws_tc5.reserve((size_t)n_pad * 2 * 4));  // ||q - q'||
            RVCR_TRY(h->ws_tc6.reserve((size_t)n_pad * 4));      // margin(q)
            tc_prep_rows_f8_kernel<<<(unsigned)((thr + 255) / 256), 256, 0, s>>>(q, n, n_pad, h->d, 0.f, h->cent_f8_scale,
                                                                                  h->ws_tc0.as<uint8_t>(), h->ws_tc1.as<float>(),
                                                                                  h->ws_tc4.as<float>(), 0.f, h->ws_tc5.as<float>(),
                                                                                  make_margin(h, 2), h->ws_tc6.as<float>());
        } else {
        RVCR_TRY(ensure_centroid_operands(h, split, s));
        RVCR_TRY(h->ws_tc0.reserve((size_t)n_pad * kdim * 2));   // bf16 queries
        RVCR_TRY(h->ws_tc1.reserve((size_t)n_pad * 4));          // ||q||^2
        RVCR_TRY(h->ws_tc5.reserve((size_t)n_pad * 2 * 4));      // ||q - q'|| (second plane: 3-way split)
        RVCR_TRY(h->ws_tc6.reserve((size_t)n_pad * 4));          // margin(q)
        tc_prep_rows_kernel<<<(unsigned)((thr + 255) / 256), 256, 0, s>>>(q, n, n_pad, h->d, split, 0,
                                                                           h->ws_tc0.as<__nv_bfloat16>(), h->ws_tc1.as<float>(), 0.f,
                                                                           h->ws_tc5.as<float>(), n_pad, make_margin(h, split),
                                                                           h->ws_tc6.as<float>());
        }
        RVCR_CHECK_LAUNCH();
        count_launch(h);
    }
    const int num_mblocks = (int)(n_pad / mblk);
    const int num_ntiles = (int)(h->cent_pad / tc::BLOCK_N);
    const int nclusters = h->num_sms / cluster;
    // Work items = m-blocks x centroid splits, dealt round-robin to the CTA pairs.  Few m-blocks: split the centroid
    // tiles so that every pair has ~3 items.  (Splitting further to fill the last round of a larger batch was measured
    // and lost: 131072 queries take 14.5 ms unsplit but 22.6 ms in 13 splits -- r01 coarse probe.)
    int nsplit = 1;
    if (num_mblocks < 4 * nclusters) {
        nsplit = (3 * nclusters + num_mblocks - 1) / num_mblocks;
        if (nsplit > num_ntiles) nsplit = num_ntiles;
        if (nsplit < 1) nsplit = 1;
    }
    if (const char* e = getenv("RVCR_TC_NSPLIT")) {   // tuning knob
        const int v = atoi(e);
        if (v >= 1 && v <= num_ntiles) nsplit = v;
    }
    int tps = (num_ntiles + nsplit - 1) / nsplit;
    nsplit = (num_ntiles + tps - 1) / tps;
    RVCR_TRY(h->ws_tc2.reserve((size_t)n_pad * nsplit * tc::NCG * tc::TOPT * 4));   // one top-8 per (query, split, column group)
    RVCR_TRY(h->ws_tc3.reserve((size_t)n_pad * nsplit * tc::NCG * tc::TOPT * 4));
    RVCR_TRY(h->ws_tc7.reserve((size_t)n_pad * nsplit * tc::NCG * 4));               // nearest value each list let go of
    CUtensorMap ta, tb;
    if (f8) {
        RVCR_TRY(make_tmap_u8(&ta, h->ws_tc0.p, n_pad, kdim, tc::BLOCK_M));
        RVCR_TRY(make_tmap_u8(&tb, h->cent_f8, h->cent_pad, kdim, tc::BLOCK_N / cluster));
    } else {
    RVCR_TRY(make_tmap(&ta, h->ws_tc0.p, n_pad, kdim, tc::BLOCK_M));
    RVCR_TRY(make_tmap(&tb, h->cent_bf16, h->cent_pad, kdim, tc::BLOCK_N / cluster));
    }
    {
        PhaseTimer pt(h, RVCR_PHASE_COARSE_MAIN, s);
        float* tv = h->ws_tc2.as<float>();
        int* tj = h->ws_tc3.as<int>();
        const int nkb = kdim / kelems;
        // option "tc_resident": A k-blocks pinned in shared memory (0 = stream both operands, the r01 v1 kernel)
        const int64_t res = h->opt.tc_resident;
        if (f8) {   // d / 128 k-blocks (6 at d = 768)
            const float* qk = h->ws_tc4.as<float>();
            // cluster 2: the 6 k-blocks of a 768-d query operand resident (d = 1024: two more streamed), 7 centroid stages in flight
            if (cluster == 2) RVCR_TRY((launch_tc<2, 6, 7, 1, 64, 1>(h, ta, tb, num_mblocks, num_ntiles, nsplit, tps, nkb, tv, tj, s, qk)));
            else RVCR_TRY((launch_tc<1, 8, 2, 2, 3, 1>(h, ta, tb, num_mblocks, num_ntiles, nsplit, tps, nkb, tv, tj, s, qk)));
        } else if (cluster == 2) {
            // tc_resident 0 streams both operands; every other value pins 9 of the 12 query k-blocks (the variants with 6 / 8 /
            // 10 resident blocks measured within 2 % of it in round 1 and were removed)
            if (res == 0) RVCR_TRY((launch_tc<2, 0, 6, 6, 1>(h, ta, tb, num_mblocks, num_ntiles, nsplit, tps, nkb, tv, tj, s)));
            else RVCR_TRY((launch_tc<2, 9, 4, 1, 4>(h, ta, tb, num_mblocks, num_ntiles, nsplit, tps, nkb, tv, tj, s)));
        } else {
            if (res == 0) RVCR_TRY((launch_tc<1, 0, 4, 4, 1>(h, ta, tb, num_mblocks, num_ntiles, nsplit, tps, nkb, tv, tj, s)));
            else RVCR_TRY((launch_tc<1, 8, 2, 2, 3>(h, ta, tb, num_mblocks, num_ntiles, nsplit, tps, nkb, tv, tj, s)));
        }
        count_launch(h);
    }
    *topv_out = h->ws_tc2.as<float>();
    *topj_out = h->ws_tc3.as<int>();
    *nsplit_out = nsplit * tc::NCG;   // top-8 lists per query
    *qnorm_out = h->ws_tc1.as<float>();
    *qmargin_out = h->ws_tc6.as<float>();
    *topd_out = h->ws_tc7.as<float>();
    return RVCR_OK;
}

int coarse_exact_selected(rvcr_index* h, const float* q, int64_t n, const int32_t* qsel, const int32_t* nsel_dev,
                          int32_t* list_out, float* dist_out, cudaStream_t s);

static int coarse_tc_split(rvcr_index* h, const float* q, int64_t n, int32_t* list_out, float* dist_out, int split, cudaStream_t s);

int coarse_tc(rvcr_index* h, const float* q, int64_t n, int32_t* list_out, float* dist_out, cudaStream_t s) {
    int split = h->opt.coarse_mode == 3 ? 1 : (h->opt.coarse_mode == 4 ? 2 : 0);
    // option coarse_f8_auto (auto mode only): decide per centroid set, on the caller's own queries, whether the e4m3
    // operands pay off -- they do when few centroids lie within the (35x wider) margin of the best one
    if (h->opt.coarse_mode == 0 && h->opt.coarse_f8_auto && h->d % 128 == 0 && h->d <= 1024) {
        constexpr int64_t kProbe = 8192;
        if (h->f8_decision == 0 && n >= 8 * kProbe) {
            int rc = coarse_tc_split(h, q, kProbe, list_out, dist_out, 2, s);   // results are overwritten by the full call below
            if (rc == RVCR_ERR_UNSUPPORTED) {
                h->f8_decision = -1;                                            // e.g. non-finite centroids
            } else {
                RVCR_TRY(rc);
                int32_t fb = 0;
                RVCR_CUDA(cudaMemcpyAsync(&fb, h->ws_misc.as<int32_t>() + 16, 4, cudaMemcpyDeviceToHost, s));
                RVCR_CUDA(cudaStreamSynchronize(s));
                h->f8_decision = fb <= kProbe / 200 ? 1 : -1;
            }
        }
        // sticky downgrade: if the previous e4m3 call sent more than 2 % of its queries to the exact kernel (the probe's
        // prefix was not representative), this centroid set goes back to bf16.  The count was copied asynchronously.
        if (h->f8_decision > 0 && h->f8_last_n > 0 && h->f8_fb_event && cudaEventQuery(h->f8_fb_event) == cudaSuccess) {
            if ((int64_t)*h->f8_fb_host * 50 > h->f8_last_n) h->f8_decision = -1;
            h->f8_last_n = 0;
        }
        cudaGetLastError();   // cudaEventQuery's cudaErrorNotReady is not an error
        if (h->f8_decision > 0) split = 2;
    }
    RVCR_TRY(coarse_tc_split(h, q, n, list_out, dist_out, split, s));
    if (split == 2 && h->opt.coarse_mode == 0 && n >= 4096) {
        if (!h->f8_fb_host) RVCR_CUDA(cudaMallocHost(&h->f8_fb_host, sizeof(int32_t)));
        if (!h->f8_fb_event) RVCR_CUDA(cudaEventCreateWithFlags(&h->f8_fb_event, cudaEventDisableTiming));
        RVCR_CUDA(cudaMemcpyAsync(h->f8_fb_host, h->ws_misc.as<int32_t>() + 16, 4, cudaMemcpyDeviceToHost, s));
        RVCR_CUDA(cudaEventRecord(h->f8_fb_event, s));
        h->f8_last_n = n;
    }
    return RVCR_OK;
}

static int coarse_tc_split(rvcr_index* h, const float* q, int64_t n, int32_t* list_out, float* dist_out, int split, cudaStream_t s) {
    const int cluster = h->opt.tc_cluster == 1 ? 1 : 2;
    float *topv, *topd, *qnorm, *qmargin;
    int* topj;
    int nsplit;
    RVCR_TRY(coarse_tc_top8(h, q, n, split, cluster, &topv, &topj, &topd, &nsplit, &qnorm, &qmargin, s));
    PhaseTimer pt(h, RVCR_PHASE_COARSE_RECHECK, s);
    // error margin between two approximate values: tc_margin(), from the exact residual norms of both operands (written per
    // query by the prep kernel, used by the tensor-core epilogue to collect and here to select)
    RVCR_TRY(h->ws_stage.reserve((size_t)n * 4 + 64));
    int32_t* fb_list = h->ws_stage.as<int32_t>();
    int32_t* counters = h->ws_misc.as<int32_t>() + 16;   // [16] fallbacks of this call; [24..27] running totals
    RVCR_CUDA(cudaMemsetAsync(counters, 0, 4, s));
    const int64_t thr = n * 32;
    tc_recheck_kernel<<<(unsigned)((thr + 255) / 256), 256, 0, s>>>(q, n, h->d, h->cent, h->nlist, qmargin, topv, topj, topd,
                                                                     nsplit, 0, list_out, dist_out, fb_list, counters);
    RVCR_CHECK_LAUNCH();
    count_launch(h);
    RVCR_TRY(coarse_exact_selected(h, q, n, fb_list, counters, list_out, dist_out, s));
    h->stats.coarse_tc_queries += n;
    h->tc_counters_dirty = true;
    return RVCR_OK;
}

}  // namespace rvcr
