// common.cuh -- shared definitions for the librvcr CUDA sources (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rvcr.h"

#define RVCR_FLT_MAX 3.402823466e+38f
#define PEER_INBOX_COUNT_OFF RVCR_PEER_INBOX_COUNT_OFF
#define PEER_INBOX_FLAGS_OFF RVCR_PEER_INBOX_FLAGS_OFF
#define PEER_INBOX_HEADER RVCR_PEER_INBOX_HEADER
#define RVCR_K 8  // k of index.search(npy, k=8), rvc/infer/pipeline.py:499

namespace rvcr {

// ---------------------------------------------------------------- error plumbing (host) ----
void set_error(const char* fmt, ...);
const char* get_error();

#define RVCR_CUDA(expr)                                                                      \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            rvcr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                            __LINE__);                                                       \
            return RVCR_ERR_CUDA;                                                            \
        }                                                                                    \
    } while (0)

#define RVCR_TRY(expr)            \
    do {                          \
        int _r = (expr);          \
        if (_r != RVCR_OK) return _r; \
    } while (0)

#define RVCR_CHECK_LAUNCH()                                                                   \
    do {                                                                                      \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) {                                                              \
            rvcr::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, \
                            __LINE__);                                                        \
            return RVCR_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the CURRENT device only, and one process may hold handles on
// several devices (a server with one index per GPU): remember per device the largest size a kernel has been opted into.
// One static instance per kernel at its launch site; concurrent callers at worst repeat the (idempotent) call.
constexpr int RVCR_MAX_DEVICES = 64;
struct SmemOptIn {
    std::atomic<size_t> bytes[RVCR_MAX_DEVICES];
    template <class Kernel>
    int ensure(Kernel kernel, size_t smem, int device) {
        if (device < 0 || device >= RVCR_MAX_DEVICES) {
            set_error("device ordinal %d outside [0, %d)", device, RVCR_MAX_DEVICES);
            return RVCR_ERR_INVALID;
        }
        if (smem > bytes[device].load(std::memory_order_relaxed)) {
            RVCR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            bytes[device].store(smem, std::memory_order_relaxed);
        }
        return RVCR_OK;
    }
};

// Device buffer grown on demand; never shrinks.  Growth synchronises the device (rare).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <class T>
    T* as() const {
        return reinterpret_cast<T*>(p);
    }
};

struct Options {
    int64_t coarse_mode = 0;  // 0 auto, 1 exact FFMA, 2 tcgen05 bf16, 3 tcgen05 3-way split, 4 tcgen05 e4m3 (experimental,
                              // explicit only: wide margin, pays off when few centroids are near the best one)
    int64_t nan_on_zero_score = 1;
    int64_t profile = 0;      // record CUDA events around each phase's kernels
    int64_t scan_impl = 0;    // 0 auto (default): 5 for batches of >= scan_tc_min_queries, else 4; 5 tcgen05 tf32 + exact
                              // re-check; 4 exact FFMA scan (packed f32x2, TQx8 tile; 1-3 select it too: the round-1
                              // kernel generations they named were removed)
    int64_t tc_cluster = 2;   // tcgen05 path: 1 = one CTA per SM, 2 = CTA pairs (cta_group::2)
    int64_t fused_blend = 1;  // scan_impl 5: re-check kernel also does the gather + weights + blend
    // Size rules of the auto modes, set from measurements (profiles/r01_small_dispatch.json, B200).  The tensor-core scan has
    // fixed costs (TMEM allocation, a re-check launch, whole 128-row TMA tiles even for 39-row lists) that the exact FFMA
    // scan does not: at C1 (500 frames) exact scan + blend take 75 us against 93 us, at C2 (3 000 frames) 175 us against
    // 117 us.  The tensor-core coarse step wins already at C1 (94 us against 116 us for 1 GFLOP).  All give identical bits.
    int64_t coarse_tc_min_flops = 64000000;     // coarse_mode 0: tcgen05 path from 64 MFLOP per call
    int64_t scan_tc_min_queries = 1024;         // scan_impl 0: tcgen05 scan from 1 Ki queries per call
    int64_t kmeans_niter = 10; // index.train: Lloyd iterations (faiss ClusteringParameters::niter of the Level1Quantizer)
    int64_t coarse_f8_auto = 1; // coarse_mode 0 only: on the first call of >= 65 536 queries per centroid set, run the first 8 192 of
                              // them with e4m3 operands (coarse_mode 4) and keep e4m3 for this centroid set if at most 0.5 % of the
                              // probe fell back to the exact kernel, else stay on bf16 (and go back to bf16 if a later call sends
                              // more than 2 % of its queries to the exact kernel).  Results are identical either way: both feed the
                              // same exact re-check with a margin built from the exact residual norms of the operands.  On since
                              // round 2 (validated on a B200: tests/test_gpu_coarse_f8.py; C5 coarse step 102 -> 74 ms).
    int64_t tc_resident = 9;  // coarse tcgen05 kernel: query-operand k-blocks pinned in shared memory per work item (0, 6, 8, 9, 10)
};

// The step AFTER the path, fused into the blend epilogue (rvc/infer/pipeline.py:463-481): 2x nearest
// upsample along time (F.interpolate(scale_factor=2): out[t] = in[t / 2]) and the pitch-protect mix against the
// un-blended features, feats * pf + feats0 * (1 - pf).
struct PostArgs {
    float* out2 = nullptr;          // (2n, d) destination rows; nullptr = no post step (blend in place as before)
    const float* pitchf = nullptr;  // 2n per-frame pitch values, or nullptr = upsample only (protect >= 0.5 / no pitch guidance)
    float protect = 0.f;
};

// Multi-GPU combine fused into the blend epilogue: the owner shard writes each blended row straight into the HOME rank's
// feature buffer over NVLink (peer-mapped memory, cudaIpc), instead of a local write + NCCL all-to-all + scatter.
// Row qi of the local batch goes to bases[rank[qi]] + row[qi] * d.
struct PushArgs {
    float* const* bases = nullptr;   // device array of G peer-mapped base pointers (own buffer included); nullptr = off
    const int32_t* rank = nullptr;   // per local row: home rank
    const int32_t* row = nullptr;    // per local row: row index inside the home rank's buffer
    int direct = 0;                  // 1: the QUERY rows are read from there too (no pull pass: the scan kernels load the
                                     // rows of their work item straight out of the home ranks' buffers over NVLink)
};

}  // namespace rvcr

// The handle behind the C ABI.
struct rvcr_index {
    int device = 0;
    int d = 0;
    int metric = 1;
    bool trained = false;
    int64_t nlist = 0;         // lists (== centroids) of the index; the quantizer is always complete
    int64_t own_lo = 0;        // a shard stores only lists [own_lo, own_hi); the others are empty here
    int64_t own_hi = 0;
    int64_t ntotal = 0;        // faiss ntotal of the whole index
    int64_t nstored = 0;       // rows in `vecs`
    int64_t nprobe = 1;
    int num_sms = 148;

    // quantizer
    float* cent = nullptr;             // nlist x d fp32
    // tcgen05 operands (built lazily by coarse_tc.cu)
    void* cent_bf16 = nullptr;         // nlist_pad x kdim bf16, K-major
    float* cent_norm = nullptr;        // nlist_pad fp32 ||c||^2 (pad rows = +inf)
    float cent_norm_max = 0.f;         // max ||c||
    float cent_res1 = 0.f;             // max ||c - bf16(c)||                     (exact residual norms: the re-check's
    float cent_res2 = 0.f;             // max ||c - bf16(c) - bf16(c - bf16(c))||  margin is built from these)
    float cent_f8_res = 0.f;           // max ||c - e4m3(c s_c) / s_c||
    int64_t cent_pad = 0;
    int cent_kdim = 0;                 // d (mode 2) or 3d (mode 3)
    int cent_split = 0;
    void* cent_f8 = nullptr;           // nlist_pad x d e4m3 (coarse_mode 4), scaled by cent_f8_scale (a power of two)
    float cent_f8_scale = 0.f;
    int f8_decision = 0;               // coarse_f8_auto: 0 = not probed yet, +1 = e4m3 operands, -1 = bf16 (per centroid set)
    int32_t* f8_fb_host = nullptr;     // pinned: fallbacks of the last e4m3 call in auto mode (copied asynchronously) ...
    cudaEvent_t f8_fb_event = nullptr; // ... valid once this event has completed; checked at the start of the next call
    int64_t f8_last_n = 0;             // queries of that call (0 = nothing pending)

    // list-ordered storage
    float* vecs = nullptr;             // nstored x d fp32
    int64_t* ids = nullptr;            // nstored
    int64_t* offsets = nullptr;        // nlist + 1 (device)
    std::vector<int64_t> offsets_host; // nlist + 1
    int64_t last_row_pos = -1;         // position of id ntotal-1 (numpy's big_npy[-1])
    size_t vecs_cap = 0;               // rows allocated

    // staged adds (id order), folded into the list layout by finalize_pending()
    float* stage_x = nullptr;          // stage_cap x d
    int32_t* stage_assign = nullptr;   // stage_cap
    int64_t stage_cap = 0;
    int64_t npending = 0;

    // per-call scratch
    // tensor-core scan (scan_tc.cu): ||y||^2 per stored row, TMA descriptors of the list storage, candidate lists
    rvcr::DevBuf row_norm, ws_cand, ws_candcnt, ws_fb, ws_cstage;
    bool row_norm_valid = false;
    float row_norm_max = 0.f;
    alignas(64) unsigned char scan_tmaps[4][128];
    bool scan_tmap_valid = false;
    const void* scan_tmap_base = nullptr;
    int64_t scan_tmap_rows = -1;
    bool scan_fb_dirty = false;

    rvcr::DevBuf ws_assign, ws_keys, ws_vals, ws_keys2, ws_vals2, ws_cub, ws_items, ws_flags, ws_misc,
        ws_D, ws_P, ws_I, ws_best, ws_q, ws_tc0, ws_tc1, ws_tc2, ws_tc3, ws_tc4, ws_tc5, ws_tc6, ws_tc7, ws_stage;
    rvcr::DevBuf ws_mp[6];             // nprobe > 1 (multiprobe.cu): probe lists, one column, per-probe D / P / I, distance matrix
    cudaStream_t copy_stream = nullptr;          // host-buffer entry points: H2D chunks overlap the coarse step
    cudaStream_t out_stream = nullptr;           // ... and blended rows stream out while later lists are scanned
    std::vector<cudaEvent_t> copy_events;
    cudaEvent_t last_done = nullptr;
    cudaStream_t last_stream = nullptr;
    bool have_last = false;

    // optional per-phase CUDA-event timing (option "profile"), read back by rvcr_get_phase_ms
    cudaEvent_t pev[RVCR_NUM_PHASES][2] = {};
    bool pev_used[RVCR_NUM_PHASES] = {};

    // set by search_device around the scan when the caller wants blended rows and option "fused_blend" is on; a scan
    // implementation that also performed the tail (scan_tc.cu) sets `done`, otherwise gather_blend runs afterwards
    struct FusedTail {
        bool active = false, done = false;
        float* out = nullptr;
        float r32 = 0.f, omr32 = 0.f;
        rvcr::PostArgs post;   // rvcr_retrieve_blend_upsample: rows go to post.out2 (2n x d), feats stay untouched
        rvcr::PushArgs push;   // rvcr_retrieve_blend_assigned_push: rows go to their home rank's buffer (peer memory)
    } fuse;

    bool tc_counters_dirty = false;
    rvcr::Options opt;
    rvcr_stats stats{};
    std::mutex mu;
};

namespace rvcr {

// ------------------------------------------------------------------ kernel entry points ----
// (host wrappers; all enqueue on `stream`, return RVCR_* codes)

// coarse_exact.cu : brute-force canonical nearest centroid (FFMA pipe)
int coarse_exact(rvcr_index* h, const float* q, int64_t n, const int32_t* qsel /*nullable*/,
                 int64_t nsel, int32_t* list_out, float* dist_out, cudaStream_t s);
// coarse_tc.cu : tcgen05 bf16 approximate top-8 + exact re-check
int coarse_tc(rvcr_index* h, const float* q, int64_t n, int32_t* list_out, float* dist_out,
              cudaStream_t s);
bool coarse_tc_available();
int coarse_tc_top8(rvcr_index* h, const float* q, int64_t n, int split, int cluster, float** topv_out, int** topj_out,
                   float** topd_out, int* nsplit_out, float** qnorm_out, float** qmargin_out, cudaStream_t s);
int coarse_tc_invalidate(rvcr_index* h);
// scan_topk.cu
int group_queries(rvcr_index* h, const int32_t* assign, int64_t n, bool only_local, cudaStream_t s);
int scan_topk(rvcr_index* h, const float* q, int64_t n, float* D, int32_t* P, int64_t* I,
              cudaStream_t s);
int scan_topk_range(rvcr_index* h, const float* q, int64_t n, float* D, int32_t* P, int64_t* I, int item_begin, int item_end,
                    int ticket_slot, bool fill, cudaStream_t s);
int fill_pad(float* D, int32_t* P, int64_t* I, int64_t nk, cudaStream_t s);
// scan_tc.cu : tcgen05 tf32 list scan + exact re-check
bool scan_tc_supported(const rvcr_index* h);
int scan_tc_invalidate(rvcr_index* h);
int scan_tc_range(rvcr_index* h, const float* q, int64_t n, float* D, int32_t* P, int64_t* I, int item_begin, int item_end,
                  int ticket_slot, bool fill, cudaStream_t s);
int gather_blend_range(rvcr_index* h, const float* feats, int64_t n, const float* D, const int32_t* P, double index_rate,
                       int item_begin, int item_end, float* out, int max_ctas, cudaStream_t s);
// multiprobe.cu : index.search / retrieve + blend when the index says nprobe > 1
int search_multiprobe(rvcr_index* h, float* feats_or_q, int64_t n, bool blend, double index_rate, float* D_dev, int64_t* I_dev,
                      cudaStream_t s);
// blend.cu
int gather_blend(rvcr_index* h, float* feats, int64_t n, const float* D, const int32_t* P,
                 const int32_t* assign /*nullable: restrict to owned*/, double index_rate,
                 cudaStream_t s);
// peer.cu : peer-memory (cudaIpc over NVLink) plumbing of the sharded path
int peer_pull_rows(const float* const* bases, const int32_t* src_rank, const int32_t* src_row, int64_t m, int d, float* out,
                   cudaStream_t s, const int32_t* m_dev = nullptr);
// inbox of the device-driven sharded step (layout: RVCR_PEER_INBOX_* in rvcr.h)
int peer_dispatch(const int32_t* assign, int64_t n, int64_t row0, int my_rank, int G, const int32_t* bounds,
                  unsigned char* const* inbox_bases, int64_t cap, cudaStream_t s);
int peer_barrier(unsigned char* const* inbox_bases, int my_rank, int G, int epoch, cudaStream_t s);
int peer_inbox_seal(int32_t* assign, const int32_t* count, int64_t cap, cudaStream_t s);
// final_proj.cu : the v1 query producer, out = x W^T + b (pipeline.py:451-453)
int final_proj(const float* x, int64_t n, int din, const float* w, const float* b, int dout, float* out, cudaStream_t s);
// post.cu : standalone 2x upsample + protect mix (pipeline.py:463-481) for callers that skipped retrieval
int upsample_protect(const float* feats, const float* feats0, const float* pitchf, int64_t n, int d, double protect, float* out,
                     cudaStream_t s);
int merge_coarse(const float* dist_all, const int32_t* list_all, int G, int64_t n, int32_t* out,
                 cudaStream_t s);
int merge_topk(const float* D_all, const int64_t* I_all, int G, int64_t n, int k, float* D, int64_t* I,
               cudaStream_t s);
// kmeans.cu : index.train (faiss Clustering::train), assignment on the coarse kernels above
int kmeans_train(rvcr_index* h, const float* x_host, int64_t n, cudaStream_t s);
void kmeans_debug_rand_perm(int64_t n, int64_t seed, int64_t* out);
int kmeans_debug_split_clusters(int d, int64_t k, int64_t n, float* hassign, float* centroids);
// lists.cu
int finalize_pending(rvcr_index* h, cudaStream_t s);
int reconstruct(rvcr_index* h, int64_t i0, int64_t ni, float* out_dev, cudaStream_t s);
int refresh_derived(rvcr_index* h, cudaStream_t s);

inline void count_launch(rvcr_index* h, int n = 1) {
    if (h) h->stats.kernels_launched += n;
}

// Phase timing: events on the launching stream, bracketing exactly the kernels of one phase.
struct PhaseTimer {
    rvcr_index* h;
    int id;
    cudaStream_t s;
    PhaseTimer(rvcr_index* h_, int id_, cudaStream_t s_) : h(h_), id(id_), s(s_) {
        if (!h->opt.profile) return;
        for (int k = 0; k < 2; ++k)
            if (!h->pev[id][k]) cudaEventCreate(&h->pev[id][k]);
        cudaEventRecord(h->pev[id][0], s);
    }
    ~PhaseTimer() {
        if (!h->opt.profile) return;
        cudaEventRecord(h->pev[id][1], s);
        h->pev_used[id] = true;
    }
};

}  // namespace rvcr

// -------------------------------------------------------------------------- device helpers --
#ifdef __CUDACC__
namespace rvcr {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// pipeline.py:476-478: pitchff = pitchf.clone(); pitchff[pitchf > 0] = 1; pitchff[pitchf < 1] = protect
__device__ __forceinline__ float protect_factor(float pitchf, float protect) {
    float pf = pitchf;
    if (pitchf > 0.f) pf = 1.f;
    if (pitchf < 1.f) pf = protect;
    return pf;
}
// pipeline.py:479-481: feats * pf + feats0 * (1 - pf), three separately rounded fp32 ops (no FMA) like the torch kernels
__device__ __forceinline__ float4 protect_mix(float4 o, float4 f0, float pf) {
    const float om = __fsub_rn(1.0f, pf);
    float4 r;
    r.x = __fadd_rn(__fmul_rn(o.x, pf), __fmul_rn(f0.x, om));
    r.y = __fadd_rn(__fmul_rn(o.y, pf), __fmul_rn(f0.y, om));
    r.z = __fadd_rn(__fmul_rn(o.z, pf), __fmul_rn(f0.z, om));
    r.w = __fadd_rn(__fmul_rn(o.w, pf), __fmul_rn(f0.w, om));
    return r;
}
// Store float4 column c of the blended row `o` of query qi to rows 2qi and 2qi+1 of post.out2 (pipeline.py:463-465),
// mixed with the un-blended row f0 when a pitch track is given.  pitchf loads are warp-uniform (L1 broadcast).
__device__ __forceinline__ void post_store(const PostArgs& pa, int64_t qi, int d, int c, float4 o, float4 f0) {
    float4* r0 = reinterpret_cast<float4*>(pa.out2 + (2 * qi) * d) + c;
    float4* r1 = reinterpret_cast<float4*>(pa.out2 + (2 * qi + 1) * d) + c;
    if (pa.pitchf) {
        *r0 = protect_mix(o, f0, protect_factor(__ldg(pa.pitchf + 2 * qi), pa.protect));
        *r1 = protect_mix(o, f0, protect_factor(__ldg(pa.pitchf + 2 * qi + 1), pa.protect));
    } else {
        *r0 = o;
        *r1 = o;
    }
}

// where the query row of local index qi is read from: the local batch, or (PushArgs::direct) its home rank's buffer
__device__ __forceinline__ const float* query_row(const float* q, const PushArgs& pa, int64_t qi, int d) {
    return (pa.bases && pa.direct) ? pa.bases[pa.rank[qi]] + (int64_t)pa.row[qi] * d : q + qi * d;
}
// destination of the blended row of local query qi when the epilogue pushes to peer memory
__device__ __forceinline__ float4* push_row_ptr(const PushArgs& pa, int64_t qi, int d) {
    return reinterpret_cast<float4*>(pa.bases[pa.rank[qi]] + (int64_t)pa.row[qi] * d);
}

// (dist, pos) strict total order used everywhere a top-k decision is made: distance ascending,
// then list position ascending (== the entry faiss scans first).
__device__ __forceinline__ bool cand_less(float da, int pa, float db, int pb) {
    return (da < db) || (da == db && pa < pb);
}

// canonical 8-lane reduction of faiss fvec_L2sqr: lanes l (0..7) hold acc[l]; afterwards every
// lane holds ((a0+a4)+(a1+a5))+((a2+a6)+(a3+a7)).  fp32 add is commutative, so the xor
// butterfly gives bit-identical results on all 8 lanes.
__device__ __forceinline__ float reduce8_canonical(float a) {
    a = __fadd_rn(a, __shfl_xor_sync(0xffffffffu, a, 4));
    a = __fadd_rn(a, __shfl_xor_sync(0xffffffffu, a, 1));
    a = __fadd_rn(a, __shfl_xor_sync(0xffffffffu, a, 2));
    return a;
}

// warp-level insertion of the candidates (cd, cpos) of the 32 lanes into the sorted top-8 held in
// lanes 0..7 (td, tp); same code as v1's inner loop
__device__ __forceinline__ void top8_update(float& td, int& tp, float cd, int cpos, bool valid, int lane) {
    const float thr_d = __shfl_sync(0xffffffffu, td, 7);
    const int thr_p = __shfl_sync(0xffffffffu, tp, 7);
    unsigned m = __ballot_sync(0xffffffffu, valid && cand_less(cd, cpos, thr_d, thr_p));
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const float xd = __shfl_sync(0xffffffffu, cd, src);
        const int xp = __shfl_sync(0xffffffffu, cpos, src);
        const bool lt = cand_less(td, tp, xd, xp);
        const int rank = __popc(__ballot_sync(0xffffffffu, lt) & 0xffu);
        const float ud = __shfl_up_sync(0xffffffffu, td, 1);
        const int up = __shfl_up_sync(0xffffffffu, tp, 1);
        if (lane < 8) {
            if (lane == rank) {
                td = xd;
                tp = xp;
            } else if (lane > rank) {
                td = ud;
                tp = up;
            }
        }
    }
}

}  // namespace rvcr
#endif
