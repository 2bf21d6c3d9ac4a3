// multiprobe.cu -- index.search(npy, k=8) for an index whose header says nprobe > 1.
//
// The reference builds its own indices with nprobe = 1 (rvc/train/process/extract_index.py:64), but the field is
// persisted in the file (offset 45) and files made by other forks arrive through the download tab
// (tabs/download/download.py:41-45) and are searched as they are (rvc/infer/pipeline.py:499, 555): faiss honours the
// stored nprobe, so must the drop-in.  faiss semantics restated:
//   IndexIVF::search: nprobe = min(nlist, nprobe); quantizer->search(n, x, nprobe) = the nprobe nearest centroids,
//   ascending; search_preassigned scans those lists in coarse-rank order into ONE k-heap with the strict replacement
//   rule -- among equal distances the entry scanned first (lower probe rank, then lower list position) wins.
// Here: (1) the canonical distance of every (query, centroid) pair on the exact FFMA kernel (coarse_exact.cu), the
// nprobe smallest per query selected by (distance, centroid id); (2) one ordinary single-list scan per probe rank
// (the same kernels as nprobe = 1, so every reported distance is the canonical one); (3) a merge of the nprobe x 8
// candidates in scan order with strict `<`; (4) the ordinary gather + weights + blend kernel on the merged (distance,
// row position) table.  This is the interop path: exact, not tuned (third-party clips are C1-sized).
#include "common.cuh"

namespace rvcr {

constexpr int MP_MAX_NPROBE = 128;

int coarse_exact_matrix(rvcr_index* h, const float* q, int64_t n, float* dmat, cudaStream_t s);   // coarse_exact.cu

// One warp per query: the np smallest entries of its row of the distance matrix, ascending by (distance, centroid id).
// A NaN distance never enters (a NaN query ends with np entries of -1).
__global__ void mp_select_kernel(const float* __restrict__ dmat, int64_t n, int64_t nlist, int np, int32_t* __restrict__ lists) {
    __shared__ float sd[8][MP_MAX_NPROBE];
    __shared__ int sj[8][MP_MAX_NPROBE];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t qi = blockIdx.x * (int64_t)(blockDim.x >> 5) + w;
    if (qi >= n) return;
    const float* row = dmat + qi * nlist;
    float* d = sd[w];
    int* j = sj[w];
    int cnt = 0;                          // warp-uniform
    for (int64_t j0 = 0; j0 < nlist; j0 += 32) {
        const int64_t jj = j0 + lane;
        const float v = jj < nlist ? row[jj] : 0.f;
        const float thr = cnt == np ? d[np - 1] : __int_as_float(0x7f800000);
        // strict: at a full list an equal distance with a higher id stays out; +inf distances enter only while not full
        unsigned m = __ballot_sync(0xffffffffu, jj < nlist && v == v && (v < thr || cnt < np));
        while (m) {                       // ascending lane == ascending centroid id
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const float x = __shfl_sync(0xffffffffu, v, src);
            if (lane == 0) {
                if (cnt < np || x < d[np - 1]) {
                    int p = cnt < np ? cnt : np - 1;
                    while (p > 0 && x < d[p - 1]) {
                        d[p] = d[p - 1];
                        j[p] = j[p - 1];
                        --p;
                    }
                    d[p] = x;
                    j[p] = (int)(j0 + src);
                }
            }
            if (cnt < np) ++cnt;
            __syncwarp();
        }
    }
    __syncwarp();
    for (int p = lane; p < np; p += 32) lists[qi * np + p] = p < cnt ? j[p] : -1;
}

__global__ void mp_column_kernel(const int32_t* __restrict__ lists, int64_t n, int np, int p, int32_t* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = lists[i * np + p];
}

// One thread per query: candidates of probe 0 (already ascending), then probe 1, ... inserted with strict `<`, so an
// equal distance scanned later stays behind (and out, at the 8th place) -- the faiss heap rule.
__global__ void mp_merge_kernel(const float* __restrict__ Dp, const int32_t* __restrict__ Pp, const int64_t* __restrict__ Ip,
                                int np, int64_t n, float* __restrict__ D, int32_t* __restrict__ P, int64_t* __restrict__ I) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float d[RVCR_K];
    int32_t ps[RVCR_K];
    int64_t id[RVCR_K];
#pragma unroll
    for (int k = 0; k < RVCR_K; ++k) {
        d[k] = RVCR_FLT_MAX;
        ps[k] = -1;
        id[k] = -1;
    }
    for (int p = 0; p < np; ++p) {
        const int64_t o = ((int64_t)p * n + i) * RVCR_K;
        for (int k = 0; k < RVCR_K; ++k) {
            const int32_t pos = Pp[o + k];
            if (pos < 0) break;                        // padding: nothing further from this list
            const float x = Dp[o + k];
            if (!(x < d[RVCR_K - 1])) break;           // ascending inside a list: the rest cannot enter either
            int t = RVCR_K - 1;
#pragma unroll
            for (int u = RVCR_K - 1; u > 0; --u) {
                if (t == u && x < d[u - 1]) {
                    d[u] = d[u - 1];
                    ps[u] = ps[u - 1];
                    id[u] = id[u - 1];
                    t = u - 1;
                }
            }
#pragma unroll
            for (int u = 0; u < RVCR_K; ++u)
                if (t == u) {
                    d[u] = x;
                    ps[u] = pos;
                    id[u] = Ip ? Ip[o + k] : -1;
                }
        }
    }
#pragma unroll
    for (int k = 0; k < RVCR_K; ++k) {
        D[i * RVCR_K + k] = d[k];
        P[i * RVCR_K + k] = ps[k];
        if (I) I[i * RVCR_K + k] = id[k];
    }
}

// index.search / the fused retrieve + blend with nprobe > 1.  Caller holds the handle lock, has checked searchability,
// finalised pending adds and entered the stream.
int search_multiprobe(rvcr_index* h, float* feats_or_q, int64_t n, bool blend, double index_rate, float* D_dev, int64_t* I_dev,
                      cudaStream_t s) {
    const int np = (int)(h->nprobe < h->nlist ? h->nprobe : h->nlist);
    if (np > MP_MAX_NPROBE) {
        set_error("nprobe=%lld unsupported (at most %d probes per query)", (long long)h->nprobe, MP_MAX_NPROBE);
        return RVCR_ERR_UNSUPPORTED;
    }
    if (h->own_lo != 0 || h->own_hi != h->nlist) {
        set_error("nprobe > 1 on a shard is unsupported (the sharded path routes each query to one owner)");
        return RVCR_ERR_UNSUPPORTED;
    }
    DevBuf &b_lists = h->ws_mp[0], &b_col = h->ws_mp[1], &b_D = h->ws_mp[2], &b_P = h->ws_mp[3], &b_I = h->ws_mp[4], &b_mat = h->ws_mp[5];
    RVCR_TRY(b_lists.reserve((size_t)n * np * 4));
    RVCR_TRY(b_col.reserve((size_t)n * 4));
    RVCR_TRY(b_D.reserve((size_t)n * np * RVCR_K * 4));
    RVCR_TRY(b_P.reserve((size_t)n * np * RVCR_K * 4));
    if (I_dev) RVCR_TRY(b_I.reserve((size_t)n * np * RVCR_K * 8));
    {   // (1) coarse: nprobe nearest centroids, in query chunks of <= 1 GB of distance matrix
        PhaseTimer pt(h, RVCR_PHASE_COARSE_MAIN, s);
        int64_t chunk = (int64_t)(1u << 28) / h->nlist;
        if (chunk < 32) chunk = 32;
        if (chunk > n) chunk = n;
        RVCR_TRY(b_mat.reserve((size_t)chunk * h->nlist * 4));
        for (int64_t q0 = 0; q0 < n; q0 += chunk) {
            const int64_t m = n - q0 < chunk ? n - q0 : chunk;
            RVCR_TRY(coarse_exact_matrix(h, feats_or_q + q0 * h->d, m, b_mat.as<float>(), s));
            mp_select_kernel<<<(unsigned)((m + 7) / 8), 256, 0, s>>>(b_mat.as<float>(), m, h->nlist, np, b_lists.as<int32_t>() + q0 * np);
            RVCR_CHECK_LAUNCH();
            count_launch(h);
        }
    }
    RVCR_TRY(h->ws_P.reserve((size_t)n * RVCR_K * 4));
    float* D = D_dev;
    if (!D) {
        RVCR_TRY(h->ws_D.reserve((size_t)n * RVCR_K * 4));
        D = h->ws_D.as<float>();
    }
    {   // (2) one single-list scan per probe rank, (3) merge in scan order
        PhaseTimer pt(h, RVCR_PHASE_SCAN, s);
        h->fuse.active = false;
        for (int p = 0; p < np; ++p) {
            mp_column_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(b_lists.as<int32_t>(), n, np, p, b_col.as<int32_t>());
            RVCR_CHECK_LAUNCH();
            RVCR_TRY(group_queries(h, b_col.as<int32_t>(), n, false, s));
            RVCR_TRY(scan_topk(h, feats_or_q, n, b_D.as<float>() + (size_t)p * n * RVCR_K, b_P.as<int32_t>() + (size_t)p * n * RVCR_K,
                               I_dev ? b_I.as<int64_t>() + (size_t)p * n * RVCR_K : nullptr, s));
        }
        mp_merge_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(b_D.as<float>(), b_P.as<int32_t>(), I_dev ? b_I.as<int64_t>() : nullptr,
                                                                     np, n, D, h->ws_P.as<int32_t>(), I_dev);
        RVCR_CHECK_LAUNCH();
        count_launch(h, np + 1);
    }
    if (blend) {   // (4) pipeline.py:500-506 on the merged table
        PhaseTimer pt(h, RVCR_PHASE_BLEND, s);
        RVCR_TRY(gather_blend(h, feats_or_q, n, D, h->ws_P.as<int32_t>(), nullptr, index_rate, s));
    }
    h->stats.calls_search += 1;
    h->stats.queries += n;
    return RVCR_OK;
}

}  // namespace rvcr
