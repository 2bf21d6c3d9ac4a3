// coarse_exact.cu -- exact (canonical fp32) nearest-centroid assignment on the FFMA pipe.
//
// Replaces quantizer->search(n, x, k=1) = faiss IndexFlatL2::search -> knn_L2sqr
// (rvc/infer/pipeline.py:499 through IndexIVF::search; rvc/train/process/extract_index.py:69
// through IndexIVF::add_with_ids -> quantizer->assign).  SURVEY a5/A4: strict "best > dis"
// replacement scanning centroids in ascending id order => lowest centroid id wins ties.
//
// This is the brute-force kernel: every (query, centroid) distance is evaluated in the canonical
// fvec_L2sqr order (pairdist.cuh), so its output defines what "FP32-exact" means for the
// tcgen05 path (coarse_tc.cu), which uses it for margin-overflow queries and which the tests
// compare against.  CTA tile = 32 queries (resident) x a range of centroids (streamed);
// grid = (query tiles, centroid splits); splits merge through a 64-bit atomicMin on
// (distance bits << 32 | centroid id) -- distances are >= +0 so the unsigned order of the bits is
// the numeric order, and the low word breaks ties towards the lowest id.
#include <algorithm>

#include "pairdist.cuh"

namespace rvcr {

constexpr int CE_TQ = 8;
constexpr int CE_QT = 4 * CE_TQ;  // 32 queries per CTA
constexpr int CE_NSTAGE = 6;

__global__ void init_best_kernel(unsigned long long* best, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) best[i] = ~0ull;
}

// grid = (query-tile slots, centroid splits).  A CTA loops over query tiles x, x + gridDim.x, ...;
// the number of queries is nq, or *nsel_ptr when given (device-side count of the selected list).
__global__ void __launch_bounds__(PD_THREADS, 1)
coarse_exact_kernel(const float* __restrict__ q, const int32_t* __restrict__ qsel, int64_t nq,
                    const int32_t* __restrict__ nsel_ptr, const float* __restrict__ cent, int64_t nlist, int d,
                    int blocks_per_split, unsigned long long* __restrict__ best, float* __restrict__ dmat) {
    extern __shared__ __align__(16) float smem[];
    const int QS = d + 8;
    float* sQ = smem;             // CE_QT x QS
    float* sR = sQ + CE_QT * QS;  // CE_NSTAGE x 32 x 128
    __shared__ unsigned long long sBest[CE_QT];

    const int tid = threadIdx.x, w = tid >> 5;
    const int l = tid & 7, qg = (tid >> 3) & 3;
    const int nkc = d / PD_KC;
    if (nsel_ptr) nq = *nsel_ptr;

    const int64_t nblk_all = (nlist + PD_RB - 1) / PD_RB;
    const int64_t blk0 = (int64_t)blockIdx.y * blocks_per_split;
    int64_t blk1 = blk0 + blocks_per_split;
    if (blk1 > nblk_all) blk1 = nblk_all;
    if (blk0 >= blk1) return;
    const int total = (int)(blk1 - blk0) * nkc;
    const int64_t rmax = nlist - 1;
    const int64_t row_base = blk0 * PD_RB;

    for (int64_t q0 = (int64_t)blockIdx.x * CE_QT; q0 < nq; q0 += (int64_t)gridDim.x * CE_QT) {
        const int cnt = (int)((nq - q0) < CE_QT ? (nq - q0) : CE_QT);
        __syncthreads();  // previous tile fully consumed (sQ, sR, sBest)
        if (tid < CE_QT) sBest[tid] = ~0ull;
        {
            const int nvec = d >> 2;
            for (int idx = tid; idx < CE_QT * nvec; idx += PD_THREADS) {
                int r = idx / nvec, c = idx - r * nvec;
                int rr = r < cnt ? r : cnt - 1;
                int64_t qi = qsel ? (int64_t)qsel[q0 + rr] : (q0 + rr);
                cp_async16(sQ + r * QS + c * 4, q + qi * d + c * 4);
            }
        }
#pragma unroll
        for (int s = 0; s < CE_NSTAGE - 1; ++s) {
            if (s < total)
                pd_load_stage(sR + s * PD_RB * PD_KC, cent, row_base + (int64_t)(s / nkc) * PD_RB, rmax, d, s % nkc, tid);
            cp_async_commit();
        }

        float bd[CE_TQ];
        int bj[CE_TQ];
#pragma unroll
        for (int a = 0; a < CE_TQ; ++a) {
            bd[a] = RVCR_FLT_MAX;
            bj[a] = -1;
        }

        PairTile<CE_TQ> tile;
        tile.clear();
        int kc = 0;
        int64_t rb = blk0;
        for (int it = 0; it < total; ++it) {
            cp_async_wait<CE_NSTAGE - 2>();
            __syncthreads();
            {
                int nx = it + CE_NSTAGE - 1;
                if (nx < total)
                    pd_load_stage(sR + (nx % CE_NSTAGE) * PD_RB * PD_KC, cent, row_base + (int64_t)(nx / nkc) * PD_RB, rmax,
                                  d, nx % nkc, tid);
                cp_async_commit();
            }
            tile.chunk(sQ + qg * QS + kc * PD_KC + l, QS, sR + (it % CE_NSTAGE) * PD_RB * PD_KC + (4 * w) * PD_KC + l);
            if (++kc == nkc) {
                kc = 0;
                tile.reduce();
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int64_t j = rb * PD_RB + 4 * w + b;  // ascending within a thread
                    if (j < nlist) {
                        if (dmat && l == 0) {   // nprobe > 1 (multiprobe.cu): the whole (query, centroid) distance matrix
#pragma unroll
                            for (int a = 0; a < CE_TQ; ++a)
                                if (qg + 4 * a < cnt) dmat[(q0 + qg + 4 * a) * nlist + j] = tile.acc[a][b];
                        }
#pragma unroll
                        for (int a = 0; a < CE_TQ; ++a) {
                            // faiss: if (best > dis) replace  -- strict, first (lowest id) wins; a NaN
                            // distance never replaces, so a NaN query keeps id -1
                            if (bd[a] > tile.acc[a][b]) {
                                bd[a] = tile.acc[a][b];
                                bj[a] = (int)j;
                            }
                        }
                    }
                }
                tile.clear();
                ++rb;
            }
        }
        cp_async_wait<0>();
        __syncthreads();
        if (l == 0) {
#pragma unroll
            for (int a = 0; a < CE_TQ; ++a)
                if (bj[a] >= 0) {
                    unsigned long long key = ((unsigned long long)__float_as_uint(bd[a]) << 32) | (unsigned)bj[a];
                    atomicMin(&sBest[qg + 4 * a], key);
                }
        }
        __syncthreads();
        if (tid < cnt) {
            unsigned long long key = sBest[tid];
            if (key != ~0ull) {
                int64_t qi = qsel ? (int64_t)qsel[q0 + tid] : (q0 + tid);
                atomicMin(&best[qi], key);
            }
        }
    }
}

__global__ void unpack_best_kernel(const unsigned long long* __restrict__ best, const int32_t* __restrict__ qsel,
                                   int64_t nsel, const int32_t* __restrict__ nsel_ptr, int64_t list_begin,
                                   int32_t* __restrict__ list_out, float* __restrict__ dist_out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (nsel_ptr) nsel = *nsel_ptr;
    if (i >= nsel) return;
    int64_t qi = qsel ? (int64_t)qsel[i] : i;
    unsigned long long key = best[qi];
    if (key == ~0ull) {
        list_out[qi] = -1;
        if (dist_out) dist_out[qi] = RVCR_FLT_MAX;
    } else {
        list_out[qi] = (int32_t)((int64_t)(unsigned)(key & 0xffffffffu) + list_begin);
        if (dist_out) dist_out[qi] = __uint_as_float((unsigned)(key >> 32));
    }
}

static size_t ce_smem_bytes(int d) {
    return sizeof(float) * ((size_t)CE_QT * (d + 8) + (size_t)CE_NSTAGE * PD_RB * PD_KC);
}

static int ce_prepare(rvcr_index* h, size_t* smem_out) {
    if (h->d % PD_KC != 0) {
        set_error("d=%d unsupported (must be a multiple of %d)", h->d, PD_KC);
        return RVCR_ERR_UNSUPPORTED;
    }
    const size_t smem = ce_smem_bytes(h->d);
    if (smem > 226 * 1024) {
        set_error("d=%d needs %zu B of shared memory", h->d, smem);
        return RVCR_ERR_UNSUPPORTED;
    }
    static SmemOptIn opt_in;   // zero-initialised (static storage)
    RVCR_TRY(opt_in.ensure(coarse_exact_kernel, smem, h->device));
    *smem_out = smem;
    return RVCR_OK;
}

// Nearest centroid for queries q[i] (i < n).  list_out / dist_out are indexed by query index.
int coarse_exact(rvcr_index* h, const float* q, int64_t n, const int32_t* qsel, int64_t nsel, int32_t* list_out,
                 float* dist_out, cudaStream_t s) {
    const int64_t nwork = qsel ? nsel : n;
    if (nwork <= 0) return RVCR_OK;
    size_t smem = 0;
    RVCR_TRY(ce_prepare(h, &smem));
    RVCR_TRY(h->ws_best.reserve((size_t)n * 8));
    unsigned long long* best = h->ws_best.as<unsigned long long>();
    init_best_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(best, n);
    RVCR_CHECK_LAUNCH();
    const int64_t qtiles = (nwork + CE_QT - 1) / CE_QT;
    const int64_t nblk = (h->nlist + PD_RB - 1) / PD_RB;
    // enough CTAs for >= 2 waves when the query tile count alone cannot fill the chip
    int64_t nsplit = (2 * h->num_sms + qtiles - 1) / qtiles;
    const int64_t max_split = (nblk + 3) / 4;  // keep >= 4 row-blocks per split
    if (nsplit > max_split) nsplit = max_split;
    if (nsplit < 1) nsplit = 1;
    if (nsplit > 65535) nsplit = 65535;
    const int bps = (int)((nblk + nsplit - 1) / nsplit);
    nsplit = (nblk + bps - 1) / bps;
    dim3 grid((unsigned)qtiles, (unsigned)nsplit);
    coarse_exact_kernel<<<grid, PD_THREADS, smem, s>>>(q, qsel, nwork, nullptr, h->cent, h->nlist, h->d, bps, best, nullptr);
    RVCR_CHECK_LAUNCH();
    unpack_best_kernel<<<(unsigned)((nwork + 255) / 256), 256, 0, s>>>(best, qsel, nwork, nullptr, 0, list_out,
                                                                         dist_out);
    RVCR_CHECK_LAUNCH();
    count_launch(h, 3);
    return RVCR_OK;
}

// Canonical distances of queries q[0 .. n) to every centroid, written to dmat (n x nlist, row-major) -- the coarse step of
// an index whose file says nprobe > 1 (multiprobe.cu selects the nprobe smallest per row).
int coarse_exact_matrix(rvcr_index* h, const float* q, int64_t n, float* dmat, cudaStream_t s) {
    if (n <= 0) return RVCR_OK;
    size_t smem = 0;
    RVCR_TRY(ce_prepare(h, &smem));
    RVCR_TRY(h->ws_best.reserve((size_t)n * 8));
    const int64_t qtiles = (n + CE_QT - 1) / CE_QT;
    const int64_t nblk = (h->nlist + PD_RB - 1) / PD_RB;
    int64_t nsplit = (2 * h->num_sms + qtiles - 1) / qtiles;
    const int64_t max_split = (nblk + 3) / 4;
    if (nsplit > max_split) nsplit = max_split;
    if (nsplit < 1) nsplit = 1;
    if (nsplit > 65535) nsplit = 65535;
    const int bps = (int)((nblk + nsplit - 1) / nsplit);
    nsplit = (nblk + bps - 1) / bps;
    dim3 grid((unsigned)qtiles, (unsigned)nsplit);
    init_best_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(h->ws_best.as<unsigned long long>(), n);
    RVCR_CHECK_LAUNCH();
    coarse_exact_kernel<<<grid, PD_THREADS, smem, s>>>(q, nullptr, n, nullptr, h->cent, h->nlist, h->d, bps,
                                                        h->ws_best.as<unsigned long long>(), dmat);
    RVCR_CHECK_LAUNCH();
    count_launch(h, 2);
    return RVCR_OK;
}

// Same, for the queries listed in qsel[0 .. *nsel_dev) where the count lives on the device (no host
// synchronisation): a fixed (64 x splits) grid whose CTAs loop over however many query tiles there
// turn out to be.  Used by coarse_tc.cu for margin-overflow queries, normally a handful.
int coarse_exact_selected(rvcr_index* h, const float* q, int64_t n, const int32_t* qsel, const int32_t* nsel_dev,
                          int32_t* list_out, float* dist_out, cudaStream_t s) {
    if (n <= 0) return RVCR_OK;
    size_t smem = 0;
    RVCR_TRY(ce_prepare(h, &smem));
    RVCR_TRY(h->ws_best.reserve((size_t)n * 8));
    unsigned long long* best = h->ws_best.as<unsigned long long>();
    init_best_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(best, n);
    RVCR_CHECK_LAUNCH();
    const int64_t nblk = (h->nlist + PD_RB - 1) / PD_RB;
    int64_t nsplit = (nblk + 3) / 4;
    if (nsplit > 64) nsplit = 64;
    if (nsplit < 1) nsplit = 1;
    const int bps = (int)((nblk + nsplit - 1) / nsplit);
    nsplit = (nblk + bps - 1) / bps;
    int64_t gx = (n + CE_QT - 1) / CE_QT;
    const int64_t gx_max = std::max<int64_t>(1, (int64_t)(8 * h->num_sms) / nsplit);
    if (gx > gx_max) gx = gx_max;
    dim3 grid((unsigned)gx, (unsigned)nsplit);
    coarse_exact_kernel<<<grid, PD_THREADS, smem, s>>>(q, qsel, n, nsel_dev, h->cent, h->nlist, h->d, bps, best, nullptr);
    RVCR_CHECK_LAUNCH();
    unpack_best_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(best, qsel, n, nsel_dev, 0, list_out, dist_out);
    RVCR_CHECK_LAUNCH();
    count_launch(h, 3);
    return RVCR_OK;
}

}  // namespace rvcr
