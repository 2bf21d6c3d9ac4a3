/*
 * oracle/rvc_oracle.c -- CPU restatement of the RVC feature-retrieval path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file, and only as the checker / the timed CPU baseline.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in faiss-cpu==1.7.3 and
 * numpy==1.26.4 (reference requirements.txt:12, :5), neither of which is vendored
 * under /root/reference nor installable here, and the reference ships no tests or
 * golden vectors for it.  This file restates the published faiss 1.7.3 algorithms
 *  anchored on the reference's own call sites:
 *
 *   rvc/infer/pipeline.py:499      index.search(npy, k=8)        -> rvco_coarse_*, rvco_scan_topk
 *   rvc/infer/pipeline.py:500-506  weights / gather-sum / blend  -> rvco_tail  (this part IS pinned:
 *                                  tests/golden holds outputs of the reference's own function body
 *                                  executed here, see tests/golden/make_golden.py)
 *   rvc/train/process/extract_index.py:67-69  index.add(...)     -> rvco_coarse_* + rvco_build_lists
 *
 * Canonical distance (faiss utils/distances_simd.cpp fvec_L2sqr, AVX build):
 *   one 8-lane accumulator; per 8 floats t = x - y; acc = fma(t, t, acc); final reduce
 *   m[l] = acc[l] + acc[l+4] (l<4); result = (m0 + m1) + (m2 + m3).
 *   The FMA contraction is what gcc emits for the intrinsics pair under -mavx2 -mfma; the
 *   non-contracted variant is available as rvco_l2sqr_nofma for tie-zone classification.
 *
 * Build: see oracle/build.py  (gcc -O3 -mavx2 -mfma -ffp-contract=off -fopenmp -shared -fPIC)
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define RVCO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* canonical squared-L2: faiss fvec_L2sqr (AVX, 8 lanes, FMA); d must be a multiple of 8       */

static inline float l2sqr_fma(const float* x, const float* y, int d) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < d; i += 8) {
        for (int l = 0; l < 8; ++l) {
            float t = x[i + l] - y[i + l];
            acc[l] = __builtin_fmaf(t, t, acc[l]);
        }
    }
    float m0 = acc[0] + acc[4], m1 = acc[1] + acc[5], m2 = acc[2] + acc[6], m3 = acc[3] + acc[7];
    return (m0 + m1) + (m2 + m3);
}

static inline float l2sqr_nofma(const float* x, const float* y, int d) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < d; i += 8) {
        for (int l = 0; l < 8; ++l) {
            float t = x[i + l] - y[i + l];
            float p = t * t; /* -ffp-contract=off keeps this a separate rounding */
            acc[l] = acc[l] + p;
        }
    }
    float m0 = acc[0] + acc[4], m1 = acc[1] + acc[5], m2 = acc[2] + acc[6], m3 = acc[3] + acc[7];
    return (m0 + m1) + (m2 + m3);
}

RVCO_API float rvco_l2sqr(const float* x, const float* y, int d) { return l2sqr_fma(x, y, d); }
RVCO_API float rvco_l2sqr_nofma(const float* x, const float* y, int d) { return l2sqr_nofma(x, y, d); }

RVCO_API double rvco_l2sqr_f64(const float* x, const float* y, int d) {
    double s = 0.0;
    for (int i = 0; i < d; ++i) {
        double t = (double)x[i] - (double)y[i];
        s += t * t;
    }
    return s;
}

RVCO_API int rvco_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the timed CPU legs of bench.py set the count
 * explicitly (faiss itself would use every core: omp_get_max_threads() under the launcher's default environment) */
RVCO_API void rvco_set_num_threads(int n) {
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* coarse quantizer, k = 1 (IndexFlatL2::search -> knn_L2sqr).                       */
/* Strict "best > dis" replacement scanning j ascending  =>  ties go to the lowest centroid id. */
/* mode 0: canonical FMA distance, mode 1: non-FMA distance.  gap_out (optional) receives the   */
/* float64 gap between the best and the second-best centroid, used for tie-zone accounting.     */

RVCO_API void rvco_coarse_assign(const float* q, int64_t nq, const float* cent, int64_t nlist, int d,
                                 int mode, int32_t* list_out, float* dist_out, double* gap_out) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < nq; ++i) {
        const float* x = q + i * d;
        float best = FLT_MAX;
        int32_t bj = -1;
        for (int64_t j = 0; j < nlist; ++j) {
            float dis = mode ? l2sqr_nofma(x, cent + j * d, d) : l2sqr_fma(x, cent + j * d, d);
            if (best > dis) {
                best = dis;
                bj = (int32_t)j;
            }
        }
        list_out[i] = bj;
        if (dist_out) dist_out[i] = best;
        if (gap_out) {
            double b1 = DBL_MAX, b2 = DBL_MAX;
            for (int64_t j = 0; j < nlist; ++j) {
                double dis = rvco_l2sqr_f64(x, cent + j * d, d);
                if (dis < b1) {
                    b2 = b1;
                    b1 = dis;
                } else if (dis < b2) {
                    b2 = dis;
                }
            }
            gap_out[i] = b2 - b1;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* coarse quantizer, k = nprobe > 1 (IndexIVF::search -> quantizer->search(n, x, nprobe); a file written  */
/* by another fork may carry any nprobe, Appendix B offset 45).  The nprobe nearest centroids by canonical distance,     */
/* ascending; among equal distances the lower centroid id first (faiss's k > 1 heap leaves that order                    */
/* implementation-defined: documented tie).  Queries with a NaN distance everywhere get -1 ids and FLT_MAX distances.    */
RVCO_API void rvco_coarse_topn(const float* q, int64_t nq, const float* cent, int64_t nlist, int d, int nprobe,
                               int32_t* list_out /* nq x nprobe */, float* dist_out /* nq x nprobe, optional */) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < nq; ++i) {
        const float* x = q + i * d;
        int32_t* lo = list_out + i * nprobe;
        float bd[nprobe];
        int cnt = 0;
        for (int64_t j = 0; j < nlist; ++j) {
            float dis = l2sqr_fma(x, cent + j * d, d);
            if (!(dis == dis)) continue;
            if (cnt == nprobe && !(dis < bd[cnt - 1])) continue;   /* strict: an equal distance with a higher id stays out */
            int p = cnt < nprobe ? cnt++ : nprobe - 1;
            while (p > 0 && dis < bd[p - 1]) {
                bd[p] = bd[p - 1];
                lo[p] = lo[p - 1];
                --p;
            }
            bd[p] = dis;
            lo[p] = (int32_t)j;
        }
        for (int p = cnt; p < nprobe; ++p) {
            bd[p] = FLT_MAX;
            lo[p] = -1;
        }
        if (dist_out)
            for (int p = 0; p < nprobe; ++p) dist_out[i * nprobe + p] = bd[p];
    }
}

/* ------------------------------------------------------------------------------------------ */
/* inverted-list scan, nprobe = 1 (IVFFlatScanner<L2,CMax>::scan_codes + heap).      */
/* The faiss max-heap keeps, among equal distances at the k-th boundary, the entry scanned      */
/* first (strict "top > dis").  The order faiss gives to EQUAL distances inside the top-k is    */
/* heap-implementation defined; this restatement fixes it as (distance asc, list position asc). */
/* Lists shorter than k are padded with (FLT_MAX, -1).  assign[i] < 0 gives an all-pad row.     */

typedef struct {
    float dis;
    int64_t pos;
} cand_t;

static inline int cand_less(float da, int64_t pa, float db, int64_t pb) {
    return (da < db) || (da == db && pa < pb);
}

RVCO_API void rvco_scan_topk(const float* q, int64_t nq, const int32_t* assign, const int64_t* list_off,
                             const float* vecs, const int64_t* ids, int d, int k, int mode, float* D,
                             int64_t* I, int64_t* P) {
#pragma omp parallel
    {
        cand_t* top = (cand_t*)malloc(sizeof(cand_t) * (size_t)k);
#pragma omp for schedule(dynamic, 8)
        for (int64_t i = 0; i < nq; ++i) {
            int n = 0;
            int32_t l = assign[i];
            if (l >= 0) {
                const float* x = q + i * d;
                int64_t b = list_off[l], e = list_off[l + 1];
                for (int64_t p = b; p < e; ++p) {
                    float dis = mode ? l2sqr_nofma(x, vecs + p * d, d) : l2sqr_fma(x, vecs + p * d, d);
                    /* sorted insertion == heap with (dist,pos) order; strict compare keeps earlier */
                    if (n == k && !cand_less(dis, p, top[k - 1].dis, top[k - 1].pos)) continue;
                    int s = (n < k) ? n++ : k - 1;
                    while (s > 0 && cand_less(dis, p, top[s - 1].dis, top[s - 1].pos)) {
                        top[s] = top[s - 1];
                        --s;
                    }
                    top[s].dis = dis;
                    top[s].pos = p;
                }
            }
            for (int j = 0; j < k; ++j) {
                if (j < n) {
                    D[i * k + j] = top[j].dis;
                    I[i * k + j] = ids[top[j].pos];
                    if (P) P[i * k + j] = top[j].pos;
                } else {
                    D[i * k + j] = FLT_MAX;
                    I[i * k + j] = -1;
                    if (P) P[i * k + j] = -1;
                }
            }
        }
        free(top);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* numpy / torch tail, rvc/infer/pipeline.py:500-506, in the exact fp32 operation order of      */
/* SURVEY Appendix C:                                                                           */
/*   weight = np.square(1 / score)                        fp32 divide, fp32 multiply            */
/*   weight /= weight.sum(axis=1, keepdims=True)          pairwise ((w0+w1)+(w2+w3))+((w4+w5)+(w6+w7)) */
/*   npy = np.sum(big_npy[ix] * weight[..., None], axis=1) products rounded, summed k = 0..7    */
/*   feats = npy * index_rate + (1 - index_rate) * feats  r32 = (float)rate, (float)(1.0 - rate) */
/* ix = -1 wraps to the last row of big_npy exactly as numpy's negative index does.             */
/* Only k == 8 reproduces numpy's pairwise order; other k sum left to right.                    */

RVCO_API void rvco_tail(const float* score, const int64_t* ix, int64_t nq, int k, const float* big_npy,
                        int64_t ntotal, int d, double index_rate, const float* feats, float* out) {
    const float r32 = (float)index_rate;
    const float omr32 = (float)(1.0 - index_rate);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nq; ++i) {
        float w[64];
        for (int j = 0; j < k; ++j) {
            float inv = 1.0f / score[i * k + j];
            w[j] = inv * inv;
        }
        float s;
        if (k == 8) {
            s = ((w[0] + w[1]) + (w[2] + w[3])) + ((w[4] + w[5]) + (w[6] + w[7]));
        } else {
            s = w[0];
            for (int j = 1; j < k; ++j) s = s + w[j];
        }
        for (int j = 0; j < k; ++j) w[j] = w[j] / s;
        for (int e = 0; e < d; ++e) {
            float acc = 0.f;
            for (int j = 0; j < k; ++j) {
                int64_t r = ix[i * k + j];
                if (r < 0) r += ntotal;
                float p = big_npy[r * d + e] * w[j];
                acc = (j == 0) ? p : acc + p;
            }
            float a = acc * r32;
            float b = omr32 * feats[i * d + e];
            out[i * d + e] = a + b;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* IndexIVFFlat::add_core: entry i of the batch gets id = id0 + i and is appended   */
/* to list assign[i]; within a list entries end up in ascending id order.  Given the            */
/* assignment of ALL ntotal vectors this builds the final list-ordered layout in one pass.      */
/* assign[i] < 0 (NaN input) entries are dropped from the lists.                                */

RVCO_API void rvco_build_lists(const float* x, int64_t n, const int32_t* assign, int64_t nlist, int d,
                               int64_t* list_off /* nlist+1 */, float* vecs_out, int64_t* ids_out) {
    memset(list_off, 0, sizeof(int64_t) * (size_t)(nlist + 1));
    for (int64_t i = 0; i < n; ++i)
        if (assign[i] >= 0) list_off[assign[i] + 1]++;
    for (int64_t l = 0; l < nlist; ++l) list_off[l + 1] += list_off[l];
    int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * (size_t)nlist);
    memcpy(cur, list_off, sizeof(int64_t) * (size_t)nlist);
    for (int64_t i = 0; i < n; ++i) {
        int32_t l = assign[i];
        if (l < 0) continue;
        int64_t p = cur[l]++;
        memcpy(vecs_out + p * d, x + i * d, sizeof(float) * (size_t)d);
        ids_out[p] = i;
    }
    free(cur);
}
