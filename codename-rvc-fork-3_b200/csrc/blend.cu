// blend.cu -- fused gather + 1/score^2 weighting + index_rate blend, written in place; plus the
// two small cross-shard merge kernels.
//
// Replaces rvc/infer/pipeline.py:500-506:
//     weight = np.square(1 / score)
//     weight /= weight.sum(axis=1, keepdims=True)
//     npy = np.sum(big_npy[ix] * np.expand_dims(weight, axis=2), axis=1)
//     feats = torch.from_numpy(npy).unsqueeze(0).to(device) * index_rate + (1 - index_rate) * feats
// The reference materialises an (n, 8, d) temporary on the host and crosses PCIe twice; here one
// kernel reads the 8 neighbour rows straight from the list-ordered storage (positions come from
// the scan, so no id -> row map and no second `big_npy` copy), and overwrites feats.
//
// Bit-exactness contract: every operation is an explicitly rounded fp32 op in numpy's order --
// divide, square, pairwise row sum ((w0+w1)+(w2+w3))+((w4+w5)+(w6+w7)), divide, then per element
// products rounded and accumulated k = 0..7 left to right (no FMA), then
// fl(fl(acc*r32) + fl(omr32*feats)) with r32 = (float)index_rate, omr32 = (float)(1.0 - index_rate).
// score == 0 (query equals an indexed vector) gives inf/inf = NaN weights and a NaN row, exactly
// like the reference; (FLT_MAX, -1) padding gives weight 0 and gathers numpy's big_npy[-1].
#include "common.cuh"

namespace rvcr {

// Default: every query qi < n, in place.  With `qorder` the kernel instead walks the sorted positions of
// work items [item_begin, item_end) (a contiguous range: items are ordered by sorted position) and writes
// the blended row of query qorder[p] to out + qorder[p]*d -- `out` may be mapped pinned host memory, in
// which case the rows go straight over PCIe while the next wave of lists is being scanned.
// MODE 1: also apply the 2x upsample + protect mix and write to post.out2; MODE 2: push the row to its home rank's buffer
// over NVLink (separate instantiations: the default kernel, MODE 0, is unchanged).
template <int MODE>
__global__ void __launch_bounds__(256)
gather_blend_kernel(float* __restrict__ feats, int64_t n, int d, const float* __restrict__ D,
                    const int32_t* __restrict__ P, const int32_t* __restrict__ assign, int64_t own_lo,
                    int64_t own_hi, const float* __restrict__ vecs, int64_t last_row_pos, float r32, float omr32,
                    const int32_t* __restrict__ qorder, const int32_t* __restrict__ items,
                    const int32_t* __restrict__ nitems_ptr, int item_begin, int item_end, float* __restrict__ out,
                    PostArgs post, PushArgs push) {
    __shared__ float sw[RVCR_K];
    __shared__ int spos[RVCR_K];
    const int tid = threadIdx.x;
    const int nvec = d >> 2;
    int64_t first = 0, last = n;
    if (qorder) {
        const int nit = *nitems_ptr;
        first = item_begin < nit ? (int64_t)items[item_begin] : n;
        last = item_end < nit ? (int64_t)items[item_end] : n;
    }
    for (int64_t it = first + blockIdx.x; it < last; it += gridDim.x) {
        const int64_t qi = qorder ? (int64_t)qorder[it] : it;
        if (assign) {
            const int64_t a = (int64_t)assign[qi];  // block-uniform
            if (a >= 0 && (a < own_lo || a >= own_hi)) continue;  // another shard's query
        }
        if (tid < 32) {
            float w = 0.f;
            int pos = -1;
            if (tid < RVCR_K) {
                const float score = D[qi * RVCR_K + tid];
                const float inv = __fdiv_rn(1.0f, score);
                w = __fmul_rn(inv, inv);
                pos = P[qi * RVCR_K + tid];
            }
            float s = __fadd_rn(w, __shfl_xor_sync(0xffffffffu, w, 1));
            s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 2));
            s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 4));
            if (tid < RVCR_K) {
                sw[tid] = __fdiv_rn(w, s);
                spos[tid] = pos < 0 ? (int)last_row_pos : pos;  // numpy: big_npy[-1]
            }
        }
        __syncthreads();
        float wk[RVCR_K];
        const float4* rows[RVCR_K];
#pragma unroll
        for (int k = 0; k < RVCR_K; ++k) {
            wk[k] = sw[k];
            rows[k] = spos[k] >= 0 ? reinterpret_cast<const float4*>(vecs + (int64_t)spos[k] * d) : nullptr;
        }
        const float4* f4 = reinterpret_cast<const float4*>(query_row(feats, push, qi, d));
        float4* o4 = MODE == 2 ? push_row_ptr(push, qi, d) : reinterpret_cast<float4*>((out ? out : feats) + qi * d);
        for (int c = tid; c < nvec; c += blockDim.x) {
            float4 v[RVCR_K];
#pragma unroll
            for (int k = 0; k < RVCR_K; ++k) v[k] = rows[k] ? __ldg(rows[k] + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 f = f4[c];
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
            if (MODE == 1) post_store(post, qi, d, c, o, f);   // fused 2x upsample + protect mix (pipeline.py:463-481)
            else o4[c] = o;
        }
        __syncthreads();
    }
}

int gather_blend(rvcr_index* h, float* feats, int64_t n, const float* D, const int32_t* P, const int32_t* assign,
                 double index_rate, cudaStream_t s) {
    if (n <= 0) return RVCR_OK;
    if (h->d % 4 != 0) {
        set_error("d=%d unsupported (must be a multiple of 4)", h->d);
        return RVCR_ERR_UNSUPPORTED;
    }
    const float r32 = (float)index_rate;
    const float omr32 = (float)(1.0 - index_rate);
    int threads = ((h->d / 4 + 31) / 32) * 32;
    if (threads > 256) threads = 256;
    if (threads < 32) threads = 32;
    int64_t grid = (int64_t)h->num_sms * (2048 / threads);
    if (grid > n) grid = n;
    auto kern = h->fuse.post.out2 ? gather_blend_kernel<1> : (h->fuse.push.bases ? gather_blend_kernel<2> : gather_blend_kernel<0>);
    kern<<<(unsigned)grid, threads, 0, s>>>(feats, n, h->d, D, P, assign, h->own_lo, h->own_hi, h->vecs,
                                                            h->last_row_pos, r32, omr32, nullptr, nullptr, nullptr, 0, 0,
                                                            nullptr, h->fuse.post, h->fuse.push);
    RVCR_CHECK_LAUNCH();
    count_launch(h);
    return RVCR_OK;
}

// Blend the queries of work items [item_begin, item_end) (after their scan wave) into `out` rows.
int gather_blend_range(rvcr_index* h, const float* feats, int64_t n, const float* D, const int32_t* P, double index_rate,
                       int item_begin, int item_end, float* out, int max_ctas, cudaStream_t s) {
    if (n <= 0) return RVCR_OK;
    const float r32 = (float)index_rate;
    const float omr32 = (float)(1.0 - index_rate);
    int threads = ((h->d / 4 + 31) / 32) * 32;
    if (threads > 256) threads = 256;
    if (threads < 32) threads = 32;
    int64_t grid = max_ctas > 0 ? max_ctas : (int64_t)h->num_sms * (2048 / threads);
    gather_blend_kernel<0><<<(unsigned)grid, threads, 0, s>>>(const_cast<float*>(feats), n, h->d, D, P, nullptr, h->own_lo, h->own_hi,
                                                            h->vecs, h->last_row_pos, r32, omr32, h->ws_vals2.as<int32_t>(),
                                                            h->ws_items.as<int32_t>(), h->ws_misc.as<int32_t>(), item_begin,
                                                            item_end, out, PostArgs{}, PushArgs{});
    RVCR_CHECK_LAUNCH();
    count_launch(h);
    return RVCR_OK;
}

// ---------------------------------------------------------------------- shard merges --------

// dist_all / list_all: G x n.  Global argmin by (distance, list id); a shard with no valid
// centroid for a query reports list -1.
__global__ void merge_coarse_kernel(const float* __restrict__ dist_all, const int32_t* __restrict__ list_all, int G,
                                    int64_t n, int32_t* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float bd = RVCR_FLT_MAX;
    int32_t bl = -1;
    for (int g = 0; g < G; ++g) {
        const int32_t li = list_all[(int64_t)g * n + i];
        const float di = dist_all[(int64_t)g * n + i];
        if (li < 0) continue;
        if (bl < 0 || di < bd || (di == bd && li < bl)) {
            bd = di;
            bl = li;
        }
    }
    out[i] = bl;
}

int merge_coarse(const float* dist_all, const int32_t* list_all, int G, int64_t n, int32_t* out, cudaStream_t s) {
    if (n <= 0) return RVCR_OK;
    merge_coarse_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dist_all, list_all, G, n, out);
    RVCR_CHECK_LAUNCH();
    return RVCR_OK;
}

// D_all / I_all: G x n x k, each shard's row sorted ascending and padded with (FLT_MAX, -1).
// One thread per query does a G-way merge by (distance, id); with nprobe = 1 exactly one shard
// holds real candidates, but the merge is written for the general case.
__global__ void merge_topk_kernel(const float* __restrict__ D_all, const int64_t* __restrict__ I_all, int G, int64_t n,
                                  int k, float* __restrict__ D, int64_t* __restrict__ I) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int head[16];
    for (int g = 0; g < G; ++g) head[g] = 0;
    for (int j = 0; j < k; ++j) {
        int bg = -1;
        float bd = RVCR_FLT_MAX;
        int64_t bi = -1;
        for (int g = 0; g < G; ++g) {
            if (head[g] >= k) continue;
            const int64_t o = ((int64_t)g * n + i) * k + head[g];
            const int64_t id = I_all[o];
            if (id < 0) continue;  // padding: nothing further in this shard
            const float dd = D_all[o];
            if (bg < 0 || dd < bd || (dd == bd && id < bi)) {
                bg = g;
                bd = dd;
                bi = id;
            }
        }
        if (bg >= 0) head[bg]++;
        D[i * k + j] = bg >= 0 ? bd : RVCR_FLT_MAX;
        I[i * k + j] = bg >= 0 ? bi : -1;
    }
}

int merge_topk(const float* D_all, const int64_t* I_all, int G, int64_t n, int k, float* D, int64_t* I, cudaStream_t s) {
    if (n <= 0) return RVCR_OK;
    if (G > 16) {
        set_error("merge_topk supports at most 16 shards");
        return RVCR_ERR_UNSUPPORTED;
    }
    merge_topk_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(D_all, I_all, G, n, k, D, I);
    RVCR_CHECK_LAUNCH();
    return RVCR_OK;
}

}  // namespace rvcr
