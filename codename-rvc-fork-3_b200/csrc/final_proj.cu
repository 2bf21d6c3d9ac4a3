// final_proj.cu -- the step right BEFORE the retrieval path for v1 models, as the query producer.
//
// Replaces rvc/infer/pipeline.py:451-453:
//     feats = model.final_proj(feats[0]).unsqueeze(0) if version == "v1" else feats
// `final_proj` is the encoder's nn.Linear(768, 256): out = x W^T + b per frame; its 256-d output rows are the queries of a v1
// index (BASELINE config C2).  The reference runs it as a cuBLAS sgemm (fp32, TF32 off by default) followed by the
// retrieval's D2H copy; here it is one fp32 FFMA kernel on the caller's stream whose output feeds rvcr_retrieve_blend
// directly (rvcr_final_proj_retrieve_blend: projection + search + gather + weights + blend in one call, nothing returns
// to the host in between).
//
// Numerics: fp32 products accumulated with fused multiply-adds in ascending k, bias added last.  A GEMM library sums in a
// different (unspecified, blocked) order, so this step is compared with the reference statement within 1e-5 relative
// (tests/test_gpu_final_proj.py), not bit for bit; everything after it is bit-exact given the projected rows.
//
// Tiling: CTA = 64 frames x 64 outputs, 256 threads, 4 x 4 accumulators per thread, k in steps of 16 through shared
// memory (both operands K-major in global memory, transposed on the way in so the inner loop reads conflict-free float4s).
#include "common.cuh"

namespace rvcr {

constexpr int FP_BM = 64, FP_BN = 64, FP_BK = 16;

__global__ void __launch_bounds__(256)
final_proj_kernel(const float* __restrict__ x, int64_t n, int din, const float* __restrict__ w, const float* __restrict__ b,
                  int dout, float* __restrict__ out) {
    __shared__ __align__(16) float sx[FP_BK][FP_BM + 4];
    __shared__ __align__(16) float sw[FP_BK][FP_BN + 4];
    const int tid = threadIdx.x;
    const int tr = (tid >> 4) * 4, tc = (tid & 15) * 4;          // this thread's 4 x 4 block inside the tile
    const int64_t m0 = (int64_t)blockIdx.x * FP_BM;
    const int n0 = blockIdx.y * FP_BN;
    const int lr = tid >> 2, lk = (tid & 3) * 4;                 // loader: row lr (0..63), k offset lk (0, 4, 8, 12)
    float acc[4][4] = {};
    for (int k0 = 0; k0 < din; k0 += FP_BK) {
        {
            const int64_t r = m0 + lr;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < n) v = __ldg(reinterpret_cast<const float4*>(x + r * din + k0 + lk));
            sx[lk + 0][lr] = v.x;
            sx[lk + 1][lr] = v.y;
            sx[lk + 2][lr] = v.z;
            sx[lk + 3][lr] = v.w;
            const float4 u = __ldg(reinterpret_cast<const float4*>(w + (int64_t)(n0 + lr) * din + k0 + lk));
            sw[lk + 0][lr] = u.x;
            sw[lk + 1][lr] = u.y;
            sw[lk + 2][lr] = u.z;
            sw[lk + 3][lr] = u.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < FP_BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&sx[k][tr]);
            const float4 c = *reinterpret_cast<const float4*>(&sw[k][tc]);
            const float av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(av[i], cv[j], acc[i][j]);
        }
        __syncthreads();
    }
    const float4 bb = b ? __ldg(reinterpret_cast<const float4*>(b + n0 + tc)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = m0 + tr + i;
        if (r < n)
            *reinterpret_cast<float4*>(out + r * dout + n0 + tc) =
                make_float4(__fadd_rn(acc[i][0], bb.x), __fadd_rn(acc[i][1], bb.y), __fadd_rn(acc[i][2], bb.z), __fadd_rn(acc[i][3], bb.w));
    }
}

int final_proj(const float* x, int64_t n, int din, const float* w, const float* b, int dout, float* out, cudaStream_t s) {
    if (n <= 0) return RVCR_OK;
    if (din % FP_BK != 0 || dout % FP_BN != 0) {
        set_error("final_proj: d_in=%d must be a multiple of %d and d_out=%d a multiple of %d", din, FP_BK, dout, FP_BN);
        return RVCR_ERR_UNSUPPORTED;
    }
    dim3 grid((unsigned)((n + FP_BM - 1) / FP_BM), (unsigned)(dout / FP_BN));
    final_proj_kernel<<<grid, 256, 0, s>>>(x, n, din, w, b, dout, out);
    RVCR_CHECK_LAUNCH();
    return RVCR_OK;
}

}  // namespace rvcr
