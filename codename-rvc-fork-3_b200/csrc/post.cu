// post.cu -- the step right after the retrieval path, as one pass.
//
// Replaces rvc/infer/pipeline.py:463-481:
//     feats  = F.interpolate(feats.permute(0, 2, 1),  scale_factor=2).permute(0, 2, 1)      # nearest: out[t] = in[t // 2]
//     feats0 = F.interpolate(feats0.permute(0, 2, 1), scale_factor=2).permute(0, 2, 1)      # (pitch guidance only)
//     pitchff = pitchf.clone(); pitchff[pitchf > 0] = 1; pitchff[pitchf < 1] = protect       # (protect < 0.5 only)
//     feats = feats * pitchff.unsqueeze(-1) + feats0 * (1 - pitchff.unsqueeze(-1))
// The reference runs two permute + interpolate passes, four elementwise kernels and keeps three (2n, d) temporaries:
// 18 row-passes of HBM traffic per input row; this kernel reads each input row once and writes the two output rows:
// 4 row-passes (2 without the protect mix).  When retrieval ran, rvcr_retrieve_blend_upsample does the same inside the
// blend epilogue (post_store in common.cuh) and this kernel is not launched at all.
//
// Bit-exactness contract: fl(fl(feats * pf) + fl(feats0 * fl(1 - pf))), no FMA, pf as protect_factor() builds it.
#include "common.cuh"

namespace rvcr {

// one CTA per group of input rows; thread -> float4 column (scalar tail for d % 4 != 0 handled by VEC = 1)
template <int VEC>
__global__ void __launch_bounds__(256)
upsample_protect_kernel(const float* __restrict__ feats, const float* __restrict__ feats0, const float* __restrict__ pitchf,
                        int64_t n, int d, float protect, float* __restrict__ out) {
    const int ncol = d / VEC;
    for (int64_t qi = blockIdx.x; qi < n; qi += gridDim.x) {
        float pf0 = 1.f, pf1 = 1.f;
        const bool mix = pitchf != nullptr && feats0 != nullptr;
        if (mix) {
            pf0 = protect_factor(__ldg(pitchf + 2 * qi), protect);
            pf1 = protect_factor(__ldg(pitchf + 2 * qi + 1), protect);
        }
        for (int c = threadIdx.x; c < ncol; c += blockDim.x) {
            if (VEC == 4) {
                const float4 o = __ldg(reinterpret_cast<const float4*>(feats + qi * d) + c);
                float4 r0 = o, r1 = o;
                if (mix) {
                    const float4 f0 = __ldg(reinterpret_cast<const float4*>(feats0 + qi * d) + c);
                    r0 = protect_mix(o, f0, pf0);
                    r1 = protect_mix(o, f0, pf1);
                }
                reinterpret_cast<float4*>(out + (2 * qi) * d)[c] = r0;
                reinterpret_cast<float4*>(out + (2 * qi + 1) * d)[c] = r1;
            } else {
                const float o = __ldg(feats + qi * d + c);
                float r0 = o, r1 = o;
                if (mix) {
                    const float f0 = __ldg(feats0 + qi * d + c);
                    r0 = __fadd_rn(__fmul_rn(o, pf0), __fmul_rn(f0, __fsub_rn(1.0f, pf0)));
                    r1 = __fadd_rn(__fmul_rn(o, pf1), __fmul_rn(f0, __fsub_rn(1.0f, pf1)));
                }
                out[(2 * qi) * d + c] = r0;
                out[(2 * qi + 1) * d + c] = r1;
            }
        }
    }
}

int upsample_protect(const float* feats, const float* feats0, const float* pitchf, int64_t n, int d, double protect, float* out,
                     cudaStream_t s) {
    if (n <= 0) return RVCR_OK;
    int dev = 0, sms = 148;
    RVCR_CUDA(cudaGetDevice(&dev));
    RVCR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const bool vec = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(feats) | reinterpret_cast<uintptr_t>(out) |
                                       reinterpret_cast<uintptr_t>(feats0)) % 16 == 0);
    const int ncol = vec ? d / 4 : d;
    int threads = ((ncol + 31) / 32) * 32;
    if (threads > 256) threads = 256;
    int64_t grid = (int64_t)sms * (2048 / threads);
    if (grid > n) grid = n;
    if (vec)
        upsample_protect_kernel<4><<<(unsigned)grid, threads, 0, s>>>(feats, feats0, pitchf, n, d, (float)protect, out);
    else
        upsample_protect_kernel<1><<<(unsigned)grid, threads, 0, s>>>(feats, feats0, pitchf, n, d, (float)protect, out);
    RVCR_CHECK_LAUNCH();
    return RVCR_OK;
}

}  // namespace rvcr
