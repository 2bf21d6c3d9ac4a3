// FP32 pipe micro-peaks on B200: scalar FFMA, packed FFMA2 (fma.rn.f32x2), and the FSUB+FFMA mix
// the canonical-distance kernels issue.  Prints TFLOP/s (2 flops per FMA lane-op) and lane-ops/s.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
    if (MODE == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = __fmaf_rn(x[i], a, b);
        }
    } else if (MODE == 1) {
        unsigned long long p[8], pa, pb;
        asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a), "f"(a));
        asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b), "f"(b));
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(x[2 * i]), "=f"(x[2 * i + 1]) : "l"(p[i]));
    } else if (MODE == 2) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float t = __fsub_rn(a, x[i] * 0.f + b);  // keep a real FSUB: t depends on loop-carried value
                x[i] = __fmaf_rn(t, t, x[i]);
            }
        }
    } else {
        unsigned long long p[8], pa;
        asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a), "f"(b));
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                unsigned long long t;
                asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(pa), "l"(p[i]));
                asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(p[i]) : "l"(t));
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(x[2 * i]), "=f"(x[2 * i + 1]) : "l"(p[i]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double lane_ops_per_iter) {
    int sms = 148, iters = 20000;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int grid = sms * 8;
    float* out;
    cudaMalloc(&out, grid * 256 * 4);
    k<MODE><<<grid, 256>>>(out, 100, 1.0001f, 0.5f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k<MODE><<<grid, 256>>>(out, iters, 1.0001f, 0.5f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double ops = (double)grid * 256 * iters * lane_ops_per_iter;
    printf("{\"probe\": \"%s\", \"ms\": %.3f, \"fp32_lane_ops_per_s\": %.4g, \"per_sm_per_clk_at_1965MHz\": %.1f}\n", name, best,
           ops / (best * 1e-3), ops / (best * 1e-3) / sms / 1.965e9);
    cudaFree(out);
}

int main() {
    run<0>("ffma_scalar (16 FFMA/iter)", 16);
    run<1>("ffma2_packed (8 FFMA2 = 16 lane-FMA/iter)", 16);
    run<2>("fsub+ffma scalar (16 pairs/iter, counted as 32 lane-ops)", 32);
    run<3>("fsub2+ffma2 packed (8+8 packed = 32 lane-ops/iter)", 32);
    return 0;
}
