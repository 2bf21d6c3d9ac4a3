// peer.cu -- peer-memory plumbing of the sharded path: one process per GPU, every rank's feature buffer is a plain
// cudaMalloc allocation exported with cudaIpcGetMemHandle and mapped into every other rank (NVLink / NVSwitch peer
// access), so the two bulk exchanges of a sharded retrieve_blend step need no collective:
//
//   dispatch   the owner shard PULLS the query rows assigned to its lists straight out of their home ranks' buffers
//              (peer_pull_rows_kernel: 3 KB coalesced row reads over NVLink -> a contiguous local batch);
//   combine    the blend epilogue PUSHES each blended row straight into its home rank's buffer (PushArgs, common.cuh).
//
// The reference has no multi-GPU retrieval; this is the B200-native widening BASELINE.json's north_star
// names.  The NCCL all-to-all path (sharded.py) stays as the transport when peer mapping is unavailable.
#include "common.cuh"

namespace rvcr {

// one CTA per row (grid-stride), thread -> float4 column
__global__ void __launch_bounds__(256)
peer_pull_rows_kernel(const float* const* __restrict__ bases, const int32_t* __restrict__ src_rank,
                      const int32_t* __restrict__ src_row, int64_t m, int d, float* __restrict__ out) {
    const int nvec = d >> 2;
    for (int64_t r = blockIdx.x; r < m; r += gridDim.x) {
        const float4* src = reinterpret_cast<const float4*>(bases[src_rank[r]] + (int64_t)src_row[r] * d);
        float4* dst = reinterpret_cast<float4*>(out + r * d);
        for (int c = threadIdx.x; c < nvec; c += blockDim.x) dst[c] = src[c];
    }
}

int peer_pull_rows(const float* const* bases, const int32_t* src_rank, const int32_t* src_row, int64_t m, int d, float* out,
                   cudaStream_t s) {
    if (m <= 0) return RVCR_OK;
    if (d % 4 != 0) {
        set_error("d=%d unsupported (must be a multiple of 4)", d);
        return RVCR_ERR_UNSUPPORTED;
    }
    int dev = 0, sms = 148;
    RVCR_CUDA(cudaGetDevice(&dev));
    RVCR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int threads = ((d / 4 + 31) / 32) * 32;
    if (threads > 256) threads = 256;
    int64_t grid = (int64_t)sms * (2048 / threads);
    if (grid > m) grid = m;
    peer_pull_rows_kernel<<<(unsigned)grid, threads, 0, s>>>(bases, src_rank, src_row, m, d, out);
    RVCR_CHECK_LAUNCH();
    return RVCR_OK;
}

}  // namespace rvcr
