// peer.cu -- peer-memory plumbing of the sharded path: one process per GPU, every rank's feature buffer is a plain
// cudaMalloc allocation exported with cudaIpcGetMemHandle and mapped into every other rank (NVLink / NVSwitch peer
// access), so the two bulk exchanges of a sharded retrieve_blend step need no collective:
//
//   dispatch   the owner shard PULLS the query rows assigned to its lists straight out of their home ranks' buffers
//              (peer_pull_rows_kernel: 3 KB coalesced row reads over NVLink -> a contiguous local batch);
//   combine    the blend epilogue PUSHES each blended row straight into its home rank's buffer (PushArgs, common.cuh).
//
// Round 2: the step is device-driven.  Round 1 found out on the host who owns what (two NCCL all-gathers, a .tolist()
// and a nonzero(): two host synchronisations per step, 3 ms of a 20.7 ms step at 8 GPUs).  Now every rank owns an INBOX in
// peer-mapped memory (header: entry count + one arrival flag per rank; then `cap` entries of list id / home rank / home
// row).  After its coarse step the home rank's dispatch kernel files each query into the inbox of the shard that owns its
// list (one remote atomicAdd per CTA and owner reserves the slots, plain NVLink stores fill them); a flag barrier kernel
// (release store to every peer, acquire spin on the own flags) says "everyone has dispatched"; the owner seals the unused
// tail of its inbox, pulls the rows, scans + blends + pushes; a second flag barrier says "every row is home".  Nothing
// returns to the host: counts stay on the device (kernels are sized by the inbox capacity and skip the sealed tail).
//
// The reference has no multi-GPU retrieval; this is the B200-native widening BASELINE.json's north_star
// names.  The NCCL all-to-all path (sharded.py) stays as the transport when peer mapping is unavailable.
#include "common.cuh"

namespace rvcr {

// one CTA per row (grid-stride), thread -> float4 column; m_dev (optional) = device-side row count (<= m)
__global__ void __launch_bounds__(256)
peer_pull_rows_kernel(const float* const* __restrict__ bases, const int32_t* __restrict__ src_rank,
                      const int32_t* __restrict__ src_row, int64_t m, const int32_t* __restrict__ m_dev, int d,
                      float* __restrict__ out) {
    const int nvec = d >> 2;
    if (m_dev) {
        const int64_t md = *m_dev;
        if (md < m) m = md;
    }
    for (int64_t r = blockIdx.x; r < m; r += gridDim.x) {
        const float4* src = reinterpret_cast<const float4*>(bases[src_rank[r]] + (int64_t)src_row[r] * d);
        float4* dst = reinterpret_cast<float4*>(out + r * d);
        for (int c = threadIdx.x; c < nvec; c += blockDim.x) dst[c] = src[c];
    }
}

int peer_pull_rows(const float* const* bases, const int32_t* src_rank, const int32_t* src_row, int64_t m, int d, float* out,
                   cudaStream_t s, const int32_t* m_dev) {
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
    peer_pull_rows_kernel<<<(unsigned)grid, threads, 0, s>>>(bases, src_rank, src_row, m, m_dev, d, out);
    RVCR_CHECK_LAUNCH();
    return RVCR_OK;
}

// ------------------------------------------------------------------------------ inbox -------

// Home rank: file query i (list assign[i], row row0 + i of this rank's feature buffer) into the inbox of the shard that
// owns the list.  bounds[r - 1] = first list of rank r (r = 1 .. G - 1); an invalid assignment (-1: NaN query) goes to
// shard 0, whose scan answers it with the all-pad / NaN row the reference produces.
__global__ void __launch_bounds__(256)
peer_dispatch_kernel(const int32_t* __restrict__ assign, int64_t n, int32_t row0, int32_t my_rank, int G,
                     const int32_t* __restrict__ bounds, unsigned char* const* __restrict__ inbox_bases, int64_t cap) {
    constexpr int PER = 4;
    __shared__ int hist[16];
    __shared__ int base[16];
    const int tid = threadIdx.x;
    if (tid < 16) hist[tid] = 0;
    __syncthreads();
    int own[PER], slot[PER];
    int32_t a[PER];
    const int64_t i0 = (int64_t)blockIdx.x * (256 * PER);
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int64_t i = i0 + tid + k * 256;
        own[k] = -1;
        if (i < n) {
            a[k] = assign[i];
            int o = 0;
            for (int r = 1; r < G; ++r) o += (a[k] >= bounds[r - 1]) ? 1 : 0;
            own[k] = a[k] < 0 ? 0 : o;
            slot[k] = atomicAdd(&hist[own[k]], 1);
        }
    }
    __syncthreads();
    if (tid < G && hist[tid] > 0) {
        // one remote atomic per (CTA, owner) reserves the range; over NVLink for tid != my_rank
        int* cnt = reinterpret_cast<int*>(inbox_bases[tid] + PEER_INBOX_COUNT_OFF);
        base[tid] = atomicAdd_system(cnt, hist[tid]);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        if (own[k] < 0) continue;
        const int64_t e = (int64_t)base[own[k]] + slot[k];
        if (e >= cap) continue;            // cannot happen when cap >= the total number of queries of a step
        unsigned char* ib = inbox_bases[own[k]];
        reinterpret_cast<int32_t*>(ib + PEER_INBOX_HEADER)[e] = a[k];
        reinterpret_cast<int32_t*>(ib + PEER_INBOX_HEADER + 4 * cap)[e] = my_rank;
        reinterpret_cast<int32_t*>(ib + PEER_INBOX_HEADER + 8 * cap)[e] = row0 + (int32_t)(i0 + tid + k * 256);
    }
}

// Cross-rank barrier on the launching streams: every rank releases `epoch` into its flag slot on every peer, then spins
// (acquire) until all of its own slots show it.  Everything the rank's earlier kernels wrote to peer memory is visible to a
// peer once that peer has seen the flag.  The spin is bounded: a rank that never arrives traps instead of hanging the GPU.
__global__ void peer_barrier_kernel(unsigned char* const* __restrict__ inbox_bases, int32_t my_rank, int G, int32_t epoch) {
    const int t = threadIdx.x;
    __threadfence_system();
    if (t < G) {
        int* flag = reinterpret_cast<int*>(inbox_bases[t] + PEER_INBOX_FLAGS_OFF) + my_rank;
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
    }
    if (t < G) {
        const int* mine = reinterpret_cast<const int*>(inbox_bases[my_rank] + PEER_INBOX_FLAGS_OFF) + t;
        int v;
        long long spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
            if (v - epoch >= 0) break;           // wrap-safe comparison
            __nanosleep(64);
            if (++spins > (1ll << 27)) __trap();   // ~10 s
        } while (true);
    }
    __threadfence_system();
}

// Owner: entries [count, cap) of the own inbox get a list id no shard owns, so the scan forms no work for them.
__global__ void peer_inbox_seal_kernel(int32_t* __restrict__ assign, const int32_t* __restrict__ count, int64_t cap) {
    const int64_t c = min((int64_t)*count, cap);
    for (int64_t i = c + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < cap; i += (int64_t)gridDim.x * blockDim.x)
        assign[i] = 0x7fffffff;
}

int peer_dispatch(const int32_t* assign, int64_t n, int64_t row0, int my_rank, int G, const int32_t* bounds,
                  unsigned char* const* inbox_bases, int64_t cap, cudaStream_t s) {
    if (n <= 0) return RVCR_OK;
    if (G < 1 || G > 16) {
        set_error("peer dispatch supports 1..16 ranks (got %d)", G);
        return RVCR_ERR_UNSUPPORTED;
    }
    if (row0 < 0 || row0 + n >= (1ll << 31)) {
        set_error("peer dispatch: rows [%lld, %lld) do not fit 31 bits", (long long)row0, (long long)(row0 + n));
        return RVCR_ERR_INVALID;
    }
    peer_dispatch_kernel<<<(unsigned)((n + 1023) / 1024), 256, 0, s>>>(assign, n, (int32_t)row0, my_rank, G, bounds, inbox_bases, cap);
    RVCR_CHECK_LAUNCH();
    return RVCR_OK;
}

int peer_barrier(unsigned char* const* inbox_bases, int my_rank, int G, int epoch, cudaStream_t s) {
    if (G < 1 || G > 16) {
        set_error("peer barrier supports 1..16 ranks (got %d)", G);
        return RVCR_ERR_UNSUPPORTED;
    }
    peer_barrier_kernel<<<1, 32, 0, s>>>(inbox_bases, my_rank, G, epoch);
    RVCR_CHECK_LAUNCH();
    return RVCR_OK;
}

int peer_inbox_seal(int32_t* assign, const int32_t* count, int64_t cap, cudaStream_t s) {
    peer_inbox_seal_kernel<<<148, 256, 0, s>>>(assign, count, cap);
    RVCR_CHECK_LAUNCH();
    return RVCR_OK;
}

}  // namespace rvcr
