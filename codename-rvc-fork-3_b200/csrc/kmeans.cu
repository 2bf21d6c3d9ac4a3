// kmeans.cu -- index.train(x): faiss Clustering::train as IndexIVF::train runs it, natively.
//
// Replaces `index_added.train(big_npy)` (rvc/train/process/extract_index.py:65) -> IndexIVF::train ->
// Level1Quantizer::train_q1 -> Clustering(d, nlist, cp).train(n, x, quantizer) [EXT]:
//
//   subsample   n > 256 * k : keep x[rand_perm(n, seed)[0 .. 256 k)]                  (max_points_per_centroid)
//   init        centroids = x[rand_perm(n, seed + 1)[0 .. k)]
//   niter x     assign   = quantizer.search(x, 1)     -> the SAME nearest-centroid kernels as the search path
//                                                        (tcgen05 bf16 + exact fp32 re-check; 2 n k d flop, all of
//                                                        the cost: 1.97 PFLOP for C4)
//               update   = per centroid, fp32 sum of its points IN POINT ORDER, times fl(1 / count)
//               split    = every empty cluster is re-seeded from a donor picked by RandomGenerator(1234) with
//                          probability (count - 1) / (n - k), copy perturbed by 1 +- 1/1024 alternating per dimension
//
// The update is deterministic and ordered like faiss's compute_centroids (one thread per (centroid, float4 column)
// walks the centroid's points in ascending point index), so the whole trajectory is a pure function of the input:
// it matches the oracle restatement bit for bit (tests/test_gpu_build_flow.py) -- atomics would not.
// std::mt19937 is faiss's RandomGenerator; rand_perm / rand_float are restated from faiss utils/random.cpp.
#include <cub/device/device_radix_sort.cuh>

#include <string.h>

#include <algorithm>
#include <random>

#include "common.cuh"

namespace rvcr {

int coarse_dispatch_for_train(rvcr_index* h, const float* q, int64_t n, int32_t* list_out, cudaStream_t s);   // api.cu

namespace {

// faiss utils/random.cpp: rand_perm(perm, n, seed)
std::vector<int64_t> faiss_rand_perm(int64_t n, int64_t seed) {
    std::vector<int64_t> perm((size_t)n);
    for (int64_t i = 0; i < n; ++i) perm[(size_t)i] = i;
    std::mt19937 mt((unsigned int)seed);
    for (int64_t i = 0; i + 1 < n; ++i) {
        const int64_t i2 = i + (int64_t)(mt() % (uint64_t)(n - i));
        std::swap(perm[(size_t)i], perm[(size_t)i2]);
    }
    return perm;
}

__global__ void gather_rows_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx, int64_t k, int d,
                                   float* __restrict__ out) {
    const int nvec = d >> 2;
    for (int64_t r = blockIdx.x; r < k; r += gridDim.x) {
        const float4* src = reinterpret_cast<const float4*>(x + idx[r] * d);
        float4* dst = reinterpret_cast<float4*>(out + r * d);
        for (int c = threadIdx.x; c < nvec; c += blockDim.x) dst[c] = src[c];
    }
}

// key = list id (k for an invalid assignment: a NaN row), value = point index; histogram of the keys
__global__ void train_keys_kernel(const int32_t* __restrict__ assign, int64_t n, uint32_t k, uint32_t* __restrict__ keys,
                                  int32_t* __restrict__ vals, unsigned int* __restrict__ counts) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t a = assign[i];
    const uint32_t key = (a >= 0 && (uint32_t)a < k) ? (uint32_t)a : k;
    keys[i] = key;
    vals[i] = (int32_t)i;
    if (key < k) atomicAdd(counts + key, 1u);
}

// One CTA per centroid, one thread per float4 column: sequential fp32 accumulation over the centroid's points in
// ascending point index (`order` comes out of a stable radix sort), then c = sum * fl(1 / count).  Four row loads are
// kept in flight; the adds stay in order.
__global__ void __launch_bounds__(256)
centroid_mean_kernel(const float* __restrict__ x, const int32_t* __restrict__ order, const int64_t* __restrict__ seg_off, int d,
                     float* __restrict__ cent) {
    const int64_t ci = blockIdx.x;
    const int64_t lo = seg_off[ci], hi = seg_off[ci + 1];
    if (hi == lo) return;   // empty cluster: split_clusters re-seeds it on the host
    const float norm = __fdiv_rn(1.0f, (float)(hi - lo));
    const int nvec = d >> 2;
    for (int c = threadIdx.x; c < nvec; c += blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int64_t p = lo;
        for (; p + 4 <= hi; p += 4) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(x + (int64_t)order[p + u] * d) + c);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc.x = __fadd_rn(acc.x, v[u].x);
                acc.y = __fadd_rn(acc.y, v[u].y);
                acc.z = __fadd_rn(acc.z, v[u].z);
                acc.w = __fadd_rn(acc.w, v[u].w);
            }
        }
        for (; p < hi; ++p) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(x + (int64_t)order[p] * d) + c);
            acc.x = __fadd_rn(acc.x, v.x);
            acc.y = __fadd_rn(acc.y, v.y);
            acc.z = __fadd_rn(acc.z, v.z);
            acc.w = __fadd_rn(acc.w, v.w);
        }
        float4 o;
        o.x = __fmul_rn(acc.x, norm);
        o.y = __fmul_rn(acc.y, norm);
        o.z = __fmul_rn(acc.z, norm);
        o.w = __fmul_rn(acc.w, norm);
        reinterpret_cast<float4*>(cent + ci * d)[c] = o;
    }
}

int ceil_log2_u64(uint64_t v) {
    int b = 0;
    while ((1ull << b) < v) ++b;
    return b;
}

// faiss Clustering.cpp split_clusters (k_frozen = 0): hassign is float, the generator is re-created per call
int split_clusters_host(int d, int64_t k, int64_t n, std::vector<float>& hassign, std::vector<float>& cent) {
    const double EPS = 1.0 / 1024.0;
    int nsplit = 0;
    std::mt19937 mt(1234u);
    for (int64_t ci = 0; ci < k; ++ci) {
        if (hassign[(size_t)ci] != 0.f) continue;
        int64_t cj = 0;
        for (;;) {
            const float p = (float)((hassign[(size_t)cj] - 1.0) / (float)(n - k));
            const float r = (float)mt() / (float)mt.max();
            if (r < p) break;
            cj = (cj + 1) % k;
        }
        float* a = cent.data() + (size_t)ci * d;
        float* b = cent.data() + (size_t)cj * d;
        memcpy(a, b, sizeof(float) * (size_t)d);
        for (int j = 0; j < d; ++j) {
            if (j % 2 == 0) {
                a[j] = (float)(a[j] * (1 + EPS));
                b[j] = (float)(b[j] * (1 - EPS));
            } else {
                a[j] = (float)(a[j] * (1 - EPS));
                b[j] = (float)(b[j] * (1 + EPS));
            }
        }
        hassign[(size_t)ci] = hassign[(size_t)cj] / 2;
        hassign[(size_t)cj] -= hassign[(size_t)ci];
        ++nsplit;
    }
    return nsplit;
}

}  // namespace

// Host-only pieces of the k-means, reachable from CPU tests (no CUDA call inside): see rvcr_debug_rand_perm /
// rvcr_debug_split_clusters in include/rvcr.h.
void kmeans_debug_rand_perm(int64_t n, int64_t seed, int64_t* out) {
    const std::vector<int64_t> p = faiss_rand_perm(n, seed);
    if (n > 0) memcpy(out, p.data(), sizeof(int64_t) * (size_t)n);
}
int kmeans_debug_split_clusters(int d, int64_t k, int64_t n, float* hassign, float* centroids) {
    std::vector<float> h(hassign, hassign + k), c(centroids, centroids + (size_t)k * d);
    const int ns = split_clusters_host(d, k, n, h, c);
    memcpy(hassign, h.data(), sizeof(float) * (size_t)k);
    memcpy(centroids, c.data(), sizeof(float) * (size_t)k * d);
    return ns;
}

// Caller holds the handle lock and has selected the device.
int kmeans_train(rvcr_index* h, const float* x_host, int64_t n, cudaStream_t s) {
    const int64_t k = h->nlist;
    const int d = h->d;
    const int64_t niter = h->opt.kmeans_niter, seed = 1234, max_per_centroid = 256;
    if (d % 4 != 0) {
        set_error("d=%d unsupported (must be a multiple of 4)", d);
        return RVCR_ERR_UNSUPPORTED;
    }
    if (n < k) {
        set_error("Number of training points (%lld) should be at least as large as number of clusters (%lld)", (long long)n,
                  (long long)k);
        return RVCR_ERR_INVALID;
    }
    // ---- subsample + upload (chunked through a pinned bounce buffer when rows are picked by a permutation) ----
    int64_t ns = n;
    std::vector<int64_t> sel;
    if (n > k * max_per_centroid) {
        sel = faiss_rand_perm(n, seed);
        ns = k * max_per_centroid;
        sel.resize((size_t)ns);
    }
    if (ns >= (1ll << 31)) {
        set_error("too many training points (%lld)", (long long)ns);
        return RVCR_ERR_UNSUPPORTED;
    }
    const size_t rowb = (size_t)d * sizeof(float);
    DevBuf xd, cent_new, assign, keys, vals, keys2, vals2, counts, seg, tmp, idx;
    bool success = false;
    auto cleanup = [&]() {
        // a failure part-way (an OOM in the coarse scratch, say) must not leave a "trained" index holding the initial or
        // half-iterated centroids: callers that catch the error (extract_index.py:74-78 does) would add / write garbage
        if (!success) {
            h->trained = false;
            coarse_tc_invalidate(h);
        }
        xd.release(); cent_new.release(); assign.release(); keys.release(); vals.release(); keys2.release(); vals2.release();
        counts.release(); seg.release(); tmp.release(); idx.release();
    };
#define KM_TRY(expr) do { int _r = (expr); if (_r != RVCR_OK) { cleanup(); return _r; } } while (0)
#define KM_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); cleanup(); return RVCR_ERR_CUDA; } } while (0)
    KM_TRY(xd.reserve((size_t)ns * rowb));
    if (sel.empty()) {
        KM_CUDA(cudaMemcpyAsync(xd.p, x_host, (size_t)ns * rowb, cudaMemcpyHostToDevice, s));
    } else {
        const int64_t chunk = std::max<int64_t>(1, (int64_t)((32u << 20) / rowb));
        char* pin[2] = {nullptr, nullptr};
        cudaEvent_t ev[2] = {nullptr, nullptr};
        bool ok = cudaMallocHost(&pin[0], (size_t)chunk * rowb) == cudaSuccess && cudaMallocHost(&pin[1], (size_t)chunk * rowb) == cudaSuccess &&
                  cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming) == cudaSuccess &&
                  cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming) == cudaSuccess;
        int cur = 0;
        bool used[2] = {false, false};
        for (int64_t r0 = 0; ok && r0 < ns; r0 += chunk) {
            const int64_t m = std::min(chunk, ns - r0);
            if (used[cur]) ok = cudaEventSynchronize(ev[cur]) == cudaSuccess;
            for (int64_t r = 0; ok && r < m; ++r) memcpy(pin[cur] + (size_t)r * rowb, x_host + sel[(size_t)(r0 + r)] * d, rowb);
            ok = ok && cudaMemcpyAsync(xd.as<char>() + (size_t)r0 * rowb, pin[cur], (size_t)m * rowb, cudaMemcpyHostToDevice, s) == cudaSuccess &&
                 cudaEventRecord(ev[cur], s) == cudaSuccess;
            used[cur] = true;
            cur ^= 1;
        }
        ok = ok && cudaStreamSynchronize(s) == cudaSuccess;
        for (int b = 0; b < 2; ++b) {
            if (pin[b]) cudaFreeHost(pin[b]);
            if (ev[b]) cudaEventDestroy(ev[b]);
        }
        if (!ok) {
            set_error("uploading the k-means training sample failed: %s", cudaGetErrorString(cudaGetLastError()));
            cleanup();
            return RVCR_ERR_CUDA;
        }
        std::vector<int64_t>().swap(sel);
    }
    const float* x = xd.as<float>();
    // ---- init: centroids = x[rand_perm(ns, seed + 1)[0 .. k)] ----
    if (!h->cent) KM_CUDA(cudaMalloc(&h->cent, (size_t)k * rowb));
    {
        std::vector<int64_t> perm = faiss_rand_perm(ns, seed + 1);
        KM_TRY(idx.reserve((size_t)k * 8));
        KM_CUDA(cudaMemcpyAsync(idx.p, perm.data(), (size_t)k * 8, cudaMemcpyHostToDevice, s));
        gather_rows_kernel<<<(unsigned)std::min<int64_t>(k, 4096), 192, 0, s>>>(x, idx.as<int64_t>(), k, d, h->cent);
        KM_CUDA(cudaGetLastError());
        KM_CUDA(cudaStreamSynchronize(s));   // perm goes out of scope
        count_launch(h);
    }
    h->trained = true;
    KM_TRY(coarse_tc_invalidate(h));
    KM_TRY(cent_new.reserve((size_t)k * rowb));
    KM_TRY(assign.reserve((size_t)ns * 4));
    KM_TRY(keys.reserve((size_t)ns * 4));
    KM_TRY(vals.reserve((size_t)ns * 4));
    KM_TRY(keys2.reserve((size_t)ns * 4));
    KM_TRY(vals2.reserve((size_t)ns * 4));
    KM_TRY(counts.reserve((size_t)k * 4));
    KM_TRY(seg.reserve((size_t)(k + 1) * 8));
    const int end_bit = ceil_log2_u64((uint64_t)k + 1);
    size_t tb = 0;
    KM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys.as<uint32_t>(), keys2.as<uint32_t>(), vals.as<int32_t>(), vals2.as<int32_t>(),
                                            (int)ns, 0, end_bit, s));
    KM_TRY(tmp.reserve(tb));
    std::vector<unsigned int> cnt((size_t)k);
    std::vector<int64_t> off((size_t)k + 1);
    std::vector<float> hassign((size_t)k), cent_host;
    const int64_t assign_chunk = 1 << 20;   // bounds the coarse kernels' per-call scratch
    for (int64_t it = 0; it < niter; ++it) {
        for (int64_t a = 0; a < ns; a += assign_chunk) {
            const int64_t m = std::min(assign_chunk, ns - a);
            KM_TRY(coarse_dispatch_for_train(h, x + a * d, m, assign.as<int32_t>() + a, s));
        }
        KM_CUDA(cudaMemsetAsync(counts.p, 0, (size_t)k * 4, s));
        train_keys_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, s>>>(assign.as<int32_t>(), ns, (uint32_t)k, keys.as<uint32_t>(),
                                                                        vals.as<int32_t>(), counts.as<unsigned int>());
        KM_CUDA(cudaGetLastError());
        KM_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys.as<uint32_t>(), keys2.as<uint32_t>(), vals.as<int32_t>(), vals2.as<int32_t>(),
                                                (int)ns, 0, end_bit, s));
        KM_CUDA(cudaMemcpyAsync(cnt.data(), counts.p, (size_t)k * 4, cudaMemcpyDeviceToHost, s));
        KM_CUDA(cudaStreamSynchronize(s));
        bool any_empty = false;
        off[0] = 0;
        for (int64_t c = 0; c < k; ++c) {
            off[(size_t)c + 1] = off[(size_t)c] + cnt[(size_t)c];
            hassign[(size_t)c] = (float)cnt[(size_t)c];
            any_empty = any_empty || cnt[(size_t)c] == 0;
        }
        KM_CUDA(cudaMemcpyAsync(seg.p, off.data(), (size_t)(k + 1) * 8, cudaMemcpyHostToDevice, s));
        // empty clusters keep their previous centroid until split_clusters overwrites them
        KM_CUDA(cudaMemcpyAsync(cent_new.p, h->cent, (size_t)k * rowb, cudaMemcpyDeviceToDevice, s));
        const int threads = std::min(256, std::max(32, ((d / 4 + 31) / 32) * 32));
        centroid_mean_kernel<<<(unsigned)k, threads, 0, s>>>(x, vals2.as<int32_t>(), seg.as<int64_t>(), d, cent_new.as<float>());
        KM_CUDA(cudaGetLastError());
        count_launch(h, 3);
        if (any_empty) {
            cent_host.resize((size_t)k * d);
            KM_CUDA(cudaMemcpyAsync(cent_host.data(), cent_new.p, (size_t)k * rowb, cudaMemcpyDeviceToHost, s));
            KM_CUDA(cudaStreamSynchronize(s));
            split_clusters_host(d, k, ns, hassign, cent_host);
            KM_CUDA(cudaMemcpyAsync(h->cent, cent_host.data(), (size_t)k * rowb, cudaMemcpyHostToDevice, s));
        } else {
            KM_CUDA(cudaMemcpyAsync(h->cent, cent_new.p, (size_t)k * rowb, cudaMemcpyDeviceToDevice, s));
        }
        KM_CUDA(cudaStreamSynchronize(s));
        KM_TRY(coarse_tc_invalidate(h));   // the bf16 operand copy of the centroids is stale
    }
    success = true;
    cleanup();
#undef KM_TRY
#undef KM_CUDA
    return RVCR_OK;
}

}  // namespace rvcr
