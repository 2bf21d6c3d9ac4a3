// lists.cu -- inverted-list storage: append (index.add), reconstruct_n, derived metadata.
//
// Replaces IndexIVFFlat::add_core / ArrayInvertedLists::add_entry behind
// rvc/train/process/extract_index.py:67-69 and IndexIVF::reconstruct_n behind
// rvc/infer/pipeline.py:556.
//
// Layout in HBM: ONE list-ordered copy of the vectors, `vecs` (nstored x d fp32, list l occupies
// rows offsets[l] .. offsets[l+1]), `ids` (nstored x int64) and `offsets` (nlist+1 x int64).
// faiss keeps a std::vector pair per list and the reference adds a second full copy (big_npy);
// here the scan emits row positions, so the gather reads this same storage.
//
// index.add() appends to an id-ordered staging area; the first consumer (search, write,
// reconstruct_n) folds it into the list layout with one stable radix sort by list id, which keeps
// every list in ascending-id order exactly like faiss's sequential append.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace rvcr {

int DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return RVCR_OK;
    if (p) {
        cudaError_t e = cudaFree(p);
        p = nullptr;
        cap = 0;
        if (e != cudaSuccess) {
            set_error("cudaFree failed: %s", cudaGetErrorString(e));
            return RVCR_ERR_CUDA;
        }
    }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        p = nullptr;
        set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? RVCR_ERR_OOM : RVCR_ERR_CUDA;
    }
    cap = want;
    return RVCR_OK;
}

void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

// ------------------------------------------------------------------------------ kernels ----

__global__ void hist_lists_kernel(const int32_t* __restrict__ assign, int64_t n, int64_t own_lo, int64_t own_hi, int64_t nlist,
                                  unsigned long long* __restrict__ counts, uint32_t* __restrict__ keys,
                                  int32_t* __restrict__ vals) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t a = (int64_t)assign[i];
    const bool ok = a >= own_lo && a < own_hi;  // a shard keeps only the rows of the lists it owns
    if (ok) atomicAdd(&counts[a], 1ull);
    keys[i] = ok ? (uint32_t)a : (uint32_t)nlist;  // dropped rows sort last
    vals[i] = (int32_t)i;
}

// one warp per old row: move it to its slot in the grown layout
__global__ void move_old_rows_kernel(const float* __restrict__ old_vecs, const int64_t* __restrict__ old_ids,
                                     const int64_t* __restrict__ old_off, const int64_t* __restrict__ new_off,
                                     int64_t nold, int64_t nlist, int d, float* __restrict__ new_vecs,
                                     int64_t* __restrict__ new_ids) {
    const int lane = threadIdx.x & 31;
    int64_t p = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (p >= nold) return;
    // list of row p: last l with old_off[l] <= p
    int64_t lo = 0, hi = nlist;
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (old_off[mid] <= p) lo = mid; else hi = mid;
    }
    const int64_t dst = new_off[lo] + (p - old_off[lo]);
    const float4* src4 = reinterpret_cast<const float4*>(old_vecs + p * d);
    float4* dst4 = reinterpret_cast<float4*>(new_vecs + dst * d);
    for (int c = lane; c < (d >> 2); c += 32) dst4[c] = src4[c];
    if (lane == 0) new_ids[dst] = old_ids[p];
}

// one warp per sorted staged row: sorted position sp holds staged row vals[sp] with key keys[sp]
__global__ void place_new_rows_kernel(const float* __restrict__ stage_x, const uint32_t* __restrict__ keys,
                                      const int32_t* __restrict__ vals, int64_t nnew, uint32_t nlist,
                                      const int64_t* __restrict__ new_off, const int64_t* __restrict__ old_sizes,
                                      int64_t id0, int d, float* __restrict__ new_vecs, int64_t* __restrict__ new_ids) {
    const int lane = threadIdx.x & 31;
    int64_t sp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (sp >= nnew) return;
    const uint32_t k = keys[sp];
    if (k >= nlist) return;  // dropped (NaN vector or foreign list)
    int64_t lo = 0, hi = sp;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < k) lo = mid + 1; else hi = mid;
    }
    const int64_t dst = new_off[k] + old_sizes[k] + (sp - lo);
    const int64_t src = vals[sp];
    const float4* src4 = reinterpret_cast<const float4*>(stage_x + src * d);
    float4* dst4 = reinterpret_cast<float4*>(new_vecs + dst * d);
    for (int c = lane; c < (d >> 2); c += 32) dst4[c] = src4[c];
    if (lane == 0) new_ids[dst] = id0 + src;
}

__global__ void reconstruct_kernel(const float* __restrict__ vecs, const int64_t* __restrict__ ids, int64_t nstored,
                                   int d, int64_t i0, int64_t ni, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    int64_t p = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (p >= nstored) return;
    const int64_t id = ids[p] - i0;
    if (id < 0 || id >= ni) return;
    const float4* src4 = reinterpret_cast<const float4*>(vecs + p * d);
    float4* dst4 = reinterpret_cast<float4*>(out + id * d);
    for (int c = lane; c < (d >> 2); c += 32) dst4[c] = src4[c];
}

__global__ void find_id_kernel(const int64_t* __restrict__ ids, int64_t nstored, int64_t want, long long* __restrict__ pos) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p < nstored && ids[p] == want) *pos = (long long)p;
}

// -------------------------------------------------------------------------------- host -----

static int ceil_log2(uint64_t v) {
    int b = 0;
    while ((1ull << b) < v) ++b;
    return b;
}

int refresh_derived(rvcr_index* h, cudaStream_t s) {
    h->last_row_pos = -1;
    scan_tc_invalidate(h);
    if (h->nstored > 0 && h->ntotal > 0) {
        RVCR_TRY(h->ws_misc.reserve(64));
        long long* dpos = h->ws_misc.as<long long>() + 4;
        long long init = -1;
        RVCR_CUDA(cudaMemcpyAsync(dpos, &init, sizeof(init), cudaMemcpyHostToDevice, s));
        find_id_kernel<<<(unsigned)((h->nstored + 255) / 256), 256, 0, s>>>(h->ids, h->nstored, h->ntotal - 1, dpos);
        RVCR_CHECK_LAUNCH();
        long long got = -1;
        RVCR_CUDA(cudaMemcpyAsync(&got, dpos, sizeof(got), cudaMemcpyDeviceToHost, s));
        RVCR_CUDA(cudaStreamSynchronize(s));
        h->last_row_pos = got;
        count_launch(h);
    }
    return RVCR_OK;
}

// Fold the staged (id-ordered) adds into the list layout.
int finalize_pending(rvcr_index* h, cudaStream_t s) {
    if (h->npending == 0) return RVCR_OK;
    if (h->d % 4 != 0) {
        set_error("d=%d unsupported (must be a multiple of 4)", h->d);
        return RVCR_ERR_UNSUPPORTED;
    }
    const int64_t nnew = h->npending;
    if (nnew >= (1ll << 31)) {
        set_error("too many staged vectors (%lld)", (long long)nnew);
        return RVCR_ERR_UNSUPPORTED;
    }
    const int64_t nlist = h->nlist;
    const int d = h->d;
    const float* stage_x = h->stage_x;
    const int32_t* stage_assign = h->stage_assign;
    const int64_t id0 = h->ntotal - nnew;

    DevBuf counts, keys, vals, keys2, vals2, tmp, new_off_d, old_sizes_d;
    int rc = RVCR_OK;
    auto cleanup = [&]() {
        counts.release(); keys.release(); vals.release(); keys2.release(); vals2.release(); tmp.release();
        new_off_d.release(); old_sizes_d.release();
    };
#define FIN_TRY(expr) do { rc = (expr); if (rc != RVCR_OK) { cleanup(); return rc; } } while (0)
#define FIN_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_error("%s failed: %s", #expr, cudaGetErrorString(_e)); cleanup(); return RVCR_ERR_CUDA; } } while (0)

    FIN_TRY(counts.reserve((size_t)(nlist + 1) * 8));
    FIN_TRY(keys.reserve((size_t)nnew * 4));
    FIN_TRY(vals.reserve((size_t)nnew * 4));
    FIN_TRY(keys2.reserve((size_t)nnew * 4));
    FIN_TRY(vals2.reserve((size_t)nnew * 4));
    FIN_CUDA(cudaMemsetAsync(counts.p, 0, (size_t)(nlist + 1) * 8, s));
    hist_lists_kernel<<<(unsigned)((nnew + 255) / 256), 256, 0, s>>>(stage_assign, nnew, h->own_lo, h->own_hi, nlist,
                                                                      counts.as<unsigned long long>(), keys.as<uint32_t>(),
                                                                      vals.as<int32_t>());
    FIN_CUDA(cudaGetLastError());
    std::vector<int64_t> cnt((size_t)nlist);
    FIN_CUDA(cudaMemcpyAsync(cnt.data(), counts.p, (size_t)nlist * 8, cudaMemcpyDeviceToHost, s));
    const int end_bit = ceil_log2((uint64_t)nlist + 1);
    size_t tb = 0;
    FIN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys.as<uint32_t>(), keys2.as<uint32_t>(), vals.as<int32_t>(),
                                             vals2.as<int32_t>(), (int)nnew, 0, end_bit, s));
    FIN_TRY(tmp.reserve(tb));
    FIN_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys.as<uint32_t>(), keys2.as<uint32_t>(), vals.as<int32_t>(),
                                             vals2.as<int32_t>(), (int)nnew, 0, end_bit, s));
    FIN_CUDA(cudaStreamSynchronize(s));

    if (h->offsets_host.size() != (size_t)nlist + 1) h->offsets_host.assign((size_t)nlist + 1, 0);
    std::vector<int64_t> old_sizes((size_t)nlist), new_off((size_t)nlist + 1, 0);
    for (int64_t l = 0; l < nlist; ++l) {
        old_sizes[l] = h->offsets_host[l + 1] - h->offsets_host[l];
        new_off[l + 1] = new_off[l] + old_sizes[l] + cnt[l];
    }
    const int64_t nstored_new = new_off[nlist];
    FIN_TRY(new_off_d.reserve((size_t)(nlist + 1) * 8));
    FIN_TRY(old_sizes_d.reserve((size_t)nlist * 8 + 8));
    FIN_CUDA(cudaMemcpyAsync(new_off_d.p, new_off.data(), (size_t)(nlist + 1) * 8, cudaMemcpyHostToDevice, s));
    FIN_CUDA(cudaMemcpyAsync(old_sizes_d.p, old_sizes.data(), (size_t)nlist * 8, cudaMemcpyHostToDevice, s));

    float* new_vecs = nullptr;
    int64_t* new_ids = nullptr;
    if (nstored_new > 0) {
        cudaError_t e = cudaMalloc(&new_vecs, (size_t)nstored_new * d * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&new_ids, (size_t)nstored_new * 8);
        if (e != cudaSuccess) {
            if (new_vecs) cudaFree(new_vecs);
            set_error("allocating %lld x %d list storage failed: %s", (long long)nstored_new, d, cudaGetErrorString(e));
            cleanup();
            return RVCR_ERR_OOM;
        }
    }
    int launches = 2;
    if (h->nstored > 0) {
        const int64_t thr = h->nstored * 32;
        move_old_rows_kernel<<<(unsigned)((thr + 255) / 256), 256, 0, s>>>(h->vecs, h->ids, h->offsets, new_off_d.as<int64_t>(),
                                                                            h->nstored, nlist, d, new_vecs, new_ids);
        FIN_CUDA(cudaGetLastError());
        ++launches;
    }
    {
        const int64_t thr = nnew * 32;
        place_new_rows_kernel<<<(unsigned)((thr + 255) / 256), 256, 0, s>>>(stage_x, keys2.as<uint32_t>(), vals2.as<int32_t>(),
                                                                             nnew, (uint32_t)nlist, new_off_d.as<int64_t>(),
                                                                             old_sizes_d.as<int64_t>(), id0, d, new_vecs, new_ids);
        FIN_CUDA(cudaGetLastError());
        ++launches;
    }
    FIN_CUDA(cudaStreamSynchronize(s));
    count_launch(h, launches);

    if (h->vecs) cudaFree(h->vecs);
    if (h->ids) cudaFree(h->ids);
    h->vecs = new_vecs;
    h->ids = new_ids;
    h->nstored = nstored_new;
    h->offsets_host = new_off;
    if (!h->offsets) FIN_CUDA(cudaMalloc(&h->offsets, (size_t)(nlist + 1) * 8));
    FIN_CUDA(cudaMemcpyAsync(h->offsets, new_off.data(), (size_t)(nlist + 1) * 8, cudaMemcpyHostToDevice, s));
    FIN_CUDA(cudaStreamSynchronize(s));
    if (h->stage_x) cudaFree(h->stage_x);
    if (h->stage_assign) cudaFree(h->stage_assign);
    h->stage_x = nullptr;
    h->stage_assign = nullptr;
    h->stage_cap = 0;
    h->npending = 0;
    cleanup();
#undef FIN_TRY
#undef FIN_CUDA
    return refresh_derived(h, s);
}

int reconstruct(rvcr_index* h, int64_t i0, int64_t ni, float* out_dev, cudaStream_t s) {
    if (ni <= 0) return RVCR_OK;
    RVCR_CUDA(cudaMemsetAsync(out_dev, 0, (size_t)ni * h->d * sizeof(float), s));
    if (h->nstored > 0) {
        const int64_t thr = h->nstored * 32;
        reconstruct_kernel<<<(unsigned)((thr + 255) / 256), 256, 0, s>>>(h->vecs, h->ids, h->nstored, h->d, i0, ni, out_dev);
        RVCR_CHECK_LAUNCH();
        count_launch(h);
    }
    return RVCR_OK;
}

}  // namespace rvcr
