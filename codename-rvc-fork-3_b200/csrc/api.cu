// api.cu -- the C ABI declared in include/rvcr.h (handle, orchestration, host-buffer variants).
//
// Call-site map (reference paths):
//   rvcr_index_read / _info / rvcr_reconstruct_n  <- rvc/infer/pipeline.py:553-561
//   rvcr_search / rvcr_retrieve_blend             <- rvc/infer/pipeline.py:497-507
//   rvcr_index_create / _set_nprobe / _add / _write <- rvc/train/process/extract_index.py:59-71
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <random>

#include "common.cuh"

namespace rvcr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

int read_index_file(const char* path, int device, int rank, int world, rvcr_index** out);
int probe_index_file(const char* path, rvcr_file_info* out, int64_t* sizes_out, int64_t sizes_cap);
void shard_range(const std::vector<int64_t>& sizes, int rank, int world, int64_t& lo, int64_t& hi);
int write_index_file(rvcr_index* h, const char* path);

namespace {

// RAII: select the handle's device for the duration of a call, restore the caller's afterwards.
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) {
            ok = false;
            return;
        }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

#define RVCR_ENTER(h)                                            \
    if (!(h)) {                                                  \
        rvcr::set_error("null index handle");                    \
        return RVCR_ERR_INVALID;                                 \
    }                                                            \
    std::lock_guard<std::mutex> _lock((h)->mu);                  \
    rvcr::DeviceGuard _guard((h)->device);                       \
    if (!_guard.ok) {                                            \
        rvcr::set_error("cannot select CUDA device %d: %s", (h)->device, cudaGetErrorString(cudaGetLastError())); \
        return RVCR_ERR_CUDA;                                    \
    }

// Scratch buffers are shared by all calls on a handle: a call on a different stream than the
// previous one first waits for the previous call's work.
int stream_enter(rvcr_index* h, cudaStream_t s) {
    if (h->have_last && h->last_stream != s) RVCR_CUDA(cudaStreamWaitEvent(s, h->last_done, 0));
    return RVCR_OK;
}
int stream_leave(rvcr_index* h, cudaStream_t s) {
    if (!h->last_done) RVCR_CUDA(cudaEventCreateWithFlags(&h->last_done, cudaEventDisableTiming));
    RVCR_CUDA(cudaEventRecord(h->last_done, s));
    h->last_stream = s;
    h->have_last = true;
    return RVCR_OK;
}

// Every entry point that touches the handle's shared scratch enters the stream BEFORE its first access (host variants
// included: their memcpys into ws_q would otherwise race with a previous asynchronous call on another stream) and records
// the leave event on every way out, error returns too, so last_stream / last_done never go stale.
struct StreamScope {
    rvcr_index* h;
    cudaStream_t s;
    bool entered = false;
    StreamScope(rvcr_index* h_, cudaStream_t s_) : h(h_), s(s_) {}
    int enter() {
        const int rc = stream_enter(h, s);
        entered = rc == RVCR_OK;
        return rc;
    }
    ~StreamScope() {
        if (entered) stream_leave(h, s);
    }
};

int check_searchable(rvcr_index* h, int32_t k) {
    if (!h->trained) {
        set_error("index is not trained (no centroids)");
        return RVCR_ERR_STATE;
    }
    if (k != RVCR_K) {
        set_error("k=%d unsupported: the RVC pipeline always searches k=8 (pipeline.py:499)", k);
        return RVCR_ERR_UNSUPPORTED;
    }
    if (h->nprobe < 1) {
        set_error("nprobe=%lld is not a valid probe count", (long long)h->nprobe);
        return RVCR_ERR_STATE;
    }
    return RVCR_OK;   // nprobe > 1 (a file made elsewhere; this fork writes 1, extract_index.py:64): multiprobe.cu
}

// nearest centroid for n device queries; dispatch FFMA-exact vs tcgen05 + exact re-check
int coarse_dispatch(rvcr_index* h, const float* q, int64_t n, int32_t* list_out, float* dist_out, cudaStream_t s) {
    int64_t mode = h->opt.coarse_mode;
    if (mode == 0) {
        // the tensor-core path pays a fixed conversion + re-check cost; use it once the
        // contraction is large enough to matter (Options::coarse_tc_min_flops)
        const double flops = 2.0 * (double)n * (double)h->nlist * (double)h->d;
        mode = (coarse_tc_available() && flops >= (double)h->opt.coarse_tc_min_flops && h->nlist >= 64 && n >= 1) ? 2 : 1;
    }
    if (mode >= 2) {
        if (!coarse_tc_available()) {
            set_error("coarse_mode=%lld requested but the tcgen05 path is not built", (long long)mode);
            return RVCR_ERR_UNSUPPORTED;
        }
        h->opt.coarse_mode = h->opt.coarse_mode;  // unchanged; split factor is read by coarse_tc
        return coarse_tc(h, q, n, list_out, dist_out, s);
    }
    PhaseTimer pt(h, RVCR_PHASE_COARSE_MAIN, s);
    return coarse_exact(h, q, n, nullptr, 0, list_out, dist_out, s);
}

int search_device(rvcr_index* h, float* feats_or_q, int64_t n, const int32_t* assign_in, bool blend, double index_rate,
                  float* D_dev, int64_t* I_dev, cudaStream_t s) {
    if (n < 0) {
        set_error("negative query count");
        return RVCR_ERR_INVALID;
    }
    RVCR_TRY(check_searchable(h, RVCR_K));
    if (n == 0) return RVCR_OK;
    StreamScope scope(h, s);
    RVCR_TRY(scope.enter());
    RVCR_TRY(finalize_pending(h, s));
    // a caller-supplied assignment came from an earlier rvcr_assign call: keep its coarse phase timings
    for (int i = assign_in ? RVCR_PHASE_GROUP : 0; i < RVCR_NUM_PHASES; ++i) h->pev_used[i] = false;
    if (h->nprobe > 1 && h->nlist > 1) {
        if (assign_in) {
            set_error("a caller-supplied list assignment carries one list per query; this index has nprobe=%lld", (long long)h->nprobe);
            return RVCR_ERR_UNSUPPORTED;
        }
        return search_multiprobe(h, feats_or_q, n, blend, index_rate, D_dev, I_dev, s);
    }
    const int32_t* assign = assign_in;
    if (!assign) {
        RVCR_TRY(h->ws_assign.reserve((size_t)n * 4));
        RVCR_TRY(coarse_dispatch(h, feats_or_q, n, h->ws_assign.as<int32_t>(), nullptr, s));
        assign = h->ws_assign.as<int32_t>();
    }
    {
        PhaseTimer pt(h, RVCR_PHASE_GROUP, s);
        RVCR_TRY(group_queries(h, assign, n, assign_in != nullptr, s));
    }
    RVCR_TRY(h->ws_P.reserve((size_t)n * RVCR_K * 4));
    float* D = D_dev;
    if (!D) {
        RVCR_TRY(h->ws_D.reserve((size_t)n * RVCR_K * 4));
        D = h->ws_D.as<float>();
    }
    h->fuse.active = blend && h->opt.fused_blend != 0;
    h->fuse.done = false;
    h->fuse.out = feats_or_q;
    h->fuse.r32 = (float)index_rate;
    h->fuse.omr32 = (float)(1.0 - index_rate);
    {
        PhaseTimer pt(h, RVCR_PHASE_SCAN, s);
        const int rc = scan_topk(h, feats_or_q, n, D, h->ws_P.as<int32_t>(), I_dev, s);
        h->fuse.active = false;
        RVCR_TRY(rc);
    }
    if (blend && !h->fuse.done) {
        PhaseTimer pt(h, RVCR_PHASE_BLEND, s);
        RVCR_TRY(gather_blend(h, feats_or_q, n, D, h->ws_P.as<int32_t>(), assign_in, index_rate, s));
    }
    h->stats.calls_search += 1;
    h->stats.queries += n;
    return RVCR_OK;
}

// Append n rows (ids ntotal .. ntotal+n-1) with their list assignment to the id-ordered staging
// area; grows geometrically (amortised O(1) per appended row).  Caller holds the handle lock.
int stage_rows(rvcr_index* h, const float* x_dev, int64_t n, const int32_t* assign_dev, cudaStream_t s) {
    if (n == 0) return RVCR_OK;
    const int64_t need = h->npending + n;
    if (need > h->stage_cap) {
        int64_t cap = h->stage_cap ? h->stage_cap : 8192;
        while (cap < need) cap *= 2;
        float* nx = nullptr;
        int32_t* na = nullptr;
        cudaError_t e = cudaMalloc(&nx, (size_t)cap * h->d * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&na, (size_t)cap * 4);
        if (e != cudaSuccess) {
            if (nx) cudaFree(nx);
            set_error("staging allocation for %lld rows failed: %s", (long long)cap, cudaGetErrorString(e));
            return RVCR_ERR_OOM;
        }
        if (h->npending) {
            RVCR_CUDA(cudaMemcpyAsync(nx, h->stage_x, (size_t)h->npending * h->d * sizeof(float), cudaMemcpyDeviceToDevice, s));
            RVCR_CUDA(cudaMemcpyAsync(na, h->stage_assign, (size_t)h->npending * 4, cudaMemcpyDeviceToDevice, s));
        }
        RVCR_CUDA(cudaStreamSynchronize(s));
        if (h->stage_x) cudaFree(h->stage_x);
        if (h->stage_assign) cudaFree(h->stage_assign);
        h->stage_x = nx;
        h->stage_assign = na;
        h->stage_cap = cap;
    }
    RVCR_CUDA(cudaMemcpyAsync(h->stage_x + (size_t)h->npending * h->d, x_dev, (size_t)n * h->d * sizeof(float),
                              cudaMemcpyDeviceToDevice, s));
    RVCR_CUDA(cudaMemcpyAsync(h->stage_assign + h->npending, assign_dev, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
    h->npending += n;
    h->ntotal += n;
    return RVCR_OK;
}

}  // namespace

// kmeans.cu's assignment step: the same dispatch (FFMA-exact vs tcgen05 + exact re-check) as the search path
int coarse_dispatch_for_train(rvcr_index* h, const float* q, int64_t n, int32_t* list_out, cudaStream_t s) {
    return coarse_dispatch(h, q, n, list_out, nullptr, s);
}
// minibatch_kmeans.cu's E-step: nearest centre and its canonical distance
int coarse_dispatch_dist(rvcr_index* h, const float* q, int64_t n, int32_t* list_out, float* dist_out, cudaStream_t s) {
    return coarse_dispatch(h, q, n, list_out, dist_out, s);
}
int minibatch_kmeans(rvcr_index* h, const float* x_host, int64_t n, int64_t batch_size, int64_t max_iter, int64_t max_no_improvement,
                     double reassignment_ratio, int64_t n_init, uint64_t seed, int verbose, float* centers_out, rvcr_mbk_stats* st_out);
}  // namespace rvcr

using namespace rvcr;

extern "C" {

const char* rvcr_last_error(void) { return get_error(); }
int rvcr_version(void) { return RVCR_VERSION; }

int rvcr_device_count(int* count_out) {
    if (!count_out) return RVCR_ERR_INVALID;
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
        *count_out = 0;
        return RVCR_ERR_CUDA;
    }
    *count_out = c;
    return RVCR_OK;
}

int rvcr_index_create(int32_t d, int64_t nlist, int device, rvcr_index** out) {
    if (!out || d <= 0 || nlist <= 0) {
        set_error("rvcr_index_create: bad arguments (d=%d nlist=%lld)", d, (long long)nlist);
        return RVCR_ERR_INVALID;
    }
    int ndev = 0;
    RVCR_TRY(rvcr_device_count(&ndev));
    if (device < 0 || device >= ndev) {
        set_error("CUDA device %d not available (%d visible); this library has no CPU fallback", device, ndev);
        return RVCR_ERR_CUDA;
    }
    cudaDeviceProp prop;
    RVCR_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; librvcr is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return RVCR_ERR_CUDA;
    }
    rvcr_index* h = new rvcr_index();
    h->device = device;
    h->d = d;
    h->nlist = nlist;
    h->own_lo = 0;
    h->own_hi = nlist;
    h->num_sms = prop.multiProcessorCount;
    h->offsets_host.assign((size_t)nlist + 1, 0);
    {
        // small fixed scratch block: work-item counters, ticket, re-check counters (layout: see users)
        DeviceGuard g(device);
        int rc = h->ws_misc.reserve(1024);
        if (rc == RVCR_OK && cudaMemset(h->ws_misc.p, 0, 1024) != cudaSuccess) rc = RVCR_ERR_CUDA;
        if (rc != RVCR_OK) {
            delete h;
            return rc;
        }
    }
    *out = h;
    return RVCR_OK;
}

int rvcr_index_read(const char* path, int device, rvcr_index** out) {
    if (!path || !out) return RVCR_ERR_INVALID;
    return read_index_file(path, device, 0, 1, out);
}

int rvcr_index_probe(const char* path, rvcr_file_info* info_out, int64_t* list_sizes_host, int64_t nlist_cap) {
    if (!path || (!info_out && !list_sizes_host) || (list_sizes_host && nlist_cap < 0)) {
        set_error("rvcr_index_probe: bad arguments");
        return RVCR_ERR_INVALID;
    }
    return probe_index_file(path, info_out, list_sizes_host, nlist_cap);
}

int rvcr_shard_range(const int64_t* list_sizes_host, int64_t nlist, int rank, int world, int64_t* lo_out, int64_t* hi_out) {
    if (!list_sizes_host || nlist <= 0 || world < 1 || rank < 0 || rank >= world || !lo_out || !hi_out) {
        set_error("rvcr_shard_range: bad arguments");
        return RVCR_ERR_INVALID;
    }
    for (int64_t l = 0; l < nlist; ++l)
        if (list_sizes_host[l] < 0) {
            set_error("rvcr_shard_range: negative list size");
            return RVCR_ERR_INVALID;
        }
    try {
        std::vector<int64_t> sizes(list_sizes_host, list_sizes_host + nlist);
        shard_range(sizes, rank, world, *lo_out, *hi_out);
    } catch (const std::exception& e) {
        set_error("rvcr_shard_range: %s", e.what());
        return RVCR_ERR_OOM;
    }
    return RVCR_OK;
}

int rvcr_index_read_shard(const char* path, int device, int rank, int world, rvcr_index** out) {
    if (!path || !out || world < 1 || rank < 0 || rank >= world) {
        set_error("rvcr_index_read_shard: bad rank/world");
        return RVCR_ERR_INVALID;
    }
    return read_index_file(path, device, rank, world, out);
}

int rvcr_index_write(rvcr_index* h, const char* path) {
    RVCR_ENTER(h);
    if (!path) return RVCR_ERR_INVALID;
    if (!h->trained) {
        // faiss would serialise an empty quantizer; the reference never does (train comes before write, extract_index.py:
        // 65-71) and this library's own reader refuses such a file, so say so here instead of writing one
        set_error("write_index on an untrained index (no centroids): train or set the centroids first");
        return RVCR_ERR_STATE;
    }
    RVCR_TRY(finalize_pending(h, nullptr));
    RVCR_CUDA(cudaDeviceSynchronize());
    return write_index_file(h, path);
}

void rvcr_index_free(rvcr_index* h) {
    if (!h) return;
    {
        DeviceGuard g(h->device);
        cudaDeviceSynchronize();
        if (h->cent) cudaFree(h->cent);
        if (h->cent_bf16) cudaFree(h->cent_bf16);
        if (h->cent_norm) cudaFree(h->cent_norm);
        if (h->cent_f8) cudaFree(h->cent_f8);
        if (h->vecs) cudaFree(h->vecs);
        if (h->ids) cudaFree(h->ids);
        if (h->offsets) cudaFree(h->offsets);
        if (h->stage_x) cudaFree(h->stage_x);
        if (h->stage_assign) cudaFree(h->stage_assign);
        DevBuf* bufs[] = {&h->ws_assign, &h->ws_keys, &h->ws_vals, &h->ws_keys2, &h->ws_vals2, &h->ws_cub, &h->ws_items,
                          &h->ws_flags, &h->ws_misc, &h->ws_D, &h->ws_P, &h->ws_I, &h->ws_best, &h->ws_q, &h->ws_tc0,
                          &h->ws_tc1, &h->ws_tc2, &h->ws_tc3, &h->ws_tc4, &h->ws_tc5, &h->ws_tc6, &h->ws_tc7, &h->ws_stage, &h->row_norm, &h->ws_cand, &h->ws_candcnt,
                          &h->ws_fb, &h->ws_cstage};
        for (DevBuf* b : bufs) b->release();
        for (DevBuf& b : h->ws_mp) b.release();
        if (h->last_done) cudaEventDestroy(h->last_done);
        if (h->f8_fb_event) cudaEventDestroy(h->f8_fb_event);
        if (h->f8_fb_host) cudaFreeHost(h->f8_fb_host);
        if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
        if (h->out_stream) cudaStreamDestroy(h->out_stream);
        for (cudaEvent_t e : h->copy_events) cudaEventDestroy(e);
        for (int i = 0; i < RVCR_NUM_PHASES; ++i)
            for (int k = 0; k < 2; ++k)
                if (h->pev[i][k]) cudaEventDestroy(h->pev[i][k]);
    }
    delete h;
}

int rvcr_index_info(rvcr_index* h, rvcr_info* out) {
    if (!h || !out) return RVCR_ERR_INVALID;
    std::lock_guard<std::mutex> lock(h->mu);
    out->d = h->d;
    out->is_trained = h->trained ? 1 : 0;
    out->ntotal = h->ntotal;
    out->nlist = h->nlist;
    out->nprobe = h->nprobe;
    out->nstored = h->nstored + h->npending;
    out->owned_begin = h->own_lo;
    out->owned_end = h->own_hi;
    out->device = h->device;
    out->metric = h->metric;
    return RVCR_OK;
}

int rvcr_index_set_nprobe(rvcr_index* h, int64_t nprobe) {
    if (!h || nprobe < 1) return RVCR_ERR_INVALID;
    std::lock_guard<std::mutex> lock(h->mu);
    h->nprobe = nprobe;
    return RVCR_OK;
}

int rvcr_index_set_shard(rvcr_index* h, int64_t owned_begin, int64_t owned_end) {
    if (!h) return RVCR_ERR_INVALID;
    std::lock_guard<std::mutex> lock(h->mu);
    if (owned_begin < 0 || owned_begin > owned_end || owned_end > h->nlist) {
        set_error("set_shard: [%lld, %lld) is not a list range of an index with %lld lists", (long long)owned_begin,
                  (long long)owned_end, (long long)h->nlist);
        return RVCR_ERR_INVALID;
    }
    h->own_lo = owned_begin;
    h->own_hi = owned_end;
    return RVCR_OK;
}

int rvcr_index_set_centroids(rvcr_index* h, const float* centroids_host, int64_t nlist) {
    RVCR_ENTER(h);
    if (!centroids_host || nlist != h->nlist) {
        set_error("set_centroids: expected %lld x %d centroids", (long long)h->nlist, h->d);
        return RVCR_ERR_INVALID;
    }
    if (!h->cent) RVCR_CUDA(cudaMalloc(&h->cent, (size_t)nlist * h->d * sizeof(float)));
    RVCR_CUDA(cudaMemcpy(h->cent, centroids_host, (size_t)nlist * h->d * sizeof(float), cudaMemcpyHostToDevice));
    h->trained = true;
    return coarse_tc_invalidate(h);
}

int rvcr_index_get_centroids(rvcr_index* h, float* centroids_host) {
    RVCR_ENTER(h);
    if (!centroids_host) return RVCR_ERR_INVALID;
    if (!h->trained) {
        set_error("index is not trained");
        return RVCR_ERR_STATE;
    }
    RVCR_CUDA(cudaMemcpy(centroids_host, h->cent, (size_t)h->nlist * h->d * sizeof(float), cudaMemcpyDeviceToHost));
    return RVCR_OK;
}

int rvcr_index_list_sizes(rvcr_index* h, int64_t* sizes_host) {
    RVCR_ENTER(h);
    if (!sizes_host) return RVCR_ERR_INVALID;
    RVCR_TRY(finalize_pending(h, nullptr));
    for (int64_t l = 0; l < h->nlist; ++l) sizes_host[l] = h->offsets_host[l + 1] - h->offsets_host[l];
    return RVCR_OK;
}

int rvcr_index_train(rvcr_index* h, const float* x_host, int64_t n) {
    RVCR_ENTER(h);
    if (!x_host || n <= 0) {
        set_error("train: need n > 0 host rows");
        return RVCR_ERR_INVALID;
    }
    if (h->nstored > 0 || h->npending > 0) {
        set_error("train on an index that already holds vectors is unsupported");
        return RVCR_ERR_STATE;
    }
    return kmeans_train(h, x_host, n, nullptr);
}

int rvcr_minibatch_kmeans(const float* x_host, int64_t n, int32_t d, int64_t k, int64_t batch_size, int64_t max_iter,
                          int64_t max_no_improvement, double reassignment_ratio, int64_t n_init, uint64_t seed, int device, int verbose,
                          float* centers_out_host, rvcr_mbk_stats* stats_out) {
    if (!x_host || !centers_out_host || n <= 0 || d <= 0 || k <= 0) {
        set_error("minibatch_kmeans: bad arguments");
        return RVCR_ERR_INVALID;
    }
    rvcr_index* h = nullptr;
    RVCR_TRY(rvcr_index_create(d, k, device, &h));
    if (seed == 0) {
        std::random_device rd;
        seed = ((uint64_t)rd() << 32) ^ (uint64_t)rd() ^ 0x9e3779b97f4a7c15ull;
    }
    const int rc = minibatch_kmeans(h, x_host, n, batch_size, max_iter, max_no_improvement, reassignment_ratio, n_init, seed, verbose,
                                    centers_out_host, stats_out);
    rvcr_index_free(h);
    return rc;
}

int rvcr_debug_rand_perm(int64_t n, int64_t seed, int64_t* perm_out_host) {
    if (n < 0 || (n > 0 && !perm_out_host)) return RVCR_ERR_INVALID;
    kmeans_debug_rand_perm(n, seed, perm_out_host);
    return RVCR_OK;
}

int rvcr_debug_split_clusters(int32_t d, int64_t k, int64_t n, float* hassign_host, float* centroids_host, int32_t* nsplit_out) {
    if (d < 1 || k < 1 || n <= k || !hassign_host || !centroids_host) return RVCR_ERR_INVALID;
    bool any_donor = false;
    for (int64_t c = 0; c < k; ++c) any_donor = any_donor || hassign_host[c] > 1.f;
    if (!any_donor) {
        set_error("split_clusters: no cluster with more than one point to split");
        return RVCR_ERR_INVALID;
    }
    const int ns = kmeans_debug_split_clusters(d, k, n, hassign_host, centroids_host);
    if (nsplit_out) *nsplit_out = ns;
    return RVCR_OK;
}

int rvcr_index_add_preassigned(rvcr_index* h, const float* x_dev, int64_t n, const int32_t* assign_dev, void* stream) {
    RVCR_ENTER(h);
    if (n < 0 || (n > 0 && (!x_dev || !assign_dev))) return RVCR_ERR_INVALID;
    return stage_rows(h, x_dev, n, assign_dev, (cudaStream_t)stream);
}

int rvcr_index_add(rvcr_index* h, const float* x, int64_t n, int x_on_device, void* stream) {
    RVCR_ENTER(h);
    cudaStream_t s = (cudaStream_t)stream;
    if (n < 0 || (n > 0 && !x)) return RVCR_ERR_INVALID;
    if (n == 0) return RVCR_OK;
    if (!h->trained) {
        set_error("index.add on an untrained index");
        return RVCR_ERR_STATE;
    }
    StreamScope scope(h, s);
    RVCR_TRY(scope.enter());
    const float* xd = x;
    if (!x_on_device) {
        RVCR_TRY(h->ws_q.reserve((size_t)n * h->d * sizeof(float)));
        RVCR_CUDA(cudaMemcpyAsync(h->ws_q.p, x, (size_t)n * h->d * sizeof(float), cudaMemcpyHostToDevice, s));
        xd = h->ws_q.as<float>();
    }
    RVCR_TRY(h->ws_assign.reserve((size_t)n * 4));
    int32_t* assign = h->ws_assign.as<int32_t>();
    RVCR_TRY(coarse_dispatch(h, xd, n, assign, nullptr, s));
    RVCR_TRY(stage_rows(h, xd, n, assign, s));
    // a pageable host buffer may be reused by the caller as soon as we return
    if (!x_on_device) RVCR_CUDA(cudaStreamSynchronize(s));
    return RVCR_OK;
}

int rvcr_index_alloc_lists(rvcr_index* h, const int64_t* sizes_host, int64_t ntotal) {
    RVCR_ENTER(h);
    if (!sizes_host || ntotal < 0) return RVCR_ERR_INVALID;
    if (h->npending) {
        set_error("alloc_lists with staged adds pending");
        return RVCR_ERR_STATE;
    }
    std::vector<int64_t> off((size_t)h->nlist + 1, 0);
    for (int64_t l = 0; l < h->nlist; ++l) {
        if (sizes_host[l] < 0) return RVCR_ERR_INVALID;
        off[l + 1] = off[l] + sizes_host[l];
    }
    const int64_t nst = off[h->nlist];
    if (h->vecs) cudaFree(h->vecs);
    if (h->ids) cudaFree(h->ids);
    h->vecs = nullptr;
    h->ids = nullptr;
    h->nstored = 0;
    if (nst > 0) {
        cudaError_t e = cudaMalloc(&h->vecs, (size_t)nst * h->d * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&h->ids, (size_t)nst * 8);
        if (e != cudaSuccess) {
            set_error("allocating %lld x %d list storage failed: %s", (long long)nst, h->d, cudaGetErrorString(e));
            return RVCR_ERR_OOM;
        }
    }
    if (!h->offsets) RVCR_CUDA(cudaMalloc(&h->offsets, (size_t)(h->nlist + 1) * 8));
    RVCR_CUDA(cudaMemcpy(h->offsets, off.data(), (size_t)(h->nlist + 1) * 8, cudaMemcpyHostToDevice));
    h->offsets_host = off;
    h->nstored = nst;
    h->ntotal = ntotal;
    h->last_row_pos = -1;
    scan_tc_invalidate(h);
    return RVCR_OK;
}

int rvcr_index_device_ptrs(rvcr_index* h, float** vecs_dev, int64_t** ids_dev, int64_t** offsets_dev, float** centroids_dev) {
    RVCR_ENTER(h);
    RVCR_TRY(finalize_pending(h, nullptr));
    if (vecs_dev) *vecs_dev = h->vecs;
    if (ids_dev) *ids_dev = h->ids;
    if (offsets_dev) *offsets_dev = h->offsets;
    if (centroids_dev) *centroids_dev = h->cent;
    return RVCR_OK;
}

int rvcr_index_commit(rvcr_index* h, void* stream) {
    RVCR_ENTER(h);
    RVCR_TRY(coarse_tc_invalidate(h));
    return refresh_derived(h, (cudaStream_t)stream);
}

int rvcr_assign(rvcr_index* h, const float* x_dev, int64_t n, int32_t* list_out_dev, float* dist_out_dev, void* stream) {
    RVCR_ENTER(h);
    cudaStream_t s = (cudaStream_t)stream;
    if (n < 0 || (n > 0 && (!x_dev || !list_out_dev))) return RVCR_ERR_INVALID;
    if (!h->trained) {
        set_error("index is not trained (no centroids)");
        return RVCR_ERR_STATE;
    }
    if (n == 0) return RVCR_OK;
    StreamScope scope(h, s);
    RVCR_TRY(scope.enter());
    for (int i = 0; i < RVCR_NUM_PHASES; ++i) h->pev_used[i] = false;
    return coarse_dispatch(h, x_dev, n, list_out_dev, dist_out_dev, s);
}

int rvcr_search(rvcr_index* h, const float* q_dev, int64_t n, int32_t k, float* D_dev, int64_t* I_dev, void* stream) {
    RVCR_ENTER(h);
    if (n > 0 && (!q_dev || !D_dev || !I_dev)) return RVCR_ERR_INVALID;
    RVCR_TRY(check_searchable(h, k));
    return search_device(h, const_cast<float*>(q_dev), n, nullptr, false, 0.0, D_dev, I_dev, (cudaStream_t)stream);
}

int rvcr_search_host(rvcr_index* h, const float* q_host, int64_t n, int32_t k, float* D_host, int64_t* I_host) {
    RVCR_ENTER(h);
    if (n > 0 && (!q_host || !D_host || !I_host)) return RVCR_ERR_INVALID;
    RVCR_TRY(check_searchable(h, k));
    if (n <= 0) return n < 0 ? RVCR_ERR_INVALID : RVCR_OK;
    cudaStream_t s = nullptr;
    StreamScope scope(h, s);
    RVCR_TRY(scope.enter());
    RVCR_TRY(h->ws_q.reserve((size_t)n * h->d * sizeof(float)));
    RVCR_TRY(h->ws_D.reserve((size_t)n * RVCR_K * 4));
    RVCR_TRY(h->ws_I.reserve((size_t)n * RVCR_K * 8));
    RVCR_CUDA(cudaMemcpyAsync(h->ws_q.p, q_host, (size_t)n * h->d * sizeof(float), cudaMemcpyHostToDevice, s));
    RVCR_TRY(search_device(h, h->ws_q.as<float>(), n, nullptr, false, 0.0, h->ws_D.as<float>(), h->ws_I.as<int64_t>(), s));
    RVCR_CUDA(cudaMemcpyAsync(D_host, h->ws_D.p, (size_t)n * RVCR_K * 4, cudaMemcpyDeviceToHost, s));
    RVCR_CUDA(cudaMemcpyAsync(I_host, h->ws_I.p, (size_t)n * RVCR_K * 8, cudaMemcpyDeviceToHost, s));
    RVCR_CUDA(cudaStreamSynchronize(s));
    return RVCR_OK;
}

int rvcr_retrieve_blend(rvcr_index* h, float* feats_dev, int64_t n, double index_rate, float* D_dev, int64_t* I_dev,
                        void* stream) {
    RVCR_ENTER(h);
    if (n > 0 && !feats_dev) return RVCR_ERR_INVALID;
    return search_device(h, feats_dev, n, nullptr, true, index_rate, D_dev, I_dev, (cudaStream_t)stream);
}

int rvcr_retrieve_blend_upsample(rvcr_index* h, const float* feats_dev, int64_t n, double index_rate, const float* pitchf_dev,
                                 double protect, float* out_dev, float* D_dev, int64_t* I_dev, void* stream) {
    RVCR_ENTER(h);
    if (n > 0 && (!feats_dev || !out_dev)) return RVCR_ERR_INVALID;
    if (n > 0 && out_dev == feats_dev) {
        set_error("retrieve_blend_upsample: out (2n x d) must not alias feats (n x d)");
        return RVCR_ERR_INVALID;
    }
    if (h->d % 4 != 0) {
        set_error("d=%d unsupported (must be a multiple of 4)", h->d);
        return RVCR_ERR_UNSUPPORTED;
    }
    // the blend epilogues (scan re-check / fallback / gather_blend) write rows 2i, 2i+1 of out_dev instead of feats
    h->fuse.post.out2 = out_dev;
    h->fuse.post.pitchf = pitchf_dev;
    h->fuse.post.protect = (float)protect;
    const int rc = search_device(h, const_cast<float*>(feats_dev), n, nullptr, true, index_rate, D_dev, I_dev, (cudaStream_t)stream);
    h->fuse.post = PostArgs{};
    return rc;
}

int rvcr_upsample_protect(const float* feats_dev, const float* feats0_dev, const float* pitchf_dev, int64_t n, int32_t d,
                          double protect, float* out_dev, int device, void* stream) {
    if (n < 0 || d < 1 || (n > 0 && (!feats_dev || !out_dev))) return RVCR_ERR_INVALID;
    if ((pitchf_dev != nullptr) != (feats0_dev != nullptr)) {
        set_error("upsample_protect: pitchf and feats0 go together (both or neither)");
        return RVCR_ERR_INVALID;
    }
    DeviceGuard g(device);
    if (!g.ok) {
        set_error("cannot select CUDA device %d: %s", device, cudaGetErrorString(cudaGetLastError()));
        return RVCR_ERR_CUDA;
    }
    return upsample_protect(feats_dev, feats0_dev, pitchf_dev, n, d, protect, out_dev, (cudaStream_t)stream);
}

int rvcr_final_proj(const float* x_dev, int64_t n, int32_t d_in, const float* w_dev, const float* b_dev, int32_t d_out,
                    float* out_dev, int device, void* stream) {
    if (n < 0 || d_in < 1 || d_out < 1 || (n > 0 && (!x_dev || !w_dev || !out_dev))) return RVCR_ERR_INVALID;
    DeviceGuard g(device);
    if (!g.ok) {
        set_error("cannot select CUDA device %d: %s", device, cudaGetErrorString(cudaGetLastError()));
        return RVCR_ERR_CUDA;
    }
    return final_proj(x_dev, n, d_in, w_dev, b_dev, d_out, out_dev, (cudaStream_t)stream);
}

int rvcr_final_proj_retrieve_blend(rvcr_index* h, const float* x_dev, int64_t n, int32_t d_in, const float* w_dev,
                                   const float* b_dev, float* feats_out_dev, double index_rate, float* D_dev, int64_t* I_dev,
                                   void* stream) {
    RVCR_ENTER(h);
    if (n < 0 || d_in < 1 || (n > 0 && (!x_dev || !w_dev || !feats_out_dev))) return RVCR_ERR_INVALID;
    if (n > 0 && (const float*)feats_out_dev == x_dev) {
        set_error("final_proj_retrieve_blend: the projected rows (n x %d) must not alias the encoder rows (n x %d)", h->d, d_in);
        return RVCR_ERR_INVALID;
    }
    RVCR_TRY(final_proj(x_dev, n, d_in, w_dev, b_dev, h->d, feats_out_dev, (cudaStream_t)stream));
    count_launch(h);
    return search_device(h, feats_out_dev, n, nullptr, true, index_rate, D_dev, I_dev, (cudaStream_t)stream);
}

int rvcr_retrieve_blend_host(rvcr_index* h, float* feats_host, int64_t n, double index_rate) {
    RVCR_ENTER(h);
    if (n > 0 && !feats_host) return RVCR_ERR_INVALID;
    if (n <= 0) return n < 0 ? RVCR_ERR_INVALID : RVCR_OK;
    RVCR_TRY(check_searchable(h, RVCR_K));
    // Host -> device copies are pipelined against the coarse quantizer: the queries go up in chunks on a
    // copy stream while the tensor cores already assign the previous chunk (the coarse step is per-query
    // independent).  The list scan needs the whole batch grouped by list, so it starts after the last
    // chunk; the blended rows come back with one device -> host copy.
    cudaStream_t s = nullptr;
    StreamScope scope(h, s);
    RVCR_TRY(scope.enter());
    const size_t rowb = (size_t)h->d * sizeof(float);
    RVCR_TRY(h->ws_q.reserve((size_t)n * rowb));
    RVCR_TRY(h->ws_assign.reserve((size_t)n * 4));
    RVCR_TRY(finalize_pending(h, s));
    if (h->nprobe > 1 && h->nlist > 1) {   // interop path (multiprobe.cu): plain copy in, search + blend, copy out
        RVCR_CUDA(cudaMemcpyAsync(h->ws_q.p, feats_host, (size_t)n * rowb, cudaMemcpyHostToDevice, s));
        RVCR_TRY(search_device(h, h->ws_q.as<float>(), n, nullptr, true, index_rate, nullptr, nullptr, s));
        RVCR_CUDA(cudaMemcpyAsync(feats_host, h->ws_q.p, (size_t)n * rowb, cudaMemcpyDeviceToHost, s));
        RVCR_CUDA(cudaStreamSynchronize(s));
        return RVCR_OK;
    }
    if (!h->copy_stream) RVCR_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    // chunk sizes ramp up (32 Ki, 64 Ki, then 128 Ki rows) so that only a small first copy is exposed
    std::vector<int64_t> coff;
    {
        int64_t off = 0, c = 32768;
        while (off < n) {
            coff.push_back(off);
            off += c;
            if (c < 131072) c *= 2;
        }
        coff.push_back(n);
    }
    const int64_t nchunks = (int64_t)coff.size() - 1;
    while ((int64_t)h->copy_events.size() < nchunks) {
        cudaEvent_t e;
        RVCR_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->copy_events.push_back(e);
    }
    for (int i = 0; i < RVCR_NUM_PHASES; ++i) h->pev_used[i] = false;
    char* dq = reinterpret_cast<char*>(h->ws_q.p);
    const bool trace = getenv("RVCR_HOST_TRACE") != nullptr;
    std::vector<cudaEvent_t> cev;   // trace only: [0] start, [1 + 2c] chunk c copied, [2 + 2c] chunk c assigned
    if (trace) {
        cev.resize(1 + 2 * (size_t)nchunks);
        for (auto& e : cev) cudaEventCreate(&e);
        cudaEventRecord(cev[0], h->copy_stream);
    }
    for (int64_t c = 0; c < nchunks; ++c) {
        const int64_t off = coff[c], m = coff[c + 1] - off;
        RVCR_CUDA(cudaMemcpyAsync(dq + (size_t)off * rowb, reinterpret_cast<const char*>(feats_host) + (size_t)off * rowb,
                                  (size_t)m * rowb, cudaMemcpyHostToDevice, h->copy_stream));
        RVCR_CUDA(cudaEventRecord(h->copy_events[c], h->copy_stream));
        if (trace) cudaEventRecord(cev[1 + 2 * c], h->copy_stream);
    }
    for (int64_t c = 0; c < nchunks; ++c) {
        const int64_t off = coff[c], m = coff[c + 1] - off;
        RVCR_CUDA(cudaStreamWaitEvent(s, h->copy_events[c], 0));
        RVCR_TRY(coarse_dispatch(h, h->ws_q.as<float>() + (size_t)off * h->d, m, h->ws_assign.as<int32_t>() + off, nullptr, s));
        if (trace) cudaEventRecord(cev[2 + 2 * c], s);
    }
    if (trace) {
        cudaStreamSynchronize(s);
        for (int64_t c = 0; c < nchunks; ++c) {
            float a, b;
            cudaEventElapsedTime(&a, cev[0], cev[1 + 2 * c]);
            cudaEventElapsedTime(&b, cev[0], cev[2 + 2 * c]);
            fprintf(stderr, "[rvcr trace] chunk %lld (%lld rows): copied %.2f ms, assigned %.2f ms\n", (long long)c,
                    (long long)(coff[c + 1] - coff[c]), a, b);
        }
        for (auto& e : cev) cudaEventDestroy(e);
    }
    // Device -> host: when the caller's buffer is pinned (device-mapped), the lists are scanned in waves of
    // work items and each wave's blended rows are written by the blend kernel straight into the host buffer
    // on a second stream, so the PCIe write-back of wave k overlaps the scan of wave k+1.
    cudaPointerAttributes pat;
    const bool mapped = cudaPointerGetAttributes(&pat, feats_host) == cudaSuccess && pat.type == cudaMemoryTypeHost &&
                        pat.devicePointer != nullptr;
    cudaGetLastError();
    if (!mapped) {
        RVCR_TRY(search_device(h, h->ws_q.as<float>(), n, h->ws_assign.as<int32_t>(), true, index_rate, nullptr, nullptr, s));
        RVCR_CUDA(cudaMemcpyAsync(feats_host, h->ws_q.p, (size_t)n * rowb, cudaMemcpyDeviceToHost, s));
        RVCR_CUDA(cudaStreamSynchronize(s));
        return RVCR_OK;
    }
    if (!h->out_stream) RVCR_CUDA(cudaStreamCreateWithFlags(&h->out_stream, cudaStreamNonBlocking));
    RVCR_TRY(group_queries(h, h->ws_assign.as<int32_t>(), n, false, s));
    RVCR_TRY(h->ws_P.reserve((size_t)n * RVCR_K * 4));
    RVCR_TRY(h->ws_D.reserve((size_t)n * RVCR_K * 4));
    int nitems = 0;
    RVCR_CUDA(cudaMemcpyAsync(&nitems, h->ws_misc.p, 4, cudaMemcpyDeviceToHost, s));
    RVCR_CUDA(cudaStreamSynchronize(s));
    int waves = nitems / 512;
    int max_waves = 16;
    if (const char* e = getenv("RVCR_HOST_WAVES")) max_waves = atoi(e);   // tuning knob (1 = no overlap)
    if (max_waves > 16) max_waves = 16;
    if (waves > max_waves) waves = max_waves;
    if (waves < 1) waves = 1;
    while ((int)h->copy_events.size() < waves) {
        cudaEvent_t e;
        RVCR_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->copy_events.push_back(e);
    }
    RVCR_CUDA(cudaMemsetAsync(h->ws_misc.as<int32_t>() + 32, 0, 4 * 16, s));   // per-wave tickets
    float* out_dev = reinterpret_cast<float*>(pat.devicePointer);
    // Enough CTAs to keep PCIe busy; more of them slow the concurrent scan (LSU back-pressure).  SM-issued writes to the
    // mapped buffer top out at ~46-48 GB/s (cudaMemcpy D2H: 57 GB/s); a cp.async.bulk (TMA store) variant of the write-back
    // was measured at the same rate (r01, 187-193 ms e2e for every CTA count / buffer depth tried) and was dropped.
    int blend_ctas = 128;
    if (const char* e = getenv("RVCR_HOST_BLEND_CTAS")) blend_ctas = atoi(e);
    std::vector<cudaEvent_t> tev;
    if (trace) {
        tev.resize(2 + 3 * (size_t)waves);
        for (auto& e : tev) cudaEventCreate(&e);
        cudaEventRecord(tev[0], s);
    }
    for (int k = 0; k < waves; ++k) {
        const int ib = (int)((int64_t)nitems * k / waves), ie = (int)((int64_t)nitems * (k + 1) / waves);
        RVCR_TRY(scan_topk_range(h, h->ws_q.as<float>(), n, h->ws_D.as<float>(), h->ws_P.as<int32_t>(), nullptr, ib, ie, 32 + k,
                                 k == 0, s));
        RVCR_CUDA(cudaEventRecord(h->copy_events[k], s));
        if (trace) cudaEventRecord(tev[1 + 3 * k], s);
        RVCR_CUDA(cudaStreamWaitEvent(h->out_stream, h->copy_events[k], 0));
        if (trace) cudaEventRecord(tev[2 + 3 * k], h->out_stream);
        RVCR_TRY(gather_blend_range(h, h->ws_q.as<float>(), n, h->ws_D.as<float>(), h->ws_P.as<int32_t>(), index_rate, ib, ie,
                                    out_dev, blend_ctas, h->out_stream));
        if (trace) cudaEventRecord(tev[3 + 3 * k], h->out_stream);
    }
    if (trace) {
        cudaStreamSynchronize(h->out_stream);
        cudaStreamSynchronize(s);
        for (int k = 0; k < waves; ++k) {
            float a, b, c;
            cudaEventElapsedTime(&a, tev[0], tev[1 + 3 * k]);
            cudaEventElapsedTime(&b, tev[0], tev[2 + 3 * k]);
            cudaEventElapsedTime(&c, tev[0], tev[3 + 3 * k]);
            fprintf(stderr, "[rvcr trace] wave %d: scan done %.2f ms, blend start %.2f, blend done %.2f\n", k, a, b, c);
        }
        for (auto& e : tev) cudaEventDestroy(e);
    }
    h->stats.calls_search += 1;
    h->stats.queries += n;
    RVCR_CUDA(cudaStreamSynchronize(h->out_stream));
    RVCR_CUDA(cudaStreamSynchronize(s));
    return RVCR_OK;
}

int rvcr_retrieve_blend_assigned(rvcr_index* h, float* feats_dev, int64_t n, const int32_t* assign_dev, double index_rate,
                                 float* D_dev, int64_t* I_dev, void* stream) {
    RVCR_ENTER(h);
    if (n > 0 && (!feats_dev || !assign_dev)) return RVCR_ERR_INVALID;
    return search_device(h, feats_dev, n, assign_dev, true, index_rate, D_dev, I_dev, (cudaStream_t)stream);
}

int rvcr_retrieve_blend_assigned_push(rvcr_index* h, const float* rows_dev, int64_t m, const int32_t* assign_dev,
                                      double index_rate, float* const* dst_bases_dev, const int32_t* dst_rank_dev,
                                      const int32_t* dst_row_dev, void* stream) {
    RVCR_ENTER(h);
    if (m > 0 && (!rows_dev || !assign_dev || !dst_bases_dev || !dst_rank_dev || !dst_row_dev)) return RVCR_ERR_INVALID;
    if (h->d % 4 != 0) {
        set_error("d=%d unsupported (must be a multiple of 4)", h->d);
        return RVCR_ERR_UNSUPPORTED;
    }
    // the blend epilogues write row qi to dst_bases[dst_rank[qi]] + dst_row[qi] * d (peer-mapped memory) instead of rows_dev
    h->fuse.push.bases = dst_bases_dev;
    h->fuse.push.rank = dst_rank_dev;
    h->fuse.push.row = dst_row_dev;
    const int rc = search_device(h, const_cast<float*>(rows_dev), m, assign_dev, true, index_rate, nullptr, nullptr, (cudaStream_t)stream);
    h->fuse.push = PushArgs{};
    return rc;
}

int rvcr_peer_dispatch(const int32_t* assign_dev, int64_t n, int64_t row0, int32_t my_rank, int32_t G, const int32_t* bounds_dev,
                       void* const* inbox_bases_dev, int64_t cap, int device, void* stream) {
    if (n < 0 || cap <= 0 || my_rank < 0 || my_rank >= G || !inbox_bases_dev || (n > 0 && !assign_dev) || (G > 1 && !bounds_dev))
        return RVCR_ERR_INVALID;
    DeviceGuard g(device);
    if (!g.ok) {
        set_error("cannot select CUDA device %d: %s", device, cudaGetErrorString(cudaGetLastError()));
        return RVCR_ERR_CUDA;
    }
    return peer_dispatch(assign_dev, n, row0, my_rank, G, bounds_dev, reinterpret_cast<unsigned char* const*>(inbox_bases_dev), cap,
                         (cudaStream_t)stream);
}

int rvcr_peer_barrier(void* const* inbox_bases_dev, int32_t my_rank, int32_t G, int32_t epoch, int device, void* stream) {
    if (!inbox_bases_dev || my_rank < 0 || my_rank >= G) return RVCR_ERR_INVALID;
    DeviceGuard g(device);
    if (!g.ok) {
        set_error("cannot select CUDA device %d: %s", device, cudaGetErrorString(cudaGetLastError()));
        return RVCR_ERR_CUDA;
    }
    return peer_barrier(reinterpret_cast<unsigned char* const*>(inbox_bases_dev), my_rank, G, epoch, (cudaStream_t)stream);
}

int rvcr_retrieve_blend_inbox_push(rvcr_index* h, void* inbox_dev, int64_t cap, float* rows_scratch_dev, double index_rate,
                                   float* const* feat_bases_dev, void* stream) {
    RVCR_ENTER(h);
    cudaStream_t s = (cudaStream_t)stream;
    if (!inbox_dev || cap <= 0 || !feat_bases_dev) return RVCR_ERR_INVALID;
    if (h->d % 4 != 0) {
        set_error("d=%d unsupported (must be a multiple of 4)", h->d);
        return RVCR_ERR_UNSUPPORTED;
    }
    unsigned char* ib = reinterpret_cast<unsigned char*>(inbox_dev);
    int32_t* count = reinterpret_cast<int32_t*>(ib + RVCR_PEER_INBOX_COUNT_OFF);
    int32_t* assign = reinterpret_cast<int32_t*>(ib + RVCR_PEER_INBOX_HEADER);
    int32_t* src_rank = assign + cap;
    int32_t* src_row = assign + 2 * cap;
    RVCR_TRY(peer_inbox_seal(assign, count, cap, s));
    // rows_scratch_dev == NULL: the scan kernels read the query rows of their work items straight out of the home ranks'
    // buffers (PushArgs::direct); else they are first pulled into the scratch (the round-2 first version, kept for A/B runs)
    const bool direct = rows_scratch_dev == nullptr;
    if (!direct) RVCR_TRY(peer_pull_rows(feat_bases_dev, src_rank, src_row, cap, h->d, rows_scratch_dev, s, count));
    count_launch(h, direct ? 1 : 2);
    // the blend epilogues write row qi to feat_bases[src_rank[qi]] + src_row[qi] * d; sealed entries form no work
    h->fuse.push.bases = feat_bases_dev;
    h->fuse.push.rank = src_rank;
    h->fuse.push.row = src_row;
    h->fuse.push.direct = direct ? 1 : 0;
    const int rc = search_device(h, rows_scratch_dev, cap, assign, true, index_rate, nullptr, nullptr, s);
    h->fuse.push = PushArgs{};
    RVCR_TRY(rc);
    RVCR_CUDA(cudaMemsetAsync(count, 0, 4, s));
    return RVCR_OK;
}

int rvcr_peer_alloc(int device, int64_t bytes, void** ptr_out, unsigned char* handle64_out) {
    if (bytes <= 0 || !ptr_out || !handle64_out) return RVCR_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    DeviceGuard g(device);
    if (!g.ok) {
        set_error("cannot select CUDA device %d: %s", device, cudaGetErrorString(cudaGetLastError()));
        return RVCR_ERR_CUDA;
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e != cudaSuccess) {
        set_error("peer buffer of %lld bytes: %s", (long long)bytes, cudaGetErrorString(e));
        cudaGetLastError();
        return RVCR_ERR_OOM;
    }
    e = cudaMemset(p, 0, (size_t)bytes);   // inboxes start empty; feature rows are overwritten anyway
    if (e != cudaSuccess) {
        set_error("clearing the peer buffer failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        cudaFree(p);
        return RVCR_ERR_CUDA;
    }
    cudaIpcMemHandle_t hd;
    e = cudaIpcGetMemHandle(&hd, p);
    if (e != cudaSuccess) {
        set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        cudaFree(p);
        return RVCR_ERR_CUDA;
    }
    memcpy(handle64_out, &hd, 64);
    *ptr_out = p;
    return RVCR_OK;
}

int rvcr_peer_open(int device, const unsigned char* handle64, void** ptr_out) {
    if (!handle64 || !ptr_out) return RVCR_ERR_INVALID;
    DeviceGuard g(device);
    if (!g.ok) {
        set_error("cannot select CUDA device %d: %s", device, cudaGetErrorString(cudaGetLastError()));
        return RVCR_ERR_CUDA;
    }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle64, 64);
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        set_error("cudaIpcOpenMemHandle failed: %s (no peer access between these GPUs / processes?)", cudaGetErrorString(e));
        cudaGetLastError();
        return RVCR_ERR_CUDA;
    }
    *ptr_out = p;
    return RVCR_OK;
}

int rvcr_peer_close(int device, void* ptr) {
    if (!ptr) return RVCR_OK;
    DeviceGuard g(device);
    if (!g.ok) return RVCR_ERR_CUDA;
    RVCR_CUDA(cudaIpcCloseMemHandle(ptr));
    return RVCR_OK;
}

int rvcr_peer_free(int device, void* ptr) {
    if (!ptr) return RVCR_OK;
    DeviceGuard g(device);
    if (!g.ok) return RVCR_ERR_CUDA;
    RVCR_CUDA(cudaFree(ptr));
    return RVCR_OK;
}

int rvcr_peer_pull_rows(const float* const* bases_dev, const int32_t* src_rank_dev, const int32_t* src_row_dev, int64_t m,
                        int32_t d, float* out_dev, int device, void* stream) {
    if (m < 0 || d < 1 || (m > 0 && (!bases_dev || !src_rank_dev || !src_row_dev || !out_dev))) return RVCR_ERR_INVALID;
    DeviceGuard g(device);
    if (!g.ok) {
        set_error("cannot select CUDA device %d: %s", device, cudaGetErrorString(cudaGetLastError()));
        return RVCR_ERR_CUDA;
    }
    return peer_pull_rows(bases_dev, src_rank_dev, src_row_dev, m, d, out_dev, (cudaStream_t)stream);
}

int rvcr_search_assigned(rvcr_index* h, const float* q_dev, int64_t n, const int32_t* assign_dev, float* D_dev, int64_t* I_dev,
                         void* stream) {
    RVCR_ENTER(h);
    if (n > 0 && (!q_dev || !assign_dev || !D_dev || !I_dev)) return RVCR_ERR_INVALID;
    return search_device(h, const_cast<float*>(q_dev), n, assign_dev, false, 0.0, D_dev, I_dev, (cudaStream_t)stream);
}

int rvcr_reconstruct_n(rvcr_index* h, int64_t i0, int64_t ni, float* out, int out_on_device) {
    RVCR_ENTER(h);
    if (i0 < 0 || ni < 0 || i0 + ni > h->ntotal || (ni > 0 && !out)) {
        set_error("reconstruct_n(%lld, %lld) out of range (ntotal=%lld)", (long long)i0, (long long)ni, (long long)h->ntotal);
        return RVCR_ERR_INVALID;
    }
    if (ni == 0) return RVCR_OK;
    cudaStream_t s = nullptr;
    StreamScope scope(h, s);          // ws_stage below is shared scratch (it is also the coarse re-check's fallback list)
    RVCR_TRY(scope.enter());
    RVCR_TRY(finalize_pending(h, s));
    if (out_on_device) {
        RVCR_TRY(reconstruct(h, i0, ni, out, s));
        RVCR_CUDA(cudaStreamSynchronize(s));
        return RVCR_OK;
    }
    // host output: reconstruct in id-range chunks of <= 256 MB through device scratch
    const size_t rowb = (size_t)h->d * sizeof(float);
    const int64_t chunk = std::max<int64_t>(1, (int64_t)((256u << 20) / rowb));
    RVCR_TRY(h->ws_stage.reserve((size_t)std::min<int64_t>(chunk, ni) * rowb));
    for (int64_t a = 0; a < ni; a += chunk) {
        const int64_t m = std::min<int64_t>(chunk, ni - a);
        RVCR_TRY(reconstruct(h, i0 + a, m, h->ws_stage.as<float>(), s));
        RVCR_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(out) + (size_t)a * rowb, h->ws_stage.p, (size_t)m * rowb,
                                  cudaMemcpyDeviceToHost, s));
        RVCR_CUDA(cudaStreamSynchronize(s));
    }
    return RVCR_OK;
}

int rvcr_merge_coarse(const float* dist_all_dev, const int32_t* list_all_dev, int32_t G, int64_t n, int32_t* list_out_dev,
                      int device, void* stream) {
    if (G < 1 || n < 0 || (n > 0 && (!dist_all_dev || !list_all_dev || !list_out_dev))) return RVCR_ERR_INVALID;
    DeviceGuard g(device);
    if (!g.ok) return RVCR_ERR_CUDA;
    return merge_coarse(dist_all_dev, list_all_dev, G, n, list_out_dev, (cudaStream_t)stream);
}

int rvcr_merge_topk(const float* D_all_dev, const int64_t* I_all_dev, int32_t G, int64_t n, int32_t k, float* D_out_dev,
                    int64_t* I_out_dev, int device, void* stream) {
    if (G < 1 || n < 0 || k < 1 || (n > 0 && (!D_all_dev || !I_all_dev || !D_out_dev || !I_out_dev))) return RVCR_ERR_INVALID;
    DeviceGuard g(device);
    if (!g.ok) return RVCR_ERR_CUDA;
    return merge_topk(D_all_dev, I_all_dev, G, n, k, D_out_dev, I_out_dev, (cudaStream_t)stream);
}

int rvcr_set_option(rvcr_index* h, const char* key, int64_t value) {
    if (!h || !key) return RVCR_ERR_INVALID;
    std::lock_guard<std::mutex> lock(h->mu);
    if (!strcmp(key, "coarse_mode")) {
        if (value < 0 || value > 4) return RVCR_ERR_INVALID;   // 4: e4m3 operands (auto picks them only through the coarse_f8_auto probe)
        h->opt.coarse_mode = value;
    } else if (!strcmp(key, "profile")) {
        h->opt.profile = value ? 1 : 0;
    } else if (!strcmp(key, "scan_impl")) {
        if (value < 0 || value > 5) return RVCR_ERR_INVALID;
        h->opt.scan_impl = value;
    } else if (!strcmp(key, "tc_cluster")) {
        if (value != 1 && value != 2) return RVCR_ERR_INVALID;
        h->opt.tc_cluster = value;
    } else if (!strcmp(key, "tc_resident")) {
        if (value != 0 && value != 6 && value != 8 && value != 9 && value != 10) return RVCR_ERR_INVALID;
        h->opt.tc_resident = value;
    } else if (!strcmp(key, "coarse_f8_auto")) {
        h->opt.coarse_f8_auto = value ? 1 : 0;
        h->f8_decision = 0;
    } else if (!strcmp(key, "coarse_tc_min_flops")) {
        if (value < 0) return RVCR_ERR_INVALID;
        h->opt.coarse_tc_min_flops = value;
    } else if (!strcmp(key, "scan_tc_min_queries")) {
        if (value < 0) return RVCR_ERR_INVALID;
        h->opt.scan_tc_min_queries = value;
    } else if (!strcmp(key, "kmeans_niter")) {
        if (value < 0 || value > 1000) return RVCR_ERR_INVALID;
        h->opt.kmeans_niter = value;
    } else if (!strcmp(key, "fused_blend")) {
        h->opt.fused_blend = value ? 1 : 0;
    } else if (!strcmp(key, "nan_on_zero_score")) {
        if (value != 1) {
            set_error("only the reference behaviour (NaN row on a zero score) is implemented");
            return RVCR_ERR_UNSUPPORTED;
        }
        h->opt.nan_on_zero_score = value;
    } else {
        set_error("unknown option '%s'", key);
        return RVCR_ERR_INVALID;
    }
    return RVCR_OK;
}

int rvcr_get_option(rvcr_index* h, const char* key, int64_t* value_out) {
    if (!h || !key || !value_out) return RVCR_ERR_INVALID;
    std::lock_guard<std::mutex> lock(h->mu);
    if (!strcmp(key, "coarse_mode")) *value_out = h->opt.coarse_mode;
    else if (!strcmp(key, "nan_on_zero_score")) *value_out = h->opt.nan_on_zero_score;
    else if (!strcmp(key, "profile")) *value_out = h->opt.profile;
    else if (!strcmp(key, "tc_cluster")) *value_out = h->opt.tc_cluster;
    else if (!strcmp(key, "scan_impl")) *value_out = h->opt.scan_impl;
    else if (!strcmp(key, "tc_resident")) *value_out = h->opt.tc_resident;
    else if (!strcmp(key, "fused_blend")) *value_out = h->opt.fused_blend;
    else if (!strcmp(key, "kmeans_niter")) *value_out = h->opt.kmeans_niter;
    else if (!strcmp(key, "coarse_f8_auto")) *value_out = h->opt.coarse_f8_auto;
    else if (!strcmp(key, "coarse_f8_decision")) *value_out = h->f8_decision;   // read only: 0 not probed, 1 e4m3, -1 bf16
    else if (!strcmp(key, "coarse_tc_min_flops")) *value_out = h->opt.coarse_tc_min_flops;
    else if (!strcmp(key, "scan_tc_min_queries")) *value_out = h->opt.scan_tc_min_queries;
    else if (!strcmp(key, "coarse_tc_available")) *value_out = coarse_tc_available() ? 1 : 0;
    else {
        set_error("unknown option '%s'", key);
        return RVCR_ERR_INVALID;
    }
    return RVCR_OK;
}

int rvcr_get_phase_ms(rvcr_index* h, float* ms_out) {
    RVCR_ENTER(h);
    if (!ms_out) return RVCR_ERR_INVALID;
    for (int i = 0; i < RVCR_NUM_PHASES; ++i) {
        ms_out[i] = -1.f;
        if (h->pev_used[i]) {
            RVCR_CUDA(cudaEventSynchronize(h->pev[i][1]));
            RVCR_CUDA(cudaEventElapsedTime(&ms_out[i], h->pev[i][0], h->pev[i][1]));
        }
    }
    return RVCR_OK;
}

int rvcr_get_stats(rvcr_index* h, rvcr_stats* out) {
    RVCR_ENTER(h);
    if (!out) return RVCR_ERR_INVALID;
    if (h->tc_counters_dirty) {   // running totals kept on the device by the re-check kernel
        unsigned long long tot[2] = {0, 0};
        RVCR_CUDA(cudaDeviceSynchronize());
        RVCR_CUDA(cudaMemcpy(tot, h->ws_misc.as<int32_t>() + 24, sizeof(tot), cudaMemcpyDeviceToHost));
        h->stats.coarse_fallback = (int64_t)tot[0];
        h->stats.coarse_candidates = (int64_t)tot[1];
        h->tc_counters_dirty = false;
    }
    if (h->scan_fb_dirty) {
        unsigned long long tot[2] = {0, 0};
        RVCR_CUDA(cudaDeviceSynchronize());
        RVCR_CUDA(cudaMemcpy(tot, h->ws_misc.as<int32_t>() + 48, sizeof(tot), cudaMemcpyDeviceToHost));
        h->stats.scan_tc_fallback = (int64_t)tot[0];
        h->stats.scan_tc_candidates = (int64_t)tot[1];
        h->scan_fb_dirty = false;
    }
    *out = h->stats;
    return RVCR_OK;
}

int rvcr_debug_coarse_tc(rvcr_index* h, const float* q_dev, int64_t n, int32_t split, int32_t cluster, int32_t cap_pairs,
                         float* v_host, int32_t* j_host, int32_t* pairs_out, float* qnorm2_host, float* cmax_out,
                         float* margin_host, float* dropped_host, void* stream) {
    RVCR_ENTER(h);
    cudaStream_t s = (cudaStream_t)stream;
    if (n <= 0 || !q_dev || !v_host || !j_host || !pairs_out || (cluster != 1 && cluster != 2)) return RVCR_ERR_INVALID;
    if (!h->trained) {
        set_error("index is not trained");
        return RVCR_ERR_STATE;
    }
    StreamScope scope(h, s);
    RVCR_TRY(scope.enter());
    float *topv, *topd, *qnorm, *qmargin;
    int* topj;
    int nsplit = 0;
    const int sp = split == 2 ? 2 : (split ? 1 : 0);   // 2: e4m3
    RVCR_TRY(coarse_tc_top8(h, q_dev, n, sp, cluster, &topv, &topj, &topd, &nsplit, &qnorm, &qmargin, s));
    const int pairs = nsplit * 8;
    *pairs_out = pairs;
    if (pairs > cap_pairs) {
        set_error("debug buffer too small: %d pairs per query needed", pairs);
        return RVCR_ERR_INVALID;
    }
    RVCR_CUDA(cudaMemcpyAsync(v_host, topv, (size_t)n * pairs * 4, cudaMemcpyDeviceToHost, s));
    RVCR_CUDA(cudaMemcpyAsync(j_host, topj, (size_t)n * pairs * 4, cudaMemcpyDeviceToHost, s));
    if (qnorm2_host) RVCR_CUDA(cudaMemcpyAsync(qnorm2_host, qnorm, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    if (margin_host)   // the margin the epilogue collected with and the re-check kernel selects with
        RVCR_CUDA(cudaMemcpyAsync(margin_host, qmargin, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    if (dropped_host) RVCR_CUDA(cudaMemcpyAsync(dropped_host, topd, (size_t)n * (pairs / 8) * 4, cudaMemcpyDeviceToHost, s));
    RVCR_CUDA(cudaStreamSynchronize(s));
    if (cmax_out) *cmax_out = h->cent_norm_max;
    return RVCR_OK;
}

}  // extern "C"
