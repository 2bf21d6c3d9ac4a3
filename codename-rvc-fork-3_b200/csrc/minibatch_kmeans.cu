// minibatch_kmeans.cu -- sklearn.cluster.MiniBatchKMeans(n_clusters=10000, batch_size=256 * cpu_count(), init="random",
// compute_labels=False).fit(big_npy).cluster_centers_ on the GPU.
//
// Replaces rvc/train/process/extract_index.py:43-56: the reference's DEFAULT build path ("Auto") first reduces any feature
// set above 200 000 rows to 10 000 mini-batch k-means centres and indexes those (nlist = 256) -- the slowest CPU step left
// in a build once train / add run on the GPU.  The algorithm is scikit-learn's (cluster/_kmeans.py, restated here):
//
//   init        n_init = 3 candidate centre sets, each k distinct rows of a random init_size subset
//               (init_size = 3 * batch, or 3 * k if that is smaller than k); the set with the lowest inertia on a random
//               validation subset of init_size rows wins
//   step        batch = `batch` rows drawn with replacement; E-step: nearest centre (the SAME coarse kernels as
//               index.search: tcgen05 bf16 + exact fp32 re-check, or the exact FFMA kernel for small problems);
//               M-step, per centre with rows in the batch:  c = (c * w + sum of its rows, in batch order) * fl(1 / (w + m)),
//               w += m  (w = rows seen so far: the per-centre learning rate)
//   reassign    whenever a centre has seen nothing, or every 10 k rows: centres with w < 0.01 * max w (at most batch / 2
//               of them) jump to distinct random rows of the batch and take the smallest remaining w
//   stop        exponentially weighted average of the per-row batch inertia (alpha = 2 batch / (n + 1)) has not improved
//               for max_no_improvement = 10 steps, or max_iter = 100 epochs
//
// The reference does not seed it (and shuffles its rows unseeded, extract_index.py:40), so no two reference runs agree and
// only distributional parity exists: tests/test_gpu_minibatch_kmeans.py compares the inertia of these centres with
// scikit-learn's on the same rows, and checks that the IVF flow downstream of them is the bit-exact path.
#include <string.h>

#include <algorithm>
#include <numeric>
#include <random>
#include <vector>

#include "common.cuh"

namespace rvcr {

int coarse_dispatch_dist(rvcr_index* h, const float* q, int64_t n, int32_t* list_out, float* dist_out, cudaStream_t s);   // api.cu

namespace {

__global__ void mbk_gather_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx, int64_t m, int d,
                                  float* __restrict__ out) {
    const int nvec = d >> 2;
    for (int64_t r = blockIdx.x; r < m; r += gridDim.x) {
        const float4* src = reinterpret_cast<const float4*>(x + idx[r] * d);
        float4* dst = reinterpret_cast<float4*>(out + r * d);
        for (int c = threadIdx.x; c < nvec; c += blockDim.x) dst[c] = src[c];
    }
}

// One CTA per centre that received rows: scikit-learn's _update_chunk_dense in fp32 -- c * w, plus its rows in ascending
// batch position, times fl(1 / (w + m)).
__global__ void __launch_bounds__(256)
mbk_update_kernel(const float* __restrict__ xb, const int32_t* __restrict__ order, const int32_t* __restrict__ seg_off,
                  const int32_t* __restrict__ touched, const float* __restrict__ w_old, int d, float* __restrict__ cent) {
    const int t = blockIdx.x;
    const int ci = touched[t];
    const int lo = seg_off[t], hi = seg_off[t + 1];
    const float w = w_old[t];
    const float alpha = __fdiv_rn(1.0f, __fadd_rn(w, (float)(hi - lo)));
    const int nvec = d >> 2;
    for (int c = threadIdx.x; c < nvec; c += blockDim.x) {
        float4 acc = reinterpret_cast<const float4*>(cent + (int64_t)ci * d)[c];
        acc.x = __fmul_rn(acc.x, w);
        acc.y = __fmul_rn(acc.y, w);
        acc.z = __fmul_rn(acc.z, w);
        acc.w = __fmul_rn(acc.w, w);
        for (int p = lo; p < hi; ++p) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(xb + (int64_t)order[p] * d) + c);
            acc.x = __fadd_rn(acc.x, v.x);
            acc.y = __fadd_rn(acc.y, v.y);
            acc.z = __fadd_rn(acc.z, v.z);
            acc.w = __fadd_rn(acc.w, v.w);
        }
        acc.x = __fmul_rn(acc.x, alpha);
        acc.y = __fmul_rn(acc.y, alpha);
        acc.z = __fmul_rn(acc.z, alpha);
        acc.w = __fmul_rn(acc.w, alpha);
        reinterpret_cast<float4*>(cent + (int64_t)ci * d)[c] = acc;
    }
}

// cent[dst[i]] = xb[src[i]]
__global__ void mbk_reassign_kernel(const float* __restrict__ xb, const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                    int m, int d, float* __restrict__ cent) {
    const int nvec = d >> 2;
    for (int r = blockIdx.x; r < m; r += gridDim.x)
        for (int c = threadIdx.x; c < nvec; c += blockDim.x)
            reinterpret_cast<float4*>(cent + (int64_t)dst[r] * d)[c] = __ldg(reinterpret_cast<const float4*>(xb + (int64_t)src[r] * d) + c);
}

}  // namespace

// h: a fresh handle made by the caller (rvcr_index_create(d, k, device)) that lends its centroid storage and scratch to the
// coarse kernels; the caller frees it.
int minibatch_kmeans(rvcr_index* h, const float* x_host, int64_t n, int64_t batch_size, int64_t max_iter,
                     int64_t max_no_improvement, double reassignment_ratio, int64_t n_init, uint64_t seed, int verbose,
                     float* centers_out, rvcr_mbk_stats* st_out) {
    const int d = h->d, device = h->device;
    const int64_t k = h->nlist;
    if (d % 4 != 0) {
        set_error("d=%d unsupported (must be a multiple of 4)", d);
        return RVCR_ERR_UNSUPPORTED;
    }
    if (n < k) {
        set_error("n_samples=%lld should be >= n_clusters=%lld.", (long long)n, (long long)k);   // scikit-learn's message
        return RVCR_ERR_INVALID;
    }
    if (k >= (1ll << 31) || batch_size < 1 || n_init < 1 || max_iter < 1) return RVCR_ERR_INVALID;
    std::lock_guard<std::mutex> lock(h->mu);
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    cudaSetDevice(device);
    cudaStream_t s = nullptr;
    const size_t rowb = (size_t)d * sizeof(float);
    const int64_t batch = std::min(batch_size, n);
    int64_t init_size = 3 * batch;
    if (init_size < k) init_size = 3 * k;
    init_size = std::min(init_size, n);
    std::mt19937_64 rng(seed);
    auto randint = [&](int64_t hi) { return (int64_t)(rng() % (uint64_t)hi); };

    DevBuf xd, xb, xv, idx_d, lab_d, dist_d, ord_d, seg_d, tch_d, w_d, best_c, rs_d, rd_d;
    float* pin = nullptr;          // host-gather mode: pinned bounce buffer for one batch / init subset
    int rc = RVCR_OK;
    auto finish = [&](int code) {
        for (DevBuf* b : {&xd, &xb, &xv, &idx_d, &lab_d, &dist_d, &ord_d, &seg_d, &tch_d, &w_d, &best_c, &rs_d, &rd_d}) b->release();
        if (pin) cudaFreeHost(pin);
        if (prev_dev >= 0) cudaSetDevice(prev_dev);
        return code;
    };
    bool failed = false;
#define MB_CHECK(expr) do { if (!failed) { int _r = (expr); if (_r != RVCR_OK) { rc = _r; failed = true; } } } while (0)
#define MB_CUDA(expr) do { if (!failed) { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); rc = RVCR_ERR_CUDA; failed = true; } } } while (0)

    // rows stay on the device when they fit next to the scratch (<= half of the free memory), else batches are gathered on
    // the host into a pinned buffer
    size_t free_b = 0, total_b = 0;
    MB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const bool resident = !failed && (size_t)n * rowb <= free_b / 2;
    const int64_t maxrows = std::max(batch, init_size);
    if (resident) {
        MB_CHECK(xd.reserve((size_t)n * rowb));
        MB_CUDA(cudaMemcpyAsync(xd.p, x_host, (size_t)n * rowb, cudaMemcpyHostToDevice, s));
    } else {
        MB_CUDA(cudaMallocHost(&pin, (size_t)maxrows * rowb));
    }
    MB_CHECK(xb.reserve((size_t)maxrows * rowb));
    MB_CHECK(xv.reserve((size_t)init_size * rowb));
    MB_CHECK(idx_d.reserve((size_t)maxrows * 8));
    MB_CHECK(lab_d.reserve((size_t)maxrows * 4));
    MB_CHECK(dist_d.reserve((size_t)maxrows * 4));
    MB_CHECK(ord_d.reserve((size_t)batch * 4));
    MB_CHECK(seg_d.reserve((size_t)(batch + 1) * 4));
    MB_CHECK(tch_d.reserve((size_t)batch * 4));
    MB_CHECK(w_d.reserve((size_t)batch * 4));
    MB_CHECK(rs_d.reserve((size_t)batch * 4));
    MB_CHECK(rd_d.reserve((size_t)batch * 4));
    MB_CHECK(best_c.reserve((size_t)k * rowb));
    if (!failed && !h->cent) MB_CUDA(cudaMalloc(&h->cent, (size_t)k * rowb));

    std::vector<int64_t> idx((size_t)maxrows);
    // rows idx[0..m) of X -> dst (device)
    auto gather = [&](int64_t m, float* dst) {
        if (resident) {
            MB_CUDA(cudaMemcpyAsync(idx_d.p, idx.data(), (size_t)m * 8, cudaMemcpyHostToDevice, s));
            if (!failed) {
                mbk_gather_kernel<<<(unsigned)std::min<int64_t>(m, 2048), 192, 0, s>>>(xd.as<float>(), idx_d.as<int64_t>(), m, d, dst);
                MB_CUDA(cudaGetLastError());
            }
        } else if (!failed) {
            for (int64_t r = 0; r < m; ++r) memcpy(pin + (size_t)r * d, x_host + idx[(size_t)r] * d, rowb);
            MB_CUDA(cudaMemcpyAsync(dst, pin, (size_t)m * rowb, cudaMemcpyHostToDevice, s));
            MB_CUDA(cudaStreamSynchronize(s));   // the bounce buffer is reused
        }
    };
    std::vector<int32_t> labels((size_t)maxrows);
    std::vector<float> dists((size_t)maxrows);
    // E-step on m rows at `rows`: labels / distances to the host, returns the inertia
    auto estep = [&](const float* rows, int64_t m) -> double {
        MB_CHECK(coarse_dispatch_dist(h, rows, m, lab_d.as<int32_t>(), dist_d.as<float>(), s));
        MB_CUDA(cudaMemcpyAsync(labels.data(), lab_d.p, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
        MB_CUDA(cudaMemcpyAsync(dists.data(), dist_d.p, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
        MB_CUDA(cudaStreamSynchronize(s));
        double in = 0.0;
        if (!failed)
            for (int64_t i = 0; i < m; ++i)
                if (labels[(size_t)i] >= 0) in += (double)dists[(size_t)i];
        return in;
    };

    // ---- init: validation subset, n_init candidate centre sets --------------------------------------------------
    for (int64_t i = 0; i < init_size; ++i) idx[(size_t)i] = randint(n);
    gather(init_size, xv.as<float>());
    double best_inertia = -1.0;
    std::vector<int64_t> sub((size_t)init_size), pick((size_t)init_size);
    for (int64_t t = 0; t < n_init && !failed; ++t) {
        for (int64_t i = 0; i < init_size; ++i) sub[(size_t)i] = randint(n);
        std::iota(pick.begin(), pick.end(), (int64_t)0);
        for (int64_t i = 0; i < k; ++i) {                       // k distinct positions of the subset (partial Fisher-Yates)
            const int64_t j = i + randint(init_size - i);
            std::swap(pick[(size_t)i], pick[(size_t)j]);
            idx[(size_t)i] = sub[(size_t)pick[(size_t)i]];
        }
        gather(k, h->cent);
        h->trained = true;
        MB_CHECK(coarse_tc_invalidate(h));
        const double inertia = estep(xv.as<float>(), init_size);
        if (verbose) fprintf(stderr, "[rvcr MiniBatchKMeans] Inertia for init %lld/%lld: %g\n", (long long)t + 1, (long long)n_init, inertia);
        if (!failed && (best_inertia < 0 || inertia < best_inertia)) {
            best_inertia = inertia;
            MB_CUDA(cudaMemcpyAsync(best_c.p, h->cent, (size_t)k * rowb, cudaMemcpyDeviceToDevice, s));
        }
    }
    MB_CUDA(cudaMemcpyAsync(h->cent, best_c.p, (size_t)k * rowb, cudaMemcpyDeviceToDevice, s));
    MB_CHECK(coarse_tc_invalidate(h));

    // ---- the optimisation -------------------------------------------------------------------------------------------
    std::vector<float> w((size_t)k, 0.f);                      // weight_sums (X dtype: float32)
    std::vector<int32_t> order((size_t)batch), seg, touched, rs, rd, head((size_t)k);
    std::vector<float> w_old;
    const int64_t n_steps = (max_iter * n) / batch;
    int64_t n_since_reassign = 0, no_improvement = 0, step = 0, n_reassigned_total = 0;
    double ewa = -1.0, ewa_min = -1.0;
    bool converged = false;
    int64_t zero_count = k;                                    // centres that have seen nothing yet
    for (step = 0; step < n_steps && !failed; ++step) {
        for (int64_t i = 0; i < batch; ++i) idx[(size_t)i] = randint(n);
        gather(batch, xb.as<float>());
        double batch_inertia = estep(xb.as<float>(), batch);
        if (failed) break;
        // _random_reassign(): every 10 k rows, or whenever a centre is still empty
        n_since_reassign += batch;
        bool random_reassign = false;
        if (zero_count > 0 || n_since_reassign >= 10 * k) {
            n_since_reassign = 0;
            random_reassign = true;
        }
        // M-step: rows grouped by label, in batch order (counting sort on the host: `batch` is a few thousand)
        std::fill(head.begin(), head.end(), 0);
        touched.clear();
        for (int64_t i = 0; i < batch; ++i) {
            const int32_t l = labels[(size_t)i];
            if (l >= 0 && head[(size_t)l]++ == 0) touched.push_back(l);
        }
        std::sort(touched.begin(), touched.end());
        seg.assign(touched.size() + 1, 0);
        for (size_t t = 0; t < touched.size(); ++t) seg[t + 1] = seg[t] + head[(size_t)touched[t]];
        for (size_t t = 0; t < touched.size(); ++t) head[(size_t)touched[t]] = seg[t];      // now: write cursor
        for (int64_t i = 0; i < batch; ++i) {
            const int32_t l = labels[(size_t)i];
            if (l >= 0) order[(size_t)head[(size_t)l]++] = (int32_t)i;
        }
        w_old.resize(touched.size());
        for (size_t t = 0; t < touched.size(); ++t) {
            float& wc = w[(size_t)touched[t]];
            w_old[t] = wc;
            if (wc == 0.f) --zero_count;
            wc += (float)(seg[t + 1] - seg[t]);
        }
        if (!touched.empty()) {
            MB_CUDA(cudaMemcpyAsync(ord_d.p, order.data(), (size_t)seg.back() * 4, cudaMemcpyHostToDevice, s));
            MB_CUDA(cudaMemcpyAsync(seg_d.p, seg.data(), seg.size() * 4, cudaMemcpyHostToDevice, s));
            MB_CUDA(cudaMemcpyAsync(tch_d.p, touched.data(), touched.size() * 4, cudaMemcpyHostToDevice, s));
            MB_CUDA(cudaMemcpyAsync(w_d.p, w_old.data(), w_old.size() * 4, cudaMemcpyHostToDevice, s));
            if (!failed) {
                const int threads = std::min(256, std::max(32, ((d / 4 + 31) / 32) * 32));
                mbk_update_kernel<<<(unsigned)touched.size(), threads, 0, s>>>(xb.as<float>(), ord_d.as<int32_t>(), seg_d.as<int32_t>(),
                                                                              tch_d.as<int32_t>(), w_d.as<float>(), d, h->cent);
                MB_CUDA(cudaGetLastError());
            }
        }
        // reassignment of centres with very low weight
        if (random_reassign && reassignment_ratio > 0) {
            const float wmax = *std::max_element(w.begin(), w.end());
            rd.clear();
            for (int64_t c = 0; c < k; ++c)
                if (w[(size_t)c] < (float)reassignment_ratio * wmax) rd.push_back((int32_t)c);
            const size_t cap = (size_t)(0.5 * (double)batch);
            if (rd.size() > cap) {                              // keep the `cap` lightest
                std::stable_sort(rd.begin(), rd.end(), [&](int32_t a, int32_t b) { return w[(size_t)a] < w[(size_t)b]; });
                rd.resize(cap);
            }
            if (!rd.empty()) {
                rs.resize((size_t)batch);
                std::iota(rs.begin(), rs.end(), 0);
                for (size_t i = 0; i < rd.size(); ++i) std::swap(rs[i], rs[i + (size_t)randint(batch - (int64_t)i)]);
                rs.resize(rd.size());
                MB_CUDA(cudaMemcpyAsync(rs_d.p, rs.data(), rs.size() * 4, cudaMemcpyHostToDevice, s));
                MB_CUDA(cudaMemcpyAsync(rd_d.p, rd.data(), rd.size() * 4, cudaMemcpyHostToDevice, s));
                if (!failed) {
                    const int threads = std::min(256, std::max(32, ((d / 4 + 31) / 32) * 32));
                    mbk_reassign_kernel<<<(unsigned)std::min<size_t>(rd.size(), 1024), threads, 0, s>>>(xb.as<float>(), rs_d.as<int32_t>(),
                                                                                                      rd_d.as<int32_t>(), (int)rd.size(), d, h->cent);
                    MB_CUDA(cudaGetLastError());
                }
                std::vector<char> is_r((size_t)k, 0);
                for (int32_t c : rd) is_r[(size_t)c] = 1;
                float wmin = -1.f;
                for (int64_t c = 0; c < k; ++c)
                    if (!is_r[(size_t)c] && (wmin < 0.f || w[(size_t)c] < wmin)) wmin = w[(size_t)c];
                if (wmin < 0.f) wmin = 0.f;
                for (int32_t c : rd) {
                    if (w[(size_t)c] == 0.f && wmin > 0.f) --zero_count;
                    w[(size_t)c] = wmin;
                }
                n_reassigned_total += (int64_t)rd.size();
                if (verbose) fprintf(stderr, "[rvcr MiniBatchKMeans] Reassigning %zu cluster centers.\n", rd.size());
            }
        }
        MB_CUDA(cudaStreamSynchronize(s));                      // host vectors are reused by the next step
        MB_CHECK(coarse_tc_invalidate(h));
        // _mini_batch_convergence()
        batch_inertia /= (double)batch;
        if (step == 0) continue;                                // the first inertia belongs to the initial centres
        if (ewa < 0) {
            ewa = batch_inertia;
        } else {
            double alpha = (double)batch * 2.0 / ((double)n + 1.0);
            if (alpha > 1.0) alpha = 1.0;
            ewa = ewa * (1.0 - alpha) + batch_inertia * alpha;
        }
        if (verbose && (step % 50 == 0)) fprintf(stderr, "[rvcr MiniBatchKMeans] step %lld/%lld: mean batch inertia %g, ewa inertia %g\n",
                                                 (long long)step + 1, (long long)n_steps, batch_inertia, ewa);
        if (ewa_min < 0 || ewa < ewa_min) {
            no_improvement = 0;
            ewa_min = ewa;
        } else {
            ++no_improvement;
        }
        if (max_no_improvement > 0 && no_improvement >= max_no_improvement) {
            converged = true;
            ++step;
            break;
        }
    }
    MB_CUDA(cudaMemcpyAsync(centers_out, h->cent, (size_t)k * rowb, cudaMemcpyDeviceToHost, s));
    MB_CUDA(cudaStreamSynchronize(s));
    if (st_out && !failed) {
        st_out->n_steps = step;
        st_out->max_steps = n_steps;
        st_out->converged = converged ? 1 : 0;
        st_out->ewa_inertia = ewa;
        st_out->init_inertia = best_inertia;
        st_out->batch_size = batch;
        st_out->init_size = init_size;
        st_out->rows_resident = resident ? 1 : 0;
        st_out->n_reassigned = n_reassigned_total;
        st_out->kernels_launched = h->stats.kernels_launched;
    }
#undef MB_CHECK
#undef MB_CUDA
    finish(rc);
    return rc;
}

}  // namespace rvcr
