// index_io.cu -- faiss `.index` (IndexIVFFlat over IndexFlat) reader / writer, straight to / from
// list-ordered device storage.
//
// Replaces faiss.read_index (rvc/infer/pipeline.py:555) and faiss.write_index
// (rvc/train/process/extract_index.py:71).  Byte layout: SURVEY Appendix B (little-endian, no
// padding; the 1-byte is_trained at offset 32 leaves every later field unaligned, so the header is
// parsed bytewise).  In the file a list is [n x d fp32 vectors][n x int64 ids]; the reader copies
// the two pieces of every list through a pinned bounce buffer into the `vecs` / `ids` arrays, so
// no host copy of the vectors is ever materialised (the reference re-reads the file AND builds a
// second host copy, big_npy, on every pipeline() call).
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace rvcr {

namespace {

struct FileReader {
    FILE* f = nullptr;
    ~FileReader() {
        if (f) fclose(f);
    }
    bool rd(void* dst, size_t n) { return fread(dst, 1, n, f) == n; }
    template <class T>
    bool get(T& v) {
        return rd(&v, sizeof(T));
    }
};

struct Header {
    int32_t d = 0;
    int64_t ntotal = 0;
    uint8_t trained = 0;
    int32_t metric = 1;
};

bool read_index_header(FileReader& r, Header& h) {
    int64_t dummy;
    if (!r.get(h.d) || !r.get(h.ntotal) || !r.get(dummy) || !r.get(dummy) || !r.get(h.trained) || !r.get(h.metric))
        return false;
    if (h.metric > 1) {
        float arg;
        if (!r.get(arg)) return false;
    }
    return true;
}

}  // namespace

// Contiguous list range [lo, hi) of shard `rank` of `world`, balanced by vector count.
void shard_range(const std::vector<int64_t>& sizes, int rank, int world, int64_t& lo, int64_t& hi) {
    const int64_t nlist = (int64_t)sizes.size();
    int64_t total = 0;
    for (int64_t s : sizes) total += s;
    auto boundary = [&](int r) -> int64_t {
        if (r <= 0) return 0;
        if (r >= world) return nlist;
        // first list index whose prefix sum reaches r * total / world
        const long double target = (long double)total * r / world;
        int64_t acc = 0;
        for (int64_t l = 0; l < nlist; ++l) {
            if ((long double)acc >= target) return l;
            acc += sizes[l];
        }
        return nlist;
    };
    if (total == 0) {  // no vectors: split lists evenly
        lo = nlist * rank / world;
        hi = nlist * (rank + 1) / world;
        return;
    }
    lo = boundary(rank);
    hi = boundary(rank + 1);
    if (hi < lo) hi = lo;
}

int read_index_file(const char* path, int device, int rank, int world, rvcr_index** out) {
    FileReader r;
    r.f = fopen(path, "rb");
    if (!r.f) {
        set_error("cannot open index file '%s'", path);
        return RVCR_ERR_IO;
    }
    char four[4];
    if (!r.rd(four, 4)) {
        set_error("'%s': truncated", path);
        return RVCR_ERR_FORMAT;
    }
    if (memcmp(four, "IwFl", 4) != 0) {
        set_error("'%s': fourcc '%.4s' is not an IndexIVFFlat ('IwFl') file", path, four);
        return RVCR_ERR_FORMAT;
    }
    Header ih;
    uint64_t nlist = 0, nprobe = 0;
    if (!read_index_header(r, ih) || !r.get(nlist) || !r.get(nprobe)) {
        set_error("'%s': truncated header", path);
        return RVCR_ERR_FORMAT;
    }
    if (ih.metric != 1) {
        set_error("'%s': metric_type %d unsupported (only METRIC_L2)", path, ih.metric);
        return RVCR_ERR_UNSUPPORTED;
    }
    if (ih.d <= 0 || nlist == 0 || nlist > (1ull << 31)) {
        set_error("'%s': implausible header (d=%d nlist=%llu)", path, ih.d, (unsigned long long)nlist);
        return RVCR_ERR_FORMAT;
    }
    if (!r.rd(four, 4) || (memcmp(four, "IxF2", 4) != 0 && memcmp(four, "IxFl", 4) != 0)) {
        set_error("'%s': quantizer fourcc '%.4s' unsupported (need IndexFlatL2)", path, four);
        return RVCR_ERR_FORMAT;
    }
    Header qh;
    uint64_t count = 0;
    if (!read_index_header(r, qh) || !r.get(count) || qh.d != ih.d || (uint64_t)qh.ntotal != nlist ||
        count != nlist * (uint64_t)ih.d) {
        set_error("'%s': inconsistent quantizer header", path);
        return RVCR_ERR_FORMAT;
    }
    std::vector<float> cent((size_t)count);
    if (!r.rd(cent.data(), (size_t)count * 4)) {
        set_error("'%s': truncated centroids", path);
        return RVCR_ERR_FORMAT;
    }
    uint8_t dm_type = 0;
    uint64_t dm_n = 0;
    if (!r.get(dm_type) || !r.get(dm_n)) {
        set_error("'%s': truncated direct map", path);
        return RVCR_ERR_FORMAT;
    }
    if (dm_n && fseek(r.f, (long)(dm_n * 8), SEEK_CUR) != 0) return RVCR_ERR_FORMAT;
    if (dm_type == 2) {
        uint64_t hn = 0;
        if (!r.get(hn) || fseek(r.f, (long)(hn * 16), SEEK_CUR) != 0) return RVCR_ERR_FORMAT;
    }
    std::vector<int64_t> sizes((size_t)nlist, 0);
    if (!r.rd(four, 4)) return RVCR_ERR_FORMAT;
    if (memcmp(four, "il00", 4) == 0) {
        uint64_t a, b;
        if (!r.get(a) || !r.get(b)) return RVCR_ERR_FORMAT;
    } else if (memcmp(four, "ilar", 4) == 0) {
        uint64_t il_nlist = 0, code_size = 0, cnt = 0;
        char lt[4];
        if (!r.get(il_nlist) || !r.get(code_size) || !r.rd(lt, 4) || !r.get(cnt)) {
            set_error("'%s': truncated inverted-list header", path);
            return RVCR_ERR_FORMAT;
        }
        if (il_nlist != nlist || code_size != 4ull * ih.d) {
            set_error("'%s': inverted lists do not match the header (nlist %llu, code_size %llu)", path,
                      (unsigned long long)il_nlist, (unsigned long long)code_size);
            return RVCR_ERR_FORMAT;
        }
        std::vector<uint64_t> raw((size_t)cnt);
        if (cnt && !r.rd(raw.data(), (size_t)cnt * 8)) return RVCR_ERR_FORMAT;
        if (memcmp(lt, "full", 4) == 0) {
            if (cnt != nlist) return RVCR_ERR_FORMAT;
            for (uint64_t l = 0; l < nlist; ++l) sizes[l] = (int64_t)raw[l];
        } else if (memcmp(lt, "sprs", 4) == 0) {
            for (uint64_t i = 0; i + 1 < cnt; i += 2) {
                if (raw[i] >= nlist) return RVCR_ERR_FORMAT;
                sizes[raw[i]] = (int64_t)raw[i + 1];
            }
        } else {
            set_error("'%s': list encoding '%.4s' unsupported", path, lt);
            return RVCR_ERR_FORMAT;
        }
    } else {
        set_error("'%s': inverted lists fourcc '%.4s' unsupported", path, four);
        return RVCR_ERR_FORMAT;
    }

    int64_t lo = 0, hi = (int64_t)nlist;
    if (world > 1) shard_range(sizes, rank, world, lo, hi);

    // the quantizer (all centroids) is replicated on every shard; only lists [lo, hi) are stored
    rvcr_index* h = nullptr;
    RVCR_TRY(rvcr_index_create(ih.d, (int64_t)nlist, device, &h));
    h->own_lo = lo;
    h->own_hi = hi;
    h->ntotal = ih.ntotal;
    h->nprobe = (int64_t)nprobe;
    auto fail = [&](int rc) {
        rvcr_index_free(h);
        return rc;
    };
    {
        int rc = rvcr_index_set_centroids(h, cent.data(), (int64_t)nlist);
        if (rc != RVCR_OK) return fail(rc);
        h->trained = ih.trained != 0;
    }
    std::vector<int64_t> local_sizes((size_t)nlist, 0);
    for (int64_t l = lo; l < hi; ++l) local_sizes[l] = sizes[l];
    {
        int64_t nst = 0;
        for (int64_t s : local_sizes) nst += s;
        int rc = rvcr_index_alloc_lists(h, local_sizes.data(), ih.ntotal);
        if (rc != RVCR_OK) return fail(rc);
        (void)nst;
    }

    // stream list payloads: two pinned bounce buffers, copies overlap the next fread
    const size_t BUF = 32u << 20;
    char* pin[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaStream_t st = nullptr;
    int prev_dev = 0;
    cudaGetDevice(&prev_dev);
    cudaSetDevice(device);
    bool ok = cudaMallocHost(&pin[0], BUF) == cudaSuccess && cudaMallocHost(&pin[1], BUF) == cudaSuccess &&
              cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming) == cudaSuccess &&
              cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
    int rc = RVCR_OK;
    if (!ok) {
        set_error("pinned staging allocation failed");
        rc = RVCR_ERR_CUDA;
    }
    int cur = 0;
    bool used[2] = {false, false};
    auto copy_piece = [&](char* dev_dst, size_t bytes) -> bool {
        size_t done = 0;
        while (done < bytes) {
            size_t n = std::min(BUF, bytes - done);
            if (used[cur] && cudaEventSynchronize(ev[cur]) != cudaSuccess) return false;
            if (!r.rd(pin[cur], n)) return false;
            if (cudaMemcpyAsync(dev_dst + done, pin[cur], n, cudaMemcpyHostToDevice, st) != cudaSuccess) return false;
            if (cudaEventRecord(ev[cur], st) != cudaSuccess) return false;
            used[cur] = true;
            cur ^= 1;
            done += n;
        }
        return true;
    };
    if (rc == RVCR_OK) {
        const size_t rowb = (size_t)ih.d * 4;
        for (int64_t l = 0; l < (int64_t)nlist && rc == RVCR_OK; ++l) {
            const int64_t n = sizes[l];
            if (n == 0) continue;
            if (l < lo || l >= hi) {
                if (fseeko(r.f, (off_t)((size_t)n * (rowb + 8)), SEEK_CUR) != 0) {
                    set_error("'%s': truncated list payload", path);
                    rc = RVCR_ERR_FORMAT;
                }
                continue;
            }
            const int64_t off = h->offsets_host[l];
            if (!copy_piece(reinterpret_cast<char*>(h->vecs) + (size_t)off * rowb, (size_t)n * rowb) ||
                !copy_piece(reinterpret_cast<char*>(h->ids) + (size_t)off * 8, (size_t)n * 8)) {
                set_error("'%s': truncated list payload or copy failure (list %lld)", path, (long long)l);
                rc = RVCR_ERR_FORMAT;
            }
        }
        if (rc == RVCR_OK && cudaStreamSynchronize(st) != cudaSuccess) {
            set_error("device copy failed while reading '%s'", path);
            rc = RVCR_ERR_CUDA;
        }
    }
    if (pin[0]) cudaFreeHost(pin[0]);
    if (pin[1]) cudaFreeHost(pin[1]);
    if (ev[0]) cudaEventDestroy(ev[0]);
    if (ev[1]) cudaEventDestroy(ev[1]);
    if (st) cudaStreamDestroy(st);
    if (rc == RVCR_OK) rc = rvcr_index_commit(h, nullptr);
    cudaSetDevice(prev_dev);
    if (rc != RVCR_OK) return fail(rc);
    *out = h;
    return RVCR_OK;
}

int write_index_file(rvcr_index* h, const char* path) {
    if (h->own_lo != 0 || h->own_hi != h->nlist) {
        set_error("write_index on a shard handle is unsupported");
        return RVCR_ERR_UNSUPPORTED;
    }
    FILE* f = fopen(path, "wb");
    if (!f) {
        set_error("cannot open '%s' for writing", path);
        return RVCR_ERR_IO;
    }
    bool ok = true;
    auto wr = [&](const void* p, size_t n) { ok = ok && fwrite(p, 1, n, f) == n; };
    auto header = [&](int32_t d, int64_t ntotal, uint8_t trained) {
        const int64_t dummy = 1 << 20;
        const int32_t metric = 1;
        wr(&d, 4); wr(&ntotal, 8); wr(&dummy, 8); wr(&dummy, 8); wr(&trained, 1); wr(&metric, 4);
    };
    const int32_t d = h->d;
    const uint64_t nlist = (uint64_t)h->nlist, nprobe = (uint64_t)h->nprobe;
    wr("IwFl", 4);
    header(d, h->ntotal, h->trained ? 1 : 0);
    wr(&nlist, 8); wr(&nprobe, 8);
    wr("IxF2", 4);
    header(d, h->trained ? (int64_t)nlist : 0, 1);
    std::vector<float> cent((size_t)nlist * d, 0.f);
    if (h->trained && cudaMemcpy(cent.data(), h->cent, cent.size() * 4, cudaMemcpyDeviceToHost) != cudaSuccess) {
        fclose(f);
        set_error("device copy failed while writing '%s'", path);
        return RVCR_ERR_CUDA;
    }
    const uint64_t count = h->trained ? nlist * (uint64_t)d : 0;
    wr(&count, 8);
    wr(cent.data(), (size_t)count * 4);
    const uint8_t dm_type = 0;
    const uint64_t zero = 0;
    wr(&dm_type, 1); wr(&zero, 8);
    wr("ilar", 4);
    const uint64_t code_size = 4ull * d;
    wr(&nlist, 8); wr(&code_size, 8);
    std::vector<uint64_t> sizes((size_t)nlist, 0);
    uint64_t non0 = 0;
    for (uint64_t l = 0; l < nlist; ++l) {
        sizes[l] = h->offsets_host.size() == nlist + 1 ? (uint64_t)(h->offsets_host[l + 1] - h->offsets_host[l]) : 0;
        non0 += sizes[l] > 0;
    }
    if (non0 > nlist / 2) {
        wr("full", 4);
        wr(&nlist, 8);
        wr(sizes.data(), (size_t)nlist * 8);
    } else {
        wr("sprs", 4);
        std::vector<uint64_t> pairs;
        for (uint64_t l = 0; l < nlist; ++l)
            if (sizes[l]) {
                pairs.push_back(l);
                pairs.push_back(sizes[l]);
            }
        const uint64_t np = pairs.size();
        wr(&np, 8);
        wr(pairs.data(), (size_t)np * 8);
    }
    // payload: device -> host in list-range chunks of <= 64 MB
    const size_t rowb = (size_t)d * 4;
    const int64_t chunk_rows = std::max<int64_t>(1, (int64_t)((64u << 20) / rowb));
    std::vector<char> hv, hi;
    int rc = RVCR_OK;
    uint64_t l = 0;
    while (l < nlist && ok && rc == RVCR_OK) {
        // take whole lists until the chunk is full (a single oversized list is taken alone)
        uint64_t l1 = l;
        int64_t rows = 0;
        while (l1 < nlist && (rows == 0 || rows + (int64_t)sizes[l1] <= chunk_rows)) rows += (int64_t)sizes[l1++];
        if (rows > 0) {
            const int64_t off = h->offsets_host[l];
            hv.resize((size_t)rows * rowb);
            hi.resize((size_t)rows * 8);
            if (cudaMemcpy(hv.data(), reinterpret_cast<char*>(h->vecs) + (size_t)off * rowb, hv.size(),
                           cudaMemcpyDeviceToHost) != cudaSuccess ||
                cudaMemcpy(hi.data(), reinterpret_cast<char*>(h->ids) + (size_t)off * 8, hi.size(), cudaMemcpyDeviceToHost) !=
                    cudaSuccess) {
                set_error("device copy failed while writing '%s'", path);
                rc = RVCR_ERR_CUDA;
                break;
            }
            size_t ro = 0;
            for (uint64_t ll = l; ll < l1; ++ll) {
                wr(hv.data() + ro * rowb, (size_t)sizes[ll] * rowb);
                wr(hi.data() + ro * 8, (size_t)sizes[ll] * 8);
                ro += (size_t)sizes[ll];
            }
        }
        l = l1;
    }
    if (fclose(f) != 0) ok = false;
    if (rc != RVCR_OK) return rc;
    if (!ok) {
        set_error("short write to '%s'", path);
        return RVCR_ERR_IO;
    }
    return RVCR_OK;
}

}  // namespace rvcr
