// index_io.cu -- faiss `.index` (IndexIVFFlat over IndexFlat) reader / writer, straight to / from
// list-ordered device storage; host-only file probe (rvcr_index_probe).
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
#include <exception>

#include "common.cuh"

namespace rvcr {

namespace {

// Bounded reader: the file length is taken once, every read / skip is checked against it BEFORE any allocation, so a
// corrupt or hostile header (users download .index files, tabs/download/download.py) can make the parse fail with
// RVCR_ERR_FORMAT but never allocate from an untrusted count, seek out of range, or abort.
struct FileReader {
    FILE* f = nullptr;
    int64_t size = 0, pos = 0;
    ~FileReader() {
        if (f) fclose(f);
    }
    bool open(const char* path) {
        f = fopen(path, "rb");
        if (!f) return false;
        if (fseeko(f, 0, SEEK_END) != 0) return false;
        size = (int64_t)ftello(f);
        if (size < 0 || fseeko(f, 0, SEEK_SET) != 0) return false;
        pos = 0;
        return true;
    }
    int64_t left() const { return size - pos; }
    // n * unit <= bytes left, without overflow
    bool fits(uint64_t n, uint64_t unit) const { return unit == 0 || n <= (uint64_t)left() / unit; }
    bool rd(void* dst, size_t n) {
        if ((uint64_t)n > (uint64_t)left()) return false;
        if (fread(dst, 1, n, f) != n) return false;
        pos += (int64_t)n;
        return true;
    }
    bool skip(uint64_t n) {
        if (n > (uint64_t)left()) return false;
        if (n && fseeko(f, (off_t)n, SEEK_CUR) != 0) return false;
        pos += (int64_t)n;
        return true;
    }
    bool seek(int64_t to) {
        if (to < 0 || to > size || fseeko(f, (off_t)to, SEEK_SET) != 0) return false;
        pos = to;
        return true;
    }
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

// Everything of an IndexIVFFlat file except the list payload, which stays in the file: where each list's vectors and
// ids sit is enough to stream them to the device (read_index_file) or to describe the file (rvcr_index_probe).
//   layout 0  "IwFl"  faiss >= 1.5: ArrayInvertedLists object ("ilar" / "il00"), per list [n x d fp32][n x int64], contiguous
//   layout 1  "IvFL"  legacy: nlist id vectors before the direct map, then nlist byte vectors of codes
//   layout 2  "IvFl"  oldest: same, the codes stored as float vectors (faiss impl/index_read.cpp, legacy branch)
struct ParsedIndex {
    Header ih;
    uint64_t nlist = 0, nprobe = 0;
    int layout = 0;
    char fourcc[5] = {0}, qfourcc[5] = {0};
    int dm_type = 0;
    int list_encoding = 0;                    // 0 "full", 1 "sprs", 2 "il00" (no lists), 3 legacy per-list vectors
    std::vector<float> cent;                  // nlist x d (empty when the caller does not want them)
    std::vector<int64_t> sizes;               // nlist
    int64_t nvec = 0;                         // sum of sizes
    int64_t payload_off = 0;                  // layout 0: offset of the first list's vectors
    std::vector<int64_t> ids_off, codes_off;  // legacy layouts: offset of each list's id / code data (past the length prefix)
    int64_t file_bytes = 0;
};

#define PARSE_FAIL(code, ...)   \
    do {                        \
        set_error(__VA_ARGS__); \
        return code;            \
    } while (0)

void printable(char* four) {
    for (int i = 0; i < 4; ++i)
        if (four[i] < 32 || four[i] > 126) four[i] = '?';
}

int parse_index_impl(FileReader& r, const char* path, bool want_centroids, ParsedIndex& p) {
    p.file_bytes = r.size;
    if (!r.rd(p.fourcc, 4)) PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': truncated", path);
    if (memcmp(p.fourcc, "IwFl", 4) == 0) {
        p.layout = 0;
    } else if (memcmp(p.fourcc, "IvFL", 4) == 0) {
        p.layout = 1;
    } else if (memcmp(p.fourcc, "IvFl", 4) == 0) {
        p.layout = 2;
    } else {
        printable(p.fourcc);
        PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': fourcc '%.4s' is not an IndexIVFFlat ('IwFl', legacy 'IvFL' / 'IvFl') file", path,
                   p.fourcc);
    }
    Header& ih = p.ih;
    if (!read_index_header(r, ih) || !r.get(p.nlist) || !r.get(p.nprobe))
        PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': truncated header", path);
    if (ih.metric != 1) PARSE_FAIL(RVCR_ERR_UNSUPPORTED, "'%s': metric_type %d unsupported (only METRIC_L2)", path, ih.metric);
    const uint64_t nlist = p.nlist;
    if (ih.d <= 0 || ih.ntotal < 0 || nlist == 0 || nlist > (1ull << 31))
        PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': implausible header (d=%d ntotal=%lld nlist=%llu)", path, ih.d, (long long)ih.ntotal,
                   (unsigned long long)nlist);
    if (!r.rd(p.qfourcc, 4) || (memcmp(p.qfourcc, "IxF2", 4) != 0 && memcmp(p.qfourcc, "IxFl", 4) != 0)) {
        printable(p.qfourcc);
        PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': quantizer fourcc '%.4s' unsupported (need IndexFlatL2)", path, p.qfourcc);
    }
    Header qh;
    uint64_t count = 0;
    if (!read_index_header(r, qh) || !r.get(count) || qh.d != ih.d || (uint64_t)qh.ntotal != nlist ||
        count != nlist * (uint64_t)ih.d)
        PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': inconsistent quantizer header", path);
    if (!r.fits(count, 4)) PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': truncated centroids", path);
    if (want_centroids) {
        p.cent.resize((size_t)count);
        if (!r.rd(p.cent.data(), (size_t)count * 4)) PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': truncated centroids", path);
    } else if (!r.skip(count * 4)) {
        PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': truncated centroids", path);
    }
    // nlist * 4 d bytes of centroids were just found in the file, so nlist-sized host arrays are bounded by its length
    p.sizes.assign((size_t)nlist, 0);
    const uint64_t rowb = 4ull * (uint64_t)ih.d;
    if (p.layout != 0) {
        // legacy: the id vectors of all lists come first (read_ivf_header(ivf, f, &ids))
        p.ids_off.assign((size_t)nlist, 0);
        p.codes_off.assign((size_t)nlist, 0);
        for (uint64_t l = 0; l < nlist; ++l) {
            uint64_t n = 0;
            if (!r.get(n) || !r.fits(n, 8))
                PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': truncated id vector of list %llu", path, (unsigned long long)l);
            p.ids_off[l] = r.pos;
            p.sizes[l] = (int64_t)n;
            r.skip(n * 8);
        }
    }
    uint8_t dm_type = 0;
    uint64_t dm_n = 0;
    if (!r.get(dm_type) || !r.get(dm_n) || !r.fits(dm_n, 8) || !r.skip(dm_n * 8))
        PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': truncated direct map", path);
    p.dm_type = dm_type;
    if (dm_type == 2) {
        uint64_t hn = 0;
        if (!r.get(hn) || !r.fits(hn, 16) || !r.skip(hn * 16)) PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': truncated direct map", path);
    } else if (dm_type > 2) {
        PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': direct map type %d unknown", path, (int)dm_type);
    }
    if (p.layout != 0) {
        p.list_encoding = 3;
        const uint64_t unit = p.layout == 1 ? 1 : 4;   // "IvFL": std::vector<uint8_t>, "IvFl": std::vector<float>
        for (uint64_t l = 0; l < nlist; ++l) {
            uint64_t m = 0;
            if (!r.get(m) || !r.fits(m, unit))
                PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': truncated code vector of list %llu", path, (unsigned long long)l);
            if (m * unit != (uint64_t)p.sizes[l] * rowb)
                PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': list %llu holds %lld ids but %llu code bytes (d=%d)", path, (unsigned long long)l,
                           (long long)p.sizes[l], (unsigned long long)(m * unit), ih.d);
            p.codes_off[l] = r.pos;
            r.skip(m * unit);
        }
    } else {
        char four[4];
        if (!r.rd(four, 4)) PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': truncated inverted-list header", path);
        if (memcmp(four, "il00", 4) == 0) {
            // faiss write_InvertedLists(nullptr) emits the fourcc and nothing else ("inverted lists not stored with IVF object")
            p.list_encoding = 2;
        } else if (memcmp(four, "ilar", 4) == 0) {
            uint64_t il_nlist = 0, code_size = 0, cnt = 0;
            char lt[4];
            if (!r.get(il_nlist) || !r.get(code_size) || !r.rd(lt, 4) || !r.get(cnt))
                PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': truncated inverted-list header", path);
            if (il_nlist != nlist || code_size != rowb)
                PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': inverted lists do not match the header (nlist %llu, code_size %llu)", path,
                           (unsigned long long)il_nlist, (unsigned long long)code_size);
            if (!r.fits(cnt, 8)) PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': truncated list-size table", path);
            std::vector<uint64_t> raw((size_t)cnt);
            if (cnt && !r.rd(raw.data(), (size_t)cnt * 8)) PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': truncated list-size table", path);
            if (memcmp(lt, "full", 4) == 0) {
                if (cnt != nlist)
                    PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': 'full' list-size table has %llu entries for %llu lists", path,
                               (unsigned long long)cnt, (unsigned long long)nlist);
                for (uint64_t l = 0; l < nlist; ++l) p.sizes[l] = (int64_t)raw[l];
                p.list_encoding = 0;
            } else if (memcmp(lt, "sprs", 4) == 0) {
                if (cnt & 1) PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': odd 'sprs' list-size table", path);
                for (uint64_t i = 0; i + 1 < cnt; i += 2) {
                    if (raw[i] >= nlist)
                        PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': 'sprs' entry names list %llu of %llu", path, (unsigned long long)raw[i],
                                   (unsigned long long)nlist);
                    p.sizes[raw[i]] = (int64_t)raw[i + 1];
                }
                p.list_encoding = 1;
            } else {
                PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': list encoding unsupported (need 'full' or 'sprs')", path);
            }
        } else {
            PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': inverted lists fourcc unsupported (need 'ilar' or 'il00')", path);
        }
        p.payload_off = r.pos;
    }
    // list sizes are untrusted too: non-negative, their payload inside the file (trailing bytes are ignored, like faiss)
    uint64_t tot = 0;
    const uint64_t cap = (uint64_t)r.size / (rowb + 8) + 1;
    for (uint64_t l = 0; l < nlist; ++l) {
        if (p.sizes[l] < 0 || (uint64_t)p.sizes[l] > cap)
            PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': list %llu claims %lld vectors", path, (unsigned long long)l, (long long)p.sizes[l]);
        tot += (uint64_t)p.sizes[l];
        if (tot > cap)
            PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': truncated list payload (the size table claims more vectors than the file holds)",
                       path);
    }
    p.nvec = (int64_t)tot;
    if (p.layout == 0 && !r.fits(tot, rowb + 8))
        PARSE_FAIL(RVCR_ERR_FORMAT, "'%s': truncated list payload (%llu vectors x %llu bytes expected, %lld bytes left)", path,
                   (unsigned long long)tot, (unsigned long long)(rowb + 8), (long long)r.left());
    return RVCR_OK;
}

int parse_index(FileReader& r, const char* path, bool want_centroids, ParsedIndex& p) {
    try {
        return parse_index_impl(r, path, want_centroids, p);
    } catch (const std::exception& e) {   // std::bad_alloc on a huge but self-consistent file: report, never abort
        set_error("'%s': out of host memory while parsing (%s)", path, e.what());
        return RVCR_ERR_OOM;
    }
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

// Host-only description of an index file (no CUDA call): rvcr_index_probe.  sizes_out (nullable) receives the nlist
// list sizes; world > 0 additionally reports the list range read_index_shard(rank, world) would load.
int probe_index_file(const char* path, rvcr_file_info* out, int64_t* sizes_out, int64_t sizes_cap) {
    FileReader r;
    if (!r.open(path)) {
        set_error("cannot open index file '%s'", path);
        return RVCR_ERR_IO;
    }
    ParsedIndex p;
    RVCR_TRY(parse_index(r, path, false, p));
    if (out) {
        memset(out, 0, sizeof(*out));
        out->d = p.ih.d;
        out->is_trained = p.ih.trained ? 1 : 0;
        out->ntotal = p.ih.ntotal;
        out->nlist = (int64_t)p.nlist;
        out->nprobe = (int64_t)p.nprobe;
        out->metric = p.ih.metric;
        out->layout = p.layout;
        out->direct_map_type = p.dm_type;
        out->list_encoding = p.list_encoding;
        out->nvectors = p.nvec;
        int64_t mx = 0, non0 = 0;
        for (int64_t s : p.sizes) {
            mx = std::max(mx, s);
            non0 += s > 0;
        }
        out->max_list_size = mx;
        out->nonempty_lists = non0;
        out->file_bytes = p.file_bytes;
        out->payload_offset = p.layout == 0 ? p.payload_off : -1;
        // what rvcr_index_read + rvcr_search accept: the kernels tile the dimension by 128; nprobe 1 (the tuned path) up
        // to 128 probes (multiprobe.cu)
        out->searchable = (p.ih.d % 128 == 0 && p.ih.trained && p.nprobe >= 1 && p.nprobe <= 128) ? 1 : 0;
        memcpy(out->fourcc, p.fourcc, 4);
        memcpy(out->quantizer_fourcc, p.qfourcc, 4);
    }
    if (sizes_out) {
        if (sizes_cap < (int64_t)p.nlist) {
            set_error("list-size buffer holds %lld entries, the index has %llu lists", (long long)sizes_cap,
                      (unsigned long long)p.nlist);
            return RVCR_ERR_INVALID;
        }
        for (uint64_t l = 0; l < p.nlist; ++l) sizes_out[l] = p.sizes[l];
    }
    return RVCR_OK;
}

int read_index_file(const char* path, int device, int rank, int world, rvcr_index** out) {
    FileReader r;
    if (!r.open(path)) {
        set_error("cannot open index file '%s'", path);
        return RVCR_ERR_IO;
    }
    ParsedIndex pi;
    RVCR_TRY(parse_index(r, path, true, pi));
    const Header& ih = pi.ih;
    const uint64_t nlist = pi.nlist, nprobe = pi.nprobe;
    const std::vector<float>& cent = pi.cent;
    const std::vector<int64_t>& sizes = pi.sizes;
    if (pi.layout == 0 && !r.seek(pi.payload_off)) {
        set_error("'%s': seek failed", path);
        return RVCR_ERR_IO;
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
                if (pi.layout == 0 && !r.skip((uint64_t)n * (rowb + 8))) {
                    set_error("'%s': truncated list payload", path);
                    rc = RVCR_ERR_FORMAT;
                }
                continue;
            }
            const int64_t off = h->offsets_host[l];
            if (pi.layout != 0) {
                // legacy files keep a list's ids and its vectors in two different places
                if (!r.seek(pi.codes_off[l]) || !copy_piece(reinterpret_cast<char*>(h->vecs) + (size_t)off * rowb, (size_t)n * rowb) ||
                    !r.seek(pi.ids_off[l]) || !copy_piece(reinterpret_cast<char*>(h->ids) + (size_t)off * 8, (size_t)n * 8)) {
                    set_error("'%s': truncated list payload or copy failure (list %lld)", path, (long long)l);
                    rc = RVCR_ERR_FORMAT;
                }
                continue;
            }
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
