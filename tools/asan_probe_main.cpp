// asan_probe_main.cpp -- sanitizer harness for the .index parser (test infrastructure, built by tests/test_index_probe_asan.py):
//   g++ -x c++ -std=c++17 -g -O1 -fsanitize=address,undefined -fno-sanitize-recover=undefined \
//       -I include -I /usr/local/cuda/include codename-rvc-fork-3_b200/csrc/index_io.cu tools/asan_probe_main.cpp \
//       -L /usr/local/cuda/lib64 -lcudart -o tools/_bin/asan_probe
// index_io.cu is host-only C++ (no kernels), so the very source that ships in librvcr.so is compiled here with
// AddressSanitizer + UBSan and fed every file named on the command line through probe_index_file (the parse that
// rvcr_index_read performs before it touches the GPU).  Exit code 0 = every file parsed or was refused cleanly.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../include/rvcr.h"

struct rvcr_index;
namespace rvcr {
static char g_err[512];
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }
int probe_index_file(const char* path, rvcr_file_info* out, int64_t* sizes_out, int64_t sizes_cap);
int scan_tc_invalidate(rvcr_index*) { return 0; }
}  // namespace rvcr

// the device half of the reader is never reached by the probe; these only satisfy the linker
extern "C" {
int rvcr_index_create(int32_t, int64_t, int, rvcr_index**) { return RVCR_ERR_CUDA; }
int rvcr_index_set_centroids(rvcr_index*, const float*, int64_t) { return RVCR_ERR_CUDA; }
int rvcr_index_alloc_lists(rvcr_index*, const int64_t*, int64_t) { return RVCR_ERR_CUDA; }
int rvcr_index_commit(rvcr_index*, void*) { return RVCR_ERR_CUDA; }
void rvcr_index_free(rvcr_index*) {}
}

int main(int argc, char** argv) {
    int ok = 0, refused = 0;
    for (int i = 1; i < argc; ++i) {
        rvcr_file_info info;
        int rc = rvcr::probe_index_file(argv[i], &info, nullptr, 0);
        if (rc == RVCR_OK) {
            std::vector<int64_t> sizes((size_t)info.nlist);
            rc = rvcr::probe_index_file(argv[i], nullptr, sizes.data(), (int64_t)sizes.size());
            int64_t tot = 0;
            for (int64_t s : sizes) tot += s;
            if (rc != RVCR_OK || tot != info.nvectors) {
                fprintf(stderr, "%s: inconsistent second parse (rc %d, %lld vs %lld vectors)\n", argv[i], rc, (long long)tot,
                        (long long)info.nvectors);
                return 2;
            }
            ++ok;
        } else if (rc == RVCR_ERR_FORMAT || rc == RVCR_ERR_UNSUPPORTED || rc == RVCR_ERR_IO) {
            ++refused;
        } else {
            fprintf(stderr, "%s: unexpected code %d (%s)\n", argv[i], rc, rvcr::get_error());
            return 3;
        }
    }
    printf("%d parsed, %d refused\n", ok, refused);
    return 0;
}
