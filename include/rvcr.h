/*
 * rvcr.h -- C ABI of the B200-native RVC feature-retrieval library (librvcr.so).
 *
 * The reference (codename0og/codename-rvc-fork-3) reaches this path only through the faiss
 * Python object and eleven lines of numpy/torch; it has no FFI of its own.  Each entry point
 * below names the reference call it replaces (paths relative to the reference root).  Plain
 * pointers and sizes only: no torch / numpy types cross this boundary.  The Python host side
 * (codename-rvc-fork-3_b200/) binds it with ctypes; INTEGRATION.md shows the stub a reference
 * maintainer would add.
 *
 * Conventions
 *   - every function returns RVCR_OK (0) or a negative RVCR_ERR_* code; rvcr_last_error() returns
 *     a thread-local message for the last failure.  Nothing aborts, nothing falls back to a CPU
 *     implementation: without a CUDA device the compute entry points fail with RVCR_ERR_CUDA.
 *   - "_dev" pointers are device memory on the index's device, "_host" pointers are host memory.
 *   - `stream` is a cudaStream_t passed as void* (0 = the legacy default stream).  Compute entry
 *     points enqueue on it and do not synchronise unless stated.
 *   - a handle may be used from several threads; calls on one handle are serialised internally.
 *   - distances are squared L2, in the canonical fp32 evaluation order of faiss fvec_L2sqr
 *     (8 lanes, fused multiply-add, (m0+m1)+(m2+m3) reduction) -- see DESIGN.md "Numerics".
 */
#ifndef RVCR_H_
#define RVCR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVCR_OK 0
#define RVCR_ERR_INVALID (-1)      /* bad argument                                   */
#define RVCR_ERR_IO (-2)           /* file could not be opened / read / written      */
#define RVCR_ERR_FORMAT (-3)       /* not an IVF-Flat .index file we understand       */
#define RVCR_ERR_CUDA (-4)         /* CUDA runtime error, or no usable device         */
#define RVCR_ERR_UNSUPPORTED (-5)  /* valid request outside this library's scope      */
#define RVCR_ERR_STATE (-6)        /* e.g. search on an untrained index               */
#define RVCR_ERR_OOM (-7)

#define RVCR_VERSION 100

/* phases timed when option "profile" is 1 (rvcr_get_phase_ms) */
#define RVCR_PHASE_COARSE_PREP 0    /* fp32 -> bf16 operands, norms                         */
#define RVCR_PHASE_COARSE_MAIN 1    /* the query x centroid contraction (tcgen05 or FFMA)   */
#define RVCR_PHASE_COARSE_RECHECK 2 /* exact re-check of tensor-core candidates + fallback  */
#define RVCR_PHASE_GROUP 3          /* sort queries by probed list, build work items        */
#define RVCR_PHASE_SCAN 4           /* inverted-list scan + top-8                           */
#define RVCR_PHASE_BLEND 5          /* gather + 1/score^2 weights + index_rate blend        */
#define RVCR_PHASE_SCAN_RECHECK 6  /* (inside SCAN) exact fp32 re-check of the tensor-core scan's candidates */
#define RVCR_PHASE_SCAN_FALLBACK 7 /* (inside SCAN) brute-force exact scan of overflowed queries            */
#define RVCR_NUM_PHASES 8

#if defined(__GNUC__)
#define RVCR_API __attribute__((visibility("default")))
#else
#define RVCR_API
#endif

typedef struct rvcr_index rvcr_index;

typedef struct rvcr_info {
    int32_t d;            /* vector dimension (768 for v2 / ContentVec, 256 for v1)             */
    int32_t is_trained;   /* quantizer holds centroids                                          */
    int64_t ntotal;       /* faiss Index::ntotal                                                */
    int64_t nlist;        /* number of inverted lists == number of centroids (always the whole index) */
    int64_t nprobe;       /* value persisted in the file (extract_index.py:64 sets 1)           */
    int64_t nstored;      /* vectors stored in this handle's lists (== ntotal unless sharded)   */
    int64_t owned_begin;  /* a shard stores lists [owned_begin, owned_end) only; 0 / nlist      */
    int64_t owned_end;    /*   for a complete index                                             */
    int32_t device;
    int32_t metric;       /* 1 = METRIC_L2                                                      */
} rvcr_info;

/* What an .index file holds, read on the host without touching a GPU (rvcr_index_probe). */
typedef struct rvcr_file_info {
    int32_t d;
    int32_t is_trained;
    int64_t ntotal;
    int64_t nlist;
    int64_t nprobe;           /* as persisted (extract_index.py:64 writes 1)                                   */
    int32_t metric;           /* 1 = METRIC_L2 (anything else is refused)                                      */
    int32_t layout;           /* 0 = "IwFl" (faiss >= 1.5), 1 = legacy "IvFL", 2 = legacy "IvFl"               */
    int32_t direct_map_type;  /* 0 none, 1 array, 2 hash table (skipped by the reader)                         */
    int32_t list_encoding;    /* 0 "full", 1 "sprs", 2 "il00" (no lists), 3 legacy per-list vectors            */
    int64_t nvectors;         /* sum of the list sizes                                                         */
    int64_t max_list_size;
    int64_t nonempty_lists;
    int64_t file_bytes;
    int64_t payload_offset;   /* layout 0: offset of the first list's vectors; -1 for the legacy layouts         */
    int32_t searchable;       /* 1 if rvcr_index_read + rvcr_search accept it as is (L2, d % 128 == 0, nprobe 1..128) */
    char fourcc[8];           /* NUL-terminated                                                                */
    char quantizer_fourcc[8];
    int32_t reserved32;
    int64_t reserved[4];
} rvcr_file_info;

typedef struct rvcr_stats {
    int64_t calls_search;       /* search / retrieve_blend calls served                          */
    int64_t queries;            /* queries processed                                             */
    int64_t kernels_launched;   /* kernels of this library launched so far (all entry points)    */
    int64_t coarse_tc_queries;  /* queries whose coarse step ran on the tcgen05 path             */
    int64_t coarse_candidates;  /* candidates exactly re-checked after the tcgen05 path           */
    int64_t coarse_fallback;    /* queries re-done by the exact FFMA kernel (margin overflow)     */
    int64_t scan_items;         /* (list, query-block) work items of the last scan               */
    int64_t scan_tc_candidates; /* rows exactly re-checked after the tcgen05 tf32 list scan       */
    int64_t scan_tc_fallback;   /* queries re-done by the brute-force exact scan (margin overflow) */
    int64_t reserved[6];
} rvcr_stats;

RVCR_API const char* rvcr_last_error(void);
RVCR_API int rvcr_version(void);
/* number of CUDA devices visible; RVCR_ERR_CUDA if the runtime cannot be initialised */
RVCR_API int rvcr_device_count(int* count_out);

/* ---------------------------------------------------------------- lifecycle / persistence -- */

/* faiss.index_factory(d, "IVF{nlist},Flat")            rvc/train/process/extract_index.py:62 */
RVCR_API int rvcr_index_create(int32_t d, int64_t nlist, int device, rvcr_index** out);
/* faiss.read_index(path)                                rvc/infer/pipeline.py:555
 * Streams the file straight into list-ordered device storage; no host copy of the vectors. */
RVCR_API int rvcr_index_read(const char* path, int device, rvcr_index** out);
/* Describe an .index file without a GPU: header fields, layout, list-size statistics; `list_sizes_host` (nullable, room for
 * `nlist_cap` entries) receives the per-list vector counts, e.g. to plan shard ranges before loading.  Parses exactly what
 * rvcr_index_read parses (same refusals, same codes); every count in the file is checked against the file length before it
 * is trusted, so a corrupt or hostile file yields RVCR_ERR_FORMAT, never a crash.  faiss.read_index, pipeline.py:555. */
RVCR_API int rvcr_index_probe(const char* path, rvcr_file_info* info_out, int64_t* list_sizes_host, int64_t nlist_cap);
/* List range [*lo_out, *hi_out) that shard `rank` of `world` loads (rvcr_index_read_shard): contiguous, balanced by
 * vector count.  Host only. */
RVCR_API int rvcr_shard_range(const int64_t* list_sizes_host, int64_t nlist, int rank, int world, int64_t* lo_out,
                              int64_t* hi_out);
/* Same file, but stores only the lists of shard `rank` of `world` (contiguous list-id ranges balanced
 * by vector count).  The quantizer (all centroids, a few hundred MB at most) is replicated on every
 * shard, so rvcr_assign on any shard gives the global assignment; list ids are global. */
RVCR_API int rvcr_index_read_shard(const char* path, int device, int rank, int world, rvcr_index** out);
/* Declare that this handle stores only lists [owned_begin, owned_end) of the index (what
 * rvcr_index_read_shard sets up from a file): adds drop rows of other lists, the *_assigned entry
 * points skip queries assigned to other lists. */
RVCR_API int rvcr_index_set_shard(rvcr_index* h, int64_t owned_begin, int64_t owned_end);
/* faiss.write_index(index, path)                        rvc/train/process/extract_index.py:71 */
RVCR_API int rvcr_index_write(rvcr_index* h, const char* path);
RVCR_API void rvcr_index_free(rvcr_index* h);
RVCR_API int rvcr_index_info(rvcr_index* h, rvcr_info* out);
/* faiss.extract_index_ivf(index).nprobe = v             rvc/train/process/extract_index.py:63-64
 * The value is persisted and honoured by search: 1 takes the tuned single-list path, 2..128 (a file made by another
 * fork) the multi-probe path -- nprobe nearest centroids by canonical distance, lists scanned in coarse-rank
 * order into one top-8 with faiss's strict replacement rule (csrc/multiprobe.cu).  More than 128: RVCR_ERR_UNSUPPORTED. */
RVCR_API int rvcr_index_set_nprobe(rvcr_index* h, int64_t nprobe);
/* Quantizer centroids (nlist x d, row-major fp32, host memory). `set` marks the index trained;
 * it is what index.train() produces (extract_index.py:65). */
RVCR_API int rvcr_index_set_centroids(rvcr_index* h, const float* centroids_host, int64_t nlist);
RVCR_API int rvcr_index_get_centroids(rvcr_index* h, float* centroids_host);
/* per-list sizes, nlist x int64, host memory */
RVCR_API int rvcr_index_list_sizes(rvcr_index* h, int64_t* sizes_host);

/* -------------------------------------------------------------------------------- build --- */

/* index.train(x)                                        rvc/train/process/extract_index.py:65
 * faiss Clustering::train as IndexIVF::train runs it (niter = 10 [option "kmeans_niter"], seed 1234,
 * subsample to 256 points per centroid with rand_perm(seed), init = x[rand_perm(seed + 1)[:nlist]], fp32
 * means accumulated in point order, split_clusters for empty clusters).  x_host: n x d fp32, host memory.
 * The assignment of every iteration runs on the same kernels as rvcr_assign; the result is deterministic. */
RVCR_API int rvcr_index_train(rvcr_index* h, const float* x_host, int64_t n);
/* index.add(x)     extract_index.py:67-69 : nearest-centroid assignment + list append; ids are
 * ntotal .. ntotal+n-1; per-list order is ascending id.  x may be host or device memory. */
RVCR_API int rvcr_index_add(rvcr_index* h, const float* x, int64_t n, int x_on_device, void* stream);
/* The same append with the assignment supplied by the caller (int32 list ids, device). */
RVCR_API int rvcr_index_add_preassigned(rvcr_index* h, const float* x_dev, int64_t n, const int32_t* assign_dev,
                               void* stream);
/* Allocate list storage for the given final list sizes (host, nlist x int64) and mark every
 * list full: the caller fills the storage in place through rvcr_index_device_ptrs.  Used for
 * zero-copy construction of very large indices (98 GB at 32M x 768). */
RVCR_API int rvcr_index_alloc_lists(rvcr_index* h, const int64_t* sizes_host, int64_t ntotal);
/* Raw device views of the list-ordered storage: vectors (nstored x d fp32), ids (nstored x
 * int64), list offsets (nlist+1 x int64), centroids (nlist x d fp32).  Any may be NULL. */
RVCR_API int rvcr_index_device_ptrs(rvcr_index* h, float** vecs_dev, int64_t** ids_dev, int64_t** offsets_dev,
                           float** centroids_dev);
/* Call after writing storage obtained from rvcr_index_device_ptrs (refreshes derived data). */
RVCR_API int rvcr_index_commit(rvcr_index* h, void* stream);

/* ------------------------------------------------------------------------------- search --- */

/* quantizer->assign(n, x): nearest centroid (list id), lowest id on ties; optional
 * canonical distance to it.                  faiss IndexFlatL2::search(k=1), pipeline.py:499 */
RVCR_API int rvcr_assign(rvcr_index* h, const float* x_dev, int64_t n, int32_t* list_out_dev,
                float* dist_out_dev /* may be NULL */, void* stream);
/* index.search(x, k=8)                                        rvc/infer/pipeline.py:499
 * D: n x k fp32 ascending, I: n x k int64, padded with (FLT_MAX, -1). k must be 8. */
RVCR_API int rvcr_search(rvcr_index* h, const float* q_dev, int64_t n, int32_t k, float* D_dev, int64_t* I_dev,
                void* stream);
/* The faiss-shaped host call: H2D copy, search, D2H copy, synchronises. */
RVCR_API int rvcr_search_host(rvcr_index* h, const float* q_host, int64_t n, int32_t k, float* D_host,
                     int64_t* I_host);
/* _retrieve_speaker_embeddings(feats, index, big_npy, index_rate)   pipeline.py:497-507
 * search(k=8) + 1/score^2 weights + gather + weighted sum + index_rate blend, written IN PLACE
 * into feats_dev (n x d fp32).  D_dev / I_dev (n x 8) are optional outputs. */
RVCR_API int rvcr_retrieve_blend(rvcr_index* h, float* feats_dev, int64_t n, double index_rate, float* D_dev,
                        int64_t* I_dev, void* stream);
/* The path plus the step after it, one call: retrieve_blend, then the 2x nearest
 * upsample along time and the pitch-protect mix against the UN-blended features
 *     feats = F.interpolate(feats.permute(0,2,1), scale_factor=2).permute(0,2,1)      pipeline.py:463-465
 *     pitchff[pitchf > 0] = 1; pitchff[pitchf < 1] = protect                          pipeline.py:476-478
 *     feats = feats * pitchff[..., None] + feats0 * (1 - pitchff[..., None])          pipeline.py:479-481
 * done in the blend kernel's epilogue.  feats_dev (n x d) is read only (it is the reference's
 * feats0 clone, pipeline.py:455); out_dev is 2n x d; pitchf_dev holds 2n values, or NULL for
 * upsample only (no pitch guidance, or protect >= 0.5: the caller applies that condition). */
RVCR_API int rvcr_retrieve_blend_upsample(rvcr_index* h, const float* feats_dev, int64_t n, double index_rate,
                                 const float* pitchf_dev, double protect, float* out_dev, float* D_dev,
                                 int64_t* I_dev, void* stream);
/* The same post step alone, for calls that skipped retrieval (no index / index_rate 0,
 * pipeline.py:456-461): feats_dev, feats0_dev n x d, out_dev 2n x d.  pitchf_dev and feats0_dev are
 * both given (mix) or both NULL (upsample only). */
RVCR_API int rvcr_upsample_protect(const float* feats_dev, const float* feats0_dev, const float* pitchf_dev, int64_t n,
                          int32_t d, double protect, float* out_dev, int device, void* stream);
/* Same with host buffers (pinned or pageable): H2D, retrieve_blend, D2H, synchronises. */
RVCR_API int rvcr_retrieve_blend_host(rvcr_index* h, float* feats_host, int64_t n, double index_rate);
/* index.reconstruct_n(i0, ni)  ->  big_npy                      rvc/infer/pipeline.py:556 */
RVCR_API int rvcr_reconstruct_n(rvcr_index* h, int64_t i0, int64_t ni, float* out, int out_on_device);

/* ------------------------------------------------------------------- multi-GPU (shards) --- */

/* Global argmin over G shards' (distance, list id) pairs, lowest list id on equal distance.
 * dist_all / list_all are G x n (shard-major, as produced by an all-gather of rvcr_assign
 * outputs).  Output: n global list ids. */
RVCR_API int rvcr_merge_coarse(const float* dist_all_dev, const int32_t* list_all_dev, int32_t G, int64_t n,
                      int32_t* list_out_dev, int device, void* stream);
/* retrieve_blend for the queries whose list (global id in assign_dev) this shard holds; rows
 * of other queries are left untouched and their D / I rows are set to (FLT_MAX, -1). */
RVCR_API int rvcr_retrieve_blend_assigned(rvcr_index* h, float* feats_dev, int64_t n, const int32_t* assign_dev,
                                 double index_rate, float* D_dev, int64_t* I_dev, void* stream);
/* The search half of the above: local top-8 (D, I) for owned queries, (FLT_MAX, -1) rows for the rest. */
RVCR_API int rvcr_search_assigned(rvcr_index* h, const float* q_dev, int64_t n, const int32_t* assign_dev, float* D_dev,
                                  int64_t* I_dev, void* stream);
/* ---- peer memory (NVLink / NVSwitch): the two bulk exchanges of a sharded step without a collective ----
 * Every rank allocates its feature buffer with rvcr_peer_alloc (cudaMalloc + cudaIpcGetMemHandle), exchanges the
 * 64-byte handles out of band (torch.distributed) and maps the others' buffers with rvcr_peer_open.  The owner
 * shard then PULLS the query rows assigned to its lists out of their home ranks' buffers (rvcr_peer_pull_rows:
 * row r of out_dev = bases[src_rank[r]] + src_row[r] * d), and rvcr_retrieve_blend_assigned_push is
 * rvcr_retrieve_blend_assigned whose blend epilogue PUSHES each blended row into its home rank's buffer
 * (dst_bases[dst_rank[r]] + dst_row[r] * d) instead of writing rows_dev -- compute and combine in one kernel.
 * bases arrays are device arrays of G pointers (own buffer included). */
RVCR_API int rvcr_peer_alloc(int device, int64_t bytes, void** ptr_out, unsigned char* handle64_out);
RVCR_API int rvcr_peer_open(int device, const unsigned char* handle64, void** ptr_out);
RVCR_API int rvcr_peer_close(int device, void* ptr);
RVCR_API int rvcr_peer_free(int device, void* ptr);
RVCR_API int rvcr_peer_pull_rows(const float* const* bases_dev, const int32_t* src_rank_dev, const int32_t* src_row_dev,
                        int64_t m, int32_t d, float* out_dev, int device, void* stream);
RVCR_API int rvcr_retrieve_blend_assigned_push(rvcr_index* h, const float* rows_dev, int64_t m, const int32_t* assign_dev,
                                      double index_rate, float* const* dst_bases_dev, const int32_t* dst_rank_dev,
                                      const int32_t* dst_row_dev, void* stream);
/* model.final_proj(feats[0])                                rvc/infer/pipeline.py:451-453 (v1 models)
 * The step before the path as the query producer: out_dev (n x d_out) = x_dev (n x d_in) W^T + b with W (d_out x d_in,
 * the nn.Linear weight as torch stores it) and b (d_out, nullable) on the device; fp32 FFMA, products accumulated in
 * ascending k.  d_in % 16 == 0, d_out % 64 == 0 (768 -> 256 in RVC).  Floating point: equals the reference's cuBLAS
 * sgemm within 1e-5 relative, not bit for bit (a GEMM library's summation order is unspecified).
 * rvcr_final_proj_retrieve_blend chains it with rvcr_retrieve_blend on the same stream: the projected rows are written to
 * feats_out_dev (n x d_out, d_out == the index's d) and blended there in place; D_dev / I_dev as in rvcr_retrieve_blend. */
RVCR_API int rvcr_final_proj(const float* x_dev, int64_t n, int32_t d_in, const float* w_dev, const float* b_dev, int32_t d_out,
                             float* out_dev, int device, void* stream);
RVCR_API int rvcr_final_proj_retrieve_blend(rvcr_index* h, const float* x_dev, int64_t n, int32_t d_in, const float* w_dev,
                                            const float* b_dev, float* feats_out_dev, double index_rate, float* D_dev,
                                            int64_t* I_dev, void* stream);
/* ---- the device-driven sharded step: inboxes in peer memory, no host synchronisation, no collective ----
 * Each rank allocates (rvcr_peer_alloc, zero-filled) and peer-maps an INBOX of RVCR_PEER_INBOX_BYTES(cap) bytes:
 *   +RVCR_PEER_INBOX_COUNT_OFF  int32   entries filed in the current step
 *   +RVCR_PEER_INBOX_FLAGS_OFF  int32[] arrival flag of every rank (rvcr_peer_barrier)
 *   +RVCR_PEER_INBOX_HEADER     int32[cap] list id | int32[cap] home rank | int32[cap] row in the home rank's feature buffer
 * rvcr_peer_dispatch (home rank, after rvcr_assign): files query i -- list assign_dev[i], row row0 + i of this rank's
 *   feature buffer -- into the inbox of the rank that owns the list (bounds_dev[r - 1] = first list of rank r); slots are
 *   reserved with one remote atomicAdd per CTA and owner, entries are plain NVLink stores.
 * rvcr_peer_barrier: stream-ordered barrier across the G ranks (flag stores with release, spin with acquire, system
 *   scope); `epoch` must increase by one per call, identically on every rank.  After it, everything the other ranks
 *   enqueued before their call of the same epoch is visible.  Every rank must call it.
 * rvcr_retrieve_blend_inbox_push (owner): seals the unused tail of the own inbox and runs scan + top-8 + gather + blend on
 *   the filed queries: the scan kernels load each work item's query rows straight out of the home ranks' buffers
 *   (feat_bases_dev, NVLink loads) and the blend epilogue pushes every blended row back to its home buffer; clears the
 *   entry count for the next step.  rows_scratch_dev: NULL, or a cap x d buffer to PULL the rows into first (one extra pass).
 * A step = assign, dispatch, barrier, inbox_push, barrier; cap must be >= the number of queries all ranks file per step. */
#define RVCR_PEER_INBOX_COUNT_OFF 0
#define RVCR_PEER_INBOX_FLAGS_OFF 256
#define RVCR_PEER_INBOX_HEADER 1024
#define RVCR_PEER_INBOX_BYTES(cap) (RVCR_PEER_INBOX_HEADER + 12 * (int64_t)(cap))
RVCR_API int rvcr_peer_dispatch(const int32_t* assign_dev, int64_t n, int64_t row0, int32_t my_rank, int32_t G,
                                const int32_t* bounds_dev, void* const* inbox_bases_dev, int64_t cap, int device, void* stream);
RVCR_API int rvcr_peer_barrier(void* const* inbox_bases_dev, int32_t my_rank, int32_t G, int32_t epoch, int device, void* stream);
RVCR_API int rvcr_retrieve_blend_inbox_push(rvcr_index* h, void* inbox_dev, int64_t cap, float* rows_scratch_dev,
                                            double index_rate, float* const* feat_bases_dev, void* stream);
/* Merge G shards' top-8 candidate lists (G x n x 8, after an all-gather) into the global
 * top-8 by (distance, id). */
RVCR_API int rvcr_merge_topk(const float* D_all_dev, const int64_t* I_all_dev, int32_t G, int64_t n, int32_t k,
                    float* D_out_dev, int64_t* I_out_dev, int device, void* stream);

/* --------------------------------------------------------------------- options / stats ---- */

/* keys: "coarse_mode" 0 = auto (tcgen05 from "coarse_tc_min_flops" = 6.4e7 flop per call, else the exact FFMA kernel:
 *                     same bits, lower fixed cost), 1 = exact FFMA kernel, 2 = tcgen05 bf16 + exact re-check,
 *                     3 = tcgen05 3-way split bf16 + exact re-check
 *                     4 = tcgen05 e4m3 operands + exact re-check: twice the tensor rate, a much wider margin -- pays off
 *                         when few centroids lie near the best one, else queries fall back to the exact kernel; results
 *                         are identical either way (auto picks it only through the probe below)
 *       "coarse_f8_auto" 1 (default) / 0: in auto mode, probe coarse_mode 4 on the first 8 192 queries of the first call of
 *                     >= 65 536 queries per centroid set and keep it when <= 0.5 % of the probe fell back (a later call that
 *                     falls back for > 2 % of its queries switches it off again);
 *                     "coarse_f8_decision" (read only) tells the outcome: 0 not probed, 1 e4m3, -1 bf16
 *       "nan_on_zero_score" 1 (default, reference behaviour: exact-match query -> NaN row)
 *       "profile" 0/1 record per-phase CUDA events (see rvcr_get_phase_ms)
 *       "tc_cluster" 1 = one CTA per SM, 2 (default) = CTA pairs issuing tcgen05.mma.cta_group::2
 *       "tc_resident" 0 / 6 / 8 / 9 (default) / 10: 64-element k-blocks of the query operand the coarse tcgen05 kernel pins
 *                     in shared memory per work item (0 streams both operands through the L2)
 *       "fused_blend" 1 (default): with the tcgen05 scan the exact re-check kernel also applies the 1/score^2 weights and
 *                     the index_rate blend (the 8 neighbour rows are still in cache); 0 = separate gather_blend kernel
 *       "scan_impl" 0 (default) auto: 5 for calls of >= "scan_tc_min_queries" (1024) queries, else the exact FFMA scan;
 *                   5 tcgen05 tf32 scan + exact re-check; 1-4 the exact FFMA scan (one kernel since round 2; the
 *                   values are kept for callers that set them)
 *       "kmeans_niter" 10 (default): Lloyd iterations of rvcr_index_train (0 = the initial centroids only) */
RVCR_API int rvcr_set_option(rvcr_index* h, const char* key, int64_t value);
RVCR_API int rvcr_get_option(rvcr_index* h, const char* key, int64_t* value_out);
RVCR_API int rvcr_get_stats(rvcr_index* h, rvcr_stats* out);
/* With option "profile" = 1 every search-type call records CUDA events (on the caller's stream)
 * around each phase; this synchronises on them and returns the last call's per-phase device time
 * in milliseconds (RVCR_NUM_PHASES floats, -1 for phases that did not run). */
RVCR_API int rvcr_get_phase_ms(rvcr_index* h, float* ms_out);
/* Diagnostics for tests: run only the tensor-core stage of the coarse quantizer (coarse_tc.cu) and return its raw
 * output to host buffers: per query `*pairs_out` = 8 * L approximate (v = ||c||^2 - 2<q,c>, centroid id) pairs (L lists:
 * centroid splits x the epilogue's column groups; each list holds up to 8 centroids whose v is within the query's
 * margin of the smallest v the list saw, unsorted, free slots v = +inf / id = -1), capacity cap_pairs per query; plus
 * ||q||^2 per query, max ||c||, (margin_host, n floats, nullable) the margin the epilogue collects with and the exact
 * re-check selects with -- two approximate values further apart than that are ordered like the true distances (built
 * from the exact norms of what the low-precision operands drop, coarse_tc.cu tc_margin) -- and (dropped_host, n x L floats,
 * nullable) per list the smallest v it had to let go of (+inf: nothing).  split: 0 = bf16, 1 = 3-way split bf16,
 * 2 = e4m3; cluster: 1 | 2. */
RVCR_API int rvcr_debug_coarse_tc(rvcr_index* h, const float* q_dev, int64_t n, int32_t split, int32_t cluster,
                                  int32_t cap_pairs, float* v_host, int32_t* j_host, int32_t* pairs_out,
                                  float* qnorm2_host, float* cmax_out, float* margin_host, float* dropped_host, void* stream);

/* MiniBatchKMeans(n_clusters=10000, batch_size=256 * cpu_count(), init="random", compute_labels=False).fit(big_npy)
 * .cluster_centers_                                         rvc/train/process/extract_index.py:43-56   
 * scikit-learn's mini-batch k-means with the E-step on the coarse kernels: n_init random centre sets scored on a
 * validation subset, steps of `batch_size` rows drawn with replacement, per-centre learning rates, random reassignment of
 * starved centres, early stop on the smoothed inertia (csrc/minibatch_kmeans.cu).  x_host: n x d float32 host rows;
 * centers_out_host: k x d.  The reference does not seed it: results are comparable in distribution only.  seed = 0 draws
 * one from std::random_device.  Defaults of the estimator: max_iter 100, max_no_improvement 10, reassignment_ratio 0.01,
 * n_init 3 (init="random"). */
typedef struct rvcr_mbk_stats {
    int64_t n_steps, max_steps, batch_size, init_size, n_reassigned, kernels_launched;
    int32_t converged, rows_resident;
    double ewa_inertia, init_inertia;
    int64_t reserved[4];
} rvcr_mbk_stats;
RVCR_API int rvcr_minibatch_kmeans(const float* x_host, int64_t n, int32_t d, int64_t k, int64_t batch_size, int64_t max_iter,
                                   int64_t max_no_improvement, double reassignment_ratio, int64_t n_init, uint64_t seed,
                                   int device, int verbose, float* centers_out_host, rvcr_mbk_stats* stats_out);

/* Host-only pieces of rvcr_index_train, exposed so that they can be checked without a GPU: faiss rand_perm
 * (identity, then swap(perm[i], perm[i + mt() % (n - i)]) with std::mt19937(seed)) and faiss split_clusters
 * (empty clusters re-seeded from donors picked by RandomGenerator(1234), copies perturbed by 1 +- 1/1024;
 * hassign: k floats, centroids: k x d, both updated in place; *nsplit_out = clusters re-seeded). */
RVCR_API int rvcr_debug_rand_perm(int64_t n, int64_t seed, int64_t* perm_out_host);
RVCR_API int rvcr_debug_split_clusters(int32_t d, int64_t k, int64_t n, float* hassign_host, float* centroids_host,
                                       int32_t* nsplit_out);

#ifdef __cplusplus
}
#endif
#endif /* RVCR_H_ */
