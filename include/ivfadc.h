/*
 * libivfadc_cuda -- C ABI of the B200-native IVFADC search / encoding engine.
 *
 * This header is the drop-in boundary for the hot path of JuliaNeighbors/IVFADC.jl.  The
 * reference has no FFI of its own (it is pure Julia); each entry point below replaces the body
 * of the Julia method cited next to it (paths relative to the reference tree) and is what the
 * Julia glue package binds with `ccall`.
 *
 * Conventions
 *   - every function returns an int status (IVFADC_OK == 0, negative on error) and never aborts
 *     the process; ivfadc_last_error() returns the message of the last failure on a handle;
 *   - host buffers are caller-owned: the library copies in / out and never retains a pointer;
 *   - matrices are passed exactly as Julia lays them out: a D x n column-major Matrix{T} is n
 *     contiguous vectors of D elements, so `pointer(data)` crosses the boundary zero-copy;
 *   - T is f32 or f64 (fixed at create), PQ codes are uint8 (k <= 256), vector ids are 0-based
 *     unsigned (reference src/index.jl:189) and cross the boundary as uint64;
 *   - cell (Voronoi cell / inverted list) numbers are 0-based int32 on output; the one input
 *     that carries cells (`assign`) has an explicit base so Julia's 1-based
 *     `KmeansResult.assignments` can be passed untouched;
 *   - a handle is bound to ONE CUDA device and is not thread-safe; host-buffer calls are
 *     synchronous, `_device` calls are asynchronous on the stream given; several GPUs are driven either by one
 *     handle per process joined by ivfadc_comm_init_rank, or by an ivfadc_group inside one process;
 *   - there is no CPU fallback: every entry point fails with IVFADC_ERR_CUDA when no device
 *     is present.
 */
#ifndef IVFADC_H
#define IVFADC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IVFADC_ABI_VERSION 3

/* status codes */
#define IVFADC_OK               0
#define IVFADC_ERR_BAD_ARG     -1   /* null pointer, wrong dimension, k < 1, w < 1 ...            */
#define IVFADC_ERR_CAPACITY    -2   /* id type cannot index one more vector (src/utils.jl:134)   */
#define IVFADC_ERR_CUDA        -3   /* CUDA runtime failure (message in ivfadc_last_error)       */
#define IVFADC_ERR_OOM         -4
#define IVFADC_ERR_UNSUPPORTED -5   /* metric / code width outside the hot-path scope            */
#define IVFADC_ERR_EMPTY       -6   /* pop from an empty index (src/utils.jl:44)                 */
#define IVFADC_ERR_NCCL        -7   /* NCCL could not be loaded, or a collective failed          */

/* element type T of data, centroids, codebooks and returned distances */
#define IVFADC_F32 0
#define IVFADC_F64 1

/* Distances.PreMetric of the coarse / residual quantizer (src/defaults.jl:6,8) */
#define IVFADC_SQEUCLIDEAN 0   /* Distances.SqEuclidean: the tensor-core paths                                  */
#define IVFADC_EUCLIDEAN   1   /* Distances.Euclidean, Cityblock, CosineDist: exact generic kernels (correct,    */
#define IVFADC_CITYBLOCK   2   /* untuned); Dc applies to coarse_search and the lookup tables                    */
#define IVFADC_COSINEDIST  3   /* (src/coarsequantizers.jl:34, src/index.jl:234), Dr to quantize_data            */

/* position argument of push!/pushfirst! and pop!/popfirst! (src/utils.jl:29,37,114,123) */
#define IVFADC_LAST  0
#define IVFADC_FIRST 1

/*
 * ivfadc_config.flags.  The list scan has two kernels with identical results: the
 * query-per-lane kernel (32 queries of one list per CTA; large batches, fp32, k <= 16) and the
 * general vector-per-lane kernel (any T, k <= 128, any batch).  By default the engine picks per
 * batch; the flags pin the choice (parity tests exercise both).
 */
#define IVFADC_FLAG_SCAN_LEGACY 1   /* always the vector-per-lane kernel                          */
#define IVFADC_FLAG_SCAN_QLANE  2   /* the query-per-lane kernel whenever the shape allows it     */
/*
 * The query-per-lane kernel builds its lookup tables on the tensor cores by default (GEMM form
 * |w|^2 - 2 r.w, 3xTF32, fp32 accumulate): returned distances agree with the reference's direct
 * form to ~1e-6 relative (bound 1e-5), neighbour ids except at such near-ties.  With
 * IVFADC_FLAG_LUT_EXACT it evaluates the reference's direct form as a sequential fp32 chain and
 * every distance is bit-identical to the CPU restatement, at roughly half the throughput.  The
 * vector-per-lane kernel is always exact.
 */
#define IVFADC_FLAG_LUT_EXACT   4
/*
 * The default table builder of the query-per-lane kernel runs on the 5th-generation tensor cores
 * (tcgen05.mma kind::tf32 into tensor memory, codebook streamed by TMA; fp32, dsub <= 8).  This
 * flag selects the older warp-level mma.sync builder instead (same numerics class, 3xTF32).
 */
#define IVFADC_FLAG_LUT_MMASYNC 8
/*
 * The default query-per-lane kernel keeps the lookup tables in tensor memory and looks them up with
 * tcgen05.ld at column = code byte; its persistent CTAs are warp-specialised (12 scanning warps, a
 * tensor-core producer warp, two loader warps, a finalizer; mbarriers only).  This flag selects the
 * previous generation of the same arithmetic (16 scanning warps that also issue the tensor-core work, one
 * CTA-wide barrier per table) -- kept for back-to-back measurements.
 */
#define IVFADC_FLAG_SCAN_TMEM_V1 16
/*
 * The fp32 coarse step (D <= 128) uses packed FP32 instructions on transposed centroids; this flag
 * selects the scalar FFMA kernel instead (identical results, both bit-exact against the reference form).
 */
#define IVFADC_FLAG_COARSE_SCALAR 32
/*
 * Test switch: the final per-query selection over candidate rows normally ranks up to 128 candidates
 * within the k-th-distance bound in registers and falls back to k filtered sweeps over the rows beyond
 * that (heavy distance ties); with this flag the fallback takes over at 4, so ordinary data exercises it.
 */
#define IVFADC_FLAG_TEST_MERGE_SWEEP 64
/*
 * The fp32 coarse step (kc >= 256, D % 16 == 0, D <= 128, w <= 32) runs on the tensor cores: TF32
 * scores prune the centroids to a provable superset of the exact top-w, which is then re-ranked with
 * the reference's direct form (results bit-identical to the FFMA kernels).  This flag keeps the
 * packed-FP32 kernel as the coarse step.
 */
#define IVFADC_FLAG_COARSE_FFMA 128
/* Test switch: the tensor-core coarse kernel flags EVERY query for the FFMA redo pass. */
#define IVFADC_FLAG_TEST_COARSE_REDO 256

typedef struct ivfadc_index ivfadc_index;   /* opaque, owns all device memory */

typedef struct ivfadc_config {
    int32_t dim;            /* D: rows of the data matrix                                         */
    int32_t kc;             /* number of coarse centroids / inverted lists                        */
    int32_t m;              /* number of PQ codebooks; sub-dimension = floor(D / m)               */
    int32_t ksub;           /* codewords per codebook (<= 256, codes are uint8)                   */
    int32_t dtype;          /* IVFADC_F32 | IVFADC_F64                                            */
    int32_t id_bytes;       /* sizeof(I) of the Julia index type: 1, 2, 4 or 8                    */
    int32_t metric_coarse;  /* Dc: IVFADC_SQEUCLIDEAN | _EUCLIDEAN | _CITYBLOCK | _COSINEDIST      */
    int32_t metric_resid;   /* Dr: same codes                                                     */
    int32_t device;         /* CUDA device ordinal                                                */
    int32_t shard_rank;     /* this handle keeps only cells with cell % shard_world == shard_rank */
    int32_t shard_world;    /* 1 = unsharded                                                      */
    int32_t flags;          /* 0 = defaults; IVFADC_FLAG_* tuning / test switches                 */
} ivfadc_config;

typedef struct ivfadc_stats {
    uint64_t searches;          /* number of search calls since reset                             */
    uint64_t queries;           /* queries processed                                              */
    uint64_t scanned_vectors;   /* sum over (query, probed list) of list length                   */
    uint64_t scan_code_bytes;   /* scanned_vectors * m : the algorithmic bytes of the list scan   */
    uint64_t gpu_launches;      /* kernels launched by this library since reset                   */
    double   coarse_ms;         /* accumulated device time, CUDA events on the launch stream      */
    double   plan_ms;
    double   scan_ms;
    double   merge_ms;
    double   encode_ms;
    uint64_t scan_launches;
    uint64_t last_scan_kernel;  /* kernel of the last search: 1 vector-per-lane, 2 scanq, 4 scanu (v1), 5 scanw */
    double   comm_ms;           /* sharded search: device time of the two grouped all-gathers (eager steps)  */
    uint64_t last_coarse_redo;  /* queries of the last tensor-core coarse step that the exact FFMA pass had to redo */
    uint64_t reserved[1];
} ivfadc_stats;

int ivfadc_abi_version(void);

/* Number of CUDA devices visible, or a negative status. */
int ivfadc_device_count(void);

/*
 * Build an empty index around trained quantizers.
 *   centroids        : T[kc][D]            (Julia: cq.vectors, D x kc)
 *   codebook_vectors : T[m][ksub][dsub]    (Julia: codebooks[i].vectors, dsub x ksub, i = 1..m)
 *   codebook_codes   : uint8[m][ksub]      (Julia: codebooks[i].codes)
 * Replaces the struct construction at src/index.jl:155-164 and src/persistency.jl:95-117.
 */
int ivfadc_create(ivfadc_index** out, const ivfadc_config* cfg, const void* centroids,
                  const void* codebook_vectors, const uint8_t* codebook_codes);

int ivfadc_destroy(ivfadc_index* h);

const char* ivfadc_last_error(const ivfadc_index* h);

/*
 * Add n vectors X[n][D].
 *   position = IVFADC_LAST : vector j gets id N + j                  (n x push!,      src/utils.jl:114)
 *   position = IVFADC_FIRST: every stored id += n, vector j gets id n-1-j
 *                            (what n successive pushfirst! calls do, src/utils.jl:123,140-141)
 *   assign == NULL : cell = coarse_search(x, 1)                      (_encode_point, src/utils.jl:148-161)
 *   assign != NULL : cell = assign[j] - assign_base                  (index build from k-means
 *                            assignments, _build_residuals + _build_inverted_index,
 *                            src/index.jl:168-194)
 * Codes are the PQ encoding of x - centroid[cell]; entries are appended at the list tail in
 * batch order (src/utils.jl:142-143).  cells_out (optional, int32[n]) receives the cells.
 * Fails with IVFADC_ERR_CAPACITY, leaving the index untouched, when N + n > 2^(8*id_bytes).
 */
int ivfadc_add(ivfadc_index* h, const void* X, int64_t n, int32_t position,
               const int64_t* assign, int32_t assign_base, int32_t* cells_out);
/*
 * Same, with the batch (and the optional assignments / cells) already in device memory of the handle's device:
 * the index build of the 10 M / 100 M-vector configurations never crosses PCIe (_build_inverted_index,
 * src/index.jl:178-194).  Synchronous like ivfadc_add; out-of-range assignments fail with IVFADC_ERR_BAD_ARG.
 */
int ivfadc_add_device(ivfadc_index* h, const void* dX, int64_t n, int32_t position,
                      const int64_t* d_assign, int32_t assign_base, int32_t* d_cells_out);

/*
 * Capacity hint before a bulk build (Julia's sizehint!; the reference grows its lists by push!, src/utils.jl:142-143):
 * room for n_total vectors over all shards -- an even share per list plus 12% -- or, with sizes != NULL (int64[kc],
 * e.g. the k-means cluster counts), exactly sizes[c] entries in list c.  One arena allocation instead of geometric
 * regrowth (each regrowth is a device allocation + a copy of everything stored so far).  Never shrinks.
 */
int ivfadc_reserve(ivfadc_index* h, int64_t n_total, const int64_t* sizes);

/*
 * Encode without mutating: cells int32[n] (0-based), codes uint8[n][m].  If assign != NULL
 * the cells are taken from it as in ivfadc_add.  Parity hook for _encode_point
 * (src/utils.jl:148-161) and QuantizedArrays.quantize_data (src/index.jl:187).
 */
int ivfadc_encode(ivfadc_index* h, const void* X, int64_t n, const int64_t* assign,
                  int32_t assign_base, int32_t* cells_out, uint8_t* codes_out);

/*
 * coarse_search for a batch (src/coarsequantizers.jl:33-37): the w nearest centroids of every
 * query in ascending (distance, cell) order.  cells int32[nq][w] 0-based, dc T[nq][w].
 * w is clamped to kc by the caller (src/index.jl:216).
 */
int ivfadc_coarse_search(ivfadc_index* h, const void* Q, int64_t nq, int32_t w,
                         int32_t* cells_out, void* dc_out);

/*
 * Batched knn_search (src/index.jl:204-273).  Q[nq][D]; k >= 1; w >= 1 (clamped to kc).
 * Outputs, row i for query i, first counts[i] (<= k) entries valid, ascending in
 * (distance, probe rank, position in list) (src/index.jl:247-257):
 *   ids   uint64[nq][k]  0-based vector ids
 *   dists T[nq][k]       dc + sum of lookup-table entries (src/index.jl:242-246)
 *   counts int32[nq]
 * Unused slots are filled with id = UINT64_MAX, dist = +inf.
 */
int ivfadc_search(ivfadc_index* h, const void* Q, int64_t nq, int32_t k, int32_t w,
                  uint64_t* ids_out, void* dists_out, int32_t* counts_out);

/*
 * Same, with every pointer a DEVICE pointer on the handle's device and the work enqueued on
 * `stream` (a cudaStream_t; NULL = the legacy default stream).  Returns without synchronising.
 */
int ivfadc_search_device(ivfadc_index* h, const void* dQ, int64_t nq, int32_t k, int32_t w,
                         uint64_t* d_ids, void* d_dists, int32_t* d_counts, void* stream);

/*
 * Sharded search, step 1 (device pointers): scan only the probed cells this shard owns and
 * return, per query, its k best local candidates with their merge keys
 *   key = (probe rank << 32) | position in list
 * so that ranks can be merged in the reference's order.  d_keys uint64[nq][k].
 */
int ivfadc_search_local_device(ivfadc_index* h, const void* dQ, int64_t nq, int32_t k, int32_t w,
                               uint64_t* d_ids, void* d_dists, uint64_t* d_keys,
                               int32_t* d_counts, void* stream);

/*
 * Sharded search with the coarse step sharded BY QUERY (device pointers): rank r runs
 * ivfadc_coarse_search_device on its slice of the batch, the host all-gathers the probe lists
 * (cells int32[nq][w], dc T[nq][w], both in query order) and every rank calls
 * ivfadc_search_probes_local_device -- step 1 with the probes supplied instead of recomputed, so
 * the replicated coarse_search (src/coarsequantizers.jl:33-37) no longer limits scaling.
 * w must already be clamped to kc (src/index.jl:216).
 */
int ivfadc_coarse_search_device(ivfadc_index* h, const void* dQ, int64_t nq, int32_t w, int32_t* d_cells,
                                void* d_dc, void* stream);
int ivfadc_search_probes_local_device(ivfadc_index* h, const void* dQ, int64_t nq, int32_t k, int32_t w,
                                      const int32_t* d_cells, const void* d_dc, uint64_t* d_ids, void* d_dists,
                                      uint64_t* d_keys, int32_t* d_counts, void* stream);

/*
 * Sharded search, step 2 (device pointers): merge `parts` candidate sets laid out
 * [parts][nq][k] (as gathered from the ranks) into the final [nq][k] by (dist, key).
 */
int ivfadc_merge_device(ivfadc_index* h, int32_t parts, int64_t nq, int32_t k,
                        const uint64_t* d_ids_in, const void* d_dists_in,
                        const uint64_t* d_keys_in, uint64_t* d_ids, void* d_dists,
                        int32_t* d_counts, void* stream);

/*
 * ---- cell-sharded search inside the library (NCCL over NVLink / NVSwitch) -----------------------------------
 * The one batched knn_search call of the reference (src/index.jl:261-273) fans out over the GPUs by itself:
 * every rank (handle with shard_rank / shard_world) runs coarse_search on its slice of the queries, ONE grouped
 * in-place all-gather distributes the probe lists (and the query slices, when they came from the host), every rank
 * scans the probed cells it owns, ONE grouped in-place all-gather brings the per-rank candidates (id, distance,
 * probe rank << 32 | position) together, and a merge kernel keeps the k smallest in the reference's order
 * (src/index.jl:247-257): bit-identical to the unsharded result.  The whole step is enqueued on one stream without
 * host synchronisation and replayed from a CUDA graph from the third call with the same shape on.
 *
 * One process per GPU: rank 0 calls ivfadc_nccl_unique_id, the IVFADC_NCCL_ID_BYTES travel over the caller's own
 * channel (torch.distributed / MPI / a file), every rank calls ivfadc_comm_init_rank on its handle.
 */
#define IVFADC_NCCL_ID_BYTES 128
int ivfadc_nccl_unique_id(void* id_out);
int ivfadc_comm_init_rank(ivfadc_index* h, const void* id, int32_t world, int32_t rank);
int ivfadc_comm_destroy(ivfadc_index* h);
/* CUDA-graph replay of the sharded step on / off (default on; per-kernel stats timing needs eager steps). */
int ivfadc_set_graph_replay(ivfadc_index* h, int32_t enable);
/*
 * Collective call: every rank passes the same batch.  Host variant: Q[nq][D] replicated on every rank, each rank
 * uploads only ITS slice (the all-gather completes the batch over NVLink) and receives the full result.
 * Device variant: dQ is the full batch on every rank's device; asynchronous on `stream`.
 */
int ivfadc_search_sharded(ivfadc_index* h, const void* Q, int64_t nq, int32_t k, int32_t w, uint64_t* ids_out,
                          void* dists_out, int32_t* counts_out);
int ivfadc_search_sharded_device(ivfadc_index* h, const void* dQ, int64_t nq, int32_t k, int32_t w, uint64_t* d_ids,
                                 void* d_dists, int32_t* d_counts, void* stream);
/* Bytes one ivfadc_search_sharded step moves on this rank: host->device, device->host, received over NVLink. */
int ivfadc_sharded_step_bytes(ivfadc_index* h, int64_t nq, int32_t k, int32_t w, int64_t* h2d_out, int64_t* d2h_out,
                              int64_t* nvlink_out);
/*
 * Which shard owns a cell: owners int32[kc] with values in [0, shard_world) (e.g. greedy bin-packing on the
 * list lengths known from the k-means assignments); default cell % shard_world.  Must be called on every shard
 * with the same map, before any vector is added.
 */
int ivfadc_set_cell_owners(ivfadc_index* h, const int32_t* owners);
/*
 * The *_device entry points are asynchronous and do not read the error flag of the tensor-core pipelines
 * (mbarrier time-outs); this call synchronises `stream` and reports (and clears) it.
 */
int ivfadc_check_async(ivfadc_index* h, void* stream);

/*
 * One process, several GPUs (what the Julia glue binds when IVFADC_DEVICES names more than one device): n handles
 * with shard_rank = i, a communicator from ncclCommInitAll, and the calls of the single-GPU API fanned out.
 * device_ids == NULL means devices 0..n_devices-1.  ivfadc_group_handle gives access to a shard for everything else
 * (list export / import, statistics).
 */
typedef struct ivfadc_group ivfadc_group;
int ivfadc_group_create(ivfadc_group** out, const ivfadc_config* cfg, int32_t n_devices, const int32_t* device_ids,
                        const void* centroids, const void* codebook_vectors, const uint8_t* codebook_codes);
int ivfadc_group_destroy(ivfadc_group* g);
int ivfadc_group_size(const ivfadc_group* g);
int ivfadc_group_handle(ivfadc_group* g, int32_t i, ivfadc_index** out);
const char* ivfadc_group_last_error(const ivfadc_group* g);
int ivfadc_group_set_cell_owners(ivfadc_group* g, const int32_t* owners);
int ivfadc_group_add(ivfadc_group* g, const void* X, int64_t n, int32_t position, const int64_t* assign,
                     int32_t assign_base, int32_t* cells_out);
int ivfadc_group_search(ivfadc_group* g, const void* Q, int64_t nq, int32_t k, int32_t w, uint64_t* ids_out,
                        void* dists_out, int32_t* counts_out);
int ivfadc_group_delete(ivfadc_group* g, const uint64_t* ids, int64_t n);
int ivfadc_group_pop(ivfadc_group* g, int32_t position, void* vec_out);
int ivfadc_group_length(const ivfadc_group* g, int64_t* n_out);

/*
 * delete_from_index! (src/utils.jl:90-105): ids are the STORED 0-based ids (the Julia glue
 * performs the `I.(points .- 1)` conversion and its InexactError).  Duplicates and unknown
 * ids are ignored; survivors keep their in-list order and are renumbered
 * new_id = old_id - |{deleted ids < old_id}|.  When sharded, `ids` must be the full global list.
 */
int ivfadc_delete(ivfadc_index* h, const uint64_t* ids, int64_t n);

/*
 * pop! / popfirst! (src/utils.jl:29-81): remove id N-1 (LAST) or id 0 (FIRST; all other ids
 * -= 1) and return centroid + decoded residual in vec_out T[D].
 * found_out (optional) is 1 when this shard owned the vector (always 1 unsharded).
 */
int ivfadc_pop(ivfadc_index* h, int32_t position, void* vec_out, int32_t* found_out);

/* length(ivfadc) (src/index.jl:56): total over ALL shards as tracked by this handle. */
int ivfadc_length(const ivfadc_index* h, int64_t* n_out);

/* Per-list lengths of the cells stored in this handle, int64[kc] (0 for cells of other shards). */
int ivfadc_list_sizes(ivfadc_index* h, int64_t* sizes_out);

/*
 * Copy one inverted list out / in (persistency, src/persistency.jl:68-78,119-131):
 * ids uint64[len], codes uint8[len][m] in list order.  Import replaces the list and adjusts
 * the length of the index.
 */
int ivfadc_export_list(ivfadc_index* h, int32_t cell, uint64_t* ids_out, uint8_t* codes_out);
int ivfadc_import_list(ivfadc_index* h, int32_t cell, const uint64_t* ids, const uint8_t* codes,
                       int64_t len);

/*
 * All lists at once (bulk persistency; the reference writes / reads element by element,
 * src/persistency.jl:68-78,119-131): sizes come from ivfadc_list_sizes, the entries of all lists follow each other
 * in ascending cell order -- ids uint64[sum sizes], codes uint8[sum sizes][m] -- gathered on the device and copied
 * in one transfer per 16 M entries.  ivfadc_import_all REPLACES every list of the handle and adjusts the length.
 */
int ivfadc_export_all(ivfadc_index* h, uint64_t* ids_out, uint8_t* codes_out);
int ivfadc_import_all(ivfadc_index* h, const int64_t* sizes, const uint64_t* ids, const uint8_t* codes);

/* Copy the quantizers back out (same layouts as ivfadc_create). */
int ivfadc_export_quantizers(ivfadc_index* h, void* centroids_out, void* codebook_vectors_out,
                             uint8_t* codebook_codes_out);

/* Override the global vector count (used by load and by sharded handles). */
int ivfadc_set_length(ivfadc_index* h, int64_t n_total);

/*
 * Diagnostics of the tensor-core table builder (tests only).  out == NULL arms a dump; after the
 * next search, out receives float[m][256][32] (the lookup tables of work item 0: subspace, code
 * value, query slot; WITHOUT the per-query constant dc + |r|^2), then int32[32] (query * w + probe
 * rank of every slot, -1 = empty) and int32 (the cell), i.e. (m * 256 * 32 + 33) * 4 bytes.
 */
int ivfadc_debug_tables(ivfadc_index* h, void* out);

/*
 * Per-kernel CUDA-event timing of the search path (ivfadc_stats.*_ms) on / off (default on).
 * Switch it off before capturing *_device calls into a CUDA graph: event queries are not
 * capturable.  Counters (queries, scanned vectors, launches) keep running.
 */
int ivfadc_set_stats_timing(ivfadc_index* h, int32_t enable);

int ivfadc_get_stats(ivfadc_index* h, ivfadc_stats* out);
int ivfadc_reset_stats(ivfadc_index* h);

/*
 * ---- quantizer training on the device (not parity-graded) ------------------------------
 * The constructor's cost at 10^6+ vectors is Clustering.kmeans (src/index.jl:129-134) and build_quantizer
 * (src/index.jl:142-147).  Lloyd on the GPU with ONE handle for all iterations: assignment = the engine's own coarse
 * kernel (ivfadc_coarse_search_device, w = 1), centre update = accumulate (fp64 atomics) + finish, then
 * ivfadc_set_centroids_device installs the new centres (index must be empty).  k-means++ seeding (D^2 sampling,
 * Philox4x32-10 keyed by `seed`: deterministic; `trials` > 1 = the greedy variant of scikit-learn: that many
 * candidates per round, the one that lowers the potential most wins) picks k rows of a device-resident sample.
 * dtype: IVFADC_F32 / IVFADC_F64 of the data; d_scratch: double[ns]; d_sums: double[kc][D] and d_counts: uint64[kc],
 * zeroed by the caller before the first accumulate of an iteration; d_empty_out (optional): int32[kc], 1 = no point.
 */
int ivfadc_set_centroids_device(ivfadc_index* h, const void* d_centroids);
int ivfadc_kmeanspp_device(const void* dS, int64_t ns, int32_t D, int32_t k, int32_t dtype, uint64_t seed, int32_t trials,
                           void* d_scratch, void* d_centres_out, int64_t* d_picked_out, void* stream);
int ivfadc_kmeans_accumulate_device(const void* dX, int64_t n, int32_t D, int32_t dtype, const int32_t* d_cells,
                                    double* d_sums, uint64_t* d_counts, void* stream);
int ivfadc_kmeans_finish_device(const double* d_sums, const uint64_t* d_counts, int32_t kc, int32_t D, int32_t dtype,
                                void* d_centroids_inout, int32_t* d_empty_out, void* stream);

/*
 * ---- synthetic inputs (harness utility) -------------------------------------------------
 * Counter-based generator, Philox4x32-10 keyed by `seed`, counter = (vector lo, vector hi, dim, stream): every
 * value depends only on (seed, vector index, dim), so a slice [first, first + n) of a data set of any size is
 * produced directly in HBM.  oracle/ivfadc_oracle.c (oracle_synth_*) computes the same values bit for bit.
 * fp32 only; the reference's README / tests draw `rand(Float32, D, N)` (README.md:32, test/index.jl:7).
 *   uniform: X[n][D] in [0, 1) (24 bits)
 *   blobs  : X = centres[blob(vector)] + t * scale, t = centred sum of four 22-bit uniforms (standard deviation
 *            2^22 / sqrt(3)): pass scale = sigma * sqrt(3) / 2^22; d_blobs_out (optional) receives the blob ids
 */
int ivfadc_synth_uniform_device(void* dX, int64_t first, int64_t n, int32_t D, uint64_t seed, void* stream);
int ivfadc_synth_blobs_device(void* dX, int32_t* d_blobs_out, int64_t first, int64_t n, int32_t D, int32_t n_blobs,
                              uint64_t seed, float scale, const void* d_centres, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IVFADC_H */
