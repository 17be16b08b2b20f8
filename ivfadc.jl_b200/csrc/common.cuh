// Internal declarations shared by the translation units of libivfadc_cuda (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <unordered_map>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "ivfadc.h"

namespace ivf {

// ---------------------------------------------------------------------------------------------
// Arithmetic contract (bit-for-bit the same as oracle/ivfadc_oracle_impl.h A1-A3): every
// distance is a strictly sequential fma chain, the compiler is never allowed to re-associate or
// contract on its own (explicit intrinsics only).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_rn(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }

__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

// colwise(Dc, a, b) for the metrics beyond SqEuclidean (oracle dist_colwise, same chains): 1 Euclidean = sqrt of
// the SqEuclidean chain, 2 Cityblock = sequential sum of |a - b|, 3 CosineDist = max(1 - dot / (|a| |b|), 0) from
// three sequential fma chains.  `a` = the centroid / codeword, `b` = the query / residual.
template <typename T>
__device__ __forceinline__ T metric_dist(int metric, const T* __restrict__ a, const T* __restrict__ b, int n) {
    if (metric == 2) {
        T s = (T)0;
        for (int i = 0; i < n; ++i) {
            const T d = sub_rn(a[i], b[i]);
            s = add_rn(s, d < (T)0 ? -d : d);
        }
        return s;
    }
    if (metric == 3) {
        T ab = (T)0, a2 = (T)0, b2 = (T)0;
        for (int i = 0; i < n; ++i) {
            ab = fma_rn(a[i], b[i], ab);
            a2 = fma_rn(a[i], a[i], a2);
            b2 = fma_rn(b[i], b[i], b2);
        }
        const T v = sub_rn((T)1, div_rn(ab, mul_rn(sqrt_rn(a2), sqrt_rn(b2))));
        return v > (T)0 ? v : (T)0;
    }
    T s = (T)0;
    for (int i = 0; i < n; ++i) {
        const T d = sub_rn(a[i], b[i]);
        s = fma_rn(d, d, s);
    }
    return metric == 1 ? sqrt_rn(s) : s;
}

template <typename T> struct Limits;
template <> struct Limits<float> {
    typedef uint32_t bits_t;
    __host__ __device__ static float inf() {
#ifdef __CUDA_ARCH__
        return __int_as_float(0x7f800000);
#else
        return __builtin_inff();
#endif
    }
};
template <> struct Limits<double> {
    typedef unsigned long long bits_t;
    __host__ __device__ static double inf() {
#ifdef __CUDA_ARCH__
        return __longlong_as_double(0x7ff0000000000000LL);
#else
        return __builtin_inf();
#endif
    }
};

__device__ __forceinline__ uint32_t to_bits(float x) { return __float_as_uint(x); }
__device__ __forceinline__ unsigned long long to_bits(double x) {
    return (unsigned long long)__double_as_longlong(x);
}
__device__ __forceinline__ float from_bits(uint32_t b) { return __uint_as_float(b); }
__device__ __forceinline__ double from_bits(unsigned long long b) {
    return __longlong_as_double((long long)b);
}
// Smallest value strictly greater than x, for finite x >= 0 (+inf stays +inf).  Turns the
// inclusive bound "keep d <= x" into the exclusive one "keep d < next_up(x)".
template <typename T> __device__ __forceinline__ T next_up_nonneg(T x) {
    return x == Limits<T>::inf() ? x : from_bits((typename Limits<T>::bits_t)(to_bits(x) + 1));
}

constexpr uint32_t kNoPos = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// Grow-only device buffer.
// ---------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename U> U* as() const { return reinterpret_cast<U*>(p); }
};

struct EventTimer {
    cudaEvent_t a = nullptr, b = nullptr;
};

}  // namespace ivf

// The opaque handle of include/ivfadc.h.
struct ivfadc_index {
    ivfadc_config cfg{};
    int dsub = 0;
    size_t tsize = 4;       // sizeof(T)
    int id_dev_bytes = 4;   // width of ids in device memory: 4 (id_bytes <= 4) or 8
    cudaStream_t stream = nullptr;

    // quantizers (device)
    void* d_centroids = nullptr;    // T[kc][D]
    void* d_centroids_t = nullptr;  // fp32 only: centroids transposed [D][kc_pad] (packed-FP32 coarse kernel)
    int kc_pad = 0;
    void* d_tcC = nullptr;          // fp32 only: centroids as tcgen05 B-operand blocks [kc_pad256 / 256][D / 8][2048] (coarse_tc.cuh)
    void* d_ccn = nullptr;          // fp32 only: squared centroid norms [kc_pad256] (+inf padding), then max |c|^2
    int kc_pad256 = 0;
    mutable ivf::DevBuf ws_coarse_redo;  // uint8[nq]: queries the tensor-core coarse kernel hands to the FFMA kernel
    void* d_cb = nullptr;           // T[m][ksub][dsub]
    uint8_t* d_cb_codes = nullptr;  // uint8[m][ksub]
    void* d_cb_norms = nullptr;     // T[m][ksub]  squared norms of the codewords (oracle A2)
    int cb_identity = 0;            // codes[i][c] == c for all i, c
    void* d_afrag = nullptr;        // codebook as mma A fragments (scanq FAST table builder)
    void* d_wnfrag = nullptr;
    int frag_ntiles = 0, frag_ksteps = 0;
    void* d_tcU = nullptr;          // codebook as per-subspace B-operand blocks, rows = code values (scanu)
    void* d_tcH = nullptr;          // the same as fp16 two-piece operand blocks (scanw): [tables][2][256 rows][16 k]
    int tch_ew = 0, tch_en = 0;     // power-of-two scales of -2w and |w|^2 in d_tcH (fp16 range), fixed at create
    int* h_err = nullptr;           // pinned host copy of d_err (read back with the results)
    int* d_err = nullptr;           // device error flag of the tcgen05 pipeline (mbarrier timeout)
    void* d_dbg_lut = nullptr;      // optional table dump of work item 0 (tests), float[m][256][32]

    // inverted lists: device-resident CSR with slack.  List c occupies entries
    // [off[c], off[c] + len[c]) of the arenas, capacity cap[c]; off[c] is a multiple of 16 so
    // that every list's code block starts 16-byte aligned for any m.
    uint8_t* d_codes = nullptr;  // uint8[arena_cap][m]
    void* d_ids = nullptr;       // uint32|uint64[arena_cap]
    int64_t arena_cap = 0;       // in vectors
    int64_t arena_used = 0;
    std::vector<int64_t> h_off, h_len, h_cap;  // host mirrors, size kc
    int64_t* d_off = nullptr;                  // int64[kc]
    int64_t* d_len = nullptr;                  // int64[kc]
    int64_t n_total = 0;                       // length(ivfadc) over all shards
    int64_t n_local = 0;                       // vectors stored in this handle
    // which shard owns a cell: owner[c] (ivfadc_set_cell_owners: balanced by list length), or c % shard_world
    std::vector<int32_t> h_owner;              // empty = modulo
    int32_t* d_owner = nullptr;                // int32[kc] or null
    bool owns(int c) const {
        if (cfg.shard_world <= 1) return true;
        return (h_owner.empty() ? c % cfg.shard_world : h_owner[c]) == cfg.shard_rank;
    }

    // search / mutation workspaces (grow-only)
    ivf::DevBuf ws_q, ws_cells, ws_dc, ws_bucket, ws_sorted, ws_pair_d, ws_pair_pos, ws_pair_cnt,
        ws_thr, ws_out_ids, ws_out_d, ws_out_cnt, ws_out_keys, ws_misc, ws_x, ws_codes, ws_assign,
        ws_sort_tmp, ws_sort_keys, ws_sort_vals, ws_del, ws_items;

    // Per-handle (hence per-device) launch state: the dynamic shared-memory limit configured for each kernel
    // (function attributes are per device) and the SM count of the handle's device.
    mutable std::unordered_map<const void*, size_t> func_smem;
    int num_sms = 0;

    cudaEvent_t ev[10] = {};
    bool stats_timing = true;
    ivfadc_stats stats{};
    std::string err;
    mutable int64_t scanw_ws_n = -1;   // scan.cu: list population the CTA shape of the scan kernel was chosen for
    mutable int scanw_ws = 0;          // 12 or 16 scanning warps (0 = not chosen yet)
    mutable int64_t last_redo_nq = 0;  // coarse.cu: queries of the last tensor-core coarse step (redo flags in ws_coarse_redo)
    // Float64 index, default flags, large batch: the search runs on a Float32 twin (same lists, quantizers rounded
    // to fp32) -- the tensor-core path within the north_star tolerance instead of the exact fp64 chain (api.cu)
    ivfadc_index* twin = nullptr;
    bool twin_failed = false, twin_used = false;
    void* extra = nullptr;      // api.cu: event ring, scanned-vector counter
    void* shard_ctx = nullptr;  // shard.cu: NCCL communicator, gathered buffers, CUDA graphs
};

namespace ivf {

// NVTX range around a C-ABI call (header-only NVTX 3: a no-op unless a profiler injects itself), so that an
// nsys / ncu timeline of a Julia or Python host shows which call owns which kernels.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
#define IVF_NVTX() ::ivf::NvtxRange nvtx_range_(__func__)

constexpr size_t kSmemMax = 227 * 1024;   // dynamic shared memory a CTA can opt in to on sm_100a

// cudaFuncAttributeMaxDynamicSharedMemorySize of `func` on the handle's device, raised once per (handle, kernel).
inline cudaError_t ensure_smem(const ivfadc_index* h, const void* func, size_t smem) {
    if (smem <= 48 * 1024) return cudaSuccess;
    size_t& cur = h->func_smem[func];
    if (smem <= cur) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) cur = smem;
    return e;
}

// ---- coarse.cu --------------------------------------------------------------------------------
// K1: w nearest centroids of every query, ascending (distance, cell); direct-form distances.
cudaError_t launch_coarse(const ivfadc_index* h, const void* dQ, int64_t nq, int w, int32_t* d_cells,
                          void* d_dc, cudaStream_t s, int* launches);
int coarse_max_w();
cudaError_t coarse_prepare(ivfadc_index* h, cudaStream_t s, int* launches);

// ---- scan.cu ----------------------------------------------------------------------------------
struct ScanPlanSizes {
    size_t bucket_bytes, sorted_bytes, pair_d_bytes, pair_pos_bytes, pair_cnt_bytes, thr_bytes, items_bytes;
};
int scan_max_k();
cudaError_t scanq_prepare(ivfadc_index* h, cudaStream_t s, int* launches);
bool scan_supported(const ivfadc_index* h, std::string* why);
bool scan_takes_tensor_path(const ivfadc_index* h, int64_t npairs, int k);   // the cost model picks the tensor-memory kernel
bool encode_supported(const ivfadc_index* h);   // encode.cu: one codeword block + one residual slice fit a CTA
ScanPlanSizes scan_plan_sizes(const ivfadc_index* h, int64_t nq, int w, int k);
// K2+K3: plan, fused LUT build + list scan + per-pair top-k, then per-query merge.
// Outputs (device): ids uint64[nq][k], dists T[nq][k], keys uint64[nq][k] (optional), counts.
cudaError_t launch_search(ivfadc_index* h, const void* dQ, int64_t nq, int k, int w,
                          const int32_t* d_cells, const void* d_dc, uint64_t* d_ids, void* d_dists,
                          uint64_t* d_keys, int32_t* d_counts, uint64_t* d_scanned, cudaStream_t s,
                          int* launches);
cudaError_t launch_merge_parts(const ivfadc_index* h, int parts, int64_t nq, int k,
                               const uint64_t* d_ids_in, const void* d_dists_in,
                               const uint64_t* d_keys_in, uint64_t* d_ids, void* d_dists,
                               int32_t* d_counts, cudaStream_t s, int* launches);

// ---- api.cu (shared with shard.cu) ---------------------------------------------------------------
int api_fail(ivfadc_index* h, int code, const char* msg, cudaError_t e = cudaSuccess);
// Batched search with every pointer on the device, asynchronous on `s`; ext_cells / ext_dc: probe lists supplied by
// the caller (coarse step done elsewhere) or null.
int api_search_core(ivfadc_index* h, const void* dQ, int64_t nq, int k, int w, uint64_t* d_ids, void* d_dists,
                    uint64_t* d_keys, int32_t* d_counts, cudaStream_t s, const int32_t* ext_cells = nullptr,
                    const void* ext_dc = nullptr);
// Reads (and clears) the device error flag of the tensor-core pipelines; IVFADC_OK or IVFADC_ERR_CUDA with a message.
int api_check_pipeline_flag(ivfadc_index* h, int flag);

// ---- shard.cu ------------------------------------------------------------------------------------
void shard_destroy_ctx(ivfadc_index* h);
void shard_flush_timing(ivfadc_index* h);

// ---- encode.cu --------------------------------------------------------------------------------
cudaError_t launch_codebook_norms(const ivfadc_index* h, cudaStream_t s, int* launches);
// K4: residual w.r.t. d_cells + per-subspace argmin (GEMM form).  codes uint8[n][m].
cudaError_t launch_encode(const ivfadc_index* h, const void* dX, int64_t n, const int32_t* d_cells,
                          uint8_t* d_codes_out, cudaStream_t s, int* launches);
cudaError_t launch_assign_check(const int64_t* d_assign, int64_t n, int base, int kc, int* d_bad, cudaStream_t s,
                                int* launches);
cudaError_t launch_assign_to_cells(const int64_t* d_assign, int64_t n, int base, int32_t* d_cells,
                                   cudaStream_t s, int* launches);

// ---- lists.cu ---------------------------------------------------------------------------------
// K5: all list-mutation kernels.
cudaError_t lists_init(ivfadc_index* h);
void lists_free(ivfadc_index* h);
// Append n encoded vectors (device cells/codes) with ids first_id + step*j (step = +1 or -1),
// keeping only cells owned by this shard; in batch order at the list tails.
cudaError_t lists_append(ivfadc_index* h, const int32_t* d_cells, const uint8_t* d_codes, int64_t n,
                         uint64_t first_id, int step, int* launches);
cudaError_t lists_shift_ids(ivfadc_index* h, int64_t by, int* launches);
// Delete the (sorted, unique, device) ids; returns number of vectors removed from this shard.
cudaError_t lists_delete(ivfadc_index* h, const uint64_t* d_sorted_ids, int64_t n, int64_t* removed,
                         int* launches);
// Locate one id: cell / position or -1.
cudaError_t lists_find(ivfadc_index* h, uint64_t id, int32_t* cell, int64_t* pos, int* launches);
cudaError_t lists_decode(ivfadc_index* h, int32_t cell, int64_t pos, void* d_vec_out, int* launches);
cudaError_t lists_export(ivfadc_index* h, int32_t cell, uint64_t* ids_out, uint8_t* codes_out);
cudaError_t lists_reserve(ivfadc_index* h, const std::vector<int64_t>& need, int* launches);
cudaError_t lists_export_all(ivfadc_index* h, uint64_t* ids_out, uint8_t* codes_out, int* launches);
cudaError_t lists_import_all(ivfadc_index* h, const int64_t* sizes, const uint64_t* ids, const uint8_t* codes,
                             int* launches);
cudaError_t lists_import(ivfadc_index* h, int32_t cell, const uint64_t* ids, const uint8_t* codes,
                         int64_t len, int* launches);
cudaError_t lists_sync_meta_to_device(ivfadc_index* h);

}  // namespace ivf
