// The C ABI of include/ivfadc.h: argument checking, host<->device staging, chunking, statistics.
// All numerics live in coarse.cu / scan*.cu / encode.cu / lists.cu.
#include <algorithm>
#include <cmath>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>

#include "common.cuh"

using namespace ivf;

namespace {

constexpr int kEventSets = 32;   // ring of per-search event sets (coarse a/b, scan a/b, end)
constexpr int kEventsPerSet = 6;

struct Timing {
    cudaEvent_t ev[kEventSets][kEventsPerSet];
    int pending[kEventSets];  // 0 = free, 1 = recorded
    int next = 0;
    bool ok = false;
};

// handle-private extras kept out of common.cuh
struct Extra {
    Timing tm;
    uint64_t* d_scanned = nullptr;  // device counter of scanned vectors
    uint64_t scanned_flushed = 0;
};

Extra* extra(ivfadc_index* h) { return static_cast<Extra*>(h->extra); }

int fail(ivfadc_index* h, int code, const char* msg, cudaError_t e = cudaSuccess) {
    if (h) {
        h->err = msg;
        if (e != cudaSuccess) {
            h->err += ": ";
            h->err += cudaGetErrorString(e);
        }
    }
    return code;
}

#define CUDA_OR_FAIL(h, call, what)                                   \
    do {                                                              \
        cudaError_t _e = (call);                                      \
        if (_e != cudaSuccess) {                                      \
            cudaGetLastError();                                       \
            return fail(h, _e == cudaErrorMemoryAllocation ? IVFADC_ERR_OOM : IVFADC_ERR_CUDA, what, _e); \
        }                                                             \
    } while (0)

void flush_set(ivfadc_index* h, int i) {
    Timing& tm = extra(h)->tm;
    if (!tm.pending[i]) return;
    cudaEventSynchronize(tm.ev[i][5]);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, tm.ev[i][0], tm.ev[i][1]) == cudaSuccess) h->stats.coarse_ms += ms;
    if (cudaEventElapsedTime(&ms, tm.ev[i][1], tm.ev[i][2]) == cudaSuccess) h->stats.plan_ms += ms;
    if (cudaEventElapsedTime(&ms, tm.ev[i][2], tm.ev[i][3]) == cudaSuccess) {
        h->stats.scan_ms += ms;
        h->stats.scan_launches += 1;
    }
    if (cudaEventElapsedTime(&ms, tm.ev[i][3], tm.ev[i][5]) == cudaSuccess) h->stats.merge_ms += ms;
    tm.pending[i] = 0;
}

void flush_all(ivfadc_index* h) {
    for (int i = 0; i < kEventSets; ++i) flush_set(h, i);
    shard_flush_timing(h);
    cudaGetLastError();
}

int check_handle(const ivfadc_index* h) { return h ? IVFADC_OK : IVFADC_ERR_BAD_ARG; }

// One chunk of queries, everything on the device, asynchronous on `s`.
int search_chunk(ivfadc_index* h, const void* dQ, int64_t nq, int k, int w, uint64_t* d_ids, void* d_dists,
                 uint64_t* d_keys, int32_t* d_counts, cudaStream_t s, const int32_t* ext_cells = nullptr,
                 const void* ext_dc = nullptr) {
    const ScanPlanSizes z = scan_plan_sizes(h, nq, w, k);
    CUDA_OR_FAIL(h, h->ws_cells.reserve(sizeof(int32_t) * (size_t)nq * w), "workspace");
    CUDA_OR_FAIL(h, h->ws_dc.reserve(h->tsize * (size_t)nq * w), "workspace");
    CUDA_OR_FAIL(h, h->ws_bucket.reserve(z.bucket_bytes), "workspace");
    CUDA_OR_FAIL(h, h->ws_sorted.reserve(z.sorted_bytes), "workspace");
    CUDA_OR_FAIL(h, h->ws_pair_d.reserve(z.pair_d_bytes), "workspace");
    CUDA_OR_FAIL(h, h->ws_pair_pos.reserve(z.pair_pos_bytes), "workspace");
    CUDA_OR_FAIL(h, h->ws_pair_cnt.reserve(z.pair_cnt_bytes), "workspace");
    CUDA_OR_FAIL(h, h->ws_thr.reserve(z.thr_bytes), "workspace");
    CUDA_OR_FAIL(h, h->ws_items.reserve(z.items_bytes), "workspace");

    Timing& tm = extra(h)->tm;
    int set = -1;
    if (h->stats_timing && tm.ok) {
        set = tm.next;
        tm.next = (tm.next + 1) % kEventSets;
        flush_set(h, set);
        h->ev[2] = tm.ev[set][2];
        h->ev[3] = tm.ev[set][3];
        cudaEventRecord(tm.ev[set][0], s);
    }
    int launches = 0;
    const int32_t* d_cells = ext_cells;
    const void* d_dc = ext_dc;
    if (!ext_cells) {  // probes supplied by the caller when the coarse step is sharded by query across ranks
        CUDA_OR_FAIL(h, launch_coarse(h, dQ, nq, w, h->ws_cells.as<int32_t>(), h->ws_dc.p, s, &launches), "coarse kernel");
        d_cells = h->ws_cells.as<int32_t>();
        d_dc = h->ws_dc.p;
    }
    if (set >= 0) cudaEventRecord(tm.ev[set][1], s);
    const bool timing = h->stats_timing;
    h->stats_timing = set >= 0;
    cudaError_t e = launch_search(h, dQ, nq, k, w, d_cells, d_dc, d_ids, d_dists, d_keys, d_counts,
                                  extra(h)->d_scanned, s, &launches);
    h->stats_timing = timing;
    CUDA_OR_FAIL(h, e, "scan kernels");
    if (set >= 0) {
        cudaEventRecord(tm.ev[set][5], s);
        tm.pending[set] = 1;
    }
    h->stats.gpu_launches += launches;
    h->stats.queries += nq;
    return IVFADC_OK;
}

// ---- Float64 indexes: the Float32 search twin ---------------------------------------------------------------
// The reference's own tests are Float64 (test/index.jl:7).  With default flags a LARGE Float64 batch -- one the
// cost model would give to the tensor-memory kernel -- is searched by a Float32 engine that shares this handle's
// lists (the CSR arenas hold bytes and ids, no floats) and holds the quantizers rounded to fp32: distances come
// back within the north_star tolerance (1e-5 relative; measured ~5e-7) widened to Float64.  Small batches, any
// tuning flag, IVFADC_FLAG_LUT_EXACT, sharded handles and caller-supplied probes keep the exact fp64 chain.
__global__ void narrow_kernel(const double* __restrict__ in, float* __restrict__ out, int64_t n, int* __restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = (float)in[i];
    out[i] = v;
    if (!(fabsf(v) <= 1e18f)) *bad = 40;   // outside what the fp32 pipeline can square: reported as a pipeline fault
}
__global__ void widen_kernel(const float* __restrict__ in, double* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)in[i];
}

struct ListsView {
    uint8_t* d_codes; void* d_ids; int64_t arena_cap, arena_used; int64_t *d_off, *d_len; int64_t n_total, n_local;
};
ListsView lists_view(const ivfadc_index* h) {
    return {h->d_codes, h->d_ids, h->arena_cap, h->arena_used, h->d_off, h->d_len, h->n_total, h->n_local};
}
void lists_set(ivfadc_index* h, const ListsView& v) {
    h->d_codes = v.d_codes; h->d_ids = v.d_ids; h->arena_cap = v.arena_cap; h->arena_used = v.arena_used;
    h->d_off = v.d_off; h->d_len = v.d_len; h->n_total = v.n_total; h->n_local = v.n_local;
}

// pipeline flag of the twin after a synchronised search (the twin's kernels report into its own flag word)
int twin_check(ivfadc_index* h) {
    if (!h->twin || !h->twin_used || !h->twin->d_err) return IVFADC_OK;
    h->twin_used = false;
    int flag = 0;
    if (cudaMemcpy(&flag, h->twin->d_err, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) {
        cudaGetLastError();
        return fail(h, IVFADC_ERR_CUDA, "D2H");
    }
    if (!flag) return IVFADC_OK;
    cudaMemset(h->twin->d_err, 0, sizeof(int));
    if (flag == 40) return fail(h, IVFADC_ERR_UNSUPPORTED, "Float64 query outside the Float32 range of the default search path: use IVFADC_FLAG_LUT_EXACT");
    return api_check_pipeline_flag(h, flag);
}

bool twin_ensure(ivfadc_index* h) {
    if (h->twin) return true;
    if (h->twin_failed) return false;
    h->twin_failed = true;   // until everything below has worked
    const ivfadc_config& c = h->cfg;
    const size_t nc = (size_t)c.kc * c.dim, nb = (size_t)c.m * c.ksub * h->dsub, ncode = (size_t)c.m * c.ksub;
    std::vector<double> cd(nc), bd(nb);
    std::vector<uint8_t> codes(ncode);
    if (cudaMemcpy(cd.data(), h->d_centroids, nc * 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(bd.data(), h->d_cb, nb * 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(codes.data(), h->d_cb_codes, ncode, cudaMemcpyDeviceToHost) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    std::vector<float> cf(nc), bf(nb);
    for (size_t i = 0; i < nc; ++i) { cf[i] = (float)cd[i]; if (!(std::fabs(cf[i]) <= 1e18f)) return false; }
    for (size_t i = 0; i < nb; ++i) { bf[i] = (float)bd[i]; if (!(std::fabs(bf[i]) <= 1e18f)) return false; }
    ivfadc_config c32 = c;
    c32.dtype = IVFADC_F32;
    ivfadc_index* t = nullptr;
    if (ivfadc_create(&t, &c32, cf.data(), bf.data(), codes.data()) != IVFADC_OK) return false;
    if (!t->d_err && (cudaMalloc(&t->d_err, sizeof(int)) != cudaSuccess || cudaMemset(t->d_err, 0, sizeof(int)) != cudaSuccess)) {
        cudaGetLastError();
        ivfadc_destroy(t);
        return false;
    }
    h->twin = t;
    h->twin_failed = false;
    return true;
}

int search_core(ivfadc_index* h, const void* dQ, int64_t nq, int k, int w, uint64_t* d_ids, void* d_dists,
                uint64_t* d_keys, int32_t* d_counts, cudaStream_t s, const int32_t* ext_cells = nullptr,
                const void* ext_dc = nullptr);

// returns 1 when the twin declined (the caller continues on the exact path), else a status code
int search_via_twin(ivfadc_index* h, const void* dQ, int64_t nq, int k, int w, uint64_t* d_ids, void* d_dists,
                    uint64_t* d_keys, int32_t* d_counts, cudaStream_t s) {
    if (!twin_ensure(h)) return 1;
    ivfadc_index* t = h->twin;
    const ListsView own = lists_view(t);
    lists_set(t, lists_view(h));
    t->h_off.swap(h->h_off); t->h_len.swap(h->h_len); t->h_cap.swap(h->h_cap);   // O(1): handed back below
    const int wc = std::min(w, h->cfg.kc);
    int rc = 1;
    if (k <= scan_max_k() && wc <= coarse_max_w() && scan_takes_tensor_path(t, nq * (int64_t)wc, k)) {
        const size_t nqd = (size_t)nq * h->cfg.dim, nk = (size_t)nq * k;
        rc = IVFADC_OK;
        if (t->ws_q.reserve(nqd * 4) != cudaSuccess || t->ws_out_d.reserve(nk * 4) != cudaSuccess) rc = fail(h, IVFADC_ERR_OOM, "workspace");
        if (rc == IVFADC_OK) {
            narrow_kernel<<<(unsigned)((nqd + 255) / 256), 256, 0, s>>>(static_cast<const double*>(dQ), t->ws_q.as<float>(),
                                                                        (int64_t)nqd, t->d_err);
            t->stats_timing = h->stats_timing;
            rc = search_core(t, t->ws_q.p, nq, k, w, d_ids, t->ws_out_d.p, d_keys, d_counts, s, nullptr, nullptr);
            if (rc != IVFADC_OK) h->err = t->err;
        }
        if (rc == IVFADC_OK) {
            widen_kernel<<<(unsigned)((nk + 255) / 256), 256, 0, s>>>(t->ws_out_d.as<float>(), static_cast<double*>(d_dists), (int64_t)nk);
            if (cudaGetLastError() != cudaSuccess) rc = fail(h, IVFADC_ERR_CUDA, "widen kernel");
            h->stats.searches += 1;
            h->stats.queries += nq;
            h->stats.gpu_launches += 2;
            h->stats.last_scan_kernel = t->stats.last_scan_kernel;
            h->twin_used = true;
        }
    }
    lists_set(t, own);   // the kernels took the device pointers by value; the twin owns nothing of the parent's
    t->h_off.swap(h->h_off); t->h_len.swap(h->h_len); t->h_cap.swap(h->h_cap);
    t->scanw_ws_n = -1;
    return rc;
}

int search_core(ivfadc_index* h, const void* dQ, int64_t nq, int k, int w, uint64_t* d_ids, void* d_dists,
                uint64_t* d_keys, int32_t* d_counts, cudaStream_t s, const int32_t* ext_cells,
                const void* ext_dc) {
    if (!dQ || !d_ids || !d_dists || !d_counts) return fail(h, IVFADC_ERR_BAD_ARG, "null pointer");
    if (nq < 0) return fail(h, IVFADC_ERR_BAD_ARG, "nq < 0");
    if (k < 1) return fail(h, IVFADC_ERR_BAD_ARG, "Number of neighbors must be k >= 1");
    if (w < 1) return fail(h, IVFADC_ERR_BAD_ARG, "Number of clusters to search in must be w >= 1");
    if (ext_cells && w > h->cfg.kc) return fail(h, IVFADC_ERR_BAD_ARG, "w > kc with caller-supplied probes");
    w = std::min(w, h->cfg.kc);  // reference src/index.jl:216
    if (k > scan_max_k()) return fail(h, IVFADC_ERR_UNSUPPORTED, "k > 128 is not supported by the scan kernel yet");
    if (w > coarse_max_w()) return fail(h, IVFADC_ERR_UNSUPPORTED, "w > 128 is not supported by the coarse kernel yet");
    if (nq == 0) return IVFADC_OK;
    if (h->cfg.dtype == IVFADC_F64 && h->cfg.flags == 0 && h->cfg.shard_world == 1 && !ext_cells &&
        h->cfg.metric_coarse == IVFADC_SQEUCLIDEAN && !h->d_dbg_lut) {
        const int rc = search_via_twin(h, dQ, nq, k, w, d_ids, d_dists, d_keys, d_counts, s);
        if (rc != 1) return rc;
    }
    h->stats.searches += 1;
    // bound the per-pair candidate workspace (~512 MB): a pair's row holds k sorted entries, or 64 candidates
    // (distance + position) when the tensor-memory kernels serve the batch (pair_stride in scan.cu)
    const size_t per_query = (size_t)w * std::max(k, 64) * (h->tsize + 4) + 64;
    int64_t chunk = (int64_t)std::max<size_t>(1, ((size_t)512 << 20) / per_query);
    chunk = std::min<int64_t>(chunk, (int64_t)1 << 20);
    for (int64_t q0 = 0; q0 < nq; q0 += chunk) {
        const int64_t n = std::min(chunk, nq - q0);
        int rc = search_chunk(h, static_cast<const char*>(dQ) + (size_t)q0 * h->cfg.dim * h->tsize, n, k, w,
                              d_ids + q0 * k, static_cast<char*>(d_dists) + (size_t)q0 * k * h->tsize,
                              d_keys ? d_keys + q0 * k : nullptr, d_counts + q0, s,
                              ext_cells ? ext_cells + q0 * w : nullptr,
                              ext_dc ? static_cast<const char*>(ext_dc) + (size_t)q0 * w * h->tsize : nullptr);
        if (rc != IVFADC_OK) return rc;
    }
    return IVFADC_OK;
}

}  // namespace

namespace ivf {
int api_fail(ivfadc_index* h, int code, const char* msg, cudaError_t e) { return fail(h, code, msg, e); }
int api_search_core(ivfadc_index* h, const void* dQ, int64_t nq, int k, int w, uint64_t* d_ids, void* d_dists,
                    uint64_t* d_keys, int32_t* d_counts, cudaStream_t s, const int32_t* ext_cells, const void* ext_dc) {
    return search_core(h, dQ, nq, k, w, d_ids, d_dists, d_keys, d_counts, s, ext_cells, ext_dc);
}
int api_check_pipeline_flag(ivfadc_index* h, int flag) {
    if (!flag) return IVFADC_OK;
    if (h->d_err) cudaMemset(h->d_err, 0, sizeof(int));
    const char* what = flag >= 20 ? "tensor-memory scan kernel: a role's mbarrier wait timed out (staging / table release / candidates)"
                     : flag >= 11 ? "tensor-core coarse kernel: pipeline wait timed out (TMA / MMA / accumulator tile)"
                     : flag == 1 ? "tensor-core table builder: TMA operand never arrived"
                     : flag == 2 ? "tensor-core table builder: MMA never completed"
                     : flag == 3 ? "tensor-core table builder: shared-memory plan does not fit"
                     : flag == 4 ? "tensor-core table builder: first A operands of a work item never written"
                     : flag == 5 ? "tensor-core table builder: codebook ring did not drain"
                                 : "tensor-core table builder: table buffer never released";
    return fail(h, IVFADC_ERR_CUDA, what);
}
}  // namespace ivf

namespace {

// cells (+ optional codes) of a device-resident batch
int cells_and_codes(ivfadc_index* h, const void* dX, int64_t n, const int64_t* d_assign, int base,
                    int32_t* d_cells, uint8_t* d_codes, int* launches) {
    cudaStream_t s = h->stream;
    if (d_assign) {
        CUDA_OR_FAIL(h, launch_assign_to_cells(d_assign, n, base, d_cells, s, launches), "assign kernel");
    } else {
        CUDA_OR_FAIL(h, h->ws_dc.reserve(h->tsize * (size_t)n), "workspace");
        CUDA_OR_FAIL(h, launch_coarse(h, dX, n, 1, d_cells, h->ws_dc.p, s, launches), "coarse kernel");
    }
    if (d_codes) CUDA_OR_FAIL(h, launch_encode(h, dX, n, d_cells, d_codes, s, launches), "encode kernel");
    return IVFADC_OK;
}

// caller-supplied assignments (reference: kmeans .assignments, src/index.jl:170-172) must name existing cells:
// an out-of-range cell would index the centroid table and the append sort out of bounds
bool assign_in_range(const int64_t* assign, int64_t n, int base, int kc) {
    uint64_t bad = 0;
    for (int64_t i = 0; i < n; ++i) bad |= (uint64_t)((uint64_t)(assign[i] - base) >= (uint64_t)kc);
    return bad == 0;
}

constexpr int64_t kAddChunk = 1 << 20;

}  // namespace

extern "C" {

int ivfadc_abi_version(void) { return IVFADC_ABI_VERSION; }

int ivfadc_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return IVFADC_ERR_CUDA;
    }
    return n;
}

int ivfadc_create(ivfadc_index** out, const ivfadc_config* cfg, const void* centroids,
                  const void* codebook_vectors, const uint8_t* codebook_codes) {
    IVF_NVTX();
    if (!out || !cfg || !centroids || !codebook_vectors || !codebook_codes) return IVFADC_ERR_BAD_ARG;
    *out = nullptr;
    if (cfg->dim < 1 || cfg->kc < 1 || cfg->m < 1 || cfg->m > cfg->dim || cfg->ksub < 1) return IVFADC_ERR_BAD_ARG;
    if (cfg->dtype != IVFADC_F32 && cfg->dtype != IVFADC_F64) return IVFADC_ERR_BAD_ARG;
    if (cfg->id_bytes != 1 && cfg->id_bytes != 2 && cfg->id_bytes != 4 && cfg->id_bytes != 8)
        return IVFADC_ERR_UNSUPPORTED;
    if (cfg->metric_coarse < IVFADC_SQEUCLIDEAN || cfg->metric_coarse > IVFADC_COSINEDIST ||
        cfg->metric_resid < IVFADC_SQEUCLIDEAN || cfg->metric_resid > IVFADC_COSINEDIST)
        return IVFADC_ERR_UNSUPPORTED;
    if (cfg->ksub > 256) return IVFADC_ERR_UNSUPPORTED;
    if (cfg->shard_world < 1 || cfg->shard_rank < 0 || cfg->shard_rank >= cfg->shard_world)
        return IVFADC_ERR_BAD_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || cfg->device < 0 || cfg->device >= ndev) {
        cudaGetLastError();
        return IVFADC_ERR_CUDA;  // no CPU fallback
    }
    if (cudaSetDevice(cfg->device) != cudaSuccess) return IVFADC_ERR_CUDA;

    ivfadc_index* h = new (std::nothrow) ivfadc_index();
    if (!h) return IVFADC_ERR_OOM;
    Extra* x = new (std::nothrow) Extra();
    if (!x) {
        delete h;
        return IVFADC_ERR_OOM;
    }
    h->cfg = *cfg;
    h->dsub = cfg->dim / cfg->m;  // QuantizedArrays.rowrange: floor(D / m)
    h->tsize = cfg->dtype == IVFADC_F32 ? 4 : 8;
    h->id_dev_bytes = cfg->id_bytes <= 4 ? 4 : 8;
    h->extra = x;
    if (cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, cfg->device) != cudaSuccess) h->num_sms = 0;

    const size_t cbytes = (size_t)cfg->kc * cfg->dim * h->tsize;
    const size_t vbytes = (size_t)cfg->m * cfg->ksub * h->dsub * h->tsize;
    const size_t kbytes = (size_t)cfg->m * cfg->ksub;
    bool ok = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_centroids, cbytes) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_cb, vbytes + 16) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_cb_codes, kbytes) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_cb_norms, kbytes * h->tsize) == cudaSuccess;
    ok = ok && cudaMalloc(&x->d_scanned, sizeof(uint64_t)) == cudaSuccess;
    ok = ok && cudaMemset(x->d_scanned, 0, sizeof(uint64_t)) == cudaSuccess;
    ok = ok && cudaMemcpy(h->d_centroids, centroids, cbytes, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(h->d_cb, codebook_vectors, vbytes, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(h->d_cb_codes, codebook_codes, kbytes, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && lists_init(h) == cudaSuccess;
    if (ok) {
        x->tm.ok = true;
        for (int i = 0; i < kEventSets && x->tm.ok; ++i) {
            x->tm.pending[i] = 0;
            for (int j = 0; j < kEventsPerSet; ++j)
                if (cudaEventCreate(&x->tm.ev[i][j]) != cudaSuccess) x->tm.ok = false;
        }
        ok = x->tm.ok;
    }
    if (cfg->dtype == IVFADC_F32) {
        // fp16 operand scales of the tensor-memory scan (scanw_impl.cuh): -2 w 2^ew and |w|^2 2^en (norm of an
        // 8-dim piece of a codeword) just below 2^14
        const float* cbv = static_cast<const float*>(codebook_vectors);
        const int dsub = h->dsub;
        double maxw = 0.0, maxn = 0.0;
        for (size_t r = 0; r < (size_t)cfg->m * cfg->ksub; ++r)
            for (int d0 = 0; d0 < dsub; d0 += 8) {
                double nn = 0.0;
                for (int d = d0; d < std::min(dsub, d0 + 8); ++d) {
                    const double v = cbv[r * dsub + d];
                    if (std::isfinite(v)) {
                        maxw = std::max(maxw, std::fabs(v));
                        nn += v * v;
                    }
                }
                maxn = std::max(maxn, nn);
            }
        int e = 0;
        if (maxw > 0.0) { std::frexp(2.0 * maxw, &e); h->tch_ew = 14 - e; }
        if (maxn > 0.0) { std::frexp(maxn, &e); h->tch_en = 14 - e; }
        h->tch_ew = std::max(-60, std::min(60, h->tch_ew));
        h->tch_en = std::max(-60, std::min(60, h->tch_en));
    }
    h->cb_identity = 1;
    for (int i = 0; i < cfg->m && h->cb_identity; ++i)
        for (int c = 0; c < cfg->ksub; ++c)
            if (codebook_codes[(size_t)i * cfg->ksub + c] != (uint8_t)c) {
                h->cb_identity = 0;
                break;
            }
    int launches = 0;
    ok = ok && launch_codebook_norms(h, h->stream, &launches) == cudaSuccess;
    ok = ok && scanq_prepare(h, h->stream, &launches) == cudaSuccess;
    ok = ok && coarse_prepare(h, h->stream, &launches) == cudaSuccess;
    ok = ok && cudaStreamSynchronize(h->stream) == cudaSuccess;
    std::string why;
    if (ok && (!scan_supported(h, &why) || !encode_supported(h))) {
        ivfadc_destroy(h);
        return IVFADC_ERR_UNSUPPORTED;
    }
    if (!ok) {
        cudaGetLastError();
        ivfadc_destroy(h);
        return IVFADC_ERR_CUDA;
    }
    h->stats.gpu_launches += launches;
    *out = h;
    return IVFADC_OK;
}

int ivfadc_destroy(ivfadc_index* h) {
    if (!h) return IVFADC_OK;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    if (h->twin) {
        ivfadc_destroy(h->twin);
        h->twin = nullptr;
    }
    shard_destroy_ctx(h);
    if (h->d_owner) cudaFree(h->d_owner);
    Extra* x = extra(h);
    if (x) {
        if (x->tm.ok)
            for (int i = 0; i < kEventSets; ++i)
                for (int j = 0; j < kEventsPerSet; ++j) cudaEventDestroy(x->tm.ev[i][j]);
        if (x->d_scanned) cudaFree(x->d_scanned);
        delete x;
    }
    lists_free(h);
    if (h->d_centroids) cudaFree(h->d_centroids);
    if (h->d_centroids_t) cudaFree(h->d_centroids_t);
    if (h->d_tcC) cudaFree(h->d_tcC);
    if (h->d_ccn) cudaFree(h->d_ccn);
    h->ws_coarse_redo.release();
    if (h->d_cb) cudaFree(h->d_cb);
    if (h->d_cb_codes) cudaFree(h->d_cb_codes);
    if (h->d_cb_norms) cudaFree(h->d_cb_norms);
    if (h->d_afrag) cudaFree(h->d_afrag);
    if (h->d_wnfrag) cudaFree(h->d_wnfrag);
    if (h->d_tcU) cudaFree(h->d_tcU);
    if (h->d_tcH) cudaFree(h->d_tcH);
    if (h->d_err) cudaFree(h->d_err);
    if (h->h_err) cudaFreeHost(h->h_err);
    if (h->d_dbg_lut) cudaFree(h->d_dbg_lut);
    DevBuf* bufs[] = {&h->ws_q, &h->ws_cells, &h->ws_dc, &h->ws_bucket, &h->ws_sorted, &h->ws_pair_d,
                      &h->ws_pair_pos, &h->ws_pair_cnt, &h->ws_thr, &h->ws_out_ids, &h->ws_out_d,
                      &h->ws_out_cnt, &h->ws_out_keys, &h->ws_misc, &h->ws_x, &h->ws_codes, &h->ws_assign,
                      &h->ws_sort_tmp, &h->ws_sort_keys, &h->ws_sort_vals, &h->ws_del, &h->ws_items};
    for (DevBuf* b : bufs) b->release();
    if (h->stream) cudaStreamDestroy(h->stream);
    cudaGetLastError();
    delete h;
    return IVFADC_OK;
}

const char* ivfadc_last_error(const ivfadc_index* h) { return h ? h->err.c_str() : "null handle"; }

int ivfadc_add(ivfadc_index* h, const void* X, int64_t n, int32_t position, const int64_t* assign,
               int32_t assign_base, int32_t* cells_out) {
    IVF_NVTX();
    if (check_handle(h)) return IVFADC_ERR_BAD_ARG;
    if (n < 0 || (n > 0 && !X)) return fail(h, IVFADC_ERR_BAD_ARG, "null data");
    if (position != IVFADC_LAST && position != IVFADC_FIRST) return fail(h, IVFADC_ERR_BAD_ARG, "bad position");
    if (n == 0) return IVFADC_OK;
    if (assign && !assign_in_range(assign, n, assign_base, h->cfg.kc))
        return fail(h, IVFADC_ERR_BAD_ARG, "assignment outside [assign_base, assign_base + kc)");
    cudaSetDevice(h->cfg.device);
    // reference src/utils.jl:134-135: bits(I) >= log2(N + 1) for every single push
    if (h->cfg.id_bytes < 8) {
        const uint64_t capacity = 1ull << (8 * h->cfg.id_bytes);
        if ((uint64_t)h->n_total + (uint64_t)n > capacity)
            return fail(h, IVFADC_ERR_CAPACITY, "Cannot index, exceeding index capacity");
    }
    const int D = h->cfg.dim, m = h->cfg.m;
    int launches = 0;
    if (position == IVFADC_FIRST)
        CUDA_OR_FAIL(h, lists_shift_ids(h, n, &launches), "shift ids");  // _shift_up_inverse_index!
    for (int64_t j0 = 0; j0 < n; j0 += kAddChunk) {
        const int64_t nb = std::min(kAddChunk, n - j0);
        CUDA_OR_FAIL(h, h->ws_x.reserve((size_t)nb * D * h->tsize), "workspace");
        CUDA_OR_FAIL(h, h->ws_cells.reserve(sizeof(int32_t) * (size_t)nb), "workspace");
        CUDA_OR_FAIL(h, h->ws_codes.reserve((size_t)nb * m), "workspace");
        CUDA_OR_FAIL(h, cudaMemcpyAsync(h->ws_x.p, static_cast<const char*>(X) + (size_t)j0 * D * h->tsize,
                                        (size_t)nb * D * h->tsize, cudaMemcpyHostToDevice, h->stream), "H2D");
        const int64_t* d_assign = nullptr;
        if (assign) {
            CUDA_OR_FAIL(h, h->ws_assign.reserve(sizeof(int64_t) * (size_t)nb), "workspace");
            CUDA_OR_FAIL(h, cudaMemcpyAsync(h->ws_assign.p, assign + j0, sizeof(int64_t) * nb,
                                            cudaMemcpyHostToDevice, h->stream), "H2D");
            d_assign = h->ws_assign.as<int64_t>();
        }
        int32_t* d_cells = h->ws_cells.as<int32_t>();
        int rc = cells_and_codes(h, h->ws_x.p, nb, d_assign, assign_base, d_cells, h->ws_codes.as<uint8_t>(),
                                 &launches);
        if (rc != IVFADC_OK) return rc;
        if (cells_out)
            CUDA_OR_FAIL(h, cudaMemcpyAsync(cells_out + j0, d_cells, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost,
                                            h->stream), "D2H");
        const uint64_t first_id = position == IVFADC_LAST ? (uint64_t)h->n_total + (uint64_t)j0
                                                          : (uint64_t)(n - 1 - j0);
        CUDA_OR_FAIL(h, lists_append(h, d_cells, h->ws_codes.as<uint8_t>(), nb, first_id,
                                     position == IVFADC_LAST ? 1 : -1, &launches), "append");
    }
    h->n_total += n;
    h->stats.gpu_launches += launches;
    return IVFADC_OK;
}

int ivfadc_add_device(ivfadc_index* h, const void* dX, int64_t n, int32_t position, const int64_t* d_assign,
                      int32_t assign_base, int32_t* d_cells_out) {
    IVF_NVTX();
    if (check_handle(h)) return IVFADC_ERR_BAD_ARG;
    if (n < 0 || (n > 0 && !dX)) return fail(h, IVFADC_ERR_BAD_ARG, "null data");
    if (position != IVFADC_LAST && position != IVFADC_FIRST) return fail(h, IVFADC_ERR_BAD_ARG, "bad position");
    if (n == 0) return IVFADC_OK;
    cudaSetDevice(h->cfg.device);
    if (h->cfg.id_bytes < 8) {   // reference src/utils.jl:134-135
        const uint64_t capacity = 1ull << (8 * h->cfg.id_bytes);
        if ((uint64_t)h->n_total + (uint64_t)n > capacity)
            return fail(h, IVFADC_ERR_CAPACITY, "Cannot index, exceeding index capacity");
    }
    const int D = h->cfg.dim, m = h->cfg.m;
    int launches = 0;
    if (d_assign) {   // same contract as ivfadc_add: cells must exist (checked on the device, one flag back)
        CUDA_OR_FAIL(h, h->ws_misc.reserve(sizeof(int)), "workspace");
        CUDA_OR_FAIL(h, cudaMemsetAsync(h->ws_misc.p, 0, sizeof(int), h->stream), "memset");
        CUDA_OR_FAIL(h, launch_assign_check(d_assign, n, assign_base, h->cfg.kc, h->ws_misc.as<int>(), h->stream,
                                            &launches), "assign check");
        int bad = 0;
        CUDA_OR_FAIL(h, cudaMemcpyAsync(&bad, h->ws_misc.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream), "D2H");
        CUDA_OR_FAIL(h, cudaStreamSynchronize(h->stream), "sync");
        if (bad) return fail(h, IVFADC_ERR_BAD_ARG, "assignment outside [assign_base, assign_base + kc)");
    }
    if (position == IVFADC_FIRST) CUDA_OR_FAIL(h, lists_shift_ids(h, n, &launches), "shift ids");
    for (int64_t j0 = 0; j0 < n; j0 += kAddChunk) {
        const int64_t nb = std::min(kAddChunk, n - j0);
        CUDA_OR_FAIL(h, h->ws_cells.reserve(sizeof(int32_t) * (size_t)nb), "workspace");
        CUDA_OR_FAIL(h, h->ws_codes.reserve((size_t)nb * m), "workspace");
        int32_t* d_cells = h->ws_cells.as<int32_t>();
        int rc = cells_and_codes(h, static_cast<const char*>(dX) + (size_t)j0 * D * h->tsize, nb,
                                 d_assign ? d_assign + j0 : nullptr, assign_base, d_cells, h->ws_codes.as<uint8_t>(),
                                 &launches);
        if (rc != IVFADC_OK) return rc;
        if (d_cells_out)
            CUDA_OR_FAIL(h, cudaMemcpyAsync(d_cells_out + j0, d_cells, sizeof(int32_t) * nb, cudaMemcpyDeviceToDevice,
                                            h->stream), "D2D");
        const uint64_t first_id = position == IVFADC_LAST ? (uint64_t)h->n_total + (uint64_t)j0
                                                          : (uint64_t)(n - 1 - j0);
        CUDA_OR_FAIL(h, lists_append(h, d_cells, h->ws_codes.as<uint8_t>(), nb, first_id,
                                     position == IVFADC_LAST ? 1 : -1, &launches), "append");
    }
    h->n_total += n;
    h->stats.gpu_launches += launches;
    return IVFADC_OK;
}

int ivfadc_encode(ivfadc_index* h, const void* X, int64_t n, const int64_t* assign, int32_t assign_base,
                  int32_t* cells_out, uint8_t* codes_out) {
    IVF_NVTX();
    if (check_handle(h)) return IVFADC_ERR_BAD_ARG;
    if (n < 0 || (n > 0 && (!X || !cells_out || !codes_out))) return fail(h, IVFADC_ERR_BAD_ARG, "null pointer");
    if (assign && !assign_in_range(assign, n, assign_base, h->cfg.kc))
        return fail(h, IVFADC_ERR_BAD_ARG, "assignment outside [assign_base, assign_base + kc)");
    cudaSetDevice(h->cfg.device);
    const int D = h->cfg.dim, m = h->cfg.m;
    int launches = 0;
    for (int64_t j0 = 0; j0 < n; j0 += kAddChunk) {
        const int64_t nb = std::min(kAddChunk, n - j0);
        CUDA_OR_FAIL(h, h->ws_x.reserve((size_t)nb * D * h->tsize), "workspace");
        CUDA_OR_FAIL(h, h->ws_cells.reserve(sizeof(int32_t) * (size_t)nb), "workspace");
        CUDA_OR_FAIL(h, h->ws_codes.reserve((size_t)nb * m), "workspace");
        CUDA_OR_FAIL(h, cudaMemcpyAsync(h->ws_x.p, static_cast<const char*>(X) + (size_t)j0 * D * h->tsize,
                                        (size_t)nb * D * h->tsize, cudaMemcpyHostToDevice, h->stream), "H2D");
        const int64_t* d_assign = nullptr;
        if (assign) {
            CUDA_OR_FAIL(h, h->ws_assign.reserve(sizeof(int64_t) * (size_t)nb), "workspace");
            CUDA_OR_FAIL(h, cudaMemcpyAsync(h->ws_assign.p, assign + j0, sizeof(int64_t) * nb,
                                            cudaMemcpyHostToDevice, h->stream), "H2D");
            d_assign = h->ws_assign.as<int64_t>();
        }
        int rc = cells_and_codes(h, h->ws_x.p, nb, d_assign, assign_base, h->ws_cells.as<int32_t>(),
                                 h->ws_codes.as<uint8_t>(), &launches);
        if (rc != IVFADC_OK) return rc;
        CUDA_OR_FAIL(h, cudaMemcpyAsync(cells_out + j0, h->ws_cells.p, sizeof(int32_t) * nb,
                                        cudaMemcpyDeviceToHost, h->stream), "D2H");
        CUDA_OR_FAIL(h, cudaMemcpyAsync(codes_out + (size_t)j0 * m, h->ws_codes.p, (size_t)nb * m,
                                        cudaMemcpyDeviceToHost, h->stream), "D2H");
        CUDA_OR_FAIL(h, cudaStreamSynchronize(h->stream), "sync");
    }
    h->stats.gpu_launches += launches;
    return IVFADC_OK;
}

int ivfadc_coarse_search(ivfadc_index* h, const void* Q, int64_t nq, int32_t w, int32_t* cells_out,
                         void* dc_out) {
    IVF_NVTX();
    if (check_handle(h)) return IVFADC_ERR_BAD_ARG;
    if (nq < 0 || (nq > 0 && (!Q || !cells_out || !dc_out))) return fail(h, IVFADC_ERR_BAD_ARG, "null pointer");
    if (w < 1) return fail(h, IVFADC_ERR_BAD_ARG, "w < 1");
    if (w > h->cfg.kc) return fail(h, IVFADC_ERR_BAD_ARG, "w > kc (clamp before calling)");
    if (w > coarse_max_w()) return fail(h, IVFADC_ERR_UNSUPPORTED, "w > 128");
    if (nq == 0) return IVFADC_OK;
    cudaSetDevice(h->cfg.device);
    int launches = 0;
    CUDA_OR_FAIL(h, h->ws_q.reserve((size_t)nq * h->cfg.dim * h->tsize), "workspace");
    CUDA_OR_FAIL(h, h->ws_cells.reserve(sizeof(int32_t) * (size_t)nq * w), "workspace");
    CUDA_OR_FAIL(h, h->ws_dc.reserve(h->tsize * (size_t)nq * w), "workspace");
    CUDA_OR_FAIL(h, cudaMemcpyAsync(h->ws_q.p, Q, (size_t)nq * h->cfg.dim * h->tsize, cudaMemcpyHostToDevice,
                                    h->stream), "H2D");
    CUDA_OR_FAIL(h, launch_coarse(h, h->ws_q.p, nq, w, h->ws_cells.as<int32_t>(), h->ws_dc.p, h->stream,
                                  &launches), "coarse kernel");
    CUDA_OR_FAIL(h, cudaMemcpyAsync(cells_out, h->ws_cells.p, sizeof(int32_t) * (size_t)nq * w,
                                    cudaMemcpyDeviceToHost, h->stream), "D2H");
    CUDA_OR_FAIL(h, cudaMemcpyAsync(dc_out, h->ws_dc.p, h->tsize * (size_t)nq * w, cudaMemcpyDeviceToHost,
                                    h->stream), "D2H");
    CUDA_OR_FAIL(h, cudaStreamSynchronize(h->stream), "sync");
    h->stats.gpu_launches += launches;
    return IVFADC_OK;
}

int ivfadc_search(ivfadc_index* h, const void* Q, int64_t nq, int32_t k, int32_t w, uint64_t* ids_out,
                  void* dists_out, int32_t* counts_out) {
    IVF_NVTX();
    if (check_handle(h)) return IVFADC_ERR_BAD_ARG;
    if (nq > 0 && (!Q || !ids_out || !dists_out || !counts_out)) return fail(h, IVFADC_ERR_BAD_ARG, "null pointer");
    if (k < 1) return fail(h, IVFADC_ERR_BAD_ARG, "Number of neighbors must be k >= 1");
    if (w < 1) return fail(h, IVFADC_ERR_BAD_ARG, "Number of clusters to search in must be w >= 1");
    if (nq <= 0) return nq == 0 ? IVFADC_OK : fail(h, IVFADC_ERR_BAD_ARG, "nq < 0");
    cudaSetDevice(h->cfg.device);
    const size_t qbytes = (size_t)nq * h->cfg.dim * h->tsize;
    CUDA_OR_FAIL(h, h->ws_q.reserve(qbytes), "workspace");
    CUDA_OR_FAIL(h, h->ws_out_ids.reserve(sizeof(uint64_t) * (size_t)nq * k), "workspace");
    CUDA_OR_FAIL(h, h->ws_out_d.reserve(h->tsize * (size_t)nq * k), "workspace");
    CUDA_OR_FAIL(h, h->ws_out_cnt.reserve(sizeof(int32_t) * (size_t)nq), "workspace");
    CUDA_OR_FAIL(h, cudaMemcpyAsync(h->ws_q.p, Q, qbytes, cudaMemcpyHostToDevice, h->stream), "H2D");
    int rc = search_core(h, h->ws_q.p, nq, k, w, h->ws_out_ids.as<uint64_t>(), h->ws_out_d.p, nullptr,
                         h->ws_out_cnt.as<int32_t>(), h->stream);
    if (rc != IVFADC_OK) return rc;
    CUDA_OR_FAIL(h, cudaMemcpyAsync(ids_out, h->ws_out_ids.p, sizeof(uint64_t) * (size_t)nq * k,
                                    cudaMemcpyDeviceToHost, h->stream), "D2H");
    CUDA_OR_FAIL(h, cudaMemcpyAsync(dists_out, h->ws_out_d.p, h->tsize * (size_t)nq * k, cudaMemcpyDeviceToHost,
                                    h->stream), "D2H");
    CUDA_OR_FAIL(h, cudaMemcpyAsync(counts_out, h->ws_out_cnt.p, sizeof(int32_t) * (size_t)nq,
                                    cudaMemcpyDeviceToHost, h->stream), "D2H");
    // the error flag of the tensor-core pipeline rides in the same asynchronous batch (pinned word)
    if (h->d_err && !h->h_err) {
        if (cudaMallocHost(reinterpret_cast<void**>(&h->h_err), sizeof(int)) != cudaSuccess) h->h_err = nullptr;
    }
    if (h->d_err && h->h_err)
        CUDA_OR_FAIL(h, cudaMemcpyAsync(h->h_err, h->d_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream), "D2H");
    CUDA_OR_FAIL(h, cudaStreamSynchronize(h->stream), "search");
    if (h->d_err) {
        int flag = 0;
        if (h->h_err) flag = *h->h_err;
        else CUDA_OR_FAIL(h, cudaMemcpy(&flag, h->d_err, sizeof(int), cudaMemcpyDeviceToHost), "D2H");
        if (flag) return api_check_pipeline_flag(h, flag);
    }
    return twin_check(h);
}

int ivfadc_search_device(ivfadc_index* h, const void* dQ, int64_t nq, int32_t k, int32_t w, uint64_t* d_ids,
                         void* d_dists, int32_t* d_counts, void* stream) {
    IVF_NVTX();
    if (check_handle(h)) return IVFADC_ERR_BAD_ARG;
    cudaSetDevice(h->cfg.device);
    return search_core(h, dQ, nq, k, w, d_ids, d_dists, nullptr, d_counts, static_cast<cudaStream_t>(stream));
}

int ivfadc_search_local_device(ivfadc_index* h, const void* dQ, int64_t nq, int32_t k, int32_t w,
                               uint64_t* d_ids, void* d_dists, uint64_t* d_keys, int32_t* d_counts,
                               void* stream) {
    IVF_NVTX();
    if (check_handle(h)) return IVFADC_ERR_BAD_ARG;
    if (!d_keys) return fail(h, IVFADC_ERR_BAD_ARG, "null keys");
    cudaSetDevice(h->cfg.device);
    return search_core(h, dQ, nq, k, w, d_ids, d_dists, d_keys, d_counts, static_cast<cudaStream_t>(stream));
}

int ivfadc_coarse_search_device(ivfadc_index* h, const void* dQ, int64_t nq, int32_t w, int32_t* d_cells,
                                void* d_dc, void* stream) {
    IVF_NVTX();
    if (check_handle(h)) return IVFADC_ERR_BAD_ARG;
    if (nq < 0 || (nq > 0 && (!dQ || !d_cells || !d_dc))) return fail(h, IVFADC_ERR_BAD_ARG, "null pointer");
    if (w < 1 || w > h->cfg.kc) return fail(h, IVFADC_ERR_BAD_ARG, "w outside 1..kc (clamp before calling)");
    if (w > coarse_max_w()) return fail(h, IVFADC_ERR_UNSUPPORTED, "w > 128");
    if (nq == 0) return IVFADC_OK;
    cudaSetDevice(h->cfg.device);
    int launches = 0;
    CUDA_OR_FAIL(h, launch_coarse(h, dQ, nq, w, d_cells, d_dc, static_cast<cudaStream_t>(stream), &launches),
                 "coarse kernel");
    h->stats.gpu_launches += launches;
    return IVFADC_OK;
}

int ivfadc_search_probes_local_device(ivfadc_index* h, const void* dQ, int64_t nq, int32_t k, int32_t w,
                                      const int32_t* d_cells, const void* d_dc, uint64_t* d_ids, void* d_dists,
                                      uint64_t* d_keys, int32_t* d_counts, void* stream) {
    IVF_NVTX();
    if (check_handle(h)) return IVFADC_ERR_BAD_ARG;
    if (!d_keys || !d_cells || !d_dc) return fail(h, IVFADC_ERR_BAD_ARG, "null pointer");
    cudaSetDevice(h->cfg.device);
    return search_core(h, dQ, nq, k, w, d_ids, d_dists, d_keys, d_counts, static_cast<cudaStream_t>(stream), d_cells,
                       d_dc);
}

int ivfadc_merge_device(ivfadc_index* h, int32_t parts, int64_t nq, int32_t k, const uint64_t* d_ids_in,
                        const void* d_dists_in, const uint64_t* d_keys_in, uint64_t* d_ids, void* d_dists,
                        int32_t* d_counts, void* stream) {
    IVF_NVTX();
    if (check_handle(h)) return IVFADC_ERR_BAD_ARG;
    if (parts < 1 || parts > 128 || k < 1 || nq < 0) return fail(h, IVFADC_ERR_BAD_ARG, "bad merge shape");
    if (!d_ids_in || !d_dists_in || !d_keys_in || !d_ids || !d_dists || !d_counts)
        return fail(h, IVFADC_ERR_BAD_ARG, "null pointer");
    if (nq == 0) return IVFADC_OK;
    cudaSetDevice(h->cfg.device);
    int launches = 0;
    CUDA_OR_FAIL(h, launch_merge_parts(h, parts, nq, k, d_ids_in, d_dists_in, d_keys_in, d_ids, d_dists, d_counts,
                                       static_cast<cudaStream_t>(stream), &launches), "merge kernel");
    h->stats.gpu_launches += launches;
    return IVFADC_OK;
}

int ivfadc_delete(ivfadc_index* h, const uint64_t* ids, int64_t n) {
    IVF_NVTX();
    if (check_handle(h)) return IVFADC_ERR_BAD_ARG;
    if (n < 0 || (n > 0 && !ids)) return fail(h, IVFADC_ERR_BAD_ARG, "null ids");
    if (n == 0) return IVFADC_OK;
    cudaSetDevice(h->cfg.device);
    // sort(unique(points)) -- reference src/utils.jl:94; ids >= N are unknown and ignored
    std::vector<uint64_t> v(ids, ids + n);
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
    while (!v.empty() && v.back() >= (uint64_t)h->n_total) v.pop_back();
    if (v.empty()) return IVFADC_OK;
    CUDA_OR_FAIL(h, h->ws_del.reserve(sizeof(uint64_t) * v.size()), "workspace");
    CUDA_OR_FAIL(h, cudaMemcpyAsync(h->ws_del.p, v.data(), sizeof(uint64_t) * v.size(), cudaMemcpyHostToDevice,
                                    h->stream), "H2D");
    int launches = 0;
    int64_t removed = 0;
    CUDA_OR_FAIL(h, lists_delete(h, h->ws_del.as<uint64_t>(), (int64_t)v.size(), &removed, &launches), "delete");
    CUDA_OR_FAIL(h, cudaStreamSynchronize(h->stream), "sync");
    h->n_total -= (int64_t)v.size();  // every id in [0, N) exists exactly once across the shards
    h->stats.gpu_launches += launches;
    if (h->cfg.shard_world == 1 && removed != (int64_t)v.size())
        return fail(h, IVFADC_ERR_CUDA, "internal: id set of the index is not dense");
    return IVFADC_OK;
}

int ivfadc_pop(ivfadc_index* h, int32_t position, void* vec_out, int32_t* found_out) {
    IVF_NVTX();
    if (check_handle(h)) return IVFADC_ERR_BAD_ARG;
    if (!vec_out) return fail(h, IVFADC_ERR_BAD_ARG, "null output");
    if (position != IVFADC_LAST && position != IVFADC_FIRST) return fail(h, IVFADC_ERR_BAD_ARG, "bad position");
    if (h->n_total <= 0) return fail(h, IVFADC_ERR_EMPTY, "Cannot pop element from empty index");
    cudaSetDevice(h->cfg.device);
    const uint64_t vecid = position == IVFADC_LAST ? (uint64_t)h->n_total - 1 : 0;  // src/utils.jl:47
    int launches = 0;
    int32_t cell = -1;
    int64_t pos = -1;
    CUDA_OR_FAIL(h, lists_find(h, vecid, &cell, &pos, &launches), "find");
    if (found_out) *found_out = cell >= 0;
    if (cell >= 0) {
        CUDA_OR_FAIL(h, h->ws_x.reserve((size_t)h->cfg.dim * h->tsize), "workspace");
        CUDA_OR_FAIL(h, lists_decode(h, cell, pos, h->ws_x.p, &launches), "decode");
        CUDA_OR_FAIL(h, cudaMemcpyAsync(vec_out, h->ws_x.p, (size_t)h->cfg.dim * h->tsize, cudaMemcpyDeviceToHost,
                                        h->stream), "D2H");
        CUDA_OR_FAIL(h, cudaStreamSynchronize(h->stream), "sync");
    } else if (h->cfg.shard_world == 1) {
        return fail(h, IVFADC_ERR_CUDA, "internal: id not found");
    }
    h->stats.gpu_launches += launches;
    return ivfadc_delete(h, &vecid, 1);  // deleteat! + _shift_down_inverse_index! (src/utils.jl:62-66)
}

int ivfadc_length(const ivfadc_index* h, int64_t* n_out) {
    if (!h || !n_out) return IVFADC_ERR_BAD_ARG;
    *n_out = h->n_total;
    return IVFADC_OK;
}

int ivfadc_list_sizes(ivfadc_index* h, int64_t* sizes_out) {
    if (check_handle(h) || !sizes_out) return IVFADC_ERR_BAD_ARG;
    std::copy(h->h_len.begin(), h->h_len.end(), sizes_out);
    return IVFADC_OK;
}

int ivfadc_export_list(ivfadc_index* h, int32_t cell, uint64_t* ids_out, uint8_t* codes_out) {
    IVF_NVTX();
    if (check_handle(h)) return IVFADC_ERR_BAD_ARG;
    if (cell < 0 || cell >= h->cfg.kc) return fail(h, IVFADC_ERR_BAD_ARG, "bad cell");
    if (h->h_len[cell] > 0 && (!ids_out || !codes_out)) return fail(h, IVFADC_ERR_BAD_ARG, "null pointer");
    cudaSetDevice(h->cfg.device);
    CUDA_OR_FAIL(h, lists_export(h, cell, ids_out, codes_out), "export");
    return IVFADC_OK;
}

int ivfadc_import_list(ivfadc_index* h, int32_t cell, const uint64_t* ids, const uint8_t* codes, int64_t len) {
    IVF_NVTX();
    if (check_handle(h)) return IVFADC_ERR_BAD_ARG;
    if (cell < 0 || cell >= h->cfg.kc || len < 0) return fail(h, IVFADC_ERR_BAD_ARG, "bad cell / length");
    if (len > 0 && (!ids || !codes)) return fail(h, IVFADC_ERR_BAD_ARG, "null pointer");
    cudaSetDevice(h->cfg.device);
    int launches = 0;
    CUDA_OR_FAIL(h, lists_import(h, cell, ids, codes, len, &launches), "import");
    h->stats.gpu_launches += launches;
    return IVFADC_OK;
}

int ivfadc_set_centroids_device(ivfadc_index* h, const void* d_centroids) {
    IVF_NVTX();
    if (check_handle(h) || !d_centroids) return IVFADC_ERR_BAD_ARG;
    if (h->n_total != 0) return fail(h, IVFADC_ERR_BAD_ARG, "centroids can only be replaced while the index is empty");
    cudaSetDevice(h->cfg.device);
    int launches = 0;
    CUDA_OR_FAIL(h, cudaMemcpyAsync(h->d_centroids, d_centroids, (size_t)h->cfg.kc * h->cfg.dim * h->tsize,
                                    cudaMemcpyDeviceToDevice, h->stream), "D2D");
    CUDA_OR_FAIL(h, coarse_prepare(h, h->stream, &launches), "coarse operands");
    CUDA_OR_FAIL(h, cudaStreamSynchronize(h->stream), "sync");
    h->stats.gpu_launches += launches;
    return IVFADC_OK;
}

int ivfadc_reserve(ivfadc_index* h, int64_t n_total, const int64_t* sizes) {
    IVF_NVTX();
    if (check_handle(h)) return IVFADC_ERR_BAD_ARG;
    if (n_total < 0) return fail(h, IVFADC_ERR_BAD_ARG, "negative size");
    cudaSetDevice(h->cfg.device);
    const int kc = h->cfg.kc;
    std::vector<int64_t> need(h->h_len);
    // without per-list sizes: an even share of this shard's part of n_total plus 12% (k-means cells are not balanced)
    const int64_t even = (n_total / std::max(1, h->cfg.shard_world) / kc) * 9 / 8 + 1;
    for (int c = 0; c < kc; ++c) {
        if (sizes && sizes[c] < 0) return fail(h, IVFADC_ERR_BAD_ARG, "negative list length");
        need[c] = std::max(need[c], sizes ? sizes[c] : even);
    }
    int launches = 0;
    CUDA_OR_FAIL(h, lists_reserve(h, need, &launches), "reserve");
    h->stats.gpu_launches += launches;
    return IVFADC_OK;
}

int ivfadc_export_all(ivfadc_index* h, uint64_t* ids_out, uint8_t* codes_out) {
    IVF_NVTX();
    if (check_handle(h)) return IVFADC_ERR_BAD_ARG;
    if (h->n_local > 0 && (!ids_out || !codes_out)) return fail(h, IVFADC_ERR_BAD_ARG, "null pointer");
    cudaSetDevice(h->cfg.device);
    int launches = 0;
    CUDA_OR_FAIL(h, lists_export_all(h, ids_out, codes_out, &launches), "export");
    h->stats.gpu_launches += launches;
    return IVFADC_OK;
}

int ivfadc_import_all(ivfadc_index* h, const int64_t* sizes, const uint64_t* ids, const uint8_t* codes) {
    IVF_NVTX();
    if (check_handle(h) || !sizes) return IVFADC_ERR_BAD_ARG;
    int64_t total = 0;
    for (int c = 0; c < h->cfg.kc; ++c) {
        if (sizes[c] < 0) return fail(h, IVFADC_ERR_BAD_ARG, "negative list length");
        total += sizes[c];
    }
    if (total > 0 && (!ids || !codes)) return fail(h, IVFADC_ERR_BAD_ARG, "null pointer");
    cudaSetDevice(h->cfg.device);
    int launches = 0;
    CUDA_OR_FAIL(h, lists_import_all(h, sizes, ids, codes, &launches), "import");
    h->stats.gpu_launches += launches;
    return IVFADC_OK;
}

int ivfadc_export_quantizers(ivfadc_index* h, void* centroids_out, void* codebook_vectors_out,
                             uint8_t* codebook_codes_out) {
    if (check_handle(h)) return IVFADC_ERR_BAD_ARG;
    cudaSetDevice(h->cfg.device);
    const ivfadc_config& c = h->cfg;
    if (centroids_out)
        CUDA_OR_FAIL(h, cudaMemcpy(centroids_out, h->d_centroids, (size_t)c.kc * c.dim * h->tsize,
                                   cudaMemcpyDeviceToHost), "D2H");
    if (codebook_vectors_out)
        CUDA_OR_FAIL(h, cudaMemcpy(codebook_vectors_out, h->d_cb, (size_t)c.m * c.ksub * h->dsub * h->tsize,
                                   cudaMemcpyDeviceToHost), "D2H");
    if (codebook_codes_out)
        CUDA_OR_FAIL(h, cudaMemcpy(codebook_codes_out, h->d_cb_codes, (size_t)c.m * c.ksub,
                                   cudaMemcpyDeviceToHost), "D2H");
    return IVFADC_OK;
}

int ivfadc_set_length(ivfadc_index* h, int64_t n_total) {
    if (check_handle(h) || n_total < 0) return IVFADC_ERR_BAD_ARG;
    h->n_total = n_total;
    return IVFADC_OK;
}

int ivfadc_set_cell_owners(ivfadc_index* h, const int32_t* owners) {
    if (check_handle(h) || !owners) return IVFADC_ERR_BAD_ARG;
    if (h->n_local != 0 || h->n_total != 0) return fail(h, IVFADC_ERR_BAD_ARG, "cell owners must be set on an empty index");
    const int kc = h->cfg.kc, world = h->cfg.shard_world;
    for (int c = 0; c < kc; ++c)
        if (owners[c] < 0 || owners[c] >= world) return fail(h, IVFADC_ERR_BAD_ARG, "cell owner outside [0, shard_world)");
    cudaSetDevice(h->cfg.device);
    if (!h->d_owner) CUDA_OR_FAIL(h, cudaMalloc(reinterpret_cast<void**>(&h->d_owner), sizeof(int32_t) * (size_t)kc), "owner map");
    CUDA_OR_FAIL(h, cudaMemcpy(h->d_owner, owners, sizeof(int32_t) * (size_t)kc, cudaMemcpyHostToDevice), "H2D");
    h->h_owner.assign(owners, owners + kc);
    return IVFADC_OK;
}

int ivfadc_check_async(ivfadc_index* h, void* stream) {
    if (check_handle(h)) return IVFADC_ERR_BAD_ARG;
    cudaSetDevice(h->cfg.device);
    CUDA_OR_FAIL(h, cudaStreamSynchronize(static_cast<cudaStream_t>(stream)), "sync");
    if (h->d_err) {
        int flag = 0;
        CUDA_OR_FAIL(h, cudaMemcpy(&flag, h->d_err, sizeof(int), cudaMemcpyDeviceToHost), "D2H");
        if (flag) return api_check_pipeline_flag(h, flag);
    }
    return twin_check(h);
}

int ivfadc_debug_tables(ivfadc_index* h, void* out) {
    if (check_handle(h)) return IVFADC_ERR_BAD_ARG;
    cudaSetDevice(h->cfg.device);
    const size_t bytes = ((size_t)h->cfg.m * 256 * 32 + 64 + 1024) * 4;  // tables, slots, cell, timeline stamps
    if (!out) {
        if (!h->d_dbg_lut) CUDA_OR_FAIL(h, cudaMalloc(&h->d_dbg_lut, bytes), "debug buffer");
        CUDA_OR_FAIL(h, cudaMemset(h->d_dbg_lut, 0, bytes), "debug buffer");
        return IVFADC_OK;
    }
    if (!h->d_dbg_lut) return fail(h, IVFADC_ERR_BAD_ARG, "debug dump not armed");
    CUDA_OR_FAIL(h, cudaDeviceSynchronize(), "sync");
    CUDA_OR_FAIL(h, cudaMemcpy(out, h->d_dbg_lut, bytes, cudaMemcpyDeviceToHost), "D2H");
    return IVFADC_OK;
}

int ivfadc_set_stats_timing(ivfadc_index* h, int32_t enable) {
    if (check_handle(h)) return IVFADC_ERR_BAD_ARG;
    cudaSetDevice(h->cfg.device);
    flush_all(h);
    h->stats_timing = enable != 0;
    return IVFADC_OK;
}

int ivfadc_get_stats(ivfadc_index* h, ivfadc_stats* out) {
    if (check_handle(h) || !out) return IVFADC_ERR_BAD_ARG;
    cudaSetDevice(h->cfg.device);
    flush_all(h);
    Extra* x = extra(h);
    uint64_t scanned = 0;
    if (cudaMemcpy(&scanned, x->d_scanned, sizeof(uint64_t), cudaMemcpyDeviceToHost) == cudaSuccess) {
        h->stats.scanned_vectors = scanned;
        h->stats.scan_code_bytes = scanned * (uint64_t)h->cfg.m;
    }
    h->stats.last_coarse_redo = 0;
    if (h->last_redo_nq > 0 && h->ws_coarse_redo.p) {   // redo flags of the last coarse step (one byte per query)
        std::vector<uint8_t> flags((size_t)h->last_redo_nq);
        if (cudaMemcpy(flags.data(), h->ws_coarse_redo.p, flags.size(), cudaMemcpyDeviceToHost) == cudaSuccess)
            for (uint8_t f : flags) h->stats.last_coarse_redo += f ? 1 : 0;
    }
    *out = h->stats;
    if (h->twin) {   // Float64 handle: the large batches ran on the Float32 twin
        ivfadc_stats ts;
        if (ivfadc_get_stats(h->twin, &ts) == IVFADC_OK) {
            out->coarse_ms += ts.coarse_ms; out->plan_ms += ts.plan_ms; out->scan_ms += ts.scan_ms; out->merge_ms += ts.merge_ms;
            out->scan_launches += ts.scan_launches; out->gpu_launches += ts.gpu_launches;
            out->scanned_vectors += ts.scanned_vectors; out->scan_code_bytes += ts.scan_code_bytes;
            if (ts.last_coarse_redo) out->last_coarse_redo = ts.last_coarse_redo;
        }
    }
    return IVFADC_OK;
}

int ivfadc_reset_stats(ivfadc_index* h) {
    if (check_handle(h)) return IVFADC_ERR_BAD_ARG;
    cudaSetDevice(h->cfg.device);
    flush_all(h);
    h->stats = ivfadc_stats{};
    cudaMemset(extra(h)->d_scanned, 0, sizeof(uint64_t));
    if (h->twin) ivfadc_reset_stats(h->twin);
    return IVFADC_OK;
}

}  // extern "C"
