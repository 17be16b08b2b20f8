// Cell-sharded multi-GPU search INSIDE the library: NCCL over NVLink / NVSwitch, CUDA-graph replay.
//
// The inverted lists shard by cell (owner map of ivfadc_set_cell_owners, or cell % world), the quantizers are
// replicated.  One batched knn_search (reference src/index.jl:261-273) on a sharded index is
//
//   coarse_search of THIS rank's slice of the queries            (launch_coarse, coarse.cu)
//   ONE grouped in-place all-gather: probe cells, probe distances (+ the query slices when they came from the host)
//   plan + list scan + per-query selection over the cells this rank owns   (api_search_core, scan.cu)
//   ONE grouped in-place all-gather: candidate ids, distances, merge keys  [world][nq][k]
//   merge by (distance, probe rank << 32 | position) -- the reference's order, src/index.jl:247-257
//   (from four ranks on: an all-to-all of candidate slices, every rank merges its slice of the queries, and a small
//    all-gather distributes the finished rows -- see slice_merge)
//
// enqueued on one stream with no host synchronisation in between, and replayed from a CUDA graph from the second
// call with the same shape on (the step is ~15 launches of 3..150 us: launch latency would otherwise set the pace).
// Grouped NCCL collectives on one communicator are aggregated into a single launch.
//
// Two ways to own the communicator:
//   * one process per GPU (bench.py under torchrun, tests): ivfadc_nccl_unique_id on rank 0, the 128 bytes travel
//     through the caller's own channel, ivfadc_comm_init_rank on every rank; then ivfadc_search_sharded[_device];
//   * one process, several GPUs (the Julia glue): ivfadc_group_* -- n handles, ncclCommInitAll, the same phases looped
//     over the devices with ncclGroupStart / ncclGroupEnd around the collectives.
// NCCL is loaded with dlopen at first use: libivfadc_cuda.so has no link-time dependency on it.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"

using namespace ivf;

namespace {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.ok ? &api : nullptr;
    tried = true;
    void* lib = nullptr;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) return nullptr;
#define IVF_SYM(field, sym) api.field = reinterpret_cast<decltype(api.field)>(dlsym(lib, sym))
    IVF_SYM(GetUniqueId, "ncclGetUniqueId");
    IVF_SYM(CommInitRank, "ncclCommInitRank");
    IVF_SYM(CommInitAll, "ncclCommInitAll");
    IVF_SYM(CommDestroy, "ncclCommDestroy");
    IVF_SYM(AllGather, "ncclAllGather");
    IVF_SYM(Send, "ncclSend");
    IVF_SYM(Recv, "ncclRecv");
    IVF_SYM(GroupStart, "ncclGroupStart");
    IVF_SYM(GroupEnd, "ncclGroupEnd");
    IVF_SYM(GetErrorString, "ncclGetErrorString");
#undef IVF_SYM
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.AllGather &&
             api.GroupStart && api.GroupEnd && api.GetErrorString;
    return api.ok ? &api : nullptr;
}

struct GraphKey {
    const void* dQ = nullptr;
    int64_t nq = 0;
    int k = 0, w = 0;
    const void *ids = nullptr, *d = nullptr, *cnt = nullptr;
    bool gather_q = false;
    bool operator==(const GraphKey& o) const {
        return dQ == o.dQ && nq == o.nq && k == o.k && w == o.w && ids == o.ids && d == o.d && cnt == o.cnt &&
               gather_q == o.gather_q;
    }
};

struct ShardCtx {
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
    bool use_graph = true;
    // gathered buffers: queries [nqp][D], probes [nqp][w], candidates [world][nq][k]
    DevBuf g_q, g_cells, g_dc, g_ids, g_d, g_keys, loc_cnt, out_ids, out_d, out_cnt;
    // slice merge (world >= 4): candidates of this rank's query slice from every rank [world][ns][k], results [nqp][k]
    DevBuf a_ids, a_d, a_keys, r_ids, r_d, r_cnt;
    // CUDA graph of the last shape (one entry: a serving loop repeats one shape)
    GraphKey key;
    int seen = 0;                 // eager runs with this key so far
    cudaGraphExec_t exec = nullptr;
    uint64_t graph_launches = 0;  // kernels of this library inside one replay
    // timing of the eager path: coarse slice, the two collectives
    cudaEvent_t ev[7] = {};
    bool ev_ok = false, ev_pending = false, ev_slices = false;
    double coarse_ms = 0, comm_ms = 0;
};

ShardCtx* ctx(ivfadc_index* h) { return static_cast<ShardCtx*>(h->shard_ctx); }

#define CUDA_OR_FAIL(h, call, what)                                                                     \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            cudaGetLastError();                                                                         \
            return api_fail(h, _e == cudaErrorMemoryAllocation ? IVFADC_ERR_OOM : IVFADC_ERR_CUDA, what, _e); \
        }                                                                                               \
    } while (0)
#define NCCL_OR_FAIL(h, call, what)                                                  \
    do {                                                                             \
        ncclResult_t _r = (call);                                                    \
        if (_r != ncclSuccess) {                                                     \
            std::string _m = std::string(what) + ": " + nccl_api()->GetErrorString(_r); \
            return api_fail(h, IVFADC_ERR_NCCL, _m.c_str());                         \
        }                                                                            \
    } while (0)

struct Step {
    int64_t nq, qs, nqp, lo, hi;  // queries, slice size, padded total, this rank's slice [lo, hi)
    int k, w;
};

Step make_step(const ivfadc_index* h, const ShardCtx* c, int64_t nq, int k, int w) {
    Step st;
    st.nq = nq;
    st.k = k;
    st.w = std::min(w, h->cfg.kc);  // reference src/index.jl:216
    st.qs = (nq + c->world - 1) / c->world;
    st.nqp = st.qs * c->world;
    st.lo = std::min(nq, (int64_t)c->rank * st.qs);
    st.hi = std::min(nq, st.lo + st.qs);
    return st;
}

// From four ranks on, every rank merges only ITS slice of the queries: the candidates travel in an all-to-all of
// [slice][k] blocks (each rank receives world x nq / world x k entries instead of world x nq x k), the merge kernel
// runs on nq / world queries, and one small all-gather distributes the finished rows.  Same kernels, same order of
// the parts (rank-major), so the result is bit-identical to the all-gather path (and to one GPU).
bool slice_merge(const ShardCtx* c) {
    NcclApi* n = nccl_api();
    return c->world >= 4 && n && n->Send && n->Recv;
}

int reserve_step(ivfadc_index* h, ShardCtx* c, const Step& st, bool gather_q) {
    const size_t T = h->tsize;
    if (gather_q) CUDA_OR_FAIL(h, c->g_q.reserve((size_t)st.nqp * h->cfg.dim * T), "workspace");
    CUDA_OR_FAIL(h, c->g_cells.reserve(sizeof(int32_t) * (size_t)st.nqp * st.w), "workspace");
    CUDA_OR_FAIL(h, c->g_dc.reserve(T * (size_t)st.nqp * st.w), "workspace");
    const size_t nk = (size_t)st.nq * st.k;
    CUDA_OR_FAIL(h, c->g_ids.reserve(sizeof(uint64_t) * nk * c->world), "workspace");
    CUDA_OR_FAIL(h, c->g_keys.reserve(sizeof(uint64_t) * nk * c->world), "workspace");
    CUDA_OR_FAIL(h, c->g_d.reserve(T * nk * c->world), "workspace");
    CUDA_OR_FAIL(h, c->loc_cnt.reserve(sizeof(int32_t) * (size_t)st.nq), "workspace");
    if (slice_merge(c)) {
        const size_t sk = (size_t)st.qs * st.k;
        CUDA_OR_FAIL(h, c->a_ids.reserve(sizeof(uint64_t) * sk * c->world), "workspace");
        CUDA_OR_FAIL(h, c->a_keys.reserve(sizeof(uint64_t) * sk * c->world), "workspace");
        CUDA_OR_FAIL(h, c->a_d.reserve(T * sk * c->world), "workspace");
        CUDA_OR_FAIL(h, c->r_ids.reserve(sizeof(uint64_t) * sk * c->world), "workspace");
        CUDA_OR_FAIL(h, c->r_d.reserve(T * sk * c->world), "workspace");
        CUDA_OR_FAIL(h, c->r_cnt.reserve(sizeof(int32_t) * (size_t)st.nqp), "workspace");
    }
    return IVFADC_OK;
}

// ---- the five phases of a step on one handle ----------------------------------------------------------
int phase_coarse(ivfadc_index* h, ShardCtx* c, const Step& st, const void* dQ, cudaStream_t s) {
    if (st.hi <= st.lo) return IVFADC_OK;
    int launches = 0;
    const char* q = static_cast<const char*>(dQ) + (size_t)st.lo * h->cfg.dim * h->tsize;
    CUDA_OR_FAIL(h, launch_coarse(h, q, st.hi - st.lo, st.w, c->g_cells.as<int32_t>() + st.lo * st.w,
                                  c->g_dc.as<char>() + (size_t)st.lo * st.w * h->tsize, s, &launches), "coarse kernel");
    h->stats.gpu_launches += launches;
    return IVFADC_OK;
}
// in-place all-gathers (the caller brackets them with ncclGroupStart / ncclGroupEnd)
int coll_probes(ivfadc_index* h, ShardCtx* c, const Step& st, bool gather_q, cudaStream_t s) {
    NcclApi* n = nccl_api();
    const size_t T = h->tsize, row = (size_t)c->rank * st.qs;
    if (gather_q) {
        const size_t b = (size_t)st.qs * h->cfg.dim * T;
        NCCL_OR_FAIL(h, n->AllGather(c->g_q.as<char>() + row * h->cfg.dim * T, c->g_q.p, b, ncclChar, c->comm, s), "all-gather (queries)");
    }
    const size_t bc = (size_t)st.qs * st.w * 4, bd = (size_t)st.qs * st.w * T;
    NCCL_OR_FAIL(h, n->AllGather(c->g_cells.as<char>() + row * st.w * 4, c->g_cells.p, bc, ncclChar, c->comm, s), "all-gather (probe cells)");
    NCCL_OR_FAIL(h, n->AllGather(c->g_dc.as<char>() + row * st.w * T, c->g_dc.p, bd, ncclChar, c->comm, s), "all-gather (probe distances)");
    return IVFADC_OK;
}
int phase_scan(ivfadc_index* h, ShardCtx* c, const Step& st, const void* dQ, cudaStream_t s) {
    const size_t nk = (size_t)st.nq * st.k;
    return api_search_core(h, dQ, st.nq, st.k, st.w, c->g_ids.as<uint64_t>() + nk * c->rank,
                           c->g_d.as<char>() + nk * c->rank * h->tsize, c->g_keys.as<uint64_t>() + nk * c->rank,
                           c->loc_cnt.as<int32_t>(), s, c->g_cells.as<int32_t>(), c->g_dc.p);
}
int coll_cands(ivfadc_index* h, ShardCtx* c, const Step& st, cudaStream_t s) {
    NcclApi* n = nccl_api();
    const size_t nk = (size_t)st.nq * st.k, T = h->tsize;
    NCCL_OR_FAIL(h, n->AllGather(c->g_ids.as<char>() + nk * c->rank * 8, c->g_ids.p, nk * 8, ncclChar, c->comm, s), "all-gather (ids)");
    NCCL_OR_FAIL(h, n->AllGather(c->g_d.as<char>() + nk * c->rank * T, c->g_d.p, nk * T, ncclChar, c->comm, s), "all-gather (distances)");
    NCCL_OR_FAIL(h, n->AllGather(c->g_keys.as<char>() + nk * c->rank * 8, c->g_keys.p, nk * 8, ncclChar, c->comm, s), "all-gather (keys)");
    return IVFADC_OK;
}
int phase_merge(ivfadc_index* h, ShardCtx* c, const Step& st, uint64_t* d_ids, void* d_dists, int32_t* d_counts,
                cudaStream_t s) {
    int launches = 0;
    CUDA_OR_FAIL(h, launch_merge_parts(h, c->world, st.nq, st.k, c->g_ids.as<uint64_t>(), c->g_d.p,
                                       c->g_keys.as<uint64_t>(), d_ids, d_dists, d_counts, s, &launches), "merge kernel");
    h->stats.gpu_launches += launches;
    return IVFADC_OK;
}

// slice merge: all-to-all of the candidate slices, merge of this rank's slice, all-gather of the finished rows
int coll_cands_slices(ivfadc_index* h, ShardCtx* c, const Step& st, cudaStream_t s) {
    NcclApi* n = nccl_api();
    const size_t nk = (size_t)st.nq * st.k, T = h->tsize, k = (size_t)st.k;
    const char* l_ids = c->g_ids.as<char>() + nk * c->rank * 8;      // this rank's candidates for all queries
    const char* l_d = c->g_d.as<char>() + nk * c->rank * T;
    const char* l_keys = c->g_keys.as<char>() + nk * c->rank * 8;
    const size_t ns = (size_t)(st.hi - st.lo);
    for (int r = 0; r < c->world; ++r) {
        const int64_t lo_r = std::min(st.nq, (int64_t)r * st.qs), hi_r = std::min(st.nq, lo_r + st.qs);
        const size_t ns_r = (size_t)(hi_r - lo_r);
        if (r == c->rank) continue;
        if (ns_r > 0) {
            NCCL_OR_FAIL(h, n->Send(l_ids + (size_t)lo_r * k * 8, ns_r * k * 8, ncclChar, r, c->comm, s), "send (ids)");
            NCCL_OR_FAIL(h, n->Send(l_d + (size_t)lo_r * k * T, ns_r * k * T, ncclChar, r, c->comm, s), "send (distances)");
            NCCL_OR_FAIL(h, n->Send(l_keys + (size_t)lo_r * k * 8, ns_r * k * 8, ncclChar, r, c->comm, s), "send (keys)");
        }
        if (ns > 0) {
            NCCL_OR_FAIL(h, n->Recv(c->a_ids.as<char>() + (size_t)r * ns * k * 8, ns * k * 8, ncclChar, r, c->comm, s), "recv (ids)");
            NCCL_OR_FAIL(h, n->Recv(c->a_d.as<char>() + (size_t)r * ns * k * T, ns * k * T, ncclChar, r, c->comm, s), "recv (distances)");
            NCCL_OR_FAIL(h, n->Recv(c->a_keys.as<char>() + (size_t)r * ns * k * 8, ns * k * 8, ncclChar, r, c->comm, s), "recv (keys)");
        }
    }
    return IVFADC_OK;
}
int phase_merge_slice(ivfadc_index* h, ShardCtx* c, const Step& st, cudaStream_t s) {
    const size_t T = h->tsize, k = (size_t)st.k, ns = (size_t)(st.hi - st.lo);
    if (ns == 0) return IVFADC_OK;
    const size_t nk = (size_t)st.nq * st.k;
    // own block of the slice: a device copy instead of a send to self
    CUDA_OR_FAIL(h, cudaMemcpyAsync(c->a_ids.as<char>() + (size_t)c->rank * ns * k * 8,
                                    c->g_ids.as<char>() + nk * c->rank * 8 + (size_t)st.lo * k * 8, ns * k * 8,
                                    cudaMemcpyDeviceToDevice, s), "D2D");
    CUDA_OR_FAIL(h, cudaMemcpyAsync(c->a_d.as<char>() + (size_t)c->rank * ns * k * T,
                                    c->g_d.as<char>() + nk * c->rank * T + (size_t)st.lo * k * T, ns * k * T,
                                    cudaMemcpyDeviceToDevice, s), "D2D");
    CUDA_OR_FAIL(h, cudaMemcpyAsync(c->a_keys.as<char>() + (size_t)c->rank * ns * k * 8,
                                    c->g_keys.as<char>() + nk * c->rank * 8 + (size_t)st.lo * k * 8, ns * k * 8,
                                    cudaMemcpyDeviceToDevice, s), "D2D");
    int launches = 0;
    CUDA_OR_FAIL(h, launch_merge_parts(h, c->world, (int64_t)ns, st.k, c->a_ids.as<uint64_t>(), c->a_d.p,
                                       c->a_keys.as<uint64_t>(), c->r_ids.as<uint64_t>() + (size_t)st.lo * k,
                                       c->r_d.as<char>() + (size_t)st.lo * k * T, c->r_cnt.as<int32_t>() + st.lo, s,
                                       &launches), "merge kernel");
    h->stats.gpu_launches += launches;
    return IVFADC_OK;
}
int coll_results(ivfadc_index* h, ShardCtx* c, const Step& st, cudaStream_t s) {
    NcclApi* n = nccl_api();
    const size_t T = h->tsize, row = (size_t)c->rank * st.qs, k = (size_t)st.k;
    NCCL_OR_FAIL(h, n->AllGather(c->r_ids.as<char>() + row * k * 8, c->r_ids.p, (size_t)st.qs * k * 8, ncclChar, c->comm, s), "all-gather (result ids)");
    NCCL_OR_FAIL(h, n->AllGather(c->r_d.as<char>() + row * k * T, c->r_d.p, (size_t)st.qs * k * T, ncclChar, c->comm, s), "all-gather (result distances)");
    NCCL_OR_FAIL(h, n->AllGather(c->r_cnt.as<char>() + row * 4, c->r_cnt.p, (size_t)st.qs * 4, ncclChar, c->comm, s), "all-gather (result counts)");
    return IVFADC_OK;
}

// one step of one rank (one process per GPU), asynchronous on `s`
int enqueue_step(ivfadc_index* h, ShardCtx* c, const Step& st, const void* dQ, bool gather_q, uint64_t* d_ids,
                 void* d_dists, int32_t* d_counts, cudaStream_t s, bool timed) {
    NcclApi* n = nccl_api();
    int rc;
    if (timed) cudaEventRecord(c->ev[0], s);
    if ((rc = phase_coarse(h, c, st, dQ, s)) != IVFADC_OK) return rc;
    if (timed) cudaEventRecord(c->ev[1], s);
    NCCL_OR_FAIL(h, n->GroupStart(), "ncclGroupStart");
    rc = coll_probes(h, c, st, gather_q, s);
    NCCL_OR_FAIL(h, n->GroupEnd(), "ncclGroupEnd");
    if (rc != IVFADC_OK) return rc;
    if (timed) cudaEventRecord(c->ev[2], s);
    if ((rc = phase_scan(h, c, st, dQ, s)) != IVFADC_OK) return rc;
    if (timed) cudaEventRecord(c->ev[3], s);
    if (slice_merge(c)) {
        NCCL_OR_FAIL(h, n->GroupStart(), "ncclGroupStart");
        rc = coll_cands_slices(h, c, st, s);
        NCCL_OR_FAIL(h, n->GroupEnd(), "ncclGroupEnd");
        if (rc != IVFADC_OK) return rc;
        if (timed) cudaEventRecord(c->ev[4], s);
        if ((rc = phase_merge_slice(h, c, st, s)) != IVFADC_OK) return rc;
        if (timed) cudaEventRecord(c->ev[5], s);
        NCCL_OR_FAIL(h, n->GroupStart(), "ncclGroupStart");
        rc = coll_results(h, c, st, s);
        NCCL_OR_FAIL(h, n->GroupEnd(), "ncclGroupEnd");
        if (rc != IVFADC_OK) return rc;
        const size_t nk = (size_t)st.nq * st.k;
        CUDA_OR_FAIL(h, cudaMemcpyAsync(d_ids, c->r_ids.p, sizeof(uint64_t) * nk, cudaMemcpyDeviceToDevice, s), "D2D");
        CUDA_OR_FAIL(h, cudaMemcpyAsync(d_dists, c->r_d.p, h->tsize * nk, cudaMemcpyDeviceToDevice, s), "D2D");
        CUDA_OR_FAIL(h, cudaMemcpyAsync(d_counts, c->r_cnt.p, sizeof(int32_t) * (size_t)st.nq, cudaMemcpyDeviceToDevice, s), "D2D");
        if (timed) {
            cudaEventRecord(c->ev[6], s);
            c->ev_pending = true;
            c->ev_slices = true;
        }
        return IVFADC_OK;
    }
    NCCL_OR_FAIL(h, n->GroupStart(), "ncclGroupStart");
    rc = coll_cands(h, c, st, s);
    NCCL_OR_FAIL(h, n->GroupEnd(), "ncclGroupEnd");
    if (rc != IVFADC_OK) return rc;
    if (timed) cudaEventRecord(c->ev[4], s);
    rc = phase_merge(h, c, st, d_ids, d_dists, d_counts, s);
    if (timed) {
        cudaEventRecord(c->ev[5], s);
        c->ev_pending = true;
        c->ev_slices = false;
    }
    return rc;
}

// Eager the first two times a shape is seen (workspaces grow, kernel attributes are set, per-kernel timing is
// recorded), then captured once and replayed.
int run_step(ivfadc_index* h, ShardCtx* c, const void* dQ, int64_t nq, int k, int w, bool gather_q, uint64_t* d_ids,
             void* d_dists, int32_t* d_counts, cudaStream_t s) {
    const Step st = make_step(h, c, nq, k, w);
    int rc = reserve_step(h, c, st, gather_q);
    if (rc != IVFADC_OK) return rc;
    const void* q = gather_q ? c->g_q.p : dQ;
    GraphKey key;
    key.dQ = q; key.nq = nq; key.k = k; key.w = st.w; key.ids = d_ids; key.d = d_dists; key.cnt = d_counts; key.gather_q = gather_q;
    if (!(key == c->key)) {
        if (c->exec) cudaGraphExecDestroy(c->exec);
        c->exec = nullptr;
        c->key = key;
        c->seen = 0;
    }
    const bool graph_ok = c->use_graph && !h->d_dbg_lut;
    if (c->exec && graph_ok) {
        CUDA_OR_FAIL(h, cudaGraphLaunch(c->exec, s), "graph launch");
        h->stats.searches += 1;
        h->stats.queries += nq;
        h->stats.gpu_launches += c->graph_launches;
        return IVFADC_OK;
    }
    if (!graph_ok || c->seen < 2) {
        shard_flush_timing(h);
        ++c->seen;
        return enqueue_step(h, c, st, q, gather_q, d_ids, d_dists, d_counts, s, h->stats_timing && c->ev_ok);
    }
    // capture
    const bool timing = h->stats_timing;
    h->stats_timing = false;
    const uint64_t launches0 = h->stats.gpu_launches;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) {
        h->stats_timing = timing;
        cudaGetLastError();
        return enqueue_step(h, c, st, q, gather_q, d_ids, d_dists, d_counts, s, false);
    }
    rc = enqueue_step(h, c, st, q, gather_q, d_ids, d_dists, d_counts, s, false);
    e = cudaStreamEndCapture(s, &graph);
    h->stats_timing = timing;
    if (rc != IVFADC_OK) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
    }
    CUDA_OR_FAIL(h, e, "graph capture");
    e = cudaGraphInstantiate(&c->exec, graph, 0);
    cudaGraphDestroy(graph);
    CUDA_OR_FAIL(h, e, "graph instantiate");
    c->graph_launches = h->stats.gpu_launches - launches0;
    CUDA_OR_FAIL(h, cudaGraphLaunch(c->exec, s), "graph launch");
    return IVFADC_OK;
}

int new_ctx(ivfadc_index* h, ncclComm_t comm, int world, int rank) {
    ShardCtx* c = new (std::nothrow) ShardCtx();
    if (!c) return api_fail(h, IVFADC_ERR_OOM, "out of host memory");
    c->comm = comm;
    c->world = world;
    c->rank = rank;
    c->ev_ok = true;
    for (auto& ev : c->ev)
        if (cudaEventCreate(&ev) != cudaSuccess) c->ev_ok = false;
    h->shard_ctx = c;
    return IVFADC_OK;
}

}  // namespace

namespace ivf {

void shard_flush_timing(ivfadc_index* h) {
    ShardCtx* c = ctx(h);
    if (!c || !c->ev_pending) return;
    c->ev_pending = false;
    if (cudaEventSynchronize(c->ev[c->ev_slices ? 6 : 5]) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]) == cudaSuccess) h->stats.coarse_ms += ms;
    if (cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]) == cudaSuccess) h->stats.comm_ms += ms;
    if (cudaEventElapsedTime(&ms, c->ev[3], c->ev[4]) == cudaSuccess) h->stats.comm_ms += ms;
    if (cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]) == cudaSuccess) h->stats.merge_ms += ms;
    if (c->ev_slices && cudaEventElapsedTime(&ms, c->ev[5], c->ev[6]) == cudaSuccess) h->stats.comm_ms += ms;
    cudaGetLastError();
}

void shard_destroy_ctx(ivfadc_index* h) {
    ShardCtx* c = ctx(h);
    if (!c) return;
    if (c->exec) cudaGraphExecDestroy(c->exec);
    for (auto& ev : c->ev)
        if (ev) cudaEventDestroy(ev);
    DevBuf* bufs[] = {&c->g_q, &c->g_cells, &c->g_dc, &c->g_ids, &c->g_d, &c->g_keys, &c->loc_cnt, &c->out_ids, &c->out_d, &c->out_cnt,
                      &c->a_ids, &c->a_d, &c->a_keys, &c->r_ids, &c->r_d, &c->r_cnt};
    for (DevBuf* b : bufs) b->release();
    if (c->comm && nccl_api()) nccl_api()->CommDestroy(c->comm);
    delete c;
    h->shard_ctx = nullptr;
    cudaGetLastError();
}

}  // namespace ivf

// The group of a single-process multi-GPU index.
struct ivfadc_group {
    std::vector<ivfadc_index*> hs;
    std::string err;
};

extern "C" {

int ivfadc_nccl_unique_id(void* id_out) {
    if (!id_out) return IVFADC_ERR_BAD_ARG;
    NcclApi* n = nccl_api();
    if (!n) return IVFADC_ERR_NCCL;
    ncclUniqueId id;
    if (n->GetUniqueId(&id) != ncclSuccess) return IVFADC_ERR_NCCL;
    static_assert(sizeof(ncclUniqueId) == IVFADC_NCCL_ID_BYTES, "ncclUniqueId size");
    memcpy(id_out, &id, sizeof(id));
    return IVFADC_OK;
}

int ivfadc_comm_init_rank(ivfadc_index* h, const void* id, int32_t world, int32_t rank) {
    IVF_NVTX();
    if (!h || !id) return IVFADC_ERR_BAD_ARG;
    if (world != h->cfg.shard_world || rank != h->cfg.shard_rank)
        return api_fail(h, IVFADC_ERR_BAD_ARG, "communicator shape differs from the handle's shard_rank / shard_world");
    if (h->shard_ctx) return api_fail(h, IVFADC_ERR_BAD_ARG, "the handle already has a communicator");
    NcclApi* n = nccl_api();
    if (!n) return api_fail(h, IVFADC_ERR_NCCL, "libnccl.so.2 could not be loaded");
    cudaSetDevice(h->cfg.device);
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm = nullptr;
    NCCL_OR_FAIL(h, n->CommInitRank(&comm, world, uid, rank), "ncclCommInitRank");
    return new_ctx(h, comm, world, rank);
}

int ivfadc_comm_destroy(ivfadc_index* h) {
    if (!h) return IVFADC_ERR_BAD_ARG;
    cudaSetDevice(h->cfg.device);
    cudaStreamSynchronize(h->stream);
    shard_destroy_ctx(h);
    return IVFADC_OK;
}

int ivfadc_set_graph_replay(ivfadc_index* h, int32_t enable) {
    if (!h || !h->shard_ctx) return IVFADC_ERR_BAD_ARG;
    ctx(h)->use_graph = enable != 0;
    return IVFADC_OK;
}

int ivfadc_search_sharded_device(ivfadc_index* h, const void* dQ, int64_t nq, int32_t k, int32_t w, uint64_t* d_ids,
                                 void* d_dists, int32_t* d_counts, void* stream) {
    IVF_NVTX();
    if (!h) return IVFADC_ERR_BAD_ARG;
    ShardCtx* c = ctx(h);
    if (!c) return api_fail(h, IVFADC_ERR_BAD_ARG, "no communicator: call ivfadc_comm_init_rank first");
    if (!dQ || !d_ids || !d_dists || !d_counts) return api_fail(h, IVFADC_ERR_BAD_ARG, "null pointer");
    if (k < 1) return api_fail(h, IVFADC_ERR_BAD_ARG, "Number of neighbors must be k >= 1");
    if (w < 1) return api_fail(h, IVFADC_ERR_BAD_ARG, "Number of clusters to search in must be w >= 1");
    if (nq <= 0) return nq == 0 ? IVFADC_OK : api_fail(h, IVFADC_ERR_BAD_ARG, "nq < 0");
    cudaSetDevice(h->cfg.device);
    return run_step(h, c, dQ, nq, k, w, false, d_ids, d_dists, d_counts, static_cast<cudaStream_t>(stream));
}

int ivfadc_search_sharded(ivfadc_index* h, const void* Q, int64_t nq, int32_t k, int32_t w, uint64_t* ids_out,
                          void* dists_out, int32_t* counts_out) {
    IVF_NVTX();
    if (!h) return IVFADC_ERR_BAD_ARG;
    ShardCtx* c = ctx(h);
    if (!c) return api_fail(h, IVFADC_ERR_BAD_ARG, "no communicator: call ivfadc_comm_init_rank first");
    if (nq > 0 && (!Q || !ids_out || !dists_out || !counts_out)) return api_fail(h, IVFADC_ERR_BAD_ARG, "null pointer");
    if (k < 1) return api_fail(h, IVFADC_ERR_BAD_ARG, "Number of neighbors must be k >= 1");
    if (w < 1) return api_fail(h, IVFADC_ERR_BAD_ARG, "Number of clusters to search in must be w >= 1");
    if (nq <= 0) return nq == 0 ? IVFADC_OK : api_fail(h, IVFADC_ERR_BAD_ARG, "nq < 0");
    cudaSetDevice(h->cfg.device);
    cudaStream_t s = h->stream;
    const Step st = make_step(h, c, nq, k, w);
    const size_t T = h->tsize, rowb = (size_t)h->cfg.dim * T, nk = (size_t)nq * k;
    int rc = reserve_step(h, c, st, true);
    if (rc != IVFADC_OK) return rc;
    CUDA_OR_FAIL(h, c->out_ids.reserve(sizeof(uint64_t) * nk), "workspace");
    CUDA_OR_FAIL(h, c->out_d.reserve(T * nk), "workspace");
    CUDA_OR_FAIL(h, c->out_cnt.reserve(sizeof(int32_t) * (size_t)nq), "workspace");
    // this rank uploads ITS slice of the batch only; the all-gather over NVLink completes it
    if (st.hi > st.lo)
        CUDA_OR_FAIL(h, cudaMemcpyAsync(c->g_q.as<char>() + st.lo * rowb, static_cast<const char*>(Q) + st.lo * rowb,
                                        (size_t)(st.hi - st.lo) * rowb, cudaMemcpyHostToDevice, s), "H2D");
    rc = run_step(h, c, nullptr, nq, k, w, true, c->out_ids.as<uint64_t>(), c->out_d.p, c->out_cnt.as<int32_t>(), s);
    if (rc != IVFADC_OK) return rc;
    CUDA_OR_FAIL(h, cudaMemcpyAsync(ids_out, c->out_ids.p, sizeof(uint64_t) * nk, cudaMemcpyDeviceToHost, s), "D2H");
    CUDA_OR_FAIL(h, cudaMemcpyAsync(dists_out, c->out_d.p, T * nk, cudaMemcpyDeviceToHost, s), "D2H");
    CUDA_OR_FAIL(h, cudaMemcpyAsync(counts_out, c->out_cnt.p, sizeof(int32_t) * (size_t)nq, cudaMemcpyDeviceToHost, s), "D2H");
    return ivfadc_check_async(h, s);
}

int ivfadc_sharded_step_bytes(ivfadc_index* h, int64_t nq, int32_t k, int32_t w, int64_t* h2d_out, int64_t* d2h_out,
                              int64_t* nvlink_out) {
    if (!h || !h->shard_ctx) return IVFADC_ERR_BAD_ARG;
    ShardCtx* c = ctx(h);
    const Step st = make_step(h, c, nq, k, w);
    const int64_t T = (int64_t)h->tsize;
    if (h2d_out) *h2d_out = (st.hi - st.lo) * h->cfg.dim * T;
    if (d2h_out) *d2h_out = nq * k * (8 + T) + nq * 4;
    // bytes this rank RECEIVES over NVLink per step: the other ranks' query slices, probes and candidates
    if (nvlink_out) {
        const int64_t probes = (int64_t)(c->world - 1) * st.qs * (h->cfg.dim * T + st.w * (4 + T));
        const int64_t ns = st.hi - st.lo;
        *nvlink_out = probes + (slice_merge(c)
                                    ? (int64_t)(c->world - 1) * (ns * k * (16 + T) + st.qs * (k * (8 + T) + 4))   // candidate slices + finished rows
                                    : (int64_t)(c->world - 1) * nq * k * (16 + T));                                // all candidates
    }
    return IVFADC_OK;
}

// ---- single process, several GPUs ------------------------------------------------------------------------
int ivfadc_group_create(ivfadc_group** out, const ivfadc_config* cfg, int32_t n_devices, const int32_t* device_ids,
                        const void* centroids, const void* codebook_vectors, const uint8_t* codebook_codes) {
    IVF_NVTX();
    if (!out || !cfg || n_devices < 1 || n_devices > 64) return IVFADC_ERR_BAD_ARG;
    *out = nullptr;
    ivfadc_group* g = new (std::nothrow) ivfadc_group();
    if (!g) return IVFADC_ERR_OOM;
    std::vector<int> devs(n_devices);
    int rc = IVFADC_OK;
    for (int i = 0; i < n_devices && rc == IVFADC_OK; ++i) {
        ivfadc_config ci = *cfg;
        ci.device = devs[i] = device_ids ? device_ids[i] : i;
        ci.shard_rank = i;
        ci.shard_world = n_devices;
        ivfadc_index* h = nullptr;
        rc = ivfadc_create(&h, &ci, centroids, codebook_vectors, codebook_codes);
        if (rc == IVFADC_OK) g->hs.push_back(h);
    }
    NcclApi* n = rc == IVFADC_OK ? nccl_api() : nullptr;
    if (rc == IVFADC_OK && !n) rc = IVFADC_ERR_NCCL;
    if (rc == IVFADC_OK) {
        std::vector<ncclComm_t> comms(n_devices, nullptr);
        if (n->CommInitAll(comms.data(), n_devices, devs.data()) != ncclSuccess) rc = IVFADC_ERR_NCCL;
        for (int i = 0; i < n_devices && rc == IVFADC_OK; ++i) {
            cudaSetDevice(devs[i]);
            rc = new_ctx(g->hs[i], comms[i], n_devices, i);
            if (rc == IVFADC_OK) ctx(g->hs[i])->use_graph = false;  // one host thread drives every device: eager
        }
    }
    if (rc != IVFADC_OK) {
        for (ivfadc_index* h : g->hs) ivfadc_destroy(h);
        delete g;
        return rc;
    }
    *out = g;
    return IVFADC_OK;
}

int ivfadc_group_destroy(ivfadc_group* g) {
    if (!g) return IVFADC_OK;
    for (ivfadc_index* h : g->hs) ivfadc_destroy(h);
    delete g;
    return IVFADC_OK;
}

int ivfadc_group_size(const ivfadc_group* g) { return g ? (int)g->hs.size() : IVFADC_ERR_BAD_ARG; }

int ivfadc_group_handle(ivfadc_group* g, int32_t i, ivfadc_index** out) {
    if (!g || !out || i < 0 || i >= (int)g->hs.size()) return IVFADC_ERR_BAD_ARG;
    *out = g->hs[i];
    return IVFADC_OK;
}

const char* ivfadc_group_last_error(const ivfadc_group* g) {
    if (!g) return "null group";
    for (const ivfadc_index* h : g->hs)
        if (!h->err.empty()) return h->err.c_str();
    return "";
}

int ivfadc_group_set_cell_owners(ivfadc_group* g, const int32_t* owners) {
    if (!g) return IVFADC_ERR_BAD_ARG;
    for (ivfadc_index* h : g->hs) {
        int rc = ivfadc_set_cell_owners(h, owners);
        if (rc != IVFADC_OK) return rc;
    }
    return IVFADC_OK;
}

int ivfadc_group_add(ivfadc_group* g, const void* X, int64_t n, int32_t position, const int64_t* assign,
                     int32_t assign_base, int32_t* cells_out) {
    IVF_NVTX();
    if (!g) return IVFADC_ERR_BAD_ARG;
    // every shard sees the whole batch and keeps the cells it owns (ids stay global); a rejected batch
    // (capacity) is rejected by the first shard before any shard has changed
    for (size_t i = 0; i < g->hs.size(); ++i) {
        int rc = ivfadc_add(g->hs[i], X, n, position, assign, assign_base, i == 0 ? cells_out : nullptr);
        if (rc != IVFADC_OK) return rc;
    }
    return IVFADC_OK;
}

int ivfadc_group_delete(ivfadc_group* g, const uint64_t* ids, int64_t n) {
    IVF_NVTX();
    if (!g) return IVFADC_ERR_BAD_ARG;
    for (ivfadc_index* h : g->hs) {
        int rc = ivfadc_delete(h, ids, n);
        if (rc != IVFADC_OK) return rc;
    }
    return IVFADC_OK;
}

int ivfadc_group_pop(ivfadc_group* g, int32_t position, void* vec_out) {
    IVF_NVTX();
    if (!g || !vec_out) return IVFADC_ERR_BAD_ARG;
    // every shard renumbers; the one that owned the vector returns it
    std::vector<char> tmp(g->hs[0]->cfg.dim * g->hs[0]->tsize);
    bool got = false;
    for (ivfadc_index* h : g->hs) {
        int32_t found = 0;
        int rc = ivfadc_pop(h, position, tmp.data(), &found);
        if (rc != IVFADC_OK) return rc;
        if (found) {
            memcpy(vec_out, tmp.data(), tmp.size());
            got = true;
        }
    }
    return got ? IVFADC_OK : api_fail(g->hs[0], IVFADC_ERR_CUDA, "internal: no shard owned the popped vector");
}

int ivfadc_group_length(const ivfadc_group* g, int64_t* n_out) {
    if (!g || !n_out) return IVFADC_ERR_BAD_ARG;
    return ivfadc_length(g->hs[0], n_out);
}

int ivfadc_group_search(ivfadc_group* g, const void* Q, int64_t nq, int32_t k, int32_t w, uint64_t* ids_out,
                        void* dists_out, int32_t* counts_out) {
    IVF_NVTX();
    if (!g) return IVFADC_ERR_BAD_ARG;
    ivfadc_index* h0 = g->hs[0];
    if (nq > 0 && (!Q || !ids_out || !dists_out || !counts_out)) return api_fail(h0, IVFADC_ERR_BAD_ARG, "null pointer");
    if (k < 1) return api_fail(h0, IVFADC_ERR_BAD_ARG, "Number of neighbors must be k >= 1");
    if (w < 1) return api_fail(h0, IVFADC_ERR_BAD_ARG, "Number of clusters to search in must be w >= 1");
    if (nq <= 0) return nq == 0 ? IVFADC_OK : api_fail(h0, IVFADC_ERR_BAD_ARG, "nq < 0");
    NcclApi* n = nccl_api();
    const int G = (int)g->hs.size();
    std::vector<Step> st(G);
    const size_t T = h0->tsize, rowb = (size_t)h0->cfg.dim * T, nk = (size_t)nq * k;
    int rc;
    for (int i = 0; i < G; ++i) {  // slice upload + coarse step of the slice, device by device (asynchronous)
        ivfadc_index* h = g->hs[i];
        ShardCtx* c = ctx(h);
        cudaSetDevice(h->cfg.device);
        st[i] = make_step(h, c, nq, k, w);
        if ((rc = reserve_step(h, c, st[i], true)) != IVFADC_OK) return rc;
        if (i == 0) {
            CUDA_OR_FAIL(h, c->out_ids.reserve(sizeof(uint64_t) * nk), "workspace");
            CUDA_OR_FAIL(h, c->out_d.reserve(T * nk), "workspace");
            CUDA_OR_FAIL(h, c->out_cnt.reserve(sizeof(int32_t) * (size_t)nq), "workspace");
        }
        if (st[i].hi > st[i].lo)
            CUDA_OR_FAIL(h, cudaMemcpyAsync(c->g_q.as<char>() + st[i].lo * rowb, static_cast<const char*>(Q) + st[i].lo * rowb,
                                            (size_t)(st[i].hi - st[i].lo) * rowb, cudaMemcpyHostToDevice, h->stream), "H2D");
        if ((rc = phase_coarse(h, c, st[i], c->g_q.p, h->stream)) != IVFADC_OK) return rc;
    }
    NCCL_OR_FAIL(h0, n->GroupStart(), "ncclGroupStart");
    for (int i = 0; i < G; ++i) {
        cudaSetDevice(g->hs[i]->cfg.device);
        if ((rc = coll_probes(g->hs[i], ctx(g->hs[i]), st[i], true, g->hs[i]->stream)) != IVFADC_OK) break;
    }
    NCCL_OR_FAIL(h0, n->GroupEnd(), "ncclGroupEnd");
    if (rc != IVFADC_OK) return rc;
    for (int i = 0; i < G; ++i) {
        cudaSetDevice(g->hs[i]->cfg.device);
        if ((rc = phase_scan(g->hs[i], ctx(g->hs[i]), st[i], ctx(g->hs[i])->g_q.p, g->hs[i]->stream)) != IVFADC_OK) return rc;
    }
    NCCL_OR_FAIL(h0, n->GroupStart(), "ncclGroupStart");
    for (int i = 0; i < G; ++i) {
        cudaSetDevice(g->hs[i]->cfg.device);
        if ((rc = coll_cands(g->hs[i], ctx(g->hs[i]), st[i], g->hs[i]->stream)) != IVFADC_OK) break;
    }
    NCCL_OR_FAIL(h0, n->GroupEnd(), "ncclGroupEnd");
    if (rc != IVFADC_OK) return rc;
    // the merged result is needed once: device 0 merges and returns it
    cudaSetDevice(h0->cfg.device);
    ShardCtx* c0 = ctx(h0);
    if ((rc = phase_merge(h0, c0, st[0], c0->out_ids.as<uint64_t>(), c0->out_d.p, c0->out_cnt.as<int32_t>(), h0->stream)) != IVFADC_OK) return rc;
    CUDA_OR_FAIL(h0, cudaMemcpyAsync(ids_out, c0->out_ids.p, sizeof(uint64_t) * nk, cudaMemcpyDeviceToHost, h0->stream), "D2H");
    CUDA_OR_FAIL(h0, cudaMemcpyAsync(dists_out, c0->out_d.p, T * nk, cudaMemcpyDeviceToHost, h0->stream), "D2H");
    CUDA_OR_FAIL(h0, cudaMemcpyAsync(counts_out, c0->out_cnt.p, sizeof(int32_t) * (size_t)nq, cudaMemcpyDeviceToHost, h0->stream), "D2H");
    for (int i = 0; i < G; ++i) {
        h0 = g->hs[i];
        h0->stats.searches += 1;
        if ((rc = ivfadc_check_async(h0, h0->stream)) != IVFADC_OK) return rc;
    }
    return IVFADC_OK;
}

}  // extern "C"
