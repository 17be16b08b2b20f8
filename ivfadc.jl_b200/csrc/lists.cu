// K5 -- the inverted lists: a device-resident CSR with per-list slack, and every mutation the
// reference performs on its Vector{InvertedList} (src/index.jl:8-23, src/utils.jl):
//   append at the list tail       push!/pushfirst!            src/utils.jl:142-143
//   add a constant to every id    _shift_up_inverse_index!    src/utils.jl:2-6
//   delete + renumber + compact   delete_from_index!          src/utils.jl:90-105, 16-20
//   find an id, decode a vector   _pop!, _decode_point        src/utils.jl:41-81
//   list export / import          persistency                 src/persistency.jl:68-78,119-131
//
// Layout: list c = entries [off[c], off[c] + len[c]) of two arenas, codes uint8[.][m] (row per
// vector, so m = 16 is one 128-bit load) and ids uint32|uint64[.]; off[c] and cap[c] are
// multiples of 16 entries so every list starts 16-byte aligned for any m.
#include <algorithm>
#include <cub/cub.cuh>

#include "common.cuh"

namespace ivf {

namespace {

inline int64_t round16(int64_t x) { return (x + 15) & ~(int64_t)15; }

template <typename IdT>
__global__ void copy_lists_kernel(const uint8_t* __restrict__ src_codes, const IdT* __restrict__ src_ids,
                                  const int64_t* __restrict__ src_off, const int64_t* __restrict__ len,
                                  uint8_t* __restrict__ dst_codes, IdT* __restrict__ dst_ids,
                                  const int64_t* __restrict__ dst_off, int m) {
    const int c = blockIdx.x;
    const int64_t n = len[c];
    if (n <= 0) return;
    const int64_t so = src_off[c], dof = dst_off[c];
    // both bases are 16-byte aligned and capacities are multiples of 16 entries
    const uint4* s4 = reinterpret_cast<const uint4*>(src_codes + (size_t)so * m);
    uint4* d4 = reinterpret_cast<uint4*>(dst_codes + (size_t)dof * m);
    const int64_t n16 = (n * m + 15) / 16;
    for (int64_t i = threadIdx.x; i < n16; i += blockDim.x) d4[i] = s4[i];
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) dst_ids[dof + i] = src_ids[so + i];
}

// histogram of the (sorted) cells of a batch + first sorted position of each cell's run
__global__ void run_bounds_kernel(const int32_t* __restrict__ sorted_cells, int64_t n, int kc,
                                  int64_t* run_start, int64_t* run_cnt) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int c = sorted_cells[t];
    if (c < 0 || c >= kc) return;
    if (t == 0 || sorted_cells[t - 1] != c) run_start[c] = t;
    if (t == n - 1 || sorted_cells[t + 1] != c) run_cnt[c] = t + 1;  // end (exclusive); fixed up on host
}

template <typename IdT>
__global__ void scatter_append_kernel(const int32_t* __restrict__ sorted_cells,
                                      const int32_t* __restrict__ perm, int64_t n, int kc,
                                      const int64_t* __restrict__ run_start,
                                      const int64_t* __restrict__ list_off,
                                      const int64_t* __restrict__ list_len,  // lengths BEFORE the append
                                      const uint8_t* __restrict__ codes_in, int m, uint8_t* codes,
                                      IdT* ids, uint64_t first_id, int step, int shard_rank,
                                      int shard_world, const int32_t* __restrict__ owner) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int c = sorted_cells[t];
    if (c < 0 || c >= kc) return;
    if (shard_world > 1 && (owner ? owner[c] : c % shard_world) != shard_rank) return;
    const int64_t j = perm[t];
    const int64_t dst = list_off[c] + list_len[c] + (t - run_start[c]);
    const uint8_t* s = codes_in + (size_t)j * m;
    uint8_t* d = codes + (size_t)dst * m;
    for (int i = 0; i < m; ++i) d[i] = s[i];
    ids[dst] = (IdT)(step > 0 ? first_id + (uint64_t)j : first_id - (uint64_t)j);
}

__global__ void iota_kernel(int32_t* p, int64_t n) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) p[t] = (int32_t)t;
}

template <typename IdT>
__global__ void shift_ids_kernel(IdT* ids, const int64_t* __restrict__ off, const int64_t* __restrict__ len,
                                 int64_t by) {
    const int c = blockIdx.x;
    const int64_t n = len[c], o = off[c];
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) ids[o + i] = (IdT)((int64_t)ids[o + i] + by);
}

__device__ __forceinline__ int64_t lower_bound_u64(const uint64_t* a, int64_t n, uint64_t x) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// delete_from_index!: one CTA per list; in-place, order-preserving compaction.  The net effect of
// the reference's descending loop of deleteat! + "every id > point -= 1" (src/utils.jl:94-101) is
// new_id = old_id - |{deleted ids < old_id}| for every survivor, survivors keeping list order.
constexpr int DTHREADS = 256;
template <typename IdT>
__global__ void __launch_bounds__(DTHREADS)
delete_compact_kernel(uint8_t* codes, IdT* ids, const int64_t* __restrict__ off, int64_t* len, int m,
                      const uint64_t* __restrict__ del, int64_t ndel) {
    extern __shared__ __align__(16) unsigned char s_stage[];  // [DTHREADS][m] code bytes of the chunk
    __shared__ int s_wsum[DTHREADS / 32];
    __shared__ int64_t s_write;
    const int c = blockIdx.x;
    const int64_t n = len[c], o = off[c];
    if (n <= 0) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_write = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += DTHREADS) {
        const int64_t i = base + tid;
        const bool valid = i < n;
        uint64_t id = 0;
        int64_t lb = 0;
        bool keep = false;
        if (valid) {
            id = (uint64_t)ids[o + i];
            lb = lower_bound_u64(del, ndel, id);
            keep = !(lb < ndel && del[lb] == id);
        }
        // stage the chunk's codes (coalesced byte copy)
        const int64_t chunk = min((int64_t)DTHREADS, n - base);
        for (int64_t x = tid; x < chunk * m; x += DTHREADS) s_stage[x] = codes[(size_t)(o + base) * m + x];
        // block-wide exclusive scan of keep flags
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_wsum[wid] = __popc(bal);
        __syncthreads();
        int before = __popc(bal & ((1u << lane) - 1u));
        int total = 0;
        for (int x = 0; x < DTHREADS / 32; ++x) {
            if (x < wid) before += s_wsum[x];
            total += s_wsum[x];
        }
        const int64_t wbase = s_write;
        if (keep) {
            const int64_t dst = o + wbase + before;
            ids[dst] = (IdT)(id - (uint64_t)lb);
            for (int x = 0; x < m; ++x) codes[(size_t)dst * m + x] = s_stage[(size_t)tid * m + x];
        }
        __syncthreads();
        if (tid == 0) s_write = wbase + total;
        __syncthreads();
    }
    if (tid == 0) len[c] = s_write;
}

template <typename IdT>
__global__ void find_id_kernel(const IdT* __restrict__ ids, const int64_t* __restrict__ off,
                               const int64_t* __restrict__ len, uint64_t id, int64_t* result) {
    const int c = blockIdx.x;
    const int64_t n = len[c], o = off[c];
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x)
        if ((uint64_t)ids[o + i] == id) {
            result[0] = c;
            result[1] = i;
        }
}

// centroid + concat_i codebook_i[code_i]   (reference src/utils.jl:58-59, 71-81)
template <typename T>
__global__ void decode_kernel(const T* __restrict__ C, const T* __restrict__ cb,
                              const uint8_t* __restrict__ cb_codes, const uint8_t* __restrict__ code,
                              int cell, int D, int m, int dsub, int ksub, T* out) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    T v = C[(size_t)cell * D + d];
    const int i = d / dsub;
    if (i < m) {
        int col = -1;
        for (int cw = 0; cw < ksub; ++cw)
            if (cb_codes[i * ksub + cw] == code[i]) { col = cw; break; }
        if (col >= 0) v = add_rn(v, cb[((size_t)i * ksub + col) * dsub + (d - i * dsub)]);
    }
    out[d] = v;
}

__global__ void widen_ids_kernel(const uint32_t* __restrict__ in, uint64_t* __restrict__ out, int64_t n) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[t] = in[t];
}
__global__ void narrow_ids_kernel(const uint64_t* __restrict__ in, uint32_t* __restrict__ out, int64_t n) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[t] = (uint32_t)in[t];
}

// Bulk persistency: packed (cell-ascending, list order) <-> arena.  Packed entry p belongs to the cell with
// pref[cell] <= p < pref[cell + 1] (binary search over the kc + 1 prefix sums); its arena slot is off[cell] + p - pref[cell].
template <typename I, bool EXPORT>
__global__ void pack_lists_kernel(uint8_t* __restrict__ arena_codes, I* __restrict__ arena_ids,
                                  const int64_t* __restrict__ off, const int64_t* __restrict__ pref, int kc, int m,
                                  int64_t p0, int64_t n, uint8_t* __restrict__ pk_codes, uint64_t* __restrict__ pk_ids) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int64_t p = p0 + t;
    int lo = 0, hi = kc;  // largest c with pref[c] <= p
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (pref[mid] <= p) lo = mid; else hi = mid;
    }
    const int64_t a = off[lo] + (p - pref[lo]);
    if (EXPORT) {
        pk_ids[t] = (uint64_t)arena_ids[a];
        for (int j = 0; j < m; ++j) pk_codes[t * m + j] = arena_codes[a * m + j];
    } else {
        arena_ids[a] = (I)pk_ids[t];
        for (int j = 0; j < m; ++j) arena_codes[a * m + j] = pk_codes[t * m + j];
    }
}

#define CK(x)                                  \
    do {                                       \
        cudaError_t _e = (x);                  \
        if (_e != cudaSuccess) return _e;      \
    } while (0)

// Make sure list c can hold need[c] entries; regrows (and re-packs) the arenas when any cannot.
cudaError_t reserve_lists(ivfadc_index* h, const std::vector<int64_t>& need, int* launches) {
    const int kc = h->cfg.kc, m = h->cfg.m;
    bool fits = true;
    for (int c = 0; c < kc; ++c)
        if (need[c] > h->h_cap[c]) { fits = false; break; }
    if (fits) return cudaSuccess;

    std::vector<int64_t> ncap(kc), noff(kc);
    int64_t total = 0;
    for (int c = 0; c < kc; ++c) {
        int64_t want = std::max(need[c], h->h_len[c]);
        // slack: 25% + 16 so that streams of single push! calls regrow rarely
        ncap[c] = round16(want > 0 ? want + want / 4 + 16 : 16);
        noff[c] = total;
        total += ncap[c];
    }
    uint8_t* ncodes = nullptr;
    void* nids = nullptr;
    int64_t* d_noff = nullptr;
    CK(cudaMalloc(&ncodes, (size_t)total * m + 16));
    cudaError_t e = cudaMalloc(&nids, (size_t)total * h->id_dev_bytes + 16);
    if (e != cudaSuccess) { cudaFree(ncodes); return e; }
    e = cudaMalloc(&d_noff, sizeof(int64_t) * kc);
    if (e != cudaSuccess) { cudaFree(ncodes); cudaFree(nids); return e; }
    CK(cudaMemcpyAsync(d_noff, noff.data(), sizeof(int64_t) * kc, cudaMemcpyHostToDevice, h->stream));
    if (h->n_local > 0) {
        if (h->id_dev_bytes == 4)
            copy_lists_kernel<uint32_t><<<kc, 256, 0, h->stream>>>(
                h->d_codes, static_cast<const uint32_t*>(h->d_ids), h->d_off, h->d_len, ncodes,
                static_cast<uint32_t*>(nids), d_noff, m);
        else
            copy_lists_kernel<uint64_t><<<kc, 256, 0, h->stream>>>(
                h->d_codes, static_cast<const uint64_t*>(h->d_ids), h->d_off, h->d_len, ncodes,
                static_cast<uint64_t*>(nids), d_noff, m);
        CK(cudaGetLastError());
        if (launches) *launches += 1;
    }
    CK(cudaMemcpyAsync(h->d_off, d_noff, sizeof(int64_t) * kc, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->d_codes) cudaFree(h->d_codes);
    if (h->d_ids) cudaFree(h->d_ids);
    cudaFree(d_noff);
    h->d_codes = ncodes;
    h->d_ids = nids;
    h->h_off = noff;
    h->h_cap = ncap;
    h->arena_cap = total;
    return cudaSuccess;
}

}  // namespace

cudaError_t lists_init(ivfadc_index* h) {
    const int kc = h->cfg.kc;
    h->h_off.assign(kc, 0);
    h->h_len.assign(kc, 0);
    h->h_cap.assign(kc, 0);
    CK(cudaMalloc(&h->d_off, sizeof(int64_t) * kc));
    CK(cudaMalloc(&h->d_len, sizeof(int64_t) * kc));
    CK(cudaMemset(h->d_off, 0, sizeof(int64_t) * kc));
    CK(cudaMemset(h->d_len, 0, sizeof(int64_t) * kc));
    h->d_codes = nullptr;
    h->d_ids = nullptr;
    h->arena_cap = 0;
    h->n_local = 0;
    return cudaSuccess;
}

void lists_free(ivfadc_index* h) {
    if (h->d_off) cudaFree(h->d_off);
    if (h->d_len) cudaFree(h->d_len);
    if (h->d_codes) cudaFree(h->d_codes);
    if (h->d_ids) cudaFree(h->d_ids);
    h->d_off = h->d_len = nullptr;
    h->d_codes = nullptr;
    h->d_ids = nullptr;
}

cudaError_t lists_sync_meta_to_device(ivfadc_index* h) {
    const int kc = h->cfg.kc;
    CK(cudaMemcpyAsync(h->d_len, h->h_len.data(), sizeof(int64_t) * kc, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_off, h->h_off.data(), sizeof(int64_t) * kc, cudaMemcpyHostToDevice, h->stream));
    return cudaStreamSynchronize(h->stream);
}

cudaError_t lists_append(ivfadc_index* h, const int32_t* d_cells, const uint8_t* d_codes, int64_t n,
                         uint64_t first_id, int step, int* launches) {
    if (n <= 0) return cudaSuccess;
    const int kc = h->cfg.kc, m = h->cfg.m;
    cudaStream_t s = h->stream;
    // stable sort of the batch by cell: keys = cells, values = batch index
    CK(h->ws_sort_keys.reserve(sizeof(int32_t) * (size_t)n * 2));
    CK(h->ws_sort_vals.reserve(sizeof(int32_t) * (size_t)n * 2));
    int32_t* keys_out = h->ws_sort_keys.as<int32_t>();
    int32_t* vals_in = h->ws_sort_vals.as<int32_t>();
    int32_t* vals_out = vals_in + n;
    iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(vals_in, n);
    int bits = 1;
    while ((1 << bits) < kc) ++bits;
    size_t tmp_bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_cells, keys_out, vals_in, vals_out, (int)n, 0,
                                       bits, s));
    CK(h->ws_sort_tmp.reserve(tmp_bytes));
    CK(cub::DeviceRadixSort::SortPairs(h->ws_sort_tmp.p, tmp_bytes, d_cells, keys_out, vals_in, vals_out,
                                       (int)n, 0, bits, s));
    // run boundaries per cell
    CK(h->ws_misc.reserve(sizeof(int64_t) * (size_t)kc * 2));
    int64_t* d_run_start = h->ws_misc.as<int64_t>();
    int64_t* d_run_end = d_run_start + kc;
    CK(cudaMemsetAsync(d_run_start, 0, sizeof(int64_t) * kc * 2, s));
    run_bounds_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(keys_out, n, kc, d_run_start, d_run_end);
    CK(cudaGetLastError());
    if (launches) *launches += 3;
    std::vector<int64_t> hs(2 * (size_t)kc);
    CK(cudaMemcpyAsync(hs.data(), d_run_start, sizeof(int64_t) * kc * 2, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    std::vector<int64_t> need(kc);
    int64_t added = 0;
    const int world = h->cfg.shard_world, rank = h->cfg.shard_rank;
    for (int c = 0; c < kc; ++c) {
        int64_t cnt = hs[kc + c] > 0 ? hs[kc + c] - hs[c] : 0;
        if (!h->owns(c)) cnt = 0;
        need[c] = h->h_len[c] + cnt;
        added += cnt;
    }
    CK(reserve_lists(h, need, launches));
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (h->id_dev_bytes == 4)
        scatter_append_kernel<uint32_t><<<grid, 256, 0, s>>>(keys_out, vals_out, n, kc, d_run_start, h->d_off,
                                                             h->d_len, d_codes, m, h->d_codes,
                                                             static_cast<uint32_t*>(h->d_ids), first_id, step,
                                                             rank, world, h->d_owner);
    else
        scatter_append_kernel<uint64_t><<<grid, 256, 0, s>>>(keys_out, vals_out, n, kc, d_run_start, h->d_off,
                                                             h->d_len, d_codes, m, h->d_codes,
                                                             static_cast<uint64_t*>(h->d_ids), first_id, step,
                                                             rank, world, h->d_owner);
    CK(cudaGetLastError());
    if (launches) *launches += 1;
    h->h_len = need;
    h->n_local += added;
    CK(cudaMemcpyAsync(h->d_len, h->h_len.data(), sizeof(int64_t) * kc, cudaMemcpyHostToDevice, s));
    return cudaStreamSynchronize(s);
}

cudaError_t lists_shift_ids(ivfadc_index* h, int64_t by, int* launches) {
    if (by == 0 || h->n_local == 0) return cudaSuccess;
    if (h->id_dev_bytes == 4)
        shift_ids_kernel<uint32_t><<<h->cfg.kc, 256, 0, h->stream>>>(static_cast<uint32_t*>(h->d_ids), h->d_off,
                                                                     h->d_len, by);
    else
        shift_ids_kernel<uint64_t><<<h->cfg.kc, 256, 0, h->stream>>>(static_cast<uint64_t*>(h->d_ids), h->d_off,
                                                                     h->d_len, by);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t lists_delete(ivfadc_index* h, const uint64_t* d_sorted_ids, int64_t n, int64_t* removed,
                         int* launches) {
    *removed = 0;
    if (n <= 0 || h->n_local == 0) return cudaSuccess;
    const int kc = h->cfg.kc, m = h->cfg.m;
    const size_t smem = (size_t)DTHREADS * m;
    if (h->id_dev_bytes == 4) {
        auto kern = delete_compact_kernel<uint32_t>;
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<kc, DTHREADS, smem, h->stream>>>(h->d_codes, static_cast<uint32_t*>(h->d_ids), h->d_off, h->d_len, m,
                                                d_sorted_ids, n);
    } else {
        auto kern = delete_compact_kernel<uint64_t>;
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<kc, DTHREADS, smem, h->stream>>>(h->d_codes, static_cast<uint64_t*>(h->d_ids), h->d_off, h->d_len, m,
                                                d_sorted_ids, n);
    }
    CK(cudaGetLastError());
    if (launches) *launches += 1;
    std::vector<int64_t> nl(kc);
    CK(cudaMemcpyAsync(nl.data(), h->d_len, sizeof(int64_t) * kc, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    int64_t rem = 0;
    for (int c = 0; c < kc; ++c) rem += h->h_len[c] - nl[c];
    h->h_len = nl;
    h->n_local -= rem;
    *removed = rem;
    return cudaSuccess;
}

cudaError_t lists_find(ivfadc_index* h, uint64_t id, int32_t* cell, int64_t* pos, int* launches) {
    *cell = -1;
    *pos = -1;
    if (h->n_local == 0) return cudaSuccess;
    CK(h->ws_misc.reserve(sizeof(int64_t) * 2));
    int64_t* d_res = h->ws_misc.as<int64_t>();
    CK(cudaMemsetAsync(d_res, 0xff, sizeof(int64_t) * 2, h->stream));
    if (h->id_dev_bytes == 4)
        find_id_kernel<uint32_t><<<h->cfg.kc, 128, 0, h->stream>>>(static_cast<const uint32_t*>(h->d_ids), h->d_off,
                                                                   h->d_len, id, d_res);
    else
        find_id_kernel<uint64_t><<<h->cfg.kc, 128, 0, h->stream>>>(static_cast<const uint64_t*>(h->d_ids), h->d_off,
                                                                   h->d_len, id, d_res);
    CK(cudaGetLastError());
    if (launches) *launches += 1;
    int64_t res[2];
    CK(cudaMemcpyAsync(res, d_res, sizeof(res), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *cell = (int32_t)res[0];
    *pos = res[1];
    return cudaSuccess;
}

cudaError_t lists_decode(ivfadc_index* h, int32_t cell, int64_t pos, void* d_vec_out, int* launches) {
    const int D = h->cfg.dim, m = h->cfg.m;
    const uint8_t* code = h->d_codes + (size_t)(h->h_off[cell] + pos) * m;
    const unsigned grid = (unsigned)((D + 127) / 128);
    if (h->cfg.dtype == IVFADC_F32)
        decode_kernel<float><<<grid, 128, 0, h->stream>>>(static_cast<const float*>(h->d_centroids),
                                                          static_cast<const float*>(h->d_cb), h->d_cb_codes, code,
                                                          cell, D, m, h->dsub, h->cfg.ksub,
                                                          static_cast<float*>(d_vec_out));
    else
        decode_kernel<double><<<grid, 128, 0, h->stream>>>(static_cast<const double*>(h->d_centroids),
                                                           static_cast<const double*>(h->d_cb), h->d_cb_codes, code,
                                                           cell, D, m, h->dsub, h->cfg.ksub,
                                                           static_cast<double*>(d_vec_out));
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t lists_export(ivfadc_index* h, int32_t cell, uint64_t* ids_out, uint8_t* codes_out) {
    const int64_t n = h->h_len[cell];
    if (n <= 0) return cudaSuccess;
    const int m = h->cfg.m;
    const int64_t o = h->h_off[cell];
    CK(cudaMemcpyAsync(codes_out, h->d_codes + (size_t)o * m, (size_t)n * m, cudaMemcpyDeviceToHost, h->stream));
    if (h->id_dev_bytes == 8) {
        CK(cudaMemcpyAsync(ids_out, static_cast<uint64_t*>(h->d_ids) + o, sizeof(uint64_t) * n,
                           cudaMemcpyDeviceToHost, h->stream));
    } else {
        CK(h->ws_misc.reserve(sizeof(uint64_t) * (size_t)n));
        widen_ids_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(
            static_cast<const uint32_t*>(h->d_ids) + o, h->ws_misc.as<uint64_t>(), n);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(ids_out, h->ws_misc.p, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, h->stream));
    }
    return cudaStreamSynchronize(h->stream);
}

cudaError_t lists_import(ivfadc_index* h, int32_t cell, const uint64_t* ids, const uint8_t* codes,
                         int64_t len, int* launches) {
    const int kc = h->cfg.kc, m = h->cfg.m;
    std::vector<int64_t> need(h->h_len);
    need[cell] = len;
    CK(reserve_lists(h, need, launches));
    const int64_t o = h->h_off[cell];
    if (len > 0) {
        CK(cudaMemcpyAsync(h->d_codes + (size_t)o * m, codes, (size_t)len * m, cudaMemcpyHostToDevice, h->stream));
        if (h->id_dev_bytes == 8) {
            CK(cudaMemcpyAsync(static_cast<uint64_t*>(h->d_ids) + o, ids, sizeof(uint64_t) * len,
                               cudaMemcpyHostToDevice, h->stream));
        } else {
            CK(h->ws_misc.reserve(sizeof(uint64_t) * (size_t)len));
            CK(cudaMemcpyAsync(h->ws_misc.p, ids, sizeof(uint64_t) * len, cudaMemcpyHostToDevice, h->stream));
            narrow_ids_kernel<<<(unsigned)((len + 255) / 256), 256, 0, h->stream>>>(
                h->ws_misc.as<uint64_t>(), static_cast<uint32_t*>(h->d_ids) + o, len);
            CK(cudaGetLastError());
            if (launches) *launches += 1;
        }
    }
    h->n_local += len - h->h_len[cell];
    h->n_total += len - h->h_len[cell];
    h->h_len[cell] = len;
    CK(cudaMemcpyAsync(h->d_len, h->h_len.data(), sizeof(int64_t) * kc, cudaMemcpyHostToDevice, h->stream));
    return cudaStreamSynchronize(h->stream);
}

namespace {
constexpr int64_t kPackChunk = 1 << 24;  // vectors per staging chunk (codes 16 m MB + ids 128 MB)

template <bool EXPORT>
cudaError_t pack_all(ivfadc_index* h, uint64_t* ids, uint8_t* codes, int* launches) {
    const int kc = h->cfg.kc, m = h->cfg.m;
    std::vector<int64_t> pref(kc + 1, 0);
    for (int c = 0; c < kc; ++c) pref[c + 1] = pref[c] + h->h_len[c];
    const int64_t total = pref[kc];
    if (total == 0) return cudaSuccess;
    CK(h->ws_sort_tmp.reserve(sizeof(int64_t) * (size_t)(kc + 1)));
    int64_t* d_pref = h->ws_sort_tmp.as<int64_t>();
    CK(cudaMemcpyAsync(d_pref, pref.data(), sizeof(int64_t) * (kc + 1), cudaMemcpyHostToDevice, h->stream));
    const int64_t chunk = std::min(total, kPackChunk);
    CK(h->ws_x.reserve((size_t)chunk * m));
    CK(h->ws_misc.reserve(sizeof(uint64_t) * (size_t)chunk));
    uint8_t* pk_codes = h->ws_x.as<uint8_t>();
    uint64_t* pk_ids = h->ws_misc.as<uint64_t>();
    for (int64_t p0 = 0; p0 < total; p0 += chunk) {
        const int64_t n = std::min(chunk, total - p0);
        const unsigned grid = (unsigned)((n + 255) / 256);
        if (!EXPORT) {
            CK(cudaMemcpyAsync(pk_codes, codes + (size_t)p0 * m, (size_t)n * m, cudaMemcpyHostToDevice, h->stream));
            CK(cudaMemcpyAsync(pk_ids, ids + p0, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, h->stream));
        }
        if (h->id_dev_bytes == 4)
            pack_lists_kernel<uint32_t, EXPORT><<<grid, 256, 0, h->stream>>>(
                h->d_codes, static_cast<uint32_t*>(h->d_ids), h->d_off, d_pref, kc, m, p0, n, pk_codes, pk_ids);
        else
            pack_lists_kernel<uint64_t, EXPORT><<<grid, 256, 0, h->stream>>>(
                h->d_codes, static_cast<uint64_t*>(h->d_ids), h->d_off, d_pref, kc, m, p0, n, pk_codes, pk_ids);
        CK(cudaGetLastError());
        if (launches) *launches += 1;
        if (EXPORT) {
            CK(cudaMemcpyAsync(codes + (size_t)p0 * m, pk_codes, (size_t)n * m, cudaMemcpyDeviceToHost, h->stream));
            CK(cudaMemcpyAsync(ids + p0, pk_ids, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, h->stream));
        }
        CK(cudaStreamSynchronize(h->stream));  // the staging buffers are reused by the next chunk
    }
    return cudaSuccess;
}
}  // namespace

// capacity for need[c] entries in every list (sizehint!): one arena allocation instead of geometric regrowth
cudaError_t lists_reserve(ivfadc_index* h, const std::vector<int64_t>& need, int* launches) {
    return reserve_lists(h, need, launches);
}

// every list, cell-ascending, packed: ids uint64[sum len], codes uint8[sum len][m] (src/persistency.jl:68-78)
cudaError_t lists_export_all(ivfadc_index* h, uint64_t* ids_out, uint8_t* codes_out, int* launches) {
    return pack_all<true>(h, ids_out, codes_out, launches);
}

// replace ALL lists by sizes[kc] + packed entries (src/persistency.jl:119-131)
cudaError_t lists_import_all(ivfadc_index* h, const int64_t* sizes, const uint64_t* ids, const uint8_t* codes,
                             int* launches) {
    const int kc = h->cfg.kc;
    std::vector<int64_t> need(sizes, sizes + kc);
    int64_t total = 0;
    for (int c = 0; c < kc; ++c) total += need[c];
    // the old contents are dropped: lengths to zero first so that a regrow copies nothing
    const int64_t old_local = h->n_local;
    std::fill(h->h_len.begin(), h->h_len.end(), 0);
    h->n_local = 0;
    CK(cudaMemsetAsync(h->d_len, 0, sizeof(int64_t) * kc, h->stream));
    CK(reserve_lists(h, need, launches));
    h->h_len = need;
    CK(pack_all<false>(h, const_cast<uint64_t*>(ids), const_cast<uint8_t*>(codes), launches));
    h->n_local = total;
    h->n_total += total - old_local;
    CK(cudaMemcpyAsync(h->d_len, h->h_len.data(), sizeof(int64_t) * kc, cudaMemcpyHostToDevice, h->stream));
    return cudaStreamSynchronize(h->stream);
}

}  // namespace ivf
