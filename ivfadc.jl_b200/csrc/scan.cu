// Search planning, the scan launcher, and the final merges.  See scan_impl.cuh for K2/K3.
#include "scan_impl.cuh"
#include "scanq_impl.cuh"
#include "scanu_impl.cuh"
#include "scanw_impl.cuh"

namespace ivf {

namespace {

// ---------------------------------------------------------------------------------------------
// Planning: group the (query, probe) pairs by inverted list so that one CTA can serve QN queries
// from one pass over a list.  Two bucket classes per list: probe rank 0 (scheduled first: the
// nearest cell gives each query a tight k-th-distance bound early) and ranks >= 1.
// ---------------------------------------------------------------------------------------------
constexpr int PLAN_PAD = 32;  // ints per bucket in the counter array: one 128-byte line each

template <typename BitsT>
__global__ void plan_count_kernel(const int32_t* __restrict__ cells, int64_t npairs, int64_t nq, int w,
                                  int kc, int split, const int64_t* __restrict__ list_len, int* bucket_cnt,
                                  BitsT* thr, BitsT inf_bits, unsigned long long* scanned) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long mylen = 0;
    if (p < nq) thr[p] = inf_bits;
    if (p < npairs) {
        const int cell = cells[p];
        const int64_t len = (cell >= 0 && cell < kc) ? list_len[cell] : 0;
        if (len > 0) {
            const int rank = (int)(p % w);
            atomicAdd(&bucket_cnt[(size_t)(((split && rank) ? kc : 0) + cell) * PLAN_PAD], 1);
            mylen = (unsigned long long)len;
        }
    }
    if (scanned) {  // one global atomic per block (a per-warp atomic on one address serialises 5 000 of them per batch)
        __shared__ unsigned long long blk;
        if (threadIdx.x == 0) blk = 0ull;
        __syncthreads();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mylen += __shfl_xor_sync(0xffffffffu, mylen, o);
        if ((threadIdx.x & 31) == 0 && mylen) atomicAdd(&blk, mylen);
        __syncthreads();
        if (threadIdx.x == 0 && blk) atomicAdd(scanned, blk);
    }
}

// Single CTA: exclusive scans of pairs-per-bucket and groups-per-bucket.
__global__ void __launch_bounds__(1024)
plan_scan_kernel(const int* __restrict__ bucket_cnt, int nb, int qn, int* bucket_off, int* group_off, int4* items) {
    __shared__ int s_p[1024], s_g[1024];
    const int t = threadIdx.x;
    const int per = (nb + 1023) / 1024;
    const int lo = min(nb, t * per), hi = min(nb, lo + per);
    int sp = 0, sg = 0;
    for (int b = lo; b < hi; ++b) {
        const int c = bucket_cnt[(size_t)b * PLAN_PAD];
        sp += c;
        sg += (c + qn - 1) / qn;
    }
    s_p[t] = sp;
    s_g[t] = sg;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan
        const int vp = t >= o ? s_p[t - o] : 0, vg = t >= o ? s_g[t - o] : 0;
        __syncthreads();
        s_p[t] += vp;
        s_g[t] += vg;
        __syncthreads();
    }
    int op = s_p[t] - sp, og = s_g[t] - sg;
    for (int b = lo; b < hi; ++b) {
        const int c = bucket_cnt[(size_t)b * PLAN_PAD];
        bucket_off[b] = op;
        group_off[b] = og;
        if (items) {  // work-item table of the tensor-core kernels: (cell, first pair slot, number of pairs, 0)
            int g = og;
            for (int p = op; p < op + c; p += qn, ++g) items[g] = make_int4(b, p, min(qn, op + c - p), 0);
        }
        op += c;
        og += (c + qn - 1) / qn;
    }
    if (t == 1023) {
        bucket_off[nb] = s_p[1023];
        group_off[nb] = s_g[1023];
    }
}

// (PLAN_PAD: the per-bucket counter and cursor sit in a 128-byte line of their own -- 32 neighbouring counters in
// one line serialise the atomics of a batch in the L2)
__global__ void plan_scatter_kernel(const int32_t* __restrict__ cells, int64_t npairs, int w, int kc,
                                    int split, const int64_t* __restrict__ list_len,
                                    const int* __restrict__ bucket_off, int* cursor,
                                    int32_t* sorted_pairs) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npairs) return;
    const int cell = cells[p];
    if (cell < 0 || cell >= kc || list_len[cell] <= 0) return;
    const int b = ((split && (p % w)) ? kc : 0) + cell;
    const int slot = bucket_off[b] + atomicAdd(&cursor[(size_t)b * PLAN_PAD], 1);
    sorted_pairs[slot] = (int32_t)p;
}

// Work-item table of the query-per-lane kernels: item -> (cell, first pair slot, number of pairs).

// ---------------------------------------------------------------------------------------------
// Final merge, one warp per query: k-way merge of L sorted candidate lists (lane l owns lists
// l, l+32, ...) by (distance, key).  Used for the w per-probe lists of one GPU
// (key = rank << 32 | position, id looked up in the list arena) and for the per-rank lists of a
// sharded search (key and id carried in the input).
// ---------------------------------------------------------------------------------------------
constexpr int MERGE_NL = 4;  // lists per lane => up to 128 lists

template <typename T> struct Cand {
    T d;
    uint64_t key;
};

template <typename T>
__device__ __forceinline__ bool cand_less(const Cand<T>& a, const Cand<T>& b) {
    return a.d < b.d || (a.d == b.d && a.key < b.key);
}

template <typename T, typename IdT>
__global__ void __launch_bounds__(128)
merge_probes_kernel(int64_t nq, int w, int k, const int32_t* __restrict__ cells,
                    const T* __restrict__ pair_d, const uint32_t* __restrict__ pair_pos,
                    const int32_t* __restrict__ pair_cnt, const int64_t* __restrict__ list_off,
                    const IdT* __restrict__ ids_arena, uint64_t* __restrict__ out_ids,
                    T* __restrict__ out_d, uint64_t* __restrict__ out_keys,
                    int32_t* __restrict__ out_cnt) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    int head[MERGE_NL], cnt[MERGE_NL];
    Cand<T> cur[MERGE_NL];
#pragma unroll
    for (int s = 0; s < MERGE_NL; ++s) {
        const int r = lane + 32 * s;
        head[s] = 0;
        cnt[s] = r < w ? pair_cnt[q * w + r] : 0;
        cur[s].d = Limits<T>::inf();
        cur[s].key = ~0ull;
        if (cnt[s] > 0) {
            const size_t o = (size_t)(q * w + r) * k;
            cur[s].d = pair_d[o];
            cur[s].key = ((uint64_t)r << 32) | pair_pos[o];
        }
    }
    int e = 0;
    for (; e < k; ++e) {
        Cand<T> best;
        best.d = Limits<T>::inf();
        best.key = ~0ull;
#pragma unroll
        for (int s = 0; s < MERGE_NL; ++s)
            if (head[s] < cnt[s] && cand_less(cur[s], best)) best = cur[s];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            Cand<T> oth;
            oth.d = __shfl_xor_sync(0xffffffffu, best.d, o);
            oth.key = __shfl_xor_sync(0xffffffffu, best.key, o);
            if (cand_less(oth, best)) best = oth;
        }
        if (best.key == ~0ull) break;  // every list exhausted
        const int r = (int)(best.key >> 32);
        const uint32_t pos = (uint32_t)best.key;
        if (lane == (r & 31)) {
#pragma unroll
            for (int s = 0; s < MERGE_NL; ++s)
                if (s == (r >> 5)) {
                    ++head[s];
                    if (head[s] < cnt[s]) {
                        const size_t o = (size_t)(q * w + r) * k + head[s];
                        cur[s].d = pair_d[o];
                        cur[s].key = ((uint64_t)r << 32) | pair_pos[o];
                    }
                }
        }
        if (lane == 0) {
            const int cell = cells[q * w + r];
            out_ids[q * k + e] = (uint64_t)ids_arena[list_off[cell] + pos];
            out_d[q * k + e] = best.d;
            if (out_keys) out_keys[q * k + e] = best.key;
        }
    }
    for (int x = e + lane; x < k; x += 32) {
        out_ids[q * k + x] = ~0ull;
        out_d[q * k + x] = Limits<T>::inf();
        if (out_keys) out_keys[q * k + x] = ~0ull;
    }
    if (lane == 0) out_cnt[q] = e;
}

// Final selection over UNSORTED candidate rows (the tensor-memory lookup kernel dumps every vector
// whose distance is within the list's k-th-distance bound; the redo kernel writes exact sorted
// rows): one warp per query picks the k smallest of the union of its w rows by
// (distance, probe rank, position) -- the reference's order, src/index.jl:247-257.
//   1. lane minima over a strided share of the candidates; the k-th smallest of the 32 minima bounds
//      the k-th distance of the query;
//   2. candidates within the bound are compacted into shared memory (ballot prefix);
//   3. k rounds of warp arg-min over the compacted set held in registers.  A set larger than
//      MC_CAP (heavy ties) is selected by k filtered sweeps over the rows in global memory.
constexpr int MC_CAP = 128;

__device__ __forceinline__ bool key_less(float da, uint64_t ka, float db, uint64_t kb) {
    return da < db || (da == db && ka < kb);
}

template <typename IdT>
__global__ void __launch_bounds__(128)
merge_cands_kernel(int64_t nq, int w, int k, int ps, int rs, int cap, const int32_t* __restrict__ cells,
                   const float* __restrict__ pair_d, const uint32_t* __restrict__ pair_pos,
                   const int32_t* __restrict__ pair_cnt, const int64_t* __restrict__ list_off,
                   const IdT* __restrict__ ids_arena, uint64_t* __restrict__ out_ids,
                   float* __restrict__ out_d, uint64_t* __restrict__ out_keys, int32_t* __restrict__ out_cnt) {
    __shared__ float s_d[4][MC_CAP];
    __shared__ uint64_t s_k[4][MC_CAP];
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    const int64_t q = (int64_t)blockIdx.x * 4 + wq;
    if (q >= nq) return;
    const float inf = Limits<float>::inf();
    int cnt[MERGE_NL];
#pragma unroll
    for (int s = 0; s < MERGE_NL; ++s) {
        const int r = lane + 32 * s;
        cnt[s] = r < w ? min(pair_cnt[q * w + r], ps) : 0;
    }
    auto count_of = [&](int r) {
        int c = 0;
#pragma unroll
        for (int s = 0; s < MERGE_NL; ++s)
            if (s == (r >> 5)) c = __shfl_sync(0xffffffffu, cnt[s], r & 31);
        return c;
    };
    // 1. bound.  Four rows at a time: the loads of a group are independent and issued together (rows are
    // short -- typically 20..30 candidates -- so one load per row and lane covers them; the tail loop is rare)
    float lmin = inf;
    for (int r0 = 0; r0 < w; r0 += 4) {
        int c[4];
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) c[i] = r0 + i < w ? count_of(r0 + i) : 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = lane < c[i] ? pair_d[(size_t)(q * w + r0 + i) * rs + lane] : inf;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            lmin = fminf(lmin, v[i]);
            for (int e = lane + 32; e < c[i]; e += 32) lmin = fminf(lmin, pair_d[(size_t)(q * w + r0 + i) * rs + e]);
        }
    }
    int rank = 0;
#pragma unroll
    for (int l = 0; l < 32; ++l) {
        const float o = __shfl_sync(0xffffffffu, lmin, l);
        rank += (o < lmin || (o == lmin && l < lane)) ? 1 : 0;
    }
    const int src = __ffs(__ballot_sync(0xffffffffu, rank == min(k, 32) - 1)) - 1;
    const float bound = __shfl_sync(0xffffffffu, lmin, src);
    // 2. compaction of the candidates within the bound (rows in rank order, so equal keys cannot occur)
    int M = 0;
    auto offer = [&](int r, int e, int c, float d, uint32_t pos) {
        const bool pred = e < c && d <= bound && d < inf;
        const unsigned mask = __ballot_sync(0xffffffffu, pred);
        if (pred) {
            const int slot = M + __popc(mask & ((1u << lane) - 1u));
            if (slot < cap) {
                s_d[wq][slot] = d;
                s_k[wq][slot] = ((uint64_t)r << 32) | pos;
            }
        }
        M += __popc(mask);
    };
    for (int r0 = 0; r0 < w; r0 += 4) {
        int c[4];
        float v[4];
        uint32_t pp[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) c[i] = r0 + i < w ? count_of(r0 + i) : 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const size_t o = (size_t)(q * w + r0 + i) * rs + lane;
            v[i] = lane < c[i] ? pair_d[o] : inf;
            pp[i] = lane < c[i] ? pair_pos[o] : 0u;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (c[i] == 0) continue;  // warp-uniform
            offer(r0 + i, lane, c[i], v[i], pp[i]);
            for (int e0 = 32; e0 < c[i]; e0 += 32) {
                const int e = e0 + lane;
                const size_t o = (size_t)(q * w + r0 + i) * rs + e;
                offer(r0 + i, e, c[i], e < c[i] ? pair_d[o] : inf, e < c[i] ? pair_pos[o] : 0u);
            }
        }
    }
    __syncwarp();
    int e = 0;
    bool done = false;
    if (M <= 32 && M <= cap) {
        // 3a. the common case -- at most one candidate per lane, no two equal distances: every lane counts the
        // smaller distances (independent shuffles of the order-preserving bit pattern) and the lanes of rank < k
        // write their result, id gathers in parallel.  Same order as 3b: (distance, probe rank, position).
        const bool have = lane < M;
        const float dm = have ? s_d[wq][lane] : inf;
        const uint64_t km = have ? s_k[wq][lane] : ~0ull;
        const uint32_t kd = have ? f_flip(dm + 0.f) : 0xFFFFFFFFu;   // + 0: -0 and +0 compare equal, as in key_less
        const unsigned same = __match_any_sync(0xffffffffu, kd);
        if (__all_sync(0xffffffffu, !have || __popc(same) == 1)) {
            int rk = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) rk += __shfl_sync(0xffffffffu, kd, j) < kd ? 1 : 0;
            if (have && rk < k) {
                const int cell = cells[q * w + (int)(km >> 32)];
                out_ids[q * k + rk] = (uint64_t)ids_arena[list_off[cell] + (uint32_t)km];
                out_d[q * k + rk] = dm;
                if (out_keys) out_keys[q * k + rk] = km;
            }
            e = min(M, k);
            done = true;
        }
    }
    if (done) {
    } else if (M <= cap) {  // cap = MC_CAP (tests lower it to drive ordinary data through the sweep path)
        // 3b. k rounds of warp arg-min; lane holds entries lane, lane + 32, lane + 64, lane + 96
        float d[4];
        uint64_t key[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int i = lane + 32 * s;
            d[s] = i < M ? s_d[wq][i] : inf;
            key[s] = i < M ? s_k[wq][i] : ~0ull;
        }
        for (; e < k; ++e) {
            float bd = inf;
            uint64_t bk = ~0ull;
#pragma unroll
            for (int s = 0; s < 4; ++s)
                if (key[s] != ~0ull && key_less(d[s], key[s], bd, bk)) { bd = d[s]; bk = key[s]; }
            const uint64_t mine = bk;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, bd, o);
                const uint64_t ok = __shfl_xor_sync(0xffffffffu, bk, o);
                if (ok != ~0ull && (bk == ~0ull || key_less(od, ok, bd, bk))) { bd = od; bk = ok; }
            }
            if (bk == ~0ull) break;  // every candidate taken
            if (mine == bk) {        // keys are unique: exactly one lane owns the winner
#pragma unroll
                for (int s = 0; s < 4; ++s)
                    if (key[s] == bk) key[s] = ~0ull;
            }
            if (lane == 0) {
                const int r = (int)(bk >> 32);
                const int cell = cells[q * w + r];
                out_ids[q * k + e] = (uint64_t)ids_arena[list_off[cell] + (uint32_t)bk];
                out_d[q * k + e] = bd;
                if (out_keys) out_keys[q * k + e] = bk;
            }
        }
    } else {
        // heavy ties: k filtered sweeps over the rows (each sweep finds the smallest candidate after the last winner)
        float ld = 0.f;
        uint64_t lk = 0;
        for (; e < k; ++e) {
            float bd = inf;
            uint64_t bk = ~0ull;
            for (int r = 0; r < w; ++r) {
                const int c = count_of(r);
                const size_t rb = (size_t)(q * w + r) * rs;
                for (int i = lane; i < c; i += 32) {
                    const float dd = pair_d[rb + i];
                    const uint64_t kk = ((uint64_t)r << 32) | pair_pos[rb + i];
                    if (dd < inf && (e == 0 || key_less(ld, lk, dd, kk)) && (bk == ~0ull || key_less(dd, kk, bd, bk))) {
                        bd = dd;
                        bk = kk;
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, bd, o);
                const uint64_t ok = __shfl_xor_sync(0xffffffffu, bk, o);
                if (ok != ~0ull && (bk == ~0ull || key_less(od, ok, bd, bk))) { bd = od; bk = ok; }
            }
            if (bk == ~0ull) break;
            ld = bd;
            lk = bk;
            if (lane == 0) {
                const int r = (int)(bk >> 32);
                const int cell = cells[q * w + r];
                out_ids[q * k + e] = (uint64_t)ids_arena[list_off[cell] + (uint32_t)bk];
                out_d[q * k + e] = bd;
                if (out_keys) out_keys[q * k + e] = bk;
            }
        }
    }
    for (int x = e + lane; x < k; x += 32) {
        out_ids[q * k + x] = ~0ull;
        out_d[q * k + x] = inf;
        if (out_keys) out_keys[q * k + x] = ~0ull;
    }
    if (lane == 0) out_cnt[q] = e;
}

// Merge `parts` candidate sets [parts][nq][k] (sorted rows padded with key = ~0) into [nq][k].
template <typename T>
__global__ void __launch_bounds__(128)
merge_parts_kernel(int parts, int64_t nq, int k, const uint64_t* __restrict__ in_ids,
                   const T* __restrict__ in_d, const uint64_t* __restrict__ in_keys,
                   uint64_t* __restrict__ out_ids, T* __restrict__ out_d,
                   int32_t* __restrict__ out_cnt) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    int head[MERGE_NL];
    Cand<T> cur[MERGE_NL];
#pragma unroll
    for (int s = 0; s < MERGE_NL; ++s) {
        const int r = lane + 32 * s;
        head[s] = 0;
        cur[s].d = Limits<T>::inf();
        cur[s].key = ~0ull;
        if (r < parts) {
            const size_t o = ((size_t)r * nq + q) * k;
            cur[s].d = in_d[o];
            cur[s].key = in_keys[o];
        }
    }
    int e = 0;
    for (; e < k; ++e) {
        Cand<T> best;
        best.d = Limits<T>::inf();
        best.key = ~0ull;
        int bs = -1;
#pragma unroll
        for (int s = 0; s < MERGE_NL; ++s)
            if (cur[s].key != ~0ull && cand_less(cur[s], best)) {
                best = cur[s];
                bs = lane + 32 * s;
            }
        int br = bs;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            Cand<T> oth;
            oth.d = __shfl_xor_sync(0xffffffffu, best.d, o);
            oth.key = __shfl_xor_sync(0xffffffffu, best.key, o);
            const int orr = __shfl_xor_sync(0xffffffffu, br, o);
            // keys of different parts never tie (a (rank, position) pair lives on one shard)
            if (cand_less(oth, best)) {
                best = oth;
                br = orr;
            }
        }
        if (best.key == ~0ull) break;
        if (lane == (br & 31)) {
#pragma unroll
            for (int s = 0; s < MERGE_NL; ++s)
                if (s == (br >> 5)) {
                    const size_t o = ((size_t)br * nq + q) * k;
                    if (lane == (br & 31)) out_ids[q * k + e] = in_ids[o + head[s]];
                    ++head[s];
                    if (head[s] < k) {
                        cur[s].d = in_d[o + head[s]];
                        cur[s].key = in_keys[o + head[s]];
                    } else {
                        cur[s].d = Limits<T>::inf();
                        cur[s].key = ~0ull;
                    }
                }
        }
        if (lane == 0) out_d[q * k + e] = best.d;
    }
    for (int x = e + lane; x < k; x += 32) {
        out_ids[q * k + x] = ~0ull;
        out_d[q * k + x] = Limits<T>::inf();
    }
    if (lane == 0) out_cnt[q] = e;
}

// Same selection without a dependent global load per result (the k-way merge above pays one memory round trip per
// round: 33 us for 10 000 queries x 2 parts x 10): the parts' sorted lists go to shared memory once, every entry
// finds its rank = its index in its own list + the number of smaller entries in every other list (binary search:
// the lists are sorted by (distance, key)), and the entries of rank < k write themselves out.  parts x k <= 256.
constexpr int MPR_MAX = 256;
template <typename T>
__global__ void __launch_bounds__(128)
merge_parts_rank_kernel(int parts, int64_t nq, int k, const uint64_t* __restrict__ in_ids,
                        const T* __restrict__ in_d, const uint64_t* __restrict__ in_keys,
                        uint64_t* __restrict__ out_ids, T* __restrict__ out_d, int32_t* __restrict__ out_cnt) {
    __shared__ T s_d[4][MPR_MAX];
    __shared__ unsigned long long s_k[4][MPR_MAX];
    const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
    const int64_t q = (int64_t)blockIdx.x * 4 + wv;
    if (q >= nq) return;
    const int n = parts * k;
    for (int x = lane; x < n; x += 32) {
        const int p = x / k, i = x - p * k;
        const size_t o = ((size_t)p * nq + q) * k + i;
        s_d[wv][x] = in_d[o];
        s_k[wv][x] = in_keys[o];
    }
    __syncwarp();
    int valid = 0;
    for (int x = lane; x < n; x += 32) {
        const int p = x / k, i = x - p * k;
        Cand<T> me;
        me.d = s_d[wv][x];
        me.key = s_k[wv][x];
        if (me.key == ~0ull) continue;   // padding behind a part's last candidate
        ++valid;
        int rank = i;
        for (int o = 0; o < parts; ++o) {
            if (o == p) continue;
            int lo = 0, hi = k;            // first index of part o that is not less than me
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                Cand<T> c;
                c.d = s_d[wv][o * k + mid];
                c.key = s_k[wv][o * k + mid];
                if (c.key != ~0ull && cand_less(c, me)) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < k) {
            out_ids[q * k + rank] = in_ids[((size_t)p * nq + q) * k + i];
            out_d[q * k + rank] = me.d;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) valid += __shfl_xor_sync(0xffffffffu, valid, o);
    const int e = min(valid, k);
    for (int x = e + lane; x < k; x += 32) {
        out_ids[q * k + x] = ~0ull;
        out_d[q * k + x] = Limits<T>::inf();
    }
    if (lane == 0) out_cnt[q] = e;
}

// ---------------------------------------------------------------------------------------------
// Scan dispatch
// ---------------------------------------------------------------------------------------------
template <typename T, int QN, int MC, int R>
cudaError_t launch_scan_inst(const ivfadc_index* h, const ScanArgs<T>& a, int grid, size_t smem, cudaStream_t s) {
    auto kern = scan_kernel<T, QN, MC, R>;
    cudaError_t e = ensure_smem(h, reinterpret_cast<const void*>(kern), smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, STHREADS, smem, s>>>(a);
    return cudaGetLastError();
}

template <typename T, int QN, int MC>
cudaError_t launch_scan_r(const ivfadc_index* h, const ScanArgs<T>& a, int grid, size_t smem, cudaStream_t s) {
    if (a.k <= 32) return launch_scan_inst<T, QN, MC, 1>(h, a, grid, smem, s);
    return launch_scan_inst<T, QN, MC, 4>(h, a, grid, smem, s);
}

template <typename T, int QN>
cudaError_t launch_scan_m(const ivfadc_index* h, const ScanArgs<T>& a, int grid, size_t smem, cudaStream_t s) {
    switch (a.m) {
        case 8: return launch_scan_r<T, QN, 8>(h, a, grid, smem, s);
        case 12: return launch_scan_r<T, QN, 12>(h, a, grid, smem, s);
        case 16: return launch_scan_r<T, QN, 16>(h, a, grid, smem, s);
        default: return launch_scan_r<T, QN, 0>(h, a, grid, smem, s);
    }
}

constexpr size_t kSmemPreferred = 74 * 1024;   // 3 CTAs per SM

template <typename T> int choose_qn(int m, int dsub, int k) {
    if (scan_smem_bytes<T, 4>(m, dsub, k) <= kSmemPreferred) return 4;
    if (scan_smem_bytes<T, 2>(m, dsub, k) <= kSmemPreferred) return 2;
    if (scan_smem_bytes<T, 1>(m, dsub, k) <= kSmemPreferred) return 1;
    if (scan_smem_bytes<T, 4>(m, dsub, k) <= kSmemMax) return 4;
    if (scan_smem_bytes<T, 2>(m, dsub, k) <= kSmemMax) return 2;
    if (scan_smem_bytes<T, 1>(m, dsub, k) <= kSmemMax) return 1;
    return 0;
}

template <typename T> size_t smem_for(int qn, int m, int dsub, int k) {
    return qn == 4 ? scan_smem_bytes<T, 4>(m, dsub, k)
                   : qn == 2 ? scan_smem_bytes<T, 2>(m, dsub, k) : scan_smem_bytes<T, 1>(m, dsub, k);
}

int choose_qn_h(const ivfadc_index* h, int k) {
    return h->cfg.dtype == IVFADC_F32 ? choose_qn<float>(h->cfg.m, h->dsub, k)
                                      : choose_qn<double>(h->cfg.m, h->dsub, k);
}

// ---------------------------------------------------------------------------------------------
// Query-per-lane path (scanq_impl.cuh): fp32, k <= 16, m in {4, 8, 12, 16}
// ---------------------------------------------------------------------------------------------
bool scanq_fast(const ivfadc_index* h) { return !(h->cfg.flags & IVFADC_FLAG_LUT_EXACT) && h->d_afrag; }

bool scanq_shape_ok(const ivfadc_index* h, int k) {
    const int m = h->cfg.m;
    if (h->cfg.dtype != IVFADC_F32 || k > QMAXK || m % QCS != 0 || m > 16 || h->cfg.ksub > 256) return false;
    return scanq_smem_layout(m, h->dsub, scanq_fast(h)).total <= kSmemMax;
}

// Which kernel serves this batch.  flags (ivfadc_config.flags): IVFADC_FLAG_SCAN_LEGACY forces the
// vector-per-lane kernel, IVFADC_FLAG_SCAN_QLANE forces the query-per-lane kernel whenever the
// shape allows it; otherwise the query-per-lane kernel is used when its 32-query groups are
// reasonably full (>= 8 probes per list on average).
int scanu_dup(const ivfadc_index* h);
bool use_scanq(const ivfadc_index* h, int64_t npairs, int k) {
    if (h->cfg.flags & IVFADC_FLAG_SCAN_LEGACY) return false;
    if (h->cfg.metric_coarse != IVFADC_SQEUCLIDEAN) return false;   // other metrics: exact tables on the general kernel
    if (!scanq_shape_ok(h, k)) return false;
    if (h->cfg.flags & IVFADC_FLAG_SCAN_QLANE) return true;
    if (npairs < (int64_t)8 * h->cfg.kc) return false;
    // Cost model fitted to the round-2 runs ("which scan kernel"; microseconds of GPU time):
    //  * query per lane: kc (n/32 + 1/2) work items, the same cost whether 5 or 32 queries share one.  A pass of
    //    `tables` table rounds costs tables x (0.0022 + 0.0040 fill) + 0.016 (fill = vectors of the pass / 1152: the
    //    lookups are bound by the tensor-memory pipe, the rest is the fixed cost of a round and of the selection);
    //    an item never takes less than the loaders need to stage it (~0.085);
    //  * vector per lane: 0.0081 per pair for its exact lookup tables (twice that at dsub = 16) + the pairs' code
    //    bytes at ~3.2 TB/s.  Few queries per (long) list, as in config D (10 per list of 6100), favour it; short
    //    lists with many queries (64 vectors, 150 queries) do not: 1.35 ms against ~0.5 ms on the 1 M / 1024 shape.
    // (a shard owns every world-th cell: its share of the cells and of the pairs, its own vectors)
    const int world = std::max(1, h->cfg.shard_world);
    const double kc = std::max(1.0, (double)h->cfg.kc / world), np_loc = (double)npairs / world, nbar = np_loc / kc;
    const double lbar = std::max(1.0, (double)h->n_local / kc);
    const int tables = h->cfg.m * ((h->dsub <= 8 || h->dsub == 16) ? scanu_dup(h) : 1);
    const double vp = (double)WShape<12>::VP, nfull = std::floor(lbar / vp), frem = (lbar - nfull * vp) / vp;
    double item = nfull * (tables * 0.0062 + 0.016);
    if (frem > 0.0) item += tables * (0.0022 + 0.0040 * frem) + 0.016;
    item = std::max(item, 0.085);
    const double est_q = kc * (nbar / 32.0 + 0.5) * item;
    const double est_v = np_loc * 0.0081 * std::max(1.0, h->dsub / 8.0) + np_loc * lbar * h->cfg.m / 3.2e6;
    return est_q < est_v;
}

// tensor-memory lookup kernel (scanu_impl.cuh), the default: fp32, k <= 16, m in {4, 8, 12, 16} with dsub <= 8,
// or m in {4, 8} with dsub = 16 (two 8-dim tables per code byte)
int scanu_dup(const ivfadc_index* h) { return h->dsub == 16 ? 2 : 1; }
bool scanu_shape_ok(const ivfadc_index* h) {
    const int mt = h->cfg.m * scanu_dup(h);
    return (h->dsub <= 8 || h->dsub == 16) && h->cfg.m % 4 == 0 && mt <= 16 && scanu_smem_layout(mt).total + 1024 <= kSmemMax;
}
bool use_scanu(const ivfadc_index* h) {
    if (h->cfg.flags & (IVFADC_FLAG_LUT_EXACT | IVFADC_FLAG_LUT_MMASYNC)) return false;
    return h->d_tcU != nullptr && scanu_shape_ok(h);
}
// warp-specialised version (scanw_impl.cuh), the default; IVFADC_FLAG_SCAN_TMEM_V1 keeps the round-1 kernel
bool use_scanw(const ivfadc_index* h) {
    if (h->cfg.flags & IVFADC_FLAG_SCAN_TMEM_V1) return false;
    return scanw_smem_layout(h->cfg.m * scanu_dup(h), h->cfg.m, 12).total + 1024 <= kSmemMax;
}
// Shape of the warp-specialised CTA: 12 scanners x 96 distances (1152 vectors per pass) or 16 x 64 (1024 per pass,
// four scanning warps per scheduler: ~6 % faster per pass).  The wider shape wins unless lists of 1025..1152 vectors
// make it pay a second pass: count the passes of both shapes over the lists this handle holds (recounted when the
// index changed) and take 16 when 0.97 x passes16 < passes12.  IVFADC_SCANW_SHAPE=12|16 pins one (A/B runs).
int scanw_shape(const ivfadc_index* h) {
    static const int pinned = [] {
        const char* e = getenv("IVFADC_SCANW_SHAPE");
        return e ? atoi(e) : 0;
    }();
    if (pinned == 12 || pinned == 16) return pinned;
    if (h->scanw_ws_n != h->n_local || h->scanw_ws == 0) {
        double p12 = 0.0, p16 = 0.0;
        for (int64_t len : h->h_len) {
            if (len <= 0) continue;
            p12 += (double)((len + WShape<12>::VP - 1) / WShape<12>::VP);
            p16 += (double)((len + WShape<16>::VP - 1) / WShape<16>::VP);
        }
        h->scanw_ws = 0.97 * p16 < p12 ? 16 : 12;
        h->scanw_ws_n = h->n_local;
    }
    return h->scanw_ws;
}

template <int NP, bool DBG, int DUP>
cudaError_t launch_scanu_inst(const ivfadc_index* h, bool v1, const ScanUArgs& ua, unsigned grid, cudaStream_t s) {
    // v1 = the round-1 kernel (every warp scans, per-table CTA barrier; IVFADC_FLAG_SCAN_TMEM_V1), else the
    // warp-specialised kernel of scanw_impl.cuh
    cudaError_t e;
    if (v1) {
        const size_t smem = scanu_smem_layout(4 * NP * DUP).total;
        if ((e = ensure_smem(h, reinterpret_cast<const void*>(&scanu_kernel<NP, DBG, DUP>), smem)) != cudaSuccess) return e;
        scanu_kernel<NP, DBG, DUP><<<grid, QTHREADS, smem, s>>>(ua);
    } else {
        if (!DBG && scanw_shape(h) == 16) {   // (the phase profile / table dump of the DBG instantiation knows the 12 x 96 shape)
            const size_t smem = scanw_smem_layout(4 * NP * DUP, 4 * NP, 16).total;
            if ((e = ensure_smem(h, reinterpret_cast<const void*>(&scanw_kernel<NP, false, DUP, 16>), smem)) != cudaSuccess) return e;
            scanw_kernel<NP, false, DUP, 16><<<grid, WShape<16>::THREADS, smem, s>>>(ua);
        } else {
            const size_t smem = scanw_smem_layout(4 * NP * DUP, 4 * NP, 12).total;
            if ((e = ensure_smem(h, reinterpret_cast<const void*>(&scanw_kernel<NP, DBG, DUP, 12>), smem)) != cudaSuccess) return e;
            scanw_kernel<NP, DBG, DUP, 12><<<grid, WShape<12>::THREADS, smem, s>>>(ua);
        }
    }
    return cudaGetLastError();
}
template <int NP, int DUP = 1>
cudaError_t launch_scanu_np(const ivfadc_index* h, bool v1, const ScanUArgs& ua, unsigned grid, cudaStream_t s) {
    if constexpr (DUP == 1) {
        if (ua.dbg) return launch_scanu_inst<NP, true, 1>(h, v1, ua, grid, s);  // table dump / timeline (bring-up, dsub <= 8)
    }
    return launch_scanu_inst<NP, false, DUP>(h, v1, ua, grid, s);
}

// Row stride of the per-pair candidate arrays: k sorted entries, or U_CAP unsorted candidates when
// the tensor-memory lookup kernel serves the batch.
int pair_stride(const ivfadc_index* h, int64_t npairs, int k) {
    return (use_scanq(h, npairs, k) && use_scanu(h)) ? std::max(k, U_CAP) : k;
}

template <typename T, int MC>
cudaError_t launch_redo_r(const ivfadc_index* h, const ScanArgs<T>& a, const int32_t* redo_pairs, const int* redo_cnt, int grid,
                          size_t smem, cudaStream_t s) {
    auto kern = scan_redo_kernel<T, MC, 1>;
    cudaError_t e = ensure_smem(h, reinterpret_cast<const void*>(kern), smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, STHREADS, smem, s>>>(a, redo_pairs, redo_cnt);
    return cudaGetLastError();
}

template <typename T>
cudaError_t search_t(ivfadc_index* h, const void* dQ, int64_t nq, int k, int w, const int32_t* d_cells,
                     const void* d_dc, uint64_t* d_ids, void* d_dists, uint64_t* d_keys,
                     int32_t* d_counts, uint64_t* d_scanned, cudaStream_t s, int* launches) {
    typedef typename Limits<T>::bits_t bits_t;
    const int kc = h->cfg.kc;
    const int64_t npairs = nq * w;
    const bool qlane = use_scanq(h, npairs, k);
    const int nb = qlane ? kc : 2 * kc;
    const int qn = qlane ? QG : choose_qn<T>(h->cfg.m, h->dsub, k);
    cudaError_t e;

    // workspaces (sizes validated / reserved by the caller through scan_plan_sizes)
    // [2kc][PLAN_PAD] (word 0 count, word 1 cursor) | redo_cnt | item_counter | [2kc+1] off | [2kc+1] groups
    int* bucket_cnt = h->ws_bucket.as<int>();
    int* cursor = bucket_cnt + 1;
    int* redo_cnt = bucket_cnt + (size_t)2 * kc * PLAN_PAD;
    int* item_counter = redo_cnt + 1;
    int* bucket_off = item_counter + 1;
    int* group_off = bucket_off + 2 * kc + 1;
    int32_t* sorted_pairs = h->ws_sorted.as<int32_t>();
    int32_t* redo_pairs = sorted_pairs + npairs;
    T* pair_d = h->ws_pair_d.as<T>();
    uint32_t* pair_pos = h->ws_pair_pos.as<uint32_t>();
    int32_t* pair_cnt = h->ws_pair_cnt.as<int32_t>();
    bits_t* thr = h->ws_thr.as<bits_t>();

    if ((e = cudaMemsetAsync(bucket_cnt, 0, sizeof(int) * (size_t)nb * PLAN_PAD, s)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(redo_cnt, 0, 2 * sizeof(int), s)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(pair_cnt, 0, sizeof(int32_t) * npairs, s)) != cudaSuccess) return e;

    const int pthreads = 256;
    const unsigned pgrid = (unsigned)((std::max<int64_t>(npairs, nq) + pthreads - 1) / pthreads);
    T inf = Limits<T>::inf();
    bits_t inf_bits;
    memcpy(&inf_bits, &inf, sizeof(T));
    const int split = qlane ? 0 : 1;
    plan_count_kernel<bits_t><<<pgrid, pthreads, 0, s>>>(d_cells, npairs, nq, w, kc, split, h->d_len, bucket_cnt,
                                                         thr, inf_bits, (unsigned long long*)d_scanned);
    const bool want_items = qlane && use_scanu(h);  // nb = kc there: bucket = cell
    plan_scan_kernel<<<1, 1024, 0, s>>>(bucket_cnt, nb, qn, bucket_off, group_off,
                                        want_items ? h->ws_items.as<int4>() : nullptr);
    plan_scatter_kernel<<<pgrid, pthreads, 0, s>>>(d_cells, npairs, w, kc, split, h->d_len, bucket_off, cursor,
                                                   sorted_pairs);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    *launches += 3;

    ScanArgs<T> a;
    a.Q = static_cast<const T*>(dQ);
    a.C = static_cast<const T*>(h->d_centroids);
    a.cb = static_cast<const T*>(h->d_cb);
    a.cb_codes = h->d_cb_codes;
    a.cb_identity = h->cb_identity;
    a.metric = h->cfg.metric_coarse;
    a.D = h->cfg.dim; a.m = h->cfg.m; a.dsub = h->dsub; a.ksub = h->cfg.ksub; a.kc = kc; a.w = w; a.k = k;
    a.list_off = h->d_off; a.list_len = h->d_len; a.codes = h->d_codes;
    a.cells = d_cells; a.dc = static_cast<const T*>(d_dc);
    a.bucket_off = bucket_off; a.group_off = group_off; a.nb = nb; a.sorted_pairs = sorted_pairs;
    a.pair_d = pair_d; a.pair_pos = pair_pos; a.pair_cnt = pair_cnt; a.thr = thr;
    // ps = capacity of a pair's row; rs = row stride.  Candidate rows (ps != k) keep the distances and the positions
    // of a pair next to each other -- row = [ps distances][ps positions] of ONE array -- so that the scan kernel
    // addresses both with one pointer and the merge finds them in neighbouring cache lines.
    const int ps = pair_stride(h, npairs, k);
    const int rs = ps != k ? 2 * ps : ps;
    if (ps != k) pair_pos = reinterpret_cast<uint32_t*>(pair_d) + ps;
    a.pair_pos = pair_pos;
    a.pstride = rs;

    // upper bound on the number of work items: every bucket adds at most one partial group
    int64_t max_items = std::min<int64_t>(npairs, npairs / qn + nb);
    if (max_items < 1) max_items = 1;
    if (h->stats_timing) cudaEventRecord(h->ev[2], s);
    h->stats.last_scan_kernel = !qlane ? 1 : use_scanu(h) ? (use_scanw(h) ? 5 : 4) : 2;
    if (qlane) {
        if constexpr (sizeof(T) == 4) {
            ScanQArgs qa;
            qa.Q = a.Q; qa.C = a.C; qa.cb = a.cb; qa.cb_codes = a.cb_codes; qa.cb_identity = a.cb_identity;
            qa.D = a.D; qa.m = a.m; qa.dsub = a.dsub; qa.ksub = a.ksub; qa.kc = kc; qa.w = w; qa.k = k;
            qa.list_off = a.list_off; qa.list_len = a.list_len; qa.codes = a.codes; qa.dc = a.dc;
            qa.bucket_off = bucket_off; qa.group_off = group_off; qa.sorted_pairs = sorted_pairs;
            qa.pair_d = pair_d; qa.pair_pos = pair_pos; qa.pair_cnt = pair_cnt;
            qa.redo_pairs = redo_pairs; qa.redo_cnt = redo_cnt;
            if (use_scanu(h)) {
                ScanUArgs uq;
                uq.q = qa;
                uq.tcU = static_cast<const float*>(h->d_tcU);
                uq.tcH = h->d_tcH;
                uq.ew = h->tch_ew;
                uq.en = h->tch_en;
                uq.items = h->ws_items.as<int4>();
                uq.item_counter = item_counter;
                uq.pstride = ps;
                uq.rstride = rs;
                uq.err = h->d_err;
                uq.dbg = static_cast<float*>(h->d_dbg_lut);
                const int num_sms = h->num_sms > 0 ? h->num_sms : 148;
                const int dup = scanu_dup(h);
                const bool v1 = !use_scanw(h);
                const unsigned ugrid = (unsigned)std::min<int64_t>(max_items, num_sms);  // persistent: one CTA per SM
                if (dup == 2) {  // dsub = 16: two 8-dim tables per code byte
                    uq.q.dsub = 8;
                    if (a.m == 4) e = launch_scanu_np<1, 2>(h, v1, uq, ugrid, s);
                    else e = launch_scanu_np<2, 2>(h, v1, uq, ugrid, s);
                } else {
                    switch (a.m) {
                        case 4: e = launch_scanu_np<1>(h, v1, uq, ugrid, s); break;
                        case 8: e = launch_scanu_np<2>(h, v1, uq, ugrid, s); break;
                        case 12: e = launch_scanu_np<3>(h, v1, uq, ugrid, s); break;
                        default: e = launch_scanu_np<4>(h, v1, uq, ugrid, s); break;
                    }
                }
                if (e != cudaSuccess) return e;
                *launches += 1;
            } else {
            const bool fast = scanq_fast(h);
            qa.afrag = static_cast<const float4*>(h->d_afrag);
            qa.wnfrag = static_cast<const float2*>(h->d_wnfrag);
            qa.ntiles = h->frag_ntiles; qa.ksteps = h->frag_ksteps;
            const size_t qsmem = scanq_smem_layout(a.m, a.dsub, fast).total;
            e = fast ? ensure_smem(h, reinterpret_cast<const void*>(&scanq_kernel<true>), qsmem)
                     : ensure_smem(h, reinterpret_cast<const void*>(&scanq_kernel<false>), qsmem);
            if (e != cudaSuccess) return e;
            if (fast) scanq_kernel<true><<<(unsigned)max_items, QTHREADS, qsmem, s>>>(qa);
            else scanq_kernel<false><<<(unsigned)max_items, QTHREADS, qsmem, s>>>(qa);
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
            *launches += 1;
            }
            // pairs whose candidate list overflowed (heavy ties): general kernel, one pair per item
            const size_t rsmem = smem_for<T>(1, h->cfg.m, h->dsub, k);
            const int rgrid = 2 * 148;
            switch (a.m) {
                case 8: e = launch_redo_r<T, 8>(h, a, redo_pairs, redo_cnt, rgrid, rsmem, s); break;
                case 12: e = launch_redo_r<T, 12>(h, a, redo_pairs, redo_cnt, rgrid, rsmem, s); break;
                case 16: e = launch_redo_r<T, 16>(h, a, redo_pairs, redo_cnt, rgrid, rsmem, s); break;
                default: e = launch_redo_r<T, 0>(h, a, redo_pairs, redo_cnt, rgrid, rsmem, s); break;
            }
            if (e != cudaSuccess) return e;
            *launches += 1;
        }
    } else {
        const size_t smem = smem_for<T>(qn, h->cfg.m, h->dsub, k);
        if (qn == 4) e = launch_scan_m<T, 4>(h, a, (int)max_items, smem, s);
        else if (qn == 2) e = launch_scan_m<T, 2>(h, a, (int)max_items, smem, s);
        else e = launch_scan_m<T, 1>(h, a, (int)max_items, smem, s);
        if (e != cudaSuccess) return e;
        *launches += 1;
    }
    if (h->stats_timing) cudaEventRecord(h->ev[3], s);

    const unsigned mgrid = (unsigned)((nq + 3) / 4);
    if (ps != k) {  // unsorted candidate rows (fp32 only)
        const int mcap = (h->cfg.flags & IVFADC_FLAG_TEST_MERGE_SWEEP) ? 4 : MC_CAP;
        if constexpr (sizeof(T) == 4) {
            if (h->id_dev_bytes == 4)
                merge_cands_kernel<uint32_t><<<mgrid, 128, 0, s>>>(
                    nq, w, k, ps, rs, mcap, d_cells, pair_d, pair_pos, pair_cnt, h->d_off, static_cast<const uint32_t*>(h->d_ids),
                    d_ids, static_cast<float*>(d_dists), d_keys, d_counts);
            else
                merge_cands_kernel<uint64_t><<<mgrid, 128, 0, s>>>(
                    nq, w, k, ps, rs, mcap, d_cells, pair_d, pair_pos, pair_cnt, h->d_off, static_cast<const uint64_t*>(h->d_ids),
                    d_ids, static_cast<float*>(d_dists), d_keys, d_counts);
        }
    } else if (h->id_dev_bytes == 4)
        merge_probes_kernel<T, uint32_t><<<mgrid, 128, 0, s>>>(
            nq, w, k, d_cells, pair_d, pair_pos, pair_cnt, h->d_off, static_cast<const uint32_t*>(h->d_ids),
            d_ids, static_cast<T*>(d_dists), d_keys, d_counts);
    else
        merge_probes_kernel<T, uint64_t><<<mgrid, 128, 0, s>>>(
            nq, w, k, d_cells, pair_d, pair_pos, pair_cnt, h->d_off, static_cast<const uint64_t*>(h->d_ids),
            d_ids, static_cast<T*>(d_dists), d_keys, d_counts);
    *launches += 1;
    return cudaGetLastError();
}

}  // namespace

int scan_max_k() { return 128; }

// Once at create: the codebook as tensor-core operand fragments for the FAST table builder.
cudaError_t scanq_prepare(ivfadc_index* h, cudaStream_t s, int* launches) {
    const int m = h->cfg.m;
    if (h->cfg.dtype != IVFADC_F32 || m % QCS != 0 || m > 16 || h->cfg.ksub > 256) return cudaSuccess;
    h->frag_ntiles = (h->cfg.ksub + 15) / 16;
    h->frag_ksteps = (h->dsub + 7) / 8;
    const size_t tiles = (size_t)m * h->frag_ntiles;
    cudaError_t e = cudaMalloc(&h->d_afrag, tiles * h->frag_ksteps * 32 * 2 * sizeof(float4));
    if (e != cudaSuccess) return e;
    e = cudaMalloc(&h->d_wnfrag, tiles * 32 * sizeof(float2));
    if (e != cudaSuccess) return e;
    const int n = (int)tiles * 32;
    prep_frags_kernel<<<(n + 127) / 128, 128, 0, s>>>(static_cast<const float*>(h->d_cb), m, h->cfg.ksub, h->dsub,
                                                     h->frag_ntiles, h->frag_ksteps,
                                                     static_cast<float4*>(h->d_afrag),
                                                     static_cast<float2*>(h->d_wnfrag));
    if (launches) *launches += 1;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    // operand blocks of the tensor-memory lookup kernel (rows = code values)
    if (scanu_shape_ok(h)) {
        const int dup = scanu_dup(h);
        const size_t bytes = (size_t)m * dup * U_BSUB;
        if ((e = cudaMalloc(&h->d_tcU, bytes)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(h->d_tcU, 0, bytes, s)) != cudaSuccess) return e;
        if (!h->d_err) {
            if ((e = cudaMalloc(&h->d_err, sizeof(int))) != cudaSuccess) return e;
            if ((e = cudaMemsetAsync(h->d_err, 0, sizeof(int), s)) != cudaSuccess) return e;
        }
        const int n = m * dup * h->cfg.ksub;
        prep_tcu_kernel<<<(n + 255) / 256, 256, 0, s>>>(static_cast<const float*>(h->d_cb), h->d_cb_codes,
                                                        h->cb_identity ? 1 : 0, m, h->cfg.ksub, h->dsub, dup,
                                                        static_cast<float*>(h->d_tcU));
        if (launches) *launches += 1;
        // fp16 two-piece operand blocks of the warp-specialised kernel
        const size_t hbytes = (size_t)m * dup * W_BSUB;
        if ((e = cudaMalloc(&h->d_tcH, hbytes)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(h->d_tcH, 0, hbytes, s)) != cudaSuccess) return e;
        prep_tch_kernel<<<(n + 255) / 256, 256, 0, s>>>(static_cast<const float*>(h->d_cb), h->d_cb_codes,
                                                        h->cb_identity ? 1 : 0, m, h->cfg.ksub, h->dsub, dup,
                                                        h->tch_ew, h->tch_en, static_cast<__half*>(h->d_tcH));
        if (launches) *launches += 1;
    }
    return cudaGetLastError();
}

bool scan_takes_tensor_path(const ivfadc_index* h, int64_t npairs, int k) { return use_scanq(h, npairs, k) && use_scanu(h); }

bool scan_supported(const ivfadc_index* h, std::string* why) {
    if (h->cfg.ksub > 256 || h->cfg.ksub < 1) {
        if (why) *why = "codebooks with more than 256 codewords (UInt16 codes) are outside the hot-path scope";
        return false;
    }
    if (choose_qn_h(h, 1) == 0) {
        if (why) *why = "m * 256 lookup-table entries do not fit in shared memory";
        return false;
    }
    return true;
}

ScanPlanSizes scan_plan_sizes(const ivfadc_index* h, int64_t nq, int w, int k) {
    ScanPlanSizes z;
    const int64_t npairs = nq * w;
    const int nb = 2 * h->cfg.kc;
    z.bucket_bytes = sizeof(int) * ((size_t)nb * PLAN_PAD + 2 * (size_t)nb + 8);
    z.sorted_bytes = sizeof(int32_t) * (size_t)npairs * 2;  // sorted pairs | redo queue
    const int ps = pair_stride(h, npairs, k);
    z.pair_d_bytes = h->tsize * (size_t)npairs * ps * (ps != k ? 2 : 1);  // candidate rows: [ps distances][ps positions]
    z.pair_pos_bytes = ps != k ? 16 : sizeof(uint32_t) * (size_t)npairs * ps;
    z.pair_cnt_bytes = sizeof(int32_t) * (size_t)npairs;
    z.thr_bytes = 8 * (size_t)nq;
    z.items_bytes = sizeof(int4) * (size_t)(std::min<int64_t>(npairs, npairs / QG + h->cfg.kc) + 1);
    return z;
}

cudaError_t launch_search(ivfadc_index* h, const void* dQ, int64_t nq, int k, int w,
                          const int32_t* d_cells, const void* d_dc, uint64_t* d_ids, void* d_dists,
                          uint64_t* d_keys, int32_t* d_counts, uint64_t* d_scanned, cudaStream_t s,
                          int* launches) {
    if (h->cfg.dtype == IVFADC_F32)
        return search_t<float>(h, dQ, nq, k, w, d_cells, d_dc, d_ids, d_dists, d_keys, d_counts, d_scanned, s, launches);
    return search_t<double>(h, dQ, nq, k, w, d_cells, d_dc, d_ids, d_dists, d_keys, d_counts, d_scanned, s, launches);
}

cudaError_t launch_merge_parts(const ivfadc_index* h, int parts, int64_t nq, int k,
                               const uint64_t* d_ids_in, const void* d_dists_in,
                               const uint64_t* d_keys_in, uint64_t* d_ids, void* d_dists,
                               int32_t* d_counts, cudaStream_t s, int* launches) {
    const unsigned mgrid = (unsigned)((nq + 3) / 4);
    if (parts * k <= MPR_MAX && !(h->cfg.flags & IVFADC_FLAG_TEST_MERGE_SWEEP)) {   // rank by binary searches (no load per round)
        if (h->cfg.dtype == IVFADC_F32)
            merge_parts_rank_kernel<float><<<mgrid, 128, 0, s>>>(parts, nq, k, d_ids_in, static_cast<const float*>(d_dists_in),
                                                                 d_keys_in, d_ids, static_cast<float*>(d_dists), d_counts);
        else
            merge_parts_rank_kernel<double><<<mgrid, 128, 0, s>>>(parts, nq, k, d_ids_in, static_cast<const double*>(d_dists_in),
                                                                  d_keys_in, d_ids, static_cast<double*>(d_dists), d_counts);
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    if (h->cfg.dtype == IVFADC_F32)
        merge_parts_kernel<float><<<mgrid, 128, 0, s>>>(parts, nq, k, d_ids_in, static_cast<const float*>(d_dists_in),
                                                        d_keys_in, d_ids, static_cast<float*>(d_dists), d_counts);
    else
        merge_parts_kernel<double><<<mgrid, 128, 0, s>>>(parts, nq, k, d_ids_in, static_cast<const double*>(d_dists_in),
                                                         d_keys_in, d_ids, static_cast<double*>(d_dists), d_counts);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace ivf
