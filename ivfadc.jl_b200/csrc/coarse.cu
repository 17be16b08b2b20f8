// K1 -- coarse assignment: for every query the w nearest centroids in ascending
// (distance, cell) order.  Replaces coarse_search(::NaiveQuantizer, point, w)
// (reference src/coarsequantizers.jl:33-37: colwise distances, stable sortperm, first w) and,
// as a documented deviation, the approximate HNSW variant (src/coarsequantizers.jl:73-76).
//
// Distances are the DIRECT form sum_d (c_d - q_d)^2 evaluated as one sequential fma chain per
// (query, centroid) pair -- bit-identical to the oracle (A1).  Three kernels, same bits:
//   * coarse_tc.cuh (default where the shape allows): tcgen05 TF32 scores prune the centroids to a
//     provable superset of the exact top-w, a second kernel re-ranks the survivors with the chain;
//   * coarse2_kernel: packed-FP32 (FADD2 / FFMA2) on transposed centroids; also the redo pass of the
//     tensor-core kernel (queries whose candidate slots overflowed);
//   * coarse_kernel: scalar FFMA, any T / D; register-tiled 2 queries x 4 centroids per thread from padded
//     shared-memory tiles (a 2 x 8 tile, TC = 128, was measured slower on config B: 0.335 vs 0.277 ms).
// The FFMA kernels fuse the top-w selection: after each 64-centroid tile every warp updates the
// warp-distributed sorted lists of its queries.
#include "common.cuh"
#include "warp_topk.cuh"
#include "coarse_tc.cuh"

#include <algorithm>

namespace ivf {

namespace {

constexpr int TQ = 32;        // queries per CTA
constexpr int TC = 64;        // centroids per tile
constexpr int TCJ = TC / 16;  // centroids per thread
constexpr int CTHREADS = 256;

template <typename T> struct CoarseCfg;
template <> struct CoarseCfg<float> {
    static constexpr int DK = 64;   // dims per shared-memory chunk
    static constexpr int VEC = 4;   // elements per 16-byte shared load
};
template <> struct CoarseCfg<double> {
    static constexpr int DK = 32;
    static constexpr int VEC = 2;
};

__device__ __forceinline__ void ld16(const float* p, float (&o)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
}
__device__ __forceinline__ void ld16(const double* p, double (&o)[2]) {
    const double2 t = *reinterpret_cast<const double2*>(p);
    o[0] = t.x; o[1] = t.y;
}

template <typename T, int R>
__global__ void __launch_bounds__(CTHREADS)
coarse_kernel(const T* __restrict__ Q, const T* __restrict__ C, int64_t nq, int kc, int D, int w,
              int32_t* __restrict__ cells_out, T* __restrict__ dc_out) {
    constexpr int DK = CoarseCfg<T>::DK;
    constexpr int VEC = CoarseCfg<T>::VEC;
    constexpr int LD = DK + VEC;  // row stride: +16 bytes => conflict-free 128-bit row reads
    constexpr int LDD = TC + 1;   // distance tile stride

    __shared__ __align__(16) T sQ[TQ * LD];
    __shared__ __align__(16) T sC[TC * LD];  // also reused as the TQ x TC distance tile
    static_assert(TQ * LDD <= TC * LD, "distance tile must fit in the centroid tile");
    T* sDist = sC;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int wid = tid >> 5;
    const int tx = tid & 15;   // centroid lane: centroids tx + 16*j
    const int ty = tid >> 4;   // query pair:   queries 2*ty, 2*ty + 1
    const int64_t q0 = (int64_t)blockIdx.x * TQ;

    // per-warp selection state: 4 queries per warp
    WarpList<T, int, R> lst[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) lst[a].init(0x7fffffff);
    T kth[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) kth[a] = Limits<T>::inf();

    for (int c0 = 0; c0 < kc; c0 += TC) {
        T acc[2][TCJ];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int j = 0; j < TCJ; ++j) acc[a][j] = (T)0;

        for (int d0 = 0; d0 < D; d0 += DK) {
            __syncthreads();  // previous chunk / distance tile fully consumed
            for (int idx = tid; idx < TQ * DK; idx += CTHREADS) {
                const int row = idx / DK, col = idx % DK;
                const int64_t q = q0 + row;
                const int d = d0 + col;
                sQ[row * LD + col] = (q < nq && d < D) ? Q[q * D + d] : (T)0;
            }
            for (int idx = tid; idx < TC * DK; idx += CTHREADS) {
                const int row = idx / DK, col = idx % DK;
                const int c = c0 + row;
                const int d = d0 + col;
                sC[row * LD + col] = (c < kc && d < D) ? C[(int64_t)c * D + d] : (T)0;
            }
            __syncthreads();
#pragma unroll 4
            for (int d = 0; d < DK; d += VEC) {
                T qv[2][VEC], cv[TCJ][VEC];
#pragma unroll
                for (int a = 0; a < 2; ++a) ld16(&sQ[(ty * 2 + a) * LD + d], qv[a]);
#pragma unroll
                for (int j = 0; j < TCJ; ++j) ld16(&sC[(tx + 16 * j) * LD + d], cv[j]);
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int j = 0; j < TCJ; ++j)
#pragma unroll
                        for (int e = 0; e < VEC; ++e) {
                            const T diff = sub_rn(cv[j][e], qv[a][e]);  // oracle A1: a[i] - b[i]
                            acc[a][j] = fma_rn(diff, diff, acc[a][j]);
                        }
            }
        }
        __syncthreads();  // all reads of sC done before it becomes the distance tile
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int j = 0; j < TCJ; ++j) {
                const int c = c0 + tx + 16 * j;
                sDist[(ty * 2 + a) * LDD + tx + 16 * j] = c < kc ? acc[a][j] : Limits<T>::inf();
            }
        __syncthreads();

        // fused selection: warp `wid` owns queries 4*wid .. 4*wid+3; candidates are offered in
        // ascending cell order, so the stable (value-only) insertion reproduces sortperm's ties.
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int ql = wid * 4 + a;
#pragma unroll
            for (int half = 0; half < TC / 32; ++half) {
                const T val = sDist[ql * LDD + lane + 32 * half];
                const int cidx = c0 + lane + 32 * half;
                unsigned mask = __ballot_sync(0xffffffffu, val < kth[a]);
                while (mask) {
                    const int src = __ffs(mask) - 1;
                    const T nv = __shfl_sync(0xffffffffu, val, src);
                    const int np = __shfl_sync(0xffffffffu, cidx, src);
                    lst[a].template insert<false>(nv, np);
                    kth[a] = lst[a].value_at(w - 1);
                    const unsigned done = (2u << src) - 1u;  // lanes <= src
                    mask = __ballot_sync(0xffffffffu, val < kth[a]) & ~done;
                }
            }
        }
    }

#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int64_t q = q0 + wid * 4 + a;
        if (q >= nq) continue;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int e = lane * R + r;
            if (e < w) {
                cells_out[q * w + e] = lst[a].p[r];
                dc_out[q * w + e] = lst[a].v[r];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// fp32 fast path (D <= 128): packed FP32 arithmetic (add.rn.f32x2 / fma.rn.f32x2, sm_100).
//   * same arithmetic as coarse_kernel -- diff = c - q, acc = fma(diff, diff, acc) in ascending d,
//     one sequential chain per (query, centroid) pair: bit-identical to the oracle (A1); the two halves
//     of a packed operation are two different centroids, never two steps of one chain;
//   * operands arrive as ready-made register pairs: centroids are kept TRANSPOSED in HBM
//     (Ct[d][kc], once at create), so a tile row holds 64 consecutive centroids of one dim and one
//     128-bit shared load yields two centroid pairs; queries are staged once per CTA (all D dims)
//     as negated duplicates {-q, -q}, so c - q is one packed add with no register shuffling;
//   * 4 queries x 4 centroids per thread: 3 LDS.128 feed 16 packed instructions per dim;
//   * TQ (queries per CTA: 16 / 24 / 32) is picked per batch so that the query blocks fill the SMs'
//     resident-CTA slots evenly (10 000 queries: TQ = 24 -> 417 blocks on 444 slots instead of 313).
// Selection as in coarse_kernel (distance tile -> warp-distributed sorted lists, ascending cell order).
// ---------------------------------------------------------------------------------------------
constexpr int PC = 64;               // centroids per tile
constexpr int PLDC = PC + 4;         // tile row stride in floats (16-byte aligned rows, conflict-free LDS.128)
constexpr int PMAXD = 256;

__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

constexpr int PDK = 64;              // dims per centroid chunk (double-buffered through cp.async)

template <int NTY, int R>
__global__ void __launch_bounds__(16 * NTY)
coarse2_kernel(const float* __restrict__ Q, const float* __restrict__ Ct, int64_t nq, int kc, int kcp, int D, int w,
               int32_t* __restrict__ cells_out, float* __restrict__ dc_out, const int32_t* __restrict__ redo_list,
               const int* __restrict__ redo_count, int list_skip) {
    constexpr int TQ2 = 4 * NTY, NT = 16 * NTY, NW = NT / 32, QPW = TQ2 / NW;  // 8 queries per warp
    constexpr int LDQ = 2 * TQ2 + 4;  // floats per dim row of the duplicated queries (16-byte aligned rows)
    constexpr int LDD = PC + 1;
    extern __shared__ __align__(16) float smem_c[];
    float* sQ = smem_c;                                  // [D][LDQ]: {-q, -q} pairs
    float* sC = smem_c + (size_t)D * LDQ;                // 2 x [PDK][PLDC]: centroid chunks, double-buffered
    float* sDist = sC + 2 * PDK * PLDC;                  // [TQ2][LDD] distance tile

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tx = tid & 15;   // centroids 4 tx .. 4 tx + 3 of the tile
    const int ty = tid >> 4;   // queries 4 ty .. 4 ty + 3 of the block
    const int64_t q0 = (int64_t)blockIdx.x * TQ2;
    // second pass behind the tensor-core kernel: only the queries it flagged (normally none), as a compacted list
    // -- the cost is proportional to their number, blocks beyond the list exit at once
    if (redo_list) {   // the first list_skip flagged queries belong to coarse_redo_small_kernel
        nq = max(0, *redo_count - list_skip);
        if (q0 >= nq) return;
        redo_list += list_skip;
    }
    const int nck = (D + PDK - 1) / PDK;                 // chunks per tile
    const int ntile = (kc + PC - 1) / PC;
    const int nchunks = ntile * nck;

    // chunk k = dims [PDK (k % nck), +PDK) of centroids [PC (k / nck), +PC): asynchronous copy, 16 bytes per request
    auto prefetch = [&](int k) {
        const int c0 = (k / nck) * PC, d0 = (k % nck) * PDK;
        float* dst = sC + (k & 1) * (PDK * PLDC);
        for (int idx = tid; idx < PDK * (PC / 4); idx += NT) {
            const int dd = idx >> 4, c4 = idx & 15;
            const uint32_t sa = (uint32_t)__cvta_generic_to_shared(dst + dd * PLDC + 4 * c4);
            if (d0 + dd < D)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(Ct + (size_t)(d0 + dd) * kcp + c0 + 4 * c4) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    prefetch(0);

    // queries: coalesced along d, stored transposed, negated and duplicated
    for (int idx = tid; idx < TQ2 * D; idx += NT) {
        const int row = idx / D, d = idx - row * D;
        int64_t q = q0 + row;
        if (q < nq && redo_list) q = redo_list[q];
        const float v = q0 + row < nq ? -Q[q * D + d] : 0.f;
        *reinterpret_cast<float2*>(&sQ[d * LDQ + 2 * row]) = make_float2(v, v);
    }

    WarpList<float, int, R> lst[QPW];
#pragma unroll
    for (int a = 0; a < QPW; ++a) lst[a].init(0x7fffffff);
    float kth[QPW];
#pragma unroll
    for (int a = 0; a < QPW; ++a) kth[a] = Limits<float>::inf();

    unsigned long long acc[4][2];
    for (int k = 0; k < nchunks; ++k) {
        const int ck = k % nck;
        if (ck == 0) {
#pragma unroll
            for (int a = 0; a < 4; ++a) acc[a][0] = acc[a][1] = 0ull;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();  // chunk k landed (and the queries, first time); every warp is done with chunk k - 1
        if (k + 1 < nchunks) prefetch(k + 1);  // into the buffer chunk k - 1 occupied
        const int d0 = ck * PDK;
        const int dn = min(PDK, D - d0);
        const float* pc = sC + (k & 1) * (PDK * PLDC) + 4 * tx;
        const float* pq = sQ + (size_t)d0 * LDQ + 8 * ty;
#pragma unroll 8
        for (int d = 0; d < dn; ++d) {
            const ulonglong2 cv = *reinterpret_cast<const ulonglong2*>(pc + d * PLDC);          // {c0, c1}, {c2, c3}
            const ulonglong2 qa = *reinterpret_cast<const ulonglong2*>(pq + d * LDQ);           // {-q0, -q0}, {-q1, -q1}
            const ulonglong2 qb = *reinterpret_cast<const ulonglong2*>(pq + d * LDQ + 4);       // {-q2, -q2}, {-q3, -q3}
            const unsigned long long qn[4] = {qa.x, qa.y, qb.x, qb.y};
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const unsigned long long d0p = add2(cv.x, qn[a]), d1p = add2(cv.y, qn[a]);  // oracle A1: c - q
                acc[a][0] = fma2(d0p, d0p, acc[a][0]);
                acc[a][1] = fma2(d1p, d1p, acc[a][1]);
            }
        }
        if (ck != nck - 1) continue;
        // ---- tile complete: distances -> shared tile -> fused selection ----
        const int c0 = (k / nck) * PC;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned long long pr = acc[a][j >> 1];
                const float v = __uint_as_float((j & 1) ? (unsigned)(pr >> 32) : (unsigned)pr);
                const int c = c0 + 4 * tx + j;
                sDist[(4 * ty + a) * LDD + 4 * tx + j] = c < kc ? v : Limits<float>::inf();
            }
        __syncthreads();
        // warp `wid` owns queries QPW * wid ..; candidates are offered in ascending cell order, so the
        // stable (value-only) insertion reproduces sortperm's ties.  (The next tile's first chunk is in flight.)
#pragma unroll
        for (int a = 0; a < QPW; ++a) {
            const int ql = wid * QPW + a;
#pragma unroll
            for (int half = 0; half < PC / 32; ++half) {
                const float val = sDist[ql * LDD + lane + 32 * half];
                const int cidx = c0 + lane + 32 * half;
                unsigned mask = __ballot_sync(0xffffffffu, val < kth[a]);
                while (mask) {
                    const int src = __ffs(mask) - 1;
                    const float nv = __shfl_sync(0xffffffffu, val, src);
                    const int np = __shfl_sync(0xffffffffu, cidx, src);
                    lst[a].template insert<false>(nv, np);
                    kth[a] = lst[a].value_at(w - 1);
                    const unsigned done = (2u << src) - 1u;  // lanes <= src
                    mask = __ballot_sync(0xffffffffu, val < kth[a]) & ~done;
                }
            }
        }
        // the distance tile is rewritten only after the next tile's chunks have passed their barriers
    }
#pragma unroll
    for (int a = 0; a < QPW; ++a) {
        int64_t q = q0 + wid * QPW + a;
        if (q >= nq) continue;
        if (redo_list) q = redo_list[q];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int e = lane * R + r;
            if (e < w) {
                cells_out[q * w + e] = lst[a].p[r];
                dc_out[q * w + e] = lst[a].v[r];
            }
        }
    }
}

// A handful of queries flagged by the tensor-core kernel (candidate overflow: about one in 10^4 on the benchmark
// data): one block per (query, 256 centroids) evaluates the reference's direct-form chain, the last block of a
// query to finish selects its w nearest by (distance, cell) -- the latency of a flagged query is a few
// microseconds instead of one CTA streaming all centroids (150 us at kc = 1024).  Queries beyond RS_MAXQ go to
// coarse2_kernel with the compacted list.
__global__ void __launch_bounds__(256)
coarse_redo_small_kernel(const float* __restrict__ Q, const float* __restrict__ C, int kc, int kcp, int D, int w,
                         int32_t* __restrict__ cells_out, float* __restrict__ dc_out,
                         const int32_t* __restrict__ redo_list, const int* __restrict__ redo_count,
                         float* __restrict__ scratch, unsigned* __restrict__ done) {
    const int slot = blockIdx.y;
    if (slot >= min(*redo_count, ctc::RS_MAXQ)) return;
    const int64_t q = redo_list[slot];
    __shared__ float sq[PMAXD];
    __shared__ int s_last;
    const int tid = threadIdx.x;
    for (int d = tid; d < D; d += 256) sq[d] = Q[q * D + d];
    __syncthreads();
    const int c = blockIdx.x * 256 + tid;
    float* row = scratch + (size_t)slot * kcp;
    if (c < kcp) {
        float acc = Limits<float>::inf();
        if (c < kc) {
            acc = 0.f;
            // the row as 128-bit loads, eight in flight (a scalar loop pays one L2 round trip per element: 26 us per
            // flagged query); the chain itself stays sequential -- oracle A1: c - q, one fma chain
            const float4* cr = reinterpret_cast<const float4*>(C + (size_t)c * D);   // D % 16 == 0 on this path
            for (int d4 = 0; d4 < (D >> 2); d4 += 8) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = d4 + u < (D >> 2) ? __ldg(cr + d4 + u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (d4 + u < (D >> 2)) {
                        const float* qd = sq + 4 * (d4 + u);
                        float diff = sub_rn(v[u].x, qd[0]); acc = fma_rn(diff, diff, acc);
                        diff = sub_rn(v[u].y, qd[1]); acc = fma_rn(diff, diff, acc);
                        diff = sub_rn(v[u].z, qd[2]); acc = fma_rn(diff, diff, acc);
                        diff = sub_rn(v[u].w, qd[3]); acc = fma_rn(diff, diff, acc);
                    }
                }
            }
        }
        __stcg(row + c, acc);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(done + slot, 1u) == gridDim.x - 1 ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // Selection without a round per result: keys = (distance bits, cell) are distinct, so the w-th smallest of the
    // 256 per-thread minima bounds the w-th smallest key; the few keys within the bound are ranked by counting.
    __shared__ unsigned long long tmin[256];
    __shared__ unsigned long long cand[ctc::RS_MAXC];
    __shared__ unsigned long long s_bound;
    __shared__ int s_nc;
    unsigned long long best = ~0ull;
    for (int i = tid; i < kcp; i += 256) {
        const unsigned long long key = ((unsigned long long)__float_as_uint(__ldcg(row + i)) << 32) | (unsigned)i;
        best = key < best ? key : best;
    }
    tmin[tid] = best;
    if (tid == 0) s_nc = 0;
    __syncthreads();
    int rk = 0;
    for (int i = 0; i < 256; ++i) rk += tmin[i] < best ? 1 : 0;
    if (rk == w - 1) s_bound = best;   // w <= 32 <= 256 distinct finite minima (kc >= 256 on this path)
    __syncthreads();
    const unsigned long long bound = s_bound;
    for (int i = tid; i < kcp; i += 256) {
        const unsigned long long key = ((unsigned long long)__float_as_uint(__ldcg(row + i)) << 32) | (unsigned)i;
        if (key <= bound) cand[atomicAdd(&s_nc, 1)] = key;   // at most w keys per thread: <= RS_MAXC
    }
    __syncthreads();
    const int nc = s_nc;
    for (int i = tid; i < nc; i += 256) {
        const unsigned long long key = cand[i];
        int r = 0;
        for (int x = 0; x < nc; ++x) r += cand[x] < key ? 1 : 0;
        if (r < w) {
            cells_out[q * w + r] = (int32_t)(unsigned)key;
            dc_out[q * w + r] = __uint_as_float((unsigned)(key >> 32));
        }
    }
}

// coarse_search for the metrics beyond SqEuclidean (Euclidean, Cityblock, CosineDist; reference
// src/coarsequantizers.jl:33-37 with D = Dc): one block per query at a time -- thread per centroid evaluates the
// metric's chain (common.cuh metric_dist, the oracle's dist_colwise), the distances go to the block's scratch row,
// then w rounds of a block-wide minimum of (distance, cell) above the previous winner = sortperm order, ties to the
// lower cell.  A compatibility path: exact, not tuned.
template <typename T>
__global__ void __launch_bounds__(256)
coarse_metric_kernel(const T* __restrict__ Q, const T* __restrict__ C, int64_t nq, int kc, int D, int w, int metric,
                     int32_t* __restrict__ cells_out, T* __restrict__ dc_out, T* __restrict__ scratch) {
    extern __shared__ __align__(16) unsigned char cm_smem[];
    T* sq = reinterpret_cast<T*>(cm_smem);
    __shared__ T red_d[8];
    __shared__ int red_c[8];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    T* row = scratch + (size_t)blockIdx.x * kc;
    for (int64_t q = blockIdx.x; q < nq; q += gridDim.x) {
        __syncthreads();
        for (int d = tid; d < D; d += 256) sq[d] = Q[q * D + d];
        __syncthreads();
        for (int c = tid; c < kc; c += 256) row[c] = metric_dist<T>(metric, C + (size_t)c * D, sq, D);
        __syncthreads();
        T pd = (T)0;
        int pc = -1;
        for (int r = 0; r < w; ++r) {
            T bd = Limits<T>::inf();
            int bc = 0x7fffffff;
            for (int c = tid; c < kc; c += 256) {
                const T v = row[c];
                const bool above = pc < 0 || v > pd || (v == pd && c > pc);
                if (above && (v < bd || (v == bd && c < bc))) { bd = v; bc = c; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const T od = __shfl_xor_sync(0xffffffffu, bd, o);
                const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
                if (od < bd || (od == bd && oc < bc)) { bd = od; bc = oc; }
            }
            if (lane == 0) { red_d[wid] = bd; red_c[wid] = bc; }
            __syncthreads();
            bd = red_d[0]; bc = red_c[0];
#pragma unroll
            for (int i = 1; i < 8; ++i)
                if (red_d[i] < bd || (red_d[i] == bd && red_c[i] < bc)) { bd = red_d[i]; bc = red_c[i]; }
            __syncthreads();
            if (tid == 0) {
                cells_out[q * w + r] = bc;
                dc_out[q * w + r] = bd;
            }
            pd = bd;
            pc = bc;
        }
    }
}

__global__ void transpose_centroids_kernel(const float* __restrict__ C, int kc, int kcp, int D, float* __restrict__ Ct) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= kcp * D) return;
    const int d = idx / kcp, c = idx - d * kcp;
    Ct[idx] = c < kc ? C[(size_t)c * D + d] : 0.f;
}

template <int NTY, int R>
cudaError_t launch_coarse2_inst(const ivfadc_index* h, const float* Q, const float* Ct, int64_t nq, int kc, int kcp, int D, int w,
                                int32_t* cells, float* dc, cudaStream_t s, const int32_t* redo_list, const int* redo_count,
                                int list_skip) {
    constexpr int TQ2 = 4 * NTY;
    const size_t smem = ((size_t)D * (2 * TQ2 + 4) + 2 * (size_t)PDK * PLDC + (size_t)TQ2 * (PC + 1)) * sizeof(float);
    cudaError_t e = ensure_smem(h, reinterpret_cast<const void*>(&coarse2_kernel<NTY, R>), smem);
    if (e != cudaSuccess) return e;
    coarse2_kernel<NTY, R><<<(unsigned)((nq + TQ2 - 1) / TQ2), 16 * NTY, smem, s>>>(Q, Ct, nq, kc, kcp, D, w, cells, dc, redo_list,
                                                                                  redo_count, list_skip);
    return cudaGetLastError();
}
template <int R>
cudaError_t launch_coarse2_r(const ivfadc_index* h, int nty, const float* Q, const float* Ct, int64_t nq, int kc, int kcp, int D, int w,
                             int32_t* cells, float* dc, cudaStream_t s, const int32_t* redo_list = nullptr,
                             const int* redo_count = nullptr, int list_skip = 0) {
    if (nty == 4) return launch_coarse2_inst<4, R>(h, Q, Ct, nq, kc, kcp, D, w, cells, dc, s, redo_list, redo_count, list_skip);
    if (nty == 6) return launch_coarse2_inst<6, R>(h, Q, Ct, nq, kc, kcp, D, w, cells, dc, s, redo_list, redo_count, list_skip);
    return launch_coarse2_inst<8, R>(h, Q, Ct, nq, kc, kcp, D, w, cells, dc, s, redo_list, redo_count, list_skip);
}

template <typename T>
cudaError_t launch_coarse_t(const ivfadc_index* h, const void* dQ, int64_t nq, int w,
                            int32_t* d_cells, void* d_dc, cudaStream_t s) {
    const dim3 grid((unsigned)((nq + TQ - 1) / TQ));
    const T* Q = static_cast<const T*>(dQ);
    const T* C = static_cast<const T*>(h->d_centroids);
    T* dc = static_cast<T*>(d_dc);
    const int kc = h->cfg.kc, D = h->cfg.dim;
    if (w <= 32)
        coarse_kernel<T, 1><<<grid, CTHREADS, 0, s>>>(Q, C, nq, kc, D, w, d_cells, dc);
    else if (w <= 64)
        coarse_kernel<T, 2><<<grid, CTHREADS, 0, s>>>(Q, C, nq, kc, D, w, d_cells, dc);
    else
        coarse_kernel<T, 4><<<grid, CTHREADS, 0, s>>>(Q, C, nq, kc, D, w, d_cells, dc);
    return cudaGetLastError();
}

}  // namespace

int coarse_max_w() { return 128; }

// Once at create (fp32): the centroids transposed, Ct[d][kc_pad], kc padded to the tile width.
cudaError_t coarse_prepare(ivfadc_index* h, cudaStream_t s, int* launches) {
    if (h->cfg.dtype != IVFADC_F32) return cudaSuccess;
    h->kc_pad = (h->cfg.kc + PC - 1) / PC * PC;
    const size_t n = (size_t)h->kc_pad * h->cfg.dim;
    cudaError_t e = cudaSuccess;   // re-entrant: ivfadc_set_centroids_device refreshes the derived operands
    if (!h->d_centroids_t && (e = cudaMalloc(&h->d_centroids_t, n * sizeof(float))) != cudaSuccess) return e;
    transpose_centroids_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(static_cast<const float*>(h->d_centroids), h->cfg.kc,
                                                                        h->kc_pad, h->cfg.dim,
                                                                        static_cast<float*>(h->d_centroids_t));
    if (launches) *launches += 1;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    // operands of the tensor-core kernel (coarse_tc.cuh): TF32 centroid blocks, squared norms, max norm
    const int kc = h->cfg.kc, D = h->cfg.dim;
    if (kc >= ctc::NC && D % 16 == 0 && D <= 128) {
        h->kc_pad256 = (kc + ctc::NC - 1) / ctc::NC * ctc::NC;
        const int ksteps = D / 8;
        const size_t words = (size_t)h->kc_pad256 * (ksteps + 1) * 8;
        if (!h->d_tcC && (e = cudaMalloc(&h->d_tcC, words * sizeof(float))) != cudaSuccess) return e;
        if (!h->d_ccn && (e = cudaMalloc(&h->d_ccn, ((size_t)h->kc_pad256 + 4) * sizeof(float))) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(h->d_ccn, 0, ((size_t)h->kc_pad256 + 4) * sizeof(float), s)) != cudaSuccess) return e;
        if (!h->d_err) {
            if ((e = cudaMalloc(&h->d_err, sizeof(int))) != cudaSuccess) return e;
            if ((e = cudaMemsetAsync(h->d_err, 0, sizeof(int), s)) != cudaSuccess) return e;
        }
        const int64_t nthr = (int64_t)h->kc_pad256 * (ksteps + 1);
        ctc::prep_tcc_kernel<<<(unsigned)((nthr + 255) / 256), 256, 0, s>>>(static_cast<const float*>(h->d_centroids), kc,
                                                                           h->kc_pad256, D, ksteps,
                                                                           static_cast<float*>(h->d_tcC));
        ctc::prep_cn_kernel<<<(unsigned)((h->kc_pad256 + 255) / 256), 256, 0, s>>>(static_cast<const float*>(h->d_centroids), kc,
                                                                                  h->kc_pad256, D,
                                                                                  static_cast<float*>(h->d_ccn));
        if (launches) *launches += 2;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

namespace {
template <int WL>
cudaError_t launch_coarse3_inst(const ivfadc_index* h, const ctc::Args& a, unsigned grid, size_t smem, cudaStream_t s) {
    cudaError_t e = ensure_smem(h, reinterpret_cast<const void*>(&ctc::coarse3_kernel<WL>), smem);
    if (e != cudaSuccess) return e;
    ctc::coarse3_kernel<WL><<<grid, ctc::THREADS, smem, s>>>(a);
    return cudaGetLastError();
}
}  // namespace

cudaError_t launch_coarse(const ivfadc_index* h, const void* dQ, int64_t nq, int w, int32_t* d_cells,
                          void* d_dc, cudaStream_t s, int* launches) {
    if (nq <= 0) return cudaSuccess;
    if (launches) *launches += 1;
    h->last_redo_nq = 0;
    if (h->cfg.metric_coarse != IVFADC_SQEUCLIDEAN) {   // Euclidean / Cityblock / CosineDist: the generic exact kernel
        const int kc = h->cfg.kc, D = h->cfg.dim;
        const unsigned grid = (unsigned)std::min<int64_t>(nq, 4 * (h->num_sms > 0 ? h->num_sms : 148));
        cudaError_t e = h->ws_coarse_redo.reserve((size_t)grid * kc * h->tsize);
        if (e != cudaSuccess) return e;
        if (h->cfg.dtype == IVFADC_F32)
            coarse_metric_kernel<float><<<grid, 256, (size_t)D * 4, s>>>(static_cast<const float*>(dQ), static_cast<const float*>(h->d_centroids),
                                                                       nq, kc, D, w, h->cfg.metric_coarse, d_cells,
                                                                       static_cast<float*>(d_dc), h->ws_coarse_redo.as<float>());
        else
            coarse_metric_kernel<double><<<grid, 256, (size_t)D * 8, s>>>(static_cast<const double*>(dQ), static_cast<const double*>(h->d_centroids),
                                                                        nq, kc, D, w, h->cfg.metric_coarse, d_cells,
                                                                        static_cast<double*>(d_dc), h->ws_coarse_redo.as<double>());
        return cudaGetLastError();
    }
    if (h->cfg.dtype == IVFADC_F32 && h->d_centroids_t && h->cfg.dim <= PMAXD && (h->cfg.dim & 3) == 0 &&
        !(h->cfg.flags & IVFADC_FLAG_COARSE_SCALAR)) {
        // queries per CTA: the block count that fills the resident-CTA slots of the SMs most evenly
        const int num_sms = h->num_sms > 0 ? h->num_sms : 148;
        const int D = h->cfg.dim;
        int best = 8;
        double best_eff = -1.0;
        for (int nty : {8, 6, 4}) {
            const int tq = 4 * nty;
            const size_t smem = ((size_t)D * (2 * tq + 4) + 2 * (size_t)PDK * PLDC + (size_t)tq * (PC + 1)) * sizeof(float) + 1024;
            const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(16, (size_t)(227 * 1024) / smem));
            const int64_t blocks = (nq + tq - 1) / tq;
            const int64_t slots = (int64_t)num_sms * per_sm;
            const int64_t rounds = (blocks + slots - 1) / slots;
            // useful query slots / provisioned query slots, smaller blocks re-stream the centroids more often
            const double eff = (double)nq / ((double)rounds * slots * tq) * (nty == 8 ? 1.0 : nty == 6 ? 0.97 : 0.93);
            if (eff > best_eff) { best_eff = eff; best = nty; }
        }
        const float* Q = static_cast<const float*>(dQ);
        const float* Ct = static_cast<const float*>(h->d_centroids_t);
        float* dc = static_cast<float*>(d_dc);
        // tensor-core pruning + exact re-rank (coarse_tc.cuh); the FFMA kernel below redoes flagged queries
        const int wl = w <= 1 ? 1 : w <= 8 ? 8 : w <= 16 ? 16 : 32;  // the bound needs wl finite group minima
        if (h->d_tcC && w <= 32 && h->cfg.kc >= wl * ctc::GRP && !(h->cfg.flags & IVFADC_FLAG_COARSE_FFMA) &&
            (reinterpret_cast<uintptr_t>(dQ) & 15) == 0 &&
            ctc::smem_layout(D / 8).total <= (size_t)(227 * 1024)) {
            // redo flags [nq] | candidate counts [nq] | candidate rows [nq][CAP]
            const size_t off_cnt = ((size_t)nq + 255) / 256 * 256, off_cand = off_cnt + (size_t)nq * 4;
            const size_t off_list = off_cand + (size_t)nq * ctc::CAP * 4;   // flagged queries, compacted | their number
            const size_t off_done = off_list + (size_t)nq * 4 + 16, off_scr = off_done + ctc::RS_MAXQ * 4;
            cudaError_t e = h->ws_coarse_redo.reserve(off_scr + (size_t)ctc::RS_MAXQ * h->kc_pad256 * 4);
            if (e != cudaSuccess) return e;
            ctc::Args ca;
            ca.Q = Q; ca.C = static_cast<const float*>(h->d_centroids);
            ca.tcC = static_cast<const float*>(h->d_tcC); ca.cn = static_cast<const float*>(h->d_ccn);
            ca.nq = nq; ca.kc = h->cfg.kc; ca.kcp = h->kc_pad256; ca.D = D; ca.ksteps = D / 8; ca.w = w;
            ca.cells_out = d_cells; ca.dc_out = dc; ca.redo = h->ws_coarse_redo.as<uint8_t>(); ca.err = h->d_err;
            ca.cnt_out = reinterpret_cast<int32_t*>(ca.redo + off_cnt);
            ca.cand_out = reinterpret_cast<int32_t*>(ca.redo + off_cand);
            ca.redo_list = reinterpret_cast<int32_t*>(ca.redo + off_list);
            ca.redo_count = reinterpret_cast<int*>(ca.redo + off_list + (size_t)nq * 4);
            ca.redo_done = reinterpret_cast<unsigned*>(ca.redo + off_done);
            ca.force_redo = (h->cfg.flags & IVFADC_FLAG_TEST_COARSE_REDO) ? 1 : 0;
            const unsigned grid = (unsigned)((nq + ctc::MQ - 1) / ctc::MQ);
            const size_t smem = ctc::smem_layout(D / 8).total;
            if (w <= 1) e = launch_coarse3_inst<1>(h, ca, grid, smem, s);
            else if (w <= 8) e = launch_coarse3_inst<8>(h, ca, grid, smem, s);
            else if (w <= 16) e = launch_coarse3_inst<16>(h, ca, grid, smem, s);
            else e = launch_coarse3_inst<32>(h, ca, grid, smem, s);
            if (e != cudaSuccess) return e;
            const size_t rsmem = ctc::RR_WARPS * ctc::rr_warp_floats(D) * sizeof(float);
            if ((e = ensure_smem(h, reinterpret_cast<const void*>(&ctc::coarse3_rerank_kernel), rsmem)) != cudaSuccess) return e;
            const unsigned rr_need = (unsigned)((nq + ctc::RR_WARPS - 1) / ctc::RR_WARPS);
            const unsigned rr_slots = (unsigned)num_sms * (unsigned)std::max<size_t>(1, (size_t)(227 * 1024) / (rsmem + 1024));
            ctc::coarse3_rerank_kernel<<<std::min(rr_need, rr_slots), ctc::RR_WARPS * 32, rsmem, s>>>(ca);
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
            if (launches) *launches += 2;
            h->last_redo_nq = nq;
            // flagged queries: the first RS_MAXQ in parallel over the centroids (when the selection's candidate
            // buffer covers the worst case), any further ones in 16-query blocks
            const bool wide = (size_t)w * (size_t)(h->kc_pad256 / 256) <= (size_t)ctc::RS_MAXC;
            if (wide) {
                coarse_redo_small_kernel<<<dim3((unsigned)(h->kc_pad256 / 256), ctc::RS_MAXQ), 256, 0, s>>>(
                    Q, ca.C, h->cfg.kc, h->kc_pad256, D, w, d_cells, dc, ca.redo_list, ca.redo_count,
                    reinterpret_cast<float*>(ca.redo + off_scr), ca.redo_done);
                if ((e = cudaGetLastError()) != cudaSuccess) return e;
                if (launches) *launches += 1;
            }
            return launch_coarse2_r<1>(h, 4, Q, Ct, nq, h->cfg.kc, h->kc_pad, D, w, d_cells, dc, s, ca.redo_list, ca.redo_count,
                                       wide ? ctc::RS_MAXQ : 0);
        }
        if (w <= 32) return launch_coarse2_r<1>(h, best, Q, Ct, nq, h->cfg.kc, h->kc_pad, D, w, d_cells, dc, s);
        if (w <= 64) return launch_coarse2_r<2>(h, best, Q, Ct, nq, h->cfg.kc, h->kc_pad, D, w, d_cells, dc, s);
        return launch_coarse2_r<4>(h, best, Q, Ct, nq, h->cfg.kc, h->kc_pad, D, w, d_cells, dc, s);
    }
    if (h->cfg.dtype == IVFADC_F32) return launch_coarse_t<float>(h, dQ, nq, w, d_cells, d_dc, s);
    return launch_coarse_t<double>(h, dQ, nq, w, d_cells, d_dc, s);
}

}  // namespace ivf
