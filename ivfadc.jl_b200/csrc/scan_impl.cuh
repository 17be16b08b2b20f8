// K2 + K3 -- the search hot loop of knn_search (reference src/index.jl:228-255), fused:
// per work item (one inverted list x up to QN queries that probe it)
//   K2  build the QN lookup tables  lut[i][code][j] = sum_d (w_icd - r_jd)^2  in shared memory
//       (src/index.jl:232-236; table 0 additionally carries dc: entry = dc + l_1, the first
//       addition of the reference's chain  d = dc; d += l_1; d += l_2 ...  src/index.jl:242-246),
//   K3  stream the list's uint8 PQ codes once (128-bit loads for m = 16), gather-add the QN
//       interleaved table entries per code byte with ONE vector shared-memory load, and keep the k
//       smallest (distance, position) per query in warp-distributed register lists,
// then merge the CTA's warps and publish the per-(query, probe) candidates.  A second kernel
// merges each query's w candidate lists in the reference's order (distance, probe rank, position)
// and looks up the ids (src/index.jl:247-257).
//
// Every distance is the same sequential fp chain as the oracle (A1/A3), so results are
// bit-identical; thresholds (per-warp, per-CTA, per-query across CTAs) only ever discard
// candidates that provably cannot enter the top k, so results do not depend on scheduling.
#pragma once

#include "common.cuh"
#include "warp_topk.cuh"

namespace ivf {

constexpr int STHREADS = 256;
constexpr int SWARPS = STHREADS / 32;

template <typename T> struct ScanArgs {
    // quantizers
    const T* Q;          // [nq][D]
    const T* C;          // [kc][D]
    const T* cb;         // [m][ksub][dsub]
    const uint8_t* cb_codes;
    int cb_identity;
    int metric;          // Dc of the lookup tables: 0 SqEuclidean (fast paths), 1 Euclidean, 2 Cityblock, 3 CosineDist
    int D, m, dsub, ksub, kc, w, k;
    // lists
    const int64_t* list_off;
    const int64_t* list_len;
    const uint8_t* codes;
    // plan
    const int32_t* cells;       // [nq][w]
    const T* dc;                // [nq][w]
    const int* bucket_off;      // [2kc+1] pairs
    const int* group_off;       // [2kc+1] work items
    int nb;                     // number of buckets (2kc: probe rank 0 / ranks >= 1 per list)
    const int32_t* sorted_pairs;
    // outputs per pair
    T* pair_d;                  // [npairs][k]
    uint32_t* pair_pos;         // [npairs][k]
    int32_t* pair_cnt;          // [npairs]
    int pstride;                // row stride of pair_d / pair_pos (>= k)
    typename Limits<T>::bits_t* thr;  // [nq] running inclusive bound on the k-th distance
};

// ---- 16-byte vector access to QN interleaved table entries -------------------------------------
template <typename T, int QN> struct LutVec;
template <> struct LutVec<float, 4> {
    __device__ static __forceinline__ void ld(const float* p, float (&o)[4]) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    }
    __device__ static __forceinline__ void st(float* p, const float (&o)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
    }
};
template <> struct LutVec<float, 2> {
    __device__ static __forceinline__ void ld(const float* p, float (&o)[2]) {
        const float2 t = *reinterpret_cast<const float2*>(p);
        o[0] = t.x; o[1] = t.y;
    }
    __device__ static __forceinline__ void st(float* p, const float (&o)[2]) {
        *reinterpret_cast<float2*>(p) = make_float2(o[0], o[1]);
    }
};
template <> struct LutVec<float, 1> {
    __device__ static __forceinline__ void ld(const float* p, float (&o)[1]) { o[0] = *p; }
    __device__ static __forceinline__ void st(float* p, const float (&o)[1]) { *p = o[0]; }
};
template <> struct LutVec<double, 4> {
    __device__ static __forceinline__ void ld(const double* p, double (&o)[4]) {
        const double2 a = *reinterpret_cast<const double2*>(p);
        const double2 b = *reinterpret_cast<const double2*>(p + 2);
        o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
    }
    __device__ static __forceinline__ void st(double* p, const double (&o)[4]) {
        *reinterpret_cast<double2*>(p) = make_double2(o[0], o[1]);
        *reinterpret_cast<double2*>(p + 2) = make_double2(o[2], o[3]);
    }
};
template <> struct LutVec<double, 2> {
    __device__ static __forceinline__ void ld(const double* p, double (&o)[2]) {
        const double2 t = *reinterpret_cast<const double2*>(p);
        o[0] = t.x; o[1] = t.y;
    }
    __device__ static __forceinline__ void st(double* p, const double (&o)[2]) {
        *reinterpret_cast<double2*>(p) = make_double2(o[0], o[1]);
    }
};
template <> struct LutVec<double, 1> {
    __device__ static __forceinline__ void ld(const double* p, double (&o)[1]) { o[0] = *p; }
    __device__ static __forceinline__ void st(double* p, const double (&o)[1]) { *p = o[0]; }
};

// ---- K2: lookup-table construction -------------------------------------------------------------
// DSUB > 0: codeword held in registers (vector global loads when 16-byte multiples), DSUB == 0:
// run-time sub-dimension.
template <typename T, int QN, int DSUB>
__device__ __forceinline__ void build_lut(const ScanArgs<T>& a, T* lut, const T* resid, const T* s_dc,
                                          int Dp) {
    const int dsub = DSUB > 0 ? DSUB : a.dsub;
    const int entries = a.m * a.ksub;
    for (int e = threadIdx.x; e < entries; e += STHREADS) {
        const int i = e / a.ksub;
        const int c = e - i * a.ksub;
        const T* wv = a.cb + (size_t)e * dsub;
        T s[QN];
#pragma unroll
        for (int j = 0; j < QN; ++j) s[j] = (T)0;
        if (a.metric != 0) {  // other metrics of Distances.jl: colwise(Dc, codeword, residual slice), generic chains
#pragma unroll
            for (int j = 0; j < QN; ++j) s[j] = metric_dist<T>(a.metric, wv, resid + j * Dp + i * dsub, dsub);
        } else if constexpr (DSUB > 0) {
            T wreg[DSUB > 0 ? DSUB : 1];
            constexpr int VEC = 16 / sizeof(T);
            if constexpr (DSUB % VEC == 0) {
#pragma unroll
                for (int d = 0; d < DSUB; d += VEC) {
                    if constexpr (sizeof(T) == 4) {
                        const float4 t = __ldg(reinterpret_cast<const float4*>(wv + d));
                        wreg[d] = t.x; wreg[d + 1] = t.y; wreg[d + 2] = t.z; wreg[d + 3] = t.w;
                    } else {
                        const double2 t = __ldg(reinterpret_cast<const double2*>(wv + d));
                        wreg[d] = t.x; wreg[d + 1] = t.y;
                    }
                }
            } else {
#pragma unroll
                for (int d = 0; d < DSUB; ++d) wreg[d] = __ldg(wv + d);
            }
#pragma unroll
            for (int j = 0; j < QN; ++j) {
                const T* r = resid + j * Dp + i * DSUB;
#pragma unroll
                for (int d = 0; d < DSUB; ++d) {
                    const T diff = sub_rn(wreg[d], r[d]);  // oracle A1: codeword - residual
                    s[j] = fma_rn(diff, diff, s[j]);
                }
            }
        } else {
            for (int d = 0; d < dsub; ++d) {
                const T wd = __ldg(wv + d);
#pragma unroll
                for (int j = 0; j < QN; ++j) {
                    const T diff = sub_rn(wd, resid[j * Dp + i * dsub + d]);
                    s[j] = fma_rn(diff, diff, s[j]);
                }
            }
        }
        if (i == 0) {
#pragma unroll
            for (int j = 0; j < QN; ++j) s[j] = add_rn(s_dc[j], s[j]);  // d = dc; d += l_1
        }
        const int code = a.cb_identity ? c : (int)a.cb_codes[e];
        LutVec<T, QN>::st(lut + ((size_t)i * 256 + code) * QN, s);
    }
}

// ---- K3 inner step: distances of one database vector to the QN queries --------------------------
template <typename T, int QN, int MC> struct CodeScan;

// run-time m: byte loads
template <typename T, int QN> struct CodeScan<T, QN, 0> {
    __device__ static __forceinline__ void run(const T* lut, const uint8_t* cp, int m, T (&acc)[QN]) {
        LutVec<T, QN>::ld(lut + (uint32_t)cp[0] * (uint32_t)QN, acc);
        for (int i = 1; i < m; ++i) {
            T v[QN];
            LutVec<T, QN>::ld(lut + ((uint32_t)i * 256u + cp[i]) * (uint32_t)QN, v);
#pragma unroll
            for (int j = 0; j < QN; ++j) acc[j] = add_rn(acc[j], v[j]);
        }
    }
};

template <typename T, int QN, int NW>
__device__ __forceinline__ void scan_words(const T* lut, const uint32_t (&wd)[NW], T (&acc)[QN]) {
#pragma unroll
    for (int i = 0; i < NW * 4; ++i) {
        const uint32_t code = (wd[i >> 2] >> (8 * (i & 3))) & 0xffu;
        if (i == 0) {
            LutVec<T, QN>::ld(lut + code * (uint32_t)QN, acc);
        } else {
            T v[QN];
            LutVec<T, QN>::ld(lut + ((uint32_t)i * 256u + code) * (uint32_t)QN, v);
#pragma unroll
            for (int j = 0; j < QN; ++j) acc[j] = add_rn(acc[j], v[j]);  // oracle A3, in order
        }
    }
}

// compile-time m, multiple of 4: widest aligned vector load the stride allows
template <typename T, int QN, int MC> struct CodeScan {
    static_assert(MC % 4 == 0 && MC > 0, "MC must be a multiple of 4");
    __device__ static __forceinline__ void run(const T* lut, const uint8_t* cp, int, T (&acc)[QN]) {
        uint32_t wd[MC / 4];
        if constexpr (MC % 16 == 0) {
#pragma unroll
            for (int x = 0; x < MC / 16; ++x) {
                const uint4 t = __ldg(reinterpret_cast<const uint4*>(cp) + x);
                wd[4 * x] = t.x; wd[4 * x + 1] = t.y; wd[4 * x + 2] = t.z; wd[4 * x + 3] = t.w;
            }
        } else if constexpr (MC % 8 == 0) {
#pragma unroll
            for (int x = 0; x < MC / 8; ++x) {
                const uint2 t = __ldg(reinterpret_cast<const uint2*>(cp) + x);
                wd[2 * x] = t.x; wd[2 * x + 1] = t.y;
            }
        } else {
#pragma unroll
            for (int x = 0; x < MC / 4; ++x) wd[x] = __ldg(reinterpret_cast<const uint32_t*>(cp) + x);
        }
        scan_words<T, QN, MC / 4>(lut, wd, acc);
    }
};

// One work item: list `cell` x the nj (<= QN) pairs pairs[0..nj).  Called by every thread of the CTA.
template <typename T, int QN, int MC, int R>
__device__ __forceinline__ void scan_item(const ScanArgs<T>& a, const int cell, const int32_t* pairs,
                                          const int nj, unsigned char* smem_raw) {
    typedef typename Limits<T>::bits_t bits_t;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int wid = tid >> 5;
    const int m = MC > 0 ? MC : a.m;
    const int k = a.k;

    const int Dp = m * a.dsub;
    // shared memory carve-up: 128-byte header | tables | residuals ; the merge area later
    // overlays tables + residuals, never the header
    T* s_dc = reinterpret_cast<T*>(smem_raw);                                   // [4]
    bits_t* s_thr = reinterpret_cast<bits_t*>(smem_raw + 32);                   // [4] inclusive bound
    int* s_pair = reinterpret_cast<int*>(smem_raw + 64);                        // [4]
    T* lut = reinterpret_cast<T*>(smem_raw + 128);
    T* resid = lut + (size_t)m * 256 * QN;

    if (tid < QN) {
        const int p = tid < nj ? pairs[tid] : -1;
        s_pair[tid] = p;
        s_dc[tid] = p >= 0 ? a.dc[p] : (T)0;
        // latest published bound of this query (other CTAs update it with atomicMin)
        s_thr[tid] = p >= 0 ? __ldcg(a.thr + p / a.w) : to_bits(Limits<T>::inf());
    }
    __syncthreads();

    // residuals r_j = q_j - centroid   (reference _closest_cluster_residuals,
    // src/coarsequantizers.jl:40-45); only the m*dsub dims the PQ covers are needed
    for (int idx = tid; idx < QN * Dp; idx += STHREADS) {
        const int j = idx / Dp, d = idx - j * Dp;
        const int p = s_pair[j];
        resid[idx] = p >= 0 ? sub_rn(a.Q[(size_t)(p / a.w) * a.D + d], a.C[(size_t)cell * a.D + d])
                            : (T)0;
    }
    __syncthreads();

    switch (a.dsub) {
        case 4: build_lut<T, QN, 4>(a, lut, resid, s_dc, Dp); break;
        case 8: build_lut<T, QN, 8>(a, lut, resid, s_dc, Dp); break;
        case 16: build_lut<T, QN, 16>(a, lut, resid, s_dc, Dp); break;
        default: build_lut<T, QN, 0>(a, lut, resid, s_dc, Dp); break;
    }
    __syncthreads();

    // ---- K3: scan ----
    const int64_t len = a.list_len[cell];
    const uint8_t* codes = a.codes + (size_t)a.list_off[cell] * m;

    WarpList<T, uint32_t, R> lst[QN];
    T lim[QN];  // exclusive acceptance bound of this warp
#pragma unroll
    for (int j = 0; j < QN; ++j) {
        lst[j].init(kNoPos);
        lim[j] = next_up_nonneg(from_bits(s_thr[j]));
    }

    for (int64_t base = (int64_t)wid * 32; base < len; base += STHREADS) {
        const int64_t p = base + lane;
        T acc[QN];
        if (p < len) {
            CodeScan<T, QN, MC>::run(lut, codes + (size_t)p * m, m, acc);
        } else {
#pragma unroll
            for (int j = 0; j < QN; ++j) acc[j] = Limits<T>::inf();
        }
#pragma unroll
        for (int j = 0; j < QN; ++j) {
            // bound published by the other warps of this CTA (inclusive -> exclusive)
            lim[j] = min(lim[j], next_up_nonneg(from_bits(*(volatile bits_t*)(s_thr + j))));
            unsigned mask = __ballot_sync(0xffffffffu, acc[j] < lim[j]);
            if (mask) {
                do {
                    const int src = __ffs(mask) - 1;
                    const T nv = __shfl_sync(0xffffffffu, acc[j], src);
                    // lanes are visited in ascending position: stable insertion == scan order
                    lst[j].template insert<false>(nv, (uint32_t)(base + src));
                    const T kv = lst[j].value_at(k - 1);
                    lim[j] = min(lim[j], kv);
                    const unsigned done = (2u << src) - 1u;
                    mask = __ballot_sync(0xffffffffu, acc[j] < lim[j]) & ~done;
                } while (mask);
                const T kv = lst[j].value_at(k - 1);
                if (lane == 0 && kv < Limits<T>::inf()) atomicMin(&s_thr[j], to_bits(kv));
            }
        }
    }
    __syncthreads();  // every warp is done with the tables: reuse them as the merge area

    T* mg_v = lut;                                                              // [SWARPS][QN][k]
    uint32_t* mg_p = reinterpret_cast<uint32_t*>(mg_v + (size_t)SWARPS * QN * k);
#pragma unroll
    for (int j = 0; j < QN; ++j)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int e = lane * R + r;
            if (e < k) {
                mg_v[((size_t)wid * QN + j) * k + e] = lst[j].v[r];
                mg_p[((size_t)wid * QN + j) * k + e] = lst[j].p[r];
            }
        }
    __syncthreads();

    // warp j merges query j's SWARPS lists by (distance, position)
    for (int j = wid; j < nj; j += SWARPS) {
        WarpList<T, uint32_t, R> fin;
        fin.init(kNoPos);
        T kv = Limits<T>::inf();
        uint32_t kp = kNoPos;
        const int total = SWARPS * k;
        for (int c0 = 0; c0 < total; c0 += 32) {
            const int c = c0 + lane;
            T val = Limits<T>::inf();
            uint32_t pos = kNoPos;
            if (c < total) {
                const int wsrc = c / k, e = c - wsrc * k;
                val = mg_v[((size_t)wsrc * QN + j) * k + e];
                pos = mg_p[((size_t)wsrc * QN + j) * k + e];
            }
            unsigned mask = __ballot_sync(0xffffffffu, pos != kNoPos && (val < kv || (val == kv && pos < kp)));
            while (mask) {
                const int src = __ffs(mask) - 1;
                const T nv = __shfl_sync(0xffffffffu, val, src);
                const uint32_t np = __shfl_sync(0xffffffffu, pos, src);
                fin.template insert<true>(nv, np);
                kv = fin.value_at(k - 1);
                kp = fin.payload_at(k - 1);
                const unsigned done = (2u << src) - 1u;
                mask = __ballot_sync(0xffffffffu, pos != kNoPos && (val < kv || (val == kv && pos < kp))) & ~done;
            }
        }
        const int pair = s_pair[j];
        int mine = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int e = lane * R + r;
            if (e < k) {
                a.pair_d[(size_t)pair * a.pstride + e] = fin.v[r];
                a.pair_pos[(size_t)pair * a.pstride + e] = fin.p[r];
                mine += fin.p[r] != kNoPos;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
        if (lane == 0) {
            a.pair_cnt[pair] = mine;
            // a full list bounds the query's k-th distance for every CTA that starts later
            if (mine == k) atomicMin(a.thr + pair / a.w, to_bits(kv));
        }
    }
}


template <typename T, int QN, int MC, int R>
__global__ void __launch_bounds__(STHREADS)
scan_kernel(const ScanArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nb = a.nb;
    // ---- work item -> (bucket, group) : last bucket with group_off[b] <= item ----
    const int item = blockIdx.x;
    if (item >= a.group_off[nb]) return;
    int lo = 0, hi = nb;  // invariant: group_off[lo] <= item < group_off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (a.group_off[mid] <= item) lo = mid; else hi = mid;
    }
    const int b = lo;
    const int cell = b >= a.kc ? b - a.kc : b;
    const int g = item - a.group_off[b];
    const int first = a.bucket_off[b] + g * QN;
    const int nj = min(QN, a.bucket_off[b + 1] - first);
    scan_item<T, QN, MC, R>(a, cell, a.sorted_pairs + first, nj, smem_raw);
}

// The (query, list) pairs the query-per-lane kernel could not finish (candidate overflow under
// heavy distance ties): a persistent grid walks the redo queue, one pair per work item.
template <typename T, int MC, int R>
__global__ void __launch_bounds__(STHREADS)
scan_redo_kernel(const ScanArgs<T> a, const int32_t* __restrict__ redo_pairs,
                 const int* __restrict__ redo_cnt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = *redo_cnt;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const int pair = redo_pairs[i];
        scan_item<T, 1, MC, R>(a, a.cells[pair], redo_pairs + i, 1, smem_raw);
        __syncthreads();
    }
}

template <typename T, int QN>
size_t scan_smem_bytes(int m, int dsub, int k) {
    const size_t lut = (size_t)m * 256 * QN * sizeof(T);
    const size_t resid = (size_t)QN * m * dsub * sizeof(T);
    const size_t merge = (size_t)SWARPS * QN * k * (sizeof(T) + sizeof(uint32_t));
    const size_t a = lut + resid;
    return 128 + (a > merge ? a : merge) + 16;
}

}  // namespace ivf
