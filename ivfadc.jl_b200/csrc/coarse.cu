// K1 -- coarse assignment: for every query the w nearest centroids in ascending
// (distance, cell) order.  Replaces coarse_search(::NaiveQuantizer, point, w)
// (reference src/coarsequantizers.jl:33-37: colwise distances, stable sortperm, first w) and,
// as a documented deviation, the approximate HNSW variant (src/coarsequantizers.jl:73-76).
//
// Distances are the DIRECT form sum_d (c_d - q_d)^2 evaluated as one sequential fma chain per
// (query, centroid) pair -- bit-identical to the oracle (A1) -- on the FFMA pipe, register-tiled
// 2 queries x 4 centroids per thread from padded shared-memory tiles (a 2 x 8 tile, TC = 128, was
// measured slower on config B: 0.335 vs 0.277 ms, fewer resident CTAs for a 313-CTA grid); the top-w selection is
// fused: after each 64-centroid tile every warp updates the warp-distributed sorted lists of its
// 4 queries.  FP32 FFMA was chosen over 3xTF32 tcgen05 because selection must agree with the
// oracle bit for bit and the whole step is < 10% of the search ("K1").
#include "common.cuh"
#include "warp_topk.cuh"

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

cudaError_t launch_coarse(const ivfadc_index* h, const void* dQ, int64_t nq, int w, int32_t* d_cells,
                          void* d_dc, cudaStream_t s, int* launches) {
    if (nq <= 0) return cudaSuccess;
    if (launches) *launches += 1;
    if (h->cfg.dtype == IVFADC_F32) return launch_coarse_t<float>(h, dQ, nq, w, d_cells, d_dc, s);
    return launch_coarse_t<double>(h, dQ, nq, w, d_cells, d_dc, s);
}

}  // namespace ivf
