// K4 -- PQ residual encoding: residual = x - centroid[cell], then per codebook i the first minimum
// over the ksub codewords of the GEMM-form squared distance
//     v = max((|w|^2 + |x_i|^2) - 2 <w, x_i>, 0)
// exactly as QuantizedArrays.quantize_data -> Distances.pairwise does it (oracle A2; reference call
// sites src/index.jl:187 and src/utils.jl:158), stored byte = codebook.codes[argmin].
//
// One CTA encodes 64 vectors; per codebook the ksub x dsub codeword block is staged in shared
// memory once and shared by all 64 vectors; each thread owns one vector and a contiguous quarter of
// the codewords (first-minimum order is preserved: ascending codewords inside a thread, quarters
// combined in ascending order with strict '<').
#include <algorithm>

#include "common.cuh"

namespace ivf {

namespace {

constexpr int EV = 64;         // vectors per CTA
constexpr int ETHREADS = 256;  // 4 codeword quarters x 64 vectors

template <typename T>
__global__ void codebook_norms_kernel(const T* __restrict__ cb, int entries, int dsub, T* __restrict__ norms) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= entries) return;
    T s = (T)0;
    for (int d = 0; d < dsub; ++d) {
        const T x = cb[(size_t)e * dsub + d];
        s = fma_rn(x, x, s);
    }
    norms[e] = s;
}

__global__ void assign_to_cells_kernel(const int64_t* __restrict__ assign, int64_t n, int base,
                                       int32_t* __restrict__ cells) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cells[i] = (int32_t)(assign[i] - base);
}

__global__ void assign_check_kernel(const int64_t* __restrict__ assign, int64_t n, int base, int kc,
                                    int* __restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (uint64_t)(assign[i] - base) >= (uint64_t)kc) *bad = 1;
}

template <typename T, int DSUB>
__global__ void __launch_bounds__(ETHREADS)
encode_kernel(const T* __restrict__ X, int64_t n, const int32_t* __restrict__ cells,
              const T* __restrict__ C, const T* __restrict__ cb, const uint8_t* __restrict__ cb_codes,
              const T* __restrict__ cb_norms, int D, int m, int dsub_rt, int ksub,
              uint8_t* __restrict__ codes_out, int tiled, int metric) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int dsub = DSUB > 0 ? DSUB : dsub_rt;
    const int Dp = m * dsub;
    // residual row stride (odd => conflict-free per-lane rows).  Wide vectors (D of several hundred) do not fit
    // 64 full residual rows in shared memory: `tiled` stages one subspace slice per codebook instead.
    const int LDR = (tiled ? dsub : Dp) + 1;
    T* s_res = reinterpret_cast<T*>(smem_raw);                 // [EV][LDR]
    T* s_cw = s_res + (size_t)EV * LDR;                        // [ksub][dsub]
    T* s_nrm = s_cw + (size_t)ksub * dsub;                     // [ksub]
    T* s_bv = s_nrm + ksub;                                    // [4][EV] partial minima
    int* s_bi = reinterpret_cast<int*>(s_bv + 4 * EV);         // [4][EV]

    const int tid = threadIdx.x;
    const int v = tid & (EV - 1);
    const int cg = tid >> 6;  // codeword quarter
    const int64_t v0 = (int64_t)blockIdx.x * EV;

    // residuals (reference src/utils.jl:157 / _build_residuals src/index.jl:168-175)
    for (int idx = tid; !tiled && idx < EV * Dp; idx += ETHREADS) {
        const int row = idx / Dp, d = idx - row * Dp;
        const int64_t gv = v0 + row;
        T r = (T)0;
        if (gv < n) r = sub_rn(X[gv * D + d], C[(size_t)cells[gv] * D + d]);
        s_res[row * LDR + d] = r;
    }

    const int per = (ksub + 3) / 4;
    const int clo = min(ksub, cg * per), chi = min(ksub, clo + per);

    for (int i = 0; i < m; ++i) {
        __syncthreads();  // previous codebook fully consumed (and residuals visible)
        for (int idx = tid; idx < ksub * dsub; idx += ETHREADS)
            s_cw[idx] = cb[(size_t)i * ksub * dsub + idx];
        for (int idx = tid; idx < ksub; idx += ETHREADS) s_nrm[idx] = cb_norms[i * ksub + idx];
        if (tiled) {
            for (int idx = tid; idx < EV * dsub; idx += ETHREADS) {
                const int row = idx / dsub, d = idx - row * dsub;
                const int64_t gv = v0 + row;
                T r = (T)0;
                if (gv < n) r = sub_rn(X[gv * D + i * dsub + d], C[(size_t)cells[gv] * D + i * dsub + d]);
                s_res[row * LDR + d] = r;
            }
        }
        __syncthreads();

        const T* xr = s_res + v * LDR + (tiled ? 0 : i * dsub);
        T best = Limits<T>::inf();
        int besti = -1;
        if (metric != 0) {
            // quantization_distance beyond SqEuclidean (oracle encode_residual): pairwise(Euclidean) = sqrt of the GEMM
            // form, pairwise(Cityblock) = the direct sum per pair, pairwise(CosineDist) = 1 - dot / (|w| |x|), clamped at 0
            T sb = (T)0;
            for (int d = 0; d < dsub; ++d) sb = fma_rn(xr[d], xr[d], sb);
            for (int c = clo; c < chi; ++c) {
                const T* wv = s_cw + c * dsub;
                T val;
                if (metric == 2) {
                    val = metric_dist<T>(2, wv, xr, dsub);
                } else {
                    T dot = (T)0;
                    for (int d = 0; d < dsub; ++d) dot = fma_rn(wv[d], xr[d], dot);
                    if (metric == 1) {
                        val = sub_rn(add_rn(s_nrm[c], sb), mul_rn((T)2, dot));
                        val = sqrt_rn(val > (T)0 ? val : (T)0);
                    } else {
                        val = sub_rn((T)1, div_rn(dot, mul_rn(sqrt_rn(s_nrm[c]), sqrt_rn(sb))));
                        val = val > (T)0 ? val : (T)0;
                    }
                }
                if (besti < 0 || val < best) {
                    best = val;
                    besti = c;
                }
            }
        } else if constexpr (DSUB > 0) {
            T x[DSUB > 0 ? DSUB : 1];
            T sb = (T)0;
#pragma unroll
            for (int d = 0; d < DSUB; ++d) {
                x[d] = xr[d];
                sb = fma_rn(x[d], x[d], sb);
            }
            for (int c = clo; c < chi; ++c) {
                const T* wv = s_cw + c * DSUB;  // same address for the whole warp: broadcast
                T dot = (T)0;
#pragma unroll
                for (int d = 0; d < DSUB; ++d) dot = fma_rn(wv[d], x[d], dot);
                T val = sub_rn(add_rn(s_nrm[c], sb), mul_rn((T)2, dot));
                val = val > (T)0 ? val : (T)0;
                if (besti < 0 || val < best) {
                    best = val;
                    besti = c;
                }
            }
        } else {
            T sb = (T)0;
            for (int d = 0; d < dsub; ++d) sb = fma_rn(xr[d], xr[d], sb);
            for (int c = clo; c < chi; ++c) {
                const T* wv = s_cw + c * dsub;
                T dot = (T)0;
                for (int d = 0; d < dsub; ++d) dot = fma_rn(wv[d], xr[d], dot);
                T val = sub_rn(add_rn(s_nrm[c], sb), mul_rn((T)2, dot));
                val = val > (T)0 ? val : (T)0;
                if (besti < 0 || val < best) {
                    best = val;
                    besti = c;
                }
            }
        }
        s_bv[cg * EV + v] = best;
        s_bi[cg * EV + v] = besti;
        __syncthreads();
        if (cg == 0 && v0 + v < n) {
            T bb = s_bv[v];
            int bi = s_bi[v];
#pragma unroll
            for (int g = 1; g < 4; ++g) {
                const int gi = s_bi[g * EV + v];
                const T gv = s_bv[g * EV + v];
                if (gi >= 0 && (bi < 0 || gv < bb)) {  // strict: earlier quarter wins ties
                    bb = gv;
                    bi = gi;
                }
            }
            codes_out[(v0 + v) * m + i] = cb_codes[i * ksub + bi];
        }
    }
}

}  // namespace

// shared memory of encode_kernel: residual rows (all subspaces, or one slice when tiled), one codeword block,
// its norms, the per-quarter partial minima
size_t encode_smem_bytes(size_t elem, int m, int dsub, int ksub, bool tiled) {
    return elem * ((size_t)EV * ((tiled ? dsub : m * dsub) + 1) + (size_t)ksub * dsub + ksub + 4 * EV) +
           sizeof(int) * 4 * EV + 16;
}

bool encode_supported(const ivfadc_index* h) {
    const size_t elem = h->cfg.dtype == IVFADC_F32 ? 4 : 8;
    return encode_smem_bytes(elem, h->cfg.m, h->dsub, h->cfg.ksub, true) <= kSmemMax;
}

namespace {

template <typename T>
cudaError_t launch_encode_t(const ivfadc_index* h, const void* dX, int64_t n, const int32_t* d_cells,
                            uint8_t* d_codes_out, cudaStream_t s) {
    const int D = h->cfg.dim, m = h->cfg.m, dsub = h->dsub, ksub = h->cfg.ksub;
    const int tiled = encode_smem_bytes(sizeof(T), m, dsub, ksub, false) > kSmemMax ? 1 : 0;
    const size_t smem = encode_smem_bytes(sizeof(T), m, dsub, ksub, tiled != 0);
    const unsigned grid = (unsigned)((n + EV - 1) / EV);
    const T* X = static_cast<const T*>(dX);
    const T* C = static_cast<const T*>(h->d_centroids);
    const T* cb = static_cast<const T*>(h->d_cb);
    const T* nrm = static_cast<const T*>(h->d_cb_norms);
#define IVF_LAUNCH_ENC(DS)                                                                          \
    do {                                                                                            \
        auto kern = encode_kernel<T, DS>;                                                           \
        cudaError_t e = ensure_smem(h, reinterpret_cast<const void*>(kern), smem);                 \
        if (e != cudaSuccess) return e;                                                             \
        kern<<<grid, ETHREADS, smem, s>>>(X, n, d_cells, C, cb, h->d_cb_codes, nrm, D, m, dsub,     \
                                          ksub, d_codes_out, tiled, h->cfg.metric_resid);           \
    } while (0)
    switch (dsub) {
        case 4: IVF_LAUNCH_ENC(4); break;
        case 8: IVF_LAUNCH_ENC(8); break;
        case 16: IVF_LAUNCH_ENC(16); break;
        default: IVF_LAUNCH_ENC(0); break;
    }
#undef IVF_LAUNCH_ENC
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_codebook_norms(const ivfadc_index* h, cudaStream_t s, int* launches) {
    const int entries = h->cfg.m * h->cfg.ksub;
    const unsigned grid = (unsigned)((entries + 255) / 256);
    if (h->cfg.dtype == IVFADC_F32)
        codebook_norms_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(h->d_cb), entries, h->dsub,
                                                          static_cast<float*>(h->d_cb_norms));
    else
        codebook_norms_kernel<double><<<grid, 256, 0, s>>>(static_cast<const double*>(h->d_cb), entries, h->dsub,
                                                           static_cast<double*>(h->d_cb_norms));
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_encode(const ivfadc_index* h, const void* dX, int64_t n, const int32_t* d_cells,
                          uint8_t* d_codes_out, cudaStream_t s, int* launches) {
    if (n <= 0) return cudaSuccess;
    if (launches) *launches += 1;
    if (h->cfg.dtype == IVFADC_F32) return launch_encode_t<float>(h, dX, n, d_cells, d_codes_out, s);
    return launch_encode_t<double>(h, dX, n, d_cells, d_codes_out, s);
}

cudaError_t launch_assign_to_cells(const int64_t* d_assign, int64_t n, int base, int32_t* d_cells,
                                   cudaStream_t s, int* launches) {
    if (n <= 0) return cudaSuccess;
    assign_to_cells_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_assign, n, base, d_cells);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_assign_check(const int64_t* d_assign, int64_t n, int base, int kc, int* d_bad, cudaStream_t s,
                                int* launches) {
    if (n <= 0) return cudaSuccess;
    assign_check_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_assign, n, base, kc, d_bad);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace ivf
