// Quantizer training on the device -- the part of IVFADCIndex(data; ...) that the reference
// spends in Clustering.kmeans (src/index.jl:129-134) and QuantizedArrays.build_quantizer (src/index.jl:142-147).
// Not parity-graded (the reference seeds k-means++ from Julia's global RNG); graded by quantisation error.
//   * k-means++ seeding of k centres from a sample: D^2 sampling with a counter-based generator (Philox4x32-10,
//     deterministic per seed), one CTA -- the k rounds are sequential by nature, a round is a distance update of
//     the sample against the newest centre plus a block-wide prefix search;
//   * Lloyd iterations: the ASSIGNMENT step is the engine's own coarse kernel (K1, exact direct-form distances,
//     tensor-core pruned where the shape allows: ivfadc_coarse_search_device with w = 1); the centre UPDATE is the
//     accumulate / finish pair below (fp64 sums by atomics, so the update is exact up to fp64 rounding whatever
//     the order of the atomics); ivfadc_set_centroids_device installs the new centres in the same handle.
#include <algorithm>

#include "common.cuh"

namespace ivf {
namespace {

__device__ __forceinline__ void philox_tr(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                          uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

constexpr int KPT = 1024;  // threads of the seeding CTA
constexpr int KTR = 8;     // most candidates per round ("greedy" k-means++)

// k-means++ (Arthur & Vassilvitskii): first centre uniform, then each next one with probability ~ D(x)^2 -- in the
// greedy variant scikit-learn uses: `trials` candidates are drawn per round and the one that lowers the potential
// sum_i min(D(x_i)^2, |x_i - c|^2) most becomes the centre (trials = 1 is the plain algorithm).
// S [ns][D] sample, d2 [ns] scratch (global), centres out [k][D], picked [k] sample indices.
template <typename T>
__global__ void __launch_bounds__(KPT) kmeanspp_kernel(const T* __restrict__ S, int64_t ns, int D, int k, int trials, uint32_t k0,
                                                       uint32_t k1, double* __restrict__ d2, T* __restrict__ centres,
                                                       int64_t* __restrict__ picked) {
    __shared__ double part[KPT];
    __shared__ double wsum[KPT / 32][KTR];
    __shared__ double s_target[KTR];
    __shared__ long long s_pick[KTR];
    __shared__ int s_owner[KTR];
    __shared__ int s_best;
    extern __shared__ __align__(16) unsigned char kpp_smem[];
    T* cand = reinterpret_cast<T*>(kpp_smem);  // [trials][D] candidate centres of the round
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t per = (ns + KPT - 1) / KPT, lo = min(ns, (int64_t)tid * per), hi = min(ns, lo + per);
    uint32_t r[4];
    philox_tr(0u, 0u, 0u, 7u, k0, k1, r);
    long long pick = (long long)(((unsigned long long)r[0] * (unsigned long long)ns) >> 32);
    for (int j = 0; j < k; ++j) {
        // the chosen sample becomes centre j; update D^2 of this thread's contiguous slice + its partial sum
        if (tid == 0) picked[j] = pick;
        for (int d = tid; d < D; d += KPT) {
            const T v = S[pick * D + d];
            cand[d] = v;
            centres[(size_t)j * D + d] = v;
        }
        __syncthreads();
        if (j == k - 1) break;
        double acc = 0.0;
        for (int64_t i = lo; i < hi; ++i) {
            double s = 0.0;
            const T* x = S + i * D;
            for (int d = 0; d < D; ++d) {
                const double df = (double)x[d] - (double)cand[d];
                s = fma(df, df, s);
            }
            const double m = j == 0 ? s : fmin(d2[i], s);
            d2[i] = m;
            acc += m;
        }
        part[tid] = acc;
        __syncthreads();
        if (tid == 0) {  // `trials` draws of a position in the cumulative D^2 mass -> owning thread + offset in its slice
            double tot = 0.0;
            for (int t = 0; t < KPT; ++t) tot += part[t];
            for (int c = 0; c < trials; ++c) {
                philox_tr((uint32_t)(j + 1), (uint32_t)c, 0u, 7u, k0, k1, r);
                const double u = ((double)r[0] * 4294967296.0 + (double)r[1]) * (1.0 / 18446744073709551616.0);
                double target = u * tot, run = 0.0;
                int owner = KPT - 1;
                for (int t = 0; t < KPT; ++t) {
                    if (run + part[t] > target) { owner = t; break; }
                    run += part[t];
                }
                s_owner[c] = owner;
                s_target[c] = target - run;
                s_pick[c] = 0;
            }
        }
        __syncthreads();
        for (int c = 0; c < trials; ++c) {
            if (tid == s_owner[c]) {
                double run = 0.0;
                long long p = hi > lo ? hi - 1 : 0;
                for (int64_t i = lo; i < hi; ++i) {
                    run += d2[i];
                    if (run > s_target[c]) { p = i; break; }
                }
                s_pick[c] = (p >= 0 && p < ns) ? p : 0;
            }
        }
        __syncthreads();
        if (trials > 1) {  // potential of every candidate; the best one wins
            for (int idx = tid; idx < trials * D; idx += KPT) cand[idx] = S[s_pick[idx / D] * D + idx % D];
            __syncthreads();
            double pot[KTR];
#pragma unroll
            for (int c = 0; c < KTR; ++c) pot[c] = 0.0;
            for (int64_t i = lo; i < hi; ++i) {
                const T* x = S + i * D;
                const double cur = d2[i];
#pragma unroll
                for (int c = 0; c < KTR; ++c) {
                    if (c < trials) {
                        double s = 0.0;
                        const T* cc = cand + c * D;
                        for (int d = 0; d < D; ++d) {
                            const double df = (double)x[d] - (double)cc[d];
                            s = fma(df, df, s);
                        }
                        pot[c] += fmin(cur, s);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < KTR; ++c) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) pot[c] += __shfl_xor_sync(0xffffffffu, pot[c], o);
                if (lane == 0) wsum[wid][c] = pot[c];
            }
            __syncthreads();
            if (tid == 0) {
                int best = 0;
                double bv = 0.0;
                for (int c = 0; c < trials; ++c) {
                    double v = 0.0;
                    for (int wv = 0; wv < KPT / 32; ++wv) v += wsum[wv][c];
                    if (c == 0 || v < bv) { bv = v; best = c; }
                }
                s_best = best;
            }
            __syncthreads();
            pick = s_pick[s_best];
        } else {
            pick = s_pick[0];
        }
        __syncthreads();
    }
}

// centre update, step 1: sums[cell][d] += x[d], counts[cell] += 1 (one warp per vector, fp64 atomics)
template <typename T>
__global__ void kmeans_accumulate_kernel(const T* __restrict__ X, int64_t n, int D, const int32_t* __restrict__ cells,
                                         double* __restrict__ sums, unsigned long long* __restrict__ counts) {
    const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (v >= n) return;
    const int c = cells[v];
    for (int d = lane; d < D; d += 32) atomicAdd(sums + (size_t)c * D + d, (double)X[v * D + d]);
    if (lane == 0) atomicAdd(counts + c, 1ull);
}

// step 2: centre = sum / count where the cell is not empty (empty cells keep their centre; empty[cell] flags them)
template <typename T>
__global__ void kmeans_finish_kernel(const double* __restrict__ sums, const unsigned long long* __restrict__ counts, int kc, int D,
                                     T* __restrict__ centroids, int32_t* __restrict__ empty) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)kc * D) return;
    const int c = (int)(i / D);
    const unsigned long long cnt = counts[c];
    if (cnt > 0) centroids[i] = (T)(sums[i] / (double)cnt);
    if (empty && i == (int64_t)c * D) empty[c] = cnt == 0 ? 1 : 0;
}

}  // namespace
}  // namespace ivf

using namespace ivf;

extern "C" {

int ivfadc_kmeanspp_device(const void* dS, int64_t ns, int32_t D, int32_t k, int32_t dtype, uint64_t seed, int32_t trials,
                           void* d_scratch, void* d_centres_out, int64_t* d_picked_out, void* stream) {
    if (!dS || !d_scratch || !d_centres_out || !d_picked_out || ns < 1 || D < 1 || k < 1 || k > ns) return IVFADC_ERR_BAD_ARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t esz = dtype == IVFADC_F32 ? 4 : 8;
    trials = std::max(1, std::min<int>(trials, KTR));
    while (trials > 1 && (size_t)trials * D * esz > 32 * 1024) --trials;   // candidate block in shared memory
    const size_t smem = (size_t)trials * D * esz;
    if (smem > 32 * 1024) return IVFADC_ERR_UNSUPPORTED;
    if (dtype == IVFADC_F32)
        kmeanspp_kernel<float><<<1, KPT, smem, s>>>(static_cast<const float*>(dS), ns, D, k, trials, (uint32_t)seed, (uint32_t)(seed >> 32),
                                                    static_cast<double*>(d_scratch), static_cast<float*>(d_centres_out), d_picked_out);
    else
        kmeanspp_kernel<double><<<1, KPT, smem, s>>>(static_cast<const double*>(dS), ns, D, k, trials, (uint32_t)seed, (uint32_t)(seed >> 32),
                                                     static_cast<double*>(d_scratch), static_cast<double*>(d_centres_out), d_picked_out);
    return cudaGetLastError() == cudaSuccess ? IVFADC_OK : IVFADC_ERR_CUDA;
}

int ivfadc_kmeans_accumulate_device(const void* dX, int64_t n, int32_t D, int32_t dtype, const int32_t* d_cells, double* d_sums,
                                    uint64_t* d_counts, void* stream) {
    if (n < 0 || (n > 0 && (!dX || !d_cells || !d_sums || !d_counts)) || D < 1) return IVFADC_ERR_BAD_ARG;
    if (n == 0) return IVFADC_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const unsigned grid = (unsigned)((n * 32 + 255) / 256);
    if (dtype == IVFADC_F32)
        kmeans_accumulate_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(dX), n, D, d_cells, d_sums,
                                                             reinterpret_cast<unsigned long long*>(d_counts));
    else
        kmeans_accumulate_kernel<double><<<grid, 256, 0, s>>>(static_cast<const double*>(dX), n, D, d_cells, d_sums,
                                                              reinterpret_cast<unsigned long long*>(d_counts));
    return cudaGetLastError() == cudaSuccess ? IVFADC_OK : IVFADC_ERR_CUDA;
}

int ivfadc_kmeans_finish_device(const double* d_sums, const uint64_t* d_counts, int32_t kc, int32_t D, int32_t dtype,
                                void* d_centroids_inout, int32_t* d_empty_out, void* stream) {
    if (!d_sums || !d_counts || !d_centroids_inout || kc < 1 || D < 1) return IVFADC_ERR_BAD_ARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t nthr = (int64_t)kc * D;
    const unsigned grid = (unsigned)((nthr + 255) / 256);
    if (dtype == IVFADC_F32)
        kmeans_finish_kernel<float><<<grid, 256, 0, s>>>(d_sums, reinterpret_cast<const unsigned long long*>(d_counts), kc, D,
                                                         static_cast<float*>(d_centroids_inout), d_empty_out);
    else
        kmeans_finish_kernel<double><<<grid, 256, 0, s>>>(d_sums, reinterpret_cast<const unsigned long long*>(d_counts), kc, D,
                                                          static_cast<double*>(d_centroids_inout), d_empty_out);
    return cudaGetLastError() == cudaSuccess ? IVFADC_OK : IVFADC_ERR_CUDA;
}

}  // extern "C"
