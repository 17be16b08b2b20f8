// Synthetic inputs for the benchmark and the tests -- a harness utility that ships with the
// library so that the 10 M / 100 M-vector configurations are generated in HBM chunk by chunk instead of crossing
// PCIe: a counter-based generator, Philox4x32-10 keyed by the seed, counter = (vector lo, vector hi, dim, stream).
// Every value depends only on (seed, vector, dim); oracle/ivfadc_oracle.c holds the same function for the CPU,
// bit for bit (integer arithmetic, one exact int -> float conversion, one multiplication, one addition).
#include "common.cuh"

namespace ivf {
namespace {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0;
        c1 = l1;
        c2 = h0 ^ c3 ^ k1;
        c3 = l0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__global__ void synth_uniform_kernel(float* __restrict__ X, int64_t first, int64_t n, int D, uint32_t k0, uint32_t k1) {
    const int dq = (D + 3) >> 2;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * dq) return;
    const int64_t i = e / dq;
    const int g = (int)(e - i * dq);
    const uint64_t v = (uint64_t)(first + i);
    uint32_t r[4];
    philox4x32_10((uint32_t)v, (uint32_t)(v >> 32), (uint32_t)g, 2u, k0, k1, r);
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (4 * g + j < D) X[i * D + 4 * g + j] = __fmul_rn((float)(r[j] >> 8), 5.9604644775390625e-08f);
}

__global__ void synth_blobs_kernel(float* __restrict__ X, int32_t* __restrict__ blobs_out, int64_t first, int64_t n,
                                   int D, int n_blobs, uint32_t k0, uint32_t k1, float scale,
                                   const float* __restrict__ centres) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * D) return;
    const int64_t i = e / D;
    const int d = (int)(e - i * D);
    const uint64_t v = (uint64_t)(first + i);
    uint32_t r[4];
    philox4x32_10((uint32_t)v, (uint32_t)(v >> 32), 0u, 1u, k0, k1, r);
    const int b = (int)(((uint64_t)r[0] * (uint64_t)n_blobs) >> 32);
    if (blobs_out && d == 0) blobs_out[i] = b;
    philox4x32_10((uint32_t)v, (uint32_t)(v >> 32), (uint32_t)d, 0u, k0, k1, r);
    const int t = (int)((r[0] >> 10) + (r[1] >> 10) + (r[2] >> 10) + (r[3] >> 10)) - 8388606;
    X[e] = __fadd_rn(centres[(size_t)b * D + d], __fmul_rn((float)t, scale));
}

}  // namespace
}  // namespace ivf

extern "C" {

int ivfadc_synth_uniform_device(void* dX, int64_t first, int64_t n, int32_t D, uint64_t seed, void* stream) {
    if (!dX || n < 0 || D <= 0) return IVFADC_ERR_BAD_ARG;
    if (n == 0) return IVFADC_OK;
    const int64_t work = n * ((D + 3) / 4);
    ivf::synth_uniform_kernel<<<(unsigned)((work + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<float*>(dX), first, n, D, (uint32_t)seed, (uint32_t)(seed >> 32));
    return cudaGetLastError() == cudaSuccess ? IVFADC_OK : IVFADC_ERR_CUDA;
}

int ivfadc_synth_blobs_device(void* dX, int32_t* d_blobs_out, int64_t first, int64_t n, int32_t D, int32_t n_blobs,
                              uint64_t seed, float scale, const void* d_centres, void* stream) {
    if (!dX || !d_centres || n < 0 || D <= 0 || n_blobs <= 0) return IVFADC_ERR_BAD_ARG;
    if (n == 0) return IVFADC_OK;
    if (n * (int64_t)D > (int64_t)0xffffffffu * 256) return IVFADC_ERR_BAD_ARG;
    const int64_t work = n * D;
    ivf::synth_blobs_kernel<<<(unsigned)((work + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<float*>(dX), d_blobs_out, first, n, D, n_blobs, (uint32_t)seed, (uint32_t)(seed >> 32), scale,
        static_cast<const float*>(d_centres));
    return cudaGetLastError() == cudaSuccess ? IVFADC_OK : IVFADC_ERR_CUDA;
}

}  // extern "C"
