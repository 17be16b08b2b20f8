// Micro-benchmark (bring-up tool, not product): tcgen05.ld.32x32b.x1 lookups (16 warps, data-dependent
// uniform column) with B lookups in flight per tcgen05.wait::ld, alone and while one thread keeps the
// tensor core busy with M128 N256 K8 kind::tf32 (or K16 kind::f16) MMAs into the other half of tensor
// memory -- the situation of the list-scan kernel.  Also times the MMAs alone.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_mma_ld tmem_mma_ld.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#ifndef NWARPS
#define NWARPS 16
#endif
constexpr int WARPS = NWARPS, THREADS = WARPS * 32;
constexpr int NCODES = 4096;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint4 lds_v4(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ float ldtm1(uint32_t taddr) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
    return __uint_as_float(v);
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    constexpr uint64_t LBO = 128, SBO = 256;
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((LBO >> 4) << 16) | ((SBO >> 4) << 32) | (1ull << 46);
}
constexpr uint32_t IDESC_TF32 = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t IDESC_F16 = (1u << 4) | (0u << 7) | (0u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

template <int KIND>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(IDESC_TF32), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(IDESC_F16), "r"(acc) : "memory");
}

// LOOK: 0 = no lookups (MMA only), else lookups in flight per wait (8, 16, 32).  NMMA: MMAs per "build"
// issued by warp 15 lane 0 every `period` lookups-chunks (0 = none).
template <int LOOK, int KIND, int SPIN>
__global__ void __launch_bounds__(THREADS, 1) k(const uint8_t* codes, float* out, long long* cycles, int reps, int nmma) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.x, lane = tid & 31;
    const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t sb = smem_u32(smem);
    const uint32_t abuf = sb, bbuf = sb + 8192, cod = sb + 8192 + 16384;
    for (int i = tid; i < (8192 + 16384) / 4; i += THREADS) asm volatile("st.shared.u32 [%0], %1;" ::"r"(sb + i * 4), "r"(0u) : "memory");
    for (int i = tid; i < WARPS * NCODES / 4; i += THREADS)
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(cod + i * 4), "r"(reinterpret_cast<const uint32_t*>(codes)[i]) : "memory");
    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_slot + ((uint32_t)((wid & 3) * 32) << 16);
    for (int c = 0; c < 256; ++c) {
        uint32_t v = __float_as_uint((float)c);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tbase + c), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    const long long t0 = clock64();
    if (wid == WARPS - 1 && nmma > 0) {
        // MMA issuer warp: lane 0 issues `nmma` MMAs, commits, waits, repeats
        if (lane == 0) {
            const uint64_t ad = smem_desc(abuf), bd = smem_desc(bbuf);
            uint32_t phase = 0;
            for (int r = 0; r < reps; ++r) {
                for (int i = 0; i < nmma; ++i) mma<KIND>(tmem_slot + 256, ad, bd, i > 0);
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
                uint32_t ok = 0;
                if (SPIN) { while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory"); }
                else { while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory"); }
                phase ^= 1;
            }
        }
        __syncwarp();
    } else if constexpr (LOOK > 0) {
        const int lreps = reps;  // one "subspace" of 64 lookups per warp per rep
        for (int r = 0; r < lreps; ++r) {
            const uint32_t cw = cod + wid * NCODES + (r & 63) * 64;
#pragma unroll
            for (int b = 0; b < 64 / LOOK; ++b) {
                float t[LOOK];
#pragma unroll
                for (int q = 0; q < LOOK / 16; ++q) {
                    const uint4 x = lds_v4(cw + b * LOOK + q * 16);
                    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int by = 0; by < 4; ++by) t[q * 16 + 4 * j + by] = ldtm1(__byte_perm(w[j], tbase, 0x7650 | by));
                }
                if (LOOK == 8) {}
                wait_ld();
#pragma unroll
                for (int i = 0; i < LOOK; ++i) acc[i & 31] += t[i];
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i];
    out[blockIdx.x * THREADS + tid] = s;
    if (tid == 0) cycles[blockIdx.x * 2] = t1 - t0;
    if (tid == THREADS - 32) cycles[blockIdx.x * 2 + 1] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
}

template <int LOOK, int KIND, int SPIN = 0>
void run(const char* name, const uint8_t* d_codes, float* d_out, long long* d_cyc, int nsm, int nmma) {
    const int smem = 8192 + 16384 + WARPS * NCODES + 1024;
    cudaFuncSetAttribute(k<LOOK, KIND, SPIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int reps = 400;
    k<LOOK, KIND, SPIN><<<nsm, THREADS, smem>>>(d_codes, d_out, d_cyc, 4, nmma);
    k<LOOK, KIND, SPIN><<<nsm, THREADS, smem>>>(d_codes, d_out, d_cyc, reps, nmma);
    cudaError_t err = cudaDeviceSynchronize();
    long long cyc[2] = {0, 0}; cudaMemcpy(cyc, d_cyc, 16, cudaMemcpyDeviceToHost);
    printf("%-44s %s  lookup warps: %7.1f cycles per 64-lookup subspace (%.2f cyc/lookup/SM) | mma warp: %7.1f cycles per build of %d MMAs\n",
           name, cudaGetErrorString(err), (double)cyc[0] / reps, (double)cyc[0] / reps / (64.0 * (WARPS - 1)), (double)cyc[1] / reps, nmma);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int nsm = p.multiProcessorCount;
    uint8_t* h = (uint8_t*)malloc(WARPS * NCODES);
    uint32_t s = 12345;
    for (int i = 0; i < WARPS * NCODES; ++i) { s = s * 1664525u + 1013904223u; h[i] = (uint8_t)(s >> 24); }
    uint8_t* d_codes; float* d_out; long long* d_cyc;
    cudaMalloc(&d_codes, WARPS * NCODES); cudaMalloc(&d_out, nsm * THREADS * 4); cudaMalloc(&d_cyc, nsm * 16);
    cudaMemcpy(d_codes, h, WARPS * NCODES, cudaMemcpyHostToDevice);
    printf("%s, %d SMs; 15 lookup warps + 1 MMA issuer warp\n", p.name, nsm);
    run<0, 0>("MMA tf32 only, 4 per build", d_codes, d_out, d_cyc, nsm, 4);
    run<0, 0>("MMA tf32 only, 1 per build", d_codes, d_out, d_cyc, nsm, 1);
    run<0, 0, 1>("MMA tf32 only, 1 per build, test_wait spin", d_codes, d_out, d_cyc, nsm, 1);
    run<0, 0, 1>("MMA tf32 only, 2 per build, test_wait spin", d_codes, d_out, d_cyc, nsm, 2);
    run<0, 0, 1>("MMA tf32 only, 4 per build, test_wait spin", d_codes, d_out, d_cyc, nsm, 4);
    run<0, 0, 1>("MMA tf32 only, 8 per build, test_wait spin", d_codes, d_out, d_cyc, nsm, 8);
    run<16, 0, 1>("lookups x16 + tf32 MMA (4), test_wait spin", d_codes, d_out, d_cyc, nsm, 4);
    run<16, 0, 1>("lookups x16 + tf32 MMA (2), test_wait spin", d_codes, d_out, d_cyc, nsm, 2);
    run<0, 1>("MMA f16 only, 4 per build", d_codes, d_out, d_cyc, nsm, 4);
    run<16, 0>("lookups x16 in flight, no MMA", d_codes, d_out, d_cyc, nsm, 0);
    run<32, 0>("lookups x32 in flight, no MMA", d_codes, d_out, d_cyc, nsm, 0);
    run<64, 0>("lookups x64 in flight, no MMA", d_codes, d_out, d_cyc, nsm, 0);
    run<16, 0>("lookups x16 + tf32 MMA back to back (4)", d_codes, d_out, d_cyc, nsm, 4);
    run<32, 0>("lookups x32 + tf32 MMA back to back (4)", d_codes, d_out, d_cyc, nsm, 4);
    run<16, 1>("lookups x16 + f16 MMA back to back (4)", d_codes, d_out, d_cyc, nsm, 4);
    run<16, 1>("lookups x16 + f16 MMA back to back (2)", d_codes, d_out, d_cyc, nsm, 2);
    run<16, 1>("lookups x16 + f16 MMA back to back (1)", d_codes, d_out, d_cyc, nsm, 1);
    run<16, 0>("lookups x16 + tf32 MMA back to back (2)", d_codes, d_out, d_cyc, nsm, 2);
    run<16, 0>("lookups x16 + tf32 MMA back to back (1)", d_codes, d_out, d_cyc, nsm, 1);
    run<0, 1>("MMA f16 only, 2 per build", d_codes, d_out, d_cyc, nsm, 2);
    run<0, 1>("MMA f16 only, 1 per build", d_codes, d_out, d_cyc, nsm, 1);
    return 0;
}
