// Micro-benchmark (bring-up tool, not product): throughput of data-dependent, warp-uniform lookups
//   (a) from shared memory  (LDS, 32 lanes x 4 B = one 128-byte wavefront per lookup), and
//   (b) from tensor memory  (tcgen05.ld.32x32b.x1, column = code byte, lane = query),
//   (c) both interleaved,
// with 16 warps per CTA and one CTA per SM, the shape of the list-scan kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_gather tmem_gather.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int WARPS = 16, THREADS = WARPS * 32;
constexpr int NCODES = 4096;      // code bytes per warp per repetition (uniform loads of 16)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float lds_f(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds_v4(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ float ldtm1(uint32_t taddr) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
    return __uint_as_float(v);
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int BY> __device__ __forceinline__ uint32_t byte_of(uint32_t x) { return (x >> (8 * BY)) & 0xFFu; }

// mode 0: LDS only, 1: LDTM only, 2: alternate words (half / half), 3: LDTM 3 of 4
template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) gather_kernel(const uint8_t* codes, float* out, long long* cycles, int reps) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, lane = tid & 31;
    const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform for ptxas: codes land in uniform registers
    const uint32_t sb = smem_u32(smem);
    // layout: table 64 KB at a 64 KB-aligned address, then code bytes 16 warps x NCODES
    const uint32_t lut = (sb + 0xFFFFu) & ~0xFFFFu;
    const uint32_t cod = lut + 65536;
    for (int i = tid; i < 16384; i += THREADS) asm volatile("st.shared.f32 [%0], %1;" ::"r"(lut + i * 4), "f"((float)(i & 1023)) : "memory");
    for (int i = tid; i < WARPS * NCODES / 4; i += THREADS)
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(cod + i * 4), "r"(reinterpret_cast<const uint32_t*>(codes)[i]) : "memory");
    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_slot + ((uint32_t)((wid & 3) * 32) << 16);
    // fill the tensor memory of this lane quarter (4 warps write the same data, harmless)
    for (int c = 0; c < 512; ++c) {
        uint32_t v = __float_as_uint((float)c);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tbase + c), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    const uint32_t lo = lut | (uint32_t)(lane * 4);
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        const uint32_t cw = cod + wid * NCODES;
#pragma unroll 1
        for (int c = 0; c < NCODES; c += 32) {
            const uint4 x = lds_v4(cw + c), y = lds_v4(cw + c + 16);
            const uint32_t w[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
            float t[16];
            // first 16 codes
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t ww = w[4 * h + j];
                    const bool use_tm = MODE == 1 || (MODE == 2 && (j & 1)) || (MODE == 3 && (j != 0));
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        if (use_tm) t[4 * j + b] = ldtm1(__byte_perm(ww, tbase + ((j & 1) ? 256u : 0u), 0x7650 | b));
                        else t[4 * j + b] = lds_f(__byte_perm(ww, lo, 0x7604 | (b << 4)));
                    }
                }
                if (MODE != 0) wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] += t[i];
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * THREADS + tid] = s;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
}

template <int MODE>
void run(const char* name, const uint8_t* d_codes, float* d_out, long long* d_cyc, int nsm) {
    const int smem = 65536 + 65536 + WARPS * NCODES + 1024;
    cudaFuncSetAttribute(gather_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int reps = 20;
    gather_kernel<MODE><<<nsm, THREADS, smem>>>(d_codes, d_out, d_cyc, 2);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    gather_kernel<MODE><<<nsm, THREADS, smem>>>(d_codes, d_out, d_cyc, reps);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long cyc = 0; cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
    float o = 0; cudaMemcpy(&o, d_out, 4, cudaMemcpyDeviceToHost);
    const double lookups_per_cta = (double)reps * WARPS * NCODES;  // warp-level lookup instructions
    printf("%-28s %s  %.3f ms  %lld cycles/CTA  %.3f cycles per warp-lookup per SM  (= %.1f B/clk/SM of 4-byte entries)  check %.1f\n",
           name, cudaGetErrorString(err), ms, cyc, cyc / lookups_per_cta, 128.0 * lookups_per_cta / cyc, o);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int nsm = p.multiProcessorCount;
    uint8_t* h = (uint8_t*)malloc(WARPS * NCODES);
    uint32_t s = 12345;
    for (int i = 0; i < WARPS * NCODES; ++i) { s = s * 1664525u + 1013904223u; h[i] = (uint8_t)(s >> 24); }
    uint8_t* d_codes; float* d_out; long long* d_cyc;
    cudaMalloc(&d_codes, WARPS * NCODES); cudaMalloc(&d_out, nsm * THREADS * 4); cudaMalloc(&d_cyc, nsm * 8);
    cudaMemcpy(d_codes, h, WARPS * NCODES, cudaMemcpyHostToDevice);
    printf("%s, %d SMs\n", p.name, nsm);
    run<0>("LDS only", d_codes, d_out, d_cyc, nsm);
    run<1>("LDTM.x1 only", d_codes, d_out, d_cyc, nsm);
    run<2>("LDS/LDTM 1:1", d_codes, d_out, d_cyc, nsm);
    run<3>("LDS/LDTM 1:3", d_codes, d_out, d_cyc, nsm);
    return 0;
}
