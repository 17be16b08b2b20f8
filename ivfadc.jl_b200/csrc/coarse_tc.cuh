// K1 on the 5th-generation tensor cores: coarse assignment as a dense query x centroid contraction
// (tcgen05.mma kind::tf32, accumulators in tensor memory, centroid operand streamed by TMA bulk
// copies), fused with a per-query candidate selection, followed by an EXACT re-rank.
// Replaces coarse_search(::NaiveQuantizer, point, w), reference src/coarsequantizers.jl:33-37
// (colwise distances, stable sortperm, first w).
//
// The returned cells and distances are bit-identical to the oracle's direct form
// sum_d (c_d - q_d)^2 (one sequential fp32 fma chain per pair): the tensor cores only PRUNE.
//
//   1. score s(c) = |c|^2 - 2 q.c with q, c rounded to TF32 (ONE piece: no hi / lo split), fp32
//      accumulate: CTA = 128 queries (accumulator lanes) x all centroids in tiles of 256 (columns),
//      K = 8 dims per MMA.  Two accumulator tiles in tensor memory: the tensor core fills tile
//      t + 1 while the four epilogue warps (lane = query) read tile t.
//   2. |s(c) + |q|^2 - d(c)| <= E with E = 2^-10 (|q|^2 + max|c|^2) (TF32 rounding of both operands,
//      Cauchy-Schwarz; fp32 accumulation and the chain's own rounding are 2^-15 of that).  The
//      kernel uses 2E = 2^-8 (|q|^2 + max|c|^2), twice the bound.
//   3. selection, branch-free per lane: the WL-th smallest (WL >= w) of the minima of groups of 8
//      columns seen so far is an upper bound B of the w-th smallest score; every centroid with
//      s <= B + 2E is a candidate.  If a centroid of the exact top-w had s > B + 2E, the >= w
//      centroids with s <= B would all have a strictly smaller exact distance -- so the candidates
//      are a superset of the exact top-w, ties included.  The tile is read twice from tensor
//      memory (minima, then filter): re-reading costs no shared-memory or HBM traffic.
//   4. exact re-rank: one warp per query, lane = candidate: the oracle's fma chain over the fp32
//      centroid row, then w rounds of warp arg-min by (distance, cell) -- sortperm's stable order.
//   5. a query with more candidates than slots (heavy ties, duplicate centroids) is flagged and
//      redone by the packed-FP32 kernel (coarse2_kernel with a redo mask, launched right after).
//
// SASS: UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UBLKCP (cp.async.bulk), SYNCS (mbarrier).
#pragma once

#include "common.cuh"
#include <cstdio>

namespace ivf {
namespace ctc {

constexpr int MQ = 128;                 // queries per CTA = accumulator lanes
constexpr int NC = 256;                 // centroids per tile = accumulator columns
constexpr int ABLK = 4096;              // A block: 128 rows x 8 k (tf32 words), canonical K-major no-swizzle layout
constexpr int BBLK = 8192;              // B block: 256 rows x 8 k
constexpr int NSLOT = 8;                // B ring depth (k-steps in flight)
constexpr int CAP = 64;                 // candidate slots per query
constexpr int CSTR = MQ + 1;            // slot stride (words) of the candidate arrays: conflict-free by row and by slot
constexpr int THREADS = 256;            // warps 0..3 epilogue (lane quarter = warp), 4 TMA producer, 5 MMA issuer
constexpr uint32_t SPIN = 1u << 22;     // bound on every mbarrier wait (error flag instead of a hang)
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

struct Smem {
    uint32_t a, b, cand_s, cand_c, norms, qn, thr, cnt, scratch, bars, total;
};
__host__ __device__ inline Smem smem_layout(int ksteps) {
    Smem s;
    uint32_t o = 0;
    s.a = o;       o += (uint32_t)ksteps * ABLK;
    s.b = o;       o += NSLOT * BBLK;
    s.cand_s = o;  o += CAP * CSTR * 4;
    s.cand_c = o;  o += CAP * CSTR * 4;
    s.norms = o;   o += 4 * NC * 4;          // per epilogue warp: squared norms of the tile's centroids
    s.qn = o;      o += MQ * 4;
    s.thr = o;     o += MQ * 4;
    s.cnt = o;     o += MQ * 4;
    s.scratch = o; o += (THREADS / 32) * CAP * 4;
    s.bars = o;    o += 256;
    s.total = o;
    return s;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
// `dead`: once a wait of this thread has timed out, its later waits return at once (a broken pipeline
// costs one bounded spin per role, not one per k-step)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int code, bool& dead) {
    if (dead) return;
#pragma unroll 1
    for (uint32_t i = 0; i < SPIN; ++i)
        if (mbar_try_wait(bar, parity)) return;
    atomicExch(err, code);
    dead = true;
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// K-major, SWIZZLE_NONE: 8-row x 16-byte core matrices; LBO = 128 (k 0..3 | 4..7), SBO = 256 (8-row groups)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    constexpr uint64_t LBO = 128, SBO = 256;
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((LBO >> 4) << 16) | ((SBO >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ uint32_t tf32_of(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

struct Args {
    const float* Q;        // [nq][D]
    const float* C;        // [kc][D] fp32 centroids (exact re-rank)
    const float* tcC;      // [kcp / 256][ksteps][2048] tf32 words: centroid operand blocks
    const float* cn;       // [kcp] squared norms (+inf beyond kc), then max |c|^2 at [kcp]
    int64_t nq;
    int kc, kcp, D, ksteps, w;
    int32_t* cells_out;    // [nq][w]
    float* dc_out;         // [nq][w]
    uint8_t* redo;         // [nq]: 1 = candidate overflow, the packed-FP32 kernel redoes the query
    int32_t* cand_out;     // [nq][CAP] cells that survive the pruning
    int32_t* cnt_out;      // [nq] their number, -1 = overflow
    int* err;
    int force_redo;        // test switch: flag every query
};

// WL: length of the per-lane sorted list of group minima (>= w)
template <int WL>
__global__ void __launch_bounds__(THREADS, 1) coarse3_kernel(const Args a) {
    extern __shared__ __align__(1024) unsigned char smem_c3[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t sb = smem_u32(smem_c3);
    const Smem L = smem_layout(a.ksteps);
    const uint32_t bar_full = sb + L.bars;                 // NSLOT x 8: B k-step landed
    const uint32_t bar_empty = bar_full + 8 * NSLOT;       // NSLOT x 8: the MMA that read the slot has completed
    const uint32_t bar_tfull = bar_empty + 8 * NSLOT;      // 2 x 8: accumulator tile complete
    const uint32_t bar_tempty = bar_tfull + 16;            // 2 x 8: the four epilogue warps are done with the tile
    const uint32_t tmem_slot = bar_tempty + 16;
    float* qn_s = reinterpret_cast<float*>(smem_c3 + L.qn);
    float* cand_s = reinterpret_cast<float*>(smem_c3 + L.cand_s);
    int* cand_c = reinterpret_cast<int*>(smem_c3 + L.cand_c);
    const int64_t q0 = (int64_t)blockIdx.x * MQ;
    const int ntiles = a.kcp / NC;
    const int KS = a.ksteps;
    bool dead = false;
#ifdef C3_STAMP
    long long stamps[40];
    int nst = 0;
#define C3S() do { if (blockIdx.x == 0 && tid == 0 && nst < 40) stamps[nst++] = clock64(); } while (0)
#else
#define C3S() do { } while (0)
#endif
    C3S();

    if (tid == 0) {
        for (int i = 0; i < NSLOT; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid < MQ) qn_s[tid] = 0.f;
    __syncthreads();

    // ---- the centroid operand starts streaming while the query operand is written ----
    if (wid == 4 && lane == 0) {
        const int total = ntiles * KS;
        for (int it = 0; it < total && it < NSLOT; ++it) {
            mbar_expect_tx(bar_full + 8 * it, BBLK);
            tma_bulk_g2s(sb + L.b + it * BBLK, a.tcC + (size_t)it * (BBLK / 4), BBLK, bar_full + 8 * it);
        }
    }
    // ---- A operand: the 128 query rows rounded to TF32, block j = dims 8 j .. 8 j + 7;
    //      word(n, k) = (n >> 3) * 64 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3)
    {
        const int nchunk = MQ * 2 * KS;  // 16-byte chunks (4 dims of one row)
        constexpr int UNR = 8;           // loads in flight per thread
        for (int base = tid; base < nchunk; base += THREADS * UNR) {
            float4 v[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int idx = base + u * THREADS;
                const int n_lo = idx & 7, k4_lo = (idx >> 3) & 3, hi = idx >> 5;
                const int n = (hi & 15) * 8 + n_lo, k4 = (hi >> 4) * 4 + k4_lo;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx < nchunk && q0 + n < a.nq) v[u] = __ldg(reinterpret_cast<const float4*>(a.Q + (size_t)(q0 + n) * a.D + 4 * k4));
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int idx = base + u * THREADS;
                if (idx >= nchunk) break;
                const int n_lo = idx & 7, k4_lo = (idx >> 3) & 3, hi = idx >> 5;
                const int n = (hi & 15) * 8 + n_lo, k4 = (hi >> 4) * 4 + k4_lo;
                atomicAdd(&qn_s[n], fmaf(v[u].x, v[u].x, fmaf(v[u].y, v[u].y, fmaf(v[u].z, v[u].z, v[u].w * v[u].w))));
                const uint32_t dst = sb + L.a + (k4 >> 1) * ABLK + (n >> 3) * 256 + (k4 & 1) * 128 + (n & 7) * 16;
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(tf32_of(v[u].x)), "r"(tf32_of(v[u].y)),
                             "r"(tf32_of(v[u].z)), "r"(tf32_of(v[u].w))
                             : "memory");
            }
        }
    }
    C3S();
    fence_proxy_async();  // generic-proxy writes of A -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    C3S();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_c3 + L.bars + 16 * NSLOT + 32);  // = tmem_slot

    if (wid == 4) {
        // ---- TMA producer: k-step `it` of the whole tile sequence goes to ring slot it % NSLOT ----
        if (lane == 0) {
            const int total = ntiles * KS;
            for (int it = NSLOT; it < total; ++it) {
                const int slot = it % NSLOT;
                mbar_wait(bar_empty + 8 * slot, ((it / NSLOT) - 1) & 1, a.err, 11, dead);
                mbar_expect_tx(bar_full + 8 * slot, BBLK);
                tma_bulk_g2s(sb + L.b + slot * BBLK, a.tcC + (size_t)it * (BBLK / 4), BBLK, bar_full + 8 * slot);
            }
        }
    } else if (wid == 5) {
        // ---- MMA issuer ----
        if (lane == 0) {
            int it = 0;
            for (int t = 0; t < ntiles; ++t) {
                const int buf = t & 1;
                if (t >= 2) mbar_wait(bar_tempty + 8 * buf, ((t >> 1) - 1) & 1, a.err, 12, dead);
                tc_fence_after();
                for (int j = 0; j < KS; ++j, ++it) {
                    const int slot = it % NSLOT;
                    mbar_wait(bar_full + 8 * slot, (it / NSLOT) & 1, a.err, 13, dead);
                    tc_fence_after();
                    tc_mma(tmem_base + buf * NC, smem_desc(sb + L.a + j * ABLK), smem_desc(sb + L.b + slot * BBLK), j > 0);
                    tc_commit(bar_empty + 8 * slot);
                }
                tc_commit(bar_tfull + 8 * buf);
            }
        }
    } else if (wid < 4) {
        // ---- epilogue: lane = query row 32 wid + lane ----
        const int row = 32 * wid + lane;
        float* nrm = reinterpret_cast<float*>(smem_c3 + L.norms) + wid * NC;
        const float cmax2 = __ldg(a.cn + a.kcp);
        const float margin = (qn_s[row] + cmax2) * 0.00390625f;  // 2E = 2^-8 (|q|^2 + max |c|^2)
        float lst[WL];
#pragma unroll
        for (int i = 0; i < WL; ++i) lst[i] = Limits<float>::inf();
        int cnt = 0;
        const uint32_t trow = tmem_base + ((uint32_t)(32 * wid) << 16);
        for (int t = 0; t < ntiles; ++t) {
            const int buf = t & 1;
            // squared norms of this tile's centroids, private to the warp
            __syncwarp();
            {
                const float4* src = reinterpret_cast<const float4*>(a.cn + (size_t)t * NC);
                reinterpret_cast<float4*>(nrm)[lane] = __ldg(src + lane);
                reinterpret_cast<float4*>(nrm)[lane + 32] = __ldg(src + lane + 32);
            }
            __syncwarp();
            C3S();
            if (lane == 0) mbar_wait(bar_tfull + 8 * buf, (t >> 1) & 1, a.err, 14, dead);
            __syncwarp();
            tc_fence_after();
            C3S();
            // pass A: minima of groups of 8 columns -> sorted list of the WL smallest group minima
#pragma unroll 1
            for (int ch = 0; ch < NC / 32; ++ch) {
                float v[32];
                tc_ld32(trow + buf * NC + ch * 32, v);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float mn = Limits<float>::inf();
#pragma unroll
                    for (int i = 0; i < 8; ++i) mn = fminf(mn, fmaf(-2.f, v[8 * g + i], nrm[ch * 32 + 8 * g + i]));
                    if (__any_sync(0xffffffffu, mn < lst[WL - 1])) {
                        float x = mn;
#pragma unroll
                        for (int i = 0; i < WL; ++i) {
                            const float lo = fminf(lst[i], x);
                            x = fmaxf(lst[i], x);
                            lst[i] = lo;
                        }
                    }
                }
            }
            C3S();
            // pass B: every column with s <= bound + 2E is a candidate (same arithmetic as pass A)
            const float cut = fminf(lst[WL - 1] + margin, 3.402823466e+38f);
            bool compacted = false;
#pragma unroll 1
            for (int ch = 0; ch < NC / 32; ++ch) {
                // A lane whose slots could run out within this chunk: every lane re-filters its candidates
                // with the current (tighter) bound; a lane that overflows nevertheless (heavy ties) is redone.
                // (Once per tile: the bound only moves between tiles.)
                if (!compacted && __any_sync(0xffffffffu, cnt > CAP - 32 && cnt <= CAP)) {
                    compacted = true;
                    int n = 0;
                    const int old = min(cnt, CAP);
#pragma unroll 4
                    for (int i = 0; i < CAP; ++i) {
                        if (i < old) {
                            const float s = cand_s[i * CSTR + row];
                            const int c = cand_c[i * CSTR + row];
                            if (s <= cut) {
                                cand_s[n * CSTR + row] = s;
                                cand_c[n * CSTR + row] = c;
                                ++n;
                            }
                        }
                    }
                    cnt = cnt > CAP ? cnt : n;
                }
                float v[32];
                tc_ld32(trow + buf * NC + ch * 32, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float s = fmaf(-2.f, v[i], nrm[ch * 32 + i]);
                    if (s <= cut) {
                        if (cnt < CAP) {
                            cand_s[cnt * CSTR + row] = s;
                            cand_c[cnt * CSTR + row] = t * NC + ch * 32 + i;
                        }
                        ++cnt;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
            C3S();
        }
        // final filter with the final bound: the surviving cells go to the query's candidate row in global
        // memory, the exact re-rank is a kernel of its own (one warp per query over the whole GPU)
        const int64_t q = q0 + row;
        if (q < a.nq) {
            const float thr = fminf(lst[WL - 1] + margin, 3.402823466e+38f);
            int n = 0;
            if (cnt <= CAP) {
                for (int i = 0; i < cnt; ++i)
                    if (cand_s[i * CSTR + row] <= thr) a.cand_out[q * CAP + n++] = cand_c[i * CSTR + row];
            }
            a.cnt_out[q] = (cnt > CAP || a.force_redo) ? -1 : n;
        }
        C3S();
#ifdef C3_STAMP
        if (blockIdx.x == 0 && tid == 0) {
            printf("coarse3 stamps (cnt %d):", cnt);
            for (int i = 1; i < nst; ++i) printf(" %lld", stamps[i] - stamps[i - 1]);
            printf("  total %lld\n", stamps[nst - 1] - stamps[0]);
        }
#endif
    }
    tc_fence_before();
    __syncthreads();
    if (wid == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// Exact re-rank: one warp per query, lane = candidate.  The oracle's chain sum_d (c_d - q_d)^2 (c - q, ascending
// d, one fma per dim), then w rounds of warp arg-min by (distance, cell) = stable sortperm order.
constexpr int RR_WARPS = 8;
__global__ void __launch_bounds__(RR_WARPS * 32) coarse3_rerank_kernel(const Args a) {
    __shared__ int scratch_s[RR_WARPS][CAP];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t q = (int64_t)blockIdx.x * RR_WARPS + wid;
    if (q >= a.nq) return;
    const int w = a.w, D4 = a.D >> 2;
    const int total = a.cnt_out[q];
    if (total < w) {  // overflow (-1); fewer than w cannot happen (the candidates contain the top w), never return garbage
        if (lane == 0) a.redo[q] = 1;
        return;
    }
    if (lane == 0) a.redo[q] = 0;
    int* scratch = scratch_s[wid];
#pragma unroll
    for (int h = 0; h < CAP / 32; ++h)
        if (lane + 32 * h < total) scratch[lane + 32 * h] = a.cand_out[q * CAP + lane + 32 * h];
    __syncwarp();
    const float4* qrow = reinterpret_cast<const float4*>(a.Q + (size_t)q * a.D);
    unsigned long long key[CAP / 32];
#pragma unroll
    for (int h = 0; h < CAP / 32; ++h) {
        key[h] = ~0ull;
        if (32 * h < total) {  // warp-uniform
            const int ci = lane + 32 * h;
            const int cell = ci < total ? scratch[ci] : scratch[0];
            const float4* crow = reinterpret_cast<const float4*>(a.C + (size_t)cell * a.D);
            float acc = 0.f;
            for (int d0 = 0; d0 < D4; d0 += 8) {
                float4 c[8], qq[8];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (d0 + u < D4) { c[u] = __ldg(crow + d0 + u); qq[u] = __ldg(qrow + d0 + u); }
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (d0 + u < D4) {
                        float df = __fsub_rn(c[u].x, qq[u].x); acc = __fmaf_rn(df, df, acc);   // oracle A1: c - q, ascending d
                        df = __fsub_rn(c[u].y, qq[u].y); acc = __fmaf_rn(df, df, acc);
                        df = __fsub_rn(c[u].z, qq[u].z); acc = __fmaf_rn(df, df, acc);
                        df = __fsub_rn(c[u].w, qq[u].w); acc = __fmaf_rn(df, df, acc);
                    }
            }
            if (ci < total) key[h] = ((unsigned long long)__float_as_uint(acc) << 32) | (unsigned)cell;
        }
    }
    // distances are >= 0, so the order of the bit patterns is the order of the values
    unsigned long long mine = ~0ull;
    for (int r = 0; r < w; ++r) {
        unsigned long long best = key[0];
#pragma unroll
        for (int h = 1; h < CAP / 32; ++h) best = key[h] < best ? key[h] : best;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other < best ? other : best;
        }
#pragma unroll
        for (int h = 0; h < CAP / 32; ++h)
            if (key[h] == best) key[h] = ~0ull;  // (distance, cell) pairs are distinct
        if (lane == r) mine = best;
    }
    if (lane < w) {
        a.cells_out[q * w + lane] = (int32_t)(unsigned)(mine & 0xffffffffull);
        a.dc_out[q * w + lane] = __uint_as_float((unsigned)(mine >> 32));
    }
}

// Centroids -> B operand blocks [tile][k-step][2048 words] (TF32-rounded), once at create.
__global__ void prep_tcc_kernel(const float* __restrict__ C, int kc, int kcp, int D, int ksteps, float* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)kcp * ksteps) return;
    const int c = (int)(idx / ksteps), j = (int)(idx - (int64_t)c * ksteps);
    const int t = c / NC, n = c - t * NC;
    float* o = out + ((size_t)t * ksteps + j) * (BBLK / 4);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int d = 8 * j + k;
        const float v = (c < kc && d < D) ? C[(size_t)c * D + d] : 0.f;
        o[(n >> 3) * 64 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3)] = __uint_as_float(tf32_of(v));
    }
}
// Squared norms (+inf for the padding columns) and their maximum at cn[kcp] (zeroed by the caller).
__global__ void prep_cn_kernel(const float* __restrict__ C, int kc, int kcp, int D, float* __restrict__ cn) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= kcp) return;
    if (c >= kc) {
        cn[c] = Limits<float>::inf();
        return;
    }
    float s = 0.f;
    for (int d = 0; d < D; ++d) s = fmaf(C[(size_t)c * D + d], C[(size_t)c * D + d], s);
    cn[c] = s;
    atomicMax(reinterpret_cast<unsigned int*>(cn + kcp), __float_as_uint(s));  // s >= 0: bit order = value order
}

}  // namespace ctc
}  // namespace ivf
