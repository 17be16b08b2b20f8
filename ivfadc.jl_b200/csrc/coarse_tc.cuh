// K1 on the 5th-generation tensor cores: coarse assignment as a dense query x centroid contraction
// (tcgen05.mma kind::tf32, accumulators in tensor memory, centroid operand streamed by TMA bulk
// copies), fused with a per-query candidate selection, followed by an EXACT re-rank.
// Replaces coarse_search(::NaiveQuantizer, point, w), reference src/coarsequantizers.jl:33-37
// (colwise distances, stable sortperm, first w).
//
// The returned cells and distances are bit-identical to the oracle's direct form
// sum_d (c_d - q_d)^2 (one sequential fp32 fma chain per pair): the tensor cores only PRUNE.
//
//   1. score s(c) = |c|^2 - 2 q.c with q, c rounded to TF32 (ONE piece: no hi / lo split; |c|^2 enters as a
//      last k-step, ones x its two TF32 pieces, so the accumulator holds the finished score), fp32
//      accumulate: CTA = 128 queries (accumulator lanes) x all centroids in tiles of 256 (columns),
//      K = 8 dims per MMA.  Two accumulator tiles in tensor memory: the tensor core fills tile
//      t + 1 while the four epilogue warps (lane = query) read tile t.
//   2. |s(c) + |q|^2 - d(c)| <= E with E = 2^-10 (|q|^2 + max|c|^2) (TF32 rounding of both operands,
//      Cauchy-Schwarz; fp32 accumulation, the split norm and the chain's own rounding are 2^-12 of that).  The
//      kernel uses 2E = 2^-8 (|q|^2 + max|c|^2), twice the bound.
//   3. selection, branch-free per lane: the WL-th smallest (WL >= w) of the minima of groups of 8 / 16
//      columns seen so far is an upper bound B of the w-th smallest score; every centroid with
//      s <= B + 2E is a candidate.  If a centroid of the exact top-w had s > B + 2E, the >= w
//      centroids with s <= B would all have a strictly smaller exact distance -- so the candidates
//      are a superset of the exact top-w, ties included.  The tile is read twice from tensor
//      memory (minima, then filter): re-reading costs no shared-memory or HBM traffic.
//   4. exact re-rank (coarse3_rerank_kernel, persistent warps, one query at a time): the surviving centroid rows
//      are fetched coalesced into a shared tile, lane = candidate runs the oracle's fma chain over its row, the
//      ranks by (distance, cell) -- sortperm's stable order -- come from counting, lanes of rank < w write.
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
constexpr int RS_MAXC = 2048;           // keys within the selection bound of the wide redo kernel (<= w x kc / 256)
constexpr int RS_MAXQ = 32;             // flagged queries served by the wide redo kernel (one block per 256 centroids)
constexpr int GRP = 8;                  // columns per group of the bound in the first GRP_FINE_TILES tiles (16 afterwards)
constexpr int GRP_FINE_TILES = 2;
constexpr int CSTR = MQ + 1;            // slot stride (words) of the candidate arrays: conflict-free by row and by slot
constexpr int THREADS = 256;            // warps 0..3 epilogue (lane quarter = warp), 4 TMA producer, 5 MMA issuer
constexpr uint32_t SPIN = 1u << 22;     // bound on every mbarrier wait (error flag instead of a hang)
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

struct Smem {
    uint32_t a, b, cand_s, cand_c, stage, bars, total;
};
__host__ __device__ inline Smem smem_layout(int ksteps) {
    Smem s;
    uint32_t o = 0;
    s.a = o;       o += (uint32_t)(ksteps + 1) * ABLK;   // + the ones block that selects the norm k-step
    s.b = o;       o += NSLOT * BBLK;
    s.cand_s = o;  o += CAP * CSTR * 4;      // candidate scores  [slot][query row]
    s.cand_c = o;  o += CAP * CSTR * 4;      // candidate cells
    s.stage = o;   o += 4 * 32 * 32 * 4;     // per epilogue warp: the 32 x 32 scores of the chunk being filtered
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
// Issue only: the registers are written asynchronously until tc_wait32 (which names them, so that the
// compiler cannot move their uses in front of the wait).
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
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
}
__device__ __forceinline__ void tc_wait32(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
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
    int32_t* redo_list;    // [nq] flagged queries, compacted (filled by the re-rank kernel)
    int* redo_count;       // their number (zeroed by coarse3_kernel)
    unsigned* redo_done;   // [RS_MAXQ] block counters of coarse_redo_small_kernel (zeroed by coarse3_kernel)
    int force_redo;        // test switch: flag every query
};

// Minima of groups of G columns of a 32-column chunk -> sorted list of the WL smallest group minima
// (branch-free insertion; skipped when no lane of the warp would change its list).
template <int WL, int G>
__device__ __forceinline__ void c3_groups(const uint32_t (&r)[32], float (&lst)[WL]) {
#pragma unroll
    for (int g = 0; g < 32 / G; ++g) {
        float m0 = __uint_as_float(r[G * g]), m1 = __uint_as_float(r[G * g + 1]);
#pragma unroll
        for (int i = 2; i < G; i += 2) {
            m0 = fminf(m0, __uint_as_float(r[G * g + i]));
            m1 = fminf(m1, __uint_as_float(r[G * g + i + 1]));
        }
        const float mn = fminf(m0, m1);
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
    float* cand_s = reinterpret_cast<float*>(smem_c3 + L.cand_s);
    int* cand_c = reinterpret_cast<int*>(smem_c3 + L.cand_c);
    const int64_t q0 = (int64_t)blockIdx.x * MQ;
    const int ntiles = a.kcp / NC;
    const int KS = a.ksteps;
    const int KB = KS + 1;  // k-steps per tile: the dims, then the squared norms (ones block x split norms)
    bool dead = false;
    if (blockIdx.x == 0 && tid == 0) *a.redo_count = 0;  // the re-rank kernel (next in the stream) appends the flagged queries
    if (blockIdx.x == 0 && tid < RS_MAXQ) a.redo_done[tid] = 0u;
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
    __syncthreads();

    // ---- the centroid operand starts streaming while the query operand is written ----
    if (wid == 4 && lane == 0) {
        const int total = ntiles * KB;
        for (int it = 0; it < total && it < NSLOT; ++it) {
            mbar_expect_tx(bar_full + 8 * it, BBLK);
            tma_bulk_g2s(sb + L.b + it * BBLK, a.tcC + (size_t)it * (BBLK / 4), BBLK, bar_full + 8 * it);
        }
    }
    // ---- A operand: the 128 query rows rounded to TF32, block j = dims 8 j .. 8 j + 7;
    //      word(n, k) = (n >> 3) * 64 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3)
    {
        const int nchunk = MQ * 2 * KS;  // 16-byte chunks (4 dims of one row)
        constexpr int UNR = 16;          // loads in flight per thread (D = 128: the whole share of a thread in one batch)
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
                const uint32_t dst = sb + L.a + (k4 >> 1) * ABLK + (n >> 3) * 256 + (k4 & 1) * 128 + (n & 7) * 16;
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(tf32_of(v[u].x)), "r"(tf32_of(v[u].y)),
                             "r"(tf32_of(v[u].z)), "r"(tf32_of(v[u].w))
                             : "memory");
            }
        }
    }
    // ones block (k slots 0, 1 select the two TF32 pieces of |c|^2)
    for (int i = tid; i < MQ * 2; i += THREADS) {
        const int n = i >> 1, half = i & 1;
        const uint32_t one = half == 0 ? 0x3f800000u : 0u;
        const uint32_t dst = sb + L.a + KS * ABLK + (n >> 3) * 256 + half * 128 + (n & 7) * 16;
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(one), "r"(one), "r"(0u), "r"(0u) : "memory");
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
            const int total = ntiles * KB;
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
                for (int j = 0; j < KB; ++j, ++it) {
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
        const float cmax2 = __ldg(a.cn + a.kcp);
        float qn = 0.f;  // |q|^2 of this lane's query (the rows are L1 / L2-hot: the A operand was just built from them)
        if (q0 + row < a.nq) {
            const float4* qr = reinterpret_cast<const float4*>(a.Q + (size_t)(q0 + row) * a.D);
            float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll 8
            for (int d4 = 0; d4 < (a.D >> 2); ++d4) {
                const float4 x = __ldg(qr + d4);
                p0 = fmaf(x.x, x.x, p0); p1 = fmaf(x.y, x.y, p1); p2 = fmaf(x.z, x.z, p2); p3 = fmaf(x.w, x.w, p3);
            }
            qn = (p0 + p1) + (p2 + p3);
        }
        const float margin = (qn + cmax2) * 0.00390625f;  // 2E = 2^-8 (|q|^2 + max |c|^2)
        float lst[WL];
#pragma unroll
        for (int i = 0; i < WL; ++i) lst[i] = Limits<float>::inf();
        int cnt = 0;
        const uint32_t trow = tmem_base + ((uint32_t)(32 * wid) << 16);
        // one group of 8 scores: minimum -> sorted list of the WL smallest group minima (branch-free insertion)
        float* stage = reinterpret_cast<float*>(smem_c3 + L.stage) + wid * (32 * 32);  // [column of the chunk][lane]
        for (int t = 0; t < ntiles; ++t) {
            const int buf = t & 1;
            C3S();
            if (lane == 0) mbar_wait(bar_tfull + 8 * buf, (t >> 1) & 1, a.err, 14, dead);
            __syncwarp();
            tc_fence_after();
            C3S();
            const uint32_t tt = trow + buf * NC;
            uint32_t va[32], vb[32];
            // pass A: the tile's group minima.  Loads run one chunk ahead of the arithmetic.
            tc_ld32_issue(tt, va);
#pragma unroll 1
            for (int ch = 0; ch < NC / 32; ch += 2) {
                // groups of 8 columns in the first tiles (a tight bound early keeps the candidate rows short),
                // of 16 later (half the insertions; the bound moves little by then)
                tc_wait32(va);
                tc_ld32_issue(tt + (ch + 1) * 32, vb);
                if (t < GRP_FINE_TILES) c3_groups<WL, 8>(va, lst); else c3_groups<WL, 16>(va, lst);
                tc_wait32(vb);
                tc_ld32_issue(tt + ((ch + 2) & (NC / 32 - 1)) * 32, va);  // wraps to chunk 0: the first load of pass B
                if (t < GRP_FINE_TILES) c3_groups<WL, 8>(vb, lst); else c3_groups<WL, 16>(vb, lst);
            }
            C3S();
            // pass B: every column with s <= bound + 2E is a candidate
            const float cut = fminf(lst[WL - 1] + margin, 3.402823466e+38f);
            // A lane whose slots could run out within this tile: every lane re-filters its candidates with the
            // current (tighter) bound; a lane that overflows nevertheless (heavy ties) is redone by the FFMA kernel.
            if (__any_sync(0xffffffffu, cnt > CAP - 24 && cnt <= CAP)) {
                int n = 0;
                const int old = min(cnt, CAP);
                int wmax = old;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
                for (int i = 0; i < wmax; ++i) {
                    if (i < old) {
                        const float sc = cand_s[i * CSTR + row];
                        const int c = cand_c[i * CSTR + row];
                        if (sc <= cut) {
                            cand_s[n * CSTR + row] = sc;
                            cand_c[n * CSTR + row] = c;
                            ++n;
                        }
                    }
                }
                cnt = cnt > CAP ? cnt : n;
            }
            // The chunk's scores go to shared memory (32 independent stores) while a pass mask is built with
            // independent compares; the few set bits are then appended from there (no dependent chain per column).
            auto filter = [&](const uint32_t (&r)[32], int col0) {
                uint32_t m = 0;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    stage[i * 32 + lane] = __uint_as_float(r[i]);
                    m |= (__uint_as_float(r[i]) <= cut ? 1u : 0u) << i;
                }
                while (m) {  // own stores only: no barrier needed
                    const int i = __ffs(m) - 1;
                    m &= m - 1;
                    if (cnt < CAP) {
                        cand_s[cnt * CSTR + row] = stage[i * 32 + lane];
                        cand_c[cnt * CSTR + row] = col0 + i;
                    }
                    ++cnt;
                }
            };
#pragma unroll 1
            for (int ch = 0; ch < NC / 32; ch += 2) {
                tc_wait32(va);
                tc_ld32_issue(tt + (ch + 1) * 32, vb);
                filter(va, t * NC + ch * 32);
                tc_wait32(vb);
                if (ch + 2 < NC / 32) tc_ld32_issue(tt + (ch + 2) * 32, va);
                filter(vb, t * NC + (ch + 1) * 32);
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
                    if (cand_s[i * CSTR + row] <= thr && cand_c[i * CSTR + row] < a.kc) a.cand_out[q * CAP + n++] = cand_c[i * CSTR + row];
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

// Exact re-rank: one warp per query.  The candidate centroid rows are fetched COALESCED (one cp.async request
// per row: every lane copies 16 bytes of it) into a padded shared-memory tile, then lane = candidate runs the
// oracle's chain sum_d (c_d - q_d)^2 (c - q, ascending d, one fma per dim) from shared memory; w rounds of warp
// arg-min by (distance, cell) give sortperm's stable order.
constexpr int RR_WARPS = 4;
__host__ __device__ inline int rr_stride(int D) { return D + 4; }                       // floats per staged row (conflict-free LDS.128 by lane)
__host__ __device__ inline size_t rr_warp_floats(int D) { return (size_t)32 * rr_stride(D) + D + CAP; }
__global__ void __launch_bounds__(RR_WARPS * 32) coarse3_rerank_kernel(const Args a) {
    extern __shared__ __align__(16) float rr_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int w = a.w, D = a.D, D4 = D >> 2, stride = rr_stride(D);
    float* rows = rr_smem + (size_t)wid * rr_warp_floats(D);
    float* qs = rows + 32 * stride;
    int* scratch = reinterpret_cast<int*>(qs + D);
    const uint32_t rows_u = smem_u32(rows), qs_u = smem_u32(qs);
    // persistent warps: candidate row and count of the NEXT query are fetched (one round trip, together)
    // while the current one is ranked
    const int64_t qstep = (int64_t)gridDim.x * RR_WARPS;
    int64_t q = (int64_t)blockIdx.x * RR_WARPS + wid;
    int cand_n[CAP / 32], total_n = 0;
    if (q < a.nq) {
#pragma unroll
        for (int h = 0; h < CAP / 32; ++h) cand_n[h] = __ldg(a.cand_out + q * CAP + lane + 32 * h);
        total_n = __ldg(a.cnt_out + q);
    }
    for (; q < a.nq; q += qstep) {
    int cand_r[CAP / 32];
#pragma unroll
    for (int h = 0; h < CAP / 32; ++h) cand_r[h] = cand_n[h];
    const int total = total_n;
    if (q + qstep < a.nq) {
#pragma unroll
        for (int h = 0; h < CAP / 32; ++h) cand_n[h] = __ldg(a.cand_out + (q + qstep) * CAP + lane + 32 * h);
        total_n = __ldg(a.cnt_out + q + qstep);
    }
    if (total < w) {  // overflow (-1); fewer than w cannot happen (the candidates contain the top w), never return garbage
        if (lane == 0) {
            a.redo[q] = 1;
            a.redo_list[atomicAdd(a.redo_count, 1)] = (int32_t)q;
        }
        continue;
    }
    if (lane == 0) a.redo[q] = 0;
    __syncwarp();  // the previous query's reads of scratch / qs are done
#pragma unroll
    for (int h = 0; h < CAP / 32; ++h)
        if (lane + 32 * h < total) scratch[lane + 32 * h] = cand_r[h];
    if (lane < D4)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(qs_u + lane * 16), "l"(a.Q + (size_t)q * D + lane * 4) : "memory");
    __syncwarp();
    unsigned long long key[CAP / 32];
#pragma unroll
    for (int h = 0; h < CAP / 32; ++h) {
        key[h] = ~0ull;
        if (32 * h < total) {  // warp-uniform
            const int nb = min(32, total - 32 * h);
            if (lane < D4) {
                for (int r = 0; r < nb; ++r) {
                    const int cell = scratch[32 * h + r];
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(rows_u + (r * stride + lane * 4) * 4),
                                 "l"(a.C + (size_t)cell * D + lane * 4)
                                 : "memory");
                }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
            if (lane < nb) {
                const float4* crow = reinterpret_cast<const float4*>(rows + lane * stride);
                const float4* qrow = reinterpret_cast<const float4*>(qs);
                float acc = 0.f;
#pragma unroll 8
                for (int d4 = 0; d4 < D4; ++d4) {
                    const float4 c = crow[d4], qq = qrow[d4];
                    float df = __fsub_rn(c.x, qq.x); acc = __fmaf_rn(df, df, acc);   // oracle A1: c - q, ascending d
                    df = __fsub_rn(c.y, qq.y); acc = __fmaf_rn(df, df, acc);
                    df = __fsub_rn(c.z, qq.z); acc = __fmaf_rn(df, df, acc);
                    df = __fsub_rn(c.w, qq.w); acc = __fmaf_rn(df, df, acc);
                }
                key[h] = ((unsigned long long)__float_as_uint(acc) << 32) | (unsigned)scratch[32 * h + lane];
            }
            __syncwarp();  // the tile is rewritten by the next batch
        }
    }
    // Rank by counting (distances are >= 0, so the order of the bit patterns is the order of the values; the
    // (distance, cell) keys are distinct): independent shuffles, no reduction chain.
    int rank[CAP / 32];
#pragma unroll
    for (int h = 0; h < CAP / 32; ++h) rank[h] = 0;
    const int nh = (total + 31) >> 5;
    const unsigned dbits = (unsigned)(key[0] >> 32);
    const unsigned same = __match_any_sync(0xffffffffu, dbits);  // every lane takes part
    const bool fast = nh == 1 && __all_sync(0xffffffffu, key[0] == ~0ull || __popc(same) == 1);
    if (fast) {  // up to 32 candidates, no two equal distances: compare the distance bits alone
        for (int j = 0; j < 32; ++j) rank[0] += __shfl_sync(0xffffffffu, dbits, j) < dbits ? 1 : 0;
    } else {
        for (int j = 0; j < 32; ++j) {
#pragma unroll
            for (int g = 0; g < CAP / 32; ++g) {
                if (g < nh) {  // warp-uniform
                    const unsigned long long other = __shfl_sync(0xffffffffu, key[g], j);
#pragma unroll
                    for (int h = 0; h < CAP / 32; ++h) rank[h] += other < key[h] ? 1 : 0;
                }
            }
        }
    }
#pragma unroll
    for (int h = 0; h < CAP / 32; ++h) {
        if (key[h] != ~0ull && rank[h] < w) {
            a.cells_out[q * w + rank[h]] = (int32_t)(unsigned)(key[h] & 0xffffffffull);
            a.dc_out[q * w + rank[h]] = __uint_as_float((unsigned)(key[h] >> 32));
        }
    }
    }  // queries of this warp
}

// Centroids -> B operand blocks [tile][k-steps + 1][2048 words], once at create: TF32(-2 c) per k-step of 8 dims,
// then the norm block (k slot 0, 1 = the two TF32 pieces of |c|^2; 1e30 for the padding columns).
__global__ void prep_tcc_kernel(const float* __restrict__ C, int kc, int kcp, int D, int ksteps, float* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int KB = ksteps + 1;
    if (idx >= (int64_t)kcp * KB) return;
    const int c = (int)(idx / KB), j = (int)(idx - (int64_t)c * KB);
    const int t = c / NC, n = c - t * NC;
    float* o = out + ((size_t)t * KB + j) * (BBLK / 4);
    if (j < ksteps) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int d = 8 * j + k;
            const float v = (c < kc && d < D) ? -2.f * C[(size_t)c * D + d] : 0.f;
            o[(n >> 3) * 64 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3)] = __uint_as_float(tf32_of(v));
        }
    } else {
        float nrm = 1e30f, lo = 0.f;
        if (c < kc) {
            nrm = 0.f;
            for (int d = 0; d < D; ++d) nrm = fmaf(C[(size_t)c * D + d], C[(size_t)c * D + d], nrm);
            const float hi = __uint_as_float(tf32_of(nrm));
            lo = __uint_as_float(tf32_of(nrm - hi));
            nrm = hi;
        } else {
            nrm = __uint_as_float(tf32_of(nrm));
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
            o[(n >> 3) * 64 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3)] = k == 0 ? nrm : k == 1 ? lo : 0.f;
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
