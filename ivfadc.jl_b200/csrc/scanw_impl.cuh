// K2 + K3, warp-specialised "tensor-memory lookup" scan -- the default batched search kernel on sm_100a
// (fp32, k <= 16, dsub <= 8 or dsub = 16, m in {4, 8, 12, 16}).
//
// Same arithmetic and the same work decomposition as scanu_impl.cuh (work item = one inverted list x up to 32
// queries that probe it, lane = query, the lookup table of one subspace is an accumulator tile
// D[128 lanes][256 columns] in tensor memory built by four tcgen05.mma kind::tf32, a lookup is
// tcgen05.ld.32x32b.x1 at column = code byte; reference: the LUT of src/index.jl:232-236 and the scan of
// src/index.jl:240-246), but the CTA is split into ROLES that only meet on mbarriers -- no CTA-wide barrier
// anywhere between the prologue and the epilogue of the persistent kernel:
//
//   * warps 0..11  SCANNERS  (152 registers via setmaxnreg): 96 partial distances per lane, so a pass covers
//     12 x 96 = 1152 vectors of the list.  Per table: wait for the build (mbarrier of tcgen05.commit), 96 lookups,
//     arrive on the table buffer's "free" mbarrier -- and straight on to the next table.  At the end of a pass:
//     two group minima per warp -> the k-th smallest of the 24 minima bounds the list's k-th distance (two
//     named barriers among the 384 scanner threads) -> every vector within the bound goes to the query's row of a
//     shared-memory staging area (slot from a per-query shared-memory counter).
//   * warp 12      ISSUER: the stream of table builds t = 0, 1, 2, ... runs across segment and item boundaries.
//     Per build only: wait for the A operand, for the 12 scanners to have released table t - 2 and for the
//     codebook operand, then four MMAs + commit and the refill of the codebook ring (cp.async.bulk, three slots).
//     Nothing else is on the path from "table buffer released" to "next build issued" (with the operand writes in
//     the same warp that path was 1.6 k cycles per table: the round's first timeline).
//   * warp 15      OPERAND WRITER: A operand of build t (TF32 hi / lo of the residuals, four row copies) into its
//     ring slot as soon as build t - 2 has completed.
//   * warps 13, 14 LOADERS: next segment (item, pass) while the scanners work on the current one -- item
//     descriptor, query rows / centroid slices / code bytes by cp.async, byte-plane transposes of the codes
//     (double-buffered planes), residuals r = q - c with their norms (double-buffered per item).  Before a
//     segment's buffers are reused they FINALIZE the segment that used them: staged candidates -> the pairs'
//     candidate rows in HBM, and after the last pass of a list the candidate counts of the item's pairs (or the
//     pair is queued for the exact redo kernel on overflow).
//
// Candidate rows, merge_cands_kernel and the redo queue are those of scanu (scan.cu).
#pragma once

#include <cuda_fp16.h>

#include "scanu_impl.cuh"

namespace ivf {

// Two shapes of the CTA (scan.cu chooses per index; IVFADC_SCANW_SHAPE=12|16 in the environment pins one for A/B runs):
//   12 scanners x 96 partial distances (152 registers), 1152 vectors per pass, 512 threads
//   16 scanners x 64 partial distances (112 registers), 1024 vectors per pass, 640 threads -- four scanning warps
//      per scheduler instead of three: a warp's table round is a chain of lookup batches (32 in flight, then
//      tcgen05.wait::ld), so the issue slots fill with the number of warps that interleave their round trips.
//      Register budget: setmaxnreg.inc is served from what the CTA's own warps released with setmaxnreg.dec (not
//      from unallocated registers of the SM): 640 threads launch with 96; the four service warps drop to 32 and
//      free 4 x 32 x 64 = 8192 = 16 x 32 x (112 - 96).  (120 / 32 and 112 / 56 wait forever in setmaxnreg.inc.)
// bring-up timing experiments (wrong results): IVF_X_NOLOOKUP skips the lookups, IVF_X_NOEXTRACT the end-of-pass selection
#ifndef IVF_X_NOLOOKUP
#define IVF_X_NOLOOKUP 0
#endif
#ifndef IVF_X_NOEXTRACT
#define IVF_X_NOEXTRACT 0
#endif
#ifndef IVF_W_HINT_NS
#define IVF_W_HINT_NS 100000u   // suspend-time hint of mbarrier.try_wait (ns)
#endif
#ifndef IVF_W_PIPE
#define IVF_W_PIPE 1   // table switch inside a lookup batch (full warps); -DIVF_W_PIPE=0 keeps one drain per table
#endif
// Both shapes are compiled (template parameter WS = scanning warps); scan.cu picks one per index from the list lengths.
template <int WS> struct WShape {
    static_assert(WS == 12 || WS == 16, "12 x 96 or 16 x 64");
    static constexpr int SCAN = WS;                      // scanning warps
    static constexpr int NV = WS == 16 ? 64 : 96;        // partial distances per lane
    static constexpr int THREADS = (WS + 4) * 32;        // + issuer, two loaders, operand writer
    static constexpr int ISSUE = WS, LOAD = WS + 1, OPER = WS + 3;
    static constexpr int NCH = NV / 16;                  // chunks of 16 vectors per scanner and pass
    static constexpr int VP = SCAN * NV;                 // vectors per pass: 1152 / 1024
    static constexpr int CSTEP = 16 * SCAN;              // byte distance of a scanner's consecutive chunks in a plane
    static constexpr int NMIN = 2 * SCAN;                // group minima per query and pass
};
// the names the kernel body uses, bound to the shape of the instantiation
#define IVF_W_SHAPE_ALIASES(WS)                                                                           \
    constexpr int W_SCAN = WShape<WS>::SCAN, W_NV = WShape<WS>::NV, W_ISSUE = WShape<WS>::ISSUE,          \
                  W_LOAD = WShape<WS>::LOAD, W_OPER = WShape<WS>::OPER, W_NCH = WShape<WS>::NCH,          \
                  W_VP = WShape<WS>::VP, W_CSTEP = WShape<WS>::CSTEP, W_NMIN = WShape<WS>::NMIN;          \
    (void)W_SCAN; (void)W_NV; (void)W_ISSUE; (void)W_LOAD; (void)W_OPER; (void)W_NCH; (void)W_VP; (void)W_CSTEP; (void)W_NMIN
constexpr int W_NLOAD = 64;                  // loader threads
constexpr int W_STATE = 8 * QG * 4;          // per item: pair | dc -> base 2^s | run | cnt | flag | 2^sr | a | 2^-s
constexpr int W_SEG = 32;                    // per segment: valid, nv, pass, ipar, nj, last
constexpr int W_CAND = QG * U_CAP * 4;       // one plane (distances or positions) of the staged candidates of a pass
constexpr int W_ABLK = 4096;                 // A block: 128 rows x 16 k (fp16)
constexpr int W_ASUB = 2 * W_ABLK;           // (hi | lo), (hi | a a 0..) of one table
constexpr int W_BBLK = 8192;                 // B block: 256 rows x 16 k (fp16)
constexpr int W_BSUB = 2 * W_BBLK;           // (hi | hi), (lo | norm pieces) of one table
constexpr int W_NB = 4;                      // B ring depth
// Instruction descriptor of tcgen05.mma kind::f16: D fp32, A/B fp16 K-major, M = 128, N = 256 (K = 16).
constexpr uint32_t W_IDESC = (1u << 4) | (0u << 7) | (0u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t W_SPIN = 1u << 20;

struct ScanWSmem {
    uint32_t aring, bring, planes, raw, resid, rawq, cand, smin, thr, rnorm, rmax, state, seg, segcnt, ldesc, cbuf, bars, total;
};

// m = tables per item (8 dims each), mc = code bytes per vector
__host__ __device__ inline ScanWSmem scanw_smem_layout(int m, int mc, int ws) {
    const int W_VP = ws == 16 ? WShape<16>::VP : WShape<12>::VP, W_NMIN = 2 * ws;
    ScanWSmem s;
    uint32_t o = 0;
    s.aring = o;   o += 2 * W_ASUB;
    s.bring = o;   o += W_NB * W_BSUB;
    s.planes = o;  o += 2u * (uint32_t)mc * W_VP;
    s.raw = o;     o += (uint32_t)mc * W_VP + 16;
    s.rawq = o;    o += (uint32_t)QG * m * 32;
    s.resid = o;   o += 2u * (uint32_t)m * 8 * T_RS * 4;
    o = (o + 15) & ~15u;
    s.cand = o;    o += 2 * W_CAND;             // [distances | positions][query][U_CAP]
    s.smin = o;    o += W_NMIN * QG * 4;
    s.thr = o;     o += QG * 4;
    s.rnorm = o;   o += (uint32_t)m * QG * 4;
    s.rmax = o;    o += (uint32_t)m * QG * 4;
    s.state = o;   o += 2 * W_STATE;
    s.seg = o;     o += 2 * W_SEG;
    s.segcnt = o;  o += 2 * QG * 4;            // candidates staged per query by the scanners of a segment
    s.ldesc = o;   o += 64;
    s.cbuf = o;    o += (uint32_t)m * 32;
    s.bars = o;    o += 160;
    s.total = o;
    return s;
}

__device__ __forceinline__ void sts_v4u(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
// 2^e as a float, e in [-126, 127]
__device__ __forceinline__ float exp2i(int e) { return __uint_as_float((uint32_t)(e + 127) << 23); }
__device__ __forceinline__ void named_bar(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait that gives up at once when another wait of this CTA has already timed out (`dead` flag in shared
// memory): a broken pipeline ends in an error code within about a second, not in a hang.  A waiting thread must
// not spin through issue slots the scanners need (the first version of this loop was 23% of the kernel's
// instructions, ncu): mbarrier.try_wait gets a suspend-time hint, so the thread sleeps in hardware until the phase
// completes (or the hint runs out), and the loop around it is four instructions.
__device__ __forceinline__ uint32_t mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(ns)
        : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait_w(uint32_t bar, uint32_t parity, uint32_t dead_u, int* err, int code) {
#pragma unroll 1
    for (uint32_t o = 0; o < (W_SPIN >> 8); ++o) {
#pragma unroll 1
        for (uint32_t i = 0; i < 256u; ++i)
            if (mbar_try_wait_hint(bar, parity, IVF_W_HINT_NS)) return;
        if (lds_u(dead_u) != 0u) return;
    }
    sts_u(dead_u, 1u);
    atomicExch(err, code);
}

// One table over the 96 slots of this warp.  tcgen05.wait::ld waits for ALL outstanding loads of the thread, so the
// number of lookups between two waits is the warp's depth in flight.  Measured (phase profile of the DBG
// instantiation + ncu): a lookup comes back after ~300 clocks under load, and with 16 in flight per warp the twelve
// scanners keep only ~190 lookups in the tensor-memory pipe -- 0.5 per clock of the 1.0 it sustains.  Two chunks
// (32 lookups) per wait.
template <bool FULL, int WS>
__device__ __forceinline__ void scanw_sub(uint32_t tb, uint32_t plane_w, int nch, float (&acc)[WShape<WS>::NV]) {
    IVF_W_SHAPE_ALIASES(WS);
#pragma unroll
    for (int jp = 0; jp < W_NCH / 2; ++jp) {
        const int j0 = 2 * jp, j1 = 2 * jp + 1;
        if (FULL || j1 < nch) {  // warp-uniform; slots beyond the list are masked after the last table
            const uint4 xa = lds_v4(plane_w + j0 * W_CSTEP), xb = lds_v4(plane_w + j1 * W_CSTEP);
            float t[32];
            scanu_issue<0>(tb, xa, t);
            scanu_issue<0>(tb, xb, t + 16);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; i += 2) add_pair(acc[16 * j0 + i], acc[16 * j0 + i + 1], t[i], t[i + 1]);
        } else if (j0 < nch) {
            const uint4 xa = lds_v4(plane_w + j0 * W_CSTEP);
            float t[16];
            scanu_issue<0>(tb, xa, t);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; i += 2) add_pair(acc[16 * j0 + i], acc[16 * j0 + i + 1], t[i], t[i + 1]);
        }
    }
}

// NP = (code bytes per vector) / 4; DUP = 2 serves dsub = 16 (two 8-dim tables share a code byte, see scanu_kernel).
// Debug buffer (DBG instantiation, CTA 0): tables of the first segment [m][256][32], int pair[32], cell, then
// the phase profile (WProf): 8 counters per warp.
// One table build = two MMAs (kind::f16, K = 16) + commit, issued by one elected lane.
__device__ __forceinline__ void tc_mma_f16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(W_IDESC), "r"(accumulate)
        : "memory");
}

// Phase profile of the DBG instantiation: every role accumulates the SM clocks it spends per phase (registers only,
// no memory traffic inside the loops); CTA 0 writes 8 counters per warp behind the table dump at the end.
template <bool ON>
struct WProf {
    uint32_t last, a[8];
    __device__ __forceinline__ void start() {
        if constexpr (ON) {
            last = (uint32_t)clock();
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = 0;
        }
    }
    __device__ __forceinline__ void tick(int i) {
        if constexpr (ON) {
            const uint32_t n = (uint32_t)clock();
            a[i] += n - last;
            last = n;
        }
    }
    __device__ __forceinline__ void store(long long* dst) {
        if constexpr (ON) {
            if (dst) {
#pragma unroll
                for (int i = 0; i < 8; ++i) dst[i] = a[i];
            }
        }
    }
};

template <int NP, bool DBG, int DUP, int WS>
__global__ void __launch_bounds__(WShape<WS>::THREADS, 1)
scanw_kernel(const ScanUArgs ua) {
    IVF_W_SHAPE_ALIASES(WS);
    const ScanQArgs& a = ua.q;
    extern __shared__ __align__(1024) unsigned char smem_w[];
    constexpr int mc = 4 * NP;      // code bytes per vector
    constexpr int m = mc * DUP;     // tables per segment
    constexpr uint32_t RESID_BYTES = (uint32_t)m * 8 * T_RS * 4;
    constexpr uint32_t PLANES_BYTES = (uint32_t)mc * W_VP;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform for ptxas
    const int k = a.k;
    const int ps = ua.pstride;
    const bool fastq = a.dsub == 8 && (a.D & 3) == 0;  // 16-byte aligned query / centroid slices
    const int nitems = a.group_off[a.kc];
    if ((int)blockIdx.x >= nitems) return;  // before any allocation

    uint32_t sb;
    asm volatile("mov.u32 %0, %1;" : "=r"(sb) : "r"(smem_u32(smem_w)));
    const ScanWSmem L = scanw_smem_layout(m, mc, WS);
    const uint32_t aring_u = sb + L.aring, bring_u = sb + L.bring, planes_u = sb + L.planes,
                   raw_u = sb + L.raw, resid_u = sb + L.resid, rawq_u = sb + L.rawq, smin_u = sb + L.smin,
                   thr_u = sb + L.thr, rnorm_u = sb + L.rnorm, rmax_u = sb + L.rmax, state_u = sb + L.state, seg_u = sb + L.seg,
                   ldesc_u = sb + L.ldesc, cbuf_u = sb + L.cbuf, cand_u = sb + L.cand, segcnt_u = sb + L.segcnt;
    const uint32_t bar_full = sb + L.bars;          // 4: codebook operand landed in ring slot
    const uint32_t bar_mma = bar_full + 32;         // 2: table build complete (tcgen05.commit)
    const uint32_t bar_free = bar_full + 48;        // 2: the 12 scanners are done with the table buffer
    const uint32_t bar_staged = bar_full + 64;      // 2: segment staged by the loaders
    const uint32_t bar_extract = bar_full + 80;     // 2: the 12 scanners have staged their candidates
    const uint32_t bar_a = bar_full + 96;           // 2: A operand of a build written
    const uint32_t bar_candfree = bar_full + 112;   // 1: staged candidates of a segment copied out
    const uint32_t tmem_slot = bar_full + 120, dead_u = bar_full + 124;
    const unsigned char* const tcH = static_cast<const unsigned char*>(ua.tcH);

    // ---- one-time setup ----
    if (tid == 0) {
        for (int i = 0; i < W_NB; ++i) mbar_init(bar_full + 8 * i, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_mma + 8 * i, 1);
            mbar_init(bar_free + 8 * i, W_SCAN);
            mbar_init(bar_staged + 8 * i, 1);
            mbar_init(bar_extract + 8 * i, W_SCAN);
            mbar_init(bar_a + 8 * i, 1);
        }
        mbar_init(bar_candfree, 1);
        sts_u(dead_u, 0u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < W_NB; ++i) {
            mbar_expect_tx(bar_full + 8 * i, W_BSUB);
            tma_bulk_g2s(bring_u + i * W_BSUB, tcH + (size_t)(i % m) * W_BSUB, W_BSUB, bar_full + 8 * i);
        }
    }
    if (wid == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "r"(U_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = lds_u(tmem_slot);

    auto warp_wait = [&](uint32_t bar, uint32_t parity, int code) {
        if (lane == 0) mbar_wait_w(bar, parity, dead_u, ua.err, code);
        __syncwarp();
    };
    long long* const stamps = (DBG && blockIdx.x == 0 && ua.dbg != nullptr)
                                  ? reinterpret_cast<long long*>(ua.dbg + (size_t)m * 256 * 32 + 64) : nullptr;

    // DBG: clocks of a window of four consecutive table builds in the middle of the launch (CTA 0), 8 events per warp
    // and table, as 32-bit words behind the 128 profile counters
    constexpr uint32_t W_TR0 = 20 * m + 6;
    const int drow = wid < 12 ? wid : wid >= W_SCAN ? wid - W_SCAN + 12 : -1;  // 12 scanners + the 4 service warps
    uint32_t* const trace = (stamps && drow >= 0) ? reinterpret_cast<uint32_t*>(stamps + 128) + drow * 32 : nullptr;
    auto tr = [&](uint32_t t, int ev) {
        if constexpr (DBG) {
            if (trace && lane == 0 && t - W_TR0 < 4u) trace[(t - W_TR0) * 8 + ev] = (uint32_t)clock();
        }
    };

    if (wid < W_SCAN) {
        // =========================================== SCANNERS ===========================================
        if constexpr (WS == 16) asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
        const uint32_t tq = tmem_base + ((uint32_t)((wid & 3) * 32) << 16);
        const uint32_t plane_w0 = planes_u + 16 * wid;
        WProf<DBG> pf;  // 0 wait staged, 1 wait table, 2 lookups, 3 release, 4 minima + barrier, 5 rank + barrier, 6 candidates
        pf.start();
        uint32_t tg = 0;
        float acc[W_NV];
#pragma unroll 1
        for (uint32_t g = 0;; ++g) {
            const uint32_t spar = g & 1;
            warp_wait(bar_staged + 8 * spar, (g >> 1) & 1, 20);
            const uint32_t sg = seg_u + spar * W_SEG;
            if (lds_u(sg) == 0u) break;
            const int nv = (int)lds_u(sg + 4), pass = (int)lds_u(sg + 8);
            const uint32_t stt = state_u + lds_u(sg + 12) * W_STATE;  // pair | base | run | cnt | flag of the item
            pf.tick(0);
            const float base = lds_f(stt + QG * 4 + lane * 4);
            // chunks j with 16 (wid + 12 j) < nv; provably warp-uniform (see below)
            const int nch = __shfl_sync(0xffffffffu, min(W_NCH, max(0, (nv - 16 * wid + W_CSTEP - 1) / W_CSTEP)), 0);
#pragma unroll
            for (int j = 0; j < W_NV; ++j) acc[j] = base;  // dc + |r|^2, then the table entries in subspace order
            // Everything that forms a lookup address must be PROVABLY warp-uniform for ptxas (shuffles from lane 0):
            // only then the code words are moved to uniform registers once (R2UR) and a lookup is UPRMT + LDTM;
            // otherwise every address is formed in a vector register and moved on its own (PRMT + R2UR + LDTM).
            // tg is a multiple of m (even), so table t = tg + s lives in buffer s & 1 and its phase flips every two tables.
            const uint32_t plane_seg = plane_w0 + spar * PLANES_BYTES;
            uint32_t par = (tg >> 1) & 1;
#if IVF_W_PIPE
            if (WS == 12 && nch == W_NCH) {   // (measured: the 16-scanner shape is faster with one drain per table)
                // Full warp (all its chunks inside the list): the table switch happens INSIDE a lookup batch.  The
                // last chunk of table s stays in flight while the warp releases nothing yet, acquires table s + 1
                // and issues its first chunk; one tcgen05.wait::ld covers both, then table s is released.  The
                // fixed cost of a switch (arrive, mbarrier wait, fences, address shuffles, code-byte loads) runs
                // under 16 lookups per warp instead of an idle tensor-memory pipe.
                float tp[16];
#pragma unroll 1
                for (int s = 0; s < m; ++s) {
                    const uint32_t b = (uint32_t)s & 1u;
                    pf.tick(3);
                    tr(tg + s, 0);
                    if (!mbar_try_wait_hint(bar_mma + 8 * b, par, IVF_W_HINT_NS)) warp_wait(bar_mma + 8 * b, par, 2);
                    tc_fence_after();
                    pf.tick(1);
                    tr(tg + s, 1);
                    const uint32_t tb = __shfl_sync(0xffffffffu, tq + b * 256, 0);
                    const uint32_t plane_w = __shfl_sync(0xffffffffu, plane_seg + (uint32_t)(s / DUP) * W_VP, 0);
                    if (DBG && g == 0 && blockIdx.x == 0 && ua.dbg != nullptr) {  // bring-up: dump the tables of the first segment
                        if (wid < 4) {
                            for (int c = wid; c < 256; c += 4) {
                                const float v = tc_ld1(tb + c);
                                tc_wait_ld();
                                ua.dbg[((size_t)s * 256 + c) * 32 + lane] = v * lds_f(stt + 7 * QG * 4 + lane * 4);
                            }
                        }
                        if (s == 0 && wid == 0) {
                            reinterpret_cast<int*>(ua.dbg + (size_t)m * 256 * 32)[lane] = (int)lds_u(stt + lane * 4);
                            if (lane == 0) reinterpret_cast<int*>(ua.dbg + (size_t)m * 256 * 32)[QG] = (int)lds_u(sg + 24);
                        }
                    }
                    {
                        float t0[16];
                        const uint4 x0 = lds_v4(plane_w);
                        scanu_issue<0>(tb, x0, t0);
                        tc_wait_ld();
                        if (s > 0) {  // the last chunk of table s - 1 (other buffer) has landed with it
#pragma unroll
                            for (int i = 0; i < 16; i += 2)
                                add_pair(acc[16 * (W_NCH - 1) + i], acc[16 * (W_NCH - 1) + i + 1], tp[i], tp[i + 1]);
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(bar_free + 8 * (b ^ 1u));
                            tr(tg + s - 1, 3);
                        }
#pragma unroll
                        for (int i = 0; i < 16; i += 2) add_pair(acc[i], acc[i + 1], t0[i], t0[i + 1]);
                    }
#pragma unroll
                    for (int jp = 0; jp < (W_NCH - 2) / 2; ++jp) {
                        const int j0 = 1 + 2 * jp, j1 = 2 + 2 * jp;
                        const uint4 xa = lds_v4(plane_w + j0 * W_CSTEP), xb = lds_v4(plane_w + j1 * W_CSTEP);
                        float t[32];
                        scanu_issue<0>(tb, xa, t);
                        scanu_issue<0>(tb, xb, t + 16);
                        tc_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; i += 2) add_pair(acc[16 * j0 + i], acc[16 * j0 + i + 1], t[i], t[i + 1]);
                    }
                    {
                        const uint4 xl = lds_v4(plane_w + (W_NCH - 1) * W_CSTEP);
                        scanu_issue<0>(tb, xl, tp);  // stays in flight across the switch
                    }
                    pf.tick(2);
                    tr(tg + s, 2);
                    par ^= b;
                }
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; i += 2)
                    add_pair(acc[16 * (W_NCH - 1) + i], acc[16 * (W_NCH - 1) + i + 1], tp[i], tp[i + 1]);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_free + 8 * ((uint32_t)(m - 1) & 1u));
            } else
#endif
#pragma unroll 1
            for (int s = 0; s < m; ++s) {
                const uint32_t b = (uint32_t)s & 1u;
                pf.tick(3);
                tr(tg + s, 0);
                // every lane polls (no divergence on the common path); the bounded loop only when the build is late
                if (!mbar_try_wait_hint(bar_mma + 8 * b, par, IVF_W_HINT_NS)) warp_wait(bar_mma + 8 * b, par, 2);
                tc_fence_after();
                pf.tick(1);
                tr(tg + s, 1);
                const uint32_t tb = __shfl_sync(0xffffffffu, tq + b * 256, 0);
                const uint32_t plane_w = __shfl_sync(0xffffffffu, plane_seg + (uint32_t)(s / DUP) * W_VP, 0);
#if IVF_X_NOLOOKUP
                acc[s & 15] += 1.0f + (float)wid;
#else
                if (nch == W_NCH) scanw_sub<true, WS>(tb, plane_w, nch, acc);
                else scanw_sub<false, WS>(tb, plane_w, nch, acc);
#endif
                pf.tick(2);
                tr(tg + s, 2);
                if (DBG && g == 0 && blockIdx.x == 0 && ua.dbg != nullptr) {  // bring-up: dump the tables of the first segment
                    if (wid < 4) {
                        for (int c = wid; c < 256; c += 4) {
                            const float v = tc_ld1(tb + c);
                            tc_wait_ld();
                            ua.dbg[((size_t)s * 256 + c) * 32 + lane] = v * lds_f(stt + 7 * QG * 4 + lane * 4);
                        }
                    }
                    if (s == 0 && wid == 0) {
                        reinterpret_cast<int*>(ua.dbg + (size_t)m * 256 * 32)[lane] = (int)lds_u(stt + lane * 4);
                        if (lane == 0) reinterpret_cast<int*>(ua.dbg + (size_t)m * 256 * 32)[QG] = (int)lds_u(sg + 24);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_free + 8 * b);
                tr(tg + s, 3);
                par ^= b;
            }
            tg += m;
            pf.tick(3);

            // ---- end of the pass: bound, candidates ----
            {   // slot 16 j + i holds vector 16 (wid + 12 j) + i of the pass: mask the slots beyond the list
                const int lim = nv - 16 * wid;
#pragma unroll
                for (int j = 0; j < W_NCH; ++j) {
                    if (W_CSTEP * j + 16 > lim) {  // warp-uniform
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (W_CSTEP * j + i >= lim) acc[16 * j + i] = Limits<float>::inf();
                    }
                }
            }
#if IVF_X_NOEXTRACT == 5
#pragma unroll
            for (int j = 0; j < W_NV; ++j) asm volatile("" ::"f"(acc[j]));   // keep the partial distances alive, nothing else
#endif
            float mn0 = Limits<float>::inf(), mn1 = Limits<float>::inf();
#pragma unroll
            for (int j = 0; j < ((IVF_X_NOEXTRACT == 1 || IVF_X_NOEXTRACT == 5) ? 2 : W_NV / 2); ++j) {
                mn0 = fminf(mn0, acc[j]);
                mn1 = fminf(mn1, acc[W_NV / 2 + j]);
            }
            sts_f(smin_u + ((2 * wid) * QG + lane) * 4, mn0);
            sts_f(smin_u + ((2 * wid + 1) * QG + lane) * 4, mn1);
            named_bar(1, W_SCAN * 32);
            pf.tick(4);
            {
                int r0 = 0, r1 = 0;
#pragma unroll
                for (int i = 0; i < W_NMIN; ++i) {
                    const float o = lds_f(smin_u + (i * QG + lane) * 4);
                    r0 += (o < mn0 || (o == mn0 && i < 2 * wid)) ? 1 : 0;
                    r1 += (o < mn1 || (o == mn1 && i < 2 * wid + 1)) ? 1 : 0;
                }
                // the k-th smallest of 24 disjoint group minima bounds the k-th smallest of all; the bound of the
                // previous passes of this list stays valid
                const int kk = min(k, W_NMIN) - 1;
                if (r0 == kk || r1 == kk) {
                    const float thr = fminf(r0 == kk ? mn0 : mn1, lds_f(stt + 2 * QG * 4 + lane * 4));
                    sts_f(thr_u + lane * 4, thr);
                    sts_f(stt + 2 * QG * 4 + lane * 4, thr);
                }
            }
            named_bar(1, W_SCAN * 32);
            pf.tick(5);
            {
                const float thr = lds_f(thr_u + lane * 4);
                // keep d <= bound; while fewer than k vectors have been seen (bound = +inf) keep every real slot
                const float cut = thr < Limits<float>::inf() ? thr : 3.402823466e+38f;
                const int pair = (int)lds_u(stt + lane * 4);
                const float unscale = lds_f(stt + 7 * QG * 4 + lane * 4);   // 2^-(ew + sr): exact
                int c = 0;
#pragma unroll
                for (int j = 0; j < ((IVF_X_NOEXTRACT == 1 || IVF_X_NOEXTRACT == 2 || IVF_X_NOEXTRACT == 5) ? 1 : W_NV); ++j) c += acc[j] <= cut ? 1 : 0;
                if (IVF_X_NOEXTRACT == 1 || IVF_X_NOEXTRACT == 2 || IVF_X_NOEXTRACT == 5) c = 0;
                if (pair < 0 || lds_u(stt + 4 * QG * 4 + lane * 4) != 0u) c = 0;
                // the staging area is single-buffered: the loaders have copied out the previous segment long ago
                if (g > 0) warp_wait(bar_candfree, (g - 1) & 1, 27);
                int old = 0;
                if (c > 0) old = atoms_add(segcnt_u + spar * (QG * 4) + lane * 4, c);
                const bool ok = c > 0 && old + c <= U_CAP && IVF_X_NOEXTRACT != 3;
                // more than a row can hold (the counter keeps the total): the loaders queue the pair for the redo kernel
                if (ok) {
                    uint32_t pd = cand_u + (uint32_t)(lane * U_CAP + old) * 4;
                    uint32_t pos = (uint32_t)pass * W_VP + 16u * wid;
#pragma unroll
                    for (int j = 0; j < W_NV; ++j) {
                        if (acc[j] <= cut) {
                            sts_f(pd, acc[j] * unscale);
                            sts_u(pd + W_CAND, pos + (uint32_t)((j >> 4) * W_CSTEP + (j & 15)));
                            pd += 4;
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_extract + 8 * spar);
            pf.tick(6);
        }
        pf.store(stamps && lane == 0 && drow >= 0 ? stamps + 8 * drow : nullptr);
    } else {
        if constexpr (WS == 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (wid == W_ISSUE) {
            // =========================================== ISSUER ===========================================
            const uint64_t descA0 = tc_smem_desc(aring_u), descB0 = tc_smem_desc(bring_u);
            uint32_t t = 0;
            uint32_t bslot = 0, bphase = 0;            // ring slot / phase of build t
            uint32_t fslot = 0, fsub = W_NB % m;       // next refill: slot (= slot of build t - 2), subspace of build t + 2
            uint32_t nfill = W_NB;                     // fills issued so far
            WProf<DBG> pf;  // 0 wait staged, 1 wait A operand (+ refill), 2 wait codebook operand, 3 wait release, 4 issue
            pf.start();
#pragma unroll 1
            for (uint32_t g = 0;; ++g) {
                const uint32_t spar = g & 1;
                warp_wait(bar_staged + 8 * spar, (g >> 1) & 1, 21);
                if (lds_u(seg_u + spar * W_SEG) == 0u) break;
                pf.tick(0);
#pragma unroll 1
                for (int s = 0; s < m; ++s, ++t) {
                    const uint32_t buf = t & 1;
                    tr(t, 0);
                    // A operand of build t written => build t - 2 has completed => its codebook ring slot is free
                    warp_wait(bar_a + 8 * buf, (t >> 1) & 1, 22);
                    tr(t, 1);
                    if (t >= 2) {
                        if (lane == 0) {
                            const uint32_t bar = bar_full + 8 * fslot;
                            mbar_expect_tx(bar, W_BSUB);
                            tma_bulk_g2s(bring_u + fslot * W_BSUB, tcH + (size_t)fsub * W_BSUB, W_BSUB, bar);
                        }
                        if (++fslot == W_NB) fslot = 0;
                        if (++fsub == (uint32_t)m) fsub = 0;
                        ++nfill;
                    }
                    const uint32_t slot_u = __shfl_sync(0xffffffffu, bslot, 0), buf_u = __shfl_sync(0xffffffffu, buf, 0);
                    const uint64_t A0 = descA0 + (uint64_t)(buf_u * (W_ASUB >> 4)), A1 = A0 + (W_ABLK >> 4);
                    const uint64_t B0 = descB0 + (uint64_t)(slot_u * (W_BSUB >> 4)), B1 = B0 + (W_BBLK >> 4);
                    const uint32_t d = tmem_base + buf_u * 256;
                    pf.tick(1);
                    tr(t, 2);
                    warp_wait(bar_full + 8 * bslot, bphase, 1);
                    pf.tick(2);
                    tr(t, 3);
                    if (t >= 2) warp_wait(bar_free + 8 * buf, ((t - 2) >> 1) & 1, 23);  // scanners released table t - 2
                    pf.tick(3);
                    tr(t, 4);
                    tc_fence_after();
                    tc_mma_f16_elect(d, A0, B0, 0);   // [rh | rl] . [wh | wh]
                    tc_mma_f16_elect(d, A1, B1, 1);   // [rh | a a 0..] . [wl | n0 n1 0..]
                    tc_commit_elect(bar_mma + 8 * buf_u);
                    tr(t, 5);
                    if (++bslot == W_NB) { bslot = 0; bphase ^= 1; }
                    pf.tick(4);
                }
            }
            pf.store(stamps && lane == 0 && drow >= 0 ? stamps + 8 * drow : nullptr);
            // drain: codebook operands fetched for builds that never ran
            if (lane == 0) {
                for (uint32_t f = t; f < nfill; ++f) mbar_wait_w(bar_full + 8 * (f % W_NB), (f / W_NB) & 1, dead_u, ua.err, 5);
            }
            __syncwarp();
        } else if (wid == W_OPER) {
            // =========================================== OPERAND WRITER ===========================================
            uint32_t t = 0;
            WProf<DBG> pf;  // 0 wait staged, 1 convert, 2 wait ring slot, 3 write
            pf.start();
#pragma unroll 1
            for (uint32_t g = 0;; ++g) {
                const uint32_t spar = g & 1;
                warp_wait(bar_staged + 8 * spar, (g >> 1) & 1, 24);
                const uint32_t sg = seg_u + spar * W_SEG;
                if (lds_u(sg) == 0u) break;
                const uint32_t res_i = resid_u + lds_u(sg + 12) * RESID_BYTES;
                const uint32_t stw = state_u + lds_u(sg + 12) * W_STATE;
                const float sc = lds_f(stw + 5 * QG * 4 + lane * 4);     // 2^sr of this lane's query
                const uint32_t aa = lds_u(stw + 6 * QG * 4 + lane * 4);  // (a, a) as two fp16
                pf.tick(0);
#pragma unroll 1
                for (int s = 0; s < m; ++s, ++t) {
                    const uint32_t buf = t & 1;
                    tr(t, 0);
                    // rows (copy, q): block 0 = [rh | rl], block 1 = [rh | a a 0 ..] with r 2^sr = rh + rl in fp16; lane = query
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int d = 0; d < 8; d += 2) {
                        const float r0 = lds_f(res_i + ((s * 8 + d) * T_RS + lane) * 4) * sc;
                        const float r1 = lds_f(res_i + ((s * 8 + d + 1) * T_RS + lane) * 4) * sc;
                        const __half h0 = __float2half_rn(r0), h1 = __float2half_rn(r1);
                        const __half l0 = __float2half_rn(r0 - __half2float(h0)), l1 = __float2half_rn(r1 - __half2float(h1));
                        hi[d >> 1] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                        lo[d >> 1] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
                    }
                    pf.tick(1);
                    tr(t, 1);
                    if (t >= 2) warp_wait(bar_mma + 8 * buf, ((t - 2) >> 1) & 1, 25);  // build t - 2 has read this ring slot
                    pf.tick(2);
                    tr(t, 2);
                    const uint32_t ph0 = aring_u + buf * W_ASUB + (lane >> 3) * 256 + (lane & 7) * 16;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {  // row = 32 c + lane: 8-row group stride 256, k halves 128 apart
                        const uint32_t ph = ph0 + c * 1024;
                        sts_v4u(ph, hi[0], hi[1], hi[2], hi[3]);
                        sts_v4u(ph + 128, lo[0], lo[1], lo[2], lo[3]);
                        sts_v4u(ph + W_ABLK, hi[0], hi[1], hi[2], hi[3]);
                        sts_v4u(ph + W_ABLK + 128, aa, 0u, 0u, 0u);
                    }
                    fence_proxy_async();  // generic-proxy writes of this lane -> async proxy
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_a + 8 * buf);
                    pf.tick(3);
                    tr(t, 3);
                }
            }
            pf.store(stamps && lane == 0 && drow >= 0 ? stamps + 8 * drow : nullptr);
        } else {
            // =========================================== LOADERS ===========================================
            const int lt = tid - W_LOAD * 32;  // 0..63
            const int lw = wid - W_LOAD;       // 0, 1
            auto cp_async = [&](uint32_t dst, const void* src, int bytes) {
                if (bytes == 16) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                else if (bytes == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
                else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
            };
            auto cp_wait = [&]() { asm volatile("cp.async.wait_all;" ::: "memory"); };
            WProf<DBG> pf;  // 0 finalize (wait + copy-out), 1 descriptor, 2 copies, 3 byte planes, 4 residuals + post
            pf.start();
            int item = blockIdx.x, pass = 0, npass = 1, ipar = 0, nj = 0, cell = 0;
            int64_t len = 0, off = 0;
            bool new_item = true;
            // Segment x is over (its scanners have staged their candidates): copy them into the pairs' candidate rows
            // -- row = [U_CAP distances][U_CAP positions], two loader threads per query -- and after the last pass of
            // the list publish the pairs.  Runs before the buffers of segment x are staged again (x + 2).
            auto finalize = [&](uint32_t x) {
                const uint32_t sp = x & 1;
                const uint32_t sgx = seg_u + sp * W_SEG;
                warp_wait(bar_extract + 8 * sp, (x >> 1) & 1, 26);
                const uint32_t stx = state_u + lds_u(sgx + 12) * W_STATE;
                const int njx = (int)lds_u(sgx + 16);
                const bool last = lds_u(sgx + 20) != 0u;
                const int q = lt & 31, half = lt >> 5;
                const int pair = (int)lds_u(stx + q * 4);
                const int n = (int)lds_u(segcnt_u + sp * (QG * 4) + q * 4);
                const int base = (int)lds_u(stx + 3 * QG * 4 + q * 4);
                const bool bad = lds_u(stx + 4 * QG * 4 + q * 4) != 0u || base + n > ps;
                if (pair >= 0 && !bad) {
                    uint32_t* row = reinterpret_cast<uint32_t*>(a.pair_d) + (size_t)pair * (2 * U_CAP) + base;
                    const uint32_t src = cand_u + (uint32_t)(q * U_CAP) * 4;
                    for (int i = half; i < (IVF_X_NOEXTRACT == 4 ? 0 : n); i += 2) {
                        row[i] = lds_u(src + i * 4);
                        row[U_CAP + i] = lds_u(src + W_CAND + i * 4);
                    }
                }
                named_bar(2, W_NLOAD);  // both halves have read the running count
                if (lt == 0) mbar_arrive(bar_candfree);
                if (half == 0) {
                    sts_u(stx + 3 * QG * 4 + q * 4, (uint32_t)(base + n));
                    if (bad) sts_u(stx + 4 * QG * 4 + q * 4, 1u);
                    if (last && q < njx && pair >= 0) {  // last pass of the list: publish the pair
                        if (bad) {
                            a.pair_cnt[pair] = 0;
                            a.redo_pairs[atomicAdd(a.redo_cnt, 1)] = pair;
                        } else {
                            a.pair_cnt[pair] = base + n;
                        }
                    }
                }
                named_bar(2, W_NLOAD);  // state / counters of the segment are free for the next staging
            };
#pragma unroll 1
            for (uint32_t g = 0;; ++g) {
                const uint32_t spar = g & 1;
                const uint32_t sg = seg_u + spar * W_SEG;
                pf.tick(4);
                if (g >= 2) finalize(g - 2);
                pf.tick(0);
                if (item >= nitems) {
                    if (lt == 0) {
                        sts_u(sg, 0u);
                        mbar_arrive(bar_staged + 8 * spar);
                    }
                    if (g >= 1) finalize(g - 1);
                    break;
                }
                if (lt < QG) sts_u(segcnt_u + spar * (QG * 4) + lt * 4, 0u);
                const uint32_t stt = state_u + ipar * W_STATE;
                const uint32_t res_i = resid_u + ipar * RESID_BYTES;
                if (new_item) {
                    if (lw == 0) {
                        // descriptor: (cell, first pair slot, number of pairs) -> list length / offset, the pairs
                        if (lane == 0) cp_async(ldesc_u, ua.items + item, 16);
                        cp_wait();
                        __syncwarp();
                        const int dcell = (int)lds_u(ldesc_u), dfirst = (int)lds_u(ldesc_u + 4), dnj = (int)lds_u(ldesc_u + 8);
                        if (lane == 0) {
                            cp_async(ldesc_u + 16, a.list_len + dcell, 8);
                            cp_async(ldesc_u + 24, a.list_off + dcell, 8);
                        }
                        if (lane < dnj) cp_async(stt + lane * 4, a.sorted_pairs + dfirst + lane, 4);
                        else sts_u(stt + lane * 4, 0xFFFFFFFFu);
                        cp_wait();
                    } else {
                        // the item after this one (work distribution by an atomic counter), and the item's running state
                        if (lane == 0) sts_u(ldesc_u + 32, (uint32_t)(gridDim.x + atomicAdd(ua.item_counter, 1)));
                        sts_f(stt + 2 * QG * 4 + lane * 4, Limits<float>::inf());
                        sts_u(stt + 3 * QG * 4 + lane * 4, 0u);
                        sts_u(stt + 4 * QG * 4 + lane * 4, 0u);
                    }
                    named_bar(2, W_NLOAD);
                    cell = (int)lds_u(ldesc_u);
                    nj = (int)lds_u(ldesc_u + 8);
                    { const uint2 v = lds_v2u(ldesc_u + 16); len = (int64_t)(((uint64_t)v.y << 32) | v.x); }
                    { const uint2 v = lds_v2u(ldesc_u + 24); off = (int64_t)(((uint64_t)v.y << 32) | v.x); }
                    npass = (int)((len + W_VP - 1) / W_VP);
                    pf.tick(1);
                    if (fastq) {
                        // Query rows, coalesced: row q = m * 32 contiguous bytes = 2 m chunks of 16, stored at
                        // chunk ^ (q & 7) (keeps the transposing reads below at 4-way bank conflicts)
                        for (int idx = lt; idx < QG * 2 * m; idx += W_NLOAD) {
                            const int q = idx / (2 * m), ch = idx - q * (2 * m);
                            const int pq = (int)lds_u(stt + q * 4);
                            const uint32_t dst = rawq_u + q * (m * 32) + ((ch ^ (q & 7)) * 16);
                            if (pq >= 0) cp_async(dst, a.Q + (size_t)(pq / a.w) * a.D + ch * 4, 16);
                            else sts_v4f(dst, 0.f, 0.f, 0.f, 0.f);
                        }
                        if (lt < 2 * m) cp_async(cbuf_u + lt * 16, a.C + (size_t)cell * a.D + lt * 4, 16);
                    } else {
                        for (int x = lt; x < m * QG; x += W_NLOAD) {
                            const int s = x >> 5, q = x & 31;
                            const int pq = (int)lds_u(stt + q * 4);
                            const float* qv = a.Q + (size_t)(pq >= 0 ? pq / a.w : 0) * a.D + s * a.dsub;
#pragma unroll
                            for (int d = 0; d < 8; ++d) {
                                const uint32_t dst = res_i + ((s * 8 + d) * T_RS + q) * 4;
                                if (pq >= 0 && d < a.dsub) cp_async(dst, qv + d, 4);
                                else sts_f(dst, 0.f);
                            }
                        }
                        for (int x = lt; x < m * 8; x += W_NLOAD) {
                            const int s = x >> 3, d = x & 7;
                            if (d < a.dsub) cp_async(cbuf_u + x * 4, a.C + (size_t)cell * a.D + s * a.dsub + d, 4);
                            else sts_f(cbuf_u + x * 4, 0.f);
                        }
                    }
                    if (lw == 0) {  // dc lands in the `base` slot; |r|^2 is added below
                        const int pq = (int)lds_u(stt + lane * 4);
                        if (pq >= 0) cp_async(stt + QG * 4 + lane * 4, a.dc + pq, 4);
                        else sts_f(stt + QG * 4 + lane * 4, 0.f);
                    }
                }
                // code bytes of the pass as they lie in the list (16-byte pieces; the list starts 16-byte aligned, a pass
                // is a multiple of 16 bytes, and the arena keeps slack behind every list)
                const int64_t vb = (int64_t)pass * W_VP;
                const int nvn = (int)min((int64_t)W_VP, len - vb);
                {
                    const uint8_t* src = a.codes + (size_t)(off + vb) * mc;
                    const int n16 = (nvn * mc + 15) >> 4;
                    for (int idx = lt; idx < n16; idx += W_NLOAD) cp_async(raw_u + idx * 16, src + (size_t)idx * 16, 16);
                }
                cp_wait();
                named_bar(2, W_NLOAD);
                pf.tick(2);
                // byte planes: plane p = code byte p of the pass's vectors (4 x 4 byte transposes)
                for (int task = lt; 4 * task < nvn; task += W_NLOAD) {
                    const int v0 = 4 * task;
                    uint32_t cw[4][NP];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if constexpr (NP == 4) {
                            const uint4 r = lds_v4(raw_u + (v0 + i) * 16);
                            cw[i][0] = r.x; cw[i][1] = r.y; cw[i][2] = r.z; cw[i][3] = r.w;
                        } else {
#pragma unroll
                            for (int c = 0; c < NP; ++c) cw[i][c] = lds_u(raw_u + ((v0 + i) * NP + c) * 4);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < NP; ++c) {
                        const uint32_t w0 = cw[0][c], w1 = cw[1][c], w2 = cw[2][c], w3 = cw[3][c];
                        const uint32_t t01l = __byte_perm(w0, w1, 0x5140), t23l = __byte_perm(w2, w3, 0x5140);
                        const uint32_t t01h = __byte_perm(w0, w1, 0x7362), t23h = __byte_perm(w2, w3, 0x7362);
                        const uint32_t pb = planes_u + spar * PLANES_BYTES + (4 * c) * W_VP + v0;
                        sts_u(pb, __byte_perm(t01l, t23l, 0x5410));
                        sts_u(pb + W_VP, __byte_perm(t01l, t23l, 0x7632));
                        sts_u(pb + 2 * W_VP, __byte_perm(t01h, t23h, 0x5410));
                        sts_u(pb + 3 * W_VP, __byte_perm(t01h, t23h, 0x7632));
                    }
                }
                pf.tick(3);
                if (new_item) {
                    // residuals r = q - c (reference _closest_cluster_residuals, src/coarsequantizers.jl:40-45) and
                    // their squared norms per table; a loader warp handles one table per round, lane = query
                    for (int x = lt; x < m * QG; x += W_NLOAD) {
                        const int s = x >> 5;
                        float qd[8];
                        if (fastq) {
                            const uint32_t rq = rawq_u + lane * (m * 32);
                            const uint4 u0 = lds_v4(rq + (((2 * s) ^ (lane & 7)) * 16)), u1 = lds_v4(rq + (((2 * s + 1) ^ (lane & 7)) * 16));
                            qd[0] = __uint_as_float(u0.x); qd[1] = __uint_as_float(u0.y); qd[2] = __uint_as_float(u0.z); qd[3] = __uint_as_float(u0.w);
                            qd[4] = __uint_as_float(u1.x); qd[5] = __uint_as_float(u1.y); qd[6] = __uint_as_float(u1.z); qd[7] = __uint_as_float(u1.w);
                        } else {
#pragma unroll
                            for (int d = 0; d < 8; ++d) qd[d] = lds_f(res_i + ((s * 8 + d) * T_RS + lane) * 4);
                        }
                        float part = 0.f, big = 0.f;
#pragma unroll
                        for (int d = 0; d < 8; ++d) {
                            const float r = sub_rn(qd[d], lds_f(cbuf_u + (s * 8 + d) * 4));
                            sts_f(res_i + ((s * 8 + d) * T_RS + lane) * 4, r);
                            part = fma_rn(r, r, part);
                            big = fmaxf(big, fabsf(r));
                        }
                        sts_f(rnorm_u + (s * QG + lane) * 4, part);
                        sts_f(rmax_u + (s * QG + lane) * 4, big);
                    }
                    named_bar(2, W_NLOAD);
                    if (lw == 0) {
                        float rn = 0.f, big = 0.f;
                        for (int s = 0; s < m; ++s) {
                            rn = add_rn(rn, lds_f(rnorm_u + (s * QG + lane) * 4));  // fixed order
                            big = fmaxf(big, lds_f(rmax_u + (s * QG + lane) * 4));
                        }
                        const uint32_t bu = stt + QG * 4 + lane * 4;
                        const float base = add_rn(lds_f(bu), rn);  // dc + |r|^2 over the PQ dims
                        // Power-of-two scales of this query (header comment): r 2^sr below 2^14, a = 2^(ew + sr - en)
                        // representable in fp16 (subnormals included), everything the scanners add scaled by 2^(ew + sr).
                        // fmaxf drops NaNs, so a NaN residual shows up in the norm (base), an infinite one in `big`.
                        const int eb = (int)((__float_as_uint(big) >> 23) & 0xffu) - 126;   // big < 2^eb
                        int sr = min(14 - eb, 15 + ua.en - ua.ew);
                        sr = max(-100, min(100, sr));
                        const int sa = ua.ew + sr - ua.en, sall = ua.ew + sr;
                        const int sallc = max(-126, min(126, sall));
                        const float scaled = base * exp2i(sallc);
                        const bool bad = !(fabsf(scaled) < 3.0e38f) || !(big < 3.0e38f) || sall > 120 || sall < -120;
                        sts_f(bu, scaled);
                        sts_f(stt + 5 * QG * 4 + lane * 4, exp2i(sr));
                        const uint32_t ah = sa >= -24 ? (uint32_t)__half_as_ushort(__float2half_rn(exp2i(max(sa, -24)))) : 0u;
                        sts_u(stt + 6 * QG * 4 + lane * 4, ah | (ah << 16));
                        sts_f(stt + 7 * QG * 4 + lane * 4, exp2i(-sallc));
                        if (bad) sts_u(stt + 4 * QG * 4 + lane * 4, 1u);   // the exact redo kernel serves this pair
                    }
                }
                if (lt == 0) {
                    sts_u(sg + 4, (uint32_t)nvn);
                    sts_u(sg + 8, (uint32_t)pass);
                    sts_u(sg + 12, (uint32_t)ipar);
                    sts_u(sg + 16, (uint32_t)nj);
                    sts_u(sg + 20, pass + 1 == npass ? 1u : 0u);
                    sts_u(sg + 24, (uint32_t)cell);
                    sts_u(sg, 1u);
                }
                named_bar(2, W_NLOAD);
                if (lt == 0) mbar_arrive(bar_staged + 8 * spar);
                pf.tick(4);
                if (pass + 1 < npass) {
                    ++pass;
                    new_item = false;
                } else {
                    item = (int)lds_u(ldesc_u + 32);
                    pass = 0;
                    ipar ^= 1;
                    new_item = true;
                }
            }
            pf.store(stamps && lane == 0 && drow >= 0 ? stamps + 8 * drow : nullptr);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (wid == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(U_TMEM_COLS)
                     : "memory");
    }
}

// Codebook -> fp16 B operand blocks of the warp-specialised kernel, once at create.  Row n of a block is the codeword
// whose code VALUE is n (as prep_tcu_kernel).  Per table two blocks of 256 rows x 16 k (K-major, no swizzle:
// half(n, k) at byte (n >> 3) * 256 + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2):
//   block 0 = [wh | wh], block 1 = [wl | n0 n1 0 ..]  with  -2 w 2^ew = wh + wl,  |w|^2 2^en = n0 + n1  (fp16 pieces).
__global__ void prep_tch_kernel(const float* __restrict__ cb, const uint8_t* __restrict__ cb_codes, int identity,
                                int m, int ksub, int dsub, int dup, int ew, int en, __half* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * dup * ksub) return;
    const int t = idx / ksub, n = idx - t * ksub;
    const int s = t / dup, d0 = (t - s * dup) * 8;
    const int row = identity ? n : (int)cb_codes[(size_t)s * ksub + n];
    __half* o = out + (size_t)t * (W_BSUB / 2);
    const float sw = exp2i(ew), sn = exp2i(en);
    const int rb = (row >> 3) * 128 + (row & 7) * 8;   // in halves
    float nrm = 0.f;
    for (int kk = 0; kk < 8; ++kk) {
        const float w = d0 + kk < dsub ? cb[((size_t)s * ksub + n) * dsub + d0 + kk] : 0.f;
        nrm = fma_rn(w, w, nrm);
        const float v = -2.f * w * sw;
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        o[rb + kk] = hi;
        o[rb + 64 + kk] = hi;
        o[W_BBLK / 2 + rb + kk] = lo;
    }
    const float ns = nrm * sn;
    const __half n0 = __float2half_rn(ns);
    o[W_BBLK / 2 + rb + 64] = n0;
    o[W_BBLK / 2 + rb + 65] = __float2half_rn(ns - __half2float(n0));
}

}  // namespace ivf
