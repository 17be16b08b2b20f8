// K2 + K3, "tensor-memory lookup" scan -- the default batched search kernel on sm_100a
// (fp32, k <= 16, dsub <= 8, m in {4, 8, 12, 16}).
//
// Same work decomposition as scanq / scant (work item = one inverted list x up to 32 queries that
// probe it, lane = query, the PQ code byte of a (vector, subspace) is warp-uniform), but the ADC
// lookup tables of src/index.jl:232-236 never leave the tensor cores' accumulator memory:
//
//   * the table of ONE subspace for the 32 queries of the item is one accumulator tile
//     D[128 lanes][256 columns] in tensor memory: lane = (copy, query) -- the four lane quarters
//     hold the same 32 queries, so that every warp (a warp can only read the quarter warp_id % 4)
//     sees the whole table -- and column = code value.  Four tcgen05.mma kind::tf32 (M128 N256 K8)
//     per subspace:  D = A_hi.B_hi + A_lo.B_hi + A_hi.B_lo + 1.|w|^2  (3xTF32 split, fp32
//     accumulate), A = residuals q - c, B = -2 * codebook and its split squared norms;
//   * tensor memory (512 columns) holds two such tiles: while the warps scan subspace s out of one,
//     the tensor core builds subspace s + 1 into the other.  B (24 KB per subspace, canonical
//     K-major no-swizzle core matrices, rows permuted to code VALUES once at create) streams
//     through a three-deep cp.async.bulk ring that never drains: the subspace sequence is periodic,
//     so the ring runs ahead across work items;
//   * K3: a lookup is  tcgen05.ld.32x32b.x1  at column = code byte.  The code bytes of 16 vectors
//     arrive by one uniform 16-byte shared load; ptxas keeps them in UNIFORM registers (the warp
//     index is made provably uniform), so one lookup costs UPRMT (address) + LDTM + FADD and no
//     shared-memory bandwidth at all.  The v6 kernel (scant) copied every table through shared
//     memory (4.1 k store wavefronts per item) and paid one 128-byte LDS wavefront per lookup
//     (15.2 k per item); here the LSU only carries the code bytes;
//   * the kernel is PERSISTENT (one CTA per SM, work items handed out by an atomic counter): at an
//     item boundary the codes / queries of the next item are loaded while the distances of the
//     current one are reduced (its descriptor was fetched by warp 0 during the scan), and the first
//     two table builds of the next item run under the candidate dump of the current one.
//
// Per-(query, list) candidates: the k-th smallest of the 16 per-warp minima bounds the k-th distance
// of the list; every vector with d <= bound goes to per-(query, warp) slots (no atomics), the serving
// warp compacts them into the pair's candidate row in global memory (<= U_CAP per pair, typically
// 2-3 k).  The exact (distance, probe rank, position) selection of src/index.jl:247-257 happens once
// per query in merge_cands_kernel (scan.cu) over the union of its w candidate rows -- a superset of
// every per-list top-k, ties at the bound included.  Overflow (heavy ties / short lists) goes to the
// redo queue served by the general kernel, which writes an exact sorted top-k into the same row.
//
// Numerics: entry = |w|^2 - 2 r.w (+ per-query constant dc + |r|^2 added first), summed in subspace
// order 1..m as the reference's chain (src/index.jl:242-246); ~1e-6 relative of the returned
// distance (bound 1e-5).
#pragma once

#include "tc_common.cuh"

namespace ivf {

constexpr int U_ABLK = 4096;                 // A block: 128 rows x 8 k (tf32)
constexpr int U_ASUB = 2 * U_ABLK;           // hi, lo of one subspace
constexpr int U_BBLK = 8192;                 // B block: 256 rows x 8 k
constexpr int U_BSUB = 3 * U_BBLK;           // hi, lo, norms of one subspace
constexpr int U_NB = 3;                      // B ring depth
constexpr uint32_t U_TMEM_COLS = 512;
constexpr int U_CW = 16;                     // candidate slots per (query, warp)
constexpr bool U_CTA_SYNC = true;             // per-table CTA barrier (true) or mbarrier release collected by the issuer (false)
#ifndef U_DEPTH
#define U_DEPTH 16                           // lookups in flight per tcgen05.wait::ld (16 or 32; 32 measured 2% slower)
#endif
// Vectors of a pass are dealt in chunks of 16: warp w owns chunks w, w + U_SW, w + 2 U_SW, w + 3 U_SW.
constexpr int U_SW = QWARPS;                  // full-share warps (U_SW = 15: the issuing warp keeps a half share and a pass
                                             // holds 992 vectors -- measured slower: lists of 993..1024 vectors take two passes)
constexpr int U_VP = U_SW == QWARPS ? 16 * 4 * QWARPS : 16 * (4 * U_SW + 2);
constexpr int U_CAP = 64;                    // candidate row of a (query, list) pair in global memory (more -> redo queue)

struct ScanUArgs {
    ScanQArgs q;
    const float* tcU;     // [m][3][2048] tf32 words: -2w hi / lo, split norms; rows = code values
    const void* tcH;      // scanw: [m][2][4096] fp16: (hi | hi), (lo | norm pieces) of -2w 2^ew, |w|^2 2^en
    int ew, en;           // scanw: the two power-of-two scales
    const int4* items;    // [nitems] (cell, first pair slot, number of pairs, 0)
    int* item_counter;    // work distribution (zeroed per launch)
    int pstride;          // candidate capacity of a pair's row
    int rstride;          // row stride of pair_d / pair_pos (words)
    int* err;             // device error flag (mbarrier timeout)
    float* dbg;           // optional: tables of the first item [m][256][32], then int pair[32], cell
};

struct ScanUSmem {
    uint32_t aone, aring, bring, resid, planes, raw, cand, smin, rnorm, cntw, thr, misc, desc, cbuf, bars, total;
};
constexpr int U_MISC = 5 * QG * 4;  // per item parity: pair, dc, run (bound), cntf (candidates so far), flag

__host__ __device__ inline ScanUSmem scanu_smem_layout(int m) {
    ScanUSmem s;
    uint32_t o = 0;
    s.aone = o;    o += U_ABLK;
    s.aring = o;   o += 2 * U_ASUB;
    s.bring = o;   o += U_NB * U_BSUB;
    s.cand = o;    o += QG * QWARPS * U_CW * 8;
    s.planes = o;  o += (uint32_t)m * T_PLANE;
    s.raw = o;     o += (uint32_t)m * U_VP;      // code words of the next pass as they lie in the list
    s.resid = o;   o += (uint32_t)m * 8 * T_RS * 4;
    o = (o + 15) & ~15u;
    s.smin = o;    o += QWARPS * QG * 4;
    s.rnorm = o;   o += QWARPS * QG * 4;
    s.cntw = o;    o += QWARPS * QG * 4;
    s.thr = o;     o += QG * 4;
    s.misc = o;    o += 2 * U_MISC;
    s.desc = o;    o += 64 + QG * 4;       // next item: cell, first, nj, -, len, off | pair[32]
    s.cbuf = o;    o += QWARPS * 8 * 4;     // centroid slices of the next item
    s.bars = o;    o += 96;
    s.total = o;
    return s;
}

// Warp-converged issue: every lane executes these with warp-uniform operands and ONE elected lane issues.
// (Under `if (lane == 0)` ptxas wraps each tcgen05.mma in an elect / R2UR.BROADCAST loop that moves the
// five operands into uniform registers one by one.)
__device__ __forceinline__ void tc_mma_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(T_IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
        : "memory");
}
__device__ __forceinline__ float tc_ld1(uint32_t taddr) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
    return __uint_as_float(v);
}
__device__ __forceinline__ void sts_v2u(uint32_t a, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint2 lds_v2u(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
// order-preserving map of a float (no NaNs; -0 canonicalised by the caller) to unsigned
__device__ __forceinline__ uint32_t f_flip(float d) {
    const uint32_t b = __float_as_uint(d);
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float f_unflip(uint32_t u) {
    return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}

// ---- K3: one subspace over the vectors of this warp ----------------------------------------------
// tb = tensor-memory address of the table (lane quarter of this warp, column 0 of the buffer; the
// low byte is zero, so ONE uniform byte-permute forms the lookup address), pa = shared address of
// the 16 code bytes of the chunk.  Everything here is warp-uniform except acc / base.
// acc[0..1] += t[0..1] as ONE packed add (add.rn.f32x2: each half is an ordinary IEEE fp32 add of its own chain)
__device__ __forceinline__ void add_pair(float& a0, float& a1, float t0, float t1) {
#if defined(IVF_X_SCALAR_ADD) && IVF_X_SCALAR_ADD
    a0 = __fadd_rn(a0, t0);
    a1 = __fadd_rn(a1, t1);
    return;
#endif
    unsigned long long a, t;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(t) : "f"(t0), "f"(t1));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(t));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(a));
}
template <int N>
__device__ __forceinline__ void scanu_issue(uint32_t tb, const uint4& x, float* t) {
    t[0] = tc_ld1(__byte_perm(x.x, tb, 0x7650));  t[1] = tc_ld1(__byte_perm(x.x, tb, 0x7651));
    t[2] = tc_ld1(__byte_perm(x.x, tb, 0x7652));  t[3] = tc_ld1(__byte_perm(x.x, tb, 0x7653));
    t[4] = tc_ld1(__byte_perm(x.y, tb, 0x7650));  t[5] = tc_ld1(__byte_perm(x.y, tb, 0x7651));
    t[6] = tc_ld1(__byte_perm(x.y, tb, 0x7652));  t[7] = tc_ld1(__byte_perm(x.y, tb, 0x7653));
    t[8] = tc_ld1(__byte_perm(x.z, tb, 0x7650));  t[9] = tc_ld1(__byte_perm(x.z, tb, 0x7651));
    t[10] = tc_ld1(__byte_perm(x.z, tb, 0x7652)); t[11] = tc_ld1(__byte_perm(x.z, tb, 0x7653));
    t[12] = tc_ld1(__byte_perm(x.w, tb, 0x7650)); t[13] = tc_ld1(__byte_perm(x.w, tb, 0x7651));
    t[14] = tc_ld1(__byte_perm(x.w, tb, 0x7652)); t[15] = tc_ld1(__byte_perm(x.w, tb, 0x7653));
}
// FULL: all four chunks of the warp are inside the list -> two chunks (32 lookups) in flight per wait;
// the wait costs a fixed ~200 cycles per warp, so the depth of a batch sets the pace of the scan.
template <bool FIRST, bool FULL>
__device__ __forceinline__ void scanu_sub(uint32_t tb, uint32_t plane_w, uint32_t cstep, int nch, float base, float (&acc)[QNV]) {
    if constexpr (FULL && U_DEPTH == 32) {
#pragma unroll
        for (int j2 = 0; j2 < QNV / 32; ++j2) {
            const uint4 x0 = lds_v4(plane_w + (2 * j2) * cstep), x1 = lds_v4(plane_w + (2 * j2 + 1) * cstep);
            float t[32];
            scanu_issue<0>(tb, x0, t);
            scanu_issue<0>(tb, x1, t + 16);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; i += 2) add_pair(acc[32 * j2 + i], acc[32 * j2 + i + 1], t[i], t[i + 1]);
        }
    } else {
#pragma unroll
        for (int j4 = 0; j4 < QNV / 16; ++j4) {
            if (FULL || j4 < nch) {  // warp-uniform; slots beyond the list are masked after the last subspace
                const uint4 x = lds_v4(plane_w + j4 * cstep);
                float t[16];
                scanu_issue<0>(tb, x, t);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; i += 2) add_pair(acc[16 * j4 + i], acc[16 * j4 + i + 1], t[i], t[i + 1]);
            }
        }
    }
}

// NP = (code bytes per vector) / 4: 32-bit code words per database vector.
// DUP = 2 serves dsub = 16 (config D: m = 8): a 16-dim subspace is scanned as TWO tables of 8 dims that share
// the code byte -- |r - w|^2 splits over the dims, so entry(code) = entry_lo(code) + entry_hi(code); the kernel
// runs m = 2 * (code bytes) tables of 8 dims, table s reads byte plane s / 2 (a.dsub is passed as 8).
template <int NP, bool DBG, int DUP = 1>
__global__ void __launch_bounds__(QTHREADS, 1)
scanu_kernel(const ScanUArgs ua) {
    const ScanQArgs& a = ua.q;
    extern __shared__ __align__(1024) unsigned char smem_u[];
    constexpr int mc = 4 * NP;      // code bytes per vector
    constexpr int m = mc * DUP;     // tables per work item
    constexpr int SB = m / 4, SC = m / 2;  // subspace iterations at which warp 0 advances the descriptor prefetch
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    // provably warp-uniform for ptxas: everything derived from it lives in uniform registers
    const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int k = a.k;
    const int ps = ua.pstride;
    const bool fastq = a.dsub == 8 && (a.D & 3) == 0;  // 16-byte aligned query / centroid slices
    const int nitems = a.group_off[a.kc];
    if ((int)blockIdx.x >= nitems) return;  // before any allocation

    uint32_t sb;
    asm volatile("mov.u32 %0, %1;" : "=r"(sb) : "r"(smem_u32(smem_u)));
    const ScanUSmem L = scanu_smem_layout(m);
    const uint32_t aone_u = sb + L.aone, aring_u = sb + L.aring, bring_u = sb + L.bring, resid_u = sb + L.resid,
                   planes_u = sb + L.planes, cand_u = sb + L.cand, smin_u = sb + L.smin, rnorm_u = sb + L.rnorm,
                   cntw_u = sb + L.cntw, thr_u = sb + L.thr, misc_u = sb + L.misc, desc_u = sb + L.desc,
                   raw_u = sb + L.raw, cbuf_u = sb + L.cbuf;
    const uint32_t bar_full = sb + L.bars;          // 3 x 8 bytes
    const uint32_t bar_mma = bar_full + 24;         // 2 x 8 bytes
    const uint32_t bar_a = bar_full + 40;           // A operands of the first two builds of a segment
    const uint32_t bar_free = bar_full + 48;        // 2 x 8 bytes: all 16 warps are done with a table buffer
    const uint32_t tmem_slot = bar_full + 64, next_slot = bar_full + 72;

    // ---- one-time setup ----
    if (wid == T_ISSUER && lane == 0) {
        for (int i = 0; i < U_NB; ++i) mbar_init(bar_full + 8 * i, 1);
        mbar_init(bar_mma, 1);
        mbar_init(bar_mma + 8, 1);
        mbar_init(bar_a, 8);
        mbar_init(bar_free, QWARPS);
        mbar_init(bar_free + 8, QWARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < U_NB; ++i) {
            mbar_expect_tx(bar_full + 8 * i, U_BSUB);
            tma_bulk_g2s(bring_u + i * U_BSUB, ua.tcU + (size_t)(i % m) * (U_BSUB / 4), U_BSUB, bar_full + 8 * i);
        }
    }
    if (wid == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "r"(U_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ones block of the A operand: every row selects the split norm (k slots 0, 1)
    for (int i = tid; i < 256; i += QTHREADS) {
        const int r = i >> 1, half = i & 1;
        const float one = half == 0 ? 1.f : 0.f;
        sts_v4f(aone_u + (r >> 3) * 256 + half * 128 + (r & 7) * 16, one, one, 0.f, 0.f);
    }

    // ---- descriptor of the NEXT work item, fetched by warp 0 in three non-blocking steps ----
    // desc (shared): cell, first, nj, -, len (int64), list offset (int64) | pair[32]
    // cp.async (global -> shared without a register in between): nothing is held across the scan
    auto cp_async = [&](uint32_t dst, const void* src, int bytes) {
        if (bytes == 16) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        else if (bytes == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
        else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
    };
    auto cp_wait = [&]() { asm volatile("cp.async.wait_all;" ::: "memory"); };
    auto desc_a = [&](int nitem) {
        if (lane == 0) cp_async(desc_u, ua.items + nitem, 16);
    };
    auto desc_b = [&]() {
        cp_wait();
        __syncwarp();
        const int dcell = (int)lds_u(desc_u), dfirst = (int)lds_u(desc_u + 4), dnj = (int)lds_u(desc_u + 8);
        if (lane == 0) {
            cp_async(desc_u + 16, a.list_len + dcell, 8);
            cp_async(desc_u + 24, a.list_off + dcell, 8);
        }
        if (lane < dnj) cp_async(desc_u + 64 + lane * 4, a.sorted_pairs + dfirst + lane, 4);
        else sts_u(desc_u + 64 + lane * 4, 0xFFFFFFFFu);
    };
    auto desc_c = [&]() { cp_wait(); };

    // ---- segment = (work item, pass of 1024 vectors) state ----
    int item = blockIdx.x;        // current item
    int ipar = 0;                 // parity of the per-item shared arrays
    int nj = 0;                   // queries of the item
    int nv = 0;                   // vectors of the running pass
    int npass = 1, pass = 0;
    uint32_t tglob = 0;           // number of the first table build of the running segment (uniform, all threads)
    uint32_t nstage = 0;          // segments staged so far (phase of bar_a)
    float base = 0.f;

    // Asynchronous loads of a segment (cp.async, global -> shared, no registers held while the
    // distances of the running segment are reduced): the code words of the pass into `raw`, and for a new
    // item the query slices (transposed, into the residual array), the centroid slices and dc.  Every
    // location is later read by the thread (or warp) that copied it.
    auto seg_load = [&](bool new_item, int par, int pass_n) {
        // desc holds the item this segment belongs to: the next item if new_item, else the running one
        const int cell_n = (int)lds_u(desc_u);
        int64_t len_n, off_n;
        { const uint2 v = lds_v2u(desc_u + 16); len_n = (int64_t)(((uint64_t)v.y << 32) | v.x); }
        { const uint2 v = lds_v2u(desc_u + 24); off_n = (int64_t)(((uint64_t)v.y << 32) | v.x); }
        if (new_item) {
            const int my_pair = (int)lds_u(desc_u + 64 + lane * 4);
            if (fastq) {
                // Query rows, coalesced: warp w copies the PQ dims of rows w and w + 16 (m * 32 contiguous bytes,
                // 16 per lane) into rawq[row][chunk ^ (row & 7)] (overlaid on the free A ring; the XOR keeps the
                // transposing reads of seg_stage at 4-way bank conflicts).  One row = 4 cache lines per request
                // instead of the 32 lines a lane-per-query gather touches.
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int q = wid + QWARPS * h;
                    const int pq = (int)lds_u(desc_u + 64 + q * 4);
                    if (lane < 2 * m) {
                        const uint32_t dst = aring_u + q * (m * 32) + ((lane ^ (q & 7)) * 16);
                        if (pq >= 0) cp_async(dst, a.Q + (size_t)(pq / a.w) * a.D + lane * 4, 16);
                        else sts_v4f(dst, 0.f, 0.f, 0.f, 0.f);
                    }
                }
                if (wid < m && lane < 2) cp_async(cbuf_u + wid * 32 + lane * 16, a.C + (size_t)cell_n * a.D + wid * 8 + lane * 4, 16);
            } else if (wid < m) {
                const float* qv = a.Q + (size_t)(my_pair >= 0 ? my_pair / a.w : 0) * a.D + wid * a.dsub;
                const float* cv = a.C + (size_t)cell_n * a.D + wid * a.dsub;
#pragma unroll
                for (int d = 0; d < 8; ++d) {
                    const uint32_t dst = resid_u + ((wid * 8 + d) * T_RS + lane) * 4;
                    if (my_pair >= 0 && d < a.dsub) cp_async(dst, qv + d, 4);
                    else sts_f(dst, 0.f);
                }
                if (lane < 8) {
                    if (lane < a.dsub) cp_async(cbuf_u + (wid * 8 + lane) * 4, cv + lane, 4);
                    else sts_f(cbuf_u + (wid * 8 + lane) * 4, 0.f);
                }
            }
            if (wid == 0) {
                const uint32_t mu = misc_u + par * U_MISC + lane * 4;
                sts_u(mu, (uint32_t)my_pair);
                if (my_pair >= 0) cp_async(mu + QG * 4, a.dc + my_pair, 4);
                else sts_f(mu + QG * 4, 0.f);
                sts_f(mu + 2 * QG * 4, Limits<float>::inf());
                sts_u(mu + 3 * QG * 4, 0u);
                sts_u(mu + 4 * QG * 4, 0u);
            }
        }
        if (tid < U_VP / 4) {
            const int64_t vb = (int64_t)pass_n * U_VP;
            const int nvn = (int)min((int64_t)U_VP, len_n - vb);
            const uint32_t* src = reinterpret_cast<const uint32_t*>(a.codes + (size_t)off_n * mc) + (size_t)vb * NP;
            const int v0 = 4 * tid;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t dst = raw_u + (v0 + i) * (4 * NP);
                if (v0 + i < nvn) {
                    if constexpr (NP == 4) cp_async(dst, src + (size_t)(v0 + i) * NP, 16);
                    else if constexpr (NP == 2) cp_async(dst, src + (size_t)(v0 + i) * NP, 8);
                    else {
#pragma unroll
                        for (int c = 0; c < NP; ++c) cp_async(dst + 4 * c, src + (size_t)(v0 + i) * NP + c, 4);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < NP; ++c) sts_u(dst + 4 * c, 0u);
                }
            }
        }
    };
    // Staging of a segment (after the previous segment's scan and reduction): byte planes of the codes
    // (plane s = code byte s of the pass's vectors; 4 x 4 byte transposes), and for a new item the
    // residuals r = q - c (reference _closest_cluster_residuals, src/coarsequantizers.jl:40-45; warp s
    // owns subspace s, lane q its query) with their squared norms.
    auto seg_stage = [&](bool new_item) {  // the caller has waited for the copies (cp_wait) and passed a CTA barrier
        if (tid < U_VP / 4) {
            const int v0 = 4 * tid;
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
                const uint32_t pb = planes_u + (4 * c) * T_PLANE + v0;
                sts_u(pb, __byte_perm(t01l, t23l, 0x5410));
                sts_u(pb + T_PLANE, __byte_perm(t01l, t23l, 0x7632));
                sts_u(pb + 2 * T_PLANE, __byte_perm(t01h, t23h, 0x5410));
                sts_u(pb + 3 * T_PLANE, __byte_perm(t01h, t23h, 0x7632));
            }
        }
        if (new_item) {
            float part = 0.f;
            if (wid < m) {
                float qd[8];
                if (fastq) {  // rows copied by other warps: visible after the barrier the caller placed before this call
                    const uint32_t rq = aring_u + lane * (m * 32);
                    const uint4 u0 = lds_v4(rq + (((2 * wid) ^ (lane & 7)) * 16)), u1 = lds_v4(rq + (((2 * wid + 1) ^ (lane & 7)) * 16));
                    qd[0] = __uint_as_float(u0.x); qd[1] = __uint_as_float(u0.y); qd[2] = __uint_as_float(u0.z); qd[3] = __uint_as_float(u0.w);
                    qd[4] = __uint_as_float(u1.x); qd[5] = __uint_as_float(u1.y); qd[6] = __uint_as_float(u1.z); qd[7] = __uint_as_float(u1.w);
                } else {
#pragma unroll
                    for (int d = 0; d < 8; ++d) qd[d] = lds_f(resid_u + ((wid * 8 + d) * T_RS + lane) * 4);
                }
#pragma unroll
                for (int d = 0; d < 8; ++d) {
                    const float r = sub_rn(qd[d], lds_f(cbuf_u + (wid * 8 + d) * 4));
                    sts_f(resid_u + ((wid * 8 + d) * T_RS + lane) * 4, r);
                    part = fma_rn(r, r, part);
                }
            }
            sts_f(rnorm_u + (wid * QG + lane) * 4, part);
        }
    };
    // The A operand of subspace s (ring slot s & 1): rows (copy, q), hi / lo of r[s][q][0..7]; lane = query,
    // one row copy per call (four warps share a table's operand).  No proxy fence here (it costs a writer
    // several hundred cycles of its scan): the issuing warp fences once, after the barrier that makes these
    // writes visible to it and before the MMAs that read them.
    auto write_A = [&](int s, int c) {
        float hi[8], lo[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            const float r = lds_f(resid_u + ((s * 8 + d) * T_RS + lane) * 4);
            hi[d] = __uint_as_float(to_tf32(r));
            lo[d] = __uint_as_float(to_tf32(r - hi[d]));
        }
        const uint32_t ph = aring_u + (s & 1) * U_ASUB + c * 1024 + (lane >> 3) * 256 + (lane & 7) * 16;  // row = 32 c + lane
        sts_v4f(ph, hi[0], hi[1], hi[2], hi[3]);
        sts_v4f(ph + 128, hi[4], hi[5], hi[6], hi[7]);
        sts_v4f(ph + U_ABLK, lo[0], lo[1], lo[2], lo[3]);
        sts_v4f(ph + U_ABLK + 128, lo[4], lo[5], lo[6], lo[7]);
    };
    // One lane waits (mbarrier.try_wait: the hardware SUSPENDS the waiting thread, so the 15 waiting warps
    // do not take issue slots from the warp that is issuing the next build -- a test_wait spin was measured
    // to starve it), the warp reconverges behind it.  Bounded: a broken pipeline raises the error flag.
    auto warp_wait = [&](uint32_t bar, uint32_t parity, int code) {
        if (lane == 0) mbar_wait(bar, parity, ua.err, code);
        __syncwarp();
    };
    uint32_t tmem_base = 0;
    // bring-up: fine-grained clock stamps of the issuing lane (DBG instantiation only)
    long long* fstamp_p = nullptr;
    int fstamp_n = 0, fstamp_on = 0;
#define U_STAMP_F() do { if (DBG && fstamp_p && fstamp_on && fstamp_n < 200) fstamp_p[fstamp_n++] = clock64(); } while (0)
    // Build number t (global count) = subspace s of the running segment, issued by ONE lane: the codebook
    // operand of build t is in ring slot t % 3 (the ring is refilled three builds ahead).  The issuing warp
    // shares its scheduler with three scanning warps, so every instruction here costs ~4 cycles of the
    // critical path: the descriptors are offsets from three precomputed ones (the address field is
    // bits 0..13 in units of 16 bytes; no carry leaves it), slot and phase are carried, not divided.
    const uint64_t descA0 = tc_smem_desc(aring_u), descB0 = tc_smem_desc(bring_u), desc1 = tc_smem_desc(aone_u);
    uint32_t bslot = 0, bphase = 0;  // ring slot / phase of the next build (issuer lane only)
    auto issue_mma = [&](uint32_t t, int s) {  // called by ALL lanes of the issuing warp
        const uint32_t slot_u = __shfl_sync(0xffffffffu, bslot, 0), par_u = __shfl_sync(0xffffffffu, bphase, 0);
        const uint32_t t_u = __shfl_sync(0xffffffffu, t, 0), s_u = (uint32_t)__shfl_sync(0xffffffffu, s, 0);
        warp_wait(bar_full + 8 * slot_u, par_u, 1);
        tc_fence_after();
        fence_proxy_async();  // A operand: generic-proxy writes of four warps (ordered before this point by a barrier) -> async proxy
        const uint64_t Ah = descA0 + (uint64_t)((s_u & 1) * (U_ASUB >> 4)), Al = Ah + (U_ABLK >> 4);
        const uint64_t Bh = descB0 + (uint64_t)(slot_u * (U_BSUB >> 4)), Bl = Bh + (U_BBLK >> 4), Bn = Bl + (U_BBLK >> 4);
        const uint32_t d = tmem_base + (t_u & 1) * 256;
        tc_mma_elect(d, Ah, Bh, 0);
        tc_mma_elect(d, Al, Bh, 1);
        tc_mma_elect(d, Ah, Bl, 1);
        tc_mma_elect(d, desc1, Bn, 1);
        tc_commit_elect(bar_mma + 8 * (t_u & 1));
        if (++bslot == U_NB) { bslot = 0; bphase ^= 1; }
    };
    uint32_t rslot = 0, rsub = U_NB % m;  // ring slot to refill next / subspace whose operand goes there (issuer lane only)
    auto refill_B = [&](uint32_t) {  // build t (in order) has completed: its ring slot takes the operand of build t + 3
        const uint32_t bar = bar_full + 8 * rslot;
        U_STAMP_F();
        mbar_expect_tx(bar, U_BSUB);
        U_STAMP_F();
        tma_bulk_g2s(bring_u + rslot * U_BSUB, ua.tcU + (size_t)rsub * (U_BSUB / 4), U_BSUB, bar);
        U_STAMP_F();
        if (++rslot == U_NB) rslot = 0;
        if (++rsub == (uint32_t)m) rsub = 0;
    };
    // after a staging barrier: warps 2, 3 write the first two A operands, the issuer starts builds 0, 1
    // (no CTA barrier: the other warps go on to the candidate dump)
    auto seg_start_builds = [&]() {
        if (wid < 8) {  // warps 0..3: operand of subspace 0, warps 4..7: subspace 1; one row copy each
            write_A(wid >> 2, wid & 3);
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_a) : "memory");
        }
        if (wid == T_ISSUER) {
            warp_wait(bar_a, nstage & 1, 4);
            tc_fence_after();
            issue_mma(tglob, 0);
            issue_mma(tglob + 1, 1);
        }
        ++nstage;
    };

    // ---- first segment ----
    {
        if (wid == 0) {
            desc_a(item);
            desc_b();
            desc_c();
        }
        __syncthreads();
        nj = (int)lds_u(desc_u + 8);
        {
            const uint2 v = lds_v2u(desc_u + 16);
            const int64_t len = (int64_t)(((uint64_t)v.y << 32) | v.x);
            npass = (int)((len + U_VP - 1) / U_VP);
            nv = (int)min((int64_t)U_VP, len);
        }
        seg_load(true, 0, 0);
        cp_wait();
        __syncthreads();
        seg_stage(true);
        if (tid == 0) sts_u(next_slot, (uint32_t)(gridDim.x + atomicAdd(ua.item_counter, 1)));
        fence_proxy_async();  // A ones block
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        tmem_base = lds_u(tmem_slot);
        float rn = 0.f;
        for (int s = 0; s < m; ++s) rn = add_rn(rn, lds_f(rnorm_u + (s * QG + lane) * 4));  // fixed order
        base = add_rn(lds_f(misc_u + QG * 4 + lane * 4), rn);  // dc + |r|^2 over the PQ dims
        seg_start_builds();
    }
    const int quarter = wid & 3;
    const uint32_t tq = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int c0 = wid < U_SW ? wid : 4 * U_SW;          // first chunk of this warp
    const int cs = wid < U_SW ? U_SW : 1;                // chunk stride
    const int maxch = wid < U_SW ? 4 : 2;                // chunks owned
    const uint32_t cstep = 16u * cs;
    const uint32_t plane_w0 = planes_u + 16 * c0;
    // bring-up (DBG instantiation only): table dump of the first item of CTA 0, and clock64 of
    // thread 0 of CTA 0 at the phase boundaries of its 4th segment
    bool dbg_first = DBG && blockIdx.x == 0;
    long long* tstamp = (DBG && blockIdx.x == 0 && tid == 0)
                            ? reinterpret_cast<long long*>(ua.dbg + (size_t)m * 256 * 32 + 64) : nullptr;
    int nstamp = 0, nseg = 0;
    auto stamp = [&]() { if (DBG && tstamp && nseg == 3 && nstamp < 30) tstamp[nstamp++] = clock64(); };
    // issuer lane of CTA 0, same segment: after the table wait / own scan done / buffer free / MMAs issued, per subspace
    long long* istamp = (DBG && blockIdx.x == 0 && tid == (U_CTA_SYNC ? 32 : T_ISSUER * 32))
                            ? reinterpret_cast<long long*>(ua.dbg + (size_t)m * 256 * 32 + 64) + 32 : nullptr;
    int nistamp = 0;
    auto stamp_i = [&]() { if (DBG && istamp && nseg == 3 && nistamp < 80) istamp[nistamp++] = clock64(); };
    // every warp of CTA 0, subspaces 5 and 6 of the same segment: before the table wait / after it / scan done / past the barrier
    long long* wstamp = (DBG && blockIdx.x == 0 && lane == 0)
                            ? reinterpret_cast<long long*>(ua.dbg + (size_t)m * 256 * 32 + 64) + 112 + wid * 8 : nullptr;
    if (DBG && blockIdx.x == 0 && tid == T_ISSUER * 32) fstamp_p = reinterpret_cast<long long*>(ua.dbg + (size_t)m * 256 * 32 + 64) + 240;
    auto stamp_w = [&](int s, int i) { if (DBG && wstamp && nseg == 3 && (s == 5 || s == 6)) wstamp[(s - 5) * 4 + i] = clock64(); };

    float acc[QNV];
    for (;;) {
        stamp();
        fstamp_on = DBG && nseg == 3;
        const int nch = min(maxch, max(0, (nv - 16 * c0 + 16 * cs - 1) / (16 * cs)));  // chunks j4 with 16 (c0 + cs j4) < nv
        const bool same_item = pass + 1 < npass;
        // the item after this one (its index was published before the last barrier of the previous boundary)
        const int nitem = same_item ? item : (int)lds_u(next_slot);
        const bool has_next = same_item || nitem < nitems;
        const bool fetch_desc = !same_item && has_next && wid == 0;

        // ---- the m subspaces of this segment ----
#pragma unroll
        for (int j = 0; j < QNV; ++j) acc[j] = base;  // dc + |r|^2, then the table entries in subspace order
#pragma unroll 1
        for (int s = 0; s < m; ++s) {
            const uint32_t t = tglob + s;
            if (fetch_desc) {  // warp 0: one step of the next item's descriptor, each consumed 2+ k cycles after it was issued
                if (s == 0) desc_a(nitem);
                else if (s == SB) desc_b();
                else if (s == SC) desc_c();
            }
            if (U_CTA_SYNC) stamp_i();
            stamp_w(s, 0);
            warp_wait(bar_mma + 8 * (t & 1), (t >> 1) & 1, 2);
            tc_fence_after();
            if (wid == 0 && lane == 0) refill_B(t);  // as early as possible (the copy needs its two table periods), and not by the issuing warp
            stamp_i();
            stamp_w(s, 1);
            if (s + 2 < m && wid >= QWARPS - 4) write_A(s + 2, wid - (QWARPS - 4));  // build s has completed: its A slot is free
            // s and t are warp-uniform, but ptxas keeps the loop counter in a vector register unless told (one SHFL each)
            const uint32_t tb = tq + (__shfl_sync(0xffffffffu, t, 0) & 1) * 256;
            const uint32_t plane_w = plane_w0 + (__shfl_sync(0xffffffffu, s, 0) / DUP) * T_PLANE;
            // ONE code path for every subspace (the accumulators start at dc + |r|^2): the loop body stays
            // within the instruction cache
            if (nch == 4) scanu_sub<false, true>(tb, plane_w, cstep, nch, base, acc);
            else scanu_sub<false, false>(tb, plane_w, cstep, nch, base, acc);
            if (DBG && dbg_first) {  // bring-up: dump the table of subspace s of the first item (the four quarters agree)
                if (wid < 4) {
                    for (int c = wid; c < 256; c += 4) {
                        const float v = tc_ld1(tb + c);
                        tc_wait_ld();
                        ua.dbg[((size_t)s * 256 + c) * 32 + lane] = v;
                    }
                }
                if (s == 0 && tid < QG) reinterpret_cast<int*>(ua.dbg + (size_t)m * 256 * 32)[tid] = (int)lds_u(misc_u + ipar * U_MISC + tid * 4);
                if (s == 0 && tid == 0) reinterpret_cast<int*>(ua.dbg + (size_t)m * 256 * 32)[QG] = ua.items[item].x;
            }
            if constexpr (U_CTA_SYNC) {
                stamp_i();
                stamp_w(s, 2);
                tc_fence_before();
                __syncthreads();  // every warp is done with this table; A operand of build s + 2 written
                stamp_i();
                // the barrier instruction defers blocking: read the clock through a shared-memory round trip that cannot pass it
                if (DBG) { sts_u(thr_u + 0 * lane, lds_u(thr_u)); }
                stamp_w(s, 3);
                if (wid == T_ISSUER && s + 2 < m) {
                    tc_fence_after();
                    issue_mma(t + 2, s + 2);
                }
            } else {
                // this warp is done with the table (and, if it was the writer, with the A operand of build s + 2):
                // no CTA-wide barrier -- the warps run ahead into the next table, only the issuer collects the arrivals
                tc_fence_before();
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_free + 8 * (t & 1)) : "memory");
                stamp_i();
                if (wid == T_ISSUER) {
                    if (s + 2 < m) {
                        warp_wait(bar_free + 8 * (t & 1), (t >> 1) & 1, 6);
                        stamp_i();
                        tc_fence_after();
                        issue_mma(t + 2, s + 2);
                        stamp_i();
                    }
                    __syncwarp();
                }
            }
            stamp();
        }
        tglob += m;
        dbg_first = false;

        // ---- segment boundary ----
        // A warp gets here once builds m - 2, m - 1 have completed, i.e. every warp has finished subspace
        // m - 3: the residuals and both A slots are free, and (m >= 8) warp 0 has published the descriptor.
        if (m < 8) __syncthreads();
        // vectors of the next pass / item (desc: the running item while it has passes left, else the next one)
        int nj_n = nj, npass_n = npass, nv_n = 0;
        if (has_next) {
            const uint2 v = lds_v2u(desc_u + 16);
            const int64_t len_n = (int64_t)(((uint64_t)v.y << 32) | v.x);
            if (!same_item) {
                nj_n = (int)lds_u(desc_u + 8);
                npass_n = (int)((len_n + U_VP - 1) / U_VP);
            }
            nv_n = (int)min((int64_t)U_VP, len_n - (same_item ? (int64_t)(pass + 1) * U_VP : 0));
        }
        int nx = 0;
        const bool fetch_next = tid == 0 && has_next && !same_item;
        if (fetch_next) nx = atomicAdd(ua.item_counter, 1);  // the item after the next one; consumed before the staging barrier
        if (has_next) seg_load(!same_item, ipar ^ 1, same_item ? pass + 1 : 0);

        // mask the slots beyond the list (chunks not reached, and the chunk that straddles the end)
        {
            const int lim = nv - 16 * c0;  // slot 16 * j4 + i holds a vector iff 16 * cs * j4 + i < lim (and j4 < maxch)
#pragma unroll
            for (int j4 = 0; j4 < QNV / 16; ++j4) {
                if (j4 >= maxch || 16 * cs * j4 + 16 > lim) {  // warp-uniform
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (j4 >= maxch || 16 * cs * j4 + i >= lim) acc[16 * j4 + i] = Limits<float>::inf();
                }
            }
        }
        const uint32_t mu = misc_u + ipar * U_MISC;  // pair | dc | run | cntf | flag of the current item
        const uint32_t pair_u = mu, run_u = mu + 2 * QG * 4, cntf_u = mu + 3 * QG * 4, flag_u = mu + 4 * QG * 4;
        float mn = Limits<float>::inf();
#pragma unroll
        for (int j = 0; j < QNV; ++j) mn = fminf(mn, acc[j]);
        sts_f(smin_u + (wid * QG + lane) * 4, mn);
        __syncthreads();
        stamp();
        {
            int rank = 0;
#pragma unroll
            for (int w2 = 0; w2 < QWARPS; ++w2) {
                const float o = lds_f(smin_u + (w2 * QG + lane) * 4);
                rank += (o < mn || (o == mn && w2 < wid)) ? 1 : 0;
            }
            // the k-th smallest of 16 disjoint minima bounds the k-th smallest of all; the bound of the
            // previous passes of this list stays valid
            if (rank == min(k, QWARPS) - 1) {
                const float thr = fminf(mn, lds_f(run_u + lane * 4));
                sts_f(thr_u + lane * 4, thr);
                sts_f(run_u + lane * 4, thr);
            }
        }
        cp_wait();  // the asynchronous copies of the next segment (issued before the previous barrier) have landed
        __syncthreads();
        stamp();
        {
            const float thr = lds_f(thr_u + lane * 4);
            // keep d <= bound; while fewer than k vectors have been seen (bound = +inf) keep every real slot
            const float cut = thr < Limits<float>::inf() ? thr : 3.402823466e+38f;
            const uint32_t cbase = cand_u + ((lane * QWARPS + wid) * U_CW) * 8;  // [query][warp][slot]
            int c = 0;
            uint32_t jj = 0;  // running slot number, kept opaque: 64 immediates would each take a register
#pragma unroll
            for (int j = 0; j < QNV; ++j) {  // record = (register slot j, distance); the dump turns j into the list position
                if (acc[j] <= cut) {
                    if (c < U_CW) {  // two 32-bit stores: a 64-bit store wants an aligned register pair and makes ptxas spill
                        sts_u(cbase + c * 8, jj);
                        sts_u(cbase + c * 8 + 4, __float_as_uint(acc[j]));
                    }
                    ++c;
                }
                asm volatile("add.u32 %0, %0, 1;" : "+r"(jj));
            }
            sts_u(cntw_u + (wid * QG + lane) * 4, (uint32_t)c);
        }
        if (has_next) seg_stage(!same_item);
        if (fetch_next) sts_u(next_slot, (uint32_t)(gridDim.x + nx));
        tc_fence_before();
        __syncthreads();
        stamp();
        float base_n = base;
        if (has_next) {
            if (!same_item) {
                float rn = 0.f;
                for (int s = 0; s < m; ++s) rn = add_rn(rn, lds_f(rnorm_u + (s * QG + lane) * 4));  // fixed order
                base_n = add_rn(lds_f(misc_u + (ipar ^ 1) * U_MISC + QG * 4 + lane * 4), rn);
            }
            seg_start_builds();  // builds 0, 1 of the next segment run under the candidate dump below
        }
        // candidate dump: warp w serves queries w and w + 16; lane = (warp of origin, half of its slots)
        for (int q = wid; q < nj; q += QWARPS) {
            const int w2 = lane >> 1, half = lane & 1;
            const int cq = (int)lds_u(cntw_u + (w2 * QG + q) * 4);
            const int nl = min(max(min(cq, U_CW) - half * (U_CW / 2), 0), U_CW / 2);  // valid slots in this lane's half
            int incl = nl;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const int n = __shfl_sync(0xffffffffu, incl, 31);
            const int cf = (int)lds_u(cntf_u + q * 4);
            const bool bad = lds_u(flag_u + q * 4) != 0u || __any_sync(0xffffffffu, cq > U_CW) || cf + n > ps;
            const int pair = (int)lds_u(pair_u + q * 4);
            if (!bad) {
                const uint32_t src = cand_u + ((q * QWARPS + w2) * U_CW + half * (U_CW / 2)) * 8;
                const size_t dst = (size_t)pair * ua.rstride + cf + (incl - nl);
                const uint32_t pbase = (uint32_t)pass * U_VP + 16 * (w2 < U_SW ? w2 : 4 * U_SW);  // position of register slot 0 of warp w2
                const uint32_t pstep = 16 * (w2 < U_SW ? U_SW : 1);
                for (int i = 0; i < nl; ++i) {
                    const uint2 v = lds_v2u(src + i * 8);
                    a.pair_pos[dst + i] = pbase + pstep * (v.x >> 4) + (v.x & 15);
                    a.pair_d[dst + i] = __uint_as_float(v.y);
                }
            }
            if (lane == 0) {
                if (bad) sts_u(flag_u + q * 4, 1u);
                else sts_u(cntf_u + q * 4, (uint32_t)(cf + n));
                if (!same_item) {  // last pass of the list: publish the pair
                    if (bad) {
                        a.pair_cnt[pair] = 0;
                        a.redo_pairs[atomicAdd(a.redo_cnt, 1)] = pair;
                    } else {
                        a.pair_cnt[pair] = cf + n;
                    }
                }
            }
        }
        stamp();
        ++nseg;
        if (!has_next) break;
        nv = nv_n;
        if (same_item) {
            ++pass;
        } else {
            item = nitem; nj = nj_n; npass = npass_n;
            pass = 0;
            ipar ^= 1;
            base = base_n;
        }
    }

    // ---- drain: the codebook ring ran three builds ahead ----
    if (wid == T_ISSUER && lane == 0) {
        for (uint32_t t = tglob; t < tglob + U_NB; ++t) mbar_wait(bar_full + 8 * (t % U_NB), (t / U_NB) & 1, ua.err, 5);
    }
    tc_fence_before();
    __syncthreads();
    if (wid == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(U_TMEM_COLS)
                     : "memory");
    }
}

// Codebook -> B operand blocks of the tensor-memory table builder, once at create.  Row n of a block
// is the codeword whose code VALUE is n (reference: LittleDict keyed by cb.codes, src/index.jl:235),
// so that the scan addresses the accumulator tile with the stored code byte directly.
// Block layout (fp32 words): word(n, k) = (n >> 3) * 64 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3).
// dup = 2 (dsub = 16): table t = 2 s + half holds dims 8 half .. 8 half + 7 of subspace s.
__global__ void prep_tcu_kernel(const float* __restrict__ cb, const uint8_t* __restrict__ cb_codes, int identity,
                                int m, int ksub, int dsub, int dup, float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * dup * ksub) return;
    const int t = idx / ksub, n = idx - t * ksub;
    const int s = t / dup, d0 = (t - s * dup) * 8;
    const int row = identity ? n : (int)cb_codes[(size_t)s * ksub + n];
    float* o = out + (size_t)t * (U_BSUB / 4);
    float nrm = 0.f;
    for (int kk = 0; kk < 8; ++kk) {
        const float w = d0 + kk < dsub ? cb[((size_t)s * ksub + n) * dsub + d0 + kk] : 0.f;
        nrm = fma_rn(w, w, nrm);
        const float v = -2.f * w;
        const float hi = __uint_as_float(to_tf32(v));
        const float lo = __uint_as_float(to_tf32(v - hi));
        const int word = (row >> 3) * 64 + (kk >> 2) * 32 + (row & 7) * 4 + (kk & 3);
        o[word] = hi;
        o[2048 + word] = lo;
    }
    const float nh = __uint_as_float(to_tf32(nrm));
    const int w0 = (row >> 3) * 64 + (row & 7) * 4;
    o[4096 + w0] = nh;
    o[4096 + w0 + 1] = __uint_as_float(to_tf32(nrm - nh));
}

}  // namespace ivf
