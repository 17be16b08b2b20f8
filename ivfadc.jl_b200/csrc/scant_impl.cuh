// K2 + K3, "query-per-lane" scan with the lookup tables built by the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, accumulators in tensor memory, codebook operand streamed by TMA bulk
// copies) -- the default batched search kernel on sm_100a (fp32, k <= 16, dsub <= 8, m % 4 == 0).
//
// Same work decomposition and scan loop as scanq_impl.cuh (work item = one inverted list x up to
// 32 queries probing it, lane = query, PQ code byte warp-uniform, every lookup one conflict-free
// 128-byte shared-memory wavefront, 64 partial distances per lane in registers); what changes is
// K2.  The mma.sync builder of scanq spends ~190 instructions per 16x32 table tile (fragment
// loads, lane rotations, 8-byte stores) and serialises with the scan; here
//
//   * the table of ONE subspace (256 codes x 32 queries, 32 KB) is an accumulator tile
//     D[128 lanes][256 columns] in tensor memory: lane = (copy, query), column = codeword index,
//         D = A_hi.B_hi + A_lo.B_hi + A_hi.B_lo  (3xTF32 split, fp32 accumulate)  + 1.|w|^2
//     with A = residuals r = q - c of that subspace (rows 32..127 repeat rows 0..31, so all four
//     lane quarters -- hence all 16 warps -- can read the tile), B = -2 * codebook and the split
//     squared norms.  Four M128 N256 K8 MMAs per subspace, issued by one thread TWO subspaces ahead
//     of the scan into a double-buffered accumulator (512 columns): nobody waits for them;
//   * B (24 KB per subspace, canonical K-major no-swizzle core-matrix layout, prepared once at
//     create) arrives by cp.async.bulk into a three-deep shared-memory ring, completion on mbarriers;
//   * the epilogue is one tcgen05.ld 32x32b.x16 per warp (lane = query: 16 consecutive codewords
//     of that query) and 16 conflict-free 128-byte STS rows into the table layout the scan wants;
//     the table is double-buffered, so the epilogue of subspace s + 1 and the scan of subspace s
//     share one barrier interval and their shared-memory traffic overlaps.
//
// Numerics: entry = |w|^2 - 2 r.w (+ per-query constant dc + |r|^2 added at scan start), the GEMM
// form of the reference's direct form (src/index.jl:234), accurate to ~1e-6 relative of the
// returned distance (bound 1e-5); the reference-exact chain remains available through
// IVFADC_FLAG_LUT_EXACT (scanq_kernel<false>).
#pragma once

#include "scanq_impl.cuh"

namespace ivf {

constexpr int TA_BLK = 4096;                 // A block: 128 rows x 8 k (tf32), canonical layout
constexpr int TB_BLK = 8192;                 // B block: 256 rows x 8 k
constexpr int TB_NBLK = 3;                   // hi, lo, norms
constexpr int TB_SUB = TB_NBLK * TB_BLK;     // bytes of B per subspace
constexpr int TB_RING = 3;                   // ring slots
constexpr int TA_BYTES = 5 * TA_BLK;         // (hi, lo) x 2 buffers, ones
constexpr uint32_t T_TMEM_COLS = 512;        // two accumulator tiles of 256 columns
constexpr int T_RS = 33;                     // row stride of the transposed residuals
constexpr uint32_t T_SPIN = 1u << 22;        // bound on every mbarrier wait (no hangs: error flag instead)

struct ScanTArgs {
    ScanQArgs q;
    const float* tcB;     // [m][3][2048] tf32 words: -2w hi | lo, split norms of one subspace
    const int4* items;    // [nitems] (cell, first pair slot, number of pairs, 0)
    int* err;             // device error flag (mbarrier timeout)
    float* dbg_lut;       // optional dump of work item 0: tables [m][256][32], then int pair[32], int cell
};

// Shared-memory plan, byte offsets from the start of dynamic shared memory.  The 64 KB table (two
// 32 KB buffers interleaved: 256-byte rows = [code][buffer][query]) is placed
// at an ABSOLUTE shared address that is a multiple of 64 KB, so that one byte-permute builds a
// complete 32-bit lookup address  base | code << 8 | lane offset  (no add per lookup); the small
// arrays fill the gap in front of it.
struct ScanTSmem {
    uint32_t abuf, resid, planes, smin, misc, bars, front_end;  // in front of the table
    uint32_t lut, bbuf, cand_d, cand_p, total;                  // from the table on (lut set at run time)
};

__host__ __device__ inline ScanTSmem scant_smem_layout(int m, uint32_t dyn_base) {
    ScanTSmem s;
    uint32_t o = 0;
    s.abuf = o;   o += TA_BYTES;
    s.resid = o;  o += (uint32_t)m * 8 * T_RS * 4;
    o = (o + 15) & ~15u;
    s.planes = o; o += (uint32_t)m * QVP;                      // byte planes: plane s = code byte s of 1024 vectors
    s.smin = o;   o += QWARPS * QG * 4;
    s.misc = o;   o += 6 * QG * 4;
    o = (o + 15) & ~15u;
    s.bars = o;   o += 96;
    s.front_end = o;
    s.lut = ((dyn_base + o + 0xFFFFu) & ~0xFFFFu) - dyn_base;
    o = s.lut + 65536;
    s.bbuf = o;   o += TB_RING * TB_SUB;
    s.cand_d = o; o += QG * QCAP * 4;
    s.cand_p = o; o += QG * QCAP * 4;
    s.total = o;
    return s;
}
constexpr uint32_t T_DYN_BASE_GUESS = 0x400;  // dynamic shared memory starts after the 1 KB the system reserves

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
// Bounded wait: a broken pipeline raises the error flag instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int code) {
    for (uint32_t i = 0; i < T_SPIN; ++i)
        if (mbar_try_wait(bar, parity)) return;
    atomicExch(err, code);
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, SWIZZLE_NONE: 8-row x 16-byte core matrices stored as
// 128 contiguous bytes; LBO = byte distance between the two K halves of one MMA (k 0..3 | 4..7),
// SBO = byte distance between 8-row groups.  (cute::UMMA::SmemDescriptor, version 1 = Blackwell.)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t addr) {
    constexpr uint64_t LBO = 128, SBO = 256;
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((LBO >> 4) << 16) | ((SBO >> 4) << 32) | (1ull << 46);
}
// Instruction descriptor of tcgen05.mma kind::tf32: D fp32, A/B tf32 K-major, M = 128, N = 256.
constexpr uint32_t T_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(T_IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- explicit shared-memory accesses -------------------------------------------------------------
// Every shared-memory access of this kernel goes through a 32-bit shared address held in a register.
// With generic C++ pointers nvcc re-materialises the shared window base (S2UR SR_CgaCtaId, UMOV,
// ULEA, IADD3) in front of almost every access once registers are tight -- 4 extra issue slots per
// load, measured as 30% of the instructions of the first version of this kernel (profiles/).
__device__ __forceinline__ float lds_f(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f(uint32_t a, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_u(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_v4f(uint32_t a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ int atoms_add(uint32_t a, int v) {
    int old;
    asm volatile("atom.shared.add.s32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v) : "memory");
    return old;
}

// ---- K3: one subspace over the 64 vectors of this warp ------------------------------------------
// Codes are staged as BYTE PLANES (plane s = code byte s of the pass's 1024 vectors), so one uniform
// 16-byte load delivers the codes of 16 consecutive vectors for the current subspace: 4 code loads
// per warp and subspace (word-planes needed 16: every uniform LDS.128 costs shared-memory wavefronts,
// measured as ~25% of the pipe in the previous version).
// lob = table base (multiple of 64 KB) | buffer * 128 | lane * 4: ONE PRMT makes the whole address
//   byte 0 <- buffer / lane offset, byte 1 <- code byte, bytes 2..3 <- table base.
template <int BY, bool FIRST>
__device__ __forceinline__ float scant_vec(uint32_t lob, uint32_t x, float acc, float base) {
    const float v = lds_f(__byte_perm(x, lob, 0x7604 | (BY << 4)));
    return FIRST ? add_rn(base, v) : add_rn(acc, v);  // subspace order, the reference's chain (src/index.jl:242-246)
}
template <bool FIRST>
__device__ __forceinline__ void scant_word(uint32_t lob, uint32_t x, float* acc, float base) {
    acc[0] = scant_vec<0, FIRST>(lob, x, acc[0], base);
    acc[1] = scant_vec<1, FIRST>(lob, x, acc[1], base);
    acc[2] = scant_vec<2, FIRST>(lob, x, acc[2], base);
    acc[3] = scant_vec<3, FIRST>(lob, x, acc[3], base);
}
// Warp w owns vectors 256 * j4 + 16 * w + i (j4 < 4, i < 16) of the pass: acc[16 * j4 + i].
template <bool FIRST>
__device__ __forceinline__ void scant_sub(uint32_t lob, uint32_t plane_w, int nch, float base, float (&acc)[QNV]) {
#pragma unroll
    for (int j4 = 0; j4 < QNV / 16; ++j4) {
        if (j4 < nch) {  // warp-uniform; slots beyond the list inside a chunk are masked after the last subspace
            const uint4 x = lds_v4(plane_w + j4 * 256);
            scant_word<FIRST>(lob, x.x, &acc[16 * j4 + 0], base);
            scant_word<FIRST>(lob, x.y, &acc[16 * j4 + 4], base);
            scant_word<FIRST>(lob, x.z, &acc[16 * j4 + 8], base);
            scant_word<FIRST>(lob, x.w, &acc[16 * j4 + 12], base);
        }
    }
}

constexpr int TTHREADS = QTHREADS + 32;  // 16 scan warps + 1 producer warp (TMA + MMA issue)
__device__ __forceinline__ void scan_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }  // the 16 scan warps
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <bool IDENT>
__global__ void __launch_bounds__(TTHREADS, 1)
scant_kernel(const ScanTArgs ta) {
    const ScanQArgs& a = ta.q;
    extern __shared__ __align__(1024) unsigned char smem_t[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int wid = tid >> 5;
    const int m = a.m;
    const int k = a.k;

    const int item = blockIdx.x;
    if (item >= a.group_off[a.kc]) return;  // before any allocation
    const int4 it = ta.items[item];
    const int cell = it.x, first = it.y, nj = it.z;
    // bring-up timeline: clock() of thread 0 of work item 300 at phase boundaries (tests/debug only)
    long long* tstamp = (ta.dbg_lut != nullptr && item == 300 && tid == 0)
                            ? reinterpret_cast<long long*>(ta.dbg_lut + (size_t)m * 256 * 32 + 64) : nullptr;
    int nstamp = 0;
    auto stamp = [&]() { if (tstamp) tstamp[nstamp++] = clock64(); };
    stamp();

    // one opaque copy of the dynamic shared-memory base; all addresses below are sb + offset
    uint32_t sb;
    asm volatile("mov.u32 %0, %1;" : "=r"(sb) : "r"(smem_u32(smem_t)));
    const ScanTSmem L = scant_smem_layout(m, sb);
    {
        uint32_t dyn_size;
        asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_size));
        if (L.total > dyn_size) {  // cannot happen with the 1 KB system reservation the host assumes
            if (tid == 0) atomicExch(ta.err, 3);
            return;
        }
    }
    const uint32_t lut_u = sb + L.lut;  // multiple of 64 KB
    const uint32_t abuf_u = sb + L.abuf, bbuf_u = sb + L.bbuf, resid_u = sb + L.resid, planes_u = sb + L.planes;
    const uint32_t cand_d_u = sb + L.cand_d, cand_p_u = sb + L.cand_p, smin_u = sb + L.smin;
    const uint32_t dc_u = sb + L.misc, thr_u = dc_u + QG * 4, run_u = thr_u + QG * 4, cnt_u = run_u + QG * 4,
                   pair_u = cnt_u + QG * 4, flag_u = pair_u + QG * 4;
    const uint32_t bar_full = sb + L.bars;         // [3] codebook operand landed in ring slot i      (TMA tx)
    const uint32_t bar_mma = bar_full + 24;        // [2] accumulator tile i complete                 (tcgen05.commit)
    const uint32_t bar_aready = bar_full + 40;     // [2] A operand buffer i written                  (1 arrival)
    const uint32_t bar_tfree = bar_full + 56;      // [2] accumulator tile i drained by the epilogue  (16 arrivals)
    const uint32_t tmem_slot = bar_full + 72;

    const int64_t len = a.list_len[cell];
    const int npass = (int)((len + QVP - 1) / QVP);
    const int T = npass * m;  // table builds (one per subspace per pass); build x is subspace x % m

    // =============================== producer warp ===============================================
    if (wid == QWARPS) {
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < TB_RING; ++i) mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_mma, 1);
            mbar_init(bar_mma + 8, 1);
            mbar_init(bar_aready, 1);
            mbar_init(bar_aready + 8, 1);
            mbar_init(bar_tfree, QWARPS);
            mbar_init(bar_tfree + 8, QWARPS);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
            for (int x = 0; x < TB_RING; ++x)
                if (x < T) {
                    mbar_expect_tx(bar_full + 8 * x, TB_SUB);
                    tma_bulk_g2s(bbuf_u + x * TB_SUB, ta.tcB + (size_t)(x % m) * (TB_SUB / 4), TB_SUB, bar_full + 8 * x);
                }
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "r"(T_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncthreads();  // [A]
        tc_fence_after();
        const uint32_t tmem_base = lds_u(tmem_slot);
        if (lane == 0) {
            const uint64_t A1 = tc_smem_desc(abuf_u + 4 * TA_BLK);
            for (int x = 0; x < T; ++x) {
                const uint32_t buf = x & 1, slot = (uint32_t)x % TB_RING;
                if (x >= 2) mbar_wait(bar_tfree + 8 * buf, (((uint32_t)x >> 1) - 1) & 1, ta.err, 4);  // tile drained
                mbar_wait(bar_aready + 8 * buf, ((uint32_t)x >> 1) & 1, ta.err, 5);                     // A(x) written
                mbar_wait(bar_full + 8 * slot, ((uint32_t)x / TB_RING) & 1, ta.err, 1);                 // B(x) landed
                tc_fence_after();
                const uint32_t ab = abuf_u + buf * (2 * TA_BLK), bb = bbuf_u + slot * TB_SUB;
                const uint32_t d = tmem_base + buf * 256;
                const uint64_t Ah = tc_smem_desc(ab), Al = tc_smem_desc(ab + TA_BLK);
                const uint64_t Bh = tc_smem_desc(bb), Bl = tc_smem_desc(bb + TB_BLK), Bn = tc_smem_desc(bb + 2 * TB_BLK);
                tc_mma(d, Ah, Bh, 0);
                tc_mma(d, Al, Bh, 1);
                tc_mma(d, Ah, Bl, 1);
                tc_mma(d, A1, Bn, 1);
                tc_commit(bar_mma + 8 * buf);
                if (x >= 1 && x + 2 < T) {  // ring slot of build x - 1 is free once that build has completed
                    mbar_wait(bar_mma + 8 * ((x - 1) & 1), ((uint32_t)(x - 1) >> 1) & 1, ta.err, 6);
                    const uint32_t s2 = (uint32_t)(x + 2) % TB_RING, bar = bar_full + 8 * s2;
                    mbar_expect_tx(bar, TB_SUB);
                    tma_bulk_g2s(bbuf_u + s2 * TB_SUB, ta.tcB + (size_t)((x + 2) % m) * (TB_SUB / 4), TB_SUB, bar);
                }
            }
        }
        __syncwarp();
        __syncthreads();  // [Z] every scan warp is done with tensor memory
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(T_TMEM_COLS)
                     : "memory");
        return;
    }

    // =============================== scan warps ==================================================
    // Stage the codes of one pass as byte planes: a thread takes 4 consecutive vectors, transposes
    // their code words 4 x 4 bytes at a time with byte permutes, and stores one word per subspace.
    auto stage_planes = [&](int64_t vbase, int nv) {
        const int nplanes = m >> 2;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(a.codes + (size_t)a.list_off[cell] * m) +
                              (size_t)vbase * nplanes;
        for (int quad = tid; 4 * quad < nv; quad += QTHREADS) {
            const int v0 = 4 * quad;
            for (int c = 0; c < nplanes; ++c) {
                uint32_t w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) w[i] = v0 + i < nv ? __ldg(src + (size_t)(v0 + i) * nplanes + c) : 0u;
                const uint32_t t01l = __byte_perm(w[0], w[1], 0x5140), t23l = __byte_perm(w[2], w[3], 0x5140);
                const uint32_t t01h = __byte_perm(w[0], w[1], 0x7362), t23h = __byte_perm(w[2], w[3], 0x7362);
                const uint32_t pb = planes_u + (4 * c) * QVP + v0;
                sts_u(pb, __byte_perm(t01l, t23l, 0x5410));
                sts_u(pb + QVP, __byte_perm(t01l, t23l, 0x7632));
                sts_u(pb + 2 * QVP, __byte_perm(t01h, t23h, 0x5410));
                sts_u(pb + 3 * QVP, __byte_perm(t01h, t23h, 0x7632));
            }
        }
    };

    if (tid < QG) {
        const int p = tid < nj ? a.sorted_pairs[first + tid] : -1;
        sts_u(pair_u + tid * 4, (uint32_t)p);
        sts_f(dc_u + tid * 4, p >= 0 ? a.dc[p] : 0.f);
        sts_f(run_u + tid * 4, Limits<float>::inf());
        sts_u(cnt_u + tid * 4, 0u);
        sts_u(flag_u + tid * 4, 0u);
    }
    // ones block: every row selects the split norm in k slots 0, 1
    for (int i = tid; i < 256; i += QTHREADS) {
        const int r = i >> 1, half = i & 1;
        const float one = half == 0 ? 1.f : 0.f;
        sts_v4f(abuf_u + 4 * TA_BLK + (r >> 3) * 256 + half * 128 + (r & 7) * 16, one, one, 0.f, 0.f);
    }
    fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
    stage_planes(0, (int)min((int64_t)QVP, len));
    tc_fence_before();
    __syncthreads();  // [A] barriers initialised, tensor memory allocated, pair slots written
    tc_fence_after();
    const uint32_t tmem_base = lds_u(tmem_slot);
    stamp();

    // residuals r_q = query - centroid (reference _closest_cluster_residuals,
    // src/coarsequantizers.jl:40-45): warp s owns subspace s, lane q its query; transposed
    // [s * 8 + d][q], zero-padded to 8 dims per subspace
    {
        float part = 0.f;
        if (wid < m) {
            const int p = (int)lds_u(pair_u + lane * 4);
            const float* qv = a.Q + (size_t)(p >= 0 ? p / a.w : 0) * a.D + wid * a.dsub;
            const float* cv = a.C + (size_t)cell * a.D + wid * a.dsub;
            float qr[8], cr[8];
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                qr[d] = (p >= 0 && d < a.dsub) ? __ldg(qv + d) : 0.f;
                cr[d] = (p >= 0 && d < a.dsub) ? __ldg(cv + d) : 0.f;
            }
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                const float r = sub_rn(qr[d], cr[d]);
                sts_f(resid_u + ((wid * 8 + d) * T_RS + lane) * 4, r);
                part = fma_rn(r, r, part);
            }
        }
        sts_f(smin_u + (wid * QG + lane) * 4, part);
    }
    scan_bar();
    stamp();
    float base;
    {
        float rn = 0.f;
        for (int s = 0; s < m; ++s) rn = add_rn(rn, lds_f(smin_u + (s * QG + lane) * 4));  // fixed order
        base = add_rn(lds_f(dc_u + lane * 4), rn);  // dc + |r|^2 over the PQ dims
    }

    // ONE warp writes the whole A operand of build x (the four row copies are the same 32 rows):
    // lane = query, 8 dims -> tf32 hi / lo, 16 stores; then tells the producer.  The caller rotates the warp.
    auto write_A = [&](int x, int s) {
        float hi[8], lo[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            const float r = lds_f(resid_u + ((s * 8 + d) * T_RS + lane) * 4);
            hi[d] = __uint_as_float(to_tf32(r));
            lo[d] = __uint_as_float(to_tf32(r - hi[d]));
        }
        const uint32_t ph = abuf_u + (x & 1) * (2 * TA_BLK) + (lane >> 3) * 256 + (lane & 7) * 16;
#pragma unroll
        for (int c = 0; c < 4; ++c) {  // row = 32 c + lane
            sts_v4f(ph + c * 1024, hi[0], hi[1], hi[2], hi[3]);
            sts_v4f(ph + c * 1024 + 128, hi[4], hi[5], hi[6], hi[7]);
            sts_v4f(ph + c * 1024 + TA_BLK, lo[0], lo[1], lo[2], lo[3]);
            sts_v4f(ph + c * 1024 + TA_BLK + 128, lo[4], lo[5], lo[6], lo[7]);
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_aready + 8 * (x & 1));
    };
    // Epilogue of build x (subspace s): tensor memory -> table buffer x & 1 in the scan's layout.
    // Warp w: lane quarter w & 3 (all quarters hold the same 32 queries), codewords 16w .. 16w + 15.
    auto epilogue = [&](int x, int s) {
        mbar_wait(bar_mma + 8 * (x & 1), ((uint32_t)x >> 1) & 1, ta.err, 2);
        tc_fence_after();
        uint32_t v[16];
        tc_ld16(tmem_base + ((uint32_t)((wid & 3) * 32) << 16) + (uint32_t)((x & 1) * 256 + 16 * wid), v);
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tfree + 8 * (x & 1));  // this warp's part of the tile is drained
        const uint32_t edst = lut_u + (x & 1) * 128 + lane * 4;
        if constexpr (IDENT) {
            // rows of codewords >= ksub are never looked up: store unconditionally
#pragma unroll
            for (int i = 0; i < 16; ++i) sts_u(edst + (16 * wid + i) * 256, v[i]);
        } else {
            const uint8_t* cvp = a.cb_codes + (size_t)s * a.ksub;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int code = 16 * wid + i;
                if (code < a.ksub) sts_u(edst + (int)cvp[code] * 256, v[i]);  // code VALUE -> entry (Q6)
            }
        }
    };

    if (wid == 2) write_A(0, 0);
    if (wid == 3 && T > 1) write_A(1, 1 % m);
    stamp();

    const uint32_t lob0 = lut_u | (uint32_t)(lane * 4), lob1 = lob0 + 128;
    const uint32_t plane_w0 = planes_u + 16 * wid;
    const bool dbg = ta.dbg_lut != nullptr && item == 0;

    float acc[QNV];
    int nv = 0, nch = 0;
    int64_t vbase = 0;
    int s = m - 1;  // subspace of build t; the loop starts one step early (t = -1: only the first epilogue)
    for (int t = -1; t < T; ++t) {
        const int s1 = s + 1 == m ? 0 : s + 1;
        if (s1 == 0 && t + 1 < T) {  // the next build opens a pass
            vbase = (int64_t)((t + 1) / m) * QVP;
        }
        // odd warps drain the next table before scanning, even warps after: the tensor-memory load
        // latency of one half hides behind the lookups of the other half
        const bool epi = t + 1 < T;
        if (epi && ((wid & 1) || t < 0)) epilogue(t + 1, s1);
        if (t >= 0) {
            if (s == 0) {
                nv = (int)min((int64_t)QVP, len - vbase);
                nch = min(4, max(0, (nv - 16 * wid + 255) >> 8));  // chunks j4 with 256 * j4 + 16 * wid < nv
            }
            if (dbg && t < m) {
                if (tid < QG) reinterpret_cast<int*>(ta.dbg_lut + (size_t)m * 256 * 32)[tid] = (int)lds_u(pair_u + tid * 4);
                if (tid == 0) reinterpret_cast<int*>(ta.dbg_lut + (size_t)m * 256 * 32)[QG] = cell;
                for (int idx = tid; idx < 256 * 32; idx += QTHREADS) {
                    const int q = idx & 31, code = idx >> 5;
                    ta.dbg_lut[((size_t)s * 256 + code) * 32 + q] = lds_f(lut_u + code * 256 + (t & 1) * 128 + q * 4);
                }
            }
            // ---- K3: scan subspace s (table buffer t & 1 == s & 1, byte plane s) ----
            const uint32_t plane_w = plane_w0 + s * QVP;
            if (s == 0) scant_sub<true>(lob0, plane_w, nch, base, acc);
            else scant_sub<false>((s & 1) ? lob1 : lob0, plane_w, nch, base, acc);
            if (epi && !(wid & 1)) epilogue(t + 1, s1);
        }
        if (t + 3 < T && wid == ((t + 5) & 15)) write_A(t + 3, (s + 3) % m);  // build t + 1 complete: its A buffer is free
        scan_bar();  // table buffer t & 1 consumed, table t + 1 complete
        s = s1;
        stamp();
        if (t < 0 || s != 0) continue;

        // ---- per-(query, list) top-k of this pass ----
        // mask the slots beyond the list (only the chunk that straddles the end has any)
        {
            const int lim = nv - 16 * wid;  // slot 16 * j4 + i holds a vector iff 256 * j4 + i < lim
#pragma unroll
            for (int j4 = 0; j4 < QNV / 16; ++j4) {
                if (256 * j4 + 16 > lim) {  // warp-uniform: chunk not completely inside the list
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (256 * j4 + i >= lim) acc[16 * j4 + i] = Limits<float>::inf();
                }
            }
        }
        if (t + 1 < T) {  // planes of the next pass (this pass's scan is complete)
            stage_planes(vbase, (int)min((int64_t)QVP, len - vbase));
        }
        const int64_t pbase = (int64_t)(t / m) * QVP + 16 * wid;  // position of slot 0 of this warp
        float mn = Limits<float>::inf();
#pragma unroll
        for (int j = 0; j < QNV; ++j) mn = fminf(mn, acc[j]);
        sts_f(smin_u + (wid * QG + lane) * 4, mn);
        scan_bar();
        {
            int rank = 0;
#pragma unroll
            for (int w2 = 0; w2 < QWARPS; ++w2) {
                const float o = lds_f(smin_u + (w2 * QG + lane) * 4);
                rank += (o < mn || (o == mn && w2 < wid)) ? 1 : 0;
            }
            // the k-th smallest of 16 distinct candidates bounds the k-th smallest of all
            if (rank == min(k, QWARPS) - 1) sts_f(thr_u + lane * 4, fminf(mn, lds_f(run_u + lane * 4)));
        }
        scan_bar();
        {
            const float thr = lds_f(thr_u + lane * 4);
            const bool any_real = thr < Limits<float>::inf();  // fewer than k vectors so far: keep every real slot
#pragma unroll
            for (int j = 0; j < QNV; ++j) {
                if (acc[j] <= thr && (any_real || acc[j] < Limits<float>::inf())) {
                    const int slot = atoms_add(cnt_u + lane * 4, 1);
                    if (slot < QCAP) {
                        sts_f(cand_d_u + (lane * QCAP + slot) * 4, acc[j]);
                        sts_u(cand_p_u + (lane * QCAP + slot) * 4, (uint32_t)(pbase + 256 * (j >> 4) + (j & 15)));
                    }
                }
            }
        }
        scan_bar();
        // exact selection by (distance, position): warp w serves queries w and w + 16
        for (int q = wid; q < QG; q += QWARPS) {
            int n = (int)lds_u(cnt_u + q * 4);
            const bool ovf = n > QCAP;
            n = min(n, QCAP);
            const uint32_t cd = cand_d_u + q * QCAP * 4, cp = cand_p_u + q * QCAP * 4;
            float d0 = Limits<float>::inf(), d1 = Limits<float>::inf();
            uint32_t p0 = kNoPos, p1 = kNoPos;
            if (lane < n) { d0 = lds_f(cd + lane * 4); p0 = lds_u(cp + lane * 4); }
            if (lane + 32 < n) { d1 = lds_f(cd + (lane + 32) * 4); p1 = lds_u(cp + (lane + 32) * 4); }
            int r0 = 0, r1 = 0;
            for (int e = 0; e < n; ++e) {
                const float de = lds_f(cd + e * 4);
                const uint32_t pe = lds_u(cp + e * 4);
                r0 += cand_before(de, pe, d0, p0) ? 1 : 0;
                r1 += cand_before(de, pe, d1, p1) ? 1 : 0;
            }
            __syncwarp();
            if (lane < n && r0 < k) { sts_f(cd + r0 * 4, d0); sts_u(cp + r0 * 4, p0); }
            if (lane + 32 < n && r1 < k) { sts_f(cd + r1 * 4, d1); sts_u(cp + r1 * 4, p1); }
            __syncwarp();
            if (lane == 0) {
                const int cnt = min(n, k);
                sts_u(cnt_u + q * 4, (uint32_t)cnt);
                sts_f(run_u + q * 4, cnt >= k ? lds_f(cd + (k - 1) * 4) : Limits<float>::inf());
                if (ovf) sts_u(flag_u + q * 4, 1u);
            }
        }
        scan_bar();
        stamp();
    }

    // ---- publish ----
    for (int idx = tid; idx < nj * k; idx += QTHREADS) {
        const int q = idx / k, e = idx - q * k;
        const int pair = (int)lds_u(pair_u + q * 4);
        if (!lds_u(flag_u + q * 4) && e < (int)lds_u(cnt_u + q * 4)) {
            a.pair_d[(size_t)pair * k + e] = lds_f(cand_d_u + (q * QCAP + e) * 4);
            a.pair_pos[(size_t)pair * k + e] = lds_u(cand_p_u + (q * QCAP + e) * 4);
        }
    }
    if (tid < nj) {
        const int pair = (int)lds_u(pair_u + tid * 4);
        if (lds_u(flag_u + tid * 4)) {
            a.pair_cnt[pair] = 0;
            a.redo_pairs[atomicAdd(a.redo_cnt, 1)] = pair;
        } else {
            a.pair_cnt[pair] = (int)lds_u(cnt_u + tid * 4);
        }
    }
    tc_fence_before();
    __syncthreads();  // [Z]
}

// Codebook -> B operand blocks of the tensor-core table builder, once at create.
// Per subspace: block 0 = tf32 hi of -2w, block 1 = lo, block 2 = split |w|^2 in k slots 0, 1.
// Block layout (fp32 words): word(n, k) = (n >> 3) * 64 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3).
__global__ void prep_tc_kernel(const float* __restrict__ cb, int m, int ksub, int dsub, float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * TB_NBLK * 2048) return;
    const int kk = idx & 7, n = (idx >> 3) & 255, b = (idx >> 11) % TB_NBLK, s = idx / (2048 * TB_NBLK);
    float val = 0.f;
    if (b < 2) {
        const float v = (n < ksub && kk < dsub) ? -2.f * cb[((size_t)s * ksub + n) * dsub + kk] : 0.f;
        const float hi = __uint_as_float(to_tf32(v));
        val = b ? __uint_as_float(to_tf32(v - hi)) : hi;
    } else if (kk < 2 && n < ksub) {
        float nrm = 0.f;
        for (int d = 0; d < dsub; ++d) {
            const float w = cb[((size_t)s * ksub + n) * dsub + d];
            nrm = fma_rn(w, w, nrm);
        }
        const float hi = __uint_as_float(to_tf32(nrm));
        val = kk ? __uint_as_float(to_tf32(nrm - hi)) : hi;
    }
    out[(size_t)(s * TB_NBLK + b) * 2048 + (n >> 3) * 64 + (kk >> 2) * 32 + (n & 7) * 4 + (kk & 3)] = val;
}

}  // namespace ivf
