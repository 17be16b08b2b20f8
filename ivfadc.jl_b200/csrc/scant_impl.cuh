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
//   * the table of a GROUP of two subspaces (2 x 256 codes x 32 queries, 64 KB) is ONE accumulator
//     tile D[128 lanes][256 columns] in tensor memory:  lane = (copy, subspace-in-group, query),
//     column = codeword index,
//         D = A_hi.B_hi + A_lo.B_hi + A_hi.B_lo  (3xTF32 split, fp32 accumulate)  + 1.|w|^2
//     with A = residuals r = q - c laid out block-diagonally (rows of the other subspace are zero,
//     rows 64..127 repeat rows 0..63 so that all four lane quarters -- hence all 16 warps -- can
//     read the tile), B = -2 * codebook and the split squared norms.  Seven M128 N256 K8 MMAs per
//     group, issued by ONE lane of the warp that carries half a scan share, while the others
//     already scan the previous group;
//   * B (40 KB per group, canonical K-major no-swizzle core-matrix layout, prepared once at
//     create) arrives by cp.async.bulk into a two-deep shared-memory ring, completion on mbarriers;
//   * the epilogue is one tcgen05.ld 32x32b.x32 per warp (lane = query: 32 consecutive codewords
//     of that query) and 32 conflict-free 128-byte STS rows into the table layout the scan wants.
// (Two other schedules were measured and rejected, see DESIGN.md: per-subspace builds two ahead
// with a double-buffered table, and a 17th producer warp -- 5 warps on one scheduler cap the kernel
// at 96 registers.)
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
constexpr int TB_NBLK = 5;                   // hi0, lo0, hi1, lo1, norms
constexpr int TB_GROUP = TB_NBLK * TB_BLK;   // bytes of B per group of two subspaces
constexpr int TA_BYTES = 5 * TA_BLK;         // hi0, lo0, hi1, lo1, ones
constexpr uint32_t T_TMEM_COLS = 256;
constexpr int T_RS = 33;                     // row stride of the transposed residuals
constexpr uint32_t T_SPIN = 1u << 22;        // bound on every mbarrier wait (no hangs: error flag instead)
// Vectors of a pass are dealt in chunks of 16: warp w owns chunks w, w + 16, w + 32, w + 48.  One lane
// of the ISSUER warp issues the TMA / MMA work of the next group.  (Giving that warp half a scan
// share -- 992-vector passes -- was measured: it turns every list of 993..1024 vectors into two passes.)
constexpr int T_ISSUER = QWARPS - 1;
constexpr int T_VP = QVP;                    // 1024 vectors per pass
constexpr int T_PLANE = 1024;                // bytes per byte plane

struct ScanTArgs {
    ScanQArgs q;
    const float* tcB;     // [m/2][5][2048] tf32 words: -2w hi/lo per subspace, split norms
    const int4* items;    // [nitems] (cell, first pair slot, number of pairs, 0)
    int* err;             // device error flag (mbarrier timeout)
    float* dbg_lut;       // optional dump of work item 0: tables [m][256][32], then int pair[32], int cell
};

// Shared-memory plan, byte offsets from the start of dynamic shared memory.  The 64 KB table
// (256-byte rows = [code][subspace-in-group][query]) is placed
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
    s.planes = o; o += (uint32_t)m * T_PLANE;                  // byte planes: plane s = code byte s of the pass's vectors
    s.smin = o;   o += QWARPS * QG * 4;
    s.misc = o;   o += 6 * QG * 4;
    o = (o + 15) & ~15u;
    s.bars = o;   o += 96;
    s.front_end = o;
    s.lut = ((dyn_base + o + 0xFFFFu) & ~0xFFFFu) - dyn_base;
    o = s.lut + 65536;
    s.bbuf = o;   o += 2 * TB_GROUP;
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
#pragma unroll 1
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

// ---- K3: one group (two subspaces) over the vectors of this warp -------------------------------
// Codes are staged as BYTE PLANES (plane s = code byte s of the pass's vectors), so one uniform
// 16-byte load delivers the codes of 16 consecutive vectors for one subspace.
// lob = table base (multiple of 64 KB) | subspace-in-group * 128 | lane * 4: ONE PRMT makes the
// whole lookup address: byte 0 <- lane offset, byte 1 <- code byte, bytes 2..3 <- table base.
template <int BY, bool FIRST>
__device__ __forceinline__ float scant_vec(uint32_t lo0, uint32_t lo1, uint32_t x0, uint32_t x1, float acc, float base) {
    const float v0 = lds_f(__byte_perm(x0, lo0, 0x7604 | (BY << 4)));
    const float v1 = lds_f(__byte_perm(x1, lo1, 0x7604 | (BY << 4)));
    acc = FIRST ? add_rn(base, v0) : add_rn(acc, v0);
    return add_rn(acc, v1);  // subspace order, the reference's chain (src/index.jl:242-246)
}
template <bool FIRST>
__device__ __forceinline__ void scant_word(uint32_t lo0, uint32_t lo1, uint32_t x0, uint32_t x1, float* acc, float base) {
    acc[0] = scant_vec<0, FIRST>(lo0, lo1, x0, x1, acc[0], base);
    acc[1] = scant_vec<1, FIRST>(lo0, lo1, x0, x1, acc[1], base);
    acc[2] = scant_vec<2, FIRST>(lo0, lo1, x0, x1, acc[2], base);
    acc[3] = scant_vec<3, FIRST>(lo0, lo1, x0, x1, acc[3], base);
}
template <bool FIRST>
__device__ __forceinline__ void scant_chunk(uint32_t lo0, uint32_t lo1, uint32_t pa, float* acc, float base) {
    const uint4 x0 = lds_v4(pa), x1 = lds_v4(pa + T_PLANE);
    scant_word<FIRST>(lo0, lo1, x0.x, x1.x, acc + 0, base);
    scant_word<FIRST>(lo0, lo1, x0.y, x1.y, acc + 4, base);
    scant_word<FIRST>(lo0, lo1, x0.z, x1.z, acc + 8, base);
    scant_word<FIRST>(lo0, lo1, x0.w, x1.w, acc + 12, base);
}
// Slot 16 * j4 + i of a warp = vector 16 * (c0 + cs * j4) + i of the pass.  FULL: all four chunks are
// inside the list -> straight-line code (the code loads of all chunks can be issued up front).
template <bool FIRST, bool FULL>
__device__ __forceinline__ void scant_group(uint32_t lo0, uint32_t lo1, uint32_t plane_w, uint32_t pstride, int nch,
                                            float base, float (&acc)[QNV]) {
#pragma unroll
    for (int j4 = 0; j4 < QNV / 16; ++j4) {
        if (FULL || j4 < nch)  // warp-uniform; slots beyond the list inside a chunk are masked after the last group
            scant_chunk<FIRST>(lo0, lo1, plane_w + j4 * pstride, &acc[16 * j4], base);
    }
}

template <bool IDENT>
__global__ void __launch_bounds__(QTHREADS, 1)
scant_kernel(const ScanTArgs ta) {
    const ScanQArgs& a = ta.q;
    extern __shared__ __align__(1024) unsigned char smem_t[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int wid = tid >> 5;
    const int m = a.m;
    const int k = a.k;
    const int ng = m >> 1;  // groups of two subspaces

    const int item = blockIdx.x;
    if (item >= a.group_off[a.kc]) return;  // before any allocation
    const int4 it = ta.items[item];
    const int cell = it.x, first = it.y, nj = it.z;
    // query slot of this lane and the start of its data: issued first, consumed after the setup
    const int my_pair = lane < nj ? a.sorted_pairs[first + lane] : -1;
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
    const uint32_t bar_full0 = sb + L.bars, bar_full1 = bar_full0 + 8, bar_mma = bar_full0 + 16,
                   tmem_slot = bar_full0 + 32;

    const int64_t len = a.list_len[cell];
    const int npass = (int)((len + T_VP - 1) / T_VP);
    const int T = npass * ng;  // table builds (one per group per pass)

    // ---- setup: barriers + first codebook operands (issuer lane), tensor memory, residuals, codes ----
    if (wid == T_ISSUER && lane == 0) {
        mbar_init(bar_full0, 1);
        mbar_init(bar_full1, 1);
        mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar_full0, TB_GROUP);
        tma_bulk_g2s(bbuf_u, ta.tcB, TB_GROUP, bar_full0);
        if (T > 1) {
            mbar_expect_tx(bar_full1, TB_GROUP);
            tma_bulk_g2s(bbuf_u + TB_GROUP, ta.tcB + (size_t)(1 % ng) * (TB_GROUP / 4), TB_GROUP, bar_full1);
        }
    }
    if (wid == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "r"(T_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // residuals r_q = query - centroid (reference _closest_cluster_residuals,
    // src/coarsequantizers.jl:40-45): warp s owns subspace s, lane q its query; transposed
    // [s * 8 + d][q], zero-padded to 8 dims per subspace
    float qr[8], cr[8];
    if (wid < m) {
        const float* qv = a.Q + (size_t)(my_pair >= 0 ? my_pair / a.w : 0) * a.D + wid * a.dsub;
        const float* cv = a.C + (size_t)cell * a.D + wid * a.dsub;
        if (a.dsub == 8 && (a.D & 3) == 0) {
            const float4 q0 = __ldg(reinterpret_cast<const float4*>(qv)), q1 = __ldg(reinterpret_cast<const float4*>(qv) + 1);
            const float4 e0 = __ldg(reinterpret_cast<const float4*>(cv)), e1 = __ldg(reinterpret_cast<const float4*>(cv) + 1);
            qr[0] = q0.x; qr[1] = q0.y; qr[2] = q0.z; qr[3] = q0.w; qr[4] = q1.x; qr[5] = q1.y; qr[6] = q1.z; qr[7] = q1.w;
            cr[0] = e0.x; cr[1] = e0.y; cr[2] = e0.z; cr[3] = e0.w; cr[4] = e1.x; cr[5] = e1.y; cr[6] = e1.z; cr[7] = e1.w;
            if (my_pair < 0) {
#pragma unroll
                for (int d = 0; d < 8; ++d) qr[d] = cr[d] = 0.f;
            }
        } else {
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                qr[d] = (my_pair >= 0 && d < a.dsub) ? __ldg(qv + d) : 0.f;
                cr[d] = (my_pair >= 0 && d < a.dsub) ? __ldg(cv + d) : 0.f;
            }
        }
    }
    if (wid == 0) {
        sts_u(pair_u + lane * 4, (uint32_t)my_pair);
        sts_f(dc_u + lane * 4, my_pair >= 0 ? a.dc[my_pair] : 0.f);
        sts_f(run_u + lane * 4, Limits<float>::inf());
        sts_u(cnt_u + lane * 4, 0u);
        sts_u(flag_u + lane * 4, 0u);
    }
    // A operand: zero rows stay zero for the whole item; ones block: row (copy, j, q) selects the
    // split norm of subspace j (k slots 2j, 2j + 1)
    for (int i = tid; i < 4 * TA_BLK / 16; i += QTHREADS) sts_v4f(abuf_u + i * 16, 0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < 256; i += QTHREADS) {
        const int r = i >> 1, half = i & 1, j = (r >> 5) & 1;
        const float a0 = (half == 0 && j == 0) ? 1.f : 0.f, a1 = (half == 0 && j == 1) ? 1.f : 0.f;
        sts_v4f(abuf_u + 4 * TA_BLK + (r >> 3) * 256 + half * 128 + (r & 7) * 16, a0, a0, a1, a1);
    }
    // Stage the codes of one pass as byte planes: a thread takes 4 consecutive vectors, transposes
    // their code words 4 x 4 bytes at a time with byte permutes, and stores one word per subspace.
    auto stage_planes = [&](int64_t vbase, int nv) {
        const int nplanes = m >> 2;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(a.codes + (size_t)a.list_off[cell] * m) +
                              (size_t)vbase * nplanes;
        auto put = [&](int c, int v0, const uint32_t (&w)[4]) {
            const uint32_t t01l = __byte_perm(w[0], w[1], 0x5140), t23l = __byte_perm(w[2], w[3], 0x5140);
            const uint32_t t01h = __byte_perm(w[0], w[1], 0x7362), t23h = __byte_perm(w[2], w[3], 0x7362);
            const uint32_t pb = planes_u + (4 * c) * T_PLANE + v0;
            sts_u(pb, __byte_perm(t01l, t23l, 0x5410));
            sts_u(pb + T_PLANE, __byte_perm(t01l, t23l, 0x7632));
            sts_u(pb + 2 * T_PLANE, __byte_perm(t01h, t23h, 0x5410));
            sts_u(pb + 3 * T_PLANE, __byte_perm(t01h, t23h, 0x7632));
        };
        for (int quad = tid; 4 * quad < nv; quad += QTHREADS) {
            const int v0 = 4 * quad;
            if (nplanes == 4) {  // m = 16: one 128-bit load per vector (lists start 16-byte aligned)
                uint4 r[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    r[i] = v0 + i < nv ? __ldg(reinterpret_cast<const uint4*>(src) + v0 + i) : make_uint4(0u, 0u, 0u, 0u);
                { const uint32_t w[4] = {r[0].x, r[1].x, r[2].x, r[3].x}; put(0, v0, w); }
                { const uint32_t w[4] = {r[0].y, r[1].y, r[2].y, r[3].y}; put(1, v0, w); }
                { const uint32_t w[4] = {r[0].z, r[1].z, r[2].z, r[3].z}; put(2, v0, w); }
                { const uint32_t w[4] = {r[0].w, r[1].w, r[2].w, r[3].w}; put(3, v0, w); }
            } else {
                for (int c = 0; c < nplanes; ++c) {
                    uint32_t w[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) w[i] = v0 + i < nv ? __ldg(src + (size_t)(v0 + i) * nplanes + c) : 0u;
                    put(c, v0, w);
                }
            }
        }
    };
    stage_planes(0, (int)min((int64_t)T_VP, len));
    {
        float part = 0.f;
        if (wid < m) {
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                const float r = sub_rn(qr[d], cr[d]);
                sts_f(resid_u + ((wid * 8 + d) * T_RS + lane) * 4, r);
                part = fma_rn(r, r, part);
            }
        }
        sts_f(smin_u + (wid * QG + lane) * 4, part);
    }
    fence_proxy_async();  // generic-proxy writes (A constants) -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = lds_u(tmem_slot);
    stamp();
    float base;
    {
        float rn = 0.f;
        for (int s = 0; s < m; ++s) rn = add_rn(rn, lds_f(smin_u + (s * QG + lane) * 4));  // fixed order
        base = add_rn(lds_f(dc_u + lane * 4), rn);  // dc + |r|^2 over the PQ dims
    }

    // ONE warp writes the A operand of group g: rows (copy, j, q), block 2j = hi, 2j + 1 = lo of
    // r[2g + j][q][0..7]; lane = query.  The caller rotates the warp.
    auto write_A = [&](int g) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float hi[8], lo[8];
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                const float r = lds_f(resid_u + (((2 * g + j) * 8 + d) * T_RS + lane) * 4);
                hi[d] = __uint_as_float(to_tf32(r));
                lo[d] = __uint_as_float(to_tf32(r - hi[d]));
            }
            const uint32_t ph = abuf_u + (2 * j) * TA_BLK + (4 * j + (lane >> 3)) * 256 + (lane & 7) * 16;
#pragma unroll
            for (int c = 0; c < 2; ++c) {  // row = 64 c + 32 j + lane
                sts_v4f(ph + c * 2048, hi[0], hi[1], hi[2], hi[3]);
                sts_v4f(ph + c * 2048 + 128, hi[4], hi[5], hi[6], hi[7]);
                sts_v4f(ph + c * 2048 + TA_BLK, lo[0], lo[1], lo[2], lo[3]);
                sts_v4f(ph + c * 2048 + TA_BLK + 128, lo[4], lo[5], lo[6], lo[7]);
            }
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
    };
    // Build t (issuer lane): the codebook operand of build t must have landed in ring slot t & 1.
    auto issue_mma = [&](int t) {
        const uint32_t slot = t & 1;
        mbar_wait(slot ? bar_full1 : bar_full0, ((uint32_t)t >> 1) & 1, ta.err, 1);
        tc_fence_after();
        const uint32_t bb = bbuf_u + slot * TB_GROUP;
        const uint64_t A0h = tc_smem_desc(abuf_u), A0l = tc_smem_desc(abuf_u + TA_BLK),
                       A1h = tc_smem_desc(abuf_u + 2 * TA_BLK), A1l = tc_smem_desc(abuf_u + 3 * TA_BLK),
                       A1s = tc_smem_desc(abuf_u + 4 * TA_BLK);
        const uint64_t B0h = tc_smem_desc(bb), B0l = tc_smem_desc(bb + TB_BLK), B1h = tc_smem_desc(bb + 2 * TB_BLK),
                       B1l = tc_smem_desc(bb + 3 * TB_BLK), Bn = tc_smem_desc(bb + 4 * TB_BLK);
        tc_mma(tmem_base, A0h, B0h, 0);
        tc_mma(tmem_base, A0l, B0h, 1);
        tc_mma(tmem_base, A0h, B0l, 1);
        tc_mma(tmem_base, A1h, B1h, 1);
        tc_mma(tmem_base, A1l, B1h, 1);
        tc_mma(tmem_base, A1h, B1l, 1);
        tc_mma(tmem_base, A1s, Bn, 1);
        tc_commit(bar_mma);
    };

    if (wid == 2) write_A(0);
    tc_fence_before();
    __syncthreads();
    if (wid == T_ISSUER && lane == 0) {
        tc_fence_after();
        issue_mma(0);
    }
    stamp();

    const uint32_t lo0 = lut_u | (uint32_t)(lane * 4), lo1 = lo0 + 128;
    // epilogue role of this warp: lane quarter -> (copy, subspace-in-group), 32 codeword columns
    const int quarter = wid & 3, ej = quarter & 1;
    const int ecol0 = 32 * ((wid >> 2) * 2 + (quarter >> 1));
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)ecol0;
    const uint32_t edst = lut_u + ej * 128 + lane * 4;  // + row * 256
    // scan role: first chunk, chunk stride, chunks owned
    const int c0 = wid, cs = QWARPS;
    const int maxch = 4;
    const uint32_t pstride = 16 * cs;
    const uint32_t plane_w0 = planes_u + 16 * c0;
    const bool dbg = ta.dbg_lut != nullptr && item == 0;

    float acc[QNV];
    int nv = 0, nch = 0;
    int64_t vbase = 0;
    int g = 0;
    for (int t = 0; t < T; ++t) {
        if (g == 0) {
            vbase = (int64_t)(t / ng) * T_VP;
            nv = (int)min((int64_t)T_VP, len - vbase);
            nch = min(maxch, max(0, (nv - 16 * c0 + (int)pstride - 1) / (int)pstride));  // chunks j4 with 16 (c0 + cs j4) < nv
        }
        // ---- epilogue of build t: tensor memory -> table layout of the scan ----
        mbar_wait(bar_mma, t & 1, ta.err, 2);
        tc_fence_after();
        {
            uint32_t v[32];
            tc_ld32(taddr, v);
            tc_wait_ld();
            if constexpr (IDENT) {
                // rows of codewords >= ksub are never looked up: store unconditionally
#pragma unroll
                for (int i = 0; i < 32; ++i) sts_u(edst + (ecol0 + i) * 256, v[i]);
            } else {
                const uint8_t* cvp = a.cb_codes + (size_t)(2 * g + ej) * a.ksub;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int code = ecol0 + i;
                    if (code < a.ksub) sts_u(edst + (int)cvp[code] * 256, v[i]);  // code VALUE -> entry (Q6)
                }
            }
        }
        const int gn = g + 1 == ng ? 0 : g + 1;
        if (t + 1 < T && wid == ((t + 3) & 7)) write_A(gn);  // build t has completed: its A operand is free
        tc_fence_before();
        __syncthreads();  // table complete, accumulator tile drained, next A operand written
        if (wid == T_ISSUER && lane == 0 && t + 1 < T) {
            tc_fence_after();
            if (t + 2 < T) {  // ring slot t & 1 was last read by build t (complete)
                const uint32_t bar = (t & 1) ? bar_full1 : bar_full0;
                const int g2 = gn + 1 == ng ? 0 : gn + 1;
                mbar_expect_tx(bar, TB_GROUP);
                tma_bulk_g2s(bbuf_u + (t & 1) * TB_GROUP, ta.tcB + (size_t)g2 * (TB_GROUP / 4), TB_GROUP, bar);
            }
            issue_mma(t + 1);
        }
        __syncwarp();
        if (dbg && t < ng) {
            if (tid < QG) reinterpret_cast<int*>(ta.dbg_lut + (size_t)m * 256 * 32)[tid] = (int)lds_u(pair_u + tid * 4);
            if (tid == 0) reinterpret_cast<int*>(ta.dbg_lut + (size_t)m * 256 * 32)[QG] = cell;
            for (int idx = tid; idx < 2 * 256 * 32; idx += QTHREADS) {
                const int q = idx & 31, code = (idx >> 5) & 255, j = idx >> 13;
                ta.dbg_lut[((size_t)(2 * g + j) * 256 + code) * 32 + q] = lds_f(lut_u + code * 256 + j * 128 + q * 4);
            }
        }
        // ---- K3: scan the two subspaces of this group (byte planes 2g, 2g + 1) ----
        {
            const uint32_t plane_w = plane_w0 + (2 * g) * T_PLANE;
            if (nch == 4) {
                if (g == 0) scant_group<true, true>(lo0, lo1, plane_w, pstride, nch, base, acc);
                else scant_group<false, true>(lo0, lo1, plane_w, pstride, nch, base, acc);
            } else {
                if (g == 0) scant_group<true, false>(lo0, lo1, plane_w, pstride, nch, base, acc);
                else scant_group<false, false>(lo0, lo1, plane_w, pstride, nch, base, acc);
            }
        }
        __syncthreads();  // every warp is done with this table
        g = gn;
        stamp();
        if (g != 0) continue;

        // ---- per-(query, list) top-k of this pass ----
        // mask the slots beyond the list (chunks not owned / not reached, and the chunk that straddles the end)
        {
            const int lim = nv - 16 * c0;  // slot 16 * j4 + i holds a vector iff pstride * j4 + i < lim
#pragma unroll
            for (int j4 = 0; j4 < QNV / 16; ++j4) {
                if (j4 >= maxch || (int)pstride * j4 + 16 > lim) {  // warp-uniform
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (j4 >= maxch || (int)pstride * j4 + i >= lim) acc[16 * j4 + i] = Limits<float>::inf();
                }
            }
        }
        const int64_t pbase = vbase + 16 * c0;  // position of slot 0 of this warp
        if (t + 1 < T) stage_planes(vbase + T_VP, (int)min((int64_t)T_VP, len - vbase - T_VP));  // next pass
        float mn = Limits<float>::inf();
#pragma unroll
        for (int j = 0; j < QNV; ++j) mn = fminf(mn, acc[j]);
        sts_f(smin_u + (wid * QG + lane) * 4, mn);
        __syncthreads();
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
        __syncthreads();
        {
            const float thr = lds_f(thr_u + lane * 4);
            // keep d <= bound; while fewer than k vectors have been seen (bound = +inf) keep every real slot
            const float cut = thr < Limits<float>::inf() ? thr : 3.402823466e+38f;
            int c = 0;
#pragma unroll
            for (int j = 0; j < QNV; ++j) c += acc[j] <= cut ? 1 : 0;
            int slot = c ? atoms_add(cnt_u + lane * 4, c) : 0;  // ONE shared-memory atomic per (warp, query)
#pragma unroll
            for (int j = 0; j < QNV; ++j) {
                if (acc[j] <= cut) {
                    if (slot < QCAP) {
                        sts_f(cand_d_u + (lane * QCAP + slot) * 4, acc[j]);
                        sts_u(cand_p_u + (lane * QCAP + slot) * 4, (uint32_t)(pbase + (int)pstride * (j >> 4) + (j & 15)));
                    }
                    ++slot;
                }
            }
        }
        __syncthreads();
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
#pragma unroll 4
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
        __syncthreads();
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
    __syncthreads();
    if (wid == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(T_TMEM_COLS)
                     : "memory");
    }
}

// Codebook -> B operand blocks of the tensor-core table builder, once at create.
// Block layout (fp32 words): word(n, k) = (n >> 3) * 64 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3).
__global__ void prep_tc_kernel(const float* __restrict__ cb, int m, int ksub, int dsub, float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int ng = m >> 1;
    if (idx >= ng * TB_NBLK * 2048) return;
    const int kk = idx & 7, n = (idx >> 3) & 255, b = (idx >> 11) % TB_NBLK, g = idx / (2048 * TB_NBLK);
    float val = 0.f;
    if (b < 4) {
        const int s = 2 * g + (b >> 1);
        const float v = (n < ksub && kk < dsub) ? -2.f * cb[((size_t)s * ksub + n) * dsub + kk] : 0.f;
        const float hi = __uint_as_float(to_tf32(v));
        val = (b & 1) ? __uint_as_float(to_tf32(v - hi)) : hi;
    } else if (kk < 4 && n < ksub) {
        const int s = 2 * g + (kk >> 1);
        float nrm = 0.f;
        for (int d = 0; d < dsub; ++d) {
            const float w = cb[((size_t)s * ksub + n) * dsub + d];
            nrm = fma_rn(w, w, nrm);
        }
        const float hi = __uint_as_float(to_tf32(nrm));
        val = (kk & 1) ? __uint_as_float(to_tf32(nrm - hi)) : hi;
    }
    out[(size_t)(g * TB_NBLK + b) * 2048 + (n >> 3) * 64 + (kk >> 2) * 32 + (n & 7) * 4 + (kk & 3)] = val;
}

}  // namespace ivf
