// K2 + K3, "query-per-lane" variant -- the batched search hot loop of knn_search (reference
// src/index.jl:228-255) for large query batches (fp32, k <= 16, m in {4, 8, 12, 16}).
//
// Work item = one inverted list x up to 32 queries that probe it.  One CTA of 16 warps per item,
// one CTA per SM.  Lane l of EVERY warp owns query l of the group; warp w owns the database
// vectors 4*(w + 16*jj) + u (jj < 16, u < 4) of the current 1024-vector pass and keeps their 64
// partial distances in registers.  Hence
//   * the PQ code byte of a (vector, subspace) is warp-uniform and the lookup
//     lut[subspace][code][lane] is one conflict-free 128-byte shared-memory wavefront: 32 lookups
//     per LDS, the shared-memory peak (vs. ~2.6-way bank conflicts of the vector-per-lane layout);
//   * every lane adds the table entries of ITS query in subspace order 1..m, so each distance is
//     the same sequential chain  ((dc + l_1) + l_2) + ...  as the reference (src/index.jl:242-246)
//     and the oracle (A3): results are bit-identical.
// A 32-query table is m*256*32*4 B (512 KB for m = 16) and cannot live in shared memory at once:
// the subspaces are processed in chunks of 4 (128 KB), the 64 partial sums per lane stay in
// registers across chunks.
//
//   K2  per chunk: lut[il][code][q] = sum_d (w_icd - r_qd)^2, r = q - centroid
//       (src/index.jl:232-236; subspace 0 additionally carries dc = first addition of the chain);
//   K3  per chunk: stream the list's code words (staged once per pass into per-chunk planes),
//       gather-add;
//   top-k per (query, list): k-th smallest of the 16 per-warp minima bounds the k-th distance,
//       candidates <= bound go to a 64-slot shared list, exact (distance, position) selection.
//       A list that overflows (heavy ties) is handed to the general kernel through a redo queue.
#pragma once

#include "common.cuh"

namespace ivf {

constexpr int QTHREADS = 512;
constexpr int QWARPS = QTHREADS / 32;
constexpr int QG = 32;                 // queries per work item (= lanes)
constexpr int QCS = 4;                 // subspaces per table chunk
constexpr int QNV = 64;                // partial distances per lane
constexpr int QVP = QWARPS * QNV;      // vectors per pass (1024)
constexpr int QCAP = 64;               // candidate slots per query
constexpr int QPLANE = QVP + 8;        // words per code plane (+8: conflict-free staging)
constexpr int QMAXK = 16;
constexpr int QRS = 33;                // row stride of the transposed residuals

struct ScanQArgs {
    const float* Q;          // [nq][D]
    const float* C;          // [kc][D]
    const float* cb;         // [m][ksub][dsub]
    const uint8_t* cb_codes;
    int cb_identity;
    int D, m, dsub, ksub, kc, w, k;
    const int64_t* list_off;
    const int64_t* list_len;
    const uint8_t* codes;
    const float* dc;             // [nq][w]
    const int* bucket_off;       // [kc+1] pairs
    const int* group_off;        // [kc+1] work items
    const int32_t* sorted_pairs;
    float* pair_d;               // [npairs][k]
    uint32_t* pair_pos;          // [npairs][k]
    int32_t* pair_cnt;           // [npairs]
    int32_t* redo_pairs;         // [npairs]
    int* redo_cnt;
};

// Table chunk layout (bytes): subspace il of the chunk, code value c, query q live at
//   (il >> 1) * 65536 + c * 256 + (il & 1) * 128 + q * 4
// i.e. 256-byte rows holding the entries of TWO subspaces for the same code value.  The row
// stride of exactly 256 B lets ONE byte-permute build the address  c << 8 | lane-offset  from the
// packed code word (no shift / mask / scale), and lanes = queries make every access a single
// conflict-free 128-byte wavefront.
__device__ __forceinline__ uint32_t lut_off(int il, int code) {
    return (uint32_t)((il >> 1) * 65536 + code * 256 + (il & 1) * 128);
}

struct ScanQSmem {
    // byte offsets into dynamic shared memory
    size_t lut, resid, planes, cand_d, cand_p, smin, misc, total;
};

__host__ __device__ inline ScanQSmem scanq_smem_layout(int m, int dsub, int ldw) {
    ScanQSmem s;
    size_t o = 0;
    s.lut = o;    o += (size_t)QCS * 256 * 32 * 4;
    s.resid = o;  o += (size_t)m * dsub * QRS * 4;
    o = (o + 15) & ~(size_t)15;
    s.planes = o; o += (size_t)(m / QCS) * QPLANE * 4;
    s.cand_d = o; o += (size_t)QG * QCAP * 4;
    s.cand_p = o; o += (size_t)QG * QCAP * 4;
    s.smin = o;   o += (size_t)QWARPS * QG * 4;
    s.misc = o;   o += 6 * QG * 4;
    s.total = o;
    return s;
}

// ---- K2 (exact): direct-form table chunk, one sequential fma chain per entry (oracle A1) --------
template <int DSUB, int LDW>
__device__ __forceinline__ void build_chunk_exact(const ScanQArgs& a, float* lut, const float* resid_t,
                                                  const float* s_dc, int c) {
    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    const int dsub = DSUB > 0 ? DSUB : a.dsub;
    const int total = QCS * a.ksub;
    const int per_warp = (total + QWARPS - 1) / QWARPS;
    const int e0 = wid * per_warp;
    const int e1 = min(total, e0 + per_warp);
    for (int il = 0; il < QCS; ++il) {
        const int lo = max(e0, il * a.ksub) - il * a.ksub;
        const int hi = min(e1, (il + 1) * a.ksub) - il * a.ksub;
        if (lo >= hi) continue;
        const int i = c * QCS + il;
        const float* rq = resid_t + (size_t)i * dsub * QRS + lane;
        const float dc0 = i == 0 ? s_dc[lane] : 0.f;
        char* lrow = reinterpret_cast<char*>(lut) + lut_off(il, 0) + lane * 4;
        if constexpr (DSUB > 0) {
            float r[DSUB > 0 ? DSUB : 1];
#pragma unroll
            for (int d = 0; d < DSUB; ++d) r[d] = rq[d * QRS];
#pragma unroll 4
            for (int code = lo; code < hi; ++code) {
                const float* wv = a.cb + ((size_t)i * a.ksub + code) * DSUB;
                float wreg[DSUB > 0 ? DSUB : 1];
                if constexpr (DSUB % 4 == 0) {
#pragma unroll
                    for (int d = 0; d < DSUB; d += 4) {
                        const float4 t = __ldg(reinterpret_cast<const float4*>(wv + d));
                        wreg[d] = t.x; wreg[d + 1] = t.y; wreg[d + 2] = t.z; wreg[d + 3] = t.w;
                    }
                } else {
#pragma unroll
                    for (int d = 0; d < DSUB; ++d) wreg[d] = __ldg(wv + d);
                }
                float s = 0.f;
#pragma unroll
                for (int d = 0; d < DSUB; ++d) {
                    const float diff = sub_rn(wreg[d], r[d]);  // oracle A1: codeword - residual
                    s = fma_rn(diff, diff, s);
                }
                if (i == 0) s = add_rn(dc0, s);  // d = dc; d += l_1
                const int cv = a.cb_identity ? code : (int)a.cb_codes[(size_t)i * a.ksub + code];
                *reinterpret_cast<float*>(lrow + cv * 256) = s;
            }
        } else {
            for (int code = lo; code < hi; ++code) {
                const float* wv = a.cb + ((size_t)i * a.ksub + code) * dsub;
                float s = 0.f;
                for (int d = 0; d < dsub; ++d) {
                    const float diff = sub_rn(__ldg(wv + d), rq[d * QRS]);
                    s = fma_rn(diff, diff, s);
                }
                if (i == 0) s = add_rn(dc0, s);
                const int cv = a.cb_identity ? code : (int)a.cb_codes[(size_t)i * a.ksub + code];
                *reinterpret_cast<float*>(lrow + cv * 256) = s;
            }
        }
    }
}

// ---- K3: one chunk of 4 subspaces over the 64 vectors of this warp ------------------------------
template <bool FIRST>
__device__ __forceinline__ float scanq_vec(const char* lutb, uint32_t lo0, uint32_t lo1, uint32_t x, float acc) {
#pragma unroll
    for (int il = 0; il < QCS; ++il) {
        // address = code << 8 | (lane * 4 [+ 128]) in one PRMT: byte 1 <- byte il of the code word
        const uint32_t off = __byte_perm(x, (il & 1) ? lo1 : lo0, 0x5504 | (il << 4));
        const float v = *reinterpret_cast<const float*>(lutb + (il >> 1) * 65536 + off);
        acc = (FIRST && il == 0) ? v : add_rn(acc, v);  // oracle A3: strictly in subspace order
    }
    return acc;
}

template <bool FIRST>
__device__ __forceinline__ void scanq_chunk(const char* lutb, uint32_t lo0, uint32_t lo1, const uint32_t* plane,
                                            int wid, int nv, float (&acc)[QNV]) {
#pragma unroll
    for (int jj = 0; jj < QNV / 4; ++jj) {
        const int g = wid + QWARPS * jj;
        if (4 * g < nv) {  // warp-uniform
            const uint4 x = *reinterpret_cast<const uint4*>(plane + 4 * g);
            acc[4 * jj + 0] = scanq_vec<FIRST>(lutb, lo0, lo1, x.x, acc[4 * jj + 0]);
            acc[4 * jj + 1] = scanq_vec<FIRST>(lutb, lo0, lo1, x.y, acc[4 * jj + 1]);
            acc[4 * jj + 2] = scanq_vec<FIRST>(lutb, lo0, lo1, x.z, acc[4 * jj + 2]);
            acc[4 * jj + 3] = scanq_vec<FIRST>(lutb, lo0, lo1, x.w, acc[4 * jj + 3]);
        }
    }
}

__device__ __forceinline__ bool cand_before(float da, uint32_t pa, float db, uint32_t pb) {
    return da < db || (da == db && pa < pb);
}

template <int LDW>
__global__ void __launch_bounds__(QTHREADS, 1)
scanq_kernel(const ScanQArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int wid = tid >> 5;
    const int m = a.m;
    const int k = a.k;
    const int nchunks = m / QCS;
    const int Dp = m * a.dsub;

    // ---- work item -> (cell, group) ----
    const int item = blockIdx.x;
    if (item >= a.group_off[a.kc]) return;
    int lo = 0, hi = a.kc;  // invariant: group_off[lo] <= item < group_off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (a.group_off[mid] <= item) lo = mid; else hi = mid;
    }
    const int cell = lo;
    const int first = a.bucket_off[cell] + (item - a.group_off[cell]) * QG;
    const int nj = min(QG, a.bucket_off[cell + 1] - first);

    const ScanQSmem L = scanq_smem_layout(m, a.dsub, LDW);
    float* lut = reinterpret_cast<float*>(smem_raw + L.lut);
    float* resid_t = reinterpret_cast<float*>(smem_raw + L.resid);
    uint32_t* planes = reinterpret_cast<uint32_t*>(smem_raw + L.planes);
    float* cand_d = reinterpret_cast<float*>(smem_raw + L.cand_d);
    uint32_t* cand_p = reinterpret_cast<uint32_t*>(smem_raw + L.cand_p);
    float* s_min = reinterpret_cast<float*>(smem_raw + L.smin);
    float* s_dc = reinterpret_cast<float*>(smem_raw + L.misc);
    float* s_thr = s_dc + QG;
    float* s_run = s_thr + QG;
    int* s_cnt = reinterpret_cast<int*>(s_run + QG);
    int* s_pair = s_cnt + QG;
    int* s_flag = s_pair + QG;

    if (tid < QG) {
        const int p = tid < nj ? a.sorted_pairs[first + tid] : -1;
        s_pair[tid] = p;
        s_dc[tid] = p >= 0 ? a.dc[p] : 0.f;
        s_run[tid] = Limits<float>::inf();
        s_cnt[tid] = 0;
        s_flag[tid] = 0;
    }
    __syncthreads();

    // residuals r_q = query - centroid (reference _closest_cluster_residuals,
    // src/coarsequantizers.jl:40-45), transposed [d][q] so that lane q reads its own column
    for (int idx = tid; idx < QG * Dp; idx += QTHREADS) {
        const int q = idx / Dp, d = idx - q * Dp;
        const int p = s_pair[q];
        resid_t[d * QRS + q] =
            p >= 0 ? sub_rn(a.Q[(size_t)(p / a.w) * a.D + d], a.C[(size_t)cell * a.D + d]) : 0.f;
    }

    const int64_t len = a.list_len[cell];
    const uint32_t* gcodes = reinterpret_cast<const uint32_t*>(a.codes + (size_t)a.list_off[cell] * m);
    const char* lutb = reinterpret_cast<const char*>(lut);
    const uint32_t lo0 = lane * 4, lo1 = lane * 4 + 128;

    for (int64_t base = 0; base < len; base += QVP) {
        const int nv = (int)min((int64_t)QVP, len - base);
        __syncthreads();  // previous pass fully consumed (planes, candidates) / residuals written
        // stage the code words of this pass: plane[c][v] = bytes 4c..4c+3 of vector v
        {
            const uint32_t* src = gcodes + (size_t)base * nchunks;
            const int nwords = nv * nchunks;
            const int padded = ((nv + 3) & ~3) * nchunks;
            for (int idx = tid; idx < padded; idx += QTHREADS) {
                const int v = idx / nchunks, c = idx - v * nchunks;
                planes[c * QPLANE + v] = idx < nwords ? __ldg(src + idx) : 0u;
            }
        }
        float acc[QNV];
#pragma unroll
        for (int j = 0; j < QNV; ++j) acc[j] = 0.f;

        for (int c = 0; c < nchunks; ++c) {
            if (c > 0) __syncthreads();  // every warp is done scanning the previous table chunk
            switch (a.dsub) {
                case 4: build_chunk_exact<4, LDW>(a, lut, resid_t, s_dc, c); break;
                case 8: build_chunk_exact<8, LDW>(a, lut, resid_t, s_dc, c); break;
                case 16: build_chunk_exact<16, LDW>(a, lut, resid_t, s_dc, c); break;
                default: build_chunk_exact<0, LDW>(a, lut, resid_t, s_dc, c); break;
            }
            __syncthreads();
            if (c == 0) scanq_chunk<true>(lutb, lo0, lo1, planes, wid, nv, acc);
            else scanq_chunk<false>(lutb, lo0, lo1, planes + c * QPLANE, wid, nv, acc);
        }

        // ---- per-(query, list) top-k of this pass ----
        const int lim = nv - 4 * wid;  // slot j of this warp holds a vector iff 64*(j/4) + j%4 < lim
        float mn = Limits<float>::inf();
#pragma unroll
        for (int j = 0; j < QNV; ++j) {
            if (16 * (j & ~3) + (j & 3) < lim) mn = fminf(mn, acc[j]);
        }
        s_min[wid * QG + lane] = mn;
        __syncthreads();
        {
            int rank = 0;
#pragma unroll
            for (int w2 = 0; w2 < QWARPS; ++w2) {
                const float o = s_min[w2 * QG + lane];
                rank += (o < mn || (o == mn && w2 < wid)) ? 1 : 0;
            }
            // the k-th smallest of 16 distinct candidates bounds the k-th smallest of all
            if (rank == min(k, QWARPS) - 1) s_thr[lane] = fminf(mn, s_run[lane]);
        }
        __syncthreads();
        {
            const float thr = s_thr[lane];
#pragma unroll
            for (int j = 0; j < QNV; ++j) {
                const int rel = 16 * (j & ~3) + (j & 3);
                if (rel < lim && acc[j] <= thr) {
                    const int slot = atomicAdd(&s_cnt[lane], 1);
                    if (slot < QCAP) {
                        cand_d[lane * QCAP + slot] = acc[j];
                        cand_p[lane * QCAP + slot] = (uint32_t)(base + 4 * wid + rel);
                    }
                }
            }
        }
        __syncthreads();
        // exact selection by (distance, position): warp w serves queries w and w + 16
        for (int q = wid; q < QG; q += QWARPS) {
            int n = s_cnt[q];
            const bool ovf = n > QCAP;
            n = min(n, QCAP);
            float d0 = Limits<float>::inf(), d1 = Limits<float>::inf();
            uint32_t p0 = kNoPos, p1 = kNoPos;
            if (lane < n) { d0 = cand_d[q * QCAP + lane]; p0 = cand_p[q * QCAP + lane]; }
            if (lane + 32 < n) { d1 = cand_d[q * QCAP + lane + 32]; p1 = cand_p[q * QCAP + lane + 32]; }
            int r0 = 0, r1 = 0;
            for (int e = 0; e < n; ++e) {
                const float de = cand_d[q * QCAP + e];
                const uint32_t pe = cand_p[q * QCAP + e];
                r0 += cand_before(de, pe, d0, p0) ? 1 : 0;
                r1 += cand_before(de, pe, d1, p1) ? 1 : 0;
            }
            __syncwarp();
            if (lane < n && r0 < k) { cand_d[q * QCAP + r0] = d0; cand_p[q * QCAP + r0] = p0; }
            if (lane + 32 < n && r1 < k) { cand_d[q * QCAP + r1] = d1; cand_p[q * QCAP + r1] = p1; }
            __syncwarp();
            if (lane == 0) {
                const int cnt = min(n, k);
                s_cnt[q] = cnt;
                s_run[q] = cnt >= k ? cand_d[q * QCAP + k - 1] : Limits<float>::inf();
                if (ovf) s_flag[q] = 1;
            }
        }
    }
    __syncthreads();

    // ---- publish ----
    for (int idx = tid; idx < nj * k; idx += QTHREADS) {
        const int q = idx / k, e = idx - q * k;
        const int pair = s_pair[q];
        if (!s_flag[q] && e < s_cnt[q]) {
            a.pair_d[(size_t)pair * k + e] = cand_d[q * QCAP + e];
            a.pair_pos[(size_t)pair * k + e] = cand_p[q * QCAP + e];
        }
    }
    if (tid < nj) {
        const int pair = s_pair[tid];
        if (s_flag[tid]) {
            a.pair_cnt[pair] = 0;
            a.redo_pairs[atomicAdd(a.redo_cnt, 1)] = pair;
        } else {
            a.pair_cnt[pair] = s_cnt[tid];
        }
    }
}

}  // namespace ivf
