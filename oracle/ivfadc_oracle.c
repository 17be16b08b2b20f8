/*
 * ============================================================================================
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 * CPU restatement ("oracle") of the hot path of JuliaNeighbors/IVFADC.jl.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and only as the checker or the timed CPU baseline; the product (libivfadc_cuda)
 * never links, imports or falls back to it.
 *
 * PARITY UNPINNED for numeric values: the reference is pure Julia and neither `julia` nor its
 * five un-vendored dependencies (Distances ^0.10, QuantizedArrays ^0.1.6, Clustering ^0.15,
 * DataStructures ^0.18, HNSW ^0.1 -- reference Project.toml:6-19, no Manifest.toml) exist in this
 * image, so the reference cannot be executed here and it ships no golden vectors.  What IS
 * pinned: the reference's only known-answer test (test/search.jl:26-49, the 2x13 toy matrix)
 * and the integer semantics of delete/renumber (test/utils.jl:58-105); both are replayed against
 * this oracle in tests/test_oracle_reference_tests.py.  Everything else follows the reference
 * source line by line (citations inline) with the dependency semantics of SURVEY.md Appendix C.
 * ============================================================================================
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* metrics of the next calls (0 SqEuclidean, 1 Euclidean, 2 Cityblock, 3 CosineDist): Dc for coarse_search and the
 * lookup tables (src/coarsequantizers.jl:34, src/index.jl:234), Dr for quantize_data (src/index.jl:187, src/utils.jl:158) */
static int g_metric_coarse = 0, g_metric_resid = 0;
void oracle_set_metrics(int coarse, int resid) { g_metric_coarse = coarse; g_metric_resid = resid; }

#define T float
#define FN(name) name##_f32
#define FMA fmaf
#include "ivfadc_oracle_impl.h"
#undef T
#undef FN
#undef FMA

#define T double
#define FN(name) name##_f64
#define FMA fma
#include "ivfadc_oracle_impl.h"
#undef T
#undef FN
#undef FMA

int oracle_abi_version(void) { return 1; }

/*
 * ---- synthetic inputs: counter-based generator ------------------------------------------
 * Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11) keyed by the seed, counter =
 * (vector index lo, hi, dim or dim / 4, stream).  Every value depends only on (seed, vector, dim), so any slice
 * of a 100 M-vector data set can be produced anywhere; ivfadc.jl_b200/csrc/synth.cu is the same function on the
 * GPU, bit for bit (integer arithmetic, then one exact int -> float conversion, one multiplication and one
 * addition, each rounded once).
 *   stream 0: noise of (vector, dim): sum of four 22-bit uniforms, centred -> t in (-2^23, 2^23), an Irwin-Hall
 *             approximation of a Gaussian with standard deviation 2^22 / sqrt(3)
 *   stream 1: the blob of a vector (dim = 0): (r0 * n_blobs) >> 32
 *   stream 2: uniforms in [0, 1): 24 bits, four dims per counter
 */
static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

/* X[n][D] = uniforms in [0, 1) of vectors first .. first + n - 1 */
void oracle_synth_uniform_f32(float* X, int64_t first, int64_t n, int D, uint64_t seed) {
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const uint64_t v = (uint64_t)(first + i);
        for (int d = 0; d < D; d += 4) {
            uint32_t r[4];
            philox4x32_10((uint32_t)v, (uint32_t)(v >> 32), (uint32_t)(d >> 2), 2u, k0, k1, r);
            for (int j = 0; j < 4 && d + j < D; ++j) X[i * D + d + j] = (float)(r[j] >> 8) * 5.9604644775390625e-08f;
        }
    }
}

/* blob of every vector (blobs_out, optional) and X[n][D] = centres[blob] + t * scale (scale = sigma sqrt(3) / 2^22) */
void oracle_synth_blobs_f32(float* X, int32_t* blobs_out, int64_t first, int64_t n, int D, int n_blobs, uint64_t seed,
                            float scale, const float* centres) {
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const uint64_t v = (uint64_t)(first + i);
        uint32_t r[4];
        philox4x32_10((uint32_t)v, (uint32_t)(v >> 32), 0u, 1u, k0, k1, r);
        const int b = (int)(((uint64_t)r[0] * (uint64_t)n_blobs) >> 32);
        if (blobs_out) blobs_out[i] = b;
        if (!X) continue;
        for (int d = 0; d < D; ++d) {
            philox4x32_10((uint32_t)v, (uint32_t)(v >> 32), (uint32_t)d, 0u, k0, k1, r);
            const int32_t t = (int32_t)((r[0] >> 10) + (r[1] >> 10) + (r[2] >> 10) + (r[3] >> 10)) - 8388606;
            const float p = (float)t * scale;
            X[i * D + d] = centres[(size_t)b * D + d] + p;
        }
    }
}
