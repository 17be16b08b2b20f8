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
