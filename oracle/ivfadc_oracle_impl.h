/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  See ivfadc_oracle.c for the header that applies.
 *
 * Type-generic body of the CPU restatement; included twice with
 *   #define T float  / double,  #define FN(name) name##_f32 / _f64,  #define FMA fmaf / fma
 *
 * Arithmetic conventions (one place, so that the CUDA path can match them bit for bit):
 *   A1  Distances.colwise(SqEuclidean(), A, b), column c:  s = 0; for i in rows: d = A[i,c] - b[i];
 *       s = fma(d, d, s)   -- rows in increasing order.  (Distances.jl evaluates the direct form
 *       in an @simd loop; @simd leaves association/contraction to the compiler, so no CPU order is
 *       "the" Julia order; sequential-with-fma is the one this oracle fixes.)
 *   A2  Distances.pairwise(SqEuclidean(), A, B; dims=2) (GEMM form used by QuantizedArrays.encode):
 *       dot = 0; for i in rows: dot = fma(A[i,a], B[i,b], dot);   sa2, sb2 likewise with fma(x,x,s);
 *       v = (sa2[a] + sb2[b]) - 2*dot;  v = max(v, 0).
 *   A3  ADC: d = dc; for i = 1..m: d = d + lut[i][code_i]           (src/index.jl:242-246)
 */

/* A1 -- colwise(SqEuclidean) for one column. */
static inline T FN(sqdist_direct)(const T* a, const T* b, int n) {
    T s = (T)0;
    for (int i = 0; i < n; ++i) {
        T d = a[i] - b[i];
        s = FMA(d, d, s);
    }
    return s;
}

/*
 * Other metrics of Distances.jl (the type parameters Dc / Dr admit any PreMetric, src/index.jl:41-42,108-109).
 * Dependency semantics restated from the published definitions of Distances ^0.10 -- UNVERIFIABLE here, one
 * swappable function each:
 *   Euclidean   colwise: sqrt(sum (a - b)^2)                                 (the SqEuclidean chain, then one sqrt)
 *   Cityblock   colwise: sum |a - b|, sequential
 *   CosineDist  colwise: max(1 - dot / (sqrt(sum a^2) sqrt(sum b^2)), 0), three sequential fma chains
 * metric codes: 0 SqEuclidean, 1 Euclidean, 2 Cityblock, 3 CosineDist (oracle_set_metrics).
 */
static inline T FN(dist_colwise)(int metric, const T* a, const T* b, int n) {
    if (metric == 0) return FN(sqdist_direct)(a, b, n);
    if (metric == 1) return (sizeof(T) == 4) ? (T)sqrtf((float)FN(sqdist_direct)(a, b, n)) : (T)sqrt((double)FN(sqdist_direct)(a, b, n));
    if (metric == 2) {
        T s = (T)0;
        for (int i = 0; i < n; ++i) {
            T d = a[i] - b[i];
            s = s + (d < (T)0 ? -d : d);
        }
        return s;
    }
    T ab = (T)0, a2 = (T)0, b2 = (T)0;
    for (int i = 0; i < n; ++i) {
        ab = FMA(a[i], b[i], ab);
        a2 = FMA(a[i], a[i], a2);
        b2 = FMA(b[i], b[i], b2);
    }
    const T na = (sizeof(T) == 4) ? (T)sqrtf((float)a2) : (T)sqrt((double)a2);
    const T nb = (sizeof(T) == 4) ? (T)sqrtf((float)b2) : (T)sqrt((double)b2);
    const T v = (T)1 - ab / (na * nb);
    return v > (T)0 ? v : (T)0;   /* a zero vector gives NaN in Julia (max(NaN, 0)); here 0 -- degenerate input, not pinned */
}

static inline T FN(sumsq)(const T* a, int n) {
    T s = (T)0;
    for (int i = 0; i < n; ++i) s = FMA(a[i], a[i], s);
    return s;
}

/*
 * coarse_search(::NaiveQuantizer, point, w)          reference src/coarsequantizers.jl:33-37
 *   coarse_distances = colwise(D(), cq.vectors, point)         :34
 *   closest_clusters = sortperm(coarse_distances)[1:w]         :35  (stable: ties -> lower cell)
 * cells are returned 0-based.  scratch: T[kc].
 */
static void FN(coarse_one)(const T* centroids, int kc, int D, const T* q, int w, int32_t* cells,
                           T* dc, T* scratch) {
    for (int c = 0; c < kc; ++c) scratch[c] = FN(dist_colwise)(g_metric_coarse, centroids + (size_t)c * D, q, D);
    /* partial stable selection of the w smallest (distance, cell): insertion into a sorted
       prefix; equivalent to sortperm(...)[1:w]. */
    int cnt = 0;
    for (int c = 0; c < kc; ++c) {
        T d = scratch[c];
        if (cnt == w && !(d < dc[cnt - 1])) continue;
        int pos = cnt < w ? cnt : w - 1;
        while (pos > 0 && d < dc[pos - 1]) {
            dc[pos] = dc[pos - 1];
            cells[pos] = cells[pos - 1];
            --pos;
        }
        dc[pos] = d;
        cells[pos] = c;
        if (cnt < w) ++cnt;
    }
}

void FN(oracle_coarse_search)(const T* centroids, int kc, int D, const T* Q, int64_t nq, int w,
                              int32_t* cells_out, T* dc_out, int nthreads) {
    if (w > kc) w = kc;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        T* scratch = (T*)malloc(sizeof(T) * (size_t)kc);
#pragma omp for schedule(dynamic, 16)
        for (int64_t i = 0; i < nq; ++i)
            FN(coarse_one)(centroids, kc, D, Q + i * D, w, cells_out + i * w, dc_out + i * w,
                           scratch);
        free(scratch);
    }
}

/*
 * QuantizedArrays.quantize_data(rq, residual)  as called at reference src/index.jl:187 and
 * src/utils.jl:158: per codebook i, rows rowrange(D, m, i) (= dsub*(i-1)+1 : dsub*i with
 * dsub = floor(D/m)), pairwise GEMM-form distances to the ksub codewords (A2), first minimum,
 * stored byte = codebook.codes[argmin].
 *   cb_vectors T[m][ksub][dsub], cb_codes uint8[m][ksub], cb_norms T[m][ksub] (= sa2)
 */
static void FN(encode_residual)(const T* resid, int D, int m, int ksub, const T* cb_vectors,
                                const uint8_t* cb_codes, const T* cb_norms, uint8_t* code_out) {
    const int dsub = D / m;
    for (int i = 0; i < m; ++i) {
        const T* x = resid + (size_t)i * dsub;
        const T sb = FN(sumsq)(x, dsub);
        T best = (T)0;
        int besti = -1;
        for (int c = 0; c < ksub; ++c) {
            const T* wv = cb_vectors + ((size_t)i * ksub + c) * dsub;
            T v;
            if (g_metric_resid == 0 || g_metric_resid == 1) {
                /* pairwise(SqEuclidean): GEMM form (A2); pairwise(Euclidean): its square root */
                T dot = (T)0;
                for (int d = 0; d < dsub; ++d) dot = FMA(wv[d], x[d], dot);
                v = (cb_norms[(size_t)i * ksub + c] + sb) - (T)2 * dot;
                v = v > (T)0 ? v : (T)0;
                if (g_metric_resid == 1) v = (sizeof(T) == 4) ? (T)sqrtf((float)v) : (T)sqrt((double)v);
            } else if (g_metric_resid == 2) {
                v = FN(dist_colwise)(2, wv, x, dsub);   /* pairwise(Cityblock): the generic loop, evaluate per pair */
            } else {
                /* pairwise(CosineDist): dot products by GEMM, norms = sqrt of the column sums of squares */
                T dot = (T)0;
                for (int d = 0; d < dsub; ++d) dot = FMA(wv[d], x[d], dot);
                const T na = (sizeof(T) == 4) ? (T)sqrtf((float)cb_norms[(size_t)i * ksub + c]) : (T)sqrt((double)cb_norms[(size_t)i * ksub + c]);
                const T nb = (sizeof(T) == 4) ? (T)sqrtf((float)sb) : (T)sqrt((double)sb);
                v = (T)1 - dot / (na * nb);
                v = v > (T)0 ? v : (T)0;
            }
            if (besti < 0 || v < best) {
                best = v;
                besti = c;
            }
        }
        code_out[i] = cb_codes[(size_t)i * ksub + besti];
    }
}

void FN(oracle_codebook_norms)(const T* cb_vectors, int m, int ksub, int dsub, T* norms_out) {
    for (size_t e = 0; e < (size_t)m * ksub; ++e) norms_out[e] = FN(sumsq)(cb_vectors + e * dsub, dsub);
}

/*
 * _encode_point (reference src/utils.jl:148-161) for a batch, or the build path
 * (_build_residuals + _build_inverted_index, src/index.jl:168-194) when `assign` is given:
 *   cell     = assign ? assign[j] - assign_base : coarse_search(x, 1)
 *   residual = x - centroid[cell]
 *   codes    = quantize_data(rq, residual)
 */
void FN(oracle_encode)(const T* centroids, int kc, int D, int m, int ksub, const T* cb_vectors,
                       const uint8_t* cb_codes, const T* X, int64_t n, const int64_t* assign,
                       int assign_base, int32_t* cells_out, uint8_t* codes_out, int nthreads) {
    const int dsub = D / m;
    T* norms = (T*)malloc(sizeof(T) * (size_t)m * ksub);
    FN(oracle_codebook_norms)(cb_vectors, m, ksub, dsub, norms);
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        T* scratch = (T*)malloc(sizeof(T) * (size_t)kc);
        T* resid = (T*)malloc(sizeof(T) * (size_t)D);
#pragma omp for schedule(dynamic, 64)
        for (int64_t j = 0; j < n; ++j) {
            const T* x = X + j * D;
            int32_t cell;
            if (assign) {
                cell = (int32_t)(assign[j] - assign_base);
            } else {
                T dc;
                FN(coarse_one)(centroids, kc, D, x, 1, &cell, &dc, scratch);
            }
            const T* c = centroids + (size_t)cell * D;
            for (int d = 0; d < D; ++d) resid[d] = x[d] - c[d];
            FN(encode_residual)(resid, D, m, ksub, cb_vectors, cb_codes, norms, codes_out + j * m);
            cells_out[j] = cell;
        }
        free(scratch);
        free(resid);
    }
    free(norms);
}

/*
 * knn_search(ivfadc, point, k; w)                      reference src/index.jl:204-258
 * Lists are given as a CSR: list c = entries offsets[c] .. offsets[c+1]-1 of codes[.][m], ids[.]
 * in list (= scan) order.  Outputs are padded to k with id = UINT64_MAX, dist = +inf.
 * scratch_T: T[kc + D*w? ...] allocated by the caller per thread (see oracle_search).
 */
static int FN(search_one)(const T* centroids, int kc, int D, int m, int ksub, const T* cb_vectors,
                          const uint8_t* cb_codes, const int64_t* offsets, const uint8_t* codes,
                          const uint64_t* ids, const T* q, int k, int w, uint64_t* ids_out,
                          T* dists_out, int32_t* cells, T* dcs, T* scratch, T* resid, T* lut,
                          int64_t* scanned) {
    const int dsub = D / m;
    /* :219  closest_clusters, coarse_distances = coarse_search(cq, point, w) */
    FN(coarse_one)(centroids, kc, D, q, w, cells, dcs, scratch);
    int cnt = 0; /* length(neighbors) */
    for (int j = 0; j < w; ++j) { /* :228  probe-rank order */
        const int cell = cells[j];
        const T dc = dcs[j]; /* :229 */
        /* :220  residuals = point .- centroids[:, closest]   (_closest_cluster_residuals) */
        const T* c = centroids + (size_t)cell * D;
        for (int d = 0; d < D; ++d) resid[d] = q[d] - c[d];
        /* :232-236  difftables[i] = LittleDict(codes_i, colwise(Dc(), vectors_i, residuals[rr, j]))
           -- keyed by code VALUE. */
        for (int i = 0; i < m; ++i)
            for (int cw = 0; cw < ksub; ++cw)
                lut[i * 256 + cb_codes[(size_t)i * ksub + cw]] = FN(dist_colwise)(
                    g_metric_coarse, cb_vectors + ((size_t)i * ksub + cw) * dsub, resid + (size_t)i * dsub, dsub);
        /* :240-255  scan the list */
        for (int64_t p = offsets[cell]; p < offsets[cell + 1]; ++p) {
            const uint8_t* cd = codes + (size_t)p * m;
            T d = dc;                                             /* :242 */
            for (int i = 0; i < m; ++i) d = d + lut[i * 256 + cd[i]]; /* :243-246 (A3) */
            /* :247-254  SortedMultiDict used as a bounded max-heap; equal keys keep insertion
               order, the evicted element is the last of the greatest keys. */
            if (cnt < k) {
                int pos = cnt;
                while (pos > 0 && d < dists_out[pos - 1]) {
                    dists_out[pos] = dists_out[pos - 1];
                    ids_out[pos] = ids_out[pos - 1];
                    --pos;
                }
                dists_out[pos] = d;
                ids_out[pos] = ids[p];
                ++cnt;
            } else if (dists_out[k - 1] > d) { /* :250 strict */
                int pos = k - 1;
                while (pos > 0 && d < dists_out[pos - 1]) {
                    dists_out[pos] = dists_out[pos - 1];
                    ids_out[pos] = ids_out[pos - 1];
                    --pos;
                }
                dists_out[pos] = d;
                ids_out[pos] = ids[p];
            }
        }
        *scanned += offsets[cell + 1] - offsets[cell];
    }
    for (int i = cnt; i < k; ++i) {
        ids_out[i] = UINT64_MAX;
        dists_out[i] = (T)INFINITY;
    }
    return cnt;
}

/* Batch knn_search (reference src/index.jl:261-273); nthreads = 1 is the faithful analogue of
   the reference (single-threaded), > 1 mirrors its commented-out Threads.@threads (:269). */
int64_t FN(oracle_search)(const T* centroids, int kc, int D, int m, int ksub, const T* cb_vectors,
                          const uint8_t* cb_codes, const int64_t* offsets, const uint8_t* codes,
                          const uint64_t* ids, const T* Q, int64_t nq, int k, int w,
                          uint64_t* ids_out, T* dists_out, int32_t* counts_out, int nthreads) {
    if (w > kc) w = kc;
    int64_t scanned_total = 0;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1) reduction(+ : scanned_total)
    {
        T* scratch = (T*)malloc(sizeof(T) * (size_t)kc);
        T* resid = (T*)malloc(sizeof(T) * (size_t)D);
        T* lut = (T*)calloc((size_t)m * 256, sizeof(T));
        T* dcs = (T*)malloc(sizeof(T) * (size_t)w);
        int32_t* cells = (int32_t*)malloc(sizeof(int32_t) * (size_t)w);
#pragma omp for schedule(dynamic, 4)
        for (int64_t i = 0; i < nq; ++i) {
            int64_t scanned = 0;
            counts_out[i] = FN(search_one)(centroids, kc, D, m, ksub, cb_vectors, cb_codes, offsets,
                                           codes, ids, Q + i * D, k, w, ids_out + i * k,
                                           dists_out + i * k, cells, dcs, scratch, resid, lut,
                                           &scanned);
            scanned_total += scanned;
        }
        free(scratch);
        free(resid);
        free(lut);
        free(dcs);
        free(cells);
    }
    return scanned_total;
}

/*
 * _pop! reconstruction (reference src/utils.jl:58-59,71-81):
 *   centroid(cell) + concat_i codebook_i[code_i]   over rowrange(D, m, i); rows beyond m*dsub get
 *   the centroid plus an undefined residual in the reference (Vector{T}(undef, n)); here 0.
 */
void FN(oracle_decode)(const T* centroids, int D, int m, int ksub, const T* cb_vectors,
                       const uint8_t* cb_codes, int cell, const uint8_t* code, T* out) {
    const int dsub = D / m;
    const T* c = centroids + (size_t)cell * D;
    for (int d = 0; d < D; ++d) out[d] = c[d];
    for (int i = 0; i < m; ++i) {
        int col = -1;
        for (int cw = 0; cw < ksub; ++cw)
            if (cb_codes[(size_t)i * ksub + cw] == code[i]) {
                col = cw; /* codemap: code value -> column */
                break;
            }
        if (col < 0) continue;
        const T* wv = cb_vectors + ((size_t)i * ksub + col) * dsub;
        for (int d = 0; d < dsub; ++d) out[i * dsub + d] = c[i * dsub + d] + wv[d];
    }
}
