"""
TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

Python face of the CPU oracle: ctypes bindings of oracle/liboracle.so (the C restatement of the
numeric hot path) plus `OracleIndex`, a literal restatement of the reference's index container and
its mutation methods (Python lists standing in for Julia's Vector{Vector{UInt8}}).

PARITY UNPINNED for numeric values (Julia and the reference's dependencies are not available in
this image; see oracle/ivfadc_oracle.c).  Pinned against the reference's own tests:
test/search.jl:26-49 and test/utils.jl:58-105 (tests/test_oracle_reference_tests.py).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
Reference citations are relative to the reference tree (src/..., test/...).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_c_i64 = ctypes.c_int64
_c_int = ctypes.c_int
_vp = ctypes.c_void_p


def build(force: bool = False) -> str:
    """Compile liboracle.so with oracle/Makefile (gcc)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("ivfadc_oracle.c", "ivfadc_oracle_impl.h", "Makefile")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _sfx(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(f"unsupported element type {dtype}")


def _p(a):
    return a.ctypes.data_as(_vp) if a is not None else None


class Quantizers:
    """The trained inputs shared by the oracle and the CUDA engine.

    centroids  T[kc][D]           (Julia cq.vectors is D x kc column-major: same bytes)
    cb_vectors T[m][ksub][dsub]   (Julia codebooks[i].vectors is dsub x ksub column-major)
    cb_codes   uint8[m][ksub]     (Julia codebooks[i].codes)
    """

    def __init__(self, centroids, cb_vectors, cb_codes=None, coarse_distance="SqEuclidean",
                 quantization_distance="SqEuclidean"):
        self.coarse_distance, self.quantization_distance = coarse_distance, quantization_distance
        self.centroids = np.ascontiguousarray(centroids)
        self.cb_vectors = np.ascontiguousarray(cb_vectors, dtype=self.centroids.dtype)
        self.kc, self.D = self.centroids.shape
        self.m, self.ksub, self.dsub = self.cb_vectors.shape
        assert self.dsub == self.D // self.m
        if cb_codes is None:
            cb_codes = np.tile(np.arange(self.ksub, dtype=np.uint8), (self.m, 1))
        self.cb_codes = np.ascontiguousarray(cb_codes, dtype=np.uint8)
        self.dtype = self.centroids.dtype


METRICS = {"SqEuclidean": 0, "Euclidean": 1, "Cityblock": 2, "CosineDist": 3}


def _set_metrics(qz):
    """Dc / Dr of the next oracle call (the reference's type parameters, src/index.jl:41-42)."""
    lib().oracle_set_metrics(_c_int(METRICS[getattr(qz, "coarse_distance", "SqEuclidean")]),
                             _c_int(METRICS[getattr(qz, "quantization_distance", "SqEuclidean")]))


def coarse_search(qz: Quantizers, Q, w: int, nthreads: int = 1):
    """coarse_search, src/coarsequantizers.jl:33-37.  Returns 0-based cells [nq, w] and dc [nq, w]."""
    Q = np.ascontiguousarray(Q, dtype=qz.dtype).reshape(-1, qz.D)
    nq = Q.shape[0]
    w = min(w, qz.kc)
    cells = np.empty((nq, w), dtype=np.int32)
    dc = np.empty((nq, w), dtype=qz.dtype)
    _set_metrics(qz)
    fn = getattr(lib(), "oracle_coarse_search_" + _sfx(qz.dtype))
    fn.restype = None
    fn(_p(qz.centroids), _c_int(qz.kc), _c_int(qz.D), _p(Q), _c_i64(nq), _c_int(w), _p(cells),
       _p(dc), _c_int(nthreads))
    return cells, dc


def encode(qz: Quantizers, X, assign=None, assign_base: int = 0, nthreads: int = 1):
    """_encode_point (src/utils.jl:148-161) / build path (src/index.jl:168-194).
    Returns 0-based cells int32[n] and codes uint8[n, m]."""
    X = np.ascontiguousarray(X, dtype=qz.dtype).reshape(-1, qz.D)
    n = X.shape[0]
    cells = np.empty(n, dtype=np.int32)
    codes = np.empty((n, qz.m), dtype=np.uint8)
    a = None if assign is None else np.ascontiguousarray(assign, dtype=np.int64)
    _set_metrics(qz)
    fn = getattr(lib(), "oracle_encode_" + _sfx(qz.dtype))
    fn.restype = None
    fn(_p(qz.centroids), _c_int(qz.kc), _c_int(qz.D), _c_int(qz.m), _c_int(qz.ksub),
       _p(qz.cb_vectors), _p(qz.cb_codes), _p(X), _c_i64(n), _p(a), _c_int(assign_base), _p(cells),
       _p(codes), _c_int(nthreads))
    return cells, codes


def search_csr(qz: Quantizers, offsets, codes, ids, Q, k: int, w: int, nthreads: int = 1):
    """Batch knn_search (src/index.jl:204-273) over a CSR copy of the lists.
    Returns ids uint64[nq,k], dists T[nq,k], counts int32[nq], scanned_vectors."""
    Q = np.ascontiguousarray(Q, dtype=qz.dtype).reshape(-1, qz.D)
    nq = Q.shape[0]
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    ids = np.ascontiguousarray(ids, dtype=np.uint64)
    ids_out = np.empty((nq, k), dtype=np.uint64)
    dists_out = np.empty((nq, k), dtype=qz.dtype)
    counts = np.empty(nq, dtype=np.int32)
    _set_metrics(qz)
    fn = getattr(lib(), "oracle_search_" + _sfx(qz.dtype))
    fn.restype = _c_i64
    scanned = fn(_p(qz.centroids), _c_int(qz.kc), _c_int(qz.D), _c_int(qz.m), _c_int(qz.ksub),
                 _p(qz.cb_vectors), _p(qz.cb_codes), _p(offsets), _p(codes), _p(ids), _p(Q),
                 _c_i64(nq), _c_int(k), _c_int(w), _p(ids_out), _p(dists_out), _p(counts),
                 _c_int(nthreads))
    return ids_out, dists_out, counts, int(scanned)


def decode(qz: Quantizers, cell: int, code):
    """centroid + _decode_point, src/utils.jl:58-59,71-81."""
    out = np.empty(qz.D, dtype=qz.dtype)
    code = np.ascontiguousarray(code, dtype=np.uint8)
    fn = getattr(lib(), "oracle_decode_" + _sfx(qz.dtype))
    fn.restype = None
    fn(_p(qz.centroids), _c_int(qz.D), _c_int(qz.m), _c_int(qz.ksub), _p(qz.cb_vectors),
       _p(qz.cb_codes), _c_int(int(cell)), _p(code), _p(out))
    return out


_ID_BITS = {1: 8, 2: 16, 4: 32, 8: 64, 16: 128}  # QuantizedArrays.TYPE_TO_BITS by sizeof


class OracleIndex:
    """Literal restatement of IVFADCIndex + src/utils.jl on Python lists (small cases only).

    inverse_index[c] = (idxs: list[int], codes: list[np.ndarray(m, uint8)])   src/index.jl:8-11
    """

    def __init__(self, qz: Quantizers, id_bytes: int = 4):
        self.qz = qz
        self.id_bytes = id_bytes
        self.lists = [([], []) for _ in range(qz.kc)]

    # -- src/index.jl:56-66 ---------------------------------------------------------------
    def __len__(self):
        return sum(len(idxs) for idxs, _ in self.lists)

    def size(self):
        return (self.qz.D, len(self))

    # -- src/index.jl:178-194 : _build_inverted_index --------------------------------------
    def build(self, data, assignments, assign_base: int = 1):
        data = np.ascontiguousarray(data, dtype=self.qz.dtype)
        n = data.shape[0]
        bits_required = int(np.ceil(np.log2(n))) if n > 1 else 0
        assert _ID_BITS[self.id_bytes] >= bits_required  # src/index.jl:124-125
        cells, codes = encode(self.qz, data, assign=assignments, assign_base=assign_base)
        for c in range(self.qz.kc):
            sel = np.flatnonzero(cells == c)  # findall(isequal(cluster), assignments): ascending
            self.lists[c] = ([int(i) for i in sel], [codes[i].copy() for i in sel])
        return self

    # -- src/utils.jl:127-145 : _push! -----------------------------------------------------
    def _push(self, point, position):
        nrows, nvectors = self.size()
        point = np.asarray(point, dtype=self.qz.dtype)
        assert nrows == point.shape[0], f"Adding to index requires {nrows}-element vectors"
        assert _ID_BITS[self.id_bytes] >= np.log2(nvectors + 1), "Cannot index, exceeding index capacity"
        cells, codes = encode(self.qz, point[None, :])  # _encode_point
        vecid, shift = (0, 1) if position == "first" else (nvectors, 0)
        for idxs, _ in self.lists:  # _shift_up_inverse_index!
            for i in range(len(idxs)):
                idxs[i] += shift
        idxs, cds = self.lists[int(cells[0])]
        idxs.append(vecid)
        cds.append(codes[0].copy())

    def push(self, point):
        self._push(point, "last")

    def pushfirst(self, point):
        self._push(point, "first")

    # -- src/utils.jl:41-68 : _pop! --------------------------------------------------------
    def _pop(self, position):
        assert len(self) > 0, "Cannot pop element from empty index"
        vecid, shift = (len(self) - 1, 0) if position == "last" else (0, 1)
        cluster, idx = None, None
        for c, (idxs, _) in enumerate(self.lists):
            if vecid in idxs:
                cluster, idx = c, idxs.index(vecid)
        idxs, cds = self.lists[cluster]
        rec = decode(self.qz, cluster, cds[idx])
        del idxs[idx]
        del cds[idx]
        for idxs, _ in self.lists:  # _shift_down_inverse_index!
            for i in range(len(idxs)):
                idxs[i] -= shift
        return rec

    def pop(self):
        return self._pop("last")

    def popfirst(self):
        return self._pop("first")

    # -- src/utils.jl:90-105 : delete_from_index! (points are 1-based) ----------------------
    def delete_from_index(self, points):
        shifted = []
        for p in points:
            v = int(p) - 1
            if v < 0 or v >= 2 ** _ID_BITS[self.id_bytes]:
                raise OverflowError("InexactError")  # I.(points .- 1)
            shifted.append(v)
        for point in sorted(set(shifted), reverse=True):
            for idxs, cds in self.lists:
                if point in idxs:
                    pidx = idxs.index(point)
                    del idxs[pidx]
                    del cds[pidx]
                    for idxs2, _ in self.lists:  # _shift_inverse_index!
                        for i in range(len(idxs2)):
                            if idxs2[i] > point:
                                idxs2[i] -= 1
                    break

    # -- CSR view + search -------------------------------------------------------------------
    def csr(self):
        lens = np.array([len(idxs) for idxs, _ in self.lists], dtype=np.int64)
        offsets = np.zeros(self.qz.kc + 1, dtype=np.int64)
        np.cumsum(lens, out=offsets[1:])
        n = int(offsets[-1])
        ids = np.empty(n, dtype=np.uint64)
        codes = np.empty((n, self.qz.m), dtype=np.uint8)
        for c, (idxs, cds) in enumerate(self.lists):
            o = int(offsets[c])
            for j in range(len(idxs)):
                ids[o + j] = idxs[j]
                codes[o + j] = cds[j]
        return offsets, codes, ids

    def knn_search(self, Q, k: int, w: int = 1, nthreads: int = 1):
        assert k >= 1, "Number of neighbors must be k >= 1"  # src/index.jl:210
        assert w >= 1, "Number of clusters to search in must be w >= 1"  # src/index.jl:211
        offsets, codes, ids = self.csr()
        return search_csr(self.qz, offsets, codes, ids, Q, k, w, nthreads)[:3]


def truth_search_fp64(qz: Quantizers, offsets, codes, ids, q, k: int, w: int):
    """fp64 'truth' of one knn_search (same algorithm, float64 numpy, exact-ish arithmetic) used
    to budget the rounding error of both the oracle and the GPU path."""
    C = qz.centroids.astype(np.float64)
    cb = qz.cb_vectors.astype(np.float64)
    q = np.asarray(q, dtype=np.float64)
    d = ((C - q[None, :]) ** 2).sum(1)
    order = np.argsort(d, kind="stable")[: min(w, qz.kc)]
    cand = []
    for rank, cell in enumerate(order):
        r = q - C[cell]
        lut = np.zeros((qz.m, 256))
        for i in range(qz.m):
            diff = cb[i] - r[i * qz.dsub:(i + 1) * qz.dsub][None, :]
            lut[i, qz.cb_codes[i]] = (diff ** 2).sum(1)
        lo, hi = int(offsets[cell]), int(offsets[cell + 1])
        if hi > lo:
            cd = codes[lo:hi]
            dist = d[cell] + lut[np.arange(qz.m)[None, :], cd].sum(1)
            for j in range(hi - lo):
                cand.append((dist[j], rank, j, int(ids[lo + j])))
    cand.sort(key=lambda t: (t[0], t[1], t[2]))
    cand = cand[:k]
    return np.array([c[3] for c in cand], dtype=np.uint64), np.array([c[0] for c in cand])


def compare_search(gi, gd, gc, oi, od, oc, rtol: float = 1e-5):
    """Parity of a search result (gi ids, gd dists, gc counts) against the oracle's (oi, od, oc) at
    the bar BASELINE.json's north_star states: counts equal, ADC distances within `rtol` relative
    rank by rank, neighbour ids equal except at near-ties (two candidates whose distances agree
    within rtol may swap ranks, or swap across the k-th boundary).  Returns a dict with the
    number of id mismatches (all verified to be near-ties), the largest relative distance error,
    and whether everything is bit-identical; raises AssertionError on a real difference."""
    gi, oi = np.asarray(gi).astype(np.uint64), np.asarray(oi).astype(np.uint64)
    gd, od = np.asarray(gd), np.asarray(od)
    gc, oc = np.asarray(gc), np.asarray(oc)
    assert np.array_equal(gc, oc), "result counts differ"
    near_ties, max_rel = 0, 0.0
    valid = np.arange(gd.shape[1])[None, :] < gc[:, None]
    denom = np.where(valid, np.maximum(np.abs(od), np.finfo(od.dtype).tiny), 1.0)
    rel = np.where(valid, np.abs(gd.astype(np.float64) - od.astype(np.float64)) / denom, 0.0)
    max_rel = float(rel.max()) if rel.size else 0.0
    assert max_rel <= rtol, f"ADC distance off by {max_rel:.3e} relative (bar {rtol:g})"
    for q in np.flatnonzero(((gi != oi) & valid).any(axis=1)):
        n = int(gc[q])
        for j in np.flatnonzero(gi[q, :n] != oi[q, :n]):
            pos = np.flatnonzero(oi[q, :n] == gi[q, j])
            ref = od[q, pos[0]] if len(pos) else od[q, n - 1]   # moved rank | crossed the k-th boundary
            assert abs(float(ref) - float(gd[q, j])) <= rtol * abs(float(ref)), \
                f"query {q} rank {j}: id {gi[q, j]} is not a near-tie of the oracle's result"
            near_ties += 1
    bit_equal = bool(np.array_equal(gi[valid], oi[valid]) and
                     np.array_equal(gd[valid].view(np.uint8), od[valid].view(np.uint8)))
    return {"near_tie_id_mismatches": near_ties, "results": int(valid.sum()), "max_rel_err": max_rel,
            "bit_identical": bit_equal}


# ---- synthetic inputs: CPU twin of ivfadc.jl_b200/csrc/synth.cu (counter-based, bit-identical) ------------------
def philox4x32_10(ctr, key):
    c = np.ascontiguousarray(ctr, dtype=np.uint32)
    k = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.empty(4, dtype=np.uint32)
    lib().oracle_philox4x32_10(_p(c), _p(k), _p(out))
    return out


def synth_uniform(first: int, n: int, D: int, seed: int):
    X = np.empty((n, D), dtype=np.float32)
    lib().oracle_synth_uniform_f32(_p(X), _c_i64(first), _c_i64(n), _c_int(D), ctypes.c_uint64(seed))
    return X


def synth_blobs(first: int, n: int, D: int, n_blobs: int, seed: int, scale, centres, want_x: bool = True):
    """(X [n, D] or None, blob ids int32 [n]) of vectors first .. first + n - 1; scale = sigma sqrt(3) / 2^22 (float32)."""
    centres = np.ascontiguousarray(centres, dtype=np.float32)
    assert centres.shape == (n_blobs, D)
    X = np.empty((n, D), dtype=np.float32) if want_x else None
    b = np.empty(n, dtype=np.int32)
    lib().oracle_synth_blobs_f32(_p(X), _p(b), _c_i64(first), _c_i64(n), _c_int(D), _c_int(n_blobs),
                                 ctypes.c_uint64(seed), ctypes.c_float(float(scale)), _p(centres))
    return X, b
