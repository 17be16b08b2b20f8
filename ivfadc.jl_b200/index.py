"""Host-side mirror of IVFADC.jl's public API over the C ABI (libivfadc_cuda).

Julia is not available in this image, so this Python module plays the role the Julia glue
package (julia/IVFADC, reviewed but not executable here) plays in production: same names, same
argument meaning, same error behaviour (AssertionError where the reference @asserts), every hot
operation a single C-ABI call.  Julia's `!` is not an identifier character in Python:

    IVFADCIndex(data; kc, k, m, ...)        -> IVFADCIndex(data, kc=, k=, m=, ...)
    knn_search(ivfadc, point(s), k; w)      -> knn_search(ivfadc, point(s), k, w=)
    push!/pushfirst!(ivfadc, point)         -> push(ivfadc, point) / pushfirst(ivfadc, point)
    pop!/popfirst!(ivfadc)                  -> pop(ivfadc) / popfirst(ivfadc)
    delete_from_index!(ivfadc, points)      -> delete_from_index(ivfadc, points)
    length(ivfadc), size(ivfadc)            -> len(ivfadc), ivfadc.size()
    save_ivfadc_index / load_ivfadc_index   -> persistency.py

`data` keeps Julia's orientation: shape (nrows, nvectors), one vector per COLUMN.
Citations are to the reference tree.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np

from . import _capi
from . import training

DEFAULT_COARSE_K = 2            # src/defaults.jl:2
DEFAULT_QUANTIZATION_K = 256    # src/defaults.jl:3
DEFAULT_QUANTIZATION_M = 1      # src/defaults.jl:4
DEFAULT_QUANTIZATION_METHOD = "pq"
DEFAULT_COARSE_DISTANCE = "SqEuclidean"
DEFAULT_COARSE_QUANTIZER = "naive"
DEFAULT_QUANTIZATION_DISTANCE = "SqEuclidean"
DEFAULT_COARSE_MAXITER = 25
DEFAULT_QUANTIZATION_MAXITER = 25

_TYPE_TO_BITS = {np.dtype(np.uint8): 8, np.dtype(np.uint16): 16, np.dtype(np.uint32): 32,
                 np.dtype(np.uint64): 64}   # QuantizedArrays.TYPE_TO_BITS (UInt128 unsupported)
_JULIA_FLOAT = {np.dtype(np.float32): "Float32", np.dtype(np.float64): "Float64"}


def _dtype_code(dt):
    dt = np.dtype(dt)
    if dt == np.float32:
        return _capi.F32
    if dt == np.float64:
        return _capi.F64
    raise TypeError(f"IVFADCIndex needs an AbstractFloat element type (Float32/Float64), got {dt}")


DEVICE_TRAINING_PAIRS = 1 << 22   # nvectors * kc from which the constructor trains on the GPU


class IVFADCIndex:
    """IVFADCIndex{U,I,Dc,Dr,T,Q} (src/index.jl:39-48) with device-resident lists."""

    # -- constructor: src/index.jl:103-165 ------------------------------------------------------
    def __init__(self, data, *, kc=DEFAULT_COARSE_K, k=DEFAULT_QUANTIZATION_K,
                 m=DEFAULT_QUANTIZATION_M, coarse_quantizer=DEFAULT_COARSE_QUANTIZER,
                 coarse_distance=DEFAULT_COARSE_DISTANCE,
                 quantization_distance=DEFAULT_QUANTIZATION_DISTANCE,
                 quantization_method=DEFAULT_QUANTIZATION_METHOD,
                 coarse_maxiter=DEFAULT_COARSE_MAXITER,
                 quantization_maxiter=DEFAULT_QUANTIZATION_MAXITER, index_type=np.uint32,
                 device=0, seed=0, shard=(0, 1), flags=0):
        data = np.asarray(data)
        _dtype_code(data.dtype)
        nrows, nvectors = data.shape
        index_type = np.dtype(index_type)
        bits_required = math.ceil(math.log2(nvectors)) if nvectors > 0 else 0
        # the reference's checks, same order and messages (src/index.jl:118-125)
        assert kc >= 2, "Number of coarse clusters has to be >= 2"
        assert k <= nvectors, f"Number of quantization levels  has to be <= {nvectors}"
        assert 1 <= m <= nrows, f"Number of codebooks has to be between 1 and {nrows}"
        assert coarse_quantizer in ("naive", "hnsw"), "Coarse quantizer can be :naive or :hnsw only"
        assert coarse_maxiter > 0, "Number of clustering iterations has to be > 0"
        assert quantization_maxiter > 0, "Number of clustering iterations has to be > 0"
        assert index_type in _TYPE_TO_BITS and _TYPE_TO_BITS[index_type] >= bits_required, \
            f"{nvectors} vectors require at least {bits_required} index bits"
        if quantization_method != "pq":
            raise NotImplementedError("only quantization_method=:pq is on the hot path (SURVEY 8f-4)")
        X = np.ascontiguousarray(data.T)  # [nvectors, nrows]: the same bytes Julia holds
        # training stays outside the engine's parity scope (Clustering.jl / QuantizedArrays.jl in production);
        # beyond toy sizes the Lloyd iterations run on the GPU, their assignment step being the engine's K1
        if nvectors * kc >= DEVICE_TRAINING_PAIRS:
            centroids, assign, cb_vectors, cb_codes = training.train_quantizers_device(
                X, kc, k, m, coarse_maxiter, quantization_maxiter, seed, device)
        else:
            centroids, assign, cb_vectors, cb_codes = training.train_quantizers(
                X, kc, k, m, coarse_maxiter, quantization_maxiter, seed)
        self._init_from_quantizers(centroids, cb_vectors, cb_codes, index_type, coarse_quantizer,
                                   coarse_distance, quantization_distance, device, shard, flags)
        # _build_residuals + _build_inverted_index (src/index.jl:168-194): k-means' own
        # assignments, ids ascending per list
        self._add(X, _capi.LAST, assign=assign, assign_base=0)

    @classmethod
    def from_quantizers(cls, centroids, cb_vectors, cb_codes=None, *, index_type=np.uint32,
                        coarse_quantizer="naive", coarse_distance="SqEuclidean",
                        quantization_distance="SqEuclidean", device=0, shard=(0, 1), flags=0):
        """An empty index around trained quantizers: centroids [kc, D], cb_vectors [m, k, dsub],
        cb_codes uint8 [m, k] (default 0..k-1).  What the Julia glue does after training."""
        self = cls.__new__(cls)
        centroids = np.ascontiguousarray(centroids)
        cb_vectors = np.ascontiguousarray(cb_vectors, dtype=centroids.dtype)
        if cb_codes is None:
            cb_codes = np.tile(np.arange(cb_vectors.shape[1], dtype=np.uint8), (cb_vectors.shape[0], 1))
        self._init_from_quantizers(centroids, cb_vectors, np.ascontiguousarray(cb_codes, dtype=np.uint8),
                                   np.dtype(index_type), coarse_quantizer, coarse_distance,
                                   quantization_distance, device, shard, flags)
        return self

    def _init_from_quantizers(self, centroids, cb_vectors, cb_codes, index_type, coarse_quantizer,
                              coarse_distance, quantization_distance, device, shard, flags=0):
        # Dc / Dr (src/index.jl:41-42,108-109): SqEuclidean runs on the tensor-core paths, Euclidean / Cityblock /
        # CosineDist on the exact generic kernels; any other Distances.jl metric is outside the engine
        for name in (str(coarse_distance), str(quantization_distance)):
            if name not in _capi.METRICS:
                raise NotImplementedError(f"distance {name}: supported are {sorted(_capi.METRICS)} (SURVEY 8f-3)")
        self.coarse_distance, self.quantization_distance = str(coarse_distance), str(quantization_distance)
        self.T = centroids.dtype
        self.I = np.dtype(index_type)
        self.coarse_quantizer = coarse_quantizer
        self.kc, self.nrows = centroids.shape
        self.m, self.k, self.dsub = cb_vectors.shape
        self.device = device
        self.shard = tuple(shard)
        self._lib = _capi.load()
        cfg = _capi.Config(dim=self.nrows, kc=self.kc, m=self.m, ksub=self.k,
                           dtype=_dtype_code(self.T), id_bytes=self.I.itemsize,
                           metric_coarse=_capi.METRICS[self.coarse_distance],
                           metric_resid=_capi.METRICS[self.quantization_distance],
                           device=device, shard_rank=shard[0], shard_world=shard[1], flags=int(flags))
        h = ctypes.c_void_p()
        rc = self._lib.ivfadc_create(ctypes.byref(h), ctypes.byref(cfg), _capi.ptr(centroids),
                                     _capi.ptr(cb_vectors), _capi.ptr(cb_codes))
        if rc != _capi.OK:
            raise _capi.IvfadcError(rc, "ivfadc_create failed (no CUDA device, or configuration "
                                        "outside the hot-path scope); there is no CPU fallback")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ivfadc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- length / size / show: src/index.jl:56-77 -------------------------------------------------
    def __len__(self):
        n = ctypes.c_int64()
        _capi.check(self._h, self._lib.ivfadc_length(self._h, ctypes.byref(n)))
        return int(n.value)

    def size(self, i=None):
        s = (self.nrows, len(self))
        return s if i is None else s[i - 1]  # size(ivfadc, i) is 1-based

    def __repr__(self):
        idxsize = self.I.itemsize
        compsize = 1
        codesize = self.m * compsize + idxsize
        cqstr = "HNSW" if self.coarse_quantizer == "hnsw" else "naive"
        return (f"IVFADCIndex, {cqstr} coarse quantizer, {codesize}-byte encoding "
                f"({idxsize} + {compsize}×{self.m}), {len(self)} {_JULIA_FLOAT[np.dtype(self.T)]} vectors")

    # -- internals ----------------------------------------------------------------------------------
    def _add(self, X, position, assign=None, assign_base=0, want_cells=False):
        X = np.ascontiguousarray(X, dtype=self.T)
        n = X.shape[0]
        a = None if assign is None else np.ascontiguousarray(assign, dtype=np.int64)
        cells = np.empty(n, dtype=np.int32) if want_cells else None
        rc = self._lib.ivfadc_add(self._h, _capi.ptr(X), n, position, _capi.ptr(a), assign_base,
                                  _capi.ptr(cells))
        _capi.check(self._h, rc)
        return cells

    def _check_point(self, point):
        point = np.asarray(point)
        if point.dtype != self.T:
            # Julia: MethodError -- the method is only defined for Vector{T} of the index's T
            raise TypeError(f"expected a vector of {_JULIA_FLOAT[np.dtype(self.T)]}, got {point.dtype}")
        return point

    def set_cell_owners(self, owners):
        """Which shard owns each cell (int32 [kc], values in [0, world)); before any vector is added."""
        owners = np.ascontiguousarray(owners, dtype=np.int32)
        assert owners.shape == (self.kc,)
        _capi.check(self._h, self._lib.ivfadc_set_cell_owners(self._h, _capi.ptr(owners)))

    def check_async(self, stream=None):
        """Synchronise `stream` and raise if a tensor-core pipeline of an asynchronous call timed out."""
        _capi.check(self._h, self._lib.ivfadc_check_async(self._h, ctypes.c_void_p(stream or 0)))

    def list_sizes(self):
        out = np.empty(self.kc, dtype=np.int64)
        _capi.check(self._h, self._lib.ivfadc_list_sizes(self._h, _capi.ptr(out)))
        return out

    def export_list(self, cell):
        """(idxs [len] in the index type, codes uint8 [len, m]) of list `cell` (0-based)."""
        n = int(self.list_sizes()[cell])
        ids = np.empty(n, dtype=np.uint64)
        codes = np.empty((n, self.m), dtype=np.uint8)
        _capi.check(self._h, self._lib.ivfadc_export_list(self._h, cell, _capi.ptr(ids), _capi.ptr(codes)))
        return ids.astype(self.I), codes

    def import_list(self, cell, ids, codes):
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        codes = np.ascontiguousarray(codes, dtype=np.uint8).reshape(len(ids), self.m)
        _capi.check(self._h, self._lib.ivfadc_import_list(self._h, cell, _capi.ptr(ids), _capi.ptr(codes),
                                                          len(ids)))

    def reserve(self, n_total, sizes=None):
        """Capacity hint before a bulk build (sizehint!): n_total vectors, or exact per-list sizes int64 [kc]."""
        s = None if sizes is None else np.ascontiguousarray(sizes, dtype=np.int64)
        _capi.check(self._h, self._lib.ivfadc_reserve(self._h, int(n_total), _capi.ptr(s)))

    def export_all(self):
        """Every list in one call (bulk persistency): (sizes int64 [kc], idxs [sum] in the index type, codes uint8
        [sum, m]); the entries of the lists follow each other in ascending cell order."""
        sizes = self.list_sizes()
        n = int(sizes.sum())
        ids = np.empty(n, dtype=np.uint64)
        codes = np.empty((n, self.m), dtype=np.uint8)
        _capi.check(self._h, self._lib.ivfadc_export_all(self._h, _capi.ptr(ids), _capi.ptr(codes)))
        return sizes, ids.astype(self.I), codes

    def import_all(self, sizes, ids, codes):
        """Replace ALL lists: sizes int64 [kc], packed ids / codes as export_all returns them."""
        sizes = np.ascontiguousarray(sizes, dtype=np.int64)
        assert sizes.shape == (self.kc,)
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        codes = np.ascontiguousarray(codes, dtype=np.uint8).reshape(len(ids), self.m)
        assert int(sizes.sum()) == len(ids)
        _capi.check(self._h, self._lib.ivfadc_import_all(self._h, _capi.ptr(sizes), _capi.ptr(ids), _capi.ptr(codes)))

    def add_device(self, dX_ptr: int, n: int, position=_capi.LAST, d_assign_ptr: int = 0, assign_base: int = 0,
                   d_cells_out_ptr: int = 0):
        """ivfadc_add_device: the batch (and optional int64 assignments / int32 cells out) are device pointers."""
        _capi.check(self._h, self._lib.ivfadc_add_device(self._h, ctypes.c_void_p(dX_ptr), n, position,
                                                         ctypes.c_void_p(d_assign_ptr or None), assign_base,
                                                         ctypes.c_void_p(d_cells_out_ptr or None)))

    def quantizers(self):
        c = np.empty((self.kc, self.nrows), dtype=self.T)
        v = np.empty((self.m, self.k, self.dsub), dtype=self.T)
        codes = np.empty((self.m, self.k), dtype=np.uint8)
        _capi.check(self._h, self._lib.ivfadc_export_quantizers(self._h, _capi.ptr(c), _capi.ptr(v),
                                                                _capi.ptr(codes)))
        return c, v, codes

    def encode(self, X, assign=None, assign_base=0):
        """Parity hook: (cells int32 [n] 0-based, codes uint8 [n, m]) of X [n, D] without mutation."""
        X = np.ascontiguousarray(X, dtype=self.T).reshape(-1, self.nrows)
        n = X.shape[0]
        cells = np.empty(n, dtype=np.int32)
        codes = np.empty((n, self.m), dtype=np.uint8)
        a = None if assign is None else np.ascontiguousarray(assign, dtype=np.int64)
        _capi.check(self._h, self._lib.ivfadc_encode(self._h, _capi.ptr(X), n, _capi.ptr(a), assign_base,
                                                     _capi.ptr(cells), _capi.ptr(codes)))
        return cells, codes

    def coarse_search(self, Q, w):
        """Parity hook for coarse_search (src/coarsequantizers.jl:33-37): 0-based cells, distances."""
        Q = np.ascontiguousarray(Q, dtype=self.T).reshape(-1, self.nrows)
        w = min(w, self.kc)
        cells = np.empty((Q.shape[0], w), dtype=np.int32)
        dc = np.empty((Q.shape[0], w), dtype=self.T)
        _capi.check(self._h, self._lib.ivfadc_coarse_search(self._h, _capi.ptr(Q), Q.shape[0], w,
                                                            _capi.ptr(cells), _capi.ptr(dc)))
        return cells, dc

    def search_packed(self, Q, k, w=1):
        """One C-ABI call on a packed [nq, D] host matrix: (ids uint64 [nq,k], dists T [nq,k],
        counts int32 [nq]); rows are padded with id = 2^64-1, dist = +inf beyond counts[i]."""
        assert k >= 1, "Number of neighbors must be k >= 1"                       # src/index.jl:210
        assert w >= 1, "Number of clusters to search in must be w >= 1"           # src/index.jl:211
        Q = np.ascontiguousarray(Q, dtype=self.T).reshape(-1, self.nrows)
        nq = Q.shape[0]
        ids = np.empty((nq, k), dtype=np.uint64)
        dists = np.empty((nq, k), dtype=self.T)
        counts = np.empty(nq, dtype=np.int32)
        _capi.check(self._h, self._lib.ivfadc_search(self._h, _capi.ptr(Q), nq, k, w, _capi.ptr(ids),
                                                     _capi.ptr(dists), _capi.ptr(counts)))
        return ids, dists, counts

    def stats(self, reset=False):
        st = _capi.Stats()
        _capi.check(self._h, self._lib.ivfadc_get_stats(self._h, ctypes.byref(st)))
        if reset:
            _capi.check(self._h, self._lib.ivfadc_reset_stats(self._h))
        return st.as_dict()


# ---- knn_search: src/index.jl:204-273 -----------------------------------------------------------
def knn_search(ivfadc: IVFADCIndex, point, k: int, w: int = 1):
    """Single query (1-D array) -> (idxs: ndarray of the index type, 0-based; dists: ndarray of T,
    ascending), at most k of them.  Batch (a list/tuple of 1-D arrays, Julia's Vector{Vector{T}})
    -> (list of idxs arrays, list of dists arrays)."""
    assert k >= 1, "Number of neighbors must be k >= 1"
    assert w >= 1, "Number of clusters to search in must be w >= 1"
    if isinstance(point, (list, tuple)):
        pts = [ivfadc._check_point(p) for p in point]
        for p in pts:
            assert p.shape == (ivfadc.nrows,)
        Q = np.stack(pts) if pts else np.empty((0, ivfadc.nrows), dtype=ivfadc.T)
        ids, dists, counts = ivfadc.search_packed(Q, k, w)
        return ([ids[i, :counts[i]].astype(ivfadc.I) for i in range(len(pts))],
                [dists[i, :counts[i]].copy() for i in range(len(pts))])
    point = ivfadc._check_point(point)
    ids, dists, counts = ivfadc.search_packed(point[None, :], k, w)
    return ids[0, :counts[0]].astype(ivfadc.I), dists[0, :counts[0]].copy()


# ---- push! / pushfirst!: src/utils.jl:114-145 ----------------------------------------------------
def _push(ivfadc: IVFADCIndex, point, position):
    nrows, nvectors = ivfadc.size()
    point = np.asarray(point)
    assert nrows == point.shape[0], f"Adding to index requires {nrows}-element vectors"   # :133
    assert _TYPE_TO_BITS[ivfadc.I] >= math.log2(nvectors + 1), \
        f"Cannot index, exceeding index capacity of {2 ** _TYPE_TO_BITS[ivfadc.I]} points"  # :134-135
    point = ivfadc._check_point(point)
    ivfadc._add(point[None, :], position)
    return None


def push(ivfadc: IVFADCIndex, point):
    return _push(ivfadc, point, _capi.LAST)


def pushfirst(ivfadc: IVFADCIndex, point):
    return _push(ivfadc, point, _capi.FIRST)


def push_batch(ivfadc: IVFADCIndex, X, first=False):
    """Extension (the reference has no batch method): X [n, D]; equivalent to n
    successive push! (or pushfirst!) calls in row order, one C-ABI call."""
    X = np.ascontiguousarray(X)
    assert X.ndim == 2 and X.shape[1] == ivfadc.nrows, f"Adding to index requires {ivfadc.nrows}-element vectors"
    assert _TYPE_TO_BITS[ivfadc.I] >= math.log2(len(ivfadc) + X.shape[0]), "Cannot index, exceeding index capacity"
    ivfadc._add(ivfadc._check_point(X), _capi.FIRST if first else _capi.LAST)


# ---- pop! / popfirst!: src/utils.jl:29-68 --------------------------------------------------------
def _pop(ivfadc: IVFADCIndex, position):
    assert len(ivfadc) > 0, "Cannot pop element from empty index"   # :44
    out = np.empty(ivfadc.nrows, dtype=ivfadc.T)
    found = ctypes.c_int32()
    _capi.check(ivfadc._h, ivfadc._lib.ivfadc_pop(ivfadc._h, position, _capi.ptr(out), ctypes.byref(found)))
    return out


def pop(ivfadc: IVFADCIndex):
    return _pop(ivfadc, _capi.LAST)


def popfirst(ivfadc: IVFADCIndex):
    return _pop(ivfadc, _capi.FIRST)


# ---- delete_from_index!: src/utils.jl:90-105 -----------------------------------------------------
def delete_from_index(ivfadc: IVFADCIndex, points):
    """`points` are 1-based integers like in the reference."""
    maxid = 2 ** _TYPE_TO_BITS[ivfadc.I] - 1
    arr = np.asarray(points)
    if arr.ndim == 1 and arr.dtype.kind in "iu" and arr.dtype.itemsize <= 8 and arr.dtype != np.uint64:
        # integer arrays: the same checks, vectorised (a 1 M-id delete is one C-ABI call, not 1 M Python steps)
        v = arr.astype(np.int64) - 1
        bad = (v < 0) | (v > maxid) if maxid < 2 ** 63 else (v < 0)
        if bad.any():
            raise OverflowError(f"InexactError: cannot convert {int(v[np.argmax(bad)])} to {ivfadc.I}")  # I.(points .- 1), :93
        ids = v.astype(np.uint64)
        if ids.size == 0:
            return None
        ids = np.ascontiguousarray(ids)
        _capi.check(ivfadc._h, ivfadc._lib.ivfadc_delete(ivfadc._h, _capi.ptr(ids), len(ids)))
        return None
    shifted = []
    for p in points:
        v = int(p) - 1
        if v < 0 or v > maxid:
            raise OverflowError(f"InexactError: cannot convert {v} to {ivfadc.I}")  # I.(points .- 1), :93
        shifted.append(v)
    if not shifted:
        return None
    ids = np.asarray(shifted, dtype=np.uint64)
    _capi.check(ivfadc._h, ivfadc._lib.ivfadc_delete(ivfadc._h, _capi.ptr(ids), len(ids)))
    return None
