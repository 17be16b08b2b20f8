"""Cell-sharded multi-GPU search: one process per GPU, `torch.distributed` for the plumbing.

The inverted lists shard by cell (cell c lives on rank c % world; the quantizers are replicated),
lists are independent, and a query's answer is the top-k over the union of its probed lists, so
the path has one exchange step for the results (and a small one for the probe lists, so that the
coarse assignment is computed once across the ranks instead of once per rank): every rank scans the probed cells it owns and produces its
k best candidates per query with their merge keys (probe rank << 32 | position); one all-gather
(NCCL over NVLink; 8 B + 8 B + 4..8 B per candidate, ~1 MB per rank for a 10k-query batch) brings
the candidate sets together and a merge kernel keeps the k smallest by (distance, key) -- the
reference's own order (src/index.jl:247-257), so the result is bit-identical to one GPU.

torch tensors are used for device memory and the collective only; ids/keys travel as int64
(bit patterns of the uint64 values).  The engine calls are the C ABI's *_device entry points.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _capi


def _tdtype(engine):
    return torch.float32 if np.dtype(engine.T) == np.float32 else torch.float64


def _stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def search_device(engine, dQ: torch.Tensor, k: int, w: int = 1, out=None):
    """ivfadc_search_device on torch's current stream.  dQ: [nq, D] device tensor of the index's T.
    Returns (ids int64 [nq,k], dists [nq,k], counts int32 [nq]) device tensors (asynchronous)."""
    assert dQ.is_cuda and dQ.is_contiguous() and dQ.dtype == _tdtype(engine)
    nq = dQ.shape[0]
    if out is None:
        out = (torch.empty((nq, k), dtype=torch.int64, device=dQ.device),
               torch.empty((nq, k), dtype=dQ.dtype, device=dQ.device),
               torch.empty((nq,), dtype=torch.int32, device=dQ.device))
    ids, dists, counts = out
    rc = engine._lib.ivfadc_search_device(engine._h, ctypes.c_void_p(dQ.data_ptr()), nq, k, w,
                                          ctypes.c_void_p(ids.data_ptr()), ctypes.c_void_p(dists.data_ptr()),
                                          ctypes.c_void_p(counts.data_ptr()), _stream_ptr())
    _capi.check(engine._h, rc)
    return ids, dists, counts


def search_local(engine, dQ: torch.Tensor, k: int, w: int = 1):
    """Step 1 on one shard: (ids, dists, keys, counts) device tensors, rows padded with key = -1."""
    assert dQ.is_cuda and dQ.is_contiguous() and dQ.dtype == _tdtype(engine)
    nq = dQ.shape[0]
    ids = torch.empty((nq, k), dtype=torch.int64, device=dQ.device)
    keys = torch.empty((nq, k), dtype=torch.int64, device=dQ.device)
    dists = torch.empty((nq, k), dtype=dQ.dtype, device=dQ.device)
    counts = torch.empty((nq,), dtype=torch.int32, device=dQ.device)
    rc = engine._lib.ivfadc_search_local_device(
        engine._h, ctypes.c_void_p(dQ.data_ptr()), nq, k, w, ctypes.c_void_p(ids.data_ptr()),
        ctypes.c_void_p(dists.data_ptr()), ctypes.c_void_p(keys.data_ptr()),
        ctypes.c_void_p(counts.data_ptr()), _stream_ptr())
    _capi.check(engine._h, rc)
    return ids, dists, keys, counts


def coarse_device(engine, dQ: torch.Tensor, w: int):
    """ivfadc_coarse_search_device: (cells int32 [nq, w], dc [nq, w]) device tensors."""
    assert dQ.is_cuda and dQ.is_contiguous() and dQ.dtype == _tdtype(engine)
    nq = dQ.shape[0]
    cells = torch.empty((nq, w), dtype=torch.int32, device=dQ.device)
    dc = torch.empty((nq, w), dtype=dQ.dtype, device=dQ.device)
    rc = engine._lib.ivfadc_coarse_search_device(engine._h, ctypes.c_void_p(dQ.data_ptr()), nq, w,
                                                 ctypes.c_void_p(cells.data_ptr()), ctypes.c_void_p(dc.data_ptr()),
                                                 _stream_ptr())
    _capi.check(engine._h, rc)
    return cells, dc


def search_local_probes(engine, dQ: torch.Tensor, k: int, w: int, cells: torch.Tensor, dc: torch.Tensor):
    """Step 1 with caller-supplied probe lists (the coarse step was sharded by query)."""
    assert dQ.is_cuda and dQ.is_contiguous() and cells.is_contiguous() and dc.is_contiguous()
    nq = dQ.shape[0]
    ids = torch.empty((nq, k), dtype=torch.int64, device=dQ.device)
    keys = torch.empty((nq, k), dtype=torch.int64, device=dQ.device)
    dists = torch.empty((nq, k), dtype=dQ.dtype, device=dQ.device)
    counts = torch.empty((nq,), dtype=torch.int32, device=dQ.device)
    rc = engine._lib.ivfadc_search_probes_local_device(
        engine._h, ctypes.c_void_p(dQ.data_ptr()), nq, k, w, ctypes.c_void_p(cells.data_ptr()),
        ctypes.c_void_p(dc.data_ptr()), ctypes.c_void_p(ids.data_ptr()), ctypes.c_void_p(dists.data_ptr()),
        ctypes.c_void_p(keys.data_ptr()), ctypes.c_void_p(counts.data_ptr()), _stream_ptr())
    _capi.check(engine._h, rc)
    return ids, dists, keys, counts


def merge_gathered(engine, ids_all, dists_all, keys_all, k: int):
    """Step 2: [parts, nq, k] gathered candidates -> final (ids, dists, counts)."""
    parts, nq, _ = ids_all.shape
    ids = torch.empty((nq, k), dtype=torch.int64, device=ids_all.device)
    dists = torch.empty((nq, k), dtype=dists_all.dtype, device=ids_all.device)
    counts = torch.empty((nq,), dtype=torch.int32, device=ids_all.device)
    rc = engine._lib.ivfadc_merge_device(
        engine._h, parts, nq, k, ctypes.c_void_p(ids_all.data_ptr()), ctypes.c_void_p(dists_all.data_ptr()),
        ctypes.c_void_p(keys_all.data_ptr()), ctypes.c_void_p(ids.data_ptr()),
        ctypes.c_void_p(dists.data_ptr()), ctypes.c_void_p(counts.data_ptr()), _stream_ptr())
    _capi.check(engine._h, rc)
    return ids, dists, counts


def merge_parts(engine, parts, k: int):
    """Convenience for candidate sets that already live on one device (tests, single process)."""
    ids_all = torch.stack([p[0] for p in parts]).contiguous()
    dists_all = torch.stack([p[1] for p in parts]).contiguous()
    keys_all = torch.stack([p[2] for p in parts]).contiguous()
    return merge_gathered(engine, ids_all, dists_all, keys_all, k)


# ---- the sharded step INSIDE the library (csrc/shard.cu): NCCL communicator owned by the handle ----------------
def balanced_owners(list_sizes, world: int) -> np.ndarray:
    """Cell -> shard by greedy bin-packing on the list lengths (longest list first onto the lightest shard):
    every GPU scans about the same number of code bytes.  Ties keep cell order, so every rank computes the same map."""
    sizes = np.asarray(list_sizes, dtype=np.int64)
    owners = np.empty(len(sizes), dtype=np.int32)
    load = np.zeros(world, dtype=np.int64)
    count = np.zeros(world, dtype=np.int64)
    for c in np.argsort(-sizes, kind="stable"):
        r = int(np.lexsort((count, load))[0])   # lightest shard, then fewest cells
        owners[c] = r
        load[r] += sizes[c]
        count[r] += 1
    return owners


def init_comm(index, group=None):
    """Rank 0 draws the NCCL unique id, torch.distributed carries its 128 bytes (plumbing), every rank joins:
    afterwards `search_sharded*` run without Python in the loop."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    buf = (ctypes.c_ubyte * _capi.NCCL_ID_BYTES)()
    if rank == 0:
        rc = index._lib.ivfadc_nccl_unique_id(buf)
        if rc != 0:
            raise _capi.IvfadcError(rc, "ivfadc_nccl_unique_id failed (libnccl.so.2 not loadable?)")
    box = [bytes(buf)]
    dist.broadcast_object_list(box, src=0, group=group)
    idb = (ctypes.c_ubyte * _capi.NCCL_ID_BYTES).from_buffer_copy(box[0])
    _capi.check(index._h, index._lib.ivfadc_comm_init_rank(index._h, idb, world, rank))


def search_sharded_device(index, dQ: torch.Tensor, k: int, w: int = 1, out=None):
    """ivfadc_search_sharded_device on torch's current stream (collective: every rank, same batch)."""
    assert dQ.is_cuda and dQ.is_contiguous() and dQ.dtype == _tdtype(index)
    nq = dQ.shape[0]
    if out is None:
        out = (torch.empty((nq, k), dtype=torch.int64, device=dQ.device),
               torch.empty((nq, k), dtype=dQ.dtype, device=dQ.device),
               torch.empty((nq,), dtype=torch.int32, device=dQ.device))
    ids, dists, counts = out
    rc = index._lib.ivfadc_search_sharded_device(index._h, ctypes.c_void_p(dQ.data_ptr()), nq, k, w,
                                                 ctypes.c_void_p(ids.data_ptr()), ctypes.c_void_p(dists.data_ptr()),
                                                 ctypes.c_void_p(counts.data_ptr()), _stream_ptr())
    _capi.check(index._h, rc)
    return ids, dists, counts


def search_sharded_host(index, Q: np.ndarray, k: int, w: int = 1, out=None):
    """ivfadc_search_sharded: host buffers in and out (this rank uploads only its slice of Q)."""
    Q = np.ascontiguousarray(Q, dtype=index.T)
    nq = Q.shape[0]
    if out is None:
        out = (np.empty((nq, k), dtype=np.uint64), np.empty((nq, k), dtype=index.T), np.empty(nq, dtype=np.int32))
    ids, dists, counts = out
    rc = index._lib.ivfadc_search_sharded(index._h, _capi.ptr(Q), nq, k, w, _capi.ptr(ids), _capi.ptr(dists),
                                          _capi.ptr(counts))
    _capi.check(index._h, rc)
    return ids, dists, counts


class CudaShardEngine:
    """Adapter: the two device entry points the distributed searcher needs."""

    def __init__(self, index):
        self.index = index

    def search_local(self, dQ, k, w):
        return search_local(self.index, dQ, k, w)[:3]

    def coarse(self, dQ, w):
        return coarse_device(self.index, dQ, w)

    def search_local_probes(self, dQ, k, w, cells, dc):
        return search_local_probes(self.index, dQ, k, w, cells, dc)[:3]

    @property
    def kc(self):
        return self.index.kc

    def set_timing(self, enable: bool):
        _capi.check(self.index._h, self.index._lib.ivfadc_set_stats_timing(self.index._h, 1 if enable else 0))

    def merge(self, ids_all, dists_all, keys_all, k):
        return merge_gathered(self.index, ids_all, dists_all, keys_all, k)


class ShardedSearcher:
    """Cell-sharded knn_search across the ranks of the default process group.

    `engine` provides search_local(Q, k, w) -> (ids, dists, keys) [nq, k] tensors and
    merge(ids_all, dists_all, keys_all, k) -> (ids, dists, counts).  In production that is
    CudaShardEngine (NCCL backend); the CPU test suite drives the same class over gloo with an
    oracle-backed engine to cover the partition / gather / merge-order logic.
    """

    def __init__(self, engine, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.engine = engine
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    @staticmethod
    def owner(cell: int, world: int) -> int:
        return cell % world

    def search_graphed(self, Q, k: int, w: int = 1):
        """`search` replayed from a CUDA graph (one graph per batch shape): the step is ~20 small
        launches and four collectives, so at multi-GPU step times below a millisecond the Python /
        launch overhead would otherwise set the pace.  Q is copied into a static buffer; the returned
        tensors are the graph's static outputs (valid until the next call)."""
        key = (tuple(Q.shape), Q.dtype, k, w)
        if not hasattr(self, "_graphs"):
            self._graphs = {}
        if key not in self._graphs:
            static_q = Q.clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):          # warm-up: workspaces grow, kernel attributes are set
                    self.search(static_q, k, w)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            if hasattr(self.engine, "set_timing"):
                self.engine.set_timing(False)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self.search(static_q, k, w)
            self._graphs[key] = (g, static_q, out)
        g, static_q, out = self._graphs[key]
        static_q.copy_(Q, non_blocking=True)
        g.replay()
        return out

    def search(self, Q, k: int, w: int = 1):
        """Q is the full (replicated / broadcast) query batch on every rank."""
        nq = Q.shape[0]
        if self.world > 1 and hasattr(self.engine, "coarse") and nq >= self.world:
            # coarse step sharded by query: rank r assigns its slice; ONE all-gather of the probe lists
            # (cells and dc are both 4-byte elements when T is float32: packed side by side as int32)
            w = min(w, self.engine.kc)
            per = (nq + self.world - 1) // self.world
            lo = min(nq, self.rank * per)
            hi = min(nq, lo + per)
            if Q.dtype == torch.float32:
                packed = torch.zeros((per, 2, w), dtype=torch.int32, device=Q.device)
                if hi > lo:
                    c, d = self.engine.coarse(Q[lo:hi].contiguous(), w)
                    packed[:hi - lo, 0] = c
                    packed[:hi - lo, 1] = d.view(torch.int32)
                allp = torch.empty((self.world * per, 2, w), dtype=torch.int32, device=Q.device)
                self.dist.all_gather_into_tensor(allp, packed, group=self.group)
                cells_all = allp[:nq, 0].contiguous()
                dc_all = allp[:nq, 1].contiguous().view(torch.float32)
            else:
                cells_l = torch.zeros((per, w), dtype=torch.int32, device=Q.device)
                dc_l = torch.zeros((per, w), dtype=Q.dtype, device=Q.device)
                if hi > lo:
                    c, d = self.engine.coarse(Q[lo:hi].contiguous(), w)
                    cells_l[:hi - lo] = c
                    dc_l[:hi - lo] = d
                cells_all = torch.empty((self.world * per, w), dtype=torch.int32, device=Q.device)
                dc_all = torch.empty((self.world * per, w), dtype=Q.dtype, device=Q.device)
                self.dist.all_gather_into_tensor(cells_all, cells_l, group=self.group)
                self.dist.all_gather_into_tensor(dc_all, dc_l, group=self.group)
                cells_all, dc_all = cells_all[:nq].contiguous(), dc_all[:nq].contiguous()
            ids, dists, keys = self.engine.search_local_probes(Q, k, w, cells_all, dc_all)
        else:
            ids, dists, keys = self.engine.search_local(Q, k, w)
        if self.world == 1:
            return self.engine.merge(ids[None], dists[None], keys[None], k)
        # rank-major concatenation along dim 0 == [world, nq, k] (the layout ivfadc_merge_device takes);
        # the three gathers are issued back to back (asynchronously) and awaited together
        ids_all = torch.empty((self.world * nq, k), dtype=ids.dtype, device=ids.device)
        dists_all = torch.empty((self.world * nq, k), dtype=dists.dtype, device=ids.device)
        keys_all = torch.empty((self.world * nq, k), dtype=keys.dtype, device=ids.device)
        works = [self.dist.all_gather_into_tensor(ids_all, ids.contiguous(), group=self.group, async_op=True),
                 self.dist.all_gather_into_tensor(dists_all, dists.contiguous(), group=self.group, async_op=True),
                 self.dist.all_gather_into_tensor(keys_all, keys.contiguous(), group=self.group, async_op=True)]
        for wk in works:
            wk.wait()
        shape = (self.world, nq, k)
        return self.engine.merge(ids_all.view(shape), dists_all.view(shape), keys_all.view(shape), k)


class DeviceGroup:
    """One process, several GPUs (ivfadc_group_*): what the Julia glue binds when IVFADC_DEVICES names more than one
    device.  Same calls as a single index; the lists are sharded by cell inside the library."""

    def __init__(self, centroids, cb_vectors, cb_codes=None, *, devices=None, n_devices=None, index_type=np.uint32, flags=0):
        centroids = np.ascontiguousarray(centroids)
        cb_vectors = np.ascontiguousarray(cb_vectors, dtype=centroids.dtype)
        if cb_codes is None:
            cb_codes = np.tile(np.arange(cb_vectors.shape[1], dtype=np.uint8), (cb_vectors.shape[0], 1))
        cb_codes = np.ascontiguousarray(cb_codes, dtype=np.uint8)
        self.T, self.I = centroids.dtype, np.dtype(index_type)
        self.kc, self.nrows = centroids.shape
        self.m, self.k, self.dsub = cb_vectors.shape
        self._lib = _capi.load()
        devs = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
        n = len(devs) if devs is not None else int(n_devices)
        cfg = _capi.Config(dim=self.nrows, kc=self.kc, m=self.m, ksub=self.k,
                           dtype=_capi.F32 if self.T == np.float32 else _capi.F64, id_bytes=self.I.itemsize,
                           metric_coarse=_capi.SQEUCLIDEAN, metric_resid=_capi.SQEUCLIDEAN, device=0, shard_rank=0,
                           shard_world=1, flags=int(flags))
        g = ctypes.c_void_p()
        rc = self._lib.ivfadc_group_create(ctypes.byref(g), ctypes.byref(cfg), n, _capi.ptr(devs), _capi.ptr(centroids),
                                           _capi.ptr(cb_vectors), _capi.ptr(cb_codes))
        if rc != 0:
            raise _capi.IvfadcError(rc, "ivfadc_group_create failed (fewer devices than asked for, or NCCL not loadable)")
        self._g = g

    def _check(self, rc):
        if rc != 0:
            msg = self._lib.ivfadc_group_last_error(self._g)
            raise _capi.IvfadcError(rc, (msg or b"").decode("utf-8", "replace"))

    def close(self):
        if getattr(self, "_g", None):
            self._lib.ivfadc_group_destroy(self._g)
            self._g = None

    def __len__(self):
        n = ctypes.c_int64()
        self._check(self._lib.ivfadc_group_length(self._g, ctypes.byref(n)))
        return int(n.value)

    def size(self):
        return int(self._lib.ivfadc_group_size(self._g))

    def set_cell_owners(self, owners):
        owners = np.ascontiguousarray(owners, dtype=np.int32)
        self._check(self._lib.ivfadc_group_set_cell_owners(self._g, _capi.ptr(owners)))

    def add(self, X, position=_capi.LAST, assign=None, assign_base=0):
        X = np.ascontiguousarray(X, dtype=self.T)
        a = None if assign is None else np.ascontiguousarray(assign, dtype=np.int64)
        self._check(self._lib.ivfadc_group_add(self._g, _capi.ptr(X), X.shape[0], position, _capi.ptr(a), assign_base, None))

    def search(self, Q, k, w=1):
        Q = np.ascontiguousarray(Q, dtype=self.T).reshape(-1, self.nrows)
        nq = Q.shape[0]
        ids = np.empty((nq, k), dtype=np.uint64)
        dists = np.empty((nq, k), dtype=self.T)
        counts = np.empty(nq, dtype=np.int32)
        self._check(self._lib.ivfadc_group_search(self._g, _capi.ptr(Q), nq, k, w, _capi.ptr(ids), _capi.ptr(dists),
                                                  _capi.ptr(counts)))
        return ids, dists, counts

    def delete(self, ids0):
        ids0 = np.ascontiguousarray(ids0, dtype=np.uint64)
        self._check(self._lib.ivfadc_group_delete(self._g, _capi.ptr(ids0), len(ids0)))

    def pop(self, position=_capi.LAST):
        out = np.empty(self.nrows, dtype=self.T)
        self._check(self._lib.ivfadc_group_pop(self._g, position, _capi.ptr(out)))
        return out

    def shard_sizes(self):
        """list lengths per shard: int64 [n_devices, kc]"""
        out = np.zeros((self.size(), self.kc), dtype=np.int64)
        for i in range(self.size()):
            h = ctypes.c_void_p()
            self._check(self._lib.ivfadc_group_handle(self._g, i, ctypes.byref(h)))
            self._check(self._lib.ivfadc_list_sizes(h, _capi.ptr(out[i])))
        return out
