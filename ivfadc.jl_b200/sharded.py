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
