"""N > 1 host logic on CPU: two gloo ranks drive ivfadc_jl_b200.sharded.ShardedSearcher with an
oracle-backed shard engine (test stand-in for the CUDA engine) and must reproduce the unsharded
oracle bit for bit: cell ownership (cell % world), merge keys (probe rank << 32 | position),
the all-gather layout [world, nq, k] and the (distance, key) merge order."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleShardEngine:
    """search_local / merge with the semantics of ivfadc_search_local_device / ivfadc_merge_device."""

    def __init__(self, qz, offsets, codes, ids, rank, world):
        self.qz, self.rank, self.world = qz, rank, world
        self.offsets, self.codes, self.ids = offsets, codes, ids

    @property
    def kc(self):
        return self.qz.kc

    def coarse(self, Q, w):
        from oracle import oracle as orc
        cells, dc = orc.coarse_search(self.qz, Q.numpy(), w)
        return torch.from_numpy(cells), torch.from_numpy(dc)

    def search_local(self, Q, k, w):
        from oracle import oracle as orc
        cells, _ = orc.coarse_search(self.qz, Q.numpy(), w)
        return self.search_local_probes(Q, k, w, torch.from_numpy(cells), None)

    def search_local_probes(self, Q, k, w, cells, dc):
        from oracle import oracle as orc
        Qn = Q.numpy()
        nq = Qn.shape[0]
        cells = cells.numpy()
        ids = np.full((nq, k), -1, dtype=np.int64)
        keys = np.full((nq, k), -1, dtype=np.int64)
        dists = np.full((nq, k), np.inf, dtype=Qn.dtype)
        for i in range(nq):
            cand = []
            for r, cell in enumerate(cells[i]):
                if cell % self.world != self.rank:
                    continue
                lo, hi = int(self.offsets[cell]), int(self.offsets[cell + 1])
                if hi == lo:
                    continue
                # one-cell oracle search = distances of the whole list in scan order
                sub_off = np.zeros(self.qz.kc + 1, dtype=np.int64)
                sub_off[cell + 1:] = hi - lo
                oi, od, oc, _ = orc.search_csr(self.qz, sub_off, self.codes[lo:hi], np.arange(hi - lo, dtype=np.uint64),
                                               Qn[i:i + 1], min(k, hi - lo), w)
                for j in range(oc[0]):
                    cand.append((od[0, j], (r << 32) | int(oi[0, j]), int(self.ids[lo + int(oi[0, j])])))
            cand.sort(key=lambda t: (t[0], t[1]))
            for j, (d, key, vid) in enumerate(cand[:k]):
                dists[i, j], keys[i, j], ids[i, j] = d, key, vid
        return torch.from_numpy(ids), torch.from_numpy(dists), torch.from_numpy(keys)

    def merge(self, ids_all, dists_all, keys_all, k):
        P, nq, _ = ids_all.shape
        ids = np.full((nq, k), -1, dtype=np.int64)
        dists = np.full((nq, k), np.inf, dtype=dists_all.numpy().dtype)
        counts = np.zeros(nq, dtype=np.int32)
        I, Dd, K = ids_all.numpy(), dists_all.numpy(), keys_all.numpy().view(np.uint64)
        for i in range(nq):
            cand = [(Dd[p, i, j], int(K[p, i, j]), int(I[p, i, j])) for p in range(P) for j in range(k)
                    if K[p, i, j] != np.uint64(2 ** 64 - 1)]
            cand.sort(key=lambda t: (t[0], t[1]))
            for j, (d, key, vid) in enumerate(cand[:k]):
                dists[i, j], ids[i, j] = d, vid
            counts[i] = min(k, len(cand))
        return torch.from_numpy(ids), torch.from_numpy(dists), torch.from_numpy(counts)


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ivfadc_jl_b200 import sharded
    from tests import helpers
    rng = np.random.default_rng(0)
    data = rng.random((16, 1500)).astype(np.float32)
    oidx, qz, assign, X = helpers.build_oracle_index(data, kc=13, k=32, m=4, seed=1)
    offsets, codes, ids = oidx.csr()
    Q = rng.random((40, 16)).astype(np.float32)
    k, w = 7, 5
    s = sharded.ShardedSearcher(OracleShardEngine(qz, offsets, codes, ids, rank, world))
    assert s.owner(11, world) == 11 % world
    gi, gd, gc = s.search(torch.from_numpy(Q), k, w)
    oi, od, oc = oidx.knn_search(Q, k, w=w)
    ok = (np.array_equal(gc.numpy(), oc) and np.array_equal(gd.numpy().view(np.uint8), od.view(np.uint8))
          and np.array_equal(gi.numpy().view(np.uint64), oi))
    open(os.path.join(tmp, f"ok{rank}"), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_search_matches_unsharded(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / f"ok{r}").read() for r in range(2)] == ["1", "1"]
