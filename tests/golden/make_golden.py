#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the CPU oracle.

The reference (pure Julia) cannot be executed in this image and ships no golden vectors
("parity unpinned"), so these fixtures pin the ORACLE: inputs (trained
quantizers, database, queries) and the oracle's outputs (cells, PQ codes, neighbour ids, distance
bits, counts) for two of the reference's own shapes.  `tests/test_golden.py` replays them against
the oracle on CPU (so the oracle cannot drift silently) and against the CUDA path on the GPU.

  python tests/golden/make_golden.py        # rewrites readme_f32.npz and reftest_f64.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from tests import helpers  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def make(name, data, kc, k, m, id_bytes, nq, searches, seed):
    oidx, qz, assign, X = helpers.build_oracle_index(data, kc=kc, k=k, m=m, id_bytes=id_bytes, seed=seed)
    cells, codes = orc.encode(qz, X, nthreads=2)                        # _encode_point path (coarse_search w = 1)
    bcells, bcodes = orc.encode(qz, X, assign=assign, assign_base=0, nthreads=2)   # build path (k-means assignments)
    Q = np.ascontiguousarray(X[:nq] + (np.random.default_rng(seed + 7).random((nq, X.shape[1])) * 0.05).astype(X.dtype))
    out = dict(centroids=qz.centroids, cb_vectors=qz.cb_vectors, cb_codes=qz.cb_codes, X=X, assign=assign, Q=Q,
               enc_cells=cells, enc_codes=codes, build_codes=bcodes,
               searches=np.array(searches, dtype=np.int64))
    ccells, cdc = orc.coarse_search(qz, Q, min(8, kc), nthreads=2)
    out["coarse_cells"], out["coarse_dc"] = ccells, cdc
    for (kk, w) in searches:
        oi, od, oc = oidx.knn_search(Q, kk, w=w, nthreads=2)
        out[f"ids_k{kk}_w{w}"], out[f"dists_k{kk}_w{w}"], out[f"counts_k{kk}_w{w}"] = oi, od, oc
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


METRIC_PAIRS = [("Euclidean", "SqEuclidean"), ("Cityblock", "Cityblock"), ("CosineDist", "Euclidean"), ("SqEuclidean", "CosineDist")]


def make_metrics(name, dtype, seed):
    """The other metrics of Distances.jl (Dc for coarse_search + lookup tables, Dr for quantize_data): one small data set,
    per (Dc, Dr) pair the oracle's cells, codes, coarse distances and search results."""
    rng = np.random.default_rng(seed)
    D, m, ksub, kc, n, nq, k, w = 12, 3, 32, 20, 600, 24, 5, 4
    X = (rng.random((n, D)) + 0.1).astype(dtype)
    cent = X[rng.choice(n, kc, replace=False)].copy()
    cb = (0.3 * rng.standard_normal((m, ksub, D // m))).astype(dtype)
    codes = np.stack([rng.permutation(256)[:ksub].astype(np.uint8) for _ in range(m)])
    Q = (rng.random((nq, D)) + 0.1).astype(dtype)
    out = dict(centroids=cent, cb_vectors=cb, cb_codes=codes, X=X, Q=Q, k=np.int64(k), w=np.int64(w))
    for dc, dr in METRIC_PAIRS:
        qz = orc.Quantizers(cent, cb, codes, coarse_distance=dc, quantization_distance=dr)
        cells, pq = orc.encode(qz, X, nthreads=2)
        ccells, cdc = orc.coarse_search(qz, Q, w, nthreads=2)
        order = np.argsort(cells, kind="stable")
        off = np.zeros(kc + 1, dtype=np.int64)
        np.cumsum(np.bincount(cells, minlength=kc), out=off[1:])
        oi, od, oc, _ = orc.search_csr(qz, off, pq[order], order.astype(np.uint64), Q, k, w, nthreads=2)
        t = f"{dc}_{dr}"
        out.update({f"cells_{t}": cells, f"codes_{t}": pq, f"ccells_{t}": ccells, f"cdc_{t}": cdc,
                    f"ids_{t}": oi, f"dists_{t}": od, f"counts_{t}": oc})
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    make_metrics("metrics_f32", np.float32, 5)
    make_metrics("metrics_f64", np.float64, 6)
    # config A, README.md:31-46 (shrunk database: 400 vectors, kc = 40, k = 64 keeps the file small)
    rng = np.random.default_rng(0)
    make("readme_f32", rng.random((50, 400)).astype(np.float32), kc=40, k=64, m=10, id_bytes=2, nq=48,
         searches=[(3, 1), (10, 16), (5, 40)], seed=0)
    # the reference's own test fixture, test/index.jl:5-28: rand(10, 243) Float64, kc = 100, k = 16, m = 2
    rng = np.random.default_rng(1)
    make("reftest_f64", rng.random((10, 243)).astype(np.float64), kc=100, k=16, m=2, id_bytes=4, nq=40,
         searches=[(1, 1), (5, 2), (10, 100)], seed=1)
