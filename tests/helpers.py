"""Shared builders for the CPU (oracle) and GPU (parity) tests."""
from __future__ import annotations

import numpy as np

from oracle import oracle as orc
from ivfadc_jl_b200 import training


def train(data_cols, kc, k, m, seed=0, maxiter=25):
    """data_cols: Julia-oriented (nrows, nvectors).  Returns (Quantizers, assignments 0-based, X [n, D])."""
    X = np.ascontiguousarray(np.asarray(data_cols).T)
    centroids, assign, cb_vectors, cb_codes = training.train_quantizers(X, kc, k, m, maxiter, maxiter, seed)
    return orc.Quantizers(centroids, cb_vectors, cb_codes), assign, X


def build_oracle_index(data_cols, kc, k, m, id_bytes=4, seed=0):
    qz, assign, X = train(data_cols, kc, k, m, seed)
    idx = orc.OracleIndex(qz, id_bytes=id_bytes).build(X, assign, assign_base=0)
    return idx, qz, assign, X


def reference_fixture(dtype=np.float64, seed=0, id_bytes=4):
    """test/index.jl:5-28: rand(10, 243), kc=100, k=16, m=2."""
    rng = np.random.default_rng(seed)
    data = rng.random((10, 243)).astype(dtype)
    return build_oracle_index(data, kc=100, k=16, m=2, id_bytes=id_bytes, seed=seed) + (data,)


TOY = np.array([[0, 0, 0, 1, 1, 1, 1, 1, 20, 20, 20, 20, 20],
                [0.1, 0.11, 0.12, 8, 10, 15, 14, 16, 5, 5.1, 5.2, 5.4, 5.5]], dtype=np.float64)
TOY_POINTS = [[1.0, 10.0], [0.0, 0.0], [20.0, 5.0]]
TOY_W1 = [[5, 4, 7, 6, 8], [1, 2, 3], [9, 10, 11, 12, 13]]
TOY_W2 = [[5, 4, 7, 6, 8], [1, 2, 3, 4, 5], [9, 10, 11, 12, 13]]


def lists_of(idx):
    """OracleIndex -> [(ids ndarray uint64, codes ndarray [len, m])] per cell."""
    out = []
    for ids, codes in idx.lists:
        c = np.array(codes, dtype=np.uint8).reshape(len(ids), idx.qz.m)
        out.append((np.array(ids, dtype=np.uint64), c))
    return out
