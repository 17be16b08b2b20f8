"""save_ivfadc_index / load_ivfadc_index in the reference's on-disk format
(src/persistency.jl:1-160, byte layout in SURVEY.md Appendix B):

  9 text lines  "{nrows} {nclusters}" / "{n} {m} {k} {d}" / coarse quantizer type / quantization
                type / U / I / Dc / Dr / T
  raw little-endian binary:
    centroids      nclusters x (nrows x T)                               (:44-49)
    per codebook   k x U codes, then d rows of k x T (row-major vectors)  (:52-61)
    rotation       nrows columns of nrows x T (identity for :pq)          (:62-64)
    per list       Int64 clsize, clsize x I ids, clsize x (m x U) codes   (:68-78)

Files written here load in the reference (type names are module-qualified, which its
_read_type_from_line accepts, :137-144) and files written by the reference load here.
The HNSW variant (:163-305) serialises a HNSW.jl graph the engine does not have (":hnsw" maps to
exact GPU search): saving such an index is left to the Julia glue, loading one is rejected.
"""
from __future__ import annotations

import numpy as np

from .index import IVFADCIndex

_JULIA_TO_NP = {"UInt8": np.uint8, "UInt16": np.uint16, "UInt32": np.uint32, "UInt64": np.uint64,
                "Float32": np.float32, "Float64": np.float64}
_NP_TO_JULIA = {np.dtype(v): k for k, v in _JULIA_TO_NP.items()}


def _type_name(line: str) -> str:
    """`Type` or `Module.Type` (src/persistency.jl:137-144) -> `Type`."""
    return line.strip().split(".")[-1]


def save_ivfadc_index(filename, ivfadc: IVFADCIndex):
    if ivfadc.coarse_quantizer != "naive":
        raise NotImplementedError("the HNSW graph is serialised by the Julia glue (HNSW.jl), not the engine")
    centroids, cb_vectors, cb_codes = ivfadc.quantizers()
    nrows, nclusters = ivfadc.nrows, ivfadc.kc
    n, m, k, d = len(ivfadc), ivfadc.m, ivfadc.k, ivfadc.dsub
    T, I = np.dtype(ivfadc.T), np.dtype(ivfadc.I)
    with open(filename, "wb") as f:
        header = [f"{nrows} {nclusters}", f"{n} {m} {k} {d}", "NaiveQuantizer",
                  "QuantizedArrays.OrthogonalQuantization", "UInt8", _NP_TO_JULIA[I],
                  "Distances." + ivfadc.coarse_distance, "Distances." + ivfadc.quantization_distance, _NP_TO_JULIA[T]]
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(centroids.astype(T.newbyteorder("<")).tobytes())
        for i in range(m):
            f.write(cb_codes[i].tobytes())
            # vectors[j, :] for j in 1:d -- the d x k matrix row by row; ours is stored [k, d]
            f.write(np.ascontiguousarray(cb_vectors[i].T).astype(T.newbyteorder("<")).tobytes())
        f.write(np.eye(nrows, dtype=T).tobytes())  # rot = I for :pq
        # one bulk device -> host transfer of all lists (ivfadc_export_all), then the per-list records of :68-78
        sizes, ids, codes = ivfadc.export_all()
        ids = ids.astype(I.newbyteorder("<"), copy=False)
        o = 0
        for c in range(nclusters):
            e = o + int(sizes[c])
            f.write(np.int64(sizes[c]).tobytes())
            f.write(ids[o:e].tobytes())
            f.write(codes[o:e].tobytes())
            o = e


def load_ivfadc_index(filename, device=0):
    with open(filename, "rb") as f:
        nrows, nclusters = (int(x) for x in f.readline().split())
        n, m, k, d = (int(x) for x in f.readline().split())
        cq = _type_name(f.readline().decode())
        quant = _type_name(f.readline().decode())
        U = _JULIA_TO_NP[_type_name(f.readline().decode())]
        I = _JULIA_TO_NP[_type_name(f.readline().decode())]
        dc = _type_name(f.readline().decode())
        dr = _type_name(f.readline().decode())
        T = np.dtype(_JULIA_TO_NP[_type_name(f.readline().decode())])
        if cq != "NaiveQuantizer":
            raise NotImplementedError(f"coarse quantizer {cq}: the HNSW graph section is handled by the Julia glue")
        if quant != "OrthogonalQuantization" or U is not np.uint8:
            raise NotImplementedError("only orthogonal (PQ) quantization with UInt8 codes is on the hot path")
        centroids = np.frombuffer(f.read(T.itemsize * nrows * nclusters), dtype=T).reshape(nclusters, nrows)
        cb_vectors = np.empty((m, k, d), dtype=T)
        cb_codes = np.empty((m, k), dtype=np.uint8)
        for i in range(m):
            cb_codes[i] = np.frombuffer(f.read(k), dtype=np.uint8)
            cb_vectors[i] = np.frombuffer(f.read(T.itemsize * k * d), dtype=T).reshape(d, k).T
        rot = np.frombuffer(f.read(T.itemsize * nrows * nrows), dtype=T).reshape(nrows, nrows)
        if not np.array_equal(rot, np.eye(nrows, dtype=T)):
            raise NotImplementedError("non-identity rotation (OPQ) is outside the hot-path scope")
        ivfadc = IVFADCIndex.from_quantizers(centroids, cb_vectors, cb_codes, index_type=I,
                                             coarse_distance=dc, quantization_distance=dr, device=device)
        # the list records of :119-131 are parsed on the host, the lists go to the device in one bulk call
        isz = np.dtype(I).itemsize
        raw = np.frombuffer(f.read(), dtype=np.uint8)
        sizes = np.zeros(nclusters, dtype=np.int64)
        ids = np.empty(n, dtype=np.uint64)
        codes = np.empty((n, m), dtype=np.uint8)
        p = o = 0
        for c in range(nclusters):
            clsize = int(raw[p:p + 8].view(np.int64)[0])
            p += 8
            if o + clsize > n:
                raise ValueError("list lengths exceed the vector count of the header")
            ids[o:o + clsize] = raw[p:p + isz * clsize].view(I)
            p += isz * clsize
            codes[o:o + clsize] = raw[p:p + m * clsize].reshape(clsize, m)
            p += m * clsize
            sizes[c] = clsize
            o += clsize
        ivfadc.import_all(sizes, ids[:o], codes[:o])
        return ivfadc
