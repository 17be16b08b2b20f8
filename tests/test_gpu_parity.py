"""GPU parity tests: the CUDA path, driven through the C ABI (ctypes), against the CPU oracle on
the same seeded inputs.  Bar (BASELINE.json north_star): coarse assignments and PQ codes
bit-exact, ADC distances within 1e-5 relative, ids equal except at distance ties.  The CUDA
kernels use the oracle's exact fp evaluation order, so these tests assert the stronger
property: everything bit-identical, zero near-tie exceptions.
"""
import numpy as np
import pytest

import ivfadc_jl_b200 as iv
from oracle import oracle as orc
from tests import helpers

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # the tolerance north_star states for ADC distances (we assert exact equality too)


LEGACY, QLANE, LUT_EXACT, LUT_MMASYNC, TMEM_V1, COARSE_SCALAR, MERGE_SWEEP = 1, 2, 4, 8, 16, 32, 64  # ivfadc_config.flags (include/ivfadc.h)
COARSE_FFMA, COARSE_REDO = 128, 256
# Default flags for the bit-exact tests: whichever scan kernel the engine picks, tables in the
# reference's direct form.  The tensor-core (3xTF32) tables are tested at the stated tolerance in
# test_search_qlane_* below.


def engine_from(qz, id_type=np.uint32, X=None, assign=None, shard=(0, 1), flags=LUT_EXACT):
    e = iv.IVFADCIndex.from_quantizers(qz.centroids, qz.cb_vectors, qz.cb_codes, index_type=id_type,
                                       shard=shard, flags=flags)
    if X is not None:
        e._add(X, 0, assign=assign, assign_base=0)
    return e


def assert_lists_equal(engine, oidx):
    sizes = engine.list_sizes()
    for c, (ids, codes) in enumerate(helpers.lists_of(oidx)):
        assert sizes[c] == len(ids), f"cell {c}"
        gi, gc = engine.export_list(c)
        np.testing.assert_array_equal(gi.astype(np.uint64), ids, err_msg=f"ids of cell {c}")
        np.testing.assert_array_equal(gc, codes, err_msg=f"codes of cell {c}")


def assert_search_equal(engine, oidx, Q, k, w, nthreads=4):
    gi, gd, gc = engine.search_packed(Q, k, w)
    oi, od, oc = oidx.knn_search(Q, k, w=w, nthreads=nthreads)
    np.testing.assert_array_equal(gc, oc)
    np.testing.assert_allclose(gd, od, rtol=RTOL)           # the stated bar
    assert np.array_equal(gd.view(np.uint8), od.view(np.uint8)), "distances not bit-identical"
    np.testing.assert_array_equal(gi, oi)
    return gi, gd, gc


# ------------------------------------------------------------------------------------------------
# K1 coarse assignment
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("D,kc,nq,w", [(50, 100, 37, 1), (10, 100, 64, 2), (128, 1024, 300, 16),
                                       (2, 3, 5, 3), (96, 257, 33, 33), (17, 70, 9, 64),
                                       (128, 300, 130, 128), (200, 64, 40, 7), (64, 500, 1000, 16),
                                       (128, 1024, 2500, 16), (32, 64, 4000, 8)])
def test_coarse_search_bit_exact(dtype, D, kc, nq, w):
    rng = np.random.default_rng(D * 1000 + kc)
    cent = rng.random((kc, D)).astype(dtype)
    # duplicate centroids force exact distance ties -> lower cell index must win (stable sortperm)
    cent[kc // 2] = cent[0]
    cb = rng.standard_normal((1, 4, D)).astype(dtype)
    qz = orc.Quantizers(cent, cb)
    Q = rng.random((nq, D)).astype(dtype)
    oc, od = orc.coarse_search(qz, Q, w, nthreads=4)
    # default: packed-FP32 kernel on transposed centroids where the shape allows it (fp32, D % 4 == 0,
    # D <= 128); COARSE_SCALAR: the scalar FFMA kernel.  Same bits either way.
    for flags in (LUT_EXACT, LUT_EXACT | COARSE_SCALAR, LUT_EXACT | COARSE_FFMA):
        e = engine_from(qz, flags=flags)
        gc, gd = e.coarse_search(Q, w)
        np.testing.assert_array_equal(gc, oc, err_msg=f"flags={flags}")
        assert np.array_equal(gd.view(np.uint8), od.view(np.uint8)), f"flags={flags}"
        e.close()


@pytest.mark.parametrize("data", ["uniform", "blobs", "large_norm", "ties"])
@pytest.mark.parametrize("D,kc,nq,w", [(128, 1024, 1000, 16), (128, 1024, 129, 1), (96, 4096, 300, 16),
                                       (64, 300, 500, 8), (128, 256, 128, 32), (16, 700, 260, 5),
                                       (128, 2048, 4000, 16), (128, 1024, 1, 16), (32, 512, 7, 32)])
def test_coarse_tensor_core_bit_exact(data, D, kc, nq, w):
    """coarse_tc.cuh: TF32 tensor-core scores only prune; cells and distances must be bit-identical to the
    oracle's direct form (src/coarsequantizers.jl:33-37), stable ties included.  COARSE_REDO sends every
    query through the FFMA redo pass that serves candidate overflows."""
    rng = np.random.default_rng(D * 7 + kc + w)
    if data == "uniform":
        cent = rng.random((kc, D)).astype(np.float32)
        Q = rng.random((nq, D)).astype(np.float32)
    elif data == "blobs":
        cent = rng.random((kc, D)).astype(np.float32)
        Q = (cent[rng.integers(0, kc, nq)] + 0.05 * rng.standard_normal((nq, D))).astype(np.float32)
    elif data == "large_norm":  # distances small against the norms: the pruning margin is wide
        cent = (100.0 + rng.random((kc, D))).astype(np.float32)
        Q = (100.0 + rng.random((nq, D))).astype(np.float32)
    else:  # many exact duplicates: more candidates than slots -> redo pass; lower cell index wins
        base = rng.random((8, D)).astype(np.float32)
        cent = base[rng.integers(0, 8, kc)].copy()
        cent[::3] = rng.random((len(cent[::3]), D)).astype(np.float32)
        Q = rng.random((nq, D)).astype(np.float32)
    cb = rng.standard_normal((1, 4, D)).astype(np.float32)
    qz = orc.Quantizers(cent, cb)
    oc, od = orc.coarse_search(qz, Q, w, nthreads=4)
    for flags in (LUT_EXACT, LUT_EXACT | COARSE_REDO):
        e = engine_from(qz, flags=flags)
        gc, gd = e.coarse_search(Q, w)
        np.testing.assert_array_equal(gc, oc, err_msg=f"flags={flags}")
        assert np.array_equal(gd.view(np.uint8), od.view(np.uint8)), f"flags={flags}"
        e.close()


# ------------------------------------------------------------------------------------------------
# K4 encoding
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("D,m,ksub,kc", [(50, 10, 256, 100), (10, 2, 16, 100), (128, 16, 256, 64),
                                         (2, 2, 8, 3), (96, 12, 256, 50), (128, 8, 256, 32),
                                         (13, 4, 7, 5), (64, 16, 33, 9)])
def test_encode_bit_exact(dtype, D, m, ksub, kc):
    rng = np.random.default_rng(D + m + ksub)
    n = 777
    X = rng.random((n, D)).astype(dtype)
    cent = X[rng.choice(n, kc, replace=False)].copy()
    dsub = D // m
    cb = (0.3 * rng.standard_normal((m, ksub, dsub))).astype(dtype)
    cb[:, -1] = cb[:, 0]  # duplicate codeword: exact tie -> first (lowest column) wins
    codes = np.stack([rng.permutation(256)[:ksub].astype(np.uint8) for _ in range(m)])  # non-identity
    qz = orc.Quantizers(cent, cb, codes)
    e = engine_from(qz)
    gcell, gcode = e.encode(X)
    ocell, ocode = orc.encode(qz, X, nthreads=4)
    np.testing.assert_array_equal(gcell, ocell)
    np.testing.assert_array_equal(gcode, ocode)
    # build path: cells given (1-based like Clustering.jl's assignments)
    assign = rng.integers(1, kc + 1, size=n)
    gcell, gcode = e.encode(X, assign=assign, assign_base=1)
    ocell, ocode = orc.encode(qz, X, assign=assign, assign_base=1, nthreads=4)
    np.testing.assert_array_equal(gcell, ocell)
    np.testing.assert_array_equal(gcode, ocode)
    e.close()


# ------------------------------------------------------------------------------------------------
# K2 + K3 search
# ------------------------------------------------------------------------------------------------
def test_search_readme_config():
    """Config A (README.md:31-46,90-91): 50-d Float32, 1000 vectors, kc=100, k=256, m=10, UInt16."""
    rng = np.random.default_rng(0)
    data = rng.random((50, 1000)).astype(np.float32)
    oidx, qz, assign, X = helpers.build_oracle_index(data, kc=100, k=256, m=10, id_bytes=2, seed=0)
    e = engine_from(qz, np.uint16, X, assign)
    assert repr(e) == "IVFADCIndex, naive coarse quantizer, 12-byte encoding (2 + 1×10), 1000 Float32 vectors"
    assert_lists_equal(e, oidx)
    point = data[:, 122].copy()
    idxs, dists = iv.knn_search(e, point, 3)
    assert idxs.dtype == np.uint16 and dists.dtype == np.float32
    oi, od, oc = oidx.knn_search(point[None, :], 3, w=1)
    np.testing.assert_array_equal(idxs.astype(np.uint64), oi[0, :oc[0]])
    assert np.array_equal(dists, od[0, :oc[0]])
    Q = np.ascontiguousarray(data.T[:200])
    for k, w in ((3, 1), (10, 16), (1, 100), (32, 7), (33, 5), (128, 40)):
        assert_search_equal(e, oidx, Q, k, w)
    e.close()


@pytest.mark.parametrize("seed", [0, 1])
def test_search_reference_fixture_float64(seed):
    """test/index.jl:5-28 shapes: Float64, 10 x 243, kc=100 (many empty / tiny lists), k=16, dsub=5."""
    oidx, qz, assign, X, data = helpers.reference_fixture(np.float64, seed)
    e = engine_from(qz, np.uint32, X, assign)
    assert_lists_equal(e, oidx)
    rng = np.random.default_rng(seed + 10)
    Q = rng.random((50, 10))
    for k, w in ((3, 2), (5, 1), (10, 100), (40, 30), (1, 1), (128, 128)):
        assert_search_equal(e, oidx, Q, k, w)   # w > kc is clamped (src/index.jl:216)
    idxs, dists = iv.knn_search(e, Q[0], 3, w=2)
    assert idxs.dtype == np.uint32 and dists.dtype == np.float64
    li, ld = iv.knn_search(e, [q for q in Q[:10]], 3, w=2)
    assert len(li) == 10 and all(x.dtype == np.uint32 for x in li)
    with pytest.raises(AssertionError):
        iv.knn_search(e, Q[0], 0)
    with pytest.raises(AssertionError):
        iv.knn_search(e, Q[0], 1, w=0)
    with pytest.raises(TypeError):
        iv.knn_search(e, Q[0].astype(np.float32), 1)
    e.close()


@pytest.mark.parametrize("seed", range(3))
def test_search_toy_known_answer(seed):
    """test/search.jl:26-49 through the CUDA path."""
    oidx, qz, assign, X = helpers.build_oracle_index(helpers.TOY, kc=3, k=8, m=2, seed=seed)
    e = engine_from(qz, np.uint32, X, assign)
    for w, expected in ((1, helpers.TOY_W1), (2, helpers.TOY_W2)):
        for point, result in zip(helpers.TOY_POINTS, expected):
            idxs, _ = iv.knn_search(e, np.array(point), 5, w=w)
            assert set(int(i) + 1 for i in idxs) <= set(result)
    assert_search_equal(e, oidx, np.array(helpers.TOY_POINTS), 5, 2)
    e.close()


@pytest.mark.parametrize("dtype,D,m,ksub,kc,n,nq,k,w", [
    (np.float32, 128, 16, 256, 64, 20000, 500, 10, 16),    # config-B-shaped, small
    (np.float32, 96, 12, 256, 128, 30000, 300, 10, 16),    # config-C-shaped (m = 12)
    (np.float32, 128, 8, 256, 32, 20000, 300, 10, 8),      # config-D-shaped (m = 8, dsub = 16)
    (np.float64, 128, 16, 256, 32, 8000, 200, 10, 8),      # Float64 fast path (QN = 2)
    (np.float32, 30, 7, 100, 16, 5000, 200, 20, 4),        # generic m / dsub / ksub
    (np.float32, 64, 32, 256, 16, 5000, 100, 5, 4),        # m = 32 (QN = 2)
    (np.float32, 160, 80, 16, 8, 3000, 100, 5, 3),         # m = 80 (QN = 1)
    (np.float32, 33, 16, 64, 20, 6000, 150, 64, 5),        # trailing dim ignored by the PQ, k = 64
])
def test_search_midsize_bit_exact(dtype, D, m, ksub, kc, n, nq, k, w):
    from ivfadc_jl_b200 import synth
    X = synth.blobs(n, D, kc, seed=11, dtype=dtype)
    cent, cb, codes = synth.random_quantizers(kc, D, m, ksub, seed=5, dtype=dtype, data=X)
    qz = orc.Quantizers(cent, cb, codes)
    cells, ocodes = orc.encode(qz, X, nthreads=8)
    oidx = None
    e = engine_from(qz, np.uint32, X)
    # CSR for the oracle from its own encoding (list order = ascending id)
    order = np.argsort(cells, kind="stable")
    offsets = np.zeros(kc + 1, dtype=np.int64)
    np.cumsum(np.bincount(cells, minlength=kc), out=offsets[1:])
    Q = synth.blobs(nq, D, kc, seed=12, dtype=dtype)
    oi, od, oc, scanned = orc.search_csr(qz, offsets, ocodes[order], order.astype(np.uint64), Q, k, w, nthreads=8)
    e.stats(reset=True)
    gi, gd, gc = e.search_packed(Q, k, w)
    np.testing.assert_array_equal(gc, oc)
    assert np.array_equal(gd.view(np.uint8), od.view(np.uint8))
    np.testing.assert_array_equal(gi, oi)
    st = e.stats()
    assert st["scanned_vectors"] == scanned and st["scan_code_bytes"] == scanned * m
    # duplicates in the database -> exact distance ties -> order by (probe rank, position)
    e.close()


def test_search_ties_and_duplicates():
    """Exact ties: many identical database vectors; the reference order is (distance, probe rank,
    position in list) with strict '>' on replacement (src/index.jl:247-254)."""
    rng = np.random.default_rng(4)
    base = rng.random((40, 16)).astype(np.float32)
    X = np.concatenate([base] * 25)  # every vector 25 times
    rng.shuffle(X)
    data = np.ascontiguousarray(X.T)
    oidx, qz, assign, Xc = helpers.build_oracle_index(data, kc=8, k=16, m=4, seed=1)
    e = engine_from(qz, np.uint32, Xc, assign)
    Q = np.concatenate([base[:20], rng.random((20, 16)).astype(np.float32)])
    for k, w in ((10, 3), (30, 8), (100, 2)):
        assert_search_equal(e, oidx, Q, k, w)
    e.close()


# ------------------------------------------------------------------------------------------------
# K2 + K3, query-per-lane kernel (large batches): same bit-exact bar, pinned with QLANE
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("D,m,ksub,kc,n,nq,k,w,identity", [
    (128, 16, 256, 64, 20000, 500, 10, 16, True),    # config-B-shaped: ~125 queries per list
    (96, 12, 256, 128, 30000, 300, 10, 16, True),    # config-C-shaped (m = 12: 3 table chunks)
    (128, 8, 256, 32, 20000, 300, 10, 8, True),      # config-D-shaped (m = 8, dsub = 16: two 8-dim tables per code byte)
    (64, 4, 200, 40, 10000, 200, 10, 8, False),      # m = 4, dsub = 16, permuted code values, ksub < 256
    (128, 16, 256, 8, 20000, 100, 16, 8, True),      # 2500-vector lists: 3 passes of 1024, k = 16
    (40, 8, 100, 16, 5000, 200, 1, 4, False),        # dsub = 5 (generic), ksub < 256, permuted codes, k = 1
    (35, 4, 37, 50, 3000, 77, 7, 50, False),         # trailing dims ignored, tiny lists, w = kc
    (64, 16, 256, 700, 3000, 40, 10, 3, True),       # mostly empty / 1-2 vector lists, nearly empty groups
])
def test_search_qlane_bit_exact(D, m, ksub, kc, n, nq, k, w, identity):
    from ivfadc_jl_b200 import synth
    X = synth.blobs(n, D, kc, seed=21)
    cent, cb, codes = synth.random_quantizers(kc, D, m, ksub, seed=6, data=X)
    if not identity:  # code VALUE -> table entry (Q6): values are a random subset of 0..255
        prng = np.random.default_rng(3)
        codes = np.stack([prng.permutation(256)[:ksub].astype(np.uint8) for _ in range(m)])
    qz = orc.Quantizers(cent, cb, codes)
    cells, ocodes = orc.encode(qz, X, nthreads=8)
    order = np.argsort(cells, kind="stable")
    offsets = np.zeros(kc + 1, dtype=np.int64)
    np.cumsum(np.bincount(cells, minlength=kc), out=offsets[1:])
    Q = synth.blobs(nq, D, kc, seed=22)
    oi, od, oc, scanned = orc.search_csr(qz, offsets, ocodes[order], order.astype(np.uint64), Q, k, w, nthreads=8)
    for flags in (QLANE | LUT_EXACT, LEGACY):
        e = engine_from(qz, np.uint32, X, flags=flags)
        gi, gd, gc = e.search_packed(Q, k, w)
        np.testing.assert_array_equal(gc, oc, err_msg=f"flags={flags}")
        assert np.array_equal(gd.view(np.uint8), od.view(np.uint8)), f"flags={flags}"
        np.testing.assert_array_equal(gi, oi, err_msg=f"flags={flags}")
        e.close()
    # tensor-core tables (the default for large batches): the north_star's tolerance bar.
    # QLANE alone = warp-specialised tensor-memory lookup kernel (tcgen05.mma tables looked up with tcgen05.ld)
    # where the shape allows it (dsub <= 8 or 16), QLANE | TMEM_V1 = the round-1 version of the same arithmetic
    # (identical bits), QLANE | LUT_MMASYNC = the warp-level mma.sync builder.
    # QLANE | MERGE_SWEEP: the heavy-tie fallback of the final selection on ordinary data.
    first = None
    for flags in (QLANE, QLANE | MERGE_SWEEP, QLANE | TMEM_V1, QLANE | LUT_MMASYNC):
        e = engine_from(qz, np.uint32, X, flags=flags)
        gi, gd, gc = e.search_packed(Q, k, w)
        if flags in (QLANE, QLANE | MERGE_SWEEP) and scanu_shape(D, m):
            assert e.stats()["last_scan_kernel"] == 5, e.stats()
        rep = orc.compare_search(gi, gd, gc, oi, od, oc, rtol=RTOL)
        assert rep["near_tie_id_mismatches"] <= max(2, rep["results"] // 200), (flags, rep)
        assert rep["max_rel_err"] < 3e-6, (flags, rep)   # measured error budget of 3xTF32
        if flags == QLANE:
            first = (gi.copy(), gd.copy(), gc.copy())
        elif flags == QLANE | MERGE_SWEEP:  # same candidates, other selection path: identical bits
            np.testing.assert_array_equal(gi, first[0])
            assert np.array_equal(gd.view(np.uint8), first[1].view(np.uint8))
            np.testing.assert_array_equal(gc, first[2])
        e.close()


def scanu_shape(D, m):
    """Shapes served by the tensor-memory lookup kernels (scan.cu, scanu_shape_ok)."""
    dsub = D // m
    return (dsub <= 8 or dsub == 16) and m % 4 == 0 and m * (2 if dsub == 16 else 1) <= 16


def test_scan_kernel_choice():
    """The engine picks the scan kernel from a cost model ("which scan kernel"): many queries per list ->
    the tensor-memory query-per-lane kernel (4), few queries per long list (config-D-like) -> vector per lane (1).
    Both answers stay within the tolerance bar of the oracle."""
    from ivfadc_jl_b200 import synth
    for (D, m, kc, n, nq, w, want) in ((128, 16, 16, 16000, 600, 8, 5), (128, 8, 64, 384000, 100, 8, 1)):
        X = synth.blobs(n, D, kc, seed=31)
        cent, cb, codes = synth.random_quantizers(kc, D, m, 256, seed=7, data=X)
        cent = synth.blob_centres(D, kc)   # balanced lists of n / kc vectors
        qz = orc.Quantizers(cent, cb, codes)
        cells, ocodes = orc.encode(qz, X, nthreads=8)
        order = np.argsort(cells, kind="stable")
        offsets = np.zeros(kc + 1, dtype=np.int64)
        np.cumsum(np.bincount(cells, minlength=kc), out=offsets[1:])
        Q = synth.blobs(nq, D, kc, seed=32)
        oi, od, oc, _ = orc.search_csr(qz, offsets, ocodes[order], order.astype(np.uint64), Q, 10, w, nthreads=8)
        e = engine_from(qz, np.uint32, X, flags=0)
        gi, gd, gc = e.search_packed(Q, 10, w)
        assert e.stats()["last_scan_kernel"] == want, (want, e.stats())
        orc.compare_search(gi, gd, gc, oi, od, oc, rtol=RTOL)
        e.close()
    # a shard owns every fourth cell: the model has to count ITS cells, pairs and vectors (a quarter of each),
    # or a config-B-like shard falls to the vector-per-lane kernel (seen at N = 4: scan 0.64 ms instead of 0.26)
    import torch
    from ivfadc_jl_b200 import sharded
    D, m, kc, n, nq, w = 128, 16, 64, 64000, 2400, 8
    X = synth.blobs(n, D, kc, seed=33)
    _, cb, codes = synth.random_quantizers(kc, D, m, 256, seed=8, data=X)
    qz = orc.Quantizers(synth.blob_centres(D, kc), cb, codes)
    e = engine_from(qz, np.uint32, X, shard=(1, 4), flags=0)
    sharded.search_local(e, torch.from_numpy(synth.blobs(nq, D, kc, seed=34)).cuda(), 10, w)
    torch.cuda.synchronize()
    assert e.stats()["last_scan_kernel"] == 5, e.stats()
    e.close()


@pytest.mark.parametrize("D,m,ksub,identity", [(128, 16, 256, True), (96, 12, 256, True), (40, 8, 100, False),
                                               (16, 4, 256, True)])
def test_tcgen05_tables_against_fp64(D, m, ksub, identity):
    """K2 in isolation: the lookup tables the tcgen05 builder leaves in tensor memory for the first work
    item (read back with tcgen05.ld and dumped through ivfadc_debug_tables) against |w|^2 - 2 r.w
    evaluated in float64."""
    import ctypes
    from ivfadc_jl_b200 import synth
    kc, n, nq, w = 8, 4000, 64, 2
    X = synth.blobs(n, D, kc, seed=31)
    cent, cb, codes = synth.random_quantizers(kc, D, m, ksub, seed=7, data=X)
    if not identity:
        prng = np.random.default_rng(4)
        codes = np.stack([prng.permutation(256)[:ksub].astype(np.uint8) for _ in range(m)])
    qz = orc.Quantizers(cent, cb, codes)
    Q = synth.blobs(nq, D, kc, seed=32)
    e = engine_from(qz, np.uint32, X, flags=QLANE)
    iv._capi.check(e._h, e._lib.ivfadc_debug_tables(e._h, None))
    e.search_packed(Q, 5, w)
    buf = np.zeros(m * 256 * 32 + 64 + 1024, dtype=np.float32)   # tables, slots, cell, timeline stamps
    iv._capi.check(e._h, e._lib.ivfadc_debug_tables(e._h, buf.ctypes.data_as(ctypes.c_void_p)))
    e.close()
    tables = buf[:m * 256 * 32].reshape(m, 256, 32)
    meta = buf[m * 256 * 32:].view(np.int32)
    pairs, cell = meta[:32], int(meta[32])
    assert 0 <= cell < kc and (pairs >= 0).any()
    dsub = D // m
    codes_np = np.asarray(qz.cb_codes)
    worst = 0.0
    for slot, p in enumerate(pairs):
        if p < 0:
            continue
        r = Q[p // w].astype(np.float64) - cent[cell].astype(np.float64)
        for s in range(m):
            rs = r[s * dsub:(s + 1) * dsub]
            wv = cb[s].astype(np.float64)                      # [ksub, dsub]
            want = (wv * wv).sum(1) - 2.0 * wv @ rs            # entry of codeword c lives at row codes[s][c]
            got = tables[s, codes_np[s].astype(np.int64), slot].astype(np.float64)
            scale = np.abs(wv * wv).sum(1) + 2.0 * np.abs(wv) @ np.abs(rs) + 1e-30
            worst = max(worst, float(np.max(np.abs(got - want) / scale)))
    assert worst < 2e-6, worst


def test_search_qlane_ties_overflow_redo():
    """Hundreds of identical database vectors nearest to the query: more candidates tie at the
    k-th-distance bound than the 64 shared-memory slots hold, so the query-per-lane kernel must
    hand those (query, list) pairs to the general kernel (redo queue) -- same answer, in the
    reference's (distance, probe rank, position) order."""
    rng = np.random.default_rng(9)
    base = rng.random((30, 32)).astype(np.float32)
    X = np.concatenate([np.repeat(base[:3], 300, axis=0), np.repeat(base[3:], 20, axis=0)])
    rng.shuffle(X)
    data = np.ascontiguousarray(X.T)
    oidx, qz, assign, Xc = helpers.build_oracle_index(data, kc=4, k=16, m=8, seed=1)
    Q = np.concatenate([base, rng.random((34, 32)).astype(np.float32)])
    for flags in (QLANE | LUT_EXACT, LUT_EXACT):
        e = engine_from(qz, np.uint32, Xc, assign, flags=flags)
        for k, w in ((10, 2), (16, 4), (1, 1)):
            assert_search_equal(e, oidx, Q, k, w)
        e.close()
    for flags in (QLANE, QLANE | MERGE_SWEEP, QLANE | TMEM_V1):
        e = engine_from(qz, np.uint32, Xc, assign, flags=flags)
        for k, w in ((10, 2), (16, 4), (1, 1)):
            gi, gd, gc = e.search_packed(Q, k, w)
            oi, od, oc = oidx.knn_search(Q, k, w=w, nthreads=4)
            orc.compare_search(gi, gd, gc, oi, od, oc, rtol=RTOL)
        e.close()


# ------------------------------------------------------------------------------------------------
# K5 mutation
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,id_type,id_bytes", [(np.float64, np.uint8, 1), (np.float32, np.uint32, 4),
                                                    (np.float32, np.uint64, 8)])
def test_mutation_sequence_matches_reference_semantics(dtype, id_type, id_bytes):
    """test/utils.jl scenarios, every step checked list-by-list against the literal restatement."""
    oidx, qz, assign, X, data = helpers.reference_fixture(dtype, seed=2, id_bytes=id_bytes)
    e = engine_from(qz, id_type, X, assign)
    rng = np.random.default_rng(3)
    assert len(e) == 243 and e.size() == (10, 243) and e.size(1) == 10
    for _ in range(13):  # push! up to 256
        p = rng.random(10).astype(dtype)
        oidx.push(p)
        assert iv.push(e, p) is None
    assert len(e) == 256
    assert_lists_equal(e, oidx)
    if id_bytes == 1:
        with pytest.raises(AssertionError):
            iv.push(e, rng.random(10).astype(dtype))  # index is full (test/utils.jl:13)
    oidx.delete_from_index([1])
    iv.delete_from_index(e, [1])
    with pytest.raises(AssertionError):
        iv.push(e, rng.random(11).astype(dtype))  # wrong dimension (test/utils.jl:15)
    for i in range(1, 13):
        oidx.delete_from_index([i])
        iv.delete_from_index(e, [i])
    assert_lists_equal(e, oidx)
    for _ in range(13):
        p = rng.random(10).astype(dtype)
        oidx.pushfirst(p)
        iv.pushfirst(e, p)
    assert len(e) == 256
    assert_lists_equal(e, oidx)
    # pop! / popfirst! (test/utils.jl:32-55)
    for _ in range(3):
        vo, vg = oidx.pop(), iv.pop(e)
        assert vg.dtype == dtype and np.array_equal(vo, vg)
        vo, vg = oidx.popfirst(), iv.popfirst(e)
        assert np.array_equal(vo, vg)
    assert len(e) == 250
    assert_lists_equal(e, oidx)
    # delete_from_index! with the reference's ranges (test/utils.jl:58-105) + duplicates + unknown ids
    n = len(e)
    dele = list(range(1, 6)) + list(range(10, 31)) + list(range(n - 5, n + 1)) + [3, 3, 12] + ([100000] if id_bytes > 1 else [255, 256])
    oidx.delete_from_index(dele)
    iv.delete_from_index(e, dele)
    assert len(e) == len(oidx)
    assert_lists_equal(e, oidx)
    with pytest.raises(OverflowError):
        iv.delete_from_index(e, [0])
    # search after all that still agrees (scan order = list order, not id order: Q9)
    Q = rng.random((40, 10)).astype(dtype)
    assert_search_equal(e, oidx, Q, 7, 20)
    # batched push == n single pushes
    P = rng.random((20, 10)).astype(dtype)
    if id_bytes > 1:
        for p in P:
            oidx.push(p)
        iv.push_batch(e, P)
        for p in P:
            oidx.pushfirst(p)
        iv.push_batch(e, P, first=True)
        assert_lists_equal(e, oidx)
    e.close()


def test_pop_until_empty_and_regrow():
    rng = np.random.default_rng(8)
    data = rng.random((8, 40)).astype(np.float32)
    oidx, qz, assign, X = helpers.build_oracle_index(data, kc=4, k=8, m=2, seed=0)
    e = engine_from(qz, np.uint32, X, assign)
    for i in range(40):
        a, b = (oidx.pop(), iv.pop(e)) if i % 2 else (oidx.popfirst(), iv.popfirst(e))
        assert np.array_equal(a, b)
    assert len(e) == 0
    with pytest.raises(AssertionError):
        iv.pop(e)
    ids, d = iv.knn_search(e, X[0], 3, w=4)
    assert len(ids) == 0 and len(d) == 0            # all probed lists empty (Q8)
    big = rng.random((5000, 8)).astype(np.float32)  # forces several arena regrows
    for s in range(0, 5000, 700):
        iv.push_batch(e, big[s:s + 700])
    cells, codes = orc.encode(qz, big)
    assert len(e) == 5000
    for c in range(4):
        gi, gc = e.export_list(c)
        sel = np.flatnonzero(cells == c)
        np.testing.assert_array_equal(gi, sel.astype(np.uint32))
        np.testing.assert_array_equal(gc, codes[sel])
    e.close()


def test_constructor_asserts_and_build():
    """test/index.jl:31-42 through the Python mirror of the constructor."""
    rng = np.random.default_rng(0)
    data = rng.random((2, 300))
    with pytest.raises(AssertionError):
        iv.IVFADCIndex(data, kc=1, k=2, m=1)
    with pytest.raises(AssertionError):
        iv.IVFADCIndex(data, kc=2, k=301, m=1)
    with pytest.raises(AssertionError):
        iv.IVFADCIndex(data, kc=2, k=300, m=3)
    with pytest.raises(AssertionError):
        iv.IVFADCIndex(data, index_type=np.uint8)
    for cq in ("naive", "hnsw"):
        d = rng.random((10, 243))
        e = iv.IVFADCIndex(d, kc=100, k=16, m=2, coarse_quantizer=cq, index_type=np.uint32)
        assert len(e) == 243
        ids, dists = iv.knn_search(e, d[:, 5].copy(), 3, w=2)
        assert len(ids) <= 3 and np.all(np.diff(dists) >= 0)
        e.close()


def test_persistency_roundtrip(tmp_path):
    """test/persistency.jl:1-35: save, load, every field equal; plus the byte layout of Appendix B."""
    oidx, qz, assign, X, data = helpers.reference_fixture(np.float64, seed=4)
    e = engine_from(qz, np.uint16, X, assign)
    fn = str(tmp_path / "index.ivfadc")
    iv.save_ivfadc_index(fn, e)
    e2 = iv.load_ivfadc_index(fn)
    assert repr(e2) == repr(e) and len(e2) == len(e)
    c1, v1, k1 = e.quantizers()
    c2, v2, k2 = e2.quantizers()
    assert np.array_equal(c1, c2) and np.array_equal(v1, v2) and np.array_equal(k1, k2)
    assert_lists_equal(e2, oidx)
    Q = np.random.default_rng(1).random((20, 10))
    assert_search_equal(e2, oidx, Q, 5, 10)
    # independent parse of the file following src/persistency.jl line by line
    raw = open(fn, "rb").read()
    lines = raw.split(b"\n", 9)
    assert lines[0] == b"10 100" and lines[1] == b"243 2 16 5" and lines[2] == b"NaiveQuantizer"
    assert lines[4] == b"UInt8" and lines[5] == b"UInt16" and lines[8] == b"Float64"
    body = lines[9]
    cent = np.frombuffer(body[:8 * 10 * 100], dtype=np.float64).reshape(100, 10)
    assert np.array_equal(cent, qz.centroids)
    off = 8 * 1000
    for i in range(2):
        assert np.array_equal(np.frombuffer(body[off:off + 16], dtype=np.uint8), qz.cb_codes[i])
        off += 16
        rows = np.frombuffer(body[off:off + 8 * 16 * 5], dtype=np.float64).reshape(5, 16)  # vectors[j, :]
        assert np.array_equal(rows.T, qz.cb_vectors[i])
        off += 8 * 16 * 5
    assert np.array_equal(np.frombuffer(body[off:off + 800], dtype=np.float64).reshape(10, 10), np.eye(10))
    off += 800
    for c, (ids, codes) in enumerate(helpers.lists_of(oidx)):
        n = int(np.frombuffer(body[off:off + 8], dtype=np.int64)[0]); off += 8
        assert n == len(ids)
        assert np.array_equal(np.frombuffer(body[off:off + 2 * n], dtype=np.uint16), ids.astype(np.uint16)); off += 2 * n
        assert np.array_equal(np.frombuffer(body[off:off + 2 * n], dtype=np.uint8).reshape(n, 2), codes); off += 2 * n
    assert off == len(body)
    e.close(); e2.close()


# ------------------------------------------------------------------------------------------------
# sharding (two shard handles on one GPU; the N>1 transport is covered by the gloo test)
# ------------------------------------------------------------------------------------------------
def test_two_shards_merge_equals_unsharded():
    import torch
    from ivfadc_jl_b200 import sharded
    rng = np.random.default_rng(6)
    data = rng.random((32, 6000)).astype(np.float32)
    oidx, qz, assign, X = helpers.build_oracle_index(data, kc=24, k=64, m=8, seed=2)
    shards = [engine_from(qz, np.uint32, X, assign, shard=(r, 2)) for r in range(2)]
    assert sum(len(s.list_sizes().nonzero()[0]) for s in shards) <= 24
    assert all(len(s) == 6000 for s in shards)
    Q = rng.random((333, 32)).astype(np.float32)
    k, w = 10, 6
    dQ = torch.from_numpy(Q).cuda()
    parts = [sharded.search_local(s, dQ, k, w) for s in shards]
    ids, dists, counts = sharded.merge_parts(shards[0], parts, k)
    oi, od, oc = oidx.knn_search(Q, k, w=w, nthreads=4)
    np.testing.assert_array_equal(counts.cpu().numpy(), oc)
    assert np.array_equal(dists.cpu().numpy().view(np.uint8), od.view(np.uint8))
    np.testing.assert_array_equal(ids.cpu().numpy().astype(np.uint64), oi)
    # coarse step sharded by query: each shard assigns half of the batch, probe lists are concatenated
    # (what the all-gather does) and handed to ivfadc_search_probes_local_device
    half = 170
    c0, d0 = sharded.coarse_device(shards[0], dQ[:half].contiguous(), w)
    c1, d1 = sharded.coarse_device(shards[1], dQ[half:].contiguous(), w)
    cells, dc = torch.cat([c0, c1]).contiguous(), torch.cat([d0, d1]).contiguous()
    ocells, odc = orc.coarse_search(qz, Q, w, nthreads=4)
    np.testing.assert_array_equal(cells.cpu().numpy(), ocells)
    assert np.array_equal(dc.cpu().numpy().view(np.uint8), odc.view(np.uint8))
    parts = [sharded.search_local_probes(s, dQ, k, w, cells, dc) for s in shards]
    ids2, dists2, counts2 = sharded.merge_parts(shards[0], parts, k)
    np.testing.assert_array_equal(counts2.cpu().numpy(), oc)
    assert np.array_equal(dists2.cpu().numpy().view(np.uint8), od.view(np.uint8))
    np.testing.assert_array_equal(ids2.cpu().numpy().astype(np.uint64), oi)
    # mutation on shards: delete + pushfirst keep the global numbering
    dele = [5, 17, 400, 5999, 6000]
    oidx.delete_from_index(dele)
    P = rng.random((3, 32)).astype(np.float32)
    for p in P:
        oidx.pushfirst(p)
    for s in shards:
        iv.delete_from_index(s, dele)
        iv.push_batch(s, P, first=True)
        assert len(s) == len(oidx)
    parts = [sharded.search_local(s, dQ, k, w) for s in shards]
    ids, dists, counts = sharded.merge_parts(shards[0], parts, k)
    oi, od, oc = oidx.knn_search(Q, k, w=w, nthreads=4)
    np.testing.assert_array_equal(ids.cpu().numpy().astype(np.uint64), oi)
    assert np.array_equal(dists.cpu().numpy().view(np.uint8), od.view(np.uint8))
    for s in shards:
        s.close()


def test_device_api_matches_host_api():
    import torch
    from ivfadc_jl_b200 import sharded
    rng = np.random.default_rng(7)
    data = rng.random((64, 4000)).astype(np.float32)
    oidx, qz, assign, X = helpers.build_oracle_index(data, kc=16, k=64, m=16, seed=3)
    e = engine_from(qz, np.uint32, X, assign)
    Q = rng.random((257, 64)).astype(np.float32)
    hi, hd, hc = e.search_packed(Q, 10, 4)
    di, dd, dcn = sharded.search_device(e, torch.from_numpy(Q).cuda(), 10, 4)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(di.cpu().numpy().astype(np.uint64), hi)
    assert np.array_equal(dd.cpu().numpy(), hd)
    np.testing.assert_array_equal(dcn.cpu().numpy(), hc)
    e.close()


# ------------------------------------------------------------------------------------------------
# Full BASELINE.json size (config B: 1M x 128-d, kc = 1024, m = 16, 10k queries, nprobe 16, k 10):
# size-independent properties + a sample against the oracle
# ------------------------------------------------------------------------------------------------
def test_config_b_full_size_properties():
    from ivfadc_jl_b200 import synth
    D, N, kc, m, ksub, nq, k, w = 128, 1_000_000, 1024, 16, 256, 10_000, 10, 16
    X = synth.blobs(N, D, kc, seed=1002)
    Q = synth.blobs(nq, D, kc, seed=2001)
    cent = synth.blob_centres(D, kc)
    _, cb, _ = synth.random_quantizers(kc, D, m, ksub, seed=5, data=X[:200_000])
    qz = orc.Quantizers(cent, cb, None)
    e = iv.IVFADCIndex.from_quantizers(cent, cb, None, index_type=np.uint32)           # engine default (tcgen05 tables)
    x = iv.IVFADCIndex.from_quantizers(cent, cb, None, index_type=np.uint32, flags=LUT_EXACT)
    iv.push_batch(e, X)
    iv.push_batch(x, X)
    assert len(e) == N and int(e.list_sizes().sum()) == N
    gi, gd, gc = e.search_packed(Q, k, w)
    xi, xd, xc = x.search_packed(Q, k, w)
    # shape / order properties (src/index.jl:247-257): k results, ascending, ids in range and unique per query
    assert (gc == k).all() and (xc == k).all()
    assert (np.diff(gd, axis=1) >= 0).all() and (np.diff(xd, axis=1) >= 0).all()
    assert gi.max() < N and all(len(set(r)) == k for r in gi[:500])
    # default tables vs reference-exact tables: the north_star tolerance on every one of the 100k results
    rep = orc.compare_search(gi, gd, gc, xi, xd, xc, rtol=RTOL)
    assert rep["near_tie_id_mismatches"] <= 50, rep
    # checksum of checksums: the exact engine is deterministic across kernels (query-per-lane vs vector-per-lane)
    y = iv.IVFADCIndex.from_quantizers(cent, cb, None, index_type=np.uint32, flags=LEGACY)
    iv.push_batch(y, X)
    yi, yd, yc = y.search_packed(Q[:2000], k, w)
    assert np.array_equal(yi, xi[:2000]) and np.array_equal(yd.view(np.uint8), xd[:2000].view(np.uint8))
    # a sample against the oracle on the exported lists
    sizes = x.list_sizes()
    off = np.zeros(kc + 1, dtype=np.int64)
    np.cumsum(sizes, out=off[1:])
    ids = np.empty(N, dtype=np.uint64)
    codes = np.empty((N, m), dtype=np.uint8)
    for c in range(kc):
        if sizes[c]:
            i, cd = x.export_list(c)
            ids[off[c]:off[c + 1]], codes[off[c]:off[c + 1]] = i, cd
    oi, od, oc, _ = orc.search_csr(qz, off, codes, ids, Q[:128], k, w, nthreads=8)
    assert np.array_equal(xi[:128], oi) and np.array_equal(xd[:128].view(np.uint8), od.view(np.uint8))
    # delete 1000 vectors: none of them may come back, counts stay k
    dele = np.unique(gi[:100].ravel())[:1000]
    iv.delete_from_index(e, (dele + 1).tolist())       # 1-based, as the reference takes them
    assert len(e) == N - len(dele)
    e.close(); x.close(); y.close()


# ------------------------------------------------------------------------------------------------
# Device trainer (not parity-graded: the reference's training is unseeded)
# ------------------------------------------------------------------------------------------------
def test_device_trainer():
    """Lloyd with the engine's coarse kernel as the assignment step: quantisation error on a par with the host
    trainer, assignments consistent with the returned centres; the constructor uses it beyond toy sizes."""
    from ivfadc_jl_b200 import synth, training
    X = synth.blobs(40000, 32, 48, seed=41)
    c_host, a_host = training.kmeans(X, 48, maxiter=15, seed=3)
    c_dev, a_dev = training.kmeans_device(X, 48, maxiter=15, seed=3)
    inertia = lambda c, a: float(((X.astype(np.float64) - c[a].astype(np.float64)) ** 2).sum())
    # two different random streams seed the two trainers (numpy / Philox on the device): on 48 separated blobs a
    # seeding that doubles up in one blob costs ~20 %, so the bar is a band, not equality
    assert inertia(c_dev, a_dev) <= 1.35 * inertia(c_host, a_host)
    d = ((X[:2000, None, :].astype(np.float64) - c_dev[None].astype(np.float64)) ** 2).sum(2)
    assert (d[np.arange(2000), a_dev[:2000]] <= d.min(1) * (1 + 1e-6)).all()   # nearest centre (up to fp32 ties)
    # the library's seeding is counter-based: the same seed gives the same centres; and the quantisation error is on a
    # par with scikit-learn's k-means++ / Lloyd on the same data ("grade by quantisation error vs. sklearn")
    c_dev2, a_dev2 = training.kmeans_device(X, 48, maxiter=15, seed=3)
    assert np.array_equal(a_dev, a_dev2)
    from sklearn.cluster import KMeans
    km = KMeans(n_clusters=48, init="k-means++", n_init=1, max_iter=15, random_state=3, algorithm="lloyd").fit(X.astype(np.float64))
    assert inertia(c_dev, a_dev) <= 1.35 * float(km.inertia_), (inertia(c_dev, a_dev), km.inertia_)
    X64 = X[:8000].astype(np.float64)
    c64, a64 = training.kmeans_device(X64, 16, maxiter=10, seed=5)
    c64h, a64h = training.kmeans(X64, 16, maxiter=10, seed=5)
    i64 = lambda c, a: float(((X64 - c[a]) ** 2).sum())
    assert c64.dtype == np.float64 and i64(c64, a64) <= 1.10 * i64(c64h, a64h)
    # constructor beyond DEVICE_TRAINING_PAIRS: 150 000 vectors x 64 cells
    n, D, kc = 150_000, 32, 64
    Xb = synth.blobs(n, D, kc, seed=42)
    idx = iv.IVFADCIndex(np.ascontiguousarray(Xb.T), kc=kc, k=256, m=8, coarse_maxiter=8, quantization_maxiter=8,
                         index_type=np.uint32)
    assert len(idx) == n
    qs = np.arange(0, n, n // 200)[:200]
    ids, dists, counts = idx.search_packed(Xb[qs], 10, 8)
    hits = sum(int(q in ids[i, :counts[i]]) for i, q in enumerate(qs))
    assert hits >= 150, hits   # a database point finds itself among its 10 nearest (PQ-approximate) neighbours
    idx.close()


# ------------------------------------------------------------------------------------------------
# round 2: wide vectors in K4, device-resident build, bulk persistency, the synthetic generator
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,D,m,ksub,kc", [(np.float32, 960, 8, 256, 20), (np.float64, 512, 8, 256, 12),
                                               (np.float32, 1024, 16, 256, 9), (np.float64, 430, 10, 64, 7)])
def test_encode_wide_vectors(dtype, D, m, ksub, kc):
    """Shapes whose 64 full residual rows exceed a CTA's shared memory (GIST-960 ...): the encode kernel stages one
    subspace slice per codebook instead; same bits as the oracle."""
    rng = np.random.default_rng(D)
    n = 300
    X = rng.random((n, D)).astype(dtype)
    cent = X[rng.choice(n, kc, replace=False)].copy()
    cb = (0.3 * rng.standard_normal((m, ksub, D // m))).astype(dtype)
    qz = orc.Quantizers(cent, cb, None)
    e = engine_from(qz, flags=LEGACY | LUT_EXACT)
    gcell, gcode = e.encode(X)
    ocell, ocode = orc.encode(qz, X, nthreads=4)
    np.testing.assert_array_equal(gcell, ocell)
    np.testing.assert_array_equal(gcode, ocode)
    e.close()


def test_assignments_out_of_range_are_rejected():
    oidx, qz, assign, X, data = helpers.reference_fixture(np.float32, seed=2)
    e = engine_from(qz)
    bad = assign.copy()
    bad[17] = qz.centroids.shape[0]          # one past the last cell
    with pytest.raises(iv._capi.IvfadcError):
        e._add(X, 0, assign=bad, assign_base=0)
    with pytest.raises(iv._capi.IvfadcError):
        e.encode(X, assign=bad - 5, assign_base=0)   # negative cells
    assert len(e) == 0 and int(e.list_sizes().sum()) == 0   # nothing was added
    e._add(X, 0, assign=assign, assign_base=0)
    assert_lists_equal(e, oidx)
    e.close()


def test_add_device_equals_add():
    """ivfadc_add_device (batch, assignments and cells in device memory) builds the same lists as ivfadc_add."""
    import torch
    rng = np.random.default_rng(5)
    D, m, kc, n = 64, 8, 37, 5000
    X = rng.random((n, D)).astype(np.float32)
    cent = X[rng.choice(n, kc, replace=False)].copy()
    cb = (0.2 * rng.standard_normal((m, 256, D // m))).astype(np.float32)
    qz = orc.Quantizers(cent, cb, None)
    a, b, c = engine_from(qz), engine_from(qz), engine_from(qz)
    cells = a._add(X, 0, want_cells=True)
    dX = torch.from_numpy(X).cuda()
    dcells = torch.empty(n, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    b.reserve(n)                                   # capacity hints (sizehint!): even share, then exact list sizes
    c.reserve(n, np.bincount(cells, minlength=kc))
    b.add_device(dX.data_ptr(), n, 0, 0, 0, dcells.data_ptr())
    np.testing.assert_array_equal(dcells.cpu().numpy(), cells)
    dassign = torch.from_numpy(cells.astype(np.int64) + 1).cuda()
    torch.cuda.synchronize()
    c.add_device(dX.data_ptr(), n, 0, dassign.data_ptr(), 1)
    for e in (b, c):
        assert len(e) == n
        np.testing.assert_array_equal(e.list_sizes(), a.list_sizes())
        for cell in range(kc):
            ia, ca = a.export_list(cell)
            ie, ce = e.export_list(cell)
            np.testing.assert_array_equal(ia, ie)
            np.testing.assert_array_equal(ca, ce)
    dassign[3] = kc + 1
    torch.cuda.synchronize()
    with pytest.raises(iv._capi.IvfadcError):
        c.add_device(dX.data_ptr(), n, 0, dassign.data_ptr(), 1)
    for e in (a, b, c):
        e.close()


@pytest.mark.parametrize("id_type,kc", [(np.uint32, 16384), (np.uint64, 300)])
def test_bulk_export_import(tmp_path, id_type, kc):
    """ivfadc_export_all / ivfadc_import_all against the per-list calls at kc = 16 384 (many empty and ragged
    lists), and the saved file byte-identical to the per-list writer's (src/persistency.jl:68-78)."""
    rng = np.random.default_rng(kc)
    D, m, n = 32, 8, 60000
    X = rng.random((n, D)).astype(np.float32)
    cent = rng.random((kc, D)).astype(np.float32)
    cb = (0.2 * rng.standard_normal((m, 256, D // m))).astype(np.float32)
    qz = orc.Quantizers(cent, cb, None)
    e = engine_from(qz, id_type, X)
    sizes, ids, codes = e.export_all()
    assert int(sizes.sum()) == n and ids.dtype == id_type
    o = 0
    for c in range(kc):
        i1, c1 = e.export_list(c)
        assert len(i1) == sizes[c]
        np.testing.assert_array_equal(ids[o:o + len(i1)], i1)
        np.testing.assert_array_equal(codes[o:o + len(i1)], c1)
        o += len(i1)
    # the file: bulk writer == a per-list writer written here after src/persistency.jl:68-78
    fn = str(tmp_path / "bulk.ivfadc")
    iv.save_ivfadc_index(fn, e)
    raw = open(fn, "rb").read()
    per_list = bytearray()
    o = 0
    for c in range(kc):
        per_list += np.int64(sizes[c]).tobytes() + ids[o:o + sizes[c]].tobytes() + codes[o:o + sizes[c]].tobytes()
        o += int(sizes[c])
    assert raw.endswith(bytes(per_list))
    # import into a fresh engine and into one that already holds other lists (everything is replaced)
    e2 = iv.load_ivfadc_index(fn)
    e3 = engine_from(qz, id_type, X[:1000])
    e3.import_all(sizes, ids, codes)
    for t in (e2, e3):
        assert len(t) == n
        s2, i2, c2 = t.export_all()
        np.testing.assert_array_equal(s2, sizes)
        np.testing.assert_array_equal(i2, ids)
        np.testing.assert_array_equal(c2, codes)
    Q = rng.random((50, D)).astype(np.float32)
    r1, r2 = e.search_packed(Q, 5, 8), e3.search_packed(Q, 5, 8)
    for x, y in zip(r1, r2):
        np.testing.assert_array_equal(x, y)
    # mutation after a bulk import keeps working (capacities were rebuilt)
    iv.push_batch(e3, X[:77])
    assert len(e3) == n + 77
    for t in (e, e2, e3):
        t.close()


def test_synth_device_equals_cpu_twin():
    """csrc/synth.cu against oracle_synth_*: the same floats bit for bit, any slice of the stream."""
    import torch
    from ivfadc_jl_b200 import synth
    for D, kb in ((128, 1024), (96, 50), (10, 7)):
        c_cpu = orc.synth_uniform(0, kb, D, 1001)
        c_dev = synth.uniform_device(0, kb, D, 1001)
        assert np.array_equal(c_dev.cpu().numpy().view(np.uint32), c_cpu.view(np.uint32))
        for first, n in ((0, 1000), (123_456_789_012, 777), (99_999_000, 1000)):
            x_cpu, b_cpu = orc.synth_blobs(first, n, D, kb, 1002, synth.blob_scale(0.05), c_cpu)
            x_dev, b_dev = synth.blobs_device(first, n, c_dev, 1002, 0.05, want_blobs=True)
            torch.cuda.synchronize()
            assert np.array_equal(b_dev.cpu().numpy(), b_cpu)
            assert np.array_equal(x_dev.cpu().numpy().view(np.uint32), x_cpu.view(np.uint32))


def test_config_c_full_size_parity():
    """Deep10M-shaped configuration at FULL size (96-d, 10 M vectors, kc = 4096, m = 12) built in HBM from the
    device generator; the default engine against the oracle on 256 queries at the north_star tolerance, the PQ codes
    of a sample bit-identical to the oracle's encoder."""
    import torch
    from ivfadc_jl_b200 import synth
    D, N, kc, m, nq, nchk, k, w = 96, 10_000_000, 4096, 12, 10_000, 256, 10, 16
    centres = synth.uniform_device(0, kc, D, 1001)
    cent = centres.cpu().numpy()
    xs = synth.blobs_device(0, 200_000, centres, 1002)
    tc, tb = synth.train_on_device_tensor(xs, kc, m, 256, iters=2, init=centres)
    cent, cb = tc.cpu().numpy(), tb.cpu().numpy()
    qz = orc.Quantizers(cent, cb, None)
    e = iv.IVFADCIndex.from_quantizers(cent, cb, None)
    buf = torch.empty((1 << 20, D), dtype=torch.float32, device="cuda")
    for s in range(0, N, 1 << 20):
        n = min(1 << 20, N - s)
        x = synth.blobs_device(s, n, centres, 1002, out=buf)
        torch.cuda.synchronize()
        e.add_device(x.data_ptr(), n)
    assert len(e) == N
    # encoding of a slice of the stream: regenerate it on the CPU (counter-based) and compare with the stored codes
    first, ns = 7_654_321, 4096
    xc, _ = orc.synth_blobs(first, ns, D, kc, 1002, synth.blob_scale(0.05), centres.cpu().numpy())
    ocell, ocode = orc.encode(qz, xc, nthreads=8)
    gcell, gcode = e.encode(xc)
    np.testing.assert_array_equal(gcell, ocell)
    np.testing.assert_array_equal(gcode, ocode)
    Q = synth.blobs_device(0, nq, centres, 2001).cpu().numpy()
    sizes, ids, codes = e.export_all()
    off = np.zeros(kc + 1, dtype=np.int64)
    np.cumsum(sizes, out=off[1:])
    # the stored entry of vector `first + j` carries exactly the oracle's code
    pos = {int(v): i for i, v in enumerate(ids[off[ocell[0]]:off[ocell[0] + 1]])}
    assert np.array_equal(codes[off[ocell[0]] + pos[first]], ocode[0])
    oi, od, oc, _ = orc.search_csr(qz, off, codes, ids.astype(np.uint64), Q[:nchk], k, w, nthreads=16)
    gi, gd, gc = e.search_packed(Q, k, w)     # the full batch: ~39 queries per list, the tensor-memory scan
    rep = orc.compare_search(gi[:nchk], gd[:nchk], gc[:nchk], oi, od, oc, rtol=RTOL)
    assert rep["near_tie_id_mismatches"] <= 8, rep
    assert int(e.stats()["last_scan_kernel"]) in (4, 5)
    e.close()


# ------------------------------------------------------------------------------------------------
# f3: the other metrics of Distances.jl for Dc (coarse_search + lookup tables) and Dr (quantize_data)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("dc,dr", [("Euclidean", "SqEuclidean"), ("Cityblock", "Cityblock"), ("CosineDist", "Euclidean"),
                                   ("SqEuclidean", "CosineDist")])
def test_other_metrics_bit_exact(dtype, dc, dr, tmp_path):
    """Euclidean / Cityblock / CosineDist (src/index.jl:108-109): coarse cells and distances, PQ codes, and every
    returned distance bit-identical to the oracle's restatement of the Distances.jl definitions; the saved file
    carries the metric names and reloads to the same results."""
    rng = np.random.default_rng(11)
    D, m, ksub, kc, n, nq, k, w = 24, 6, 64, 40, 3000, 60, 7, 5
    X = (rng.random((n, D)) + 0.1).astype(dtype)
    cent = X[rng.choice(n, kc, replace=False)].copy()
    cb = (0.3 * rng.standard_normal((m, ksub, D // m))).astype(dtype)
    cb[:, -1] = cb[:, 0]   # an exact tie between two codewords: the first one wins
    codes = np.stack([rng.permutation(256)[:ksub].astype(np.uint8) for _ in range(m)])
    qz = orc.Quantizers(cent, cb, codes, coarse_distance=dc, quantization_distance=dr)
    Q = (rng.random((nq, D)) + 0.1).astype(dtype)
    e = iv.IVFADCIndex.from_quantizers(cent, cb, codes, coarse_distance=dc, quantization_distance=dr)
    gcell, gdc = e.coarse_search(Q, w)
    ocell, odc = orc.coarse_search(qz, Q, w, nthreads=2)
    np.testing.assert_array_equal(gcell, ocell)
    assert np.array_equal(gdc.view(np.uint8), odc.view(np.uint8))
    gc, gcode = e.encode(X)
    oc, ocode = orc.encode(qz, X, nthreads=4)
    np.testing.assert_array_equal(gc, oc)
    np.testing.assert_array_equal(gcode, ocode)
    iv.push_batch(e, X)
    order = np.argsort(oc, kind="stable")
    off = np.zeros(kc + 1, dtype=np.int64)
    np.cumsum(np.bincount(oc, minlength=kc), out=off[1:])
    oi, od, ocnt, _ = orc.search_csr(qz, off, ocode[order], order.astype(np.uint64), Q, k, w, nthreads=4)
    for eng in (e,):
        gi, gd, gcnt = eng.search_packed(Q, k, w)
        np.testing.assert_array_equal(gcnt, ocnt)
        assert np.array_equal(gd.view(np.uint8), od.view(np.uint8))
        np.testing.assert_array_equal(gi, oi)
    fn = str(tmp_path / "metric.ivfadc")
    iv.save_ivfadc_index(fn, e)
    lines = open(fn, "rb").read().split(b"\n", 9)
    assert lines[6] == ("Distances." + dc).encode() and lines[7] == ("Distances." + dr).encode()
    e2 = iv.load_ivfadc_index(fn)
    gi2, gd2, gcnt2 = e2.search_packed(Q, k, w)
    assert np.array_equal(gi2, gi) and np.array_equal(gd2.view(np.uint8), gd.view(np.uint8))
    e.close(); e2.close()


def test_config_d_full_size_parity():
    """The 100 M-vector configuration at FULL size on one GPU (128-d, kc = 16384, m = 8, UInt32 ids; codes 800 MB):
    built in HBM from the device generator, 64 queries of a full 10 000-query batch against the oracle over the lists
    they probe (the cost model sends this shape -- 10 queries per list of 6 100 vectors -- to the exact vector-per-lane
    kernel, so ids and distance bits are the oracle's), and the stored codes of a slice of the stream."""
    import torch
    from ivfadc_jl_b200 import synth
    D, N, kc, m, nq, nchk, k, w = 128, 100_000_000, 16384, 8, 10_000, 64, 10, 16
    centres = synth.uniform_device(0, kc, D, 1001)
    xs = synth.blobs_device(0, 1 << 20, centres, 1002)
    tc, tb = synth.train_on_device_tensor(xs, kc, m, 256, iters=1, init=centres)
    del xs
    cent, cb = tc.cpu().numpy(), tb.cpu().numpy()
    qz = orc.Quantizers(cent, cb, None)
    e = iv.IVFADCIndex.from_quantizers(cent, cb, None)
    e.reserve(N)
    buf = torch.empty((1 << 20, D), dtype=torch.float32, device="cuda")
    for s in range(0, N, 1 << 20):
        n = min(1 << 20, N - s)
        x = synth.blobs_device(s, n, centres, 1002, out=buf)
        torch.cuda.synchronize()
        e.add_device(x.data_ptr(), n)
    assert len(e) == N
    Q = synth.blobs_device(0, nq, centres, 2001).cpu().numpy()
    gi, gd, gc = e.search_packed(Q, k, w)
    assert int(e.stats()["last_scan_kernel"]) == 1
    # oracle over exactly the probed lists
    ocell, _ = orc.coarse_search(qz, Q[:nchk], w, nthreads=8)
    sizes = e.list_sizes()
    off = np.zeros(kc + 1, dtype=np.int64)
    need = np.unique(ocell)
    lens = np.zeros(kc, dtype=np.int64)
    lens[need] = sizes[need]
    np.cumsum(lens, out=off[1:])
    ids = np.empty(int(off[-1]), dtype=np.uint64)
    codes = np.empty((int(off[-1]), m), dtype=np.uint8)
    for c in need:
        i, cd = e.export_list(int(c))
        ids[off[c]:off[c + 1]] = i
        codes[off[c]:off[c + 1]] = cd
    oi, od, oc, _ = orc.search_csr(qz, off, codes, ids, Q[:nchk], k, w, nthreads=16)
    np.testing.assert_array_equal(gc[:nchk], oc)
    assert np.array_equal(gd[:nchk].view(np.uint8), od.view(np.uint8))
    np.testing.assert_array_equal(gi[:nchk], oi)
    first, ns = 87_654_321, 2048
    xc, _ = orc.synth_blobs(first, ns, D, kc, 1002, synth.blob_scale(0.05), centres.cpu().numpy())
    ocells, ocodes = orc.encode(qz, xc, nthreads=8)
    gcells, gcodes = e.encode(xc)
    np.testing.assert_array_equal(gcells, ocells)
    np.testing.assert_array_equal(gcodes, ocodes)
    e.close()


def test_float64_large_batch_takes_the_float32_twin():
    """Float64 index (the reference's own tests are Float64), default flags: a large batch is searched by the Float32
    twin (tensor-memory kernel; distances within the north_star tolerance, widened to Float64); a small batch and
    IVFADC_FLAG_LUT_EXACT keep the exact fp64 chain (bit-identical); the twin follows mutations of the lists."""
    from ivfadc_jl_b200 import synth
    D, m, kc, n, nq, k, w = 128, 16, 64, 60000, 1500, 10, 8
    X = synth.blobs(n, D, kc, seed=51, dtype=np.float64)
    Q = synth.blobs(nq, D, kc, seed=52, dtype=np.float64)
    _, cb, codes = synth.random_quantizers(kc, D, m, 256, seed=9, dtype=np.float64, data=X)
    cent = synth.blob_centres(D, kc, dtype=np.float64)
    qz = orc.Quantizers(cent, cb, codes)
    e = iv.IVFADCIndex.from_quantizers(cent, cb, codes)          # default flags
    iv.push_batch(e, X)

    def oracle_of(engine, queries):
        sizes, ids, codes_all = engine.export_all()
        off = np.zeros(kc + 1, dtype=np.int64)
        np.cumsum(sizes, out=off[1:])
        return orc.search_csr(qz, off, codes_all, ids.astype(np.uint64), queries, k, w, nthreads=8)[:3]

    oi, od, oc = oracle_of(e, Q)
    gi, gd, gc = e.search_packed(Q, k, w)
    assert gd.dtype == np.float64 and int(e.stats()["last_scan_kernel"]) in (4, 5)
    rep = orc.compare_search(gi, gd, gc, oi, od, oc, rtol=RTOL)
    assert rep["max_rel_err"] > 0 and rep["near_tie_id_mismatches"] <= 4, rep
    # a small batch stays on the exact chain
    si, sd, sc = e.search_packed(Q[:20], k, w)
    assert int(e.stats()["last_scan_kernel"]) == 1
    assert np.array_equal(sd.view(np.uint8), od[:20].view(np.uint8)) and np.array_equal(si, oi[:20])
    # the twin sees the lists as they are now
    iv.delete_from_index(e, np.arange(1, 5001))
    iv.push_batch(e, X[:777])
    oi2, od2, oc2 = oracle_of(e, Q)
    gi2, gd2, gc2 = e.search_packed(Q, k, w)
    orc.compare_search(gi2, gd2, gc2, oi2, od2, oc2, rtol=RTOL)
    assert not np.array_equal(gi2, gi)
    e.close()
    ex = iv.IVFADCIndex.from_quantizers(cent, cb, codes, flags=LUT_EXACT)
    iv.push_batch(ex, X)
    xi, xd, xc = ex.search_packed(Q, k, w)
    assert np.array_equal(xd.view(np.uint8), od.view(np.uint8)) and np.array_equal(xi, oi)
    ex.close()


def test_device_group_equals_one_gpu():
    """ivfadc_group_* (one process, several GPUs: what the Julia glue binds with IVFADC_DEVICES): build, search, delete,
    pop and length against the single-GPU engine, bit for bit.  Needs two devices (skipped on a one-GPU box)."""
    from ivfadc_jl_b200 import sharded, synth
    if iv._capi.load().ivfadc_device_count() < 2:
        pytest.skip("needs two GPUs")
    D, m, kc, n, nq, k, w = 64, 8, 48, 40000, 700, 10, 8
    X = synth.blobs(n, D, kc, seed=61)
    Q = synth.blobs(nq, D, kc, seed=62)
    cent, cb, codes = synth.random_quantizers(kc, D, m, 256, seed=10, data=X)
    one = iv.IVFADCIndex.from_quantizers(cent, cb, codes)
    cells = one._add(X, iv._capi.LAST, want_cells=True)
    g = sharded.DeviceGroup(cent, cb, codes, n_devices=2)
    g.set_cell_owners(sharded.balanced_owners(np.bincount(cells, minlength=kc), 2))
    g.add(X, assign=cells.astype(np.int64), assign_base=0)
    assert len(g) == len(one) == n
    ss = g.shard_sizes()
    np.testing.assert_array_equal(ss.sum(0), one.list_sizes())
    assert abs(int(ss[0].sum()) - int(ss[1].sum())) < 0.05 * n      # balanced by list length
    for kk, ww in ((k, w), (3, 1), (16, 48)):
        ui, ud, uc = one.search_packed(Q, kk, ww)
        gi, gd, gc = g.search(Q, kk, ww)
        np.testing.assert_array_equal(gc, uc)
        np.testing.assert_array_equal(gi, ui)
        assert np.array_equal(gd.view(np.uint8), ud.view(np.uint8))
    dele = np.random.default_rng(3).choice(n, 3000, replace=False)
    iv.delete_from_index(one, dele + 1)
    g.delete(np.sort(dele).astype(np.uint64))
    assert len(g) == len(one) == n - 3000
    v1, v2 = iv.pop(one), g.pop()
    np.testing.assert_array_equal(v1, v2)
    v1, v2 = iv.popfirst(one), g.pop(iv._capi.FIRST)
    np.testing.assert_array_equal(v1, v2)
    g.add(X[:500])
    iv.push_batch(one, X[:500])
    ui, ud, uc = one.search_packed(Q, k, w)
    gi, gd, gc = g.search(Q, k, w)
    np.testing.assert_array_equal(gi, ui)
    assert np.array_equal(gd.view(np.uint8), ud.view(np.uint8))
    g.close(); one.close()
