"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle).

CPU: the oracle must reproduce its own pinned outputs bit for bit (guards the checker against
silent drift -- the reference itself ships no golden vectors).
GPU: the CUDA path, through the C ABI, must reproduce the same fixtures bit for bit
(reference-exact tables) and within 1e-5 relative / ids equal except near-ties (default tables).
"""
import os

import numpy as np
import pytest

from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = ["readme_f32", "reftest_f64"]


def load(name):
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    qz = orc.Quantizers(z["centroids"], z["cb_vectors"], z["cb_codes"])
    return z, qz


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_reproduces_golden(name):
    z, qz = load(name)
    X, Q, assign = z["X"], z["Q"], z["assign"]
    cells, codes = orc.encode(qz, X, nthreads=2)
    np.testing.assert_array_equal(cells, z["enc_cells"])
    np.testing.assert_array_equal(codes, z["enc_codes"])
    _, bcodes = orc.encode(qz, X, assign=assign, assign_base=0, nthreads=2)
    np.testing.assert_array_equal(bcodes, z["build_codes"])
    ccells, cdc = orc.coarse_search(qz, Q, z["coarse_cells"].shape[1], nthreads=2)
    np.testing.assert_array_equal(ccells, z["coarse_cells"])
    assert np.array_equal(cdc.view(np.uint8), z["coarse_dc"].view(np.uint8))
    oidx = orc.OracleIndex(qz, id_bytes=4).build(X, assign, assign_base=0)
    for kk, w in z["searches"]:
        oi, od, oc = oidx.knn_search(Q, int(kk), w=int(w), nthreads=2)
        np.testing.assert_array_equal(oc, z[f"counts_k{kk}_w{w}"])
        np.testing.assert_array_equal(oi, z[f"ids_k{kk}_w{w}"])
        assert np.array_equal(od.view(np.uint8), z[f"dists_k{kk}_w{w}"].view(np.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("name", FIXTURES)
@pytest.mark.parametrize("flags", [4, 0, 1])   # reference-exact tables | engine default | vector-per-lane kernel
def test_cuda_reproduces_golden(name, flags):
    import ivfadc_jl_b200 as iv
    z, qz = load(name)
    X, Q, assign = z["X"], z["Q"], z["assign"]
    e = iv.IVFADCIndex.from_quantizers(qz.centroids, qz.cb_vectors, qz.cb_codes, index_type=np.uint32, flags=flags)
    gcell, gcode = e.encode(X)
    np.testing.assert_array_equal(gcell, z["enc_cells"])
    np.testing.assert_array_equal(gcode, z["enc_codes"])
    w8 = z["coarse_cells"].shape[1]
    ccells, cdc = e.coarse_search(Q, w8)
    np.testing.assert_array_equal(ccells, z["coarse_cells"])
    assert np.array_equal(cdc.view(np.uint8), z["coarse_dc"].view(np.uint8))
    e._add(X, 0, assign=assign, assign_base=0)
    for kk, w in z["searches"]:
        gi, gd, gc = e.search_packed(Q, int(kk), int(w))
        oi, od, oc = z[f"ids_k{kk}_w{w}"], z[f"dists_k{kk}_w{w}"], z[f"counts_k{kk}_w{w}"]
        if flags in (4, 1) or X.dtype == np.float64:
            np.testing.assert_array_equal(gc, oc)
            np.testing.assert_array_equal(gi, oi)
            assert np.array_equal(gd.view(np.uint8), od.view(np.uint8))
        else:
            orc.compare_search(gi, gd, gc, oi, od, oc, rtol=1e-5)
    e.close()


# ---- synthetic generator (bench / scale tests): Philox4x32-10 known answers and the statistics of the streams ----
def test_philox_known_answers():
    """Random123 known-answer vectors of Philox4x32-10 (kat_vectors: zero, all-ones and the pi digits)."""
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kat:
        assert [int(x) for x in orc.philox4x32_10(ctr, key)] == want


def test_synth_streams_are_counter_based():
    from ivfadc_jl_b200 import synth
    D, kb = 24, 13
    c = orc.synth_uniform(0, kb, D, 1001)
    assert c.min() >= 0.0 and c.max() < 1.0 and abs(c.mean() - 0.5) < 0.05
    scale = synth.blob_scale(0.05)
    X, b = orc.synth_blobs(0, 50000, D, kb, 1002, scale, c)
    r = X - c[b]
    assert abs(r.std() - 0.05) < 1e-3 and abs(r.mean()) < 1e-3
    assert b.min() == 0 and b.max() == kb - 1
    assert np.bincount(b, minlength=kb).min() > 50000 / kb * 0.9
    # any slice equals the same rows of the whole stream; other seeds give other data
    X2, b2 = orc.synth_blobs(1234, 100, D, kb, 1002, scale, c)
    assert np.array_equal(X2, X[1234:1334]) and np.array_equal(b2, b[1234:1334])
    X3, _ = orc.synth_blobs(0, 100, D, kb, 1003, scale, c)
    assert not np.array_equal(X3, X[:100])


def test_validation_bundle_against_oracle():
    """tests/golden/validation/ (written on a B200 by make_validation_bundle.py for julia/validate_against_reference.jl):
    the index file this engine saved in the reference's format, parsed here independently after src/persistency.jl,
    searched by the oracle -- the engine's recorded exact-mode results must be the oracle's, bit for bit."""
    d = os.path.join(HERE, "golden", "validation")
    raw = open(os.path.join(d, "index.ivfadc"), "rb").read()
    lines = raw.split(b"\n", 9)
    nrows, kc = (int(x) for x in lines[0].split())
    n, m, k, dsub = (int(x) for x in lines[1].split())
    assert lines[2] == b"NaiveQuantizer" and lines[4] == b"UInt8" and lines[5] == b"UInt16" and lines[8] == b"Float32"
    body = np.frombuffer(lines[9], dtype=np.uint8)
    p = 4 * nrows * kc
    cent = body[:p].view(np.float32).reshape(kc, nrows)
    cbv = np.empty((m, k, dsub), dtype=np.float32)
    cbc = np.empty((m, k), dtype=np.uint8)
    for i in range(m):
        cbc[i] = body[p:p + k]; p += k
        cbv[i] = body[p:p + 4 * k * dsub].view(np.float32).reshape(dsub, k).T; p += 4 * k * dsub
    assert np.array_equal(body[p:p + 4 * nrows * nrows].view(np.float32).reshape(nrows, nrows), np.eye(nrows, dtype=np.float32))
    p += 4 * nrows * nrows
    offsets = np.zeros(kc + 1, dtype=np.int64)
    ids, codes = [], []
    for c in range(kc):
        ln = int(body[p:p + 8].view(np.int64)[0]); p += 8
        ids.append(body[p:p + 2 * ln].view(np.uint16).astype(np.uint64)); p += 2 * ln
        codes.append(body[p:p + m * ln].reshape(ln, m)); p += m * ln
        offsets[c + 1] = offsets[c] + ln
    assert p == len(body) and offsets[-1] == n
    with open(os.path.join(d, "queries.bin"), "rb") as f:
        nq, D, kk, w = (int(x) for x in f.readline().split())
        Q = np.frombuffer(f.read(4 * nq * D), dtype="<f4").reshape(nq, D)
        gc = np.frombuffer(f.read(4 * nq), dtype="<i4")
        gi = np.frombuffer(f.read(8 * nq * kk), dtype="<u8").reshape(nq, kk)
        gd = np.frombuffer(f.read(4 * nq * kk), dtype="<f4").reshape(nq, kk)
    qz = orc.Quantizers(np.ascontiguousarray(cent), cbv, cbc)
    oi, od, oc, _ = orc.search_csr(qz, offsets, np.concatenate(codes), np.concatenate(ids), np.ascontiguousarray(Q), kk, w, nthreads=2)
    assert np.array_equal(gc, oc)
    for j in range(nq):
        assert np.array_equal(gi[j, :gc[j]], oi[j, :oc[j]])
        assert np.array_equal(gd[j, :gc[j]].view(np.uint32), od[j, :oc[j]].view(np.uint32))


# ---- the other metrics (Euclidean / Cityblock / CosineDist): fixtures of the oracle's restatement -----------------
METRIC_PAIRS = [("Euclidean", "SqEuclidean"), ("Cityblock", "Cityblock"), ("CosineDist", "Euclidean"), ("SqEuclidean", "CosineDist")]


def _metric_case(z, dc, dr):
    qz = orc.Quantizers(z["centroids"], z["cb_vectors"], z["cb_codes"], coarse_distance=dc, quantization_distance=dr)
    t = f"{dc}_{dr}"
    return qz, t


@pytest.mark.parametrize("name", ["metrics_f32", "metrics_f64"])
@pytest.mark.parametrize("dc,dr", METRIC_PAIRS)
def test_oracle_reproduces_metric_golden(name, dc, dr):
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    qz, t = _metric_case(z, dc, dr)
    X, Q, k, w = z["X"], z["Q"], int(z["k"]), int(z["w"])
    cells, pq = orc.encode(qz, X, nthreads=2)
    np.testing.assert_array_equal(cells, z[f"cells_{t}"])
    np.testing.assert_array_equal(pq, z[f"codes_{t}"])
    ccells, cdc = orc.coarse_search(qz, Q, w, nthreads=2)
    np.testing.assert_array_equal(ccells, z[f"ccells_{t}"])
    assert np.array_equal(cdc.view(np.uint8), z[f"cdc_{t}"].view(np.uint8))
    order = np.argsort(cells, kind="stable")
    off = np.zeros(qz.kc + 1, dtype=np.int64)
    np.cumsum(np.bincount(cells, minlength=qz.kc), out=off[1:])
    oi, od, oc, _ = orc.search_csr(qz, off, pq[order], order.astype(np.uint64), Q, k, w, nthreads=2)
    np.testing.assert_array_equal(oc, z[f"counts_{t}"])
    np.testing.assert_array_equal(oi, z[f"ids_{t}"])
    assert np.array_equal(od.view(np.uint8), z[f"dists_{t}"].view(np.uint8))
    # and against the plain numpy float64 definition of the metric (loose: different summation order / precision)
    c, q = z["centroids"].astype(np.float64), Q[0].astype(np.float64)
    ref = {"SqEuclidean": ((c - q) ** 2).sum(1), "Euclidean": np.sqrt(((c - q) ** 2).sum(1)), "Cityblock": np.abs(c - q).sum(1),
           "CosineDist": 1 - (c @ q) / np.sqrt((c * c).sum(1) * (q @ q))}[dc]
    np.testing.assert_allclose(cdc[0], np.sort(ref)[:w], rtol=1e-4 if z["X"].dtype == np.float32 else 1e-11, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["metrics_f32", "metrics_f64"])
@pytest.mark.parametrize("dc,dr", METRIC_PAIRS)
def test_cuda_reproduces_metric_golden(name, dc, dr):
    import ivfadc_jl_b200 as iv
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    qz, t = _metric_case(z, dc, dr)
    X, Q, k, w = z["X"], z["Q"], int(z["k"]), int(z["w"])
    e = iv.IVFADCIndex.from_quantizers(qz.centroids, qz.cb_vectors, qz.cb_codes, coarse_distance=dc, quantization_distance=dr)
    gcell, gcode = e.encode(X)
    np.testing.assert_array_equal(gcell, z[f"cells_{t}"])
    np.testing.assert_array_equal(gcode, z[f"codes_{t}"])
    ccells, cdc = e.coarse_search(Q, w)
    np.testing.assert_array_equal(ccells, z[f"ccells_{t}"])
    assert np.array_equal(cdc.view(np.uint8), z[f"cdc_{t}"].view(np.uint8))
    iv.push_batch(e, X)
    gi, gd, gc = e.search_packed(Q, k, w)
    np.testing.assert_array_equal(gc, z[f"counts_{t}"])
    np.testing.assert_array_equal(gi, z[f"ids_{t}"])
    assert np.array_equal(gd.view(np.uint8), z[f"dists_{t}"].view(np.uint8))
    e.close()
