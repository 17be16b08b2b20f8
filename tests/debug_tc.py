"""Bring-up diagnostics for the tcgen05 table builder (run on the GPU box; not a pytest test)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ivfadc_jl_b200 as iv
from ivfadc_jl_b200 import synth
from oracle import oracle as orc

D, m, ksub, kc, n, nq, w, k = 128, 16, 256, 8, 4000, 64, 2, 5
X = synth.blobs(n, D, kc, seed=31)
cent, cb, codes = synth.random_quantizers(kc, D, m, ksub, seed=7, data=X)
qz = orc.Quantizers(cent, cb, codes)
Q = synth.blobs(nq, D, kc, seed=32)
e = iv.IVFADCIndex.from_quantizers(cent, cb, codes, flags=2)
e._add(X, 0)
iv._capi.check(e._h, e._lib.ivfadc_debug_tables(e._h, None))
try:
    gi, gd, gc = e.search_packed(Q, k, w)
except Exception as ex:
    print("SEARCH FAILED:", ex)
    gi = None
buf = np.zeros(m * 256 * 32 + 64 + 1024, dtype=np.float32)
iv._capi.check(e._h, e._lib.ivfadc_debug_tables(e._h, buf.ctypes.data_as(ctypes.c_void_p)))
tables = buf[:m * 256 * 32].reshape(m, 256, 32)
meta = buf[m * 256 * 32:].view(np.int32)
pairs, cell = meta[:32], int(meta[32])
print("cell", cell, "pairs", pairs.tolist())
dsub = D // m
for s in range(m):
    errs = []
    for slot, p in enumerate(pairs):
        if p < 0:
            continue
        r = Q[p // w].astype(np.float64) - cent[cell].astype(np.float64)
        rs = r[s * dsub:(s + 1) * dsub]
        wv = cb[s].astype(np.float64)
        want = (wv * wv).sum(1) - 2.0 * wv @ rs
        got = tables[s, :, slot].astype(np.float64)
        errs.append(np.max(np.abs(got - want)))
        if s < 2 and slot < 2:
            print(f" s={s} slot={slot} want[:4]={want[:4]} got[:4]={got[:4]}")
    print(f"subspace {s}: max abs err {max(errs):.3e}  (table magnitude {np.abs(tables[s]).max():.3e})")
if gi is not None:
    cells, ocodes = orc.encode(qz, X, nthreads=4)
    order = np.argsort(cells, kind="stable")
    off = np.zeros(kc + 1, dtype=np.int64)
    np.cumsum(np.bincount(cells, minlength=kc), out=off[1:])
    oi, od, oc, _ = orc.search_csr(qz, off, ocodes[order], order.astype(np.uint64), Q, k, w, nthreads=4)
    try:
        print(orc.compare_search(gi, gd, gc, oi, od, oc, rtol=1e-5))
    except AssertionError as ex:
        print("COMPARE FAILED:", str(ex)[:500])
        print("gd[0]", gd[0], "od[0]", od[0])
e.close()
