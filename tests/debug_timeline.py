"""Bring-up: phase timeline (SM clocks) of one work item of the tcgen05 scan kernel on config B."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ivfadc_jl_b200 as iv
from ivfadc_jl_b200 import synth
D, N, kc, m, ksub, nq, k, w = 128, 1_000_000, 1024, 16, 256, 10_000, 10, 16
X = synth.blobs(N, D, kc, seed=1002); Q = synth.blobs(nq, D, kc, seed=2001)
cent = synth.blob_centres(D, kc)
_, cb, codes = synth.random_quantizers(kc, D, m, ksub, seed=5, data=X[:100000])
e = iv.IVFADCIndex.from_quantizers(cent, cb, None, flags=int(sys.argv[1]) if len(sys.argv) > 1 else 0)
iv.push_batch(e, X)
for _ in range(2): e.search_packed(Q, k, w)
iv._capi.check(e._h, e._lib.ivfadc_debug_tables(e._h, None))
e.search_packed(Q, k, w)
buf = np.zeros(m * 256 * 32 + 64 + 1024, dtype=np.float32)
iv._capi.check(e._h, e._lib.ivfadc_debug_tables(e._h, buf.ctypes.data_as(ctypes.c_void_p)))
allts = buf[m * 256 * 32 + 64:].view(np.int64)
its = allts[32:112]; its = its[its != 0]
fs = allts[240:440]; fs = fs[fs != 0]
if len(fs):
    print("issuer lane fine stamps (per build: refill[3] at top, then issue[4]) deltas:", np.diff(fs).tolist())
ws = allts[112:112 + 128].reshape(16, 8)
if ws.any():
    w0 = ws[ws != 0].min()
    print("per-warp stamps, subspaces 5 and 6 (before wait, after wait, scan done, past barrier) relative to the earliest:")
    for w in range(16):
        print("  warp %2d" % w, (ws[w] - w0).tolist())
if len(its):
    print("warp stamps relative:", (its - its[0]).tolist())
    d = np.diff(its)
    print("as rows of 4 deltas (wait, scan, sync, loop-back):", d[:4 * (len(d) // 4)].reshape(-1, 4).tolist())
ts = allts[:32]
ts = ts[ts != 0]
print("stamps:", len(ts)); d = np.diff(ts); print("deltas:", d.tolist()); print("total", int(ts[-1] - ts[0]))
e.close()
