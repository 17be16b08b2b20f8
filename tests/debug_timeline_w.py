"""Bring-up: role timeline (SM clocks) of one segment of the warp-specialised scan kernel on config B.
Stamp layout: scanw_impl.cuh (W_DBG_SEG)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
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
ts = buf[m * 256 * 32 + 64:].view(np.int64)
print("stats", e.stats())
sc = ts[:192].reshape(12, 16)
t0 = sc[:, 0][sc[:, 0] != 0].min() if (sc[:, 0] != 0).any() else 0
names = ["start", "tables_done", "bar1", "bar2", "extract_done"]
print("scanner stamps relative to the earliest segment start:", names, "| s=5: before wait, after wait, scan done | s=6: same")
for wv in range(12):
    r = sc[wv]
    print("  warp %2d" % wv, [int(x - t0) if x else None for x in r[:5]], [int(x - t0) if x else None for x in r[6:12]])
pr = ts[192:256].reshape(16, 4)
ow = ts[272:304].reshape(16, 2)
print("issuer per table (A operand seen + refill issued, codebook operand seen, table buffer released, MMAs issued) | operand writer (ring slot free, A written), relative:")
for s in range(16):
    print("  s=%2d" % s, [int(x - t0) if x else None for x in pr[s]], [int(x - t0) if x else None for x in ow[s]])
print("loader (segment after): start, finalize of the segment two back done, desc, copies landed, planes, staged:",
      [int(x - t0) if x else None for x in (ts[256], ts[261], ts[257], ts[258], ts[259], ts[260])])
e.close()
