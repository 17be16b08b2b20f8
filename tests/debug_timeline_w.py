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
pr = ts[:128].reshape(16, 8).copy()
names = {
    "scanner": ["wait staged", "wait table", "lookups", "release", "minima+rank+bars", "-", "candidates", "-"],
    "issuer": ["wait staged", "wait A", "wait codebook", "wait release", "refill + descriptors + issue", "-", "-", "-"],
    "loader": ["finalize", "descriptor", "copies", "byte planes", "residuals+post", "-", "-", "-"],
    "writer": ["wait staged", "convert", "wait ring slot", "write", "-", "-", "-", "-"],
}
role = ["scanner"] * 12 + ["issuer", "loader", "loader", "writer"]
print("phase profile of CTA 0 (SM clocks summed over the launch):")
for wv in range(12):
    if pr[wv][5]:
        print("  warp %2d waited > 300 clocks for %d tables; build complete -> warp running again %.0f clocks on average"
              % (wv, pr[wv][5], pr[wv][7] / pr[wv][5]))
    pr[wv][5] = pr[wv][7] = 0
if pr[12][7]:
    print("  issuer: last release -> build issued %.0f clocks, issued -> complete %.0f clocks on average over %d builds"
          % (pr[12][5] / pr[12][7], pr[12][6] / pr[12][7], pr[12][7]))
pr[12][5] = pr[12][6] = pr[12][7] = 0
for wv in range(16):
    tot = int(pr[wv].sum()) or 1
    print("  warp %2d %-8s total %9d | " % (wv, role[wv], tot) + ", ".join(
        "%s %4.1f%%" % (n, 100.0 * int(v) / tot) for n, v in zip(names[role[wv]], pr[wv]) if n != "-"))
tr = ts[128:128 + 256].view(np.uint32).reshape(16, 4, 8).astype(np.int64)
if tr.any():
    t0 = tr[tr != 0].min()
    rel = np.where(tr != 0, tr - t0, -1)
    print("trace of four consecutive tables (clocks relative to the first event):")
    print("  scanners: before wait, table seen, lookups done, released")
    for wv in range(12):
        print("   warp %2d " % wv + " | ".join(" ".join("%5d" % x for x in rel[wv, k, :4]) for k in range(4)))
    print("  issuer: loop top, A seen, refill + descriptors done, codebook seen, release seen, issued")
    print("           " + " | ".join(" ".join("%5d" % x for x in rel[12, k, :6]) for k in range(4)))
    print("  writer: loop top, converted, ring slot free, written")
    print("           " + " | ".join(" ".join("%5d" % x for x in rel[15, k, :4]) for k in range(4)))
e.close()
