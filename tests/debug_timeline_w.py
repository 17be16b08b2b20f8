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
pr = ts[:128].reshape(16, 8)
names = {
    "scanner": ["wait staged", "wait table", "lookups", "release", "minima+bar", "rank+bar", "candidates", "-"],
    "issuer": ["wait staged", "wait A (+refill)", "wait codebook", "wait release", "issue", "-", "-", "-"],
    "loader": ["finalize", "descriptor", "copies", "byte planes", "residuals+post", "-", "-", "-"],
    "writer": ["wait staged", "convert", "wait ring slot", "write", "-", "-", "-", "-"],
}
role = ["scanner"] * 12 + ["issuer", "loader", "loader", "writer"]
print("phase profile of CTA 0 (SM clocks summed over the launch):")
for wv in range(16):
    tot = int(pr[wv].sum()) or 1
    print("  warp %2d %-8s total %9d | " % (wv, role[wv], tot) + ", ".join(
        "%s %4.1f%%" % (n, 100.0 * int(v) / tot) for n, v in zip(names[role[wv]], pr[wv]) if n != "-"))
e.close()
