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
buf = np.zeros(m * 256 * 32 + 64 + 256, dtype=np.float32)
iv._capi.check(e._h, e._lib.ivfadc_debug_tables(e._h, buf.ctypes.data_as(ctypes.c_void_p)))
ts = buf[m * 256 * 32 + 64:].view(np.int64)
ts = ts[ts != 0]
print("stamps:", len(ts)); d = np.diff(ts); print("deltas:", d.tolist()); print("total", int(ts[-1] - ts[0]))
e.close()
