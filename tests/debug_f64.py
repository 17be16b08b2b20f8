"""Bring-up: Float64 index, 10 000-query batch: default flags (Float32 twin on the tensor-memory kernel) against
IVFADC_FLAG_LUT_EXACT (exact fp64 chain on the vector-per-lane kernel)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ivfadc_jl_b200 as iv
from ivfadc_jl_b200 import synth
D, m, kc, n, nq, k, w = 128, 16, 512, 500_000, 10_000, 10, 16
X = synth.blobs(n, D, kc, seed=51, dtype=np.float64)
Q = synth.blobs(nq, D, kc, seed=52, dtype=np.float64)
_, cb, codes = synth.random_quantizers(kc, D, m, 256, seed=9, dtype=np.float64, data=X)
cent = synth.blob_centres(D, kc, dtype=np.float64)
for name, flags in (("default (Float32 twin)", 0), ("LUT_EXACT (fp64 chain)", 4)):
    e = iv.IVFADCIndex.from_quantizers(cent, cb, codes, flags=flags)
    iv.push_batch(e, X)
    for _ in range(2):
        e.search_packed(Q, k, w)
    e.stats(reset=True)
    t = time.perf_counter()
    for _ in range(5):
        ids, d, c = e.search_packed(Q, k, w)
    dt = (time.perf_counter() - t) / 5
    st = e.stats()
    print(f"{name}: {1e3 * dt:.3f} ms per batch through the host API ({nq / dt / 1e6:.2f} M QPS), scan {st['scan_ms'] / 5:.3f} ms, "
          f"kernel {st['last_scan_kernel']}, d[0,:3] = {d[0, :3]}", flush=True)
    e.close()
