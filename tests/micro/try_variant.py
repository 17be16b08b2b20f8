"""Bring-up: one small search on a variant library (IVFADC_LIB=path), checked against the oracle; meant to run under
`timeout -s KILL` so that a deadlocked variant costs seconds, not the call's whole limit."""
import os, sys, shutil
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, root)
if os.environ.get("IVFADC_LIB"):
    shutil.copy(os.environ["IVFADC_LIB"], os.path.join(root, "ivfadc.jl_b200", "libivfadc_cuda.so"))
import numpy as np
import ivfadc_jl_b200 as iv
from ivfadc_jl_b200 import synth
from oracle import oracle as orc
D, n, kc, m, nq, k, w = 64, 262144, 256, 8, 4096, 10, 8
X = synth.blobs(n, D, kc, seed=1)
cent, cb, codes = synth.random_quantizers(kc, D, m, 256, seed=2, data=X)
cent = synth.blob_centres(D, kc)
e = iv.IVFADCIndex.from_quantizers(cent, cb, codes)
iv.push_batch(e, X)
print("built", flush=True)
ids, dists, counts = e.search_packed(Q := synth.blobs(nq, D, kc, seed=3), k, w)
print("searched, kernel", e.stats()["last_scan_kernel"], flush=True)
sizes, ia, ca = e.export_all()
off = np.zeros(kc + 1, dtype=np.int64); np.cumsum(sizes, out=off[1:])
oi, od, oc, _ = orc.search_csr(orc.Quantizers(cent, cb, codes), off, ca, ia.astype(np.uint64), Q, k, w, nthreads=8)
print(orc.compare_search(ids, dists, counts, oi, od, oc, rtol=1e-5), flush=True)
