"""Writes the bundle julia/validate_against_reference.jl consumes (run on a GPU box):
an index saved by this engine in the reference's on-disk format + queries + the engine's knn_search results in the
exact-table mode.  README-shaped data (config A: 50-d Float32, 1000 vectors, kc = 100, k = 256, m = 10, UInt16 ids,
reference README.md:31-46) plus a w = 4 batch.  Usage: python tests/golden/make_validation_bundle.py [outdir]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

import ivfadc_jl_b200 as iv
from ivfadc_jl_b200 import _capi, persistency, synth

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "validation")
os.makedirs(out, exist_ok=True)
D, N, kc, m, ksub, nq, k, w = 50, 1000, 100, 10, 256, 64, 3, 4
X = synth.uniform(N, D, seed=1003)
Q = synth.uniform(nq, D, seed=2003)
cent, cb, codes = synth.random_quantizers(kc, D, m, ksub, seed=5, data=X, resid_scale=0.15)
e = iv.IVFADCIndex.from_quantizers(cent, cb, codes, index_type=np.uint16, flags=_capi.FLAG_LUT_EXACT)
iv.push_batch(e, X)
persistency.save_ivfadc_index(os.path.join(out, "index.ivfadc"), e)
ids, dists, counts = e.search_packed(Q, k, w)
with open(os.path.join(out, "queries.bin"), "wb") as f:
    f.write(f"{nq} {D} {k} {w}\n".encode())
    f.write(Q.astype("<f4").tobytes())
    f.write(counts.astype("<i4").tobytes())
    f.write(ids.astype("<u8").tobytes())
    f.write(dists.astype("<f4").tobytes())
print("bundle written to", out, "index bytes", os.path.getsize(os.path.join(out, "index.ivfadc")))
e.close()
