"""Bring-up: one coarse search on config-B shapes (prints the phase stamps of a -DC3_STAMP build)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ivfadc_jl_b200 as iv
from ivfadc_jl_b200 import synth
D, kc, nq, w = 128, 1024, 10000, 16
cent = synth.blob_centres(D, kc)
Q = synth.blobs(nq, D, kc, seed=2001)
cb = np.zeros((16, 256, 8), dtype=np.float32)
e = iv.IVFADCIndex.from_quantizers(cent, cb, None)
for _ in range(3):
    c, d = e.coarse_search(Q, w)
print("cells", c[:2].tolist())
e.close()
