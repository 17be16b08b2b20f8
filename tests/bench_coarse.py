"""Bring-up / measurement: the coarse step alone (K1) at the centroid counts of configs B, C and D, tensor-core
pruning + exact re-rank (default) against the packed-FP32 kernel (IVFADC_FLAG_COARSE_FFMA), device-resident
queries, CUDA events, L2 flushed between repetitions.  Prints one JSON line per shape."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ivfadc_jl_b200 as iv
from ivfadc_jl_b200 import synth, sharded, _capi

nq, reps = 10000, 10
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, kc, D, w in (("B", 1024, 128, 16), ("C", 4096, 96, 16), ("D", 16384, 128, 16)):
    cent = synth.blob_centres(D, kc)
    Q = synth.blobs(nq, D, kc, seed=2001)
    cb = np.zeros((D // 8, 256, 8), dtype=np.float32)
    dQ = torch.from_numpy(Q).cuda()
    out = {"shape": name, "kc": kc, "D": D, "w": w, "nq": nq}
    res = {}
    for label, flags in (("tensor_core_ms", 0), ("ffma_ms", _capi.FLAG_COARSE_FFMA)):
        e = iv.IVFADCIndex.from_quantizers(cent, cb, None, flags=flags)
        for _ in range(3):
            c, d = sharded.coarse_device(e, dQ, w)
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            c, d = sharded.coarse_device(e, dQ, w)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        out[label] = round(sum(ts) / len(ts), 4)
        res[label] = (c.cpu().numpy(), d.cpu().numpy())
        e.close()
    a, b = res["tensor_core_ms"], res["ffma_ms"]
    out["identical_bits"] = bool(np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint8), b[1].view(np.uint8)))
    out["gflop_direct_form"] = round(3e-9 * nq * kc * D, 1)
    print(json.dumps(out))
