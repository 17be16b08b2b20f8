"""Bring-up: per-chunk time of ivfadc_add_device against ivfadc_add on workload E's shape."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ivfadc_jl_b200 as iv
from ivfadc_jl_b200 import synth
D, kc, m, ksub, CH = 128, 1024, 16, 256, 1 << 20
centres = synth.uniform_device(0, kc, D, 1001)
xs = synth.blobs_device(0, 262144, centres, 1002)
tc, tb = synth.train_on_device_tensor(xs, kc, m, ksub, iters=4, init=centres)
cent, cb = tc.cpu().numpy(), tb.cpu().numpy()
buf = torch.empty((CH, D), dtype=torch.float32, device="cuda")
for mode in ("device", "host", "device-same-chunk"):
    e = iv.IVFADCIndex.from_quantizers(cent, cb, None)
    ts = []
    xh = synth.blobs_device(0, CH, centres, 1002).cpu().pin_memory().numpy()
    for i in range(8):
        x = synth.blobs_device((0 if mode == "device-same-chunk" else i) * CH, CH, centres, 1002, out=buf)
        torch.cuda.synchronize()
        t = time.perf_counter()
        if mode == "host":
            iv.push_batch(e, xh)
        else:
            e.add_device(x.data_ptr(), CH)
        ts.append(1e3 * (time.perf_counter() - t))
    print(mode, " ".join("%.1f" % v for v in ts), "ms; launches", e.stats()["gpu_launches"], flush=True)
    e.close()
