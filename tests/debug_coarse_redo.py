"""Bring-up: coarse step of workload B on the bench's data (device generator) and on the round-1 data (numpy blobs):
time per step and the number of queries the exact redo pass had to take."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ivfadc_jl_b200 as iv
from ivfadc_jl_b200 import sharded, synth
D, kc, m, ksub, nq, w = 128, 1024, 16, 256, 10000, 16
dev = torch.device("cuda", 0)

def run(tag, cent, cb, dQ):
    e = iv.IVFADCIndex.from_quantizers(cent, cb, None)
    for _ in range(3):
        sharded.coarse_device(e, dQ, w)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        c, dc = sharded.coarse_device(e, dQ, w)
    b.record(); torch.cuda.synchronize()
    st = e.stats()
    cn = (cent.astype(np.float64) ** 2).sum(1)
    dd = np.sort(((cent[:, None, :].astype(np.float64) - cent[None, :64, :]) ** 2).sum(-1), axis=0)[1]
    print(f"{tag}: coarse {a.elapsed_time(b) / 10:.4f} ms/step, redo queries {st['last_coarse_redo']} of {nq}; "
          f"|c|^2 max {cn.max():.2f} mean {cn.mean():.2f}; nearest other centroid d2 min {dd.min():.3f} mean {dd.mean():.3f}; "
          f"dc[:,0] mean {float(dc[:, 0].mean()):.3f} dc[:,15] mean {float(dc[:, 15].mean()):.3f}", flush=True)
    e.close()

centres = synth.uniform_device(0, kc, D, 1001)
xs = synth.blobs_device(0, 262144, centres, 1002)
tc, tb = synth.train_on_device_tensor(xs, kc, m, ksub, iters=8, init=centres)
Qd = synth.blobs_device(0, nq, centres, 2001).contiguous()
run("new data, trained centroids", tc.cpu().numpy(), tb.cpu().numpy(), Qd)
run("new data, blob centres as centroids", centres.cpu().numpy(), tb.cpu().numpy(), Qd)
X = synth.blobs(1_000_000, D, kc, seed=1002); Q = synth.blobs(nq, D, kc, seed=2001)
cent, cb = synth.train_on_device(X, kc, m, ksub, init=synth.blob_centres(D, kc))
run("old data, trained centroids", cent, cb, torch.from_numpy(Q).to(dev))
run("old centroids, new queries", cent, cb, Qd)
