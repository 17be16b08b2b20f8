#!/usr/bin/env python
"""Build / mutation throughput (BASELINE.json configs[4]): push! of 10 M 128-d vectors (coarse assign w = 1 +
PQ residual encode, m = 16, + append to the device-resident lists) and delete_from_index! compaction.

  python bench_build.py [--n 10000000] [--chunk 1000000] [--delete 1000000]

Everything goes through the host API the reference's callers use (ivfadc_add / ivfadc_delete on HOST buffers:
the host->device copy of the vectors is inside the timed region).  The reference has no batch push!
(src/utils.jl:114-145 is one vector per call, O(N) per call); the batched entry is this engine's extension
(SURVEY 3.4).  The CPU figure beside it is the oracle's encode (the C restatement of _encode_point,
src/utils.jl:148-161) on the host cores over a bounded sample.  One JSON line on stdout.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--chunk", type=int, default=1_000_000)
    ap.add_argument("--delete", type=int, default=1_000_000)
    ap.add_argument("--cpu-sample", type=int, default=200_000)
    args = ap.parse_args()

    import torch
    import ivfadc_jl_b200 as iv
    from ivfadc_jl_b200 import synth
    from oracle import oracle as orc

    if not torch.cuda.is_available():
        raise SystemExit("bench_build.py needs a CUDA device: the engine has no CPU fallback")
    D, kc, m, ksub = 128, 1024, 16, 256
    # two distinct chunks of synthetic vectors, pushed alternately (host RAM: 1 GB instead of 5 GB)
    blocks = [synth.blobs(args.chunk, D, kc, seed=1002 + i) for i in range(2)]
    cent, cb = synth.train_on_device(blocks[0], kc, m, ksub, init=synth.blob_centres(D, kc))
    pinned = [torch.from_numpy(b).pin_memory() for b in blocks]
    e = iv.IVFADCIndex.from_quantizers(cent, cb, None, index_type=np.uint32)
    iv.push_batch(e, blocks[0][:1000])     # warm-up: kernels loaded, workspaces sized
    iv.delete_from_index(e, np.arange(1, 1001))
    assert len(e) == 0
    torch.cuda.synchronize()
    e.stats(reset=True)

    times = []
    nchunks = args.n // args.chunk
    for i in range(nchunks):
        x = pinned[i & 1].numpy()
        t = time.perf_counter()
        iv.push_batch(e, x)               # synchronous: H2D + coarse (w = 1) + encode + append
        times.append(time.perf_counter() - t)
    total = nchunks * args.chunk
    assert len(e) == total
    push_s = sum(times)
    st = e.stats()

    rng = np.random.default_rng(7)
    del_ids = rng.choice(total, size=min(args.delete, total), replace=False).astype(np.int64) + 1   # 1-based (src/utils.jl:93)
    t = time.perf_counter()
    iv.delete_from_index(e, del_ids)
    del_s = time.perf_counter() - t
    assert len(e) == total - len(del_ids)

    # spot check against the oracle: codes and cells of a sample of the first chunk, bit-exact
    ns = min(args.cpu_sample, args.chunk)
    qz = orc.Quantizers(cent, cb, None)
    nth = os.cpu_count() or 1
    t = time.perf_counter()
    ocells, ocodes = orc.encode(qz, blocks[0][:ns], nthreads=nth)
    cpu_s = time.perf_counter() - t
    e2 = iv.IVFADCIndex.from_quantizers(cent, cb, None, index_type=np.uint32)
    gcells, gcodes = e2.encode(blocks[0][:ns])
    parity = bool(np.array_equal(np.asarray(gcells).astype(np.int64), np.asarray(ocells).astype(np.int64)) and
                  np.array_equal(gcodes, ocodes))
    e2.close()

    line = {
        "metric": "push! throughput (coarse assign + PQ residual encode + append), host buffers",
        "value": total / push_s, "unit": "vectors/s", "n_gpus": 1,
        "config": {"workload": "Build/encode throughput: push! of 10M 128-d vectors (coarse assign + PQ residual encode, "
                               "m=16) plus delete_from_index! compaction",
                   "n": total, "chunk": args.chunk, "D": D, "kc": kc, "m": m, "ksub": ksub, "ids": "UInt32"},
        "push": {"seconds": push_s, "per_chunk_ms": [round(1e3 * x, 2) for x in times],
                 "h2d_bytes": int(total) * D * 4, "gpu_launches": int(st["gpu_launches"])},
        "delete": {"ids": int(len(del_ids)), "seconds": del_s, "ids_per_s": len(del_ids) / del_s,
                   "vectors_compacted_per_s": total / del_s},
        "cpu_baseline": {"value": ns / cpu_s, "unit": "vectors/s", "cores": nth, "kind": "port",
                         "sample": f"oracle encode of {ns} vectors ({cpu_s:.2f} s)"},
        "parity": {"vectors": ns, "cells_and_codes_bit_exact": parity},
        "data": "synthetic",
    }
    print(json.dumps(line))
    e.close()


if __name__ == "__main__":
    main()
