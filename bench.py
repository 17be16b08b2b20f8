#!/usr/bin/env python
"""Benchmark of the IVFADC hot path: batched knn_search QPS (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload B] [--extras auto]

A "step" is one batched knn_search of the whole query batch (10 000 queries, k = 10, nprobe = 16)
over the SIFT1M-shaped synthetic index (128-d, 1M vectors, kc = 1024, m = 16, 256 codewords).
  value  queries/s with the query batch already resident in HBM (CUDA events on the launch stream,
         L2 flushed between steps, max over ranks);
  e2e    queries/s through the public host API (ivfadc_search / ivfadc_search_sharded on pinned host buffers:
         H2D of the queries and D2H of ids/distances/counts inside the timed region, host clock around the call);
  roofline  list-scan kernel: algorithmic PQ-code bytes per launch / its CUDA-event time, against
         the measured HBM copy bandwidth (MEASURED_PEAKS.json);
  cpu_baseline  the C restatement of the reference's CPU path (oracle/, Julia is not available in
         this image) on the box's host cores over the same index and queries;
  extra  the other named shapes of BASELINE.json measured in the same run: C (Deep10M-shaped) at every N,
         D (100 M vectors) at N = 1 (code array beyond the L2: the HBM-streaming scan) and at N = 8,
         E (build: push! of 10 M vectors + delete_from_index!) at N = 1.
Data: counter-based generator (Philox4x32-10 keyed (seed, vector, dim); csrc/synth.cu on the device, the same
function bit for bit in oracle/ for the CPU arm), so the 10 M / 100 M-vector sets are produced in HBM chunk by chunk.
N > 1 (torchrun, one rank per GPU): inverted lists sharded by cell (owners balanced by list length), the whole
step inside libivfadc_cuda -- coarse slice, one grouped NCCL all-gather of the probe lists, local scan, one grouped
all-gather of the candidates, merge -- replayed from a CUDA graph; total work is fixed -> "strong" scaling.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "knn_search QPS (batch 10k, k=10, nprobe=16)"
SEED_CENTRES, SEED_DATA, SEED_QUERIES = 1001, 1002, 2001
SIGMA = 0.05
TRAIN_ITERS = 8
CHUNK = 1 << 20

WORKLOADS = {
    "B": dict(name="SIFT1M-shaped synthetic: 128-d, 1M vectors, kc=1024, m=16, k=256, 10k-query batch, nprobe=16, k=10",
              D=128, N=1_000_000, kc=1024, m=16, ksub=256, nq=10_000, k=10, w=16),
    "C": dict(name="Deep10M-shaped synthetic: 96-d, 10M vectors, kc=4096, m=12, SqEuclidean, 10k-query batch, nprobe=16, k=10",
              D=96, N=10_000_000, kc=4096, m=12, ksub=256, nq=10_000, k=10, w=16),
    "D": dict(name="100M-vector 128-d synthetic, kc=16384, m=8 (dsub=16), UInt32 ids, 10k-query batch, nprobe=16, k=10",
              D=128, N=100_000_000, kc=16384, m=8, ksub=256, nq=10_000, k=10, w=16),
    "S": dict(name="small smoke workload (not a bench line)",
              D=64, N=200_000, kc=256, m=16, ksub=256, nq=2_000, k=10, w=8),
}
SCAN_KERNELS = {1: "scan_kernel (vector per lane, exact tables)", 2: "scanq_kernel (query per lane, shared-memory tables)",
                3: "scant_kernel", 4: "scanu_kernel (tensor-memory lookups)",
                5: "scanw_kernel (warp-specialised tensor-memory lookups)"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_config(wl):
    """The part of `config` both arms (ours / --impl reference) print identically."""
    code_mb = wl["N"] * wl["m"] / 1e6
    return {"workload": wl["name"], "nq": wl["nq"], "k": wl["k"], "nprobe": wl["w"],
            "data": f"Philox4x32-10 blob mixture (kc blobs, sigma {SIGMA}), seeds {SEED_CENTRES}/{SEED_DATA}/{SEED_QUERIES}; "
                    f"quantizers: Lloyd x {TRAIN_ITERS} from the blob centres on the first max(262144, 64 kc) vectors",
            "l2": f"GPU arm: L2 flushed between steps (256 MiB write); the {code_mb:.0f} MB code array is "
                  + ("L2-resident within a step" if code_mb < 100 else "larger than the 126 MB L2 (streams from HBM)"),
            "timing": "GPU arm: CUDA events on the launch stream, per step, mean, max over ranks; CPU arm: host clock"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append((float(f[0]), float(f[1])))
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(s[0] for s in self.samples)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0][1], "reasons": sorted(self.reasons),
                "samples": len(sm)}


# ---- CPU arm helpers (oracle) ------------------------------------------------------------------------------
def cpu_reference_qps(qz, offsets, codes, ids, Q, k, w, nthreads, budget_s=12.0):
    """Time the oracle (C restatement of the reference's CPU path) on a bounded sample."""
    from oracle import oracle as orc
    n0 = min(len(Q), 64 * nthreads)
    t = time.perf_counter()
    orc.search_csr(qz, offsets, codes, ids, Q[:n0], k, w, nthreads=nthreads)
    dt = time.perf_counter() - t
    n1 = int(min(len(Q), max(n0, n0 * budget_s / max(dt, 1e-6))))
    t = time.perf_counter()
    orc.search_csr(qz, offsets, codes, ids, Q[:n1], k, w, nthreads=nthreads)
    dt = time.perf_counter() - t
    return n1 / dt, n1, dt


def train_numpy(xs, kc, m, ksub, init, iters=TRAIN_ITERS, seed=3001):
    """The benchmark trainer on the host (same procedure as synth.train_on_device_tensor): Lloyd x iters from the
    blob centres, then one Lloyd per subspace on the residuals (random-sample seeding)."""
    def lloyd(x, c):
        for _ in range(iters):
            cn = (c * c).sum(1)
            a = np.empty(len(x), dtype=np.int64)
            for s in range(0, len(x), 1 << 15):
                a[s:s + (1 << 15)] = (cn[None, :] - 2.0 * (x[s:s + (1 << 15)] @ c.T)).argmin(1)
            cnt = np.bincount(a, minlength=len(c))
            order = np.argsort(a, kind="stable")
            start = np.concatenate(([0], np.cumsum(cnt)))[:-1]
            nz = cnt > 0
            sums = np.add.reduceat(x[order], np.minimum(start, len(x) - 1), axis=0)
            c = c.copy()
            c[nz] = sums[nz] / cnt[nz, None]
        return c.astype(np.float32), a
    cent, a = lloyd(xs, init.astype(np.float32))
    cn = (cent * cent).sum(1)
    for s in range(0, len(xs), 1 << 15):
        a[s:s + (1 << 15)] = (cn[None, :] - 2.0 * (xs[s:s + (1 << 15)] @ cent.T)).argmin(1)
    resid = xs - cent[a]
    dsub = xs.shape[1] // m
    rng = np.random.default_rng(seed)
    cb = np.empty((m, ksub, dsub), dtype=np.float32)
    for i in range(m):
        r = np.ascontiguousarray(resid[:, i * dsub:(i + 1) * dsub])
        cb[i], _ = lloyd(r, r[rng.choice(len(r), ksub, replace=False)])
    return cent, cb


def run_reference(args, wl):
    """--impl reference: the reference's CPU algorithm (oracle port; Julia unavailable) on the host cores, on the
    same workload: same generator (CPU twin of the device generator, bit-identical data and queries), same training
    procedure, index built by the oracle's own encoder, the full query batch per step.  No GPU code on this arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    from ivfadc_jl_b200 import synth
    nthreads = os.cpu_count() or 1
    D, N, kc, m, ksub, nq, k, w = (wl[x] for x in ("D", "N", "kc", "m", "ksub", "nq", "k", "w"))
    t0 = time.perf_counter()
    scale = synth.blob_scale(SIGMA)
    centres = orc.synth_uniform(0, kc, D, SEED_CENTRES)
    ns = min(N, max(262144, 64 * kc))
    xs, _ = orc.synth_blobs(0, ns, D, kc, SEED_DATA, scale, centres)
    cent, cb = train_numpy(xs, kc, m, ksub, centres)
    qz = orc.Quantizers(cent, cb, None)
    Q, _ = orc.synth_blobs(0, nq, D, kc, SEED_QUERIES, scale, centres)
    cells = np.empty(N, dtype=np.int64)
    ocodes = np.empty((N, m), dtype=np.uint8)
    exact_build = N <= 2_000_000
    for s0 in range(0, N, 1 << 18):
        n = min(1 << 18, N - s0)
        xb, _ = orc.synth_blobs(s0, n, D, kc, SEED_DATA, scale, centres)
        if exact_build:   # the oracle's own _encode_point: coarse w = 1 + PQ residual encoding
            cells[s0:s0 + n], ocodes[s0:s0 + n] = orc.encode(qz, xb, nthreads=nthreads)
        else:             # bounded build for the big shapes: BLAS assignment, uniform random codes (cost depends on list lengths only)
            cn = (cent.astype(np.float64) ** 2).sum(1)
            cells[s0:s0 + n] = (cn[None, :] - 2.0 * (xb @ cent.T)).argmin(1)
            ocodes[s0:s0 + n] = np.random.default_rng(7 + s0).integers(0, ksub, size=(n, m), dtype=np.uint8)
    order = np.argsort(cells, kind="stable")
    offsets = np.zeros(kc + 1, dtype=np.int64)
    np.cumsum(np.bincount(cells, minlength=kc), out=offsets[1:])
    codes_csr, ids_csr = ocodes[order], order.astype(np.uint64)
    build_s = time.perf_counter() - t0
    vals = []
    for s in range(args.warmup + args.steps):
        t = time.perf_counter()
        orc.search_csr(qz, offsets, codes_csr, ids_csr, Q, k, w, nthreads=nthreads)
        dt = time.perf_counter() - t
        if s >= args.warmup:
            vals.append(dt)
    ms = 1e3 * sum(vals) / len(vals)
    qps = nq / (ms / 1e3)
    sample = (f"all {nq} queries per step on the full index ({N} vectors); index built on the host in {build_s:.1f}s untimed "
              + ("(oracle encoder)" if exact_build else "(BLAS assignment, random codes)"))
    line = {"impl": "reference", "metric": METRIC, "value": qps,
            "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(wl),
            "note": "C restatement of IVFADC.jl's CPU path (Julia is not installed in this image); real Julia would be "
                    "slower (LittleDict / SortedMultiDict)",
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": nthreads, "kind": "port", "sample": sample},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---- GPU arm ---------------------------------------------------------------------------------------------------
class Ctx:
    pass


def build_index(cx, wl, flags):
    """Quantizers trained on rank 0 (torch: plumbing, outside the hot path) and broadcast; the data set generated in
    HBM chunk by chunk and added through ivfadc_add_device.  N > 1: a first pass assigns every vector (the engine's
    coarse kernel, w = 1) so that the cells can be dealt to the GPUs by list length; the second pass encodes."""
    import torch
    import ivfadc_jl_b200 as iv
    from ivfadc_jl_b200 import _capi, sharded, synth
    D, N, kc, m, ksub = (wl[x] for x in ("D", "N", "kc", "m", "ksub"))
    dev, dist = cx.dev, cx.dist
    t0 = time.perf_counter()
    centres = synth.uniform_device(0, kc, D, SEED_CENTRES, device=dev)
    if cx.rank == 0:
        ns = min(N, max(262144, 64 * kc))
        xs = synth.blobs_device(0, ns, centres, SEED_DATA, SIGMA)
        tc, tb = synth.train_on_device_tensor(xs, kc, m, ksub, iters=TRAIN_ITERS, init=centres)
        tc, tb = tc.contiguous(), tb.contiguous()
        del xs
    else:
        tc = torch.empty((kc, D), dtype=torch.float32, device=dev)
        tb = torch.empty((m, ksub, D // m), dtype=torch.float32, device=dev)
    if cx.world > 1:
        dist.broadcast(tc, 0)
        dist.broadcast(tb, 0)
    cent, cb = tc.cpu().numpy(), tb.cpu().numpy()
    del tc, tb
    prep_s = time.perf_counter() - t0

    engine = iv.IVFADCIndex.from_quantizers(cent, cb, None, index_type=np.uint32, device=cx.local_rank,
                                            shard=(cx.rank, cx.world), flags=flags)
    buf = torch.empty((min(CHUNK, N), D), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev)
    t0 = time.perf_counter()
    assign_s = 0.0
    if cx.world > 1:
        cells_all = torch.empty(N, dtype=torch.int32, device=dev)
        counts = torch.zeros(kc, dtype=torch.int64, device=dev)
        for s in range(0, N, CHUNK):
            n = min(CHUNK, N - s)
            x = synth.blobs_device(s, n, centres, SEED_DATA, SIGMA, out=buf)
            c, _ = sharded.coarse_device(engine, x, 1)
            cells_all[s:s + n] = c[:, 0]
            counts += torch.bincount(c[:, 0].long(), minlength=kc)
        torch.cuda.synchronize(dev)
        engine.check_async(stream.cuda_stream)
        owners = sharded.balanced_owners(counts.cpu().numpy(), cx.world)
        engine.set_cell_owners(owners)
        engine.reserve(N, np.where(owners == cx.rank, counts.cpu().numpy(), 0))   # exact list lengths: one allocation
        assign_s = time.perf_counter() - t0
        for s in range(0, N, CHUNK):
            n = min(CHUNK, N - s)
            x = synth.blobs_device(s, n, centres, SEED_DATA, SIGMA, out=buf)
            a = cells_all[s:s + n].long()
            stream.synchronize()
            engine.add_device(x.data_ptr(), n, _capi.LAST, a.data_ptr(), 0)
        del cells_all
        sharded.init_comm(engine)
    else:
        engine.reserve(N)
        for s in range(0, N, CHUNK):
            n = min(CHUNK, N - s)
            x = synth.blobs_device(s, n, centres, SEED_DATA, SIGMA, out=buf)
            stream.synchronize()
            engine.add_device(x.data_ptr(), n)
    build_s = time.perf_counter() - t0
    del buf
    Qd = synth.blobs_device(0, wl["nq"], centres, SEED_QUERIES, SIGMA).contiguous()
    torch.cuda.synchronize(dev)
    info = {"vectors": N, "seconds": build_s, "vectors_per_s": N / build_s, "prep_s": prep_s,
            "how": "device generator -> ivfadc_add_device (coarse w=1 + PQ encode + append), chunks of 2^20"
                   + (f"; first pass (assignment for the balanced cell owners) {assign_s:.2f}s of it" if cx.world > 1 else "")}
    return engine, cent, cb, Qd, info


def gather_probed_csr(cx, engine, qz, Q, w, kc, m):
    """CSR of exactly the lists the checked queries probe, assembled on rank 0 from the shards that own them."""
    from oracle import oracle as orc
    need = None
    if cx.rank == 0:
        cells, _ = orc.coarse_search(qz, Q, w, nthreads=os.cpu_count())
        need = np.unique(cells)
    if cx.world > 1:
        box = [need]
        cx.dist.broadcast_object_list(box, src=0)
        need = box[0]
    sizes = engine.list_sizes()
    mine = {int(c): engine.export_list(int(c)) for c in need if sizes[int(c)] > 0}
    if cx.world > 1:
        parts = [None] * cx.world if cx.rank == 0 else None
        cx.dist.gather_object(mine, parts, dst=0)
    else:
        parts = [mine]
    if cx.rank != 0:
        return None
    lists = {}
    for p in parts:
        lists.update(p)
    lens = np.zeros(kc, dtype=np.int64)
    for c, (i, _) in lists.items():
        lens[c] = len(i)
    offsets = np.zeros(kc + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    ids = np.empty(int(offsets[-1]), dtype=np.uint64)
    codes = np.empty((int(offsets[-1]), m), dtype=np.uint8)
    for c, (i, cd) in lists.items():
        ids[offsets[c]:offsets[c + 1]] = i
        codes[offsets[c]:offsets[c + 1]] = cd
    return offsets, codes, ids


def run_workload(cx, args, name, steps, headline):
    """Build the index of one workload, time the batched search (device-resident and through the host API), check a
    sample against the oracle.  Returns the record (rank 0) or None."""
    import torch
    from ivfadc_jl_b200 import _capi, sharded
    wl = dict(WORKLOADS[name])
    if os.environ.get("IVFADC_BENCH_N"):   # bring-up: shrink the data set
        wl["N"] = int(os.environ["IVFADC_BENCH_N"])
    dev, dist, world, rank = cx.dev, cx.dist, cx.world, cx.rank
    nq, k, w, D, kc, m = wl["nq"], wl["k"], wl["w"], wl["D"], wl["kc"], wl["m"]
    engine, cent, cb, dQ, build = build_index(cx, wl, args.flags)
    lib, h = engine._lib, engine._h
    out = (torch.empty((nq, k), dtype=torch.int64, device=dev), torch.empty((nq, k), dtype=torch.float32, device=dev),
           torch.empty((nq,), dtype=torch.int32, device=dev))
    stream = torch.cuda.current_stream(dev)

    def step_device():
        if world > 1:
            return sharded.search_sharded_device(engine, dQ, k, w, out=out)
        return sharded.search_device(engine, dQ, k, w, out=out)

    # ---- per-kernel breakdown (N > 1: eager steps; the headline replays the step from a CUDA graph) -------------
    breakdown, nbreak = None, 0
    if world > 1:
        _capi.check(h, lib.ivfadc_set_graph_replay(h, 0))
        for _ in range(2):
            step_device()
        torch.cuda.synchronize(dev)
        engine.stats(reset=True)
        nbreak = 5
        for _ in range(nbreak):
            cx.flush.zero_()
            step_device()
        torch.cuda.synchronize(dev)
        breakdown = engine.stats()
        _capi.check(h, lib.ivfadc_set_graph_replay(h, 0 if args.no_graph else 1))

    # ---- device-resident timing --------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        cx.flush.zero_()
        step_device()
    torch.cuda.synchronize(dev)
    engine.check_async(stream.cuda_stream)
    engine.stats(reset=True)
    if world > 1:
        dist.barrier()
    sampler = None
    if headline:
        sampler = ClockSampler(cx.local_rank)
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda.synchronize(dev)
    for a, b in evs:
        cx.flush.zero_()
        a.record()
        step_device()
        b.record()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs) / steps
    engine.check_async(stream.cuda_stream)
    st = engine.stats()
    clocks = sampler.stop() if sampler else None
    launches = int(st["gpu_launches"])
    if breakdown is not None:
        st = dict(st)
        for key_ in ("coarse_ms", "plan_ms", "scan_ms", "merge_ms", "comm_ms", "scan_launches", "scan_code_bytes",
                     "scanned_vectors", "last_scan_kernel", "last_coarse_redo"):
            st[key_] = breakdown[key_]
        launches = int(breakdown["gpu_launches"] / nbreak * steps)
    else:
        nbreak = steps
    if world > 1:
        t = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    result = tuple(x.clone() for x in out)

    # ---- end to end through the host API (pinned host buffers, host clock around the synchronous call) ----------
    hQ = dQ.cpu().pin_memory()
    h_ids = torch.empty((nq, k), dtype=torch.int64).pin_memory()
    h_d = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    h_c = torch.empty((nq,), dtype=torch.int32).pin_memory()
    fn = lib.ivfadc_search_sharded if world > 1 else lib.ivfadc_search

    def step_host():
        rc = fn(h, ctypes.c_void_p(hQ.data_ptr()), nq, k, w, ctypes.c_void_p(h_ids.data_ptr()),
                ctypes.c_void_p(h_d.data_ptr()), ctypes.c_void_p(h_c.data_ptr()))
        _capi.check(h, rc)
    for _ in range(max(args.warmup, 3)):
        step_host()
    ts = []
    for _ in range(steps):
        cx.flush.zero_()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)
        t = time.perf_counter()
        step_host()  # synchronous: H2D + kernels (+ collectives) + D2H
        ts.append(time.perf_counter() - t)
    e2e_ms = 1e3 * sum(ts) / len(ts)
    h2d, d2h = int(nq * D * 4), int(nq * k * 12 + nq * 4)
    note = None
    if world > 1:
        t2 = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e_ms = float(t2.item())
        b_h2d, b_d2h, b_nv = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        lib.ivfadc_sharded_step_bytes(h, nq, k, w, ctypes.byref(b_h2d), ctypes.byref(b_d2h), ctypes.byref(b_nv))
        h2d, d2h = int(b_h2d.value) * world, int(b_d2h.value)
        note = (f"one ivfadc_search_sharded call per rank: every rank uploads its 1/{world} slice of the batch, the "
                f"all-gathers complete it over NVLink ({int(b_nv.value)} B received per rank and step); d2h = rank 0's result")
    e2e = {"value": nq / (e2e_ms / 1e3), "unit": "queries/s", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}
    if note:
        e2e["note"] = note

    # ---- parity sample against the oracle (+ CPU baseline on the headline) ---------------------------------------
    parity, cpu = None, None
    nchk = min(args.check, nq, 256 if world == 1 else (64 if wl["N"] <= 20_000_000 else 32))
    if nchk:
        from oracle import oracle as orc
        qz = orc.Quantizers(cent, cb, None)
        Qh = hQ.numpy()
        full = world == 1 and headline and not args.no_cpu_baseline
        if full:     # the CPU baseline needs the whole index: one bulk export
            sizes, ids_csr, codes_csr = engine.export_all()
            offsets = np.zeros(kc + 1, dtype=np.int64)
            np.cumsum(sizes, out=offsets[1:])
            csr = (offsets, codes_csr, ids_csr.astype(np.uint64))
        else:
            csr = gather_probed_csr(cx, engine, qz, Qh[:nchk], w, kc, m)
        if rank == 0:
            oi, od, oc, _ = orc.search_csr(qz, csr[0], csr[1], csr[2], Qh[:nchk], k, w, nthreads=os.cpu_count())
            gi = result[0][:nchk].cpu().numpy().view(np.uint64)
            gd = result[1][:nchk].cpu().numpy()
            gcn = result[2][:nchk].cpu().numpy()
            parity = {"queries": nchk, "rtol": 1e-5}
            try:
                parity.update(orc.compare_search(gi, gd, gcn, oi, od, oc, rtol=1e-5))
                parity["ok"] = True
                hi = h_ids[:nchk].numpy().view(np.uint64)
                parity["host_api_equals_device_api"] = bool(np.array_equal(hi, gi) and np.array_equal(h_d[:nchk].numpy(), gd))
            except AssertionError as ex:
                parity.update({"ok": False, "error": str(ex)[:200]})
            if full:
                nth = os.cpu_count() or 1
                qps1, n1, dt1 = cpu_reference_qps(qz, *csr, Qh, k, w, 1, budget_s=8.0)
                qpsN, nN, dtN = cpu_reference_qps(qz, *csr, Qh, k, w, nth, budget_s=12.0)
                cpu = {"value": qpsN, "unit": "queries/s", "cores": nth, "kind": "port",
                       "sample": f"{nN} of {nq} queries, full index, {dtN:.1f}s; single thread: {qps1:.0f} q/s on {n1} queries",
                       "single_thread_value": qps1}

    ranks = None
    if world > 1:
        mine = [float(st["scan_code_bytes"]) / max(1, st["scan_launches"]), float(st["scan_ms"]) / max(1, st["scan_launches"]),
                float(engine.list_sizes().sum())]
        ranks = [None] * world
        dist.all_gather_object(ranks, mine)
    rec = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        scan_ms = st["scan_ms"] / max(1, st["scan_launches"])
        bytes_per_launch = st["scan_code_bytes"] / max(1, st["scan_launches"])
        achieved = bytes_per_launch / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else 0.0
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "scan_dram_traffic.json")
        if os.path.exists(tp) and world == 1:
            tj = json.load(open(tp))
            traffic = tj.get(name)
            traffic_src = tj.get(name + "_source", tj.get("source", "ncu capture kept in profiles/ (not re-measured in this run)")) if traffic else None
        kern = SCAN_KERNELS.get(int(st.get("last_scan_kernel", 0)), "?")
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel": kern + " (K2 lookup tables + K3 list scan + per-list candidate selection)",
                "algorithmic_bytes_per_launch": bytes_per_launch, "kernel_ms": scan_ms, "peak_source": peak_src,
                "per_rank": world > 1}
        if ranks:
            roof["ranks"] = [{"algorithmic_bytes": r[0], "kernel_ms": r[1], "frac": (r[0] / (r[1] * 1e-3) / 1e9 / peak) if r[1] > 0 else 0.0,
                              "vectors": r[2]} for r in ranks]
        bd = {"coarse": st["coarse_ms"] / nbreak, "plan": st["plan_ms"] / nbreak, "scan": st["scan_ms"] / nbreak,
              "merge": st["merge_ms"] / nbreak}
        bd["coarse_redo_queries_last_step"] = int(st.get("last_coarse_redo", 0))
        if world > 1:
            bd["comm"] = st.get("comm_ms", 0.0) / nbreak
            bd["note"] = "eager steps (per-kernel CUDA events); the timed steps replay the same work from a CUDA graph"
        rec = {"value": nq / (dev_ms / 1e3), "ms_per_step": dev_ms, "steps": steps,
               "config": workload_config(wl),
               "engine": {"lists": f"cell-sharded over {world} GPUs, owners balanced by list length" if world > 1 else "one GPU",
                          "launch": ("CUDA graph replay of the sharded step inside the library" if world > 1 and not args.no_graph
                                     else "eager"), "flags": args.flags, "scan": kern,
                          "coarse": ("packed-FP32 FFMA kernel" if (args.flags & 160) else
                                     "tcgen05 kind::tf32 scores prune to a provable superset of the top-w, exact direct-form re-rank"),
                          "tables": ("exact direct form (fp32 chain)" if (args.flags & 5) else
                                     "mma.sync 3xTF32 GEMM form" if (args.flags & 8) else
                                     "tcgen05 GEMM form (fp16 two-piece operands, fp32 accumulators in tensor memory)")},
               "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
               "breakdown_ms": bd, "build": build, "parity": parity}
    if world > 1:
        torch.cuda.synchronize(dev)
        dist.barrier()
        lib.ivfadc_comm_destroy(h)
    engine.close()
    del dQ, out, result
    torch.cuda.empty_cache()
    return rec


def run_build_extra(cx, args):
    """BASELINE.json configs[4] inside the bench line (N = 1): push! of 10 M 128-d vectors (coarse assign w = 1 + PQ
    residual encode, m = 16, + append) device-resident and through the host API, then delete_from_index! of 1 M ids;
    the cells / codes of a sample against the oracle's _encode_point, bit for bit."""
    import torch
    import ivfadc_jl_b200 as iv
    from ivfadc_jl_b200 import synth
    from oracle import oracle as orc
    D, N, kc, m, ksub = 128, int(os.environ.get("IVFADC_BENCH_E_N", 10_000_000)), 1024, 16, 256
    dev = cx.dev
    centres = synth.uniform_device(0, kc, D, SEED_CENTRES, device=dev)
    xs = synth.blobs_device(0, 262144, centres, SEED_DATA, SIGMA)
    tc, tb = synth.train_on_device_tensor(xs, kc, m, ksub, iters=TRAIN_ITERS, init=centres)
    cent, cb = tc.cpu().numpy(), tb.cpu().numpy()
    del xs, tc, tb
    stream = torch.cuda.current_stream(dev)
    buf = torch.empty((CHUNK, D), dtype=torch.float32, device=dev)
    # (1) device-resident: the batch is already in HBM
    e = iv.IVFADCIndex.from_quantizers(cent, cb, None, index_type=np.uint32, device=cx.local_rank)
    x = synth.blobs_device(0, 4096, centres, SEED_DATA, SIGMA, out=buf)
    stream.synchronize()
    e.add_device(x.data_ptr(), 4096)                      # warm-up: kernels loaded, workspaces sized
    iv.delete_from_index(e, np.arange(1, 4097))
    e.reserve(N)                                          # sizehint!: one arena allocation for the bulk build
    dev_s = 0.0
    for s0 in range(0, N, CHUNK):
        n = min(CHUNK, N - s0)
        x = synth.blobs_device(s0, n, centres, SEED_DATA, SIGMA, out=buf)
        stream.synchronize()
        t = time.perf_counter()
        e.add_device(x.data_ptr(), n)                     # synchronous: coarse (w = 1) + encode + append
        dev_s += time.perf_counter() - t
    assert len(e) == N
    rng = np.random.default_rng(7)
    del_ids = rng.choice(N, size=min(1_000_000, N // 2), replace=False).astype(np.int64) + 1   # 1-based (src/utils.jl:93)
    t = time.perf_counter()
    iv.delete_from_index(e, del_ids)
    del_s = time.perf_counter() - t
    assert len(e) == N - len(del_ids)
    e.close()
    # (2) through the host API: two pinned host chunks pushed alternately, H2D inside the timed region
    e = iv.IVFADCIndex.from_quantizers(cent, cb, None, index_type=np.uint32, device=cx.local_rank)
    pinned = [synth.blobs_device(i * CHUNK, CHUNK, centres, SEED_DATA, SIGMA).cpu().pin_memory() for i in range(2)]
    iv.push_batch(e, pinned[0].numpy()[:4096])
    iv.delete_from_index(e, np.arange(1, 4097))
    e.reserve(N)
    host_s, nhost = 0.0, 0
    for i in range(max(1, N // CHUNK)):
        xh = pinned[i & 1].numpy()
        t = time.perf_counter()
        iv.push_batch(e, xh)
        host_s += time.perf_counter() - t
        nhost += CHUNK
    # (3) parity of a sample: regenerate the vectors on the CPU (counter-based) and encode them with the oracle
    ns = 50_000
    xc, _ = orc.synth_blobs(0, ns, D, kc, SEED_DATA, synth.blob_scale(SIGMA), centres.cpu().numpy())
    qz = orc.Quantizers(cent, cb, None)
    nth = os.cpu_count() or 1
    t = time.perf_counter()
    ocells, ocodes = orc.encode(qz, xc, nthreads=nth)
    cpu_s = time.perf_counter() - t
    gcells, gcodes = e.encode(xc)
    same_input = bool(np.array_equal(xc, pinned[0].numpy()[:ns]))
    parity = bool(np.array_equal(np.asarray(gcells).astype(np.int64), np.asarray(ocells).astype(np.int64))
                  and np.array_equal(gcodes, ocodes))
    e.close()
    del buf, pinned
    torch.cuda.empty_cache()
    flop = (2.0 * kc * D + 3.0 * D * ksub) * N     # direct-form work of K1 (w = 1) + K4 per vector
    return {"config": {"workload": "Build/encode throughput: push! of 10M 128-d vectors (coarse assign + PQ residual encode, "
                                   "m=16) plus delete_from_index! compaction", "n": N, "chunk": CHUNK, "D": D, "kc": kc,
                       "m": m, "ksub": ksub, "ids": "UInt32"},
            "device_resident": {"value": N / dev_s, "unit": "vectors/s", "seconds": dev_s,
                                "algorithmic_tflops": flop / dev_s / 1e12,
                                "note": "ivfadc_add_device on batches already in HBM (coarse w=1 on the tensor cores + exact "
                                        "re-rank, PQ encode, sort by cell + append)"},
            "host_api": {"value": nhost / host_s, "unit": "vectors/s", "seconds": host_s, "h2d_bytes": int(nhost) * D * 4,
                         "note": "ivfadc_add on pinned host buffers (the host->device copy inside the timed region)"},
            "delete": {"ids": int(len(del_ids)), "seconds": del_s, "vectors_compacted_per_s": N / del_s},
            "cpu_baseline": {"value": ns / cpu_s, "unit": "vectors/s", "cores": nth, "kind": "port",
                             "sample": f"oracle encode of {ns} vectors ({cpu_s:.2f} s)"},
            "parity": {"vectors": ns, "cells_and_codes_bit_exact": parity, "cpu_generator_equals_device": same_input}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="B", choices=sorted(WORKLOADS))
    ap.add_argument("--extras", default="auto", help="auto | none | comma list of workloads measured after the headline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--check", type=int, default=256, help="queries verified against the oracle")
    ap.add_argument("--no-graph", action="store_true", help="N > 1: launch the sharded step eagerly instead of replaying its CUDA graph")
    ap.add_argument("--flags", type=int, default=0,
                    help="ivfadc_config.flags (1 vector-per-lane scan, 2 query-per-lane scan, 4 exact tables, 8 mma.sync tables, "
                         "16 round-1 tensor-memory kernel, 32 scalar coarse, 128 packed-FP32 coarse)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl)
    args.warmup = max(args.warmup, 3)

    import torch
    cx = Ctx()
    cx.world = int(os.environ.get("WORLD_SIZE", "1"))
    cx.rank = int(os.environ.get("RANK", "0"))
    cx.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(cx.local_rank)
    cx.dev = torch.device("cuda", cx.local_rank)
    cx.dist = None
    if cx.world > 1:
        # rank 0 prints ONE JSON line on stdout: keep NCCL's version banner (NCCL_DEBUG=VERSION in this image) off it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=cx.dev)
        cx.dist = dist
    cx.flush = torch.empty(256 << 20, dtype=torch.uint8, device=cx.dev)  # > 126 MB L2

    head = run_workload(cx, args, args.workload, args.steps, headline=True)
    extras = {}
    if args.extras == "auto":
        names = []
        if args.workload == "B":
            names = ["C"] + (["D"] if cx.world in (1, 8) else []) + (["E"] if cx.world == 1 else [])
    elif args.extras == "none":
        names = []
    else:
        names = [x for x in args.extras.split(",") if x]
    for name in names:
        try:
            rec = (run_build_extra(cx, args) if name == "E" else
                   run_workload(cx, args, name, min(args.steps, 10), headline=False))
            if cx.rank == 0:
                extras[name] = rec
        except Exception as ex:   # an extra must never take the headline down
            if cx.rank == 0:
                extras[name] = {"error": f"{type(ex).__name__}: {str(ex)[:300]}"}
            if cx.world > 1:
                break   # the ranks may be out of step: no further collectives

    if cx.rank == 0:
        line = {"metric": METRIC, "value": head["value"], "unit": "queries/s", "n_gpus": cx.world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
        for key_ in ("config", "engine", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "breakdown_ms",
                     "build", "parity"):
            line[key_] = head[key_]
        line["extra"] = extras
        print(json.dumps(line))
    if cx.world > 1:
        torch.cuda.synchronize()
        cx.dist.barrier()
        cx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
