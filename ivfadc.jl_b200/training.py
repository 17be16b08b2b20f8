"""
Quantizer training for the host-side constructor -- OUTSIDE the parity scope.

The reference trains with Clustering.kmeans (k-means++ seeding from Julia's global RNG,
reference src/index.jl:129-134) and QuantizedArrays.build_quantizer (one k-means per sub-space,
src/index.jl:142-147).  Both are unseeded, so no output of theirs is reproducible or pinned by any
test; in the Julia integration they stay on the Julia side.
This module is the stand-in the Python mirror of the constructor and the benchmark harness use:
plain Lloyd iterations with k-means++ seeding, numpy on the host, or torch on the GPU for the
benchmark shapes.  Its results are INPUTS to both the oracle and the CUDA engine.
"""
from __future__ import annotations

import numpy as np


def _kmpp_init(X: np.ndarray, k: int, rng: np.random.Generator) -> np.ndarray:
    n = X.shape[0]
    centers = np.empty((k, X.shape[1]), dtype=np.float64)
    centers[0] = X[rng.integers(n)]
    d2 = ((X - centers[0]) ** 2).sum(1)
    for j in range(1, k):
        tot = d2.sum()
        if tot <= 0:
            idx = rng.integers(n)
        else:
            idx = min(int(np.searchsorted(np.cumsum(d2), rng.random() * tot)), n - 1)
        centers[j] = X[idx]
        d2 = np.minimum(d2, ((X - centers[j]) ** 2).sum(1))
    return centers


def kmeans(X, k: int, maxiter: int = 25, seed: int = 0):
    """Lloyd's algorithm.  X is [n, d]; returns (centers [k, d] in X.dtype, assignments int64[n]
    0-based, consistent with the returned centers as Clustering.jl's are)."""
    X64 = np.asarray(X, dtype=np.float64)
    n = X64.shape[0]
    assert 1 <= k <= n
    rng = np.random.default_rng(seed)
    centers = _kmpp_init(X64, k, rng)
    assign = np.zeros(n, dtype=np.int64)
    xn = (X64 ** 2).sum(1)
    for it in range(maxiter + 1):
        d = xn[:, None] - 2.0 * X64 @ centers.T + (centers ** 2).sum(1)[None, :]
        new_assign = d.argmin(1)
        if it == maxiter or (it > 0 and np.array_equal(new_assign, assign)):
            assign = new_assign
            break
        assign = new_assign
        counts = np.bincount(assign, minlength=k)
        sums = np.zeros_like(centers)
        np.add.at(sums, assign, X64)
        empty = counts == 0
        centers[~empty] = sums[~empty] / counts[~empty, None]
        if empty.any():  # re-seed empty clusters on the points farthest from their centre
            far = np.argsort(-d[np.arange(n), assign])[: int(empty.sum())]
            centers[empty] = X64[far]
    centers = centers.astype(np.asarray(X).dtype)
    # final assignments consistent with the (rounded) returned centres
    c64 = centers.astype(np.float64)
    d = xn[:, None] - 2.0 * X64 @ c64.T + (c64 ** 2).sum(1)[None, :]
    return centers, d.argmin(1).astype(np.int64)


def train_quantizers(data, kc: int, k: int, m: int, coarse_maxiter: int = 25,
                     quantization_maxiter: int = 25, seed: int = 0):
    """data [n, D] -> (centroids [kc, D], assignments int64[n] 0-based,
    cb_vectors [m, k, dsub], cb_codes uint8[m, k]).  Mirrors the constructor's training phase
    (reference src/index.jl:129-147): coarse k-means, residuals w.r.t. the assigned centre,
    one k-means per rowrange(D, m, i) of the residuals; codes are 0..k-1 in column order."""
    data = np.asarray(data)
    n, D = data.shape
    dsub = D // m
    centroids, assign = kmeans(data, kc, coarse_maxiter, seed)
    resid = data - centroids[assign]
    cb_vectors = np.empty((m, k, dsub), dtype=data.dtype)
    for i in range(m):
        cb_vectors[i], _ = kmeans(resid[:, i * dsub:(i + 1) * dsub], k, quantization_maxiter,
                                  seed + 1 + i)
    cb_codes = np.tile(np.arange(k, dtype=np.uint8), (m, 1))
    return centroids, assign, cb_vectors, cb_codes


def kmeans_torch(X, k: int, iters: int = 10, seed: int = 0, chunk: int = 1 << 18, init=None):
    """Benchmark-harness trainer: Lloyd on whatever device X (a torch tensor [n, d]) lives on,
    random-sample seeding.  Not graded, not on the hot path."""
    import torch

    g = torch.Generator(device="cpu").manual_seed(seed)
    n, d = X.shape
    if init is not None:
        centers = init.clone().float()
    else:
        centers = X[torch.randperm(n, generator=g)[:k].to(X.device)].clone().float()
    for _ in range(iters):
        sums = torch.zeros(k, d, device=X.device, dtype=torch.float32)
        counts = torch.zeros(k, device=X.device, dtype=torch.float32)
        cn = (centers ** 2).sum(1)
        for s in range(0, n, chunk):
            xb = X[s:s + chunk].float()
            a = (cn[None, :] - 2.0 * xb @ centers.T).argmin(1)
            sums.index_add_(0, a, xb)
            counts.index_add_(0, a, torch.ones_like(a, dtype=torch.float32))
        nz = counts > 0
        centers[nz] = sums[nz] / counts[nz, None]
        if (~nz).any():
            centers[~nz] = X[torch.randint(0, n, (int((~nz).sum()),), generator=g).to(X.device)].float()
    return centers


# ---------------------------------------------------------------------------------------------
# Device trainer: Lloyd iterations whose assignment step is the engine's own coarse
# kernel (K1: exact nearest centre in the reference's direct form, tensor-core pruned where the
# shape allows) -- what makes IVFADCIndex(data; ...) practical at 10^6+ vectors, where the host
# trainer above would need an n x kc float64 distance matrix.  Seeding (k-means++ on a sample) and the
# centre update are kernels of the library too (csrc/train.cu); torch carries the device buffers.  Not parity-graded (the
# reference's training is unseeded); graded by quantisation error against the host trainer
# (tests/test_gpu_parity.py::test_device_trainer).
# ---------------------------------------------------------------------------------------------
def _dummy_codebook(d: int, dtype):
    """The engine wants a product quantizer next to the centroids; training only uses the coarse step."""
    m = max(x for x in range(1, min(d, 16) + 1) if d % x == 0)
    return np.zeros((m, 2, d // m), dtype=dtype)


def kmeans_device(X, k: int, maxiter: int = 25, seed: int = 0, device: int = 0, chunk: int = 1 << 20):
    """Lloyd on the GPU.  X [n, d] (numpy, float32 / float64) -> (centers [k, d] in X.dtype, assignments
    int64[n] 0-based, consistent with the returned centers).

    Everything that touches the data is a kernel of the library: k-means++ seeding (ivfadc_kmeanspp_device, D^2
    sampling on a sample, Philox keyed by `seed`), the assignment step (the engine's coarse kernel through ONE handle
    whose centroids are replaced in place every iteration, ivfadc_set_centroids_device), the centre update
    (ivfadc_kmeans_accumulate_device / ivfadc_kmeans_finish_device).  torch only owns the device buffers."""
    import ctypes

    import torch

    from . import _capi, sharded
    from .index import IVFADCIndex

    X = np.ascontiguousarray(X)
    n, d = X.shape
    assert 1 <= k <= n
    lib = _capi.load()
    dev = torch.device("cuda", device)
    dX = torch.from_numpy(X).to(dev)
    dt = _capi.F32 if X.dtype == np.float32 else _capi.F64
    rng = np.random.default_rng(seed)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731

    def ok(rc, what):
        if rc != 0:
            raise _capi.IvfadcError(rc, what)

    # k-means++ seeding on a sample
    ns = min(n, max(32768, 8 * k))
    S = dX[torch.from_numpy(rng.choice(n, ns, replace=False)).to(dev)].contiguous()
    centers = torch.empty((k, d), dtype=dX.dtype, device=dev)
    scratch = torch.empty(ns, dtype=torch.float64, device=dev)
    picked = torch.empty(k, dtype=torch.int64, device=dev)
    # greedy k-means++ like scikit-learn (2 + ln k candidates per round) while the seeding stays affordable on one CTA
    trials = 2 + int(np.log(k)) if float(ns) * k * d <= 4e10 else 1
    ok(lib.ivfadc_kmeanspp_device(vp(S), ns, d, k, dt, seed, trials, vp(scratch), vp(centers), vp(picked), stream),
       "ivfadc_kmeanspp_device")
    torch.cuda.synchronize(dev)

    eng = IVFADCIndex.from_quantizers(centers.cpu().numpy(), _dummy_codebook(d, X.dtype), None, device=device)
    assign = torch.zeros(n, dtype=torch.int32, device=dev)
    new_assign = torch.empty_like(assign)
    dist = torch.empty(n, dtype=dX.dtype, device=dev)
    sums = torch.empty((k, d), dtype=torch.float64, device=dev)
    counts = torch.empty(k, dtype=torch.int64, device=dev)
    empty = torch.empty(k, dtype=torch.int32, device=dev)
    try:
        for it in range(maxiter + 1):
            sums.zero_()
            counts.zero_()
            for s in range(0, n, chunk):  # K1: nearest centre of every point (w = 1), then the sums of the chunk
                xb = dX[s:s + chunk]
                c, dc = sharded.coarse_device(eng, xb, 1)
                new_assign[s:s + chunk] = c[:, 0]
                dist[s:s + chunk] = dc[:, 0]
                ok(lib.ivfadc_kmeans_accumulate_device(vp(xb), xb.shape[0], d, dt, vp(new_assign[s:s + chunk]), vp(sums),
                                                       vp(counts), stream), "ivfadc_kmeans_accumulate_device")
            same = it > 0 and bool(torch.equal(new_assign, assign))
            assign, new_assign = new_assign, assign
            if it == maxiter or same:
                break
            ok(lib.ivfadc_kmeans_finish_device(vp(sums), vp(counts), k, d, dt, vp(centers), vp(empty), stream),
               "ivfadc_kmeans_finish_device")
            nempty = int(empty.sum())
            if nempty:  # re-seed empty clusters on the points farthest from their centre
                far = torch.topk(dist.double(), nempty).indices
                centers[empty.bool()] = dX[far]
            torch.cuda.synchronize(dev)
            eng.check_async(stream.value)
            _capi.check(eng._h, lib.ivfadc_set_centroids_device(eng._h, vp(centers)))
        torch.cuda.synchronize(dev)
    finally:
        eng.close()
    return centers.cpu().numpy(), assign.cpu().numpy().astype(np.int64)


def train_quantizers_device(data, kc: int, k: int, m: int, coarse_maxiter: int = 25,
                            quantization_maxiter: int = 25, seed: int = 0, device: int = 0):
    """train_quantizers with the Lloyd iterations on the GPU (same outputs, same layout)."""
    data = np.ascontiguousarray(data)
    n, D = data.shape
    dsub = D // m
    centroids, assign = kmeans_device(data, kc, coarse_maxiter, seed, device)
    resid = data - centroids[assign]
    cb_vectors = np.empty((m, k, dsub), dtype=data.dtype)
    for i in range(m):
        cb_vectors[i], _ = kmeans_device(np.ascontiguousarray(resid[:, i * dsub:(i + 1) * dsub]), k,
                                         quantization_maxiter, seed + 1 + i, device)
    cb_codes = np.tile(np.arange(k, dtype=np.uint8), (m, 1))
    return centroids, assign, cb_vectors, cb_codes
