"""Seeded synthetic inputs for tests and the benchmark (harness utility, not on the hot path).

Database = mixture of Gaussian blobs (centres U[0,1)^D, sigma 0.05) so that lists are roughly
balanced and residuals have structure; queries are fresh draws from the same mixture; a plain
U[0,1)^D variant mirrors what the reference's README/tests use (README.md:32, test/index.jl:7).
"""
from __future__ import annotations

import numpy as np


def blobs(n: int, D: int, n_blobs: int, seed: int, dtype=np.float32, sigma: float = 0.05,
          centre_seed: int = 1001, balanced: bool = False):
    centres = np.random.default_rng(centre_seed).random((n_blobs, D))
    rng = np.random.default_rng(seed)
    which = rng.integers(0, n_blobs, size=n)
    if balanced:  # bring-up experiments: every blob gets the same number of points
        which = np.arange(n) % n_blobs
    out = np.empty((n, D), dtype=dtype)
    step = 1 << 18
    for s in range(0, n, step):
        e = min(n, s + step)
        out[s:e] = (centres[which[s:e]] + sigma * rng.standard_normal((e - s, D))).astype(dtype)
    return out


def uniform(n: int, D: int, seed: int, dtype=np.float32):
    return np.random.default_rng(seed).random((n, D)).astype(dtype)


def random_quantizers(kc: int, D: int, m: int, ksub: int, seed: int, dtype=np.float32,
                      data=None, resid_scale: float = 0.05):
    """Cheap stand-in for training when only shapes matter: centroids = random data points (or
    uniform), codewords = Gaussian of the residual scale."""
    rng = np.random.default_rng(seed)
    if data is not None:
        centroids = np.ascontiguousarray(data[rng.choice(len(data), kc, replace=False)]).astype(dtype)
    else:
        centroids = rng.random((kc, D)).astype(dtype)
    dsub = D // m
    cb = (resid_scale * rng.standard_normal((m, ksub, dsub))).astype(dtype)
    codes = np.tile(np.arange(ksub, dtype=np.uint8), (m, 1))
    return centroids, cb, codes


def blob_centres(D: int, n_blobs: int, dtype=np.float32, centre_seed: int = 1001):
    """The mixture centres `blobs` draws around (same generator, same seed)."""
    return np.random.default_rng(centre_seed).random((n_blobs, D)).astype(dtype)


def train_on_device(X, kc: int, m: int, ksub: int, seed: int = 3001, iters: int = 8,
                    sample: int = 262144, init=None):
    """Benchmark trainer: Lloyd on the GPU through torch (plumbing; training is outside the hot
    path).  X: numpy [n, D].  `init` (optional [kc, D]) seeds the coarse Lloyd iterations -- the
    benchmark passes the mixture centres, which makes the cells the (balanced) blobs, as a
    converged k-means on this data would.  Returns numpy centroids [kc, D], codebooks
    [m, ksub, dsub]."""
    import torch

    dev = torch.device("cuda")
    n, D = X.shape
    rng = np.random.default_rng(seed)
    sel = rng.choice(n, min(n, max(sample, 64 * kc)), replace=False)
    xs = torch.from_numpy(np.ascontiguousarray(X[sel])).to(dev).float()
    cent, cb = train_on_device_tensor(xs, kc, m, ksub, seed, iters,
                                      init=None if init is None else torch.from_numpy(init).to(dev))
    return cent.cpu().numpy().astype(X.dtype), cb.cpu().numpy().astype(X.dtype)


def train_on_device_tensor(xs, kc: int, m: int, ksub: int, seed: int = 3001, iters: int = 8, init=None):
    """Same trainer on a sample that already lives on the device (torch float32 [ns, D]); returns device tensors
    (centroids [kc, D], codebooks [m, ksub, dsub])."""
    import torch

    from .training import kmeans_torch

    ns, D = xs.shape
    chunk = int(min(1 << 18, max(4096, (1 << 30) // kc)))   # bound the ns x kc distance block to ~4 GB
    cent = kmeans_torch(xs, kc, iters, seed, chunk=chunk, init=init)
    a = torch.empty(ns, dtype=torch.long, device=xs.device)
    cn = (cent ** 2).sum(1)
    for s in range(0, ns, chunk):
        a[s:s + chunk] = (cn[None, :] - 2.0 * xs[s:s + chunk] @ cent.T).argmin(1)
    resid = xs - cent[a]
    dsub = D // m
    cbs = []
    for i in range(m):
        cbs.append(kmeans_torch(resid[:, i * dsub:(i + 1) * dsub].contiguous(), ksub, iters, seed + 1 + i))
    return cent, torch.stack(cbs)


# ---- counter-based generator on the device (csrc/synth.cu; CPU twin: oracle.synth_*) ---------------------------
def blob_scale(sigma: float = 0.05):
    """sigma * sqrt(3) / 2^22 as float32: the factor that turns the centred sum of four 22-bit uniforms into noise
    of standard deviation sigma (passed to both generators so that they multiply by the same float)."""
    return np.float32(np.float32(sigma) * np.float32(1.7320508075688772) / np.float32(4194304.0))


def uniform_device(first: int, n: int, D: int, seed: int, device=None):
    """torch float32 [n, D] in [0, 1): vectors first .. first + n - 1 of the stream `seed`."""
    import ctypes

    import torch

    from . import _capi
    out = torch.empty((n, D), dtype=torch.float32, device=device or torch.device("cuda", torch.cuda.current_device()))
    rc = _capi.load().ivfadc_synth_uniform_device(ctypes.c_void_p(out.data_ptr()), first, n, D, seed,
                                                  ctypes.c_void_p(torch.cuda.current_stream(out.device).cuda_stream))
    if rc != 0:
        raise _capi.IvfadcError(rc, "ivfadc_synth_uniform_device")
    return out


def blobs_device(first: int, n: int, centres, seed: int, sigma: float = 0.05, out=None, want_blobs: bool = False):
    """torch float32 [n, D]: vectors first .. first + n - 1 of the blob mixture around `centres` (device tensor
    [n_blobs, D]); `out` reuses a buffer of at least n rows."""
    import ctypes

    import torch

    from . import _capi
    n_blobs, D = centres.shape
    if out is None:
        out = torch.empty((n, D), dtype=torch.float32, device=centres.device)
    x = out[:n]
    blobs = torch.empty(n, dtype=torch.int32, device=centres.device) if want_blobs else None
    rc = _capi.load().ivfadc_synth_blobs_device(
        ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(blobs.data_ptr() if want_blobs else None), first, n, D, n_blobs,
        seed, ctypes.c_float(float(blob_scale(sigma))), ctypes.c_void_p(centres.data_ptr()),
        ctypes.c_void_p(torch.cuda.current_stream(centres.device).cuda_stream))
    if rc != 0:
        raise _capi.IvfadcError(rc, "ivfadc_synth_blobs_device")
    return (x, blobs) if want_blobs else x
