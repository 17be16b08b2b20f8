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

    from .training import kmeans_torch

    dev = torch.device("cuda")
    n, D = X.shape
    rng = np.random.default_rng(seed)
    sel = rng.choice(n, min(n, max(sample, 64 * kc)), replace=False)
    xs = torch.from_numpy(np.ascontiguousarray(X[sel])).to(dev).float()
    cent = kmeans_torch(xs, kc, iters, seed, init=None if init is None else torch.from_numpy(init).to(dev))
    a = torch.empty(xs.shape[0], dtype=torch.long, device=dev)
    cn = (cent ** 2).sum(1)
    for s in range(0, xs.shape[0], 1 << 16):
        a[s:s + (1 << 16)] = (cn[None, :] - 2.0 * xs[s:s + (1 << 16)] @ cent.T).argmin(1)
    resid = xs - cent[a]
    dsub = D // m
    cbs = []
    for i in range(m):
        cbs.append(kmeans_torch(resid[:, i * dsub:(i + 1) * dsub].contiguous(), ksub, iters, seed + 1 + i))
    cb = torch.stack(cbs)
    return cent.cpu().numpy().astype(X.dtype), cb.cpu().numpy().astype(X.dtype)
