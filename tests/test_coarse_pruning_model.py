"""CPU model of the tensor-core coarse step's pruning (ivfadc.jl_b200/csrc/coarse_tc.cuh): TF32-rounded scores,
the group-minima bound and the 2E margin, checked against the oracle's exact top-w -- the claim the kernel's
correctness rests on ("the candidates are a superset of the exact top-w for any data") and the error bound
E = 2^-10 (|q|^2 + max |c|^2) it is derived from.  The GPU tests check the kernel itself bit for bit; this one
pins the arithmetic of the bound where no GPU is needed."""
import numpy as np
import pytest

from oracle import oracle as orc


def tf32(x):
    """cvt.rna.tf32.f32: round to 10 explicit mantissa bits, ties away from zero."""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def make(kind, kc, D, nq, rng):
    if kind == "uniform":
        return rng.random((kc, D)).astype(np.float32), rng.random((nq, D)).astype(np.float32)
    if kind == "blobs":
        c = rng.random((kc, D)).astype(np.float32)
        return c, (c[rng.integers(0, kc, nq)] + 0.05 * rng.standard_normal((nq, D))).astype(np.float32)
    if kind == "large_norm":
        return (100.0 + rng.random((kc, D))).astype(np.float32), (100.0 + rng.random((nq, D))).astype(np.float32)
    c = (0.001 * rng.standard_normal((kc, D)) + 3.0).astype(np.float32)      # "tight": distances << norms
    return c, (0.001 * rng.standard_normal((nq, D)) + 3.0).astype(np.float32)


@pytest.mark.parametrize("kind", ["uniform", "blobs", "large_norm", "tight"])
@pytest.mark.parametrize("kc,D,w", [(1024, 128, 16), (512, 96, 32), (300, 64, 8), (2048, 32, 1)])
def test_candidates_contain_the_exact_top_w(kind, kc, D, w):
    rng = np.random.default_rng(kc + D + w)
    nq = 200
    C, Q = make(kind, kc, D, nq, rng)
    qz = orc.Quantizers(C, np.zeros((1, 4, D), dtype=np.float32))
    ocells, odist = orc.coarse_search(qz, Q, w, nthreads=4)          # exact direct form, stable order
    # what the kernel computes: TF32 operands, -2c exact scaling, norm as hi + lo TF32 pieces, fp32 accumulate
    cn = (C.astype(np.float32) ** 2).sum(1, dtype=np.float32)
    cn_hi = tf32(cn)
    cn_s = cn_hi.astype(np.float64) + tf32(cn - cn_hi).astype(np.float64)
    S = (cn_s[None, :] - 2.0 * (tf32(Q).astype(np.float64) @ tf32(C).astype(np.float64).T)).astype(np.float32)
    qn = (Q.astype(np.float64) ** 2).sum(1)
    cmax2 = float(cn.max())
    # (a) the error bound the margin is built on
    d_all = ((C[None, :, :].astype(np.float64) - Q[:, None, :].astype(np.float64)) ** 2).sum(2)
    err = np.abs(S.astype(np.float64) + qn[:, None] - d_all)
    E = 2.0 ** -10 * (qn + cmax2)
    assert (err.max(1) <= E).all(), float((err.max(1) / E).max())
    # (b) the bound: WL-th smallest minimum of groups of 8 columns in the first two tiles of 256, of 16 afterwards
    WL = 1 if w <= 1 else 8 if w <= 8 else 16 if w <= 16 else 32
    kcp = -(-kc // 256) * 256
    Sp = np.full((nq, kcp), np.inf, dtype=np.float32)
    Sp[:, :kc] = S
    fine = Sp[:, :512].reshape(nq, -1, 8).min(2)
    mins = np.concatenate([fine, Sp[:, 512:].reshape(nq, -1, 16).min(2)], axis=1) if kcp > 512 else fine
    assert mins.shape[1] >= WL
    B = np.sort(mins, axis=1)[:, WL - 1]
    cut = B.astype(np.float64) + 2.0 ** -8 * (qn + cmax2)
    # (c) superset, and a useful one
    cand = S.astype(np.float64) <= cut[:, None]
    assert cand[np.arange(nq)[:, None], ocells].all()
    if kind in ("uniform", "blobs"):
        assert cand.sum(1).mean() <= 4 * WL + 8
