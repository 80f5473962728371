"""CAVIAR-1.0's first downstream steps served from the device.

What follows the distance operator in the reference is: the pooled per-pair mean (CAVIAR-1.0.py:342, :378), the
centred matrices ``dist - mu_pool`` (:343, :379-380) and the lagged covariance blocks ``X0.T @ X0 / n`` of the VAMP-2
scan (:396-401).  The means come from the fused kernel's fp64 sums (no pass over (T, P)); centring is one elementwise
pass over a matrix that is already resident; the covariance blocks are PLAIN dense contractions and go to cuBLAS
through ``torch.matmul`` in float32 with TF32 off (the reference multiplies float32 matrices with numpy's sgemm) --
a library GEMM, deliberately not a hand-written kernel: nothing is fused into it.
Everything here takes and returns CUDA tensors; nothing falls back to the CPU.
"""
from __future__ import annotations

from typing import Tuple


def _cuda(t, what):
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{what} must be a CUDA tensor (caviar_b200 has no CPU path)")
    return t


def pooled_mean(sum_a, n_a: int, sum_b, n_b: int):
    """``(distA.mean(0) + distB.mean(0)) / 2`` (CAVIAR-1.0.py:342, :378) from the fp64 ``sum`` aggregates of two runs,
    as float32 like the reference's means."""
    import torch
    sa, sb = _cuda(sum_a, "sum_a"), _cuda(sum_b, "sum_b")
    return ((sa.to(torch.float64) / float(n_a) + sb.to(torch.float64) / float(n_b)) * 0.5).to(torch.float32)


def centred(dist, mu, out=None):
    """``dist - mu`` per column, float32 (CAVIAR-1.0.py:343, :379-380).  ``out=dist`` centres in place."""
    import torch
    d, m = _cuda(dist, "dist"), _cuda(mu, "mu")
    if d.dim() != 2 or m.shape != (d.shape[1],):
        raise ValueError("dist must be (T, K) and mu (K,)")
    return torch.sub(d, m.to(d.dtype), out=out)


def cov_blocks(x0, xt) -> Tuple[object, object, object]:
    """``C00, Ctt, C0t = X0.T @ X0 / n, Xt.T @ Xt / n, X0.T @ Xt / n`` (CAVIAR-1.0.py:396-401) on the device."""
    import torch
    a, b = _cuda(x0, "x0"), _cuda(xt, "xt")
    if a.dim() != 2 or a.shape != b.shape:
        raise ValueError("x0 and xt must be (n, K) matrices of the same shape")
    n = max(1, a.shape[0])
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False           # float32 products like numpy's sgemm, not TF32
    try:
        a32, b32 = a.to(torch.float32), b.to(torch.float32)
        return (a32.T @ a32) / n, (b32.T @ b32) / n, (a32.T @ b32) / n
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
