"""CAVIAR-1.0's first downstream steps served from the device.

What follows the distance operator in the reference is: the pooled per-pair mean (CAVIAR-1.0.py:342, :378), the
centred matrices ``dist - mu_pool`` (:343, :379-382) and the lagged covariance blocks ``X0.T @ X0 / n`` of the VAMP-2
scan (:396-402).  The means come from the fused kernel's fp64 sums (no pass over (T, P)); centring is fused into the
operand preparation of the covariance kernel; the covariance blocks -- the one dense contraction next to the hot path --
run on the tensor cores through the library's own tcgen05 / TMEM / TMA kernel (``csrc/cv_cov.cu``): every float32
operand is split once into two TF32-exact halves and each block is accumulated as hi.hi + hi.lo + lo.hi in float32
("3xTF32"), which keeps the float32 accuracy of the reference's sgemm.  ``cov_blocks_cublas`` (torch.matmul, float32,
TF32 off) is kept as the comparator of the tests and of ``tools/gpu_cov.py`` only.
Everything here takes and returns CUDA tensors; nothing falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

from . import _lib
from ._lib import check


def _cuda(t, what):
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{what} must be a CUDA tensor (caviar_b200 has no CPU path)")
    return t


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def pooled_mean(sum_a, n_a: int, sum_b, n_b: int):
    """``(distA.mean(0) + distB.mean(0)) / 2`` (CAVIAR-1.0.py:342, :378) from the fp64 ``sum`` aggregates of two runs,
    as float32 like the reference's means."""
    import torch
    sa, sb = _cuda(sum_a, "sum_a"), _cuda(sum_b, "sum_b")
    return ((sa.to(torch.float64) / float(n_a) + sb.to(torch.float64) / float(n_b)) * 0.5).to(torch.float32)


def centred(dist, mu, out=None):
    """``dist - mu`` per column, float32 (CAVIAR-1.0.py:343, :379-382).  ``out=dist`` centres in place.  (The covariance
    path does not need this: ``split_operand(dist, mu)`` subtracts the mean on the fly.)"""
    import torch
    d, m = _cuda(dist, "dist"), _cuda(mu, "mu")
    if d.dim() != 2 or m.shape != (d.shape[1],):
        raise ValueError("dist must be (T, K) and mu (K,)")
    return torch.sub(d, m.to(d.dtype), out=out)


class SplitOperand:
    """An (n, K) float32 matrix prepared for the tensor-core covariance kernel: two (n, Kp) planes with
    ``hi + lo == x - mu`` exactly, ``hi`` rounded to TF32, Kp = K padded to the 128-column tile.  Row views
    (``rows(lo, hi, step)``) share the planes: the lagged pair X[:-tau], X[tau:] of the reference and its even / odd
    frame selections cost no copy."""

    def __init__(self, hi, lo, n_cols: int, row0: int = 0, n_rows: Optional[int] = None, step: int = 1):
        self.hi, self.lo, self.K = hi, lo, int(n_cols)
        self.row0, self.step = int(row0), int(step)
        self.n = int(hi.shape[0] - row0 if n_rows is None else n_rows)

    def rows(self, lo: int, hi: int, step: int = 1) -> "SplitOperand":
        """The view ``self[lo:hi:step]``."""
        if not (0 <= lo <= hi <= self.n) or step < 1:
            raise ValueError("row range out of bounds")
        return SplitOperand(self.hi, self.lo, self.K, self.row0 + lo * self.step, len(range(lo, hi, step)), self.step * step)

    def _ptrs(self):
        off = self.row0 * self.hi.shape[1] * 4
        return self.hi.data_ptr() + off, self.lo.data_ptr() + off


def split_operand(x, mu=None) -> SplitOperand:
    """``x`` (n, K) float32 CUDA (any row stride), optional ``mu`` (K,) to subtract -> TF32 hi / lo planes (one pass)."""
    import torch
    x = _cuda(x, "x")
    if x.dim() != 2 or x.dtype != torch.float32 or x.stride(1) != 1:
        raise ValueError("x must be a (n, K) float32 matrix with contiguous rows")
    n, K = x.shape
    if mu is not None:
        mu = _cuda(mu, "mu").to(torch.float32).contiguous()
        if mu.shape != (K,):
            raise ValueError("mu must have one entry per column")
    Kp = int(_lib.lib.caviar_cov_padded_columns(K))
    hi = torch.empty((n, Kp), dtype=torch.float32, device=x.device)
    lo = torch.empty((n, Kp), dtype=torch.float32, device=x.device)
    check(_lib.lib.caviar_cov_split(x.data_ptr(), x.stride(0) if n > 1 else K, None if mu is None else mu.data_ptr(), n, K,
                                    hi.data_ptr(), lo.data_ptr(), _stream()))
    return SplitOperand(hi, lo, K)


def cov_gemm(a: SplitOperand, b: SplitOperand, out=None, scale: Optional[float] = None, accumulate: bool = False):
    """``out (+)= scale * a.T @ b`` as float32 (K, K) on the tensor cores (tcgen05, 3xTF32); scale defaults to
    ``1 / max(1, n)`` (CAVIAR-1.0.py:398-401)."""
    import torch
    if a.n != b.n or a.K != b.K or a.hi.shape[1] != b.hi.shape[1] or a.step != b.step:
        raise ValueError("operands must have the same shape and row step")
    if a.n <= 0:
        raise ValueError("need at least one row")
    if out is None:
        if accumulate:
            raise ValueError("accumulate needs an output matrix")
        out = torch.empty((a.K, a.K), dtype=torch.float32, device=a.hi.device)
    elif out.shape != (a.K, a.K) or out.dtype != torch.float32 or out.stride(1) != 1 or not out.is_cuda:
        raise ValueError("out must be a (K, K) float32 CUDA matrix with contiguous rows")
    ah, al = a._ptrs()
    bh, bl = b._ptrs()
    check(_lib.lib.caviar_cov_gemm(ah, al, bh, bl, a.n, a.step * a.hi.shape[1], a.K,
                                   1.0 / max(1, a.n) if scale is None else float(scale), 1 if accumulate else 0,
                                   out.data_ptr(), out.stride(0), _stream()))
    return out


def cov_blocks(x0, xt) -> Tuple[object, object, object]:
    """``C00, Ctt, C0t = X0.T @ X0 / n, Xt.T @ Xt / n, X0.T @ Xt / n`` (CAVIAR-1.0.py:396-402) on the tensor cores."""
    a, b = _cuda(x0, "x0"), _cuda(xt, "xt")
    if a.dim() != 2 or a.shape != b.shape:
        raise ValueError("x0 and xt must be (n, K) matrices of the same shape")
    sa, sb = split_operand(a), split_operand(b)
    return cov_gemm(sa, sa), cov_gemm(sb, sb), cov_gemm(sa, sb)


def vamp_cov_blocks(operands, tau: int, parity: Optional[int] = None) -> Tuple[object, object, object]:
    """The covariance blocks exactly as the reference's VAMP-2 scan builds them (CAVIAR-1.0.py:396-421): for every
    ensemble ``X`` in ``operands`` (``SplitOperand`` of the centred matrix) the lagged pair ``X0 = X[:-tau]``,
    ``Xt = X[tau:]``; with ``parity`` 0 / 1 only the pairs whose index has that parity (``_stack_even_odd``); the
    ensembles stacked (``np.vstack``) -- here: accumulated tile by tile -- and divided by the total row count.
    Ensembles with ``len(X) <= tau`` are skipped like in the reference.  Returns (C00, Ctt, C0t), or None when no
    ensemble is long enough."""
    tau = int(tau)
    views = []
    for s in operands:
        if s.n <= tau:
            continue
        n = s.n - tau
        lo, step = (0, 1) if parity is None else (int(parity), 2)
        if len(range(lo, n, step)) == 0:
            continue
        views.append((s.rows(lo, n, step), s.rows(tau + lo, tau + n, step)))
    if not views:
        return None
    total = sum(a.n for a, _ in views)
    out = [None, None, None]
    for k, (a, b) in enumerate(views):
        for slot, (p, q) in enumerate(((a, a), (b, b), (a, b))):
            out[slot] = cov_gemm(p, q, out=out[slot], scale=1.0 / max(1, total), accumulate=k > 0)
    return tuple(out)


def cov_blocks_cublas(x0, xt) -> Tuple[object, object, object]:
    """Comparator only (tests, tools/gpu_cov.py): the same blocks through torch.matmul -> cuBLAS float32, TF32 off."""
    import torch
    a, b = _cuda(x0, "x0"), _cuda(xt, "xt")
    n = max(1, a.shape[0])
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        a32, b32 = a.to(torch.float32), b.to(torch.float32)
        return (a32.T @ a32) / n, (b32.T @ b32) / n, (a32.T @ b32) / n
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
