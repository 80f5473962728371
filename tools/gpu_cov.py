#!/usr/bin/env python3
"""Covariance blocks (CAVIAR-1.0.py:396-402) on the tensor cores against cuBLAS float32 on the same box:
python tools/gpu_cov.py [n] [K]  ->  one JSON line (times, TFLOP/s of useful 2*n*K*K flops, errors vs float64)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caviar_b200 import downstream as ds


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 14000
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    x = torch.randn((n, K), device="cuda", generator=g) * 0.7 + torch.sin(torch.arange(n, device="cuda")[:, None] * 1e-3)
    x = (x - x.mean(0)).contiguous()
    s = ds.split_operand(x)
    out = torch.empty((K, K), device="cuda")
    ms_split = timed(lambda: ds.split_operand(x))
    ms_ours = timed(lambda: ds.cov_gemm(s, s, out=out))
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    ms_fp32 = timed(lambda: torch.matmul(x.T, x, out=out))
    torch.backends.cuda.matmul.allow_tf32 = True
    ms_tf32 = timed(lambda: torch.matmul(x.T, x, out=out))
    torch.backends.cuda.matmul.allow_tf32 = prev
    flops = 2.0 * n * K * K
    # accuracy on a corner block against float64
    kk = min(K, 512)
    ref = (x[:, :kk].double().T @ x[:, :kk].double() / n).cpu().numpy()
    ours = ds.cov_gemm(s, s).cpu().numpy()[:kk, :kk]
    torch.backends.cuda.matmul.allow_tf32 = False
    lib = ((x.T @ x) / n).cpu().numpy()[:kk, :kk]
    torch.backends.cuda.matmul.allow_tf32 = True
    lib_tf32 = ((x.T @ x) / n).cpu().numpy()[:kk, :kk]
    torch.backends.cuda.matmul.allow_tf32 = prev
    scale = np.sqrt(np.outer(np.diag(ref), np.diag(ref))).max()
    print(json.dumps({"n": n, "K": K, "split_ms": ms_split, "tcgen05_3xtf32_ms": ms_ours, "cublas_fp32_ms": ms_fp32, "cublas_tf32_ms": ms_tf32,
                      "tcgen05_3xtf32_tflops": flops / ms_ours / 1e9, "cublas_fp32_tflops": flops / ms_fp32 / 1e9,
                      "cublas_tf32_tflops": flops / ms_tf32 / 1e9, "tensor_core_tflops_issued": 3 * flops / ms_ours / 1e9,
                      "speedup_vs_cublas_fp32": ms_fp32 / ms_ours,
                      "rel_err_vs_fp64": {"tcgen05_3xtf32": float(np.abs(ours - ref).max() / scale),
                                          "cublas_fp32": float(np.abs(lib - ref).max() / scale),
                                          "cublas_tf32": float(np.abs(lib_tf32 - ref).max() / scale)}}))


if __name__ == "__main__":
    main()
