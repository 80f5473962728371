#!/usr/bin/env python3
"""Write-only and copy HBM rates on this box (torch fill / copy kernels over 8 GB), next to MEASURED_PEAKS.json's copy
figure: config 4 is a write-dominated workload (4 bytes out per evaluation, ~0.03 in), so its ceiling is the write-only
rate, not the copy rate.  python tools/gpu_write_peak.py"""
import json
import torch


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = 2 << 30                                   # floats: 8 GB
    a = torch.empty(n, dtype=torch.float32, device="cuda")
    b = torch.empty(n, dtype=torch.float32, device="cuda")
    res = {}
    ms = timed(lambda: a.fill_(1.0)); res["fill_8GB_GBps"] = 4 * n / ms / 1e6
    ms = timed(lambda: torch.cuda.current_stream().synchronize() or a.zero_()); res["memset_8GB_GBps"] = 4 * n / ms / 1e6
    ms = timed(lambda: b.copy_(a)); res["copy_8GB_read_plus_write_GBps"] = 8 * n / ms / 1e6
    ms = timed(lambda: a.sum()); res["read_8GB_GBps"] = 4 * n / ms / 1e6
    print(json.dumps(res))


if __name__ == "__main__":
    main()
