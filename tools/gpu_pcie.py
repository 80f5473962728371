#!/usr/bin/env python3
"""GPU box: what the PCIe link gives a bare pinned cudaMemcpyAsync loop (the ceiling of bench.py's e2e number)."""
import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
def t_one(chunks, streams):
    ss = [torch.cuda.Stream() for _ in range(streams)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    step = n // chunks
    for i in range(chunks):
        with torch.cuda.stream(ss[i % streams]):
            d[i*step:(i+1)*step].copy_(h[i*step:(i+1)*step], non_blocking=True)
    torch.cuda.synchronize(); return n / (time.perf_counter() - t0) / 1e9
for chunks, streams in ((1,1),(4,1),(4,2),(8,2),(8,4),(16,4)):
    best = max(t_one(chunks, streams) for _ in range(4))
    print(f"H2D {chunks} chunks on {streams} streams: {best:.1f} GB/s")
# D2H concurrently with H2D
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n // 4, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n // 4, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"H2D 1 GiB + D2H 0.25 GiB concurrently: {dt*1e3:.2f} ms -> H2D-equivalent {n/dt/1e9:.1f} GB/s")
