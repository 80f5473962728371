#!/usr/bin/env python3
"""GPU box: python tests/components_report.py   (lives under tests/ because it uses the oracle as the checker)
f4 (CAVIAR-1.0.py:183-199): residue components of the thresholded mean-distance graph -- device kernel fed by the fused
kernel's fp64 sums, against the reference's algorithm (dict of pair means + Python DFS, oracle port) on the host (GPU box only)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import caviar_b200 as cb
from oracle import caviar_oracle as orc

for n, minsep in ((300, 5), (1000, 5)):
    pairs = cb.all_pairs(n, minsep)
    P = len(pairs)
    rng = np.random.default_rng(n)
    T, cut = 10, 8.0
    mean = rng.uniform(9.0, 50.0, size=P)
    mean[rng.random(P) < 2e-3] = 5.0
    chain = (pairs[:, 1] - pairs[:, 0] == minsep) & (pairs[:, 0] % 3 != 0)
    mean[chain] = 7.5
    with cb.FeaturePlan(n, pairs, None, box=None) as plan:
        sums = torch.from_numpy(mean * T).cuda()
        for _ in range(3):
            lab = plan.residue_components(sums, T, cut)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20):
            lab = plan.residue_components(sums, T, cut)
        e1.record(); torch.cuda.synchronize()
        dev_ms = e0.elapsed_time(e1) / 20
        got = lab.cpu().numpy().tolist()
    m32 = (mean * T / T).astype(np.float32)
    t0 = time.perf_counter()
    want = orc.residue_components(n, {(int(i), int(j)): float(m) for (i, j), m in zip(pairs, m32)}, cut)
    ref_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    vec = cb.residue_components(n, pairs, m32, cut)
    vec_s = time.perf_counter() - t0
    print(json.dumps({"n_res": n, "pairs": P, "components": max(got) + 1, "device_ms": round(dev_ms, 4),
                      "reference_algorithm_python_s": round(ref_s, 3), "host_vectorised_s": round(vec_s, 4),
                      "equal_to_reference_algorithm": got == want, "equal_to_host_vectorised": got == vec}), flush=True)
