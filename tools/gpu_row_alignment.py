#!/usr/bin/env python3
"""All-pairs lists whose (T, P) output rows are multiples of 32 bytes (P % 8 == 0) against lists whose rows are not: the
stores of a warp cover 4 sectors in the first case, 5 (two of them partial) in the second.  python tools/gpu_row_alignment.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import caviar_b200 as cb
def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
T = 20000
for n, m in ((1024, 1), (1024, 2), (1024, 3), (1000, 5), (1008, 1)):
    pairs = cb.all_pairs(n, m)
    P = len(pairs)
    x = (torch.rand((T, n, 3), device="cuda") * 60).contiguous()
    with cb.FeaturePlan(n, pairs, None, pair_cutoff=8.0) as plan:
        out = plan.alloc_outputs(T, want=("dist",))
        ms = timed(lambda: plan.run(x, None, out=out))
        out2 = plan.alloc_outputs(T, want=("dist", "flags", "score"))
        agg = plan.new_aggregates()
        ms2 = timed(lambda: plan.run(x, None, out=out2, aggregates=agg))
    print(json.dumps({"n": n, "minsep": m, "P": P, "row_bytes_mod_32": (4 * P) % 32, "dist_only_evals_per_s": T * P / ms * 1e3, "full_evals_per_s": T * P / ms2 * 1e3}), flush=True)
    del out, out2, x
