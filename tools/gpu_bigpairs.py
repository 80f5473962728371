#!/usr/bin/env python3
"""All pairs of n atoms (minsep 5, no box) for n beyond the small-frame regime: which plan mode, how fast."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import caviar_b200 as cb

for n, T in ((1000, 2048), (3000, 256), (5000, 96), (9000, 32)):
    ii, jj = np.triu_indices(n, 5)
    pairs = np.stack([ii, jj], 1).astype(np.int32)
    x = (torch.rand((T, n, 3), device="cuda") * 80.0).contiguous()
    for force in (None, "direct"):
        with cb.FeaturePlan(n, pairs, None, box=None, force=force) as plan:
            coords = plan.stage(x)
            out = plan.alloc_outputs(T, want=("dist", "flags", "score"))
            agg = plan.new_aggregates()
            for _ in range(2):
                plan.run(coords, None, out=out, aggregates=agg)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(3):
                plan.run(coords, None, out=out, aggregates=agg)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            info = plan.info
            print(json.dumps({"n": n, "P": len(pairs), "T": T, "force": force, "mode": info["mode_name"], "ctas": info["n_ctas"], "smem": info["smem_bytes"],
                              "ms": round(ms, 3), "evals_per_s": T * len(pairs) / (ms * 1e-3), "store_GBs": 4 * len(pairs) * T / ms / 1e6}), flush=True)
        del out
