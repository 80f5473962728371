#!/usr/bin/env python3
"""Config 4 (all pairs of 1000 sites, minsep 5 -> P = 495 510, no box) with different output sets and plan modes:
which side bounds it -- the (T,P) store, the L2->shared-memory refill of the coordinates, or instruction issue?"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import caviar_b200 as cb
from caviar_b200 import synthetic as syn

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
top = syn.make_topology(4)
pairs, triplets = syn.compact_indices(top)
U = len(top.referenced)
raw = syn.device_frames(top, T, only_referenced=True)
for force in (None, "staged"):
    plan = cb.FeaturePlan(U, pairs, None, box=top.cfg.box, force=force)
    coords = plan.stage(raw)
    for want, agg_on in ((("dist", "flags", "score"), True), (("flags", "score"), True), (("score",), True), ((), True),
                         (("dist",), False), (("flags",), False), (("score",), False),
                         (("flags", "score"), "occ"), ((), "occ")):       # "occ": occupancy without moments = lean run
        out = plan.alloc_outputs(T, want=want)
        agg = plan.new_aggregates(moments=agg_on != "occ") if agg_on else None
        for _ in range(2):
            plan.run(coords, None, out=out, aggregates=agg)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(5):
            plan.run(coords, None, out=out, aggregates=agg)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        info = plan.info
        print(json.dumps({"mode": info["mode_name"], "tiles": info["n_tiles"], "fb": info["frames_per_stage"], "stages": info["stages"],
                          "smem": info["smem_bytes"], "want": list(want), "agg": agg_on, "ms": round(ms, 3), "ms_per_50k": round(ms * 50000 / T, 2),
                          "evals_per_s": T * plan.P / (ms * 1e-3), "store_GBs": (4 * plan.P * T if "dist" in want else 0) / ms / 1e6}), flush=True)
        del out
    plan.close()
