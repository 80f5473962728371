#!/usr/bin/env python3
"""Device-timed parts of a BASELINE config: fused kernel on plan-order frames, K0 (caller-order -> plan-order permutation),
and both back to back.  python tools/gpu_probe.py [config] [frames]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import caviar_b200 as cb
from caviar_b200 import synthetic as syn

CUTS = dict(pair_cutoff=8.0, hbond_dist_cutoff=3.0, hbond_angle_cutoff=135.0)


def timed(fn, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    cfg_id = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
    top = syn.make_topology(cfg_id)
    pairs, triplets = syn.compact_indices(top)
    U = len(top.referenced)
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    with cb.FeaturePlan(U, pairs, triplets if len(triplets) else None, box=top.cfg.box, **CUTS) as plan:
        raw = syn.device_frames(top, T, only_referenced=True)
        box = syn.boxes_for(top, T)
        rec = plan.prepare_box(torch.from_numpy(box).cuda()) if box is not None else None
        coords = plan.stage(raw)
        out = plan.alloc_outputs(T)
        agg = plan.new_aggregates()
        ws = plan.new_workspace()
        ms_run = timed(lambda: plan.run(coords, rec, out=out, aggregates=agg, workspace=ws))
        res = {"config": cfg_id, "frames": T, "mode": plan.info["mode_name"], "fused_ms": ms_run}
        alg = plan.info["algorithmic_bytes_per_frame"] * T
        res["fused_frac"] = alg / ms_run / 1e6 / peak
        res["fused_ms_per_100k"] = ms_run * 1e5 / T
        if plan.info["mode_name"] == "staged":
            ms_k0 = timed(lambda: plan.stage(raw, out=coords))
            ms_both = timed(lambda: plan.run(plan.stage(raw, out=coords), rec, out=out, aggregates=agg, workspace=ws))
            res.update(k0_ms=ms_k0, k0_gbs=(12 * U + 4 * plan.info["staged_floats_per_frame"]) * T / ms_k0 / 1e6,
                       caller_order_ms=ms_both, caller_order_frac=alg / ms_both / 1e6 / peak)
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
