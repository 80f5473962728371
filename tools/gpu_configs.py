#!/usr/bin/env python3
"""Throughput of the device-resident path on every BASELINE.json configuration (GPU box only; parity for these
shapes is in tests/).  Frames are capped so that inputs + outputs stay within a few GB."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import caviar_b200 as cb
from caviar_b200 import synthetic as syn

peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
for cfg_id, T in ((1, 1000), (2, 10000), (3, 32768), (4, 2048), (5, 4096)):
    top = syn.make_topology(cfg_id)
    pairs, triplets = syn.compact_indices(top)
    U = len(top.referenced)
    raw = syn.device_frames(top, T, only_referenced=True)
    with cb.FeaturePlan(U, pairs, triplets if len(triplets) else None, box=top.cfg.box) as plan:
        b = syn.boxes_for(top, T)
        rec = plan.prepare_box(torch.from_numpy(b).cuda()) if b is not None else None
        coords = plan.stage(raw, rec)
        out = plan.alloc_outputs(T)
        agg = plan.new_aggregates()
        for _ in range(3):
            plan.run(coords, rec, out=out, aggregates=agg)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        reps = 10
        for _ in range(reps):
            plan.run(coords, rec, out=out, aggregates=agg)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        info = plan.info
        alg = info["algorithmic_bytes_per_frame"] * T
        print(json.dumps({"config": top.cfg.name, "frames": T, "P": plan.P, "A": plan.A, "U": U, "mode": info["mode_name"],
                          "tiles": info["n_tiles"], "stages": info["stages"], "fb": info["frames_per_stage"], "ms": round(ms, 4),
                          "evals_per_s": T * (plan.P + plan.A) / (ms * 1e-3), "alg_GBs": alg / ms / 1e6,
                          "frac_of_measured_peak": alg / ms / 1e6 / peak}), flush=True)
    del raw, coords, out
