#!/usr/bin/env python3
"""Short runs (BASELINE config 1 and 2 and shorter pieces of them): time per pass against the floor of the tiny-run
chunking (GPU box only).  CAVIAR_TINY_CHUNK is read by make_chunks() (cv_api.cu) on every run; 100000 = rule off.
Timed as CUDA-graph replays of one pass (FeaturePlan.capture)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import caviar_b200 as cb
from caviar_b200 import synthetic as syn

peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6557.4
KNOBS = [("default", {}),
         ("tiny8", {"CAVIAR_TINY_CHUNK": "8"}),
         ("tiny_off", {"CAVIAR_TINY_CHUNK": "100000"}),
         ("default_again", {})]
for cfg_id, T in ((1, 1000), (1, 250), (1, 2000), (1, 4000), (1, 6000), (2, 500), (2, 1000), (2, 2000), (2, 10000)):
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
        ws = plan.new_workspace()
        info = plan.info
        print(json.dumps({"config": cfg_id, "tiles": info["n_tiles"], "stages": info["stages"], "fb": info["frames_per_stage"],
                          "n_ctas": info["n_ctas"], "partial_slots": info["partial_slots"]}), flush=True)
        ref = None
        for name, env in KNOBS:
            for k in ("CAVIAR_TINY_CHUNK",):
                os.environ.pop(k, None)
            os.environ.update(env)
            agg.occupancy.zero_(); agg.sum.zero_(); agg.sumsq.zero_()
            plan.run(coords, rec, out=out, aggregates=agg, workspace=ws)
            torch.cuda.synchronize()
            chk = (out.dist.view(torch.int32).sum().item(), int(agg.occupancy.sum().item()), float(agg.sum.sum().item()))
            if ref is None:
                ref = chk
            graph, _ = plan.capture(coords, rec, out=out, aggregates=agg, workspace=ws)     # the knobs are read while recording
            for _ in range(5):
                graph.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 200
            torch.cuda.synchronize(); e0.record()
            for _ in range(reps):
                graph.replay()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            alg = info["algorithmic_bytes_per_frame"] * T
            print(f"  cfg{cfg_id} T={T:<6d} {name:22s} {ms * 1e3:8.1f} us  frac {alg / ms / 1e6 / peak:.3f}  same_dist={chk[0] == ref[0]} same_occ={chk[1] == ref[1]} sum_rel={abs(chk[2] - ref[2]) / abs(ref[2]):.1e}", flush=True)
    del raw, coords, out
