#!/usr/bin/env python3
"""Time the fused kernel on the cfg3 topology with different box kinds / outputs (GPU box only)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import caviar_b200 as cb
from caviar_b200 import synthetic as syn

T = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
top = syn.make_topology(3)
pairs, triplets = syn.compact_indices(top)
U = len(top.referenced)
raw = syn.device_frames(top, T, only_referenced=True)
res = {}
for name, box, prs, trs in (("tric_full", "triclinic", pairs, triplets), ("tric_pairs", "triclinic", pairs, None),
                            ("tric_near_pairs", "triclinic", pairs[:5000], None), ("tric_rand_pairs", "triclinic", pairs[5000:], None),
                            ("tric_trip", "triclinic", None, triplets), ("ortho_full", "ortho", pairs, triplets),
                            ("none_full", None, pairs, triplets), ("none_pairs", None, pairs, None)):
    plan = cb.FeaturePlan(U, prs, trs, box=box)
    if box == "triclinic":
        b = torch.from_numpy(syn.boxes_for(top, T)).cuda()
    elif box == "ortho":
        b = torch.from_numpy(np.broadcast_to(np.array([100., 100., 100.], np.float32), (T, 3)).copy()).cuda()
    else:
        b = None
    rec = plan.prepare_box(b) if b is not None else None
    staged = plan.stage(raw, rec)
    out = plan.alloc_outputs(T)
    agg = plan.new_aggregates()
    for _ in range(3):
        plan.run(staged, rec, out=out, aggregates=agg)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5):
        plan.run(staged, rec, out=out, aggregates=agg)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    info = plan.info
    S = info["staged_floats_per_frame"] or 3 * U
    streamed = T * (S * 4 + 4 * plan.P + 4 * plan.A + 4 * plan.W + 4)
    res[name] = dict(ms=ms, ms_per_100k=ms * 1e5 / T, streamed_GBs=streamed / ms / 1e6, tiles=info["n_tiles"], stages=info["stages"], fb=info["frames_per_stage"])
    print(name, json.dumps(res[name]), flush=True)
    plan.close(); del staged, out
