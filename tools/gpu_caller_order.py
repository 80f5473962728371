#!/usr/bin/env python3
"""Config 3 frames compact in the CALLER's atom order, resident in HBM: K0 + fused kernel against the direct-gather
kernel on the same frames (GPU box only)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import caviar_b200 as cb
from caviar_b200 import synthetic as syn

T = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
top = syn.make_topology(3)
pairs, triplets = syn.compact_indices(top)
U = len(top.referenced)
raw = syn.device_frames(top, T, only_referenced=True)
box = torch.from_numpy(syn.boxes_for(top, T)).cuda()


def timed(fn, n=5):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = {}
ref = None
for name, force in (("k0_plus_fused", None), ("direct_gather", "direct")):
    kw = {} if force is None else {"force": force}
    with cb.FeaturePlan(U, pairs, triplets, box="triclinic", **kw) as plan:
        rec = plan.prepare_box(box)
        out = plan.alloc_outputs(T)
        agg = plan.new_aggregates()
        ws = plan.new_workspace()
        if plan.info["mode_name"] == "staged":
            staged = torch.empty((T, int(plan.info["staged_floats_per_frame"])), dtype=torch.float32, device="cuda")
            fn = lambda: plan.run(plan.stage(raw, out=staged), rec, out=out, aggregates=agg, workspace=ws)
        else:
            fn = lambda: plan.run(raw, rec, out=out, aggregates=agg, workspace=ws)
        ms = timed(fn)
        alg = int(plan.info["algorithmic_bytes_per_frame"]) * T
        res[name] = {"mode": plan.info["mode_name"], "ms": ms, "ms_per_100k": ms * 1e5 / T, "alg_GBs": alg / ms / 1e6}
        d = out.dist.clone()
        if ref is None:
            ref = d
        else:
            res[name]["dist_bit_equal"] = bool(torch.equal(ref.view(torch.int32), d.view(torch.int32)))
        print(name, json.dumps(res[name]), flush=True)
        del out, agg, ws
