#!/usr/bin/env python3
"""BASELINE config 4 (all pairs of 1 000 residues, no box) device-resident: full outputs, distances without moments,
contact map (lean), occupancy only.  python tools/gpu_cfg4.py [frames]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import caviar_b200 as cb
from caviar_b200 import synthetic as syn


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    top = syn.make_topology(4)
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    x = syn.device_frames(top, T, only_referenced=True)
    P = len(top.pairs)
    with cb.FeaturePlan(top.cfg.n_atoms, top.pairs, None, box=None, pair_cutoff=8.0) as plan:
        ws = plan.new_workspace()
        res = {"frames": T, "pairs": P, "mode": plan.info["mode_name"]}
        for name, want, moments, agg_on in (("full", ("dist", "flags", "score"), True, True), ("dist_flags_occ", ("dist", "flags", "score"), False, True),
                                            ("dist_only", ("dist",), False, False), ("contact_map", ("flags", "score"), False, True),
                                            ("occupancy_only", (), False, True)):
            out = plan.alloc_outputs(T, want=want)
            agg = plan.new_aggregates(moments=moments) if agg_on else None
            ms = timed(lambda: plan.run(x, None, out=out, aggregates=agg, workspace=ws))
            alg = T * (12 * 1000 + (4 * P if "dist" in want else 0) + (4 * plan.W if "flags" in want else 0) + (4 if "score" in want else 0))
            res[name] = {"ms": ms, "evals_per_s": T * P / ms * 1e3, "hbm_frac": alg / ms / 1e6 / peak}
            del out
        print(json.dumps(res))


if __name__ == "__main__":
    main()
