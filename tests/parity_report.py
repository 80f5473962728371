#!/usr/bin/env python3
"""Parity report (run on the GPU box):  python tests/parity_report.py > profiles/rNN_parity_report.json

For every BASELINE.json configuration (frame subsets the fp64 oracle finishes in seconds) the CUDA path is compared
with the CPU oracle: worst distance / angle deviation, and every contact / H-bond flag that differs -- each one is
listed with its oracle value so that it can be seen to lie within tolerance of the cut-off (north_star: "pairs
within that tolerance of a cutoff ... are enumerated in the report").
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import caviar_b200 as cb                      # noqa: E402
from caviar_b200 import synthetic as syn      # noqa: E402
from oracle import caviar_oracle as orc       # noqa: E402

CUTS = (8.0, 3.0, 135.0)
DIST_TOL, ANGLE_TOL = 1e-4, 1e-3
CASES = [(1, 1000, False), (2, 400, False), (3, 300, True), (4, 16, False), (5, 48, True)]


def main():
    report = {"tolerances": {"distance_A": DIST_TOL, "angle_deg": ANGLE_TOL}, "cutoffs": CUTS, "cases": []}
    for cfg_id, T, compact in CASES:
        top = syn.make_topology(cfg_id)
        x = syn.parity_frames(top, T, only_referenced=compact)
        pairs, trips = syn.compact_indices(top) if compact else (top.pairs, top.triplets)
        box = syn.boxes_for(top, T)
        with cb.FeaturePlan(x.shape[1], pairs, trips if len(trips) else None, box=top.cfg.box, pair_cutoff=CUTS[0],
                            hbond_dist_cutoff=CUTS[1], hbond_angle_cutoff=CUTS[2]) as plan:
            res = plan.features_host(x, box)
            mode = plan.info["mode_name"]
        P, A = len(pairs), len(trips)
        d64 = orc.pbc_pair_distances(x, pairs, box)
        case = {"config": top.cfg.name, "frames": T, "pairs": P, "triplets": A, "box": top.cfg.box or "none",
                "mode": mode, "input": "compact (T,U,3)" if compact else "full (T,N,3)"}
        case["max_abs_distance_error_A"] = float(np.abs(res["dist"] - d64).max())
        if top.cfg.box is None:
            ref = orc.pair_distances(x, [tuple(p) for p in np.asarray(pairs).tolist()], n_jobs=4)
            case["distances_bit_exact_vs_reference_algorithm"] = bool(res["dist"].tobytes() == ref.tobytes())
        fp, ft = cb.unpack_flags(res["flags"], P, A)
        want_p = orc.contact_flags(d64, CUTS[0])
        diff = np.argwhere(fp != want_p)
        case["pair_flags_differing"] = int(len(diff))
        case["pair_flags_differing_list"] = [
            {"frame": int(t), "pair": int(k), "d_oracle_A": float(d64[t, k]), "d_device_A": float(res["dist"][t, k]),
             "within_tolerance_of_cutoff": bool(abs(d64[t, k] - CUTS[0]) <= DIST_TOL)} for t, k in diff[:200]]
        case["pair_flags_outside_tolerance"] = int(sum(abs(d64[t, k] - CUTS[0]) > DIST_TOL for t, k in diff))
        case["pairs_within_tolerance_of_cutoff"] = int(orc.near_cutoff(d64, CUTS[0], DIST_TOL).sum())
        if A:
            a64 = orc.triplet_angles(x, trips, box)
            dac = orc.triplet_end_distances(x, trips, box)
            case["max_abs_angle_error_deg"] = float(np.abs(res["angle"] - a64).max())
            want_t = orc.hbond_flags(dac, a64, CUTS[1], CUTS[2])
            dt = np.argwhere(ft != want_t)
            case["hbond_flags_differing"] = int(len(dt))
            case["hbond_flags_differing_list"] = [
                {"frame": int(t), "triplet": int(k), "d_ac_oracle_A": float(dac[t, k]), "angle_oracle_deg": float(a64[t, k]),
                 "within_tolerance_of_cutoff": bool(abs(dac[t, k] - CUTS[1]) <= DIST_TOL or abs(a64[t, k] - CUTS[2]) <= ANGLE_TOL)}
                for t, k in dt[:200]]
            case["hbond_flags_outside_tolerance"] = int(sum(
                not (abs(dac[t, k] - CUTS[1]) <= DIST_TOL or abs(a64[t, k] - CUTS[2]) <= ANGLE_TOL) for t, k in dt))
        flags_all = np.concatenate([fp, ft], axis=1)
        case["occupancy_equals_flag_sums"] = bool(np.array_equal(res["occupancy"], flags_all.sum(0)))
        case["scores_equal_flag_sums"] = bool(np.array_equal(res["score"], flags_all.sum(1)))
        report["cases"].append(case)
    json.dump(report, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
