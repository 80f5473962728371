"""f4 on the device: thresholded mean-distance graph -> residue components (CAVIAR-1.0.py:183-199), fed by the fused
kernel's fp64 sums, against label vectors produced by the reference's own build_residue_components (tests/golden)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    import __graft_entry__ as g
    g.build()
    import caviar_b200
    return caviar_b200


def test_device_components_match_reference_vectors(cb, golden_dir):
    import torch
    gold = np.load(os.path.join(golden_dir, "distances.npz"))
    with open(os.path.join(golden_dir, "reference_cases.json")) as fh:
        agg_meta = json.load(fh)["aggregates"]
    n, m = agg_meta["n_res"], agg_meta["minsep"]
    pairs = cb.all_pairs(n, m)
    x = torch.from_numpy(gold["B_x"]).cuda()                           # the frames the golden distances came from
    with cb.FeaturePlan(n, pairs, None, box=None) as plan:
        agg = plan.new_aggregates()
        out = plan.run(x, aggregates=agg)
        assert out.dist.cpu().numpy().tobytes() == gold["B_d"].tobytes()
        T = x.shape[0]
        means = (agg.sum[:len(pairs)].cpu().numpy() / T).astype(np.float32)
        checked = 0
        for cut_repr, want in agg_meta["components"].items():
            cut = float(cut_repr)
            # the reference thresholds ITS float32-accumulated mean; skip cut-offs within float32 rounding of one
            if np.any(np.abs(means.astype(np.float64) - cut) < 4e-6 * cut):
                continue
            got = plan.residue_components(agg.sum, T, cut).cpu().numpy().tolist()
            assert got == want, cut_repr
            assert got == cb.residue_components(n, pairs, means, cut)
            checked += 1
        assert checked >= 3
        # pooled run (:496): 0.5 * (meanA + meanB) in float32, here with the same ensemble twice scaled differently
        agg_b = plan.new_aggregates()
        plan.run((x * 1.5).contiguous(), aggregates=agg_b)
        mean_b = (agg_b.sum[:len(pairs)].cpu().numpy() / T).astype(np.float32)
        pooled = np.float32(0.5) * (means + mean_b)
        for cut in (10.0, 15.0, 22.0):
            got = plan.residue_components(agg.sum, T, cut, sum_b=agg_b.sum, n_frames_b=T).cpu().numpy().tolist()
            assert got == cb.residue_components(n, pairs, pooled, cut)


def test_device_components_on_a_large_sparse_graph(cb):
    """1 000 residues, 495 510 pairs (the config-4 list): chains, isolated residues and merged blobs, many sweeps."""
    import torch
    n = 1000
    pairs = cb.all_pairs(n, 5)
    rng = np.random.default_rng(5)
    with cb.FeaturePlan(n, pairs, None, box=None) as plan:
        # synthetic sums: T = 10, mean below the cut for a random 0.05 % of the pairs plus a long chain i -- i+5
        T, cut = 10, 8.0
        mean = rng.uniform(9.0, 50.0, size=len(pairs))
        mean[rng.random(len(pairs)) < 5e-4] = 5.0
        chain = (pairs[:, 1] - pairs[:, 0] == 5) & (pairs[:, 0] >= 300) & (pairs[:, 0] < 700)
        mean[chain] = 7.5
        sums = torch.from_numpy(mean * T).cuda()
        got = plan.residue_components(sums, T, cut).cpu().numpy().tolist()
        want = cb.residue_components(n, pairs, (mean * T / T).astype(np.float32), cut)
        assert got == want and max(got) > 100
        # strictness: a mean exactly at the cut-off is NOT an edge (CAVIAR-1.0.py:186)
        sums2 = torch.full((len(pairs),), cut * T, dtype=torch.float64, device="cuda")
        assert plan.residue_components(sums2, T, cut).cpu().numpy().tolist() == list(range(n))
        sums2 -= 1e-3
        assert plan.residue_components(sums2, T, cut).cpu().numpy().max() == 0      # every pair an edge: one component
