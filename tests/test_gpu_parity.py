"""GPU parity: the sm_100a path (through the C-ABI) against the CPU oracle and the reference's golden vectors.

Tolerances (BASELINE.json north_star): no-box float32 distances BIT-EXACT (they reproduce the reference's own
arithmetic); periodic distances within 1e-4 A and angles within 1e-3 deg of the fp64 oracle; contact flags,
occupancy counts and frame scores bit-exact except for features within that tolerance of a cut-off, which the
tests enumerate.
"""
import os

import numpy as np
import pytest

from oracle import caviar_oracle as orc

pytestmark = pytest.mark.gpu

DIST_TOL = 1e-4      # Angstrom
ANGLE_TOL = 1e-3     # degrees


@pytest.fixture(scope="module")
def cb():
    import caviar_b200
    return caviar_b200


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "distances.npz"))


def _bits(a):
    return np.ascontiguousarray(a).tobytes()


# ---------------------------------------------------------------------------------------------------------
# 1. the reference operator: compute_distances_parallel (CAVIAR-1.0.py:110-122)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["A", "B", "C", "D"])
def test_compute_distances_parallel_golden_bit_exact(cb, gold, case):
    pairs = [(int(a), int(b)) for a, b in gold[f"{case}_pairs"]]
    got = cb.compute_distances_parallel(gold[f"{case}_x"], pairs, n_jobs=3, chunk_size=97)
    want = gold[f"{case}_d"]
    assert got.dtype == np.float32 and got.shape == want.shape
    assert _bits(got) == _bits(want)


def test_compute_distances_parallel_float64_input(cb, gold):
    pairs = [(int(a), int(b)) for a, b in gold["E_pairs"]]
    got = cb.compute_distances_parallel(gold["E_x"], pairs)
    assert got.dtype == np.float64                      # reference: float64 in, float64 out
    assert _bits(got) == _bits(gold["E_d"])             # float64 kernel, same operation order: bit-exact
    rng = np.random.default_rng(12)
    x = rng.normal(0, 30, size=(9, 41, 3))
    pr = [(int(a), int(b)) for a, b in rng.integers(-41, 41, size=(500, 2))]
    assert _bits(cb.compute_distances_parallel(x, pr)) == _bits(orc.pair_distances(x, pr, n_jobs=1))


def test_compute_distances_parallel_errors(cb, gold):
    x = gold["C_x"]
    with pytest.raises(ValueError):
        cb.compute_distances_parallel(x, [])
    with pytest.raises(IndexError):
        cb.compute_distances_parallel(x, [(0, 17)])
    with pytest.raises(IndexError):
        cb.compute_distances_parallel(x, [(-18, 0)])


def test_known_answers(cb):
    x = np.zeros((1, 4, 3), np.float32)
    x[0, 1] = (3, 4, 0)
    x[0, 2] = (3, 4, 12)
    d = cb.compute_distances_parallel(x, [(0, 1), (0, 2), (1, 1)])
    assert d.tolist() == [[5.0, 13.0, 0.0]]


@pytest.mark.parametrize("n,T,force", [(300, 257, None), (300, 64, "direct"), (300, 64, "staged"),
                                       (297, 33, None), (64, 1, None), (1000, 40, None)])
def test_all_pairs_every_mode_bit_exact(cb, n, T, force):
    """The reference's own pair list (all i, j >= i+5) in zero-copy / staged / direct mode."""
    rng = np.random.default_rng(n + T)
    x = (np.cumsum(rng.normal(0, 3.8, size=(n, 3)), axis=0)[None] + rng.normal(0, 0.5, size=(T, n, 3))).astype(np.float32)
    pairs = orc.all_pairs(n, 5)
    ii = np.array([p[0] for p in pairs]); jj = np.array([p[1] for p in pairs])
    want = orc.pair_distances_f32_steps(x, ii, jj)
    with cb.FeaturePlan(n, pairs, None, box=None, force=force) as plan:
        res = plan.features_host(x, want=("dist",))
    assert _bits(res["dist"]) == _bits(want)
    if force is None and n <= 300:
        assert _bits(cb.compute_distances_parallel(x, pairs)) == _bits(orc.pair_distances(x, pairs, n_jobs=2))


# ---------------------------------------------------------------------------------------------------------
# 2. full feature set against the oracle
# ---------------------------------------------------------------------------------------------------------
def _oracle_features(x, box, pairs, triplets, cuts):
    """fp64 oracle: distances, angles, end distances, flags."""
    pc, hd, ha = cuts
    d = orc.pbc_pair_distances(x, pairs, box) if len(pairs) else np.zeros((x.shape[0], 0))
    if len(triplets):
        ang = orc.triplet_angles(x, triplets, box)
        dac = orc.triplet_end_distances(x, triplets, box)
    else:
        ang = dac = np.zeros((x.shape[0], 0))
    return d, ang, dac


def _check_features(cb, res, x, box, pairs, triplets, cuts, exact_dist=False):
    """Compare one host result dict with the oracle; returns the number of near-cut-off features skipped."""
    pc, hd, ha = cuts
    T, P, A = x.shape[0], len(pairs), len(triplets)
    d64, a64, dac64 = _oracle_features(x, box, pairs, triplets, cuts)
    if P:
        if exact_dist:
            ii, jj = np.asarray(pairs)[:, 0], np.asarray(pairs)[:, 1]
            assert _bits(res["dist"]) == _bits(orc.pair_distances_f32_steps(x, ii, jj))
        err = np.abs(res["dist"].astype(np.float64) - d64)
        assert err.max() <= DIST_TOL, f"distance error {err.max():.3g} A"
    amb = np.zeros((T, A), dtype=bool)
    if A:
        # An arm within a rounding error of a Voronoi face of the lattice has two images of the same length to
        # within the distance tolerance: its direction -- and the angle -- is not defined at float32 resolution
        # (one such arm in 2e7 of the random-shape sweep).  Those are set aside, and there must be next to none.
        amb = orc.triplet_arm_gaps(x, triplets, box) < DIST_TOL
        if amb.any():       # count ambiguous (frame, middle atom, end atom) arms once, however many triplets share them
            tri = np.asarray(triplets, dtype=np.int64)
            n_at = x.shape[1]
            arms = set()
            for t, k in np.argwhere(amb):
                arms.add((t, tri[k, 1] * n_at + tri[k, 0])); arms.add((t, tri[k, 1] * n_at + tri[k, 2]))
            assert len(arms) <= max(4, 2e-3 * T * A), f"{len(arms)} image-ambiguous arms"
        err = np.abs(res["angle"].astype(np.float64) - a64)
        assert err[~amb].max(initial=0.0) <= ANGLE_TOL, f"angle error {err[~amb].max():.3g} deg"
    fp, ft = cb.unpack_flags(res["flags"], P, A)
    # the flags the device reports must equal the double-compare of ITS OWN float32 features (bit-exact) ...
    if P:
        assert np.array_equal(fp, orc.contact_flags(res["dist"], pc))
    # ... and equal the oracle's flags except within tolerance of a cut-off
    want_p = orc.contact_flags(d64, pc)
    near_p = orc.near_cutoff(d64, pc, DIST_TOL)
    assert np.array_equal(fp[~near_p], want_p[~near_p])
    want_t = orc.hbond_flags(dac64, a64, hd, ha)
    near_t = orc.near_cutoff(dac64, hd, DIST_TOL) | orc.near_cutoff(a64, ha, ANGLE_TOL) | amb
    assert np.array_equal(ft[~near_t], want_t[~near_t])
    # integer aggregates are exact functions of the flags
    flags_all = np.concatenate([fp, ft], axis=1)
    assert np.array_equal(res["score"], orc.frame_scores(flags_all))
    assert np.array_equal(res["occupancy"], orc.occupancy(flags_all))
    # fp64 moment sums of the float32 features
    feats = np.concatenate([res["dist"] if P else np.zeros((T, 0), np.float32),
                            res["angle"] if A else np.zeros((T, 0), np.float32)], axis=1)
    s1, s2 = orc.moment_sums(feats)
    np.testing.assert_allclose(res["sum"], s1, rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(res["sumsq"], s2, rtol=1e-12, atol=1e-9)
    # the reference's per-pair time mean (CAVIAR-1.0.py:342) accumulates in float32 down a strided axis (no
    # pairwise blocking), so IT carries ~T*eps/4 relative error; our fp64 sum is the accurate side.
    if P:
        np.testing.assert_allclose(res["sum"][:P] / T, orc.pair_time_mean(res["dist"]), rtol=max(1e-6, 3e-8 * T))
    return int(near_p.sum() + near_t.sum())


CUTS = (8.0, 3.0, 135.0)


def _run_cfg(cb, cfg_id, T, force, n_pairs=None, n_triplets=None, compact=False, frames_per_block=0):
    from caviar_b200 import synthetic as syn
    top = syn.make_topology(cfg_id, n_pairs=n_pairs, n_triplets=n_triplets)
    x = syn.parity_frames(top, T, only_referenced=compact)
    pairs, triplets = syn.compact_indices(top) if compact else (top.pairs, top.triplets)
    box = syn.boxes_for(top, T)
    with cb.FeaturePlan(x.shape[1], pairs, triplets, box=top.cfg.box, pair_cutoff=CUTS[0],
                        hbond_dist_cutoff=CUTS[1], hbond_angle_cutoff=CUTS[2], force=force) as plan:
        res = plan.features_host(x, box, frames_per_block=frames_per_block)
        info = plan.info
    return top, x, box, pairs, triplets, res, info


@pytest.mark.parametrize("force", [None, "direct", "staged"])
def test_cfg1_full(cb, force):
    top, x, box, pairs, triplets, res, info = _run_cfg(cb, 1, 1000, force)
    _check_features(cb, res, x, box, pairs, triplets, CUTS, exact_dist=True)


@pytest.mark.parametrize("force,T,fpb", [(None, 300, 0), ("direct", 120, 0), (None, 301, 64)])
def test_cfg2_ortho(cb, force, T, fpb):
    top, x, box, pairs, triplets, res, info = _run_cfg(cb, 2, T, force, frames_per_block=fpb)
    near = _check_features(cb, res, x, box, pairs, triplets, CUTS)
    assert near < 0.01 * res["flags"].size * 32


@pytest.mark.parametrize("force,T,compact", [(None, 48, False), ("direct", 32, False), (None, 200, True)])
def test_cfg3_triclinic(cb, force, T, compact):
    top, x, box, pairs, triplets, res, info = _run_cfg(cb, 3, T, force, compact=compact)
    assert info["mode_name"] == ("direct" if force == "direct" else "staged")
    _check_features(cb, res, x, box, pairs, triplets, CUTS)


def test_cfg4_all_pairs_contact_map(cb):
    top, x, box, pairs, triplets, res, info = _run_cfg(cb, 4, 24, None)
    assert info["mode_name"] == "allpairs" and len(pairs) == 495510
    _check_features(cb, res, x, box, pairs, triplets, CUTS, exact_dist=True)


def test_cfg5_membrane_subset(cb):
    # 1M atoms: the compact form of the referenced atoms (what reaches the device in production), 50k pairs
    top, x, box, pairs, triplets, res, info = _run_cfg(cb, 5, 40, None, compact=True)
    _check_features(cb, res, x, box, pairs, triplets, CUTS)


def test_triclinic_long_range_image_search(cb):
    """Uniformly random atoms in a truncated octahedron: most pairs need the lattice-translation search."""
    from caviar_b200 import synthetic as syn
    top = syn.make_topology(3, n_pairs=64, n_triplets=8)
    rng = np.random.default_rng(5)
    n, T = 512, 16
    x = (rng.random((T, n, 3)) @ top.cell).astype(np.float32)
    x[:, ::7] += (top.cell[0] - 2 * top.cell[2]).astype(np.float32)        # lattice translations change nothing
    pairs = rng.integers(0, n, size=(3000, 2))
    triplets = rng.integers(0, n, size=(700, 3))
    box = np.broadcast_to(top.cell.astype(np.float32), (T, 3, 3))
    cuts = (45.0, 40.0, 60.0)
    with cb.FeaturePlan(n, pairs, triplets, box="triclinic", pair_cutoff=cuts[0], hbond_dist_cutoff=cuts[1],
                        hbond_angle_cutoff=cuts[2]) as plan:
        res = plan.features_host(x, box)
    _check_features(cb, res, x, box, pairs, triplets, cuts)


def test_ortho_known_answer_across_face(cb):
    L = 20.0
    x = np.array([[[0.5, 1, 1], [L - 0.5, 1, 1], [10.5, 1, 1], [0.5, 2, 1]]], dtype=np.float32)
    with cb.FeaturePlan(4, [(0, 1), (0, 2), (1, 0)], [(1, 0, 3), (0, 1, 2)], box="ortho") as plan:
        res = plan.features_host(x, np.array([L, L, L]))
    assert res["dist"][0] == pytest.approx([1.0, 10.0, 1.0], abs=1e-5)
    assert res["angle"][0] == pytest.approx([90.0, 180.0], abs=1e-3)


def test_angle_conditioning_near_0_and_180(cb):
    eps = np.array([0.0, 1e-4, 1e-3, 1e-2])
    n = 3 * len(eps) * 2
    x = np.zeros((1, n, 3), np.float64)
    trip = []
    for k, e in enumerate(eps):
        for s, sign in enumerate((1.0, -1.0)):
            o = 3 * (2 * k + s)
            x[0, o] = (50 + 1.0, 50, 50)
            x[0, o + 1] = (50, 50, 50)
            x[0, o + 2] = (50 + sign * 2.0 * np.cos(e), 50 + 2.0 * np.sin(e), 50)
            trip.append((o, o + 1, o + 2))
    x32 = x.astype(np.float32)
    with cb.FeaturePlan(n, None, trip, box=None) as plan:
        res = plan.features_host(x32)
    want = orc.triplet_angles(x32, trip)
    assert np.abs(res["angle"] - want).max() <= ANGLE_TOL


def test_unsupported_and_invalid_inputs(cb):
    x = np.zeros((2, 8, 3), np.float32)
    with pytest.raises(IndexError):
        cb.FeaturePlan(8, [(0, 8)])
    with pytest.raises(cb.CaviarError):
        cb.FeaturePlan(8, None, None)                  # no features
    upper = np.array([[10, 1, 0], [0, 10, 0], [0, 0, 10]], np.float32)
    with cb.FeaturePlan(8, [(0, 1)], box="triclinic") as plan:
        with pytest.raises(cb.CaviarError):
            plan.features_host(x, upper)               # not lower triangular
    with cb.FeaturePlan(8, [(0, 1)], box="ortho") as plan:
        with pytest.raises(ValueError):
            plan.features_host(x, None)
        res = plan.features_host(x[:0], np.zeros((0, 3)))       # empty trajectory
        assert res["dist"].shape == (0, 1) and res["occupancy"].tolist() == [0]


# ---------------------------------------------------------------------------------------------------------
# 3. device-resident path (what bench.py times) and block accumulation
# ---------------------------------------------------------------------------------------------------------
def test_device_path_matches_host_path_and_accumulates(cb):
    import torch
    from caviar_b200 import synthetic as syn
    top = syn.make_topology(3, n_pairs=3000, n_triplets=600)
    T = 96
    x = syn.parity_frames(top, T, only_referenced=True)
    pairs, triplets = syn.compact_indices(top)
    box = syn.boxes_for(top, T)
    with cb.FeaturePlan(x.shape[1], pairs, triplets, box="triclinic", pair_cutoff=8.0) as plan:
        host = plan.features_host(x, box)
        xd = torch.from_numpy(x).cuda()
        rec = plan.prepare_box(torch.from_numpy(box).cuda())
        staged = plan.stage(xd, rec)
        agg = plan.new_aggregates()
        ws = plan.new_workspace()
        out1 = plan.run(staged[:40], rec[:40], aggregates=agg, workspace=ws)
        out2 = plan.run(staged[40:], rec[40:], aggregates=agg, workspace=ws)
        torch.cuda.synchronize()
        dist = torch.cat([out1.dist, out2.dist]).cpu().numpy()
        ang = torch.cat([out1.angle, out2.angle]).cpu().numpy()
        flags = torch.cat([out1.flags, out2.flags]).cpu().numpy().view(np.uint32)
        score = torch.cat([out1.score, out2.score]).cpu().numpy()
    assert _bits(dist) == _bits(host["dist"]) and _bits(ang) == _bits(host["angle"])
    assert np.array_equal(flags, host["flags"]) and np.array_equal(score, host["score"])
    assert np.array_equal(agg.occupancy.cpu().numpy().astype(np.uint64), host["occupancy"])
    np.testing.assert_allclose(agg.sum.cpu().numpy(), host["sum"], rtol=1e-13)
    np.testing.assert_allclose(agg.sumsq.cpu().numpy(), host["sumsq"], rtol=1e-13)


def test_size_independent_properties_at_scale(cb):
    """cfg3-shaped device run at a larger frame count: properties that need no oracle."""
    import torch
    from caviar_b200 import synthetic as syn
    top = syn.make_topology(3)
    T = 2048
    pairs, triplets = syn.compact_indices(top)
    x = syn.device_frames(top, T)
    box = torch.from_numpy(syn.boxes_for(top, T)).cuda()
    with cb.FeaturePlan(x.shape[1], pairs, triplets, box="triclinic") as plan:
        rec = plan.prepare_box(box)
        staged = plan.stage(x, rec)
        agg = plan.new_aggregates()
        out = plan.run(staged, rec, aggregates=agg)
        # translating every atom by a lattice vector must not change anything beyond float32 rounding
        shift = torch.from_numpy((top.cell[1] - top.cell[2]).astype(np.float32)).cuda()
        x2 = x.clone()
        x2[:, ::3] += shift * box[:, None, 0, 0:1] / float(top.cell[0, 0])
        out_b = plan.run(plan.stage(x2, rec), rec)
        # swapping (i, j) must not change any distance bit
        with cb.FeaturePlan(x.shape[1], pairs[:, ::-1].copy(), triplets[:, ::-1].copy(), box="triclinic") as plan_r:
            rec_r = plan_r.prepare_box(box)
            out_r = plan_r.run(plan_r.stage(x, rec_r), rec_r)
        torch.cuda.synchronize()
    P, A = len(pairs), len(triplets)
    fl = out.flags.cpu().numpy().view(np.uint32)
    fp, ft = cb.unpack_flags(fl, P, A)
    flags_all = np.concatenate([fp, ft], axis=1)
    assert np.array_equal(out.score.cpu().numpy(), flags_all.sum(1))
    assert np.array_equal(agg.occupancy.cpu().numpy(), flags_all.sum(0))
    assert int(out.score.sum()) == int(agg.occupancy.sum())           # checksum of checksums
    assert (out.dist.max() < 100.0) and (out.dist.min() >= 0.0)
    assert torch.equal(out.dist, out_r.dist)
    assert (out.angle - out_r.angle).abs().max() <= 1e-4               # same angle, arms swapped
    assert (out.dist - out_b.dist).abs().max() <= DIST_TOL
    d64 = out.dist.double()
    np.testing.assert_allclose(agg.sum[:P].cpu().numpy(), d64.sum(0).cpu().numpy(), rtol=1e-12)


# ---------------------------------------------------------------------------------------------------------
# 4. cell shapes: the image search must find the TRUE minimum image, not just agree with a 27-image oracle
# ---------------------------------------------------------------------------------------------------------
CELLS = {
    "cubic_as_triclinic": ((40.0, 40.0, 40.0), (90.0, 90.0, 90.0)),
    "hexagonal": ((35.0, 35.0, 50.0), (90.0, 90.0, 120.0)),
    "rhombic_dodecahedron": ((45.0, 45.0, 45.0), (60.0, 60.0, 90.0)),
    "truncated_octahedron": ((60.0, 60.0, 60.0), (109.4712206, 109.4712206, 109.4712206)),
    "monoclinic": ((30.0, 44.0, 38.0), (90.0, 112.0, 90.0)),
    "general": ((33.0, 41.0, 37.0), (75.0, 85.0, 70.0)),
    "rhombohedral_60": ((40.0, 40.0, 40.0), (60.0, 60.0, 60.0)),
}


@pytest.mark.parametrize("name", sorted(CELLS))
def test_true_minimum_image_for_cell_shapes(cb, name):
    lengths, angles = CELLS[name]
    h = orc.box_matrix(lengths, angles)
    h32 = h.astype(np.float32)
    h = h32.astype(np.float64)                           # the cell the device sees
    rng = np.random.default_rng(abs(hash(name)) % 1000)
    n, T = 300, 6
    x = ((rng.random((T, n, 3)) * 3.0 - 1.0) @ h).astype(np.float32)     # atoms spread over several cells
    pairs = rng.integers(0, n, size=(1200, 2))
    triplets = rng.integers(0, n, size=(300, 3))
    box = np.broadcast_to(h32, (T, 3, 3))
    with cb.FeaturePlan(n, pairs, triplets, box="triclinic", pair_cutoff=12.0, hbond_dist_cutoff=15.0,
                        hbond_angle_cutoff=90.0) as plan:
        res = plan.features_host(x, box)
    # brute force over 7x7x7 images in fp64
    rngs = np.arange(-3, 4)
    shifts = np.array([(a, b, c) for a in rngs for b in rngs for c in rngs], dtype=np.float64) @ h
    x64 = x.astype(np.float64)
    delta = x64[:, pairs[:, 0], :] - x64[:, pairs[:, 1], :]
    brute = np.linalg.norm(delta[:, :, None, :] + shifts[None, None], axis=3).min(axis=2)
    assert np.abs(res["dist"] - brute).max() <= DIST_TOL
    # angles: arms by brute-force minimum image, skipping arms that are (nearly) equidistant to two images
    def arm(i, j):
        d = x64[:, triplets[:, i], :] - x64[:, triplets[:, j], :]
        cand = d[:, :, None, :] + shifts[None, None]
        ln = np.linalg.norm(cand, axis=3)
        order = np.argsort(ln, axis=2)[:, :, :2]
        best = np.take_along_axis(cand, order[:, :, :1, None], axis=2)[:, :, 0, :]
        gap = np.take_along_axis(ln, order, axis=2)
        return best, (gap[:, :, 1] - gap[:, :, 0]) > 1e-3, gap[:, :, 0]
    u, ok_u, lu = arm(0, 1)
    v, ok_v, lv = arm(2, 1)
    cosang = np.einsum("tkc,tkc->tk", u, v) / np.maximum(lu * lv, 1e-300)
    want = np.degrees(np.arccos(np.clip(cosang, -1, 1)))
    ok = ok_u & ok_v & (lu > 1e-6) & (lv > 1e-6)
    assert ok.mean() > 0.95
    assert np.abs(res["angle"] - want)[ok].max() <= ANGLE_TOL


def test_aggregates_are_bit_reproducible_run_to_run(cb):
    """The work queue hands items to CTAs in a timing-dependent order, but every item writes its own partial row
    and finalize sums the rows in a fixed order: repeated runs must agree bit for bit (fp64 sums included)."""
    import torch
    from caviar_b200 import synthetic as syn
    top = syn.make_topology(3, n_pairs=4000, n_triplets=900)
    T = 3000
    pairs, triplets = syn.compact_indices(top)
    x = syn.device_frames(top, T)
    box = torch.from_numpy(syn.boxes_for(top, T)).cuda()
    with cb.FeaturePlan(x.shape[1], pairs, triplets, box="triclinic") as plan:
        rec = plan.prepare_box(box)
        staged = plan.stage(x, rec)
        runs = []
        for _ in range(3):
            agg = plan.new_aggregates()
            out = plan.run(staged, rec, aggregates=agg)
            torch.cuda.synchronize()
            runs.append((agg.occupancy.clone(), agg.sum.clone(), agg.sumsq.clone(), out.score.clone(), out.dist.clone()))
    for r in runs[1:]:
        for a, b in zip(runs[0], r):
            assert torch.equal(a, b)


def test_graph_replay_equals_eager_runs_on_a_short_pass(cb):
    """``FeaturePlan.capture``: a short pass (config-1 shape: fewer work items than CTAs, i.e. the tiny-chunk rule of
    make_chunks) recorded as a CUDA graph; two replays must leave exactly what two eager runs leave -- per-frame outputs
    and accumulated aggregates, bit for bit -- and the capture itself must not touch the caller's totals."""
    import torch
    from caviar_b200 import synthetic as syn
    top = syn.make_topology(1)
    T = 1000
    pairs, triplets = syn.compact_indices(top)
    x = syn.device_frames(top, T, only_referenced=True)
    with cb.FeaturePlan(x.shape[1], pairs, triplets) as plan:
        coords = plan.stage(x)
        agg_e = plan.new_aggregates()
        out_e = plan.run(coords, aggregates=agg_e)
        plan.run(coords, out=out_e, aggregates=agg_e)
        torch.cuda.synchronize()
        agg_g = plan.new_aggregates()
        graph, out_g = plan.capture(coords, aggregates=agg_g)
        torch.cuda.synchronize()
        assert int(agg_g.occupancy.sum()) == 0 and float(agg_g.sum.abs().sum()) == 0.0
        graph.replay(); graph.replay()
        torch.cuda.synchronize()
        for a, b in ((out_e.dist, out_g.dist), (out_e.angle, out_g.angle), (out_e.flags, out_g.flags), (out_e.score, out_g.score),
                     (agg_e.occupancy, agg_g.occupancy), (agg_e.sum, agg_g.sum), (agg_e.sumsq, agg_g.sumsq)):
            assert torch.equal(a, b)
        assert int(agg_e.occupancy.sum()) > 0


def test_integration_md_ctypes_stub(cb, gold):
    """The raw-ctypes binding INTEGRATION.md shows a maintainer (no caviar_b200 import on the call path)."""
    import ctypes
    import re
    from caviar_b200.build import LIB_PATH
    text = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "INTEGRATION.md")).read()
    stub = re.search(r"```python\nimport ctypes, numpy as np\n(.*?)```", text, re.S).group(0)
    code = stub.strip("`").replace("python\n", "", 1).replace("/path/to/libcaviar_b200.so", LIB_PATH)
    ns = {}
    exec(code, ns)
    pairs = [(int(a), int(b)) for a, b in gold["C_pairs"]]              # includes negative indices
    got = ns["compute_distances_parallel"](gold["C_x"], pairs)
    assert _bits(got) == _bits(gold["C_d"])
    got = ns["compute_distances_parallel"](gold["A_x"], [(int(a), int(b)) for a, b in gold["A_pairs"]])
    assert _bits(got) == _bits(gold["A_d"])


def test_float64_occupancy_buffer_equals_integer_occupancy(cb):
    """Aggregates accumulated into ONE float64 buffer [occupancy | sum | sumsq] (the multi-GPU exchange buffer): the same
    numbers as the uint64 + float64 triple, occupancy exactly integral, across several accumulating runs."""
    import torch
    from caviar_b200 import synthetic as syn
    top = syn.make_topology(2, n_pairs=1500, n_triplets=300)
    T = 200
    pairs, triplets = syn.compact_indices(top)
    x = torch.from_numpy(syn.parity_frames(top, T, only_referenced=True)).cuda()
    box = torch.from_numpy(syn.boxes_for(top, T)).cuda()
    with cb.FeaturePlan(x.shape[1], pairs, triplets, box="ortho") as plan:
        rec = plan.prepare_box(box)
        coords = plan.stage(x)
        ref = plan.new_aggregates()
        buf, packed = plan.new_packed_aggregates()
        lean_buf, lean = plan.new_packed_aggregates(moments=False)
        for lo, hi in ((0, 80), (80, 200)):
            plan.run(coords[lo:hi], rec[lo:hi], aggregates=ref)
            plan.run(coords[lo:hi], rec[lo:hi], aggregates=packed)
            plan.run(coords[lo:hi], rec[lo:hi], out=plan.alloc_outputs(hi - lo, want=("flags",)), aggregates=lean)
        torch.cuda.synchronize()
        F = plan.F
        assert buf.shape == (3 * F,) and lean_buf.shape == (F,) and buf.dtype == torch.float64
        assert torch.equal(buf[:F], ref.occupancy.double()) and torch.equal(lean_buf, ref.occupancy.double())
        assert torch.equal(buf[F:2 * F], ref.sum) and torch.equal(buf[2 * F:], ref.sumsq)
        with pytest.raises(ValueError):
            plan.run(coords, rec, aggregates=cb.Aggregates(torch.zeros(F, dtype=torch.float32, device="cuda"), None, None))
