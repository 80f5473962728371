"""Randomised shape sweep of the sm_100a path against the oracle: ragged sizes, heavy atom sharing between features
(the staged block layout has to keep shared atoms consistent across tiles), all box kinds, every plan mode; and the
plan cache behind the reference operator (CAVIAR-1.0.py:337-338 call it with one list for two trajectories)."""
import numpy as np
import pytest

from oracle import caviar_oracle as orc
from test_gpu_parity import _bits, _check_features

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    import caviar_b200
    return caviar_b200


def _cell(kind, L):
    if kind == "ortho":
        return np.diag([L, 1.1 * L, 0.9 * L])
    # truncated octahedron in Amber's lower-triangular setting
    return L * np.array([[1.0, 0.0, 0.0], [-1.0 / 3.0, 2.0 * np.sqrt(2.0) / 3.0, 0.0],
                         [-1.0 / 3.0, -np.sqrt(2.0) / 3.0, np.sqrt(6.0) / 3.0]])


@pytest.mark.parametrize("seed", range(24))
def test_random_shapes_match_oracle(cb, seed):
    rng = np.random.default_rng(9000 + seed)
    box_kind = (None, "ortho", "triclinic")[seed % 3]
    force = (None, "staged", "direct", None)[(seed // 3) % 4]
    n = int(rng.choice([5, 31, 64, 257, 1000, 4099]))
    T = int(rng.choice([1, 2, 3, 17, 70]))
    P = int(rng.choice([0, 1, 31, 33, 895, 897, 2500]))
    A = int(rng.choice([0, 1, 32, 225, 700]))
    if P + A == 0:
        P = 7
    hot = max(2, n // int(rng.choice([1, 4, 50])))          # draw from a subset: many features share atoms
    L = 30.0
    cell = _cell(box_kind, L) if box_kind else None
    base = rng.random((n, 3)) * L if cell is None else rng.random((n, 3)) @ cell
    x = (base[None] + rng.normal(0, 0.4, size=(T, n, 3))).astype(np.float32)
    pairs = rng.integers(0, hot, size=(P, 2))
    triplets = rng.integers(0, hot, size=(A, 3))
    if A:                                                    # some hydrogen-bond sized triplets
        k = min(A, n // 3)
        for t in range(k):
            a, b, c = triplets[t]
            x[:, b] = x[:, a] + rng.normal(0, 0.6, size=(T, 3)).astype(np.float32)
            x[:, c] = x[:, b] + rng.normal(0, 1.2, size=(T, 3)).astype(np.float32)
    box = None
    if box_kind == "ortho":
        box = np.broadcast_to(np.diag(cell).astype(np.float32), (T, 3)) * (1 + 1e-3 * np.arange(T, dtype=np.float32))[:, None]
    elif box_kind == "triclinic":
        box = np.broadcast_to(cell.astype(np.float32), (T, 3, 3)) * (1 + 1e-3 * np.arange(T, dtype=np.float32))[:, None, None]
    cuts = (9.0, 6.0, 80.0)
    with cb.FeaturePlan(n, pairs if P else None, triplets if A else None, box=box_kind, pair_cutoff=cuts[0],
                        hbond_dist_cutoff=cuts[1], hbond_angle_cutoff=cuts[2], force=force) as plan:
        res = plan.features_host(x, box, frames_per_block=int(rng.choice([0, 1, 5])))
        info = plan.info
        lean = plan.features_host(x, box, want=("flags", "score", "occ"))          # contact-map flavour: same bits
        for key in ("flags", "score", "occupancy"):
            assert np.array_equal(lean[key], res[key]), f"lean run differs in {key}"
    if info["mode_name"] == "staged":
        assert info["gather_bank_conflicts"] >= 0
    if not P:
        res["dist"] = np.zeros((T, 0), np.float32)
    if not A:
        res["angle"] = np.zeros((T, 0), np.float32)
    _check_features(cb, res, x, box, [tuple(p) for p in pairs], [tuple(t) for t in triplets], cuts, exact_dist=box_kind is None)


def test_operator_plan_cache_switches_lists(cb):
    """Same list twice (cache hit), another list (replaced), the first again; then an explicit release."""
    from caviar_b200 import _lib
    rng = np.random.default_rng(77)
    xa = rng.normal(0, 20, size=(37, 50, 3)).astype(np.float32)
    xb = rng.normal(0, 20, size=(11, 50, 3)).astype(np.float32)
    xc = rng.normal(0, 20, size=(5, 64, 3)).astype(np.float32)
    p1 = orc.all_pairs(50, 5)
    p2 = [(int(a), int(b)) for a, b in rng.integers(0, 50, size=(len(p1), 2))]      # same length, other content
    p3 = orc.all_pairs(64, 3)
    for x, p in ((xa, p1), (xb, p1), (xa, p2), (xc, p3), (xb, p1)):
        assert _bits(cb.compute_distances_parallel(x, p)) == _bits(orc.pair_distances(x, p, n_jobs=1))
    _lib.lib.caviar_release_cache()
    _lib.lib.caviar_release_cache()                           # idempotent
    assert _bits(cb.compute_distances_parallel(xa, p1)) == _bits(orc.pair_distances(xa, p1, n_jobs=1))


@pytest.mark.parametrize("cfg_id,T", [(2, 2000), (3, 1000), (5, 200)])
def test_survey_parity_sets_against_c_oracle(cb, cfg_id, T):
    """SURVEY.md 8d parity sets at (close to) their full frame counts: affordable with the multi-threaded C oracle."""
    from caviar_b200 import synthetic as syn
    from oracle import c_oracle as co
    top = syn.make_topology(cfg_id)
    x = syn.parity_frames(top, T, only_referenced=True)
    pairs, triplets = syn.compact_indices(top)
    box = syn.boxes_for(top, T)
    cuts = (8.0, 3.0, 135.0)
    with cb.FeaturePlan(x.shape[1], pairs, triplets if len(triplets) else None, box=top.cfg.box, pair_cutoff=cuts[0],
                        hbond_dist_cutoff=cuts[1], hbond_angle_cutoff=cuts[2]) as plan:
        res = plan.features_host(x, box)
    P, A = len(pairs), len(triplets)
    d64 = co.pbc_pair_distances(x, pairs, box)
    assert np.abs(res["dist"].astype(np.float64) - d64).max() <= 1e-4
    fp, ft = cb.unpack_flags(res["flags"], P, A)
    assert np.array_equal(fp, orc.contact_flags(res["dist"], cuts[0]))               # exact on the device's own features
    near = orc.near_cutoff(d64, cuts[0], 1e-4)
    assert np.array_equal(fp[~near], orc.contact_flags(d64, cuts[0])[~near])
    assert near.mean() < 1e-3
    flags_all = fp
    if A:
        a64, e64 = co.triplet_angles(x, triplets, box, with_end_distances=True)
        assert np.abs(res["angle"].astype(np.float64) - a64).max() <= 1e-3
        near_t = orc.near_cutoff(e64, cuts[1], 1e-4) | orc.near_cutoff(a64, cuts[2], 1e-3)
        assert np.array_equal(ft[~near_t], orc.hbond_flags(e64, a64, cuts[1], cuts[2])[~near_t])
        flags_all = np.concatenate([fp, ft], axis=1)
    assert np.array_equal(res["score"], orc.frame_scores(flags_all))
    assert np.array_equal(res["occupancy"], orc.occupancy(flags_all))
    np.testing.assert_allclose(res["sum"][:P], res["dist"].astype(np.float64).sum(axis=0), rtol=1e-12)


def test_plan_order_input_equals_caller_order_input(cb):
    """The hand-over format of staged plans: frames sliced with plan.atom_order (raw floats, duplicates included) give
    bit-identical results to the caller-order frames permuted on the device by K0, with and without a box."""
    import torch
    rng = np.random.default_rng(1)
    n, T = 3000, 9
    pairs = rng.integers(0, n, size=(2500, 2))
    trips = rng.integers(0, n, size=(700, 3))
    x = (rng.random((T, n, 3)) * 20).astype(np.float32)
    x[:, ::7] += np.float32(60.0)                               # some atoms three cells away
    for box in (None, "ortho", "triclinic"):
        cell = {None: None, "ortho": np.full((T, 3), 20.0), "triclinic": np.broadcast_to(orc.box_matrix([20.0] * 3, [109.4712206] * 3), (T, 3, 3)).copy()}[box]
        with cb.FeaturePlan(n, pairs, trips, box=box) as plan:
            assert plan.info["mode_name"] == "staged" and plan.info["staged_fixedpoint_floats"] == 0
            order = plan.atom_order
            xp = np.ascontiguousarray(x[:, order, :])
            a = plan.features_host(x, cell)
            la = plan.last_launches
            b = plan.features_host(xp, cell, plan_order=True)
            assert plan.last_launches < la                      # no permutation kernel on the plan-order path
            for k in ("dist", "angle", "flags", "score", "occupancy", "sum", "sumsq"):
                assert _bits(a[k]) == _bits(b[k]), (box, k)
            xd = torch.from_numpy(x).cuda()
            staged = plan.stage(xd)                             # K0 = the same permutation, bit for bit, box or not
            assert torch.equal(staged.view(T, -1, 3), torch.from_numpy(xp).cuda())
            rec = plan.prepare_box(torch.from_numpy(cell.astype(np.float32)).cuda()) if box else None
            out = plan.run(torch.from_numpy(xp).cuda(), rec)    # (T, S/3, 3) frames in plan order, straight in
            assert _bits(out.dist.cpu().numpy()) == _bits(a["dist"])
            with pytest.raises(ValueError):
                plan.features_host(x, cell, plan_order=True)    # caller-order shape is refused, not guessed


def test_unwrapped_coordinates_far_outside_the_cell(cb):
    """Molecules that diffused many box lengths away (unwrapped trajectories): displacements that span several cells are
    re-taken and reduced in fp64 inside the fused kernel (far_displacement)."""
    rng = np.random.default_rng(2)
    n, T, L = 96, 7, 25.0
    base = rng.random((n, 3)) * L
    x = base[None] + rng.normal(0, 0.3, size=(T, n, 3))
    x += rng.integers(-40, 40, size=(1, n, 3)) * L             # every atom in its own distant image
    x = x.astype(np.float32)
    pairs = rng.integers(0, n, size=(300, 2))
    trips = rng.integers(0, n, size=(100, 3))
    box = np.full((T, 3), L)
    with cb.FeaturePlan(n, pairs, trips, box="ortho", pair_cutoff=6.0, hbond_dist_cutoff=7.0, hbond_angle_cutoff=70.0) as plan:
        res = plan.features_host(x, box)
    a64 = orc.triplet_angles(x, trips, box)
    d64 = orc.pbc_pair_distances(x, pairs, box)
    # float32 INPUT at |x| ~ 1000 A carries 6e-5 A of quantisation; the path itself adds nothing visible on top
    assert np.abs(res["dist"] - d64).max() <= 1e-4
    assert np.abs(res["angle"] - a64).max() <= 1e-3


@pytest.mark.parametrize("n,box,A,mode,stages", [(1000, None, 0, "zerocopy", 2), (3000, None, 0, "direct", 0),
                                                   (5000, "ortho", 0, "direct", 0), (1000, "triclinic", 64, "zerocopy", 2),
                                                   (5000, "triclinic", 200, "direct", 0), (9000, None, 0, "direct", 0)])
def test_large_overlapping_selection_modes(cb, n, box, A, mode, stages):
    """Rows of a big all-pairs list: per-tile blocks would multiply the frame hundreds of times.  Small frames are
    streamed whole (with or without a box); large ones are gathered from the caller's frames in global memory."""
    from oracle import c_oracle as co
    rng = np.random.default_rng(n + A)
    rows = np.arange(7, n, 211 if n > 1000 else 37)
    pairs = np.stack([np.repeat(rows, n), np.tile(np.arange(n), len(rows))], 1).astype(np.int32)
    trips = rng.integers(0, n, size=(A, 3)) if A else None
    T, L = 6, 80.0
    cell = None
    if box == "ortho":
        cell = np.full((T, 3), L)
    elif box == "triclinic":
        cell = np.broadcast_to(orc.box_matrix([L, L, L], [109.4712206] * 3), (T, 3, 3)).copy()
    x = (rng.random((T, n, 3)) * L).astype(np.float32)
    x[:, ::5] += np.float32(3 * L)                              # some atoms a few cells away
    cuts = (10.0, 30.0, 60.0)
    with cb.FeaturePlan(n, pairs, trips, box=box, pair_cutoff=cuts[0], hbond_dist_cutoff=cuts[1], hbond_angle_cutoff=cuts[2]) as plan:
        assert plan.info["mode_name"] == mode and plan.info["stages"] == stages
        res = plan.features_host(x, cell, want=("dist", "angle", "flags", "score", "agg"))
    if box is None:
        assert _bits(res["dist"]) == _bits(orc.pair_distances_f32_steps(x, pairs[:, 0], pairs[:, 1]))
    else:
        assert np.abs(res["dist"] - co.pbc_pair_distances(x, pairs, cell)).max() <= 1e-4
    fp, ft = cb.unpack_flags(res["flags"], len(pairs), A)
    assert np.array_equal(fp, orc.contact_flags(res["dist"], cuts[0]))
    flags = fp
    if A:
        a64, e64 = co.triplet_angles(x, trips, cell, with_end_distances=True)
        assert np.abs(res["angle"] - a64).max() <= 1e-3
        near = orc.near_cutoff(e64, cuts[1], 1e-4) | orc.near_cutoff(a64, cuts[2], 1e-3)
        assert np.array_equal(ft[~near], orc.hbond_flags(e64, a64, cuts[1], cuts[2])[~near])
        flags = np.concatenate([fp, ft], axis=1)
    assert np.array_equal(res["occupancy"], flags.sum(axis=0).astype(np.uint64))
    assert np.array_equal(res["score"], flags.sum(axis=1).astype(np.int32))


def test_non_finite_coordinates_propagate(cb):
    """NaN in, NaN out on every path (plan-order blocks are raw floats: nothing can hide a blown-up frame)."""
    n, T = 32, 4
    rng = np.random.default_rng(3)
    x = (rng.random((T, n, 3)) * 20).astype(np.float32)
    x[2, 5, 1] = np.nan
    box = np.full((T, 3), 20.0)
    for force in (None, "staged", "direct"):
        with cb.FeaturePlan(n, [(5, 6), (1, 2)], None, box="ortho", force=force) as plan:
            d = plan.features_host(x, box, want=("dist",))["dist"]
            assert np.isnan(d[2, 0]) and np.isfinite(np.delete(d.ravel(), 4)).all()
        with cb.FeaturePlan(n, [(1, 2)], [(4, 5, 6), (7, 8, 9)], box="ortho", force=force) as plan:
            res = plan.features_host(x, box)
            assert np.isnan(res["angle"][2, 0]) and np.isfinite(np.delete(res["angle"].ravel(), 4)).all()
            assert np.isfinite(res["dist"]).all()
    with cb.FeaturePlan(n, None, [(4, 5, 6)], box=None) as plan:
        assert np.isnan(plan.features_host(x, want=("angle",))["angle"][2, 0])
    cb.release_cache()


def test_outputs_beyond_2_31_elements(cb):
    """Maximum sizes: a (T, P) distance matrix of 2.5e9 elements (9.9 GB) -- row offsets need 64 bits everywhere."""
    import torch
    if torch.cuda.mem_get_info()[0] < 40 << 30:
        pytest.skip("needs 40 GB of free device memory")
    n, T = 1000, 5003
    pairs = orc.all_pairs(n, 5)
    P = len(pairs)
    assert T * P > 2 ** 31
    g = torch.Generator(device="cuda").manual_seed(4)
    x = (torch.rand((T, n, 3), device="cuda", generator=g) * 60.0).contiguous()
    with cb.FeaturePlan(n, pairs, None, box=None, pair_cutoff=8.0) as plan:
        out = plan.alloc_outputs(T, want=("dist", "flags", "score"))
        agg = plan.new_aggregates()
        plan.run(x, None, out=out, aggregates=agg)
        torch.cuda.synchronize()
    ii, jj = np.asarray(pairs)[:, 0], np.asarray(pairs)[:, 1]
    for t in (0, 2171, 4334, T - 1):                      # rows on both sides of element 2^31 (row 4334) and the last one
        want = orc.pair_distances_f32_steps(x[t:t + 1].cpu().numpy(), ii, jj)
        assert _bits(out.dist[t].cpu().numpy()) == _bits(want[0])
        fp, _ = cb.unpack_flags(out.flags[t:t + 1].cpu().numpy(), P, 0)
        assert np.array_equal(fp[0], orc.contact_flags(want, 8.0)[0])
        assert int(out.score[t]) == int(fp.sum())
    thr = float(orc.f32_at_or_above(8.0))
    occ = torch.zeros(P, dtype=torch.int64, device="cuda")
    for t0 in range(0, T, 512):                           # the device's own distances, in slices
        occ += (out.dist[t0:t0 + 512] < thr).sum(dim=0)
    assert torch.equal(occ, agg.occupancy)
    assert int(out.score.sum()) == int(occ.sum())


@pytest.mark.parametrize("box", [None, "triclinic"])
def test_weighted_frame_scores(cb, box):
    """per-frame score weighted by CAVIAR's per-pair loadings -- exact fixed-point sums, bit for bit."""
    import torch
    from caviar_b200 import synthetic as syn
    top = syn.make_topology(1 if box is None else 3, n_pairs=1500, n_triplets=333)
    T = 41
    x = syn.parity_frames(top, T, only_referenced=True)
    pairs, trips = syn.compact_indices(top)
    cell = syn.boxes_for(top, T)
    rng = np.random.default_rng(6)
    w = np.round(rng.random(len(pairs) + len(trips)) * 2.5, 6)              # abs_loading is printed with 6 decimals
    with cb.FeaturePlan(x.shape[1], pairs, trips, box=top.cfg.box) as plan:
        res = plan.features_host(x, cell, weights=w)
        fp, ft = cb.unpack_flags(res["flags"], len(pairs), len(trips))
        want = orc.weighted_frame_scores(np.concatenate([fp, ft], axis=1), w)
        assert np.array_equal(res["score_weighted"], want)
        res2 = plan.features_host(x, cell, want=("score",), weights=w, frames_per_block=7)     # flags not requested
        assert np.array_equal(res2["score_weighted"], want) and "flags" not in res2
        assert np.array_equal(plan.features_host(x, cell, want=("score",), weights=np.ones_like(w))["score_weighted"],
                              res["score"].astype(np.float64))
        xd = torch.from_numpy(x).cuda()
        rec = plan.prepare_box(torch.from_numpy(cell).cuda()) if cell is not None else None
        out = plan.run(plan.stage(xd, rec), rec)
        assert np.array_equal(plan.weighted_scores(out.flags, w).cpu().numpy(), want)
        with pytest.raises(ValueError):
            plan.weighted_scores(out.flags, w[:-1])


@pytest.mark.parametrize("cfg_id,force", [(1, None), (1, "direct"), (2, None), (3, None), (3, "direct"), (4, None)])
def test_lean_contact_map_runs_match_the_full_kernel(cb, cfg_id, force):
    """Flags / scores / occupancy without a feature matrix and without moments take the lean kernel flavour (no box:
    cut-off on the squared distance, no square root): every bit must equal the full run's."""
    import torch
    from caviar_b200 import synthetic as syn
    top = syn.make_topology(cfg_id) if cfg_id != 3 else syn.make_topology(3, n_pairs=3000, n_triplets=700)
    T = 24 if cfg_id == 4 else 150
    x = syn.parity_frames(top, T, only_referenced=True)
    pairs, trips = syn.compact_indices(top)
    cell = syn.boxes_for(top, T)
    with cb.FeaturePlan(x.shape[1], pairs, trips if len(trips) else None, box=top.cfg.box, force=force) as plan:
        full = plan.features_host(x, cell)
        lean = plan.features_host(x, cell, want=("flags", "score", "occ"), frames_per_block=37)
        assert set(lean) == {"flags", "score", "occupancy"}
        for k in ("flags", "score", "occupancy"):
            assert np.array_equal(full[k], lean[k]), k
        only_occ = plan.features_host(x, cell, want=("occ",))
        assert np.array_equal(only_occ["occupancy"], full["occupancy"])
        xd = torch.from_numpy(x).cuda()
        rec = plan.prepare_box(torch.from_numpy(cell).cuda()) if cell is not None else None
        agg = plan.new_aggregates(moments=False)
        out = plan.run(plan.stage(xd, rec), rec, out=plan.alloc_outputs(T, want=("flags", "score")), aggregates=agg)
        assert np.array_equal(out.flags.cpu().numpy().view(np.uint32), full["flags"])
        assert np.array_equal(agg.occupancy.cpu().numpy().astype(np.uint64), full["occupancy"])
    # thresholds that sit exactly on representable distances: the squared-distance compare must flip on the same bit
    xk = np.zeros((1, 4, 3), np.float32); xk[0, 1] = (3, 4, 0); xk[0, 2] = (3, 4, 12); xk[0, 3] = (0.1, 0.2, 0.2)
    for cut in (5.0, np.nextafter(np.float32(5.0), np.float32(6.0)), 13.0, 0.3, np.float32(0.3), 0.0, -1.0):
        with cb.FeaturePlan(4, [(0, 1), (0, 2), (0, 3), (1, 1)], None, pair_cutoff=float(cut)) as plan:
            a = plan.features_host(xk)
            b = plan.features_host(xk, want=("flags", "score", "occ"))
            assert np.array_equal(a["flags"], b["flags"]) and np.array_equal(a["occupancy"], b["occupancy"]), cut


def test_downstream_pooled_mean_centring_and_covariance(cb):
    """f3: the steps that follow the operator in CAVIAR-1.0.py (:342-343, :378-380, :396-401), on the device."""
    import torch
    from caviar_b200 import downstream as ds
    rng = np.random.default_rng(12)
    n, TA, TB = 40, 300, 260
    pairs = orc.all_pairs(n, 5)
    xa = (np.cumsum(rng.normal(0, 3.8, size=(n, 3)), 0)[None] + rng.normal(0, 0.6, size=(TA, n, 3))).astype(np.float32)
    xb = (np.cumsum(rng.normal(0, 3.8, size=(n, 3)), 0)[None] + rng.normal(0, 0.6, size=(TB, n, 3))).astype(np.float32)
    with cb.FeaturePlan(n, pairs, None) as plan:
        da, db = torch.from_numpy(xa).cuda(), torch.from_numpy(xb).cuda()
        agg_a, agg_b = plan.new_aggregates(), plan.new_aggregates()
        out_a = plan.run(plan.stage(da), None, aggregates=agg_a)
        out_b = plan.run(plan.stage(db), None, aggregates=agg_b)
        mu = ds.pooled_mean(agg_a.sum, TA, agg_b.sum, TB)
        XA, XB = ds.centred(out_a.dist, mu), ds.centred(out_b.dist, mu)
        lag = 7
        C00, Ctt, C0t = ds.cov_blocks(XA[:-lag], XA[lag:])
        torch.cuda.synchronize()
    # the reference's own lines, in numpy
    distA, distB = orc.pair_distances(xa, pairs, n_jobs=1), orc.pair_distances(xb, pairs, n_jobs=1)
    mu_ref = (distA.mean(axis=0) + distB.mean(axis=0)) / 2.0
    np.testing.assert_allclose(mu.cpu().numpy(), mu_ref, rtol=1e-6)
    XA_ref = distA - mu_ref
    np.testing.assert_allclose(XA.cpu().numpy(), XA_ref, atol=5e-6)
    np.testing.assert_allclose(XB.cpu().numpy(), distB - mu_ref, atol=5e-6)
    X0, Xt = XA_ref[:-lag].astype(np.float64), XA_ref[lag:].astype(np.float64)
    for got, want in ((C00, X0.T @ X0), (Ctt, Xt.T @ Xt), (C0t, X0.T @ Xt)):
        np.testing.assert_allclose(got.cpu().numpy(), want / len(X0), rtol=2e-4, atol=2e-5)
    with pytest.raises(TypeError):
        ds.cov_blocks(XA_ref, XA_ref)                      # host arrays: refused, no CPU path


def test_every_combination_of_requested_outputs(cb):
    """Any subset of {dist, flags, score} x {no aggregates, occupancy, occupancy + moments}: same values whatever else
    is requested (a distance matrix with occupancy but WITHOUT moments once wrote through a null pointer)."""
    import itertools
    import torch
    rng = np.random.default_rng(12)
    n, T = 400, 40
    x = torch.from_numpy((rng.random((T, n, 3)) * 30).astype(np.float32)).cuda()
    for pairs, box in ((cb.all_pairs(n, 5), None), (rng.integers(0, n, size=(3000, 2)), "ortho")):
        with cb.FeaturePlan(n, pairs, None, box=box, pair_cutoff=9.0) as plan:
            rec = plan.prepare_box(torch.full((T, 3), 30.0, device="cuda")) if box else None
            coords = plan.stage(x)
            ref_out = plan.alloc_outputs(T)
            ref_agg = plan.new_aggregates()
            plan.run(coords, rec, out=ref_out, aggregates=ref_agg)
            for r in range(0, 4):
                for want in itertools.combinations(("dist", "flags", "score"), r):
                    for agg_kind in (None, "occ", "all"):
                        out = plan.alloc_outputs(T, want=want)
                        agg = None if agg_kind is None else plan.new_aggregates(moments=agg_kind == "all")
                        if not want and agg is None:
                            continue
                        plan.run(coords, rec, out=out, aggregates=agg)
                        torch.cuda.synchronize()
                        for name in want:
                            assert torch.equal(getattr(out, name), getattr(ref_out, name)), (box, want, agg_kind, name)
                        if agg is not None:
                            assert torch.equal(agg.occupancy, ref_agg.occupancy), (box, want, agg_kind)
                            if agg_kind == "all":
                                assert torch.equal(agg.sum, ref_agg.sum) and torch.equal(agg.sumsq, ref_agg.sumsq)
