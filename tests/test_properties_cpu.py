"""Property tests (hypothesis) of the oracle's extension semantics and of the host-side planner."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import caviar_oracle as orc


@st.composite
def systems(draw):
    n = draw(st.integers(4, 24))
    t = draw(st.integers(1, 4))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    lengths = rng.uniform(12.0, 30.0, size=3)
    angles = rng.uniform(75.0, 105.0, size=3)        # always a valid (non-degenerate) cell
    h = orc.box_matrix(lengths, angles)
    x = rng.random((t, n, 3)) @ h
    pairs = rng.integers(0, n, size=(draw(st.integers(1, 12)), 2))
    trips = rng.integers(0, n, size=(draw(st.integers(1, 6)), 3))
    return x, h, pairs, trips, rng


@settings(max_examples=40, deadline=None)
@given(systems())
def test_oracle_pbc_invariances(sysm):
    x, h1, pairs, trips, rng = sysm
    h = np.broadcast_to(h1, (x.shape[0], 3, 3))          # explicit per-frame cells ((3,3) is ambiguous when T == 3)
    d = orc.pbc_pair_distances(x, pairs, h)
    # swapping (i, j) leaves the distance unchanged
    assert np.allclose(orc.pbc_pair_distances(x, pairs[:, ::-1], h), d, atol=1e-9)
    # translating any atom by a lattice vector leaves every minimum-image quantity unchanged
    x2 = x.copy()
    shift = rng.integers(-2, 3, size=(x.shape[1], 3)).astype(np.float64) @ h1
    x2 += shift[None]
    assert np.allclose(orc.pbc_pair_distances(x2, pairs, h), d, atol=1e-8)
    # angles: skip arms that are (nearly) equidistant to two images -- there the angle is not well defined
    ang1, ang2 = orc.triplet_angles(x, trips, h), orc.triplet_angles(x2, trips, h)
    arm_u = orc.pbc_pair_distances(x, trips[:, [0, 1]], h)
    arm_v = orc.pbc_pair_distances(x, trips[:, [2, 1]], h)
    coef = np.array([(a, b, c) for a in range(-2, 3) for b in range(-2, 3) for c in range(-2, 3) if (a, b, c) != (0, 0, 0)])
    half = 0.5 * np.linalg.norm(coef @ h1, axis=1).min()        # below this an arm has a unique minimum image
    safe = (arm_u < 0.9 * half) & (arm_v < 0.9 * half) & (arm_u > 1e-3) & (arm_v > 1e-3)
    assert np.allclose(ang1[safe], ang2[safe], atol=1e-5)
    # never longer than the plain Euclidean distance; never longer than half the cell diagonal sum
    assert (d <= orc.pbc_pair_distances(x, pairs, None) + 1e-9).all()


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 60), st.integers(1, 70), st.integers(0, 2 ** 31 - 1), st.floats(2.0, 14.0))
def test_flag_aggregates_are_consistent(t, p, seed, cut):
    rng = np.random.default_rng(seed)
    d = rng.uniform(1.0, 15.0, size=(t, p)).astype(np.float32)
    f = orc.contact_flags(d, cut)
    assert np.array_equal(f, d < orc.f32_at_or_above(cut))               # float32 threshold == double compare
    words = orc.pack_flags(f)
    assert np.array_equal(orc.unpack_flags(words, p), f)
    assert int(orc.occupancy(f).sum()) == int(orc.frame_scores(f).sum())  # checksum of checksums


@settings(max_examples=30, deadline=None)
@given(st.integers(2, 3000), st.integers(0, 4000), st.integers(0, 900), st.integers(0, 2 ** 31 - 1),
       st.sampled_from([None, "ortho", "triclinic"]), st.sampled_from([None, "direct", "staged"]))
def test_planner_invariants(n_atoms, n_pairs, n_trip, seed, box, force):
    import __graft_entry__ as g
    g.build()
    import caviar_b200 as cb
    if n_pairs + n_trip == 0:
        n_pairs = 1
    rng = np.random.default_rng(seed)
    pairs = rng.integers(0, n_atoms, size=(n_pairs, 2))
    trips = rng.integers(0, n_atoms, size=(n_trip, 3))
    info = cb.FeaturePlan(n_atoms, pairs, trips, box=box, force=force, describe_only_sms=148).info
    used = np.unique(np.concatenate([pairs.ravel(), trips.ravel()]))
    assert info["n_features"] == n_pairs + n_trip
    assert info["n_distinct_atoms"] == len(used)
    assert info["n_flag_words"] == (n_pairs + 31) // 32 + (n_trip + 31) // 32
    assert info["n_tiles"] == -(-n_pairs // 896) + -(-n_trip // 896) or info["n_tiles"] >= 1
    assert 1 <= info["n_ctas"] <= 296
    assert info["workspace_bytes"] >= 256 + 24 * (n_pairs + n_trip)
    if info["mode_name"] == "direct":
        assert force == "direct" and info["smem_bytes"] == 0
    else:
        assert 2 <= info["stages"] <= 8 and info["frames_per_stage"] >= 1
        assert info["smem_bytes"] <= 113 * 1024 or info["n_ctas"] <= 148
    if info["mode_name"] == "staged":
        s = info["staged_floats_per_frame"]
        assert s % 12 == 0 and s >= 3 * len(used) and s <= 3 * (2 * n_pairs + 3 * n_trip) + 12 * info["n_tiles"]
    if info["mode_name"] == "zerocopy":
        assert force is None and n_atoms % 4 == 0 and n_atoms * 12 <= 48 * 1024


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.sampled_from([None, "ortho", "triclinic"]), st.sampled_from([None, "staged", "direct"]))
def test_planner_layout_property(seed, box, force):
    """Whatever the list looks like (sizes, sharing, degenerate features), every feature slot of the plan resolves to
    the caller's atoms -- checked through caviar_plan_layout, no GPU involved."""
    import caviar_b200 as cb
    from test_abi_cpu import _check_layout
    rng = np.random.default_rng(seed)
    n = int(rng.choice([1, 2, 5, 33, 64, 300, 2048, 5000]))
    hot = max(1, n // int(rng.choice([1, 3, 40])))
    P = int(rng.choice([0, 1, 32, 897, 1793, 4000]))
    A = int(rng.choice([0, 1, 225, 673, 1500]))
    if P + A == 0:
        P = 3
    pairs = rng.integers(0, hot, size=(P, 2)) if P else None
    trips = rng.integers(0, hot, size=(A, 3)) if A else None
    _check_layout(cb, n, pairs, trips, box, force=force)


@settings(max_examples=40, deadline=None)
@given(systems(), st.sampled_from(["none", "ortho", "triclinic"]))
def test_c_oracle_agrees_with_numpy_oracle_property(sysm, kind):
    """The two independent restatements (numpy, C) agree on random cells and lists: distances to 1e-9 A, angles to
    1e-5 deg (acos noise at 0 / 180), the fused C pass with the flag / score / occupancy definitions."""
    from oracle import c_oracle as co
    x, h1, pairs, trips, rng = sysm
    x = x.astype(np.float32)
    T = x.shape[0]
    if kind == "none":
        box = None
    elif kind == "ortho":
        box = np.broadcast_to(np.diag(h1), (T, 3)).copy()
    else:
        box = np.broadcast_to(h1, (T, 3, 3)).copy()
    d_np, d_c = orc.pbc_pair_distances(x, pairs, box), co.pbc_pair_distances(x, pairs, box, n_threads=2)
    np.testing.assert_allclose(d_c, d_np, rtol=0, atol=1e-9)
    gap = orc.triplet_arm_gaps(x, trips, box)
    a_np = orc.triplet_angles(x, trips, box)
    a_c, e_c = co.triplet_angles(x, trips, box, n_threads=2, with_end_distances=True)
    ok = gap > 1e-6                                      # an arm on a Voronoi face may legitimately pick either image
    np.testing.assert_allclose(a_c[ok], a_np[ok], rtol=0, atol=1e-5)
    np.testing.assert_allclose(e_c, orc.triplet_end_distances(x, trips, box), rtol=0, atol=1e-9)
    cuts = (9.0, 8.0, 70.0)
    f = co.fused_pass(x, pairs, trips, box, *cuts, n_threads=3)
    flags = np.concatenate([orc.contact_flags(f["dist"], cuts[0]),
                            orc.hbond_flags(e_c.astype(np.float32), f["angle"], cuts[1], cuts[2])], axis=1)
    assert np.array_equal(f["flags"].astype(bool), flags)
    assert np.array_equal(f["score"], orc.frame_scores(flags)) and np.array_equal(f["occupancy"], orc.occupancy(flags))


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 3), st.integers(10, 160), st.sampled_from([100.0, 1000.0, 10000.0]),
       st.sampled_from([0.02, 0.3, 3.0, 300.0]), st.sampled_from([1995, 2023]))
def test_xtc_round_trip_is_exact_for_any_mixture(tmp_path_factory, seed, T, N, precision, spread, magic):
    """The compressed XTC stream (large numbers, runs of small offsets with swapped first pairs, radix changes, ranges too
    wide for one mixed-radix number) always decodes to exactly the rounded positions the writer stored."""
    import __graft_entry__ as g
    g.build()
    from caviar_b200 import trajio
    rng = np.random.default_rng(seed)
    # clusters of 1..4 atoms around random centres: some atoms close to their predecessor, some far
    centres = rng.random((T, N, 3)) * spread * 20.0
    keep = rng.random(N) < 0.5
    xyz = centres.copy()
    for a in range(1, N):
        if keep[a]:
            xyz[:, a] = xyz[:, a - 1] + rng.normal(0, spread * 0.02, size=(T, 3))
    path = str(tmp_path_factory.mktemp("xtc") / "p.xtc")
    want = trajio.write_xtc(path, xyz.astype(np.float32), np.diag([1.0, 2.0, 3.0]), precision=precision, magic=magic)
    tr = trajio.Trajectory(path)
    assert (tr.n_frames, tr.n_atoms) == (T, N)
    assert np.array_equal(tr.coords(0, T), want * np.float32(10.0))
    tr.close()
