"""Pin the CPU oracle against vectors produced by the reference itself (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from oracle import caviar_oracle as orc


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "distances.npz"))


@pytest.fixture(scope="module")
def meta(golden_dir):
    with open(os.path.join(golden_dir, "reference_cases.json")) as fh:
        return json.load(fh)


def _pairs(arr):
    return [(int(a), int(b)) for a, b in arr]


@pytest.mark.parametrize("case,kw", [("A", dict(n_jobs=1)), ("B", dict(n_jobs=3, chunk_size=97)),
                                     ("C", dict(n_jobs=2)), ("D", dict(n_jobs=4, chunk_size=64)),
                                     ("E", dict(n_jobs=1))])
def test_pair_distances_bit_exact(gold, case, kw):
    want = gold[f"{case}_d"]
    got = orc.pair_distances(gold[f"{case}_x"], _pairs(gold[f"{case}_pairs"]), **kw)
    assert got.dtype == want.dtype and got.shape == want.shape
    assert np.ascontiguousarray(got).tobytes() == np.ascontiguousarray(want).tobytes()


@pytest.mark.parametrize("case", ["A", "B", "C", "D"])
def test_f32_step_order_matches_reference_bits(gold, case):
    """The explicit sub/mul/(x+y)+z/sqrt float32 sequence is what the reference's norm() evaluates."""
    pairs = gold[f"{case}_pairs"]
    n = gold[f"{case}_x"].shape[1]
    i, j = pairs[:, 0] % n, pairs[:, 1] % n
    got = orc.pair_distances_f32_steps(gold[f"{case}_x"], i, j)
    assert np.ascontiguousarray(got).tobytes() == np.ascontiguousarray(gold[f"{case}_d"]).tobytes()


def test_chunking_and_threads_do_not_change_bits(gold):
    x, pairs = gold["B_x"], _pairs(gold["B_pairs"])
    base = orc.pair_distances(x, pairs, n_jobs=1, chunk_size=8000)
    for n_jobs, chunk in ((2, 1), (3, 50), (None, 8000), (0, 333)):
        assert np.array_equal(orc.pair_distances(x, pairs, n_jobs=n_jobs, chunk_size=chunk), base)


def test_error_behaviour_matches_reference(gold):
    x = gold["C_x"]
    with pytest.raises(ValueError):
        orc.pair_distances(x, [])
    with pytest.raises(IndexError):
        orc.pair_distances(x, [(0, 17)])


def test_all_pairs_and_trim(meta):
    agg = meta["aggregates"]
    n, m = agg["n_res"], agg["minsep"]
    pairs = orc.all_pairs(n, m)
    assert len(pairs) == (n - m) * (n - m + 1) // 2
    assert pairs[0] == (0, m) and pairs[-1] == (n - m - 1, n - 1)
    idx = list(range(100, 130))
    for key, want in agg["trim"].items():
        tn, tc = (int(v) for v in key.split(","))
        assert list(orc.trim_ca_indices(idx, tn, tc)) == want
    with pytest.raises(SystemExit):
        orc.trim_ca_indices(idx, 20, 10)


def test_mean_and_components(gold, meta):
    agg = meta["aggregates"]
    n = agg["n_res"]
    pairs = orc.all_pairs(n, agg["minsep"])
    d = orc.pair_distances(gold["B_x"], pairs, n_jobs=1)
    mean32 = orc.pair_time_mean(d)
    assert mean32.dtype == np.float32
    assert np.array_equal(mean32, gold["B_mean32"])
    pm = {pairs[k]: float(mean32[k]) for k in range(len(pairs))}
    for cut_repr, want in agg["components"].items():
        assert orc.residue_components(n, pm, float(cut_repr)) == want


def test_cohens_d_matches_formula():
    rng = np.random.default_rng(3)
    a = rng.normal(5, 1, 200).astype(np.float32)
    b = rng.normal(5.5, 2, 150).astype(np.float32)
    a64, b64 = a.astype(float), b.astype(float)
    sp = np.sqrt((199 * a64.var(ddof=1) + 149 * b64.var(ddof=1)) / 348)
    assert orc.cohens_d(a, b) == pytest.approx((a64.mean() - b64.mean()) / (sp + 1e-12), rel=1e-15)


def test_parser_matches_reference(meta, tmp_path):
    par = meta["parser"]
    for tok, want in par["tokens"]:
        assert [list(p) for p in orc.parse_pair_token(tok)] == want, tok
    for s, want in par["strings"]:
        assert [list(p) for p in orc.parse_pairs_str(s)] == want, s
    for case in par["files"]:
        f = tmp_path / "pairs.txt"
        f.write_text(case["body"])
        assert [list(p) for p in orc.read_pairs(case["pairs"], str(f))] == case["result"]
    with pytest.raises(SystemExit):
        orc.read_pairs("nothing here", "")


def test_tic_csv_rows(meta):
    c = meta["tic_csv"]
    labels = {int(k): v for k, v in c["labels"].items()}
    rows = orc.tic_csv_rows([tuple(p) for p in c["sel_pairs"]], np.array(c["tic_vec"]), labels)
    assert "\n".join(rows) + "\n" == c["csv"]
    # and the Extractor-side parser reads those rows back
    got = [orc.parse_pair_token(r) for r in rows[1:]]
    assert all(len(g) == 1 for g in got)


# ---- extension semantics: known-answer tests (no reference vectors exist for these) -------------

def test_known_answers_plain():
    x = np.zeros((1, 3, 3), np.float32)
    x[0, 1] = (3, 4, 0)
    x[0, 2] = (3, 4, 12)
    d = orc.pair_distances(x, [(0, 1), (0, 2), (1, 1)], n_jobs=1)
    assert d.tolist() == [[5.0, 13.0, 0.0]]


def test_ortho_min_image_across_face():
    L = 20.0
    x = np.array([[[0.5, 1, 1], [L - 0.5, 1, 1], [10.5, 1, 1]]])
    d = orc.pbc_pair_distances(x, [(0, 1), (0, 2), (1, 0)], box=np.array([L, L, L]))
    assert d[0] == pytest.approx([1.0, 10.0, 1.0])
    assert orc.pbc_pair_distances(x, [(0, 1)], box=None)[0, 0] == pytest.approx(L - 1.0)


def test_triclinic_truncated_octahedron_image():
    a = 100.0
    h = orc.box_matrix([a, a, a], [109.4712206, 109.4712206, 109.4712206])
    assert np.allclose(np.linalg.norm(h, axis=1), a)
    rng = np.random.default_rng(11)
    frac = rng.random((1, 400, 3))
    x = frac @ h
    pairs = [(i, i + 200) for i in range(200)]
    d = orc.pbc_pair_distances(x, pairs, box=h)
    # brute force over a 5x5x5 neighbourhood must agree with the 27-image rule
    delta = x[0, :200] - x[0, 200:]
    shifts = np.array([(p, q, r) for p in range(-2, 3) for q in range(-2, 3) for r in range(-2, 3)]) @ h
    brute = np.linalg.norm(delta[:, None, :] + shifts[None], axis=2).min(axis=1)
    assert np.allclose(d[0], brute, atol=1e-9)
    # lattice translations of single atoms leave distances unchanged
    x2 = x.copy()
    x2[0, ::3] += h[0] - 2 * h[2]
    assert np.allclose(orc.pbc_pair_distances(x2, pairs, box=h), d, atol=1e-9)
    # the single-rint shortcut is NOT enough for this cell (this is why 27 images are searched)
    s = delta @ np.linalg.inv(h)
    naive = np.linalg.norm((s - np.rint(s)) @ h, axis=1)
    assert (naive > brute + 1e-6).any()


def test_angle_known_answers_and_clamp():
    x = np.array([[[1, 0, 0], [0, 0, 0], [0, 1, 0], [2, 0, 0], [-3, 0, 0], [0.5, np.sqrt(3) / 2, 0]]], float)
    ang = orc.triplet_angles(x, [(0, 1, 2), (0, 1, 3), (0, 1, 4), (0, 1, 5), (0, 1, 1)])
    assert ang[0] == pytest.approx([90.0, 0.0, 180.0, 60.0, 0.0], abs=1e-10)
    big = np.array([[[1e8 + 1, 0, 0], [1e8, 0, 0], [1e8 + 2, 0, 0]]], float)
    assert np.isfinite(orc.triplet_angles(big, [(0, 1, 2)])).all()


def test_cutoff_thresholds_equal_double_compare():
    for cut in (3.5, 4.5, 8.0, 0.8, 0.35, 135.0, 1e-3):
        up, lo = orc.f32_at_or_above(cut), orc.f32_at_or_below(cut)
        around = np.float32(cut)
        cand = np.array([np.nextafter(around, np.float32(-np.inf)), around,
                         np.nextafter(around, np.float32(np.inf))], dtype=np.float32)
        assert np.array_equal(cand < up, cand.astype(np.float64) < cut)
        assert np.array_equal(cand > lo, cand.astype(np.float64) > cut)


def test_flags_occupancy_scores_packing():
    rng = np.random.default_rng(5)
    d = rng.uniform(2, 14, size=(37, 70)).astype(np.float32)
    f = orc.contact_flags(d, 8.0)
    words = orc.pack_flags(f)
    assert words.shape == (37, 3) and words.dtype == np.uint32
    assert np.array_equal(orc.unpack_flags(words, 70), f)
    assert np.array_equal(orc.occupancy(f), f.sum(0))
    assert np.array_equal(orc.frame_scores(f), f.sum(1))
    assert int(orc.frame_scores(f).sum()) == int(orc.occupancy(f).sum())
    bits = np.array([bin(int(w)).count("1") for w in words.ravel()]).reshape(words.shape).sum(1)
    assert np.array_equal(bits, orc.frame_scores(f))
    s1, s2 = orc.moment_sums(d)
    assert np.allclose(s1 / 37, d.mean(0), rtol=1e-6)
    assert np.allclose(s2, (d.astype(np.float64) ** 2).sum(0), rtol=1e-14)


def test_min_image_gap_flags_displacements_on_a_voronoi_face():
    L = 20.0
    d = np.array([[[10.0, 1.0, 2.0], [9.0, 1.0, 2.0], [3.0, 0.0, 0.0]]])          # exactly half a box, 1 A off, short
    gap = orc.min_image_gap(d, np.array([L, L, L]))
    assert gap[0, 0] == 0.0 and gap[0, 1] == pytest.approx(np.sqrt(11 ** 2 + 5) - np.sqrt(9 ** 2 + 5)) and gap[0, 2] > 10
    assert np.isinf(orc.min_image_gap(d, None)).all()
    x = np.zeros((1, 3, 3)); x[0, 0] = (10.0, 0, 0); x[0, 2] = (0, 3.0, 0)
    assert orc.triplet_arm_gaps(x, [(0, 1, 2)], np.array([L, L, L]))[0, 0] == 0.0


def test_weighted_frame_scores_are_exact_fixed_point_sums():
    rng = np.random.default_rng(1)
    flags = rng.random((7, 50)) < 0.4
    w = rng.random(50) * 3
    s = orc.weighted_frame_scores(flags, w)
    np.testing.assert_allclose(s, flags @ w, rtol=0, atol=50 * 2.0 ** -33)
    perm = rng.permutation(50)
    assert np.array_equal(s, orc.weighted_frame_scores(flags[:, perm], w[perm]))       # order-independent, bit for bit
    assert np.array_equal(orc.weighted_frame_scores(flags, np.ones(50)), orc.frame_scores(flags).astype(np.float64))


@pytest.mark.skipif(not os.path.exists("/root/reference/CAVIAR-1.0.py"), reason="the reference is only present in the build container")
def test_cov_golden_is_reproducible_from_the_reference(tmp_path, golden_dir):
    """tests/golden/cov_blocks.npz = outputs of the reference's own _cov_blocks / _stack_even_odd (cut out with ast)."""
    import subprocess
    import sys
    out = tmp_path / "cov.npz"
    subprocess.run([sys.executable, os.path.join(golden_dir, "make_cov_golden.py"), str(out)], check=True, stdout=subprocess.DEVNULL)
    a, b = np.load(out), np.load(os.path.join(golden_dir, "cov_blocks.npz"))
    assert sorted(a.files) == sorted(b.files)
    for k in a.files:
        assert a[k].tobytes() == b[k].tobytes(), k
