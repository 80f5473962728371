"""CPU-side checks of the C-ABI library: it loads, exports every declared symbol, and plans sensibly."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cb():
    import __graft_entry__ as g
    g.build()
    import caviar_b200
    return caviar_b200


def test_header_symbols_all_exported(cb):
    from caviar_b200 import _lib
    header = open(os.path.join(ROOT, "include", "caviar_b200.h")).read()
    declared = set(re.findall(r"\b(caviar_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(_lib.lib, name)
    assert _lib.lib.caviar_abi_version() == 3


def test_no_cpu_fallback(cb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cb.CaviarError) as e:
        cb.FeaturePlan(10, [(0, 1)])
    assert e.value.code == -3
    with pytest.raises(cb.CaviarError):
        cb.compute_distances_parallel(np.zeros((2, 4, 3), np.float32), [(0, 1)])


def test_reference_error_behaviour_needs_no_gpu(cb):
    x = np.zeros((2, 4, 3), np.float32)
    with pytest.raises(ValueError):
        cb.compute_distances_parallel(x, [])
    with pytest.raises(IndexError):
        cb.compute_distances_parallel(x, [(0, 4)])


@pytest.mark.parametrize("cfg_id,mode", [(1, "staged"), (2, "staged"), (3, "staged"), (4, "allpairs")])
def test_plan_describe_configs(cb, cfg_id, mode):
    from caviar_b200 import synthetic as syn
    top = syn.make_topology(cfg_id)
    plan = cb.FeaturePlan(top.cfg.n_atoms, top.pairs, top.triplets, box=top.cfg.box, describe_only_sms=148)
    info = plan.info
    assert info["mode_name"] == mode
    P, A = len(top.pairs), len(top.triplets)
    assert info["n_features"] == P + A
    assert info["n_flag_words"] == (P + 31) // 32 + (A + 31) // 32
    assert info["n_distinct_atoms"] == len(top.referenced)
    box_b = {None: 0, "ortho": 12, "triclinic": 36}[top.cfg.box]
    assert info["algorithmic_bytes_per_frame"] == 12 * len(top.referenced) + box_b + 4 * P + 4 * A + 4 * info["n_flag_words"] + 4
    assert info["n_ctas"] <= 296 and info["smem_bytes"] <= 113 * 1024 * (2 if info["n_ctas"] <= 148 else 1)
    if mode == "staged":
        assert info["staged_floats_per_frame"] % 12 == 0
        assert info["staged_floats_per_frame"] >= 3 * len(top.referenced)
        assert info["staged_floats_per_frame"] <= 3 * (2 * P + 3 * A) + 12 * info["n_tiles"]


def test_plan_forced_modes_and_direct(cb):
    rng = np.random.default_rng(0)
    pairs = rng.integers(0, 4096, size=(1000, 2))              # two tiles that share few atoms
    for force, mode in (("direct", "direct"), ("staged", "staged"), (None, "staged")):
        info = cb.FeaturePlan(4096, pairs, None, describe_only_sms=148, force=force).info
        assert info["mode_name"] == mode
    # six tiles that each touch a third of a 48-KB frame: per-tile blocks (2.4x the bytes) beat uncoalesced gathers
    pairs = rng.integers(0, 4096, size=(5000, 2))
    assert cb.FeaturePlan(4096, pairs, None, describe_only_sms=148).info["mode_name"] == "staged"
    # ... of a 12-KB frame: stream the frame as it is
    pairs = rng.integers(0, 1024, size=(5000, 2))
    assert cb.FeaturePlan(1024, pairs, None, describe_only_sms=148).info["mode_name"] == "zerocopy"
    assert cb.FeaturePlan(1024, pairs, None, box="ortho", describe_only_sms=148).info["mode_name"] == "zerocopy"   # raw floats: a box changes nothing


def test_thresholds_match_double_compare():
    """smallest float32 >= cut: d < t  <=>  (double)d < cut -- the CAVIAR-1.0.py:186 comparison."""
    from oracle import caviar_oracle as orc
    for cut in (3.5, 4.5, 8.0, 0.8, 135.0):
        t = orc.f32_at_or_above(cut)
        for d in (np.nextafter(np.float32(cut), np.float32(0)), np.float32(cut), np.nextafter(np.float32(cut), np.float32(1e9))):
            assert (d < t) == (float(d) < cut)


def test_index_array_fast_path_equals_numpy(cb):
    """A list of plain-int tuples (the reference's pair list, CAVIAR-1.0.py:331) takes a flat-iterator path;
    everything else goes through np.asarray -- same result, same errors."""
    from caviar_b200.features import _index_array
    pairs = [(i, j) for i in range(40) for j in range(i + 5, 40)]
    a = _index_array(pairs, 2, 40, "pairs")
    assert a.dtype == np.int32 and a.flags.c_contiguous
    assert np.array_equal(a, np.asarray(pairs))
    assert np.array_equal(_index_array([(-1, 0), (3, -40)], 2, 40, "pairs"), [[39, 0], [3, 0]])
    assert np.array_equal(_index_array([(np.int64(1), 2)], 2, 40, "pairs"), [[1, 2]])       # slow path
    assert np.array_equal(_index_array(np.asarray(pairs), 2, 40, "pairs"), a)
    with pytest.raises(IndexError):
        _index_array([(0, 40)], 2, 40, "pairs")
    with pytest.raises(IndexError):
        _index_array([(0.5, 1)], 2, 40, "pairs")
    with pytest.raises(ValueError):
        _index_array([(0, 1, 2)], 2, 40, "pairs")
    with pytest.raises((ValueError, IndexError)):
        _index_array([(0, 1), (2,)], 2, 40, "pairs")
    with pytest.raises((IndexError, OverflowError)):
        _index_array([(0, 2 ** 70)], 2, 40, "pairs")


def test_staged_blocks_are_bank_conflict_free(cb):
    """Planner: a tile's atoms are ordered so that the 32 lanes of every gather hit distinct shared-memory banks."""
    rng = np.random.default_rng(3)
    n = 20000
    pairs = rng.integers(0, n, size=(5000, 2))
    trips = rng.integers(0, n, size=(1000, 3))
    info = cb.FeaturePlan(n, pairs, trips, box="triclinic", describe_only_sms=148).info
    # greedy colouring: atoms shared between gathers may keep a few collisions (13 000 references here)
    assert info["mode_name"] == "staged" and info["gather_bank_conflicts"] <= 40
    once = rng.permutation(n)[:10000].reshape(-1, 2)       # every atom referenced once: always conflict-free
    info = cb.FeaturePlan(n, once, None, box=None, force="staged", describe_only_sms=148).info
    assert info["gather_bank_conflicts"] == 0
    # heavy sharing (few atoms, many features) cannot always be conflict-free, but must still plan
    pairs = rng.integers(0, 6000, size=(40000, 2))
    info = cb.FeaturePlan(6000, pairs, None, box=None, force="staged", describe_only_sms=148).info
    assert info["gather_bank_conflicts"] >= 0 and info["staged_floats_per_frame"] % 4 == 0


def test_plan_order_is_raw_float_blocks_whatever_the_box(cb):
    """The staged layout is a pure permutation of the caller's atoms (ABI 3): no fixed-point blocks, the same plan
    order with and without a box, and small aligned frames are streamed as they are even with a box."""
    rng = np.random.default_rng(8)
    n = 256                                                     # small, aligned frames: zero-copy candidates
    pairs = rng.integers(0, n, size=(4000, 2))
    trips = rng.integers(0, n, size=(500, 3))
    for box in (None, "ortho", "triclinic"):
        for tr in (None, trips):
            info = cb.FeaturePlan(n, pairs, tr, box=box, describe_only_sms=148).info
            assert info["staged_fixedpoint_floats"] == 0 and info["mode_name"] == "zerocopy"
    n = 5000
    pairs, trips = rng.integers(0, n, size=(3001, 2)), rng.integers(0, n, size=(700, 3))
    orders = []
    for box in (None, "ortho", "triclinic"):
        plan = cb.FeaturePlan(n, pairs, trips, box=box, describe_only_sms=148)
        assert plan.info["mode_name"] == "staged" and plan.info["staged_fixedpoint_floats"] == 0
        order = plan.atom_order
        assert order.dtype == np.int32 and 3 * len(order) == plan.info["staged_floats_per_frame"] and len(order) % 4 == 0
        assert np.array_equal(np.unique(order), np.unique(np.concatenate([pairs.ravel(), trips.ravel()])))
        assert np.array_equal(3 * order, plan.layout()["src_off"][0::3])
        orders.append(order)
    assert np.array_equal(orders[0], orders[1]) and np.array_equal(orders[0], orders[2])
    assert cb.FeaturePlan(n, pairs, None, box=None, force="direct", describe_only_sms=148).atom_order is None


def test_large_overlapping_selections_do_not_blow_up_the_staged_layout(cb):
    """All pairs of thousands of atoms: every tile needs ~900 atoms of the same frame.  Per-tile blocks would stage
    the frame thousands of times over; the planner must share one block (<= 100 KB) or gather directly."""
    def rows_against_all(n):                                      # every 50th atom against every atom, i-major
        i = np.repeat(np.arange(0, n, 50), n)
        return np.stack([i, np.tile(np.arange(n), len(i) // n)], 1).astype(np.int32)

    for n, want in ((1000, "zerocopy"), (1001, "staged"), (3000, "direct"), (9000, "direct")):
        pairs = rows_against_all(n)
        info = cb.FeaturePlan(n, pairs, None, box=None, describe_only_sms=148).info
        assert info["mode_name"] == want, (n, info)
        assert info["staged_floats_per_frame"] <= 3 * (n + 3)
    pairs = rows_against_all(5000)
    trips = np.stack([np.arange(0, 300), np.arange(1, 301), np.arange(2, 302)], 1)
    info = cb.FeaturePlan(5000, pairs, trips, box="triclinic", describe_only_sms=148).info
    assert info["mode_name"] == "direct" and info["stages"] == 0 and info["smem_bytes"] == 0       # LDG from the caller's frames
    info = cb.FeaturePlan(5000, pairs, None, box="ortho", describe_only_sms=148).info
    assert info["mode_name"] == "direct" and info["staged_floats_per_frame"] == 0


def _check_layout(cb, n, pairs, trips, box, force=None):
    """Every feature slot of every tile must resolve, through the index tables and the staged layout, to the atoms
    the caller's list names -- whatever the planner decided (per-tile blocks in bank order, shared block, compact
    per-kind blocks, zero-copy, direct)."""
    plan = cb.FeaturePlan(n, pairs, trips, box=box, force=force, describe_only_sms=148)
    lay, info = plan.layout(), plan.info
    staged = info["mode_name"] == "staged"
    src = lay["src_off"]
    if staged:
        assert len(src) % 12 == 0 and (src.reshape(-1, 3) % 3 == [0, 1, 2]).all()
        assert (np.diff(src.reshape(-1, 3), axis=1) == 1).all() and src.min() >= 0 and src.max() < 3 * n
    P = 0 if pairs is None else len(pairs)
    seen_pairs = seen_trips = 0
    for kind, f_begin, n_feat, idx_off, grp_off, grp_floats, fixed, _ in lay["tiles"]:
        want = (np.asarray(pairs)[f_begin:f_begin + n_feat] if kind == 0 else np.asarray(trips)[f_begin:f_begin + n_feat])
        for role in range(want.shape[1]):
            off = lay["idx"][role, idx_off:idx_off + n_feat]
            assert (off % 3 == 0).all() and off.min() >= 0
            if staged:
                assert off.max() + 2 < grp_floats and grp_off + grp_floats <= len(src)
                atoms = src[grp_off + off] // 3
            else:
                atoms = off // 3
            assert np.array_equal(atoms, want[:, role]), (kind, f_begin, role)
        assert fixed == 0                                          # ABI 3: raw float32 blocks only
        seen_pairs += n_feat if kind == 0 else 0
        seen_trips += n_feat if kind == 1 else 0
    assert seen_pairs == P and seen_trips == (0 if trips is None else len(trips))
    return info


def test_planner_layout_resolves_every_feature_to_its_atoms(cb):
    rng = np.random.default_rng(21)
    # random lists: per-tile blocks in bank-conflict-free order, with and without a box / triplets
    n = 5000
    pairs, trips = rng.integers(0, n, size=(3001, 2)), rng.integers(0, n, size=(1234, 3))
    for box in (None, "ortho", "triclinic"):
        info = _check_layout(cb, n, pairs, trips, box)
        assert info["mode_name"] == "staged" and info["stages"] > 0
    _check_layout(cb, n, pairs, None, "ortho", force="direct")
    # heavy sharing over few atoms: holes in the bank order, shared block, zero-copy
    few = rng.integers(0, 60, size=(5000, 2))
    assert _check_layout(cb, 60, few, None, None)["mode_name"] == "zerocopy"
    assert _check_layout(cb, 61, few, None, None)["mode_name"] == "staged"
    _check_layout(cb, 60, few, rng.integers(0, 60, size=(300, 3)), "triclinic")
    _check_layout(cb, 60, few, None, None, force="staged")
    # large overlapping selection: gathered from the caller's frames in global memory, with or without a box
    rows = np.repeat(np.arange(0, 3000, 50), 3000)
    big = np.stack([rows, np.tile(np.arange(3000), 60)], 1)
    info = _check_layout(cb, 3000, big, rng.integers(0, 3000, size=(100, 3)), "ortho")
    assert info["mode_name"] == "direct" and info["stages"] == 0
    assert _check_layout(cb, 3000, big, None, None)["mode_name"] == "direct"


def test_plain_c_host_links_against_the_abi(cb, tmp_path):
    """examples/pair_distances.c: a C program binds the same header and library (no Python below the ABI).  Without a
    GPU the library refuses loudly (CV_E_NO_DEVICE) -- there is no CPU fallback to fall into."""
    import shutil
    import subprocess
    import torch
    from caviar_b200.build import LIB_PATH
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "pair_distances")
    libdir = os.path.dirname(LIB_PATH)
    subprocess.run([gcc, "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "pair_distances.c"),
                    "-o", exe, "-L", libdir, "-lcaviar_b200", f"-Wl,-rpath,{libdir}"], check=True)
    res = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert "ABI version 3" in res.stdout
    if torch.cuda.is_available():
        assert res.returncode == 0 and "frame 1: 5.0000 20.0000 5.0000" in res.stdout
    else:
        assert res.returncode == 2 and "no CPU fallback" in res.stdout


def test_run_chunks_cover_every_frame_once(cb):
    """caviar_plan_chunks: the frame chunks of the work queue (make_chunks in cv_api.cu) tile [0, n_frames) exactly, fit the
    partial-aggregate slots, and runs too short to give every CTA an item are cut finer (down to 4 frames)."""
    from caviar_b200 import synthetic as syn
    rng = np.random.default_rng(11)
    for cfg_id in (1, 2, 3):
        top = syn.make_topology(cfg_id)
        pairs, trips = syn.compact_indices(top)
        plan = cb.FeaturePlan(len(top.referenced), pairs, trips if len(trips) else None, box=top.cfg.box, describe_only_sms=148)
        slots, n_tiles, n_ctas = plan.info["partial_slots"], plan.info["n_tiles"], plan.info["n_ctas"]
        for T in [1, 2, 3, 5, 31, 32, 33, 250, 1000, 4000, 4096, 10000, 100000, 1234567] + [int(t) for t in rng.integers(1, 300000, 40)]:
            ck = plan.chunks(T)
            cf, tf, nb, nc = ck["chunk_frames"], ck["tail_frames"], ck["n_big_chunks"], ck["n_chunks"]
            assert cf >= 1 and tf >= 1 and 0 <= nb <= nc <= slots, (cfg_id, T, ck)
            start = lambda c: c * cf if c < nb else nb * cf + (c - nb) * tf
            assert start(nc - 1) < T <= start(nc), (cfg_id, T, ck)          # last chunk is non-empty and reaches the end
            if nb < nc:
                assert nb * cf <= T                                          # the big chunks never run past the end
            if nb == nc and cf < 32:                                         # the fine chunks of a short run:
                assert cf >= 4 and (nc * n_tiles <= n_ctas + n_tiles or cf == 4), (cfg_id, T, ck)   # about one item per CTA
        if cfg_id == 1:                                                      # 2 tiles on 296 CTAs: 148 chunks wanted
            assert plan.chunks(1000)["chunk_frames"] == 7 and plan.chunks(4000)["chunk_frames"] == 28
            assert plan.chunks(250)["chunk_frames"] == 4 and plan.chunks(100000)["chunk_frames"] >= 32
        with pytest.raises(cb.CaviarError):
            plan.chunks(0)
