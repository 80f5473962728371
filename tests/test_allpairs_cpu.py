"""The all-pairs flavour's planner (cv_plan.hpp: build_allpairs) without a GPU: the C side runs the kernel's own per-thread
set-up (cv_allpairs.cuh: ap_thread_setup) for every tile, warp and lane and checks that every pair of the reference's
list (CAVIAR-1.0.py:331) is evaluated exactly once, from its own atoms, inside the tile's atom ranges."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def cb():
    import caviar_b200
    return caviar_b200


@pytest.mark.parametrize("n,m", [(1000, 5), (1000, 1), (1024, 3), (704, 5), (900, 7), (1200, 40), (2000, 5), (1536, 1),
                                 (1001, 5), (1003, 2), (2222, 9)])      # any selection size: frames are padded to 4-atom multiples
def test_every_pair_is_evaluated_once_from_its_atoms(cb, n, m):
    pairs = cb.all_pairs(n, m)
    plan = cb.FeaturePlan(n, pairs, None, describe_only_sms=148)
    assert plan.info["mode_name"] == "allpairs" and plan.info["threads_per_cta"] == 256
    st = plan.allpairs_check()                                    # raises on any violation
    assert st["n_tiles"] == plan.info["n_tiles"] > 0
    assert st["windows"] == (len(pairs) + 127) // 128
    assert st["fast_windows"] >= 0.6 * st["windows"]
    # 2-D tiles: a tile streams a few hundred atoms, not the frame
    assert st["max_tile_atoms"] <= 560 and st["mean_tile_atoms"] <= 280
    assert st["frames_per_stage"] * st["stages"] * 12 * st["max_tile_atoms"] <= st["smem_bytes"] <= 110 * 1024
    with pytest.raises(cb.CaviarError):
        plan.layout()                                             # no index tables in this mode


def test_lists_that_are_not_the_reference_list_stay_on_the_generic_kernel(cb):
    n = 1000
    full = cb.all_pairs(n, 5)
    swapped = full.copy(); swapped[[1000, 1001]] = swapped[[1001, 1000]]
    flipped = full.copy(); flipped[77] = flipped[77][::-1]
    for pairs in (full[:-1], full[1:], swapped, flipped, np.concatenate([full, full[:1]])):
        plan = cb.FeaturePlan(n, pairs, None, describe_only_sms=148)
        assert plan.info["mode_name"] != "allpairs" and plan.allpairs_check()["n_tiles"] == 0
    # a box, triplets, forced modes, short rows
    assert cb.FeaturePlan(n, full, None, box="ortho", describe_only_sms=148).info["mode_name"] != "allpairs"
    assert cb.FeaturePlan(n, full, [(0, 1, 2)], describe_only_sms=148).info["mode_name"] != "allpairs"
    for force in ("direct", "staged"):
        assert cb.FeaturePlan(n, full, None, force=force, describe_only_sms=148).info["mode_name"] == force
    assert cb.FeaturePlan(300, cb.all_pairs(300, 5), None, describe_only_sms=148).info["mode_name"] == "zerocopy"


def test_environment_switch_disables_the_flavour(cb, monkeypatch):
    monkeypatch.setenv("CAVIAR_NO_ALLPAIRS", "1")
    assert cb.FeaturePlan(1000, cb.all_pairs(1000, 5), None, describe_only_sms=148).info["mode_name"] == "zerocopy"


def test_random_geometries(cb):
    """Any selection size (multiple of 4) and separation the planner accepts: the geometry check holds."""
    rng = np.random.default_rng(7)
    seen = 0
    for _ in range(60):
        n = int(rng.integers(600, 2800))
        m = int(rng.integers(1, 200))
        plan = cb.FeaturePlan(n, cb.all_pairs(n, m), None, describe_only_sms=int(rng.choice([148, 132, 8])))
        st = plan.allpairs_check()
        if plan.info["mode_name"] == "allpairs":
            assert st["n_tiles"] > 0 and st["max_tile_atoms"] <= 700 and st["smem_bytes"] <= 110 * 1024, (n, m, st)
            seen += 1
        else:
            assert st["n_tiles"] == 0
    assert seen >= 20


def test_shared_column_geometry(cb, monkeypatch):
    monkeypatch.setenv("CAVIAR_AP_SHARE", "1")
    for n, m in ((1000, 5), (1536, 1), (2000, 5), (1200, 40)):
        plan = cb.FeaturePlan(n, cb.all_pairs(n, m), None, describe_only_sms=148)
        assert plan.info["mode_name"] == "allpairs" and plan.allpairs_check()["n_tiles"] > 0
