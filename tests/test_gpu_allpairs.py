"""The all-pairs flavour (cv_allpairs.cuh: 2-D tiles, 8 pairs per lane, packed arithmetic) on the GPU: bit for bit the
reference's float32 arithmetic (CAVIAR-1.0.py:118-119) for the reference's own list (:331), and identical to the generic
kernel in every output -- distances, flag words, frame scores, occupancy -- for every combination of requested outputs,
ragged frame counts, strided frames, coincident atoms and non-finite coordinates."""
import itertools

import numpy as np
import pytest

from oracle import caviar_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    import caviar_b200
    return caviar_b200


def _frames(rng, T, n, scale=0.5):
    return (np.cumsum(rng.normal(0, 3.8, size=(n, 3)), axis=0)[None] + rng.normal(0, scale, size=(T, n, 3))).astype(np.float32)


def _run(cb, plan, x, want=("dist", "flags", "score"), moments=True):
    import torch
    out = plan.alloc_outputs(x.shape[0], want=want)
    agg = plan.new_aggregates(moments=moments)
    plan.run(x, None, out=out, aggregates=agg)
    torch.cuda.synchronize()
    return out, agg


@pytest.mark.parametrize("n,m,T", [(1000, 5, 37), (704, 1, 9), (1024, 3, 130), (1200, 40, 3)])
def test_same_bits_as_the_reference_arithmetic_and_as_the_generic_kernel(cb, monkeypatch, n, m, T):
    import torch
    rng = np.random.default_rng(n + m + T)
    xh = _frames(rng, T, n)
    x = torch.from_numpy(xh).cuda()
    pairs = cb.all_pairs(n, m)
    P = len(pairs)
    with cb.FeaturePlan(n, pairs, None, pair_cutoff=9.0) as plan:
        assert plan.info["mode_name"] == "allpairs"
        out, agg = _run(cb, plan, x)
        assert plan.last_launches == 2                       # fast tiles, general tiles (aggregates are added in place)
    # the reference's arithmetic, operation by operation
    want = orc.pair_distances_f32_steps(xh, pairs[:, 0], pairs[:, 1])
    got = out.dist.cpu().numpy()
    assert got.tobytes() == want.tobytes()
    fp, _ = cb.unpack_flags(out.flags.cpu().numpy(), P, 0)
    flags_want = orc.contact_flags(want, 9.0)
    assert np.array_equal(fp, flags_want)
    assert np.array_equal(out.score.cpu().numpy(), flags_want.sum(axis=1))
    assert np.array_equal(agg.occupancy.cpu().numpy(), flags_want.sum(axis=0))
    d64 = want.astype(np.float64)
    np.testing.assert_allclose(agg.sum.cpu().numpy(), d64.sum(axis=0), rtol=1e-12)
    np.testing.assert_allclose(agg.sumsq.cpu().numpy(), (d64 * d64).sum(axis=0), rtol=1e-12)
    # the generic kernel on the same plan description
    monkeypatch.setenv("CAVIAR_NO_ALLPAIRS", "1")
    with cb.FeaturePlan(n, pairs, None, pair_cutoff=9.0) as plan:
        assert plan.info["mode_name"] != "allpairs"
        ref, ref_agg = _run(cb, plan, x)
    assert torch.equal(out.dist, ref.dist) and torch.equal(out.flags, ref.flags) and torch.equal(out.score, ref.score)
    assert torch.equal(agg.occupancy, ref_agg.occupancy)


def test_every_combination_of_requested_outputs(cb):
    import torch
    rng = np.random.default_rng(3)
    n, T = 1000, 19
    x = torch.from_numpy((rng.random((T, n, 3)) * 40).astype(np.float32)).cuda()
    with cb.FeaturePlan(n, cb.all_pairs(n, 5), None, pair_cutoff=9.0) as plan:
        assert plan.info["mode_name"] == "allpairs"
        ref_out, ref_agg = _run(cb, plan, x)
        for r in range(0, 4):
            for want in itertools.combinations(("dist", "flags", "score"), r):
                for agg_kind in (None, "occ", "all"):
                    if not want and agg_kind is None:
                        continue
                    out = plan.alloc_outputs(T, want=want)
                    agg = None if agg_kind is None else plan.new_aggregates(moments=agg_kind == "all")
                    plan.run(x, None, out=out, aggregates=agg)
                    torch.cuda.synchronize()
                    for name in want:
                        assert torch.equal(getattr(out, name), getattr(ref_out, name)), (want, agg_kind, name)
                    if agg is not None:
                        assert torch.equal(agg.occupancy, ref_agg.occupancy), (want, agg_kind)
                        if agg_kind == "all":
                            assert torch.equal(agg.sum, ref_agg.sum) and torch.equal(agg.sumsq, ref_agg.sumsq)


def test_many_chunks_tail_chunks_and_repeated_runs_accumulate(cb):
    """A long run on few CTAs: many (chunk, tile) items per CTA, quarter-size tail chunks, stages with fewer frames than
    the ring holds; aggregates accumulate over calls; bit-reproducible."""
    import torch
    rng = np.random.default_rng(8)
    n, T = 704, 1237
    xh = _frames(rng, T, n, scale=1.5)
    x = torch.from_numpy(xh).cuda()
    pairs = cb.all_pairs(n, 5)
    want = orc.pair_distances_f32_steps(xh, pairs[:, 0], pairs[:, 1])
    flags_want = orc.contact_flags(want, 8.0)
    results = []
    for n_ctas in (0, 5):
        with cb.FeaturePlan(n, pairs, None, pair_cutoff=8.0, n_ctas=n_ctas) as plan:
            assert plan.info["mode_name"] == "allpairs"
            out = plan.alloc_outputs(T, want=("dist", "score"))
            agg = plan.new_aggregates()
            plan.run(x[:500], None, out=out, aggregates=agg, n_frames=500)
            o2 = plan.alloc_outputs(T - 500, want=("dist", "score"))
            plan.run(x[500:], None, out=o2, aggregates=agg)
            torch.cuda.synchronize()
            got = np.concatenate([out.dist[:500].cpu().numpy(), o2.dist.cpu().numpy()])
            assert got.tobytes() == want.tobytes()
            assert np.array_equal(np.concatenate([out.score[:500].cpu().numpy(), o2.score.cpu().numpy()]), flags_want.sum(axis=1))
            assert np.array_equal(agg.occupancy.cpu().numpy(), flags_want.sum(axis=0))
            results.append((agg.sum.clone(), agg.sumsq.clone()))
            agg2 = plan.new_aggregates()
            plan.run(x[:500], None, out=out, aggregates=agg2, n_frames=500)
            plan.run(x[500:], None, out=o2, aggregates=agg2)
            torch.cuda.synchronize()
            assert torch.equal(agg2.sum, agg.sum) and torch.equal(agg2.sumsq, agg.sumsq)     # same run, same bits
    np.testing.assert_allclose(results[0][0].cpu().numpy(), want.astype(np.float64).sum(axis=0), rtol=1e-12)


def test_strided_frames_and_the_fallback_for_unaligned_strides(cb):
    import torch
    rng = np.random.default_rng(5)
    n, T = 1000, 11
    xh = _frames(rng, T, n)
    pairs = cb.all_pairs(n, 5)
    want = orc.pair_distances_f32_steps(xh, pairs[:, 0], pairs[:, 1])
    with cb.FeaturePlan(n, pairs, None) as plan:
        for pad in (8, 3):                                   # 16-byte aligned frames: all-pairs kernel; otherwise the generic one
            wide = torch.zeros((T, 3 * n + pad), dtype=torch.float32, device="cuda")
            wide[:, :3 * n] = torch.from_numpy(xh.reshape(T, -1)).cuda()
            view = wide.as_strided((T, n, 3), (3 * n + pad, 3, 1))
            out = plan.alloc_outputs(T, want=("dist",))
            if pad % 4:
                with pytest.raises(cb.CaviarError):            # neither kernel streams frames that are not 16-byte aligned
                    plan.run(view, None, out=out)
                continue
            plan.run(view, None, out=out)
            torch.cuda.synchronize()
            assert plan.last_launches == 2
            assert out.dist.cpu().numpy().tobytes() == want.tobytes()


def test_coincident_atoms_and_non_finite_coordinates(cb):
    """Zero distances leave the fast square-root path (MUFU.RSQ of 0), NaN / inf coordinates propagate like numpy's."""
    import torch
    rng = np.random.default_rng(6)
    n, T = 1000, 6
    xh = _frames(rng, T, n)
    xh[1, 100] = xh[1, 900]                   # a fast window of row 100
    xh[2, 5] = xh[2, 10]                      # first words of a row
    xh[3, 998] = xh[3, 993]                   # the ragged end of the list
    xh[4, 400] = np.nan
    xh[5, 7, 1] = np.inf
    xh[5, 600] = 1e-20                        # pair (600, 610): the squares are denormal numbers
    xh[5, 610] = -1e-20
    xh[5, 620] = 1e-30                        # pair (620, 630): the squares underflow to zero
    xh[5, 630] = -1e-30
    pairs = cb.all_pairs(n, 5)
    with np.errstate(all="ignore"):
        want = orc.pair_distances_f32_steps(xh, pairs[:, 0], pairs[:, 1])
    with cb.FeaturePlan(n, pairs, None, pair_cutoff=8.0) as plan:
        out, agg = _run(cb, plan, torch.from_numpy(xh).cuda())
    got = out.dist.cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert got[ok].tobytes() == want[ok].tobytes()
    fp, _ = cb.unpack_flags(out.flags.cpu().numpy(), len(pairs), 0)
    with np.errstate(all="ignore"):
        assert np.array_equal(fp, orc.contact_flags(want, 8.0))
    assert np.array_equal(agg.occupancy.cpu().numpy(), fp.sum(axis=0))


def test_host_entry_point_and_operator_take_the_flavour(cb):
    """features_host (H2D, kernels, D2H inside) and the reference operator itself on a 1 000-residue selection."""
    rng = np.random.default_rng(9)
    n, T = 1000, 70
    xh = _frames(rng, T, n)
    pairs = cb.all_pairs(n, 5)
    want = orc.pair_distances_f32_steps(xh, pairs[:, 0], pairs[:, 1])
    with cb.FeaturePlan(n, pairs, None, pair_cutoff=8.0) as plan:
        res = plan.features_host(xh, want=("dist", "flags", "score", "agg"), frames_per_block=32)
    assert res["dist"].tobytes() == want.tobytes()
    flags_want = orc.contact_flags(want, 8.0)
    fp, _ = cb.unpack_flags(res["flags"], len(pairs), 0)
    assert np.array_equal(fp, flags_want) and np.array_equal(res["score"], flags_want.sum(axis=1))
    assert np.array_equal(res["occupancy"], flags_want.sum(axis=0))
    got = cb.compute_distances_parallel(xh[:9], [tuple(p) for p in pairs.tolist()])
    assert got.tobytes() == want[:9].tobytes()


def test_packed_float64_aggregates_and_frame_sharding(cb):
    """The multi-GPU exchange format: occupancy as float64 next to the moments in one buffer (added in place by the kernel);
    two frame shards accumulated into the same buffer equal the single run -- integers exactly, moments to rounding."""
    import torch
    rng = np.random.default_rng(11)
    n, T = 1000, 301
    xh = _frames(rng, T, n, scale=1.0)
    x = torch.from_numpy(xh).cuda()
    with cb.FeaturePlan(n, cb.all_pairs(n, 5), None, pair_cutoff=8.5) as plan:
        _, ref = _run(cb, plan, x, want=("flags",))
        buf, agg = plan.new_packed_aggregates()
        out = plan.alloc_outputs(T, want=("score",))
        plan.run(x[:130], None, out=out, aggregates=agg, n_frames=130)
        plan.run(x[130:], None, out=plan.alloc_outputs(T - 130, want=("score",)), aggregates=agg)
        torch.cuda.synchronize()
        F = plan.F
        assert torch.equal(buf[:F].to(torch.int64), ref.occupancy)
        assert torch.allclose(buf[F:2 * F], ref.sum, rtol=1e-13, atol=0) and torch.allclose(buf[2 * F:], ref.sumsq, rtol=1e-13, atol=0)
        lean_buf, lean = plan.new_packed_aggregates(moments=False)
        plan.run(x, None, out=plan.alloc_outputs(T, want=("flags",)), aggregates=lean)
        torch.cuda.synchronize()
        assert torch.equal(lean_buf.to(torch.int64), ref.occupancy)


def test_random_selections_separations_and_grids(cb):
    """Randomised: selection size, minimum separation, frame count, grid size, cut-off, outputs -- against the reference
    arithmetic (bits) and numpy aggregates."""
    import torch
    rng = np.random.default_rng(2024)
    done = 0
    for trial in range(40):
        n = int(rng.integers(176, 420)) * 4
        m = int(rng.integers(1, 80))
        T = int(rng.integers(1, 90))
        pairs = cb.all_pairs(n, m)
        n_ctas = int(rng.choice([0, 0, 3, 17, 64]))
        cut = float(rng.uniform(5.0, 25.0))
        with cb.FeaturePlan(n, pairs, None, pair_cutoff=cut, n_ctas=n_ctas) as plan:
            if plan.info["mode_name"] != "allpairs":
                continue
            xh = _frames(rng, T, n, scale=float(rng.uniform(0.1, 2.0)))
            lean = bool(rng.integers(0, 2))
            want_out = ("flags", "score") if lean else ("dist", "flags", "score")
            out, agg = _run(cb, plan, torch.from_numpy(xh).cuda(), want=want_out, moments=not lean)
        want = orc.pair_distances_f32_steps(xh, pairs[:, 0], pairs[:, 1])
        flags_want = orc.contact_flags(want, cut)
        fp, _ = cb.unpack_flags(out.flags.cpu().numpy(), len(pairs), 0)
        assert np.array_equal(fp, flags_want), (n, m, T, n_ctas, lean)
        assert np.array_equal(out.score.cpu().numpy(), flags_want.sum(axis=1))
        assert np.array_equal(agg.occupancy.cpu().numpy(), flags_want.sum(axis=0))
        if not lean:
            assert out.dist.cpu().numpy().tobytes() == want.tobytes(), (n, m, T, n_ctas)
            np.testing.assert_allclose(agg.sum.cpu().numpy(), want.astype(np.float64).sum(axis=0), rtol=1e-12)
        done += 1
    assert done >= 10


def test_shared_column_geometry_stays_correct(cb, monkeypatch):
    """CAVIAR_AP_SHARE=1: warps that hold the same 128 columns of rows i and i + 256 (their window grids coincide) and load
    atom j once for both -- measured slower, off by default, kept correct."""
    import torch
    monkeypatch.setenv("CAVIAR_AP_SHARE", "1")
    rng = np.random.default_rng(77)
    n, T = 1000, 45
    xh = _frames(rng, T, n)
    pairs = cb.all_pairs(n, 5)
    with cb.FeaturePlan(n, pairs, None, pair_cutoff=8.0) as plan:
        out, agg = _run(cb, plan, torch.from_numpy(xh).cuda())
        assert plan.last_launches == 3
    want = orc.pair_distances_f32_steps(xh, pairs[:, 0], pairs[:, 1])
    assert out.dist.cpu().numpy().tobytes() == want.tobytes()
    flags_want = orc.contact_flags(want, 8.0)
    fp, _ = cb.unpack_flags(out.flags.cpu().numpy(), len(pairs), 0)
    assert np.array_equal(fp, flags_want) and np.array_equal(out.score.cpu().numpy(), flags_want.sum(axis=1))
    assert np.array_equal(agg.occupancy.cpu().numpy(), flags_want.sum(axis=0))


@pytest.mark.parametrize("n", [1001, 1003, 1402])
def test_selection_sizes_that_are_not_multiples_of_four(cb, n):
    """Tiles round their atom ranges to 4-atom multiples (16-byte TMA granularity): the host entry point pads the frames on
    the way to the device, device-resident callers pad with stage(); dense frames are refused with a message when the
    generic plan of the list has a different hand-over format, and run on the generic kernel otherwise."""
    import torch
    rng = np.random.default_rng(n)
    T = 23
    xh = _frames(rng, T, n)
    pairs = cb.all_pairs(n, 5)
    want = orc.pair_distances_f32_steps(xh, pairs[:, 0], pairs[:, 1])
    flags_want = orc.contact_flags(want, 8.0)
    with cb.FeaturePlan(n, pairs, None, pair_cutoff=8.0) as plan:
        assert plan.info["mode_name"] == "allpairs"
        res = plan.features_host(xh, want=("dist", "score", "agg"), frames_per_block=10)
        assert plan.last_launches >= 3 * 2                       # three blocks, two all-pairs launches each
        assert res["dist"].tobytes() == want.tobytes()
        assert np.array_equal(res["score"], flags_want.sum(axis=1)) and np.array_equal(res["occupancy"], flags_want.sum(axis=0))
        x = torch.from_numpy(xh).cuda()
        out, agg = _run(cb, plan, plan.stage(x))
        assert plan.last_launches == 2
        assert out.dist.cpu().numpy().tobytes() == want.tobytes()
        assert np.array_equal(agg.occupancy.cpu().numpy(), flags_want.sum(axis=0))
        one = plan.alloc_outputs(1, want=("dist",))
        plan.run(plan.stage(x[3:4]), None, out=one)
        torch.cuda.synchronize()
        assert one.dist.cpu().numpy().tobytes() == want[3:4].tobytes()
        try:                                                     # dense frames: generic kernel, or a clear refusal
            out2 = plan.alloc_outputs(T, want=("dist",))
            plan.run(x, None, out=out2)
            torch.cuda.synchronize()
            assert out2.dist.cpu().numpy().tobytes() == want.tobytes()
        except cb.CaviarError as e:
            assert "pad the frame stride" in str(e)
    got = cb.compute_distances_parallel(xh[:5], [tuple(p) for p in pairs.tolist()])
    assert got.tobytes() == want[:5].tobytes()
