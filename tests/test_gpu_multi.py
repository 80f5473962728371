"""Two-rank NCCL run of the frame-sharded path: sharded result == single-GPU result (integers bit-exact).
Needs >= 2 GPUs (gpurun --gpus 2); skipped otherwise."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, T, q):
    import torch
    import torch.distributed as dist
    import caviar_b200 as cb
    from caviar_b200 import synthetic as syn
    from caviar_b200.sharding import frame_shard, gather_frames, sharded_features_host
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        top = syn.make_topology(3, n_pairs=3000, n_triplets=500)
        x = syn.parity_frames(top, T, only_referenced=True)
        pairs, triplets = syn.compact_indices(top)
        box = syn.boxes_for(top, T)
        lo, hi = frame_shard(T, rank, world)
        with cb.FeaturePlan(x.shape[1], pairs, triplets, box="triclinic", device=rank) as plan:
            res = sharded_features_host(plan, x[lo:hi], box[lo:hi])
            # contact-map (lean) run: occupancy only -- the exchange must not expect the moments
            lean = sharded_features_host(plan, x[lo:hi], box[lo:hi], want=("flags", "score", "occ"))
            assert "sum" not in lean
        dist_all = gather_frames(res["dist"], T)
        score_all = gather_frames(res["score"], T)
        if rank == 0:
            q.put((res["occupancy"], res["sum"], res["sumsq"], dist_all, score_all, lean["occupancy"]))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import caviar_b200 as cb
    from caviar_b200 import synthetic as syn
    T, world = 101, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    occ, s1, s2, dist_all, score_all, occ_lean = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    top = syn.make_topology(3, n_pairs=3000, n_triplets=500)
    x = syn.parity_frames(top, T, only_referenced=True)
    pairs, triplets = syn.compact_indices(top)
    with cb.FeaturePlan(x.shape[1], pairs, triplets, box="triclinic", device=0) as plan:
        one = plan.features_host(x, syn.boxes_for(top, T))
    assert dist_all.tobytes() == one["dist"].tobytes()
    assert np.array_equal(score_all, one["score"])
    assert np.array_equal(occ, one["occupancy"])                 # integer aggregates: bit-exact across shardings
    assert np.array_equal(occ_lean, one["occupancy"])
    np.testing.assert_allclose(s1, one["sum"], rtol=1e-13)
    np.testing.assert_allclose(s2, one["sumsq"], rtol=1e-13)
