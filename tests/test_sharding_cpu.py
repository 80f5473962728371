"""World-size-2 gloo test of the multi-GPU host logic (frame shards + the one all-reduce of the aggregates).

The compute inside each rank is the CPU oracle here (tests only); on the GPUs the same sharding code wraps the
CUDA path (tests/test_gpu_parity.py::test_two_rank_sharding_* run it under gpurun --gpus 2)."""
import os
import socket

import numpy as np
import pytest

from caviar_b200.sharding import frame_shard


def test_frame_shards_partition_the_trajectory():
    for T in (0, 1, 7, 100, 100_001):
        for world in (1, 2, 3, 8):
            spans = [frame_shard(T, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == T
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        frame_shard(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, T, q):
    import torch
    import torch.distributed as dist
    from caviar_b200 import synthetic as syn
    from caviar_b200.sharding import gather_frames, reduce_aggregates
    from oracle import caviar_oracle as orc
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        top = syn.make_topology(1)
        x = syn.parity_frames(top, T)
        lo, hi = frame_shard(T, rank, world)
        d = orc.pair_distances(x[lo:hi], [tuple(p) for p in top.pairs.tolist()], n_jobs=1)
        flags = orc.contact_flags(d, 8.0)
        occ = torch.from_numpy(orc.occupancy(flags).astype(np.int64))
        s1, s2 = orc.moment_sums(d)
        sums = torch.from_numpy(np.stack([s1, s2]))
        reduce_aggregates(occ, sums)
        full = gather_frames(d, T)
        # occupancy-only (contact-map) exchange, with counts far beyond 2^31 to show the float64 buffer keeps them exact
        big = torch.tensor([(1 << 52) - 3 + rank, 5 * rank, 0], dtype=torch.int64)
        reduce_aggregates(big, None)
        if rank == 0:
            q.put((occ.numpy(), sums.numpy(), full, big.numpy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_reduce_matches_single_process():
    import torch.multiprocessing as mp
    from caviar_b200 import synthetic as syn
    from oracle import caviar_oracle as orc
    T, world = 37, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    occ, sums, full, big = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    top = syn.make_topology(1)
    x = syn.parity_frames(top, T)
    d = orc.pair_distances(x, [tuple(p) for p in top.pairs.tolist()], n_jobs=1)
    assert full.tobytes() == d.tobytes()                                   # rank-order concatenation
    assert np.array_equal(occ, orc.occupancy(orc.contact_flags(d, 8.0)).astype(np.int64))   # integers: exact
    s1, s2 = orc.moment_sums(d)
    np.testing.assert_allclose(sums[0], s1, rtol=1e-14)
    np.testing.assert_allclose(sums[1], s2, rtol=1e-14)
    assert big.tolist() == [2 * ((1 << 52) - 3) + 1, 5, 0]


def test_packed_exchange_buffer_round_trips():
    import torch
    from caviar_b200.sharding import pack_aggregates, unpack_aggregates
    occ = torch.tensor([0, 7, (1 << 53) - 1], dtype=torch.int64)
    s1, s2 = torch.tensor([0.5, -1.25, 3e300]), torch.tensor([1.0, 2.0, 3.0])
    s1, s2 = s1.double(), s2.double()
    buf = pack_aggregates(occ, s1, s2)
    assert buf.dtype == torch.float64 and buf.shape == (9,)
    o2, a2, b2 = torch.zeros(3, dtype=torch.int64), torch.zeros(3, dtype=torch.float64), torch.zeros(3, dtype=torch.float64)
    unpack_aggregates(buf, o2, a2, b2)
    assert torch.equal(o2, occ) and torch.equal(a2, s1) and torch.equal(b2, s2)
    assert pack_aggregates(occ).shape == (3,)


def test_numa_binding_helper_is_a_no_op_without_topology():
    from caviar_b200.sharding import _parse_cpulist, bind_to_gpu_numa_node
    assert _parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert _parse_cpulist("") == []
    import os
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa_node(0) is None or isinstance(bind_to_gpu_numa_node(0), int)      # no GPU here: None
    if bind_to_gpu_numa_node(0) is None:
        assert os.sched_getaffinity(0) == before
