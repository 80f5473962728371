"""Frame sharding across the GPUs of one box.

Frames are independent (every output row depends on one frame, CAVIAR-1.0.py:118-119), so rank r owns the
contiguous frame range ``frame_shard(T, r, world)``, the pair / triplet lists are replicated, per-frame outputs
stay rank-local (concatenated in rank order when gathered) and the path has exactly ONE exchange step: ONE sum
all-reduce of the per-feature aggregates {occupancy, sum, sum of squares} packed into a float64 buffer -- NCCL over
NVLink on the GPUs, gloo in the CPU tests.  Integer occupancies stay bit-exact (counts < 2^53 are exact in float64).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def _parse_cpulist(text: str):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device: int = 0) -> Optional[int]:
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off (Linux sysfs), so that the page-locked
    host buffers it allocates afterwards are node-local: with one process per GPU pushing ~50 GB/s over PCIe each, a
    buffer on the other socket sends every byte across the inter-socket link.  Returns the node, or None when the
    topology is not visible (single node, container without sysfs, ...) -- then nothing is changed."""
    import os
    try:
        import torch
        bus = torch.cuda.get_device_properties(device).pci_bus_id
        dev = torch.cuda.get_device_properties(device).pci_device_id
        dom = torch.cuda.get_device_properties(device).pci_domain_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set(_parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read()))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def frame_shard(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first ``n_frames % world`` ranks get one extra frame."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    base, extra = divmod(int(n_frames), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_aggregates(occupancy, sum_=None, sumsq=None, out=None):
    """The exchange buffer of the path: ONE float64 tensor [occupancy | sum | sumsq] (3F, or F without the moments).

    Occupancy counts are frame counts: integers below 2^53 are exact in float64 and so is their sum in any order, so
    the integer bit-exactness of the reduced occupancy survives the single-dtype buffer (checked: ``unpack`` refuses a
    value that is not an integer).  One collective instead of two halves the latency-bound exchange step.
    """
    import torch
    F = occupancy.shape[0]
    n = F if sum_ is None else 3 * F
    if out is None:
        out = torch.empty(n, dtype=torch.float64, device=occupancy.device)
    out[:F].copy_(occupancy)                  # int64 -> float64, exact below 2^53
    if sum_ is not None:
        out[F:2 * F].copy_(sum_)
        out[2 * F:].copy_(sumsq)
    return out


def unpack_aggregates(buf, occupancy, sum_=None, sumsq=None):
    """Inverse of ``pack_aggregates`` into the caller's tensors (occupancy int64)."""
    F = occupancy.shape[0]
    occupancy.copy_(buf[:F])                  # float64 -> int64: exact for the integers a sum of counts can be
    if sum_ is not None:
        sum_.copy_(buf[F:2 * F])
        sumsq.copy_(buf[2 * F:])
    return occupancy, sum_, sumsq


def reduce_aggregates(occupancy, sums=None, group=None, buf=None):
    """In-place sum all-reduce of the aggregates over the process group: ONE collective over the packed float64
    buffer (8*F or 24*F bytes) -- latency-bound on NVLink 5, so NCCL's stock all-reduce is used.

    ``occupancy``: int64 tensor (F,); ``sums``: float64 tensor (2, F) = [sum, sumsq], or None for occupancy-only
    (contact-map) runs.
    """
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return occupancy, sums
    buf = pack_aggregates(occupancy, None if sums is None else sums[0], None if sums is None else sums[1], out=buf)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    unpack_aggregates(buf, occupancy, None if sums is None else sums[0], None if sums is None else sums[1])
    return occupancy, sums


def gather_frames(local, n_frames: int, group=None, dst: int = 0):
    """Concatenate rank-local per-frame arrays (numpy, first axis = the rank's frames) in rank order on ``dst``."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts = [None] * world if rank == dst else None
    dist.gather_object(local, parts, dst=dst, group=group)
    if rank != dst:
        return None
    out = np.concatenate(parts, axis=0)
    assert out.shape[0] == n_frames
    return out


def sharded_features_host(plan, xyz_shard: np.ndarray, box_shard=None, group=None, device=None, **kw) -> dict:
    """Run ``plan.features_host`` on this rank's frame shard and all-reduce the aggregates.

    Returns the rank-local per-frame outputs plus GLOBAL occupancy / sum / sumsq (identical on every rank).
    """
    import torch
    res = plan.features_host(xyz_shard, box_shard, **kw)
    if "occupancy" in res:
        import torch.distributed as dist
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            backend = dist.get_backend(group)
            dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
            occ = torch.from_numpy(res["occupancy"].astype(np.int64)).to(dev)
            # lean (contact-map) runs return the occupancy alone: the moments are reduced only when they exist
            sums = torch.from_numpy(np.stack([res["sum"], res["sumsq"]])).to(dev) if "sum" in res else None
            reduce_aggregates(occ, sums, group)
            res["occupancy"] = occ.cpu().numpy().astype(np.uint64)
            if sums is not None:
                res["sum"], res["sumsq"] = sums[0].cpu().numpy(), sums[1].cpu().numpy()
    return res
