"""Synthetic workloads of the five BASELINE.json configurations.

Units are Angstrom, coordinates float32 (T, N, 3) C-contiguous.  The same topology generator (base positions,
pair and triplet lists, box) feeds the parity sets (host, numpy) and the throughput sets (device, torch).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np


@dataclass
class Config:
    name: str
    n_atoms: int
    n_frames: int
    n_pairs: int
    n_triplets: int
    box: Optional[str]            # None | "ortho" | "triclinic"
    cell: tuple                   # edge lengths used to place atoms (and as the box when periodic)
    all_pairs_minsep: int = 0     # >0: the reference's all-pairs list (CAVIAR-1.0.py:331) instead of random pairs
    seed: int = 0
    description: str = ""


TRUNC_OCT_ANGLE = 109.4712206

CONFIGS = {
    1: Config("cfg1", 5_000, 1_000, 200, 50, None, (36.8, 36.8, 36.8), seed=1001,
              description="5k-atom protein-ligand system, 1,000 frames, 200 pairs + 50 triplets, no box"),
    2: Config("cfg2", 20_000, 10_000, 2_000, 500, "ortho", (58.5, 58.5 * 1.05, 58.5 * 0.95), seed=1002,
              description="20k-atom solvated complex, 10,000 frames, 2,000 pairs + 500 triplets, orthorhombic PBC"),
    3: Config("cfg3", 100_000, 100_000, 10_000, 2_000, "triclinic", (100.0, 100.0, 100.0), seed=1003,
              description="100k-atom system, 100,000 frames, 10,000 pairs + 2,000 triplets, triclinic PBC "
                          "(truncated octahedron a=100 A), full scoring"),
    4: Config("cfg4", 1_000, 50_000, 495_510, 0, None, (60.0, 60.0, 60.0), all_pairs_minsep=5, seed=1004,
              description="residue-level all-pairs contact map: 1,000 residues (495,510 pairs, minsep 5), 50,000 frames"),
    5: Config("cfg5", 1_000_000, 500_000, 50_000, 0, "ortho", (215.0, 215.0, 215.0), seed=1005,
              description="1M-atom membrane system, 500,000 frames, 50,000 pairs, orthorhombic PBC, frame-sharded"),
}


def cell_matrix(cfg: Config) -> np.ndarray:
    """3x3, rows = lattice vectors, lower triangular (a along x, b in the xy plane)."""
    a, b, c = cfg.cell
    if cfg.box != "triclinic":
        return np.diag([a, b, c]).astype(np.float64)
    al = be = ga = np.deg2rad(TRUNC_OCT_ANGLE)
    h = np.zeros((3, 3))
    h[0, 0] = a
    h[1, 0] = b * np.cos(ga)
    h[1, 1] = b * np.sin(ga)
    h[2, 0] = c * np.cos(be)
    h[2, 1] = c * (np.cos(al) - np.cos(be) * np.cos(ga)) / np.sin(ga)
    h[2, 2] = np.sqrt(c * c - h[2, 0] ** 2 - h[2, 1] ** 2)
    return h


def _wrap(pos: np.ndarray, h: np.ndarray) -> np.ndarray:
    frac = pos @ np.linalg.inv(h)
    frac -= np.floor(frac)
    return frac @ h


def _unit(rng, n):
    v = rng.normal(size=(n, 3))
    return v / np.linalg.norm(v, axis=1, keepdims=True)


@dataclass
class Topology:
    cfg: Config
    base: np.ndarray                  # (N, 3) float64 base positions inside the cell
    pairs: np.ndarray                 # (P, 2) int64
    triplets: np.ndarray              # (A, 3) int64
    cell: np.ndarray                  # (3, 3)
    referenced: np.ndarray = field(default=None)   # sorted distinct atom ids


def make_topology(cfg_id: int, n_pairs: Optional[int] = None, n_triplets: Optional[int] = None) -> Topology:
    cfg = CONFIGS[cfg_id]
    rng = np.random.default_rng(cfg.seed)
    N = cfg.n_atoms
    P = cfg.n_pairs if n_pairs is None else n_pairs
    A = cfg.n_triplets if n_triplets is None else n_triplets
    h = cell_matrix(cfg)
    # 3.8 A-step random walk wrapped into the cell, so that selected pairs straddle faces
    base = _wrap(np.cumsum(3.8 * _unit(rng, N), axis=0), h)
    if cfg.all_pairs_minsep > 0:
        m = cfg.all_pairs_minsep
        ii, jj = np.triu_indices(N, k=m)
        pairs = np.stack([ii, jj], axis=1).astype(np.int64)          # i-major, j >= i + m (CAVIAR-1.0.py:331)
        if n_pairs is not None:
            pairs = pairs[:n_pairs]
        triplets = np.zeros((0, 3), dtype=np.int64)
    else:
        # all atoms that get re-placed are drawn without replacement so that placements do not undo each other
        n_near = P // 2
        placed = rng.choice(N, size=n_near + 2 * A, replace=False)
        near_j, trip_h, trip_a = placed[:n_near], placed[n_near:n_near + A], placed[n_near + A:]
        free = np.setdiff1d(np.arange(N), placed, assume_unique=False)
        near_i = rng.choice(free, size=n_near, replace=True)
        base[near_j] = base[near_i] + _unit(rng, n_near) * rng.uniform(2.0, 14.0, size=(n_near, 1))
        rnd = rng.integers(0, N, size=(P - n_near, 2))
        pairs = np.concatenate([np.stack([near_i, near_j], axis=1), rnd], axis=0).astype(np.int64)
        # (D, H, A): H at 1.0 A from D, A at U(1.6, 3.5) A from H, random angle
        trip_d = rng.choice(free, size=A, replace=True)
        base[trip_h] = base[trip_d] + 1.0 * _unit(rng, A)
        base[trip_a] = base[trip_h] + _unit(rng, A) * rng.uniform(1.6, 3.5, size=(A, 1))
        triplets = np.stack([trip_d, trip_h, trip_a], axis=1).astype(np.int64)
        base = _wrap(base, h)
    ref = np.unique(np.concatenate([pairs.ravel(), triplets.ravel()]))
    return Topology(cfg, base, pairs, triplets, h, ref)


def boxes_for(top: Topology, n_frames: int, breathing: bool = True) -> Optional[np.ndarray]:
    """Per-frame boxes: (T,3) edge lengths or (T,3,3) matrices; a 0.1 % 'NPT' breathing keeps them distinct."""
    cfg = top.cfg
    if cfg.box is None:
        return None
    scale = 1.0 + (1e-3 * np.sin(0.37 * np.arange(n_frames)) if breathing else np.zeros(n_frames))
    if cfg.box == "ortho":
        return (np.diag(top.cell)[None, :] * scale[:, None]).astype(np.float32)
    return (top.cell[None, :, :] * scale[:, None, None]).astype(np.float32)


def parity_frames(top: Topology, n_frames: int, seed_offset: int = 0, sigma: float = 0.5,
                  only_referenced: bool = False) -> np.ndarray:
    """Host frames: base + N(0, sigma) per frame and coordinate, float32.

    ``only_referenced=True`` returns the compact (T, U, 3) array of the referenced atoms (what the reference's
    atom_slice hands to compute_distances_parallel, CAVIAR-1.0.py:335); remap indices with ``compact_indices``.
    """
    rng = np.random.default_rng(top.cfg.seed * 7919 + seed_offset)
    base = top.base[top.referenced] if only_referenced else top.base
    out = np.empty((n_frames,) + base.shape, dtype=np.float32)
    step = max(1, (64 << 20) // (base.size * 8))
    for t0 in range(0, n_frames, step):
        t1 = min(n_frames, t0 + step)
        out[t0:t1] = (base[None] + rng.normal(0.0, sigma, size=(t1 - t0,) + base.shape)).astype(np.float32)
    return out


def compact_indices(top: Topology):
    """Pair / triplet lists re-expressed in the compact numbering of ``top.referenced``."""
    return (np.searchsorted(top.referenced, top.pairs), np.searchsorted(top.referenced, top.triplets))


def device_frames(top: Topology, n_frames: int, device="cuda", seed: int = 0, sigma: float = 0.5,
                  only_referenced: bool = True, out=None):
    """Throughput set, generated on the device: (T, U, 3) compact or (T, N, 3) full float32 frames."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(top.cfg.seed * 104729 + seed)
    base_np = top.base[top.referenced] if only_referenced else top.base
    base = torch.from_numpy(base_np.astype(np.float32)).to(device)
    if out is None:
        out = torch.empty((n_frames,) + tuple(base.shape), dtype=torch.float32, device=device)
    step = max(1, (256 << 20) // (base.numel() * 4))
    for t0 in range(0, n_frames, step):
        t1 = min(n_frames, t0 + step)
        blk = out[t0:t1]
        blk.normal_(0.0, sigma, generator=g)
        blk.add_(base)
    return out
