"""Residue clusters from the fused aggregates.

The reference thresholds the per-pair TIME-MEAN distance and labels the connected components of the resulting
residue graph (``build_residue_components``, CAVIAR-1.0.py:183-199, called at :497-499 and :785-786 with
``nm_cut = cluster_cut_A / 10``).  Here the mean comes straight from the kernel's fp64 ``sum`` aggregate
(``sum[p] / T``, no second pass over the (T, P) matrix) and the components are found with a union-find instead of
the reference's DFS; labels are assigned in order of the lowest residue index, which is exactly the order in which
the reference's outer loop discovers components, so the label vectors are identical.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def pair_means_from_aggregates(sum_per_pair, n_frames: int, dtype=np.float32) -> np.ndarray:
    """``sum / T`` rounded to float32 -- the dtype the reference's ``dist.mean(axis=0)`` yields (:342, :496)."""
    return (np.asarray(sum_per_pair, dtype=np.float64) / float(n_frames)).astype(dtype)


def residue_components(n_res: int, pairs: Sequence[Tuple[int, int]], pair_means, cut: float) -> List[int]:
    """Component label of every residue; an edge (i, j) exists iff ``float(mean_ij) < cut`` (strict, in double)."""
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    means = np.asarray(pair_means, dtype=np.float64)
    if len(pairs) != len(means):
        raise ValueError("one mean per pair expected")
    parent = np.arange(n_res, dtype=np.int64)

    def find(a: int) -> int:
        root = a
        while parent[root] != root:
            root = parent[root]
        while parent[a] != root:
            parent[a], a = root, parent[a]
        return root

    for (i, j) in pairs[means < float(cut)]:
        ri, rj = find(int(i)), find(int(j))
        if ri != rj:
            parent[max(ri, rj)] = min(ri, rj)
    labels, seen = [0] * n_res, {}
    for r in range(n_res):
        root = find(r)
        labels[r] = seen.setdefault(root, len(seen))
    return labels


def cohens_d_from_aggregates(sum_a, sumsq_a, n_a: int, sum_b, sumsq_b, n_b: int) -> np.ndarray:
    """Per-pair Cohen's d between two ensembles from the kernel's fp64 sums (CAVIAR-1.0.py:539-545, :559-561):
    ``(mean_a - mean_b) / (pooled_sd + 1e-12)`` with ``var(ddof=1)`` -- no pass over the (T, P) matrices needed."""
    sa, qa = np.asarray(sum_a, dtype=np.float64), np.asarray(sumsq_a, dtype=np.float64)
    sb, qb = np.asarray(sum_b, dtype=np.float64), np.asarray(sumsq_b, dtype=np.float64)
    ma, mb = sa / n_a, sb / n_b
    va = (qa - n_a * ma * ma) / (n_a - 1) if n_a > 1 else np.zeros_like(ma)
    vb = (qb - n_b * mb * mb) / (n_b - 1) if n_b > 1 else np.zeros_like(mb)
    va, vb = np.maximum(va, 0.0), np.maximum(vb, 0.0)
    pooled = np.sqrt(((n_a - 1) * va + (n_b - 1) * vb) / max(n_a + n_b - 2, 1))
    return (ma - mb) / (pooled + 1e-12)
