"""Residue clusters from the fused aggregates.

The reference thresholds the per-pair TIME-MEAN distance and labels the connected components of the resulting
residue graph (``build_residue_components``, CAVIAR-1.0.py:183-199, called at :497-499 and :785-786 with
``nm_cut = cluster_cut_A / 10``).  Here the mean comes straight from the kernel's fp64 ``sum`` aggregate
(``sum[p] / T``, no second pass over the (T, P) matrix) and the components are found on the device
(``FeaturePlan.residue_components`` -> ``caviar_residue_components``: label propagation in shared memory) or, for host
arrays, with scipy's connected components; labels are assigned in order of the lowest residue index, which is exactly
the order in which the reference's outer loop discovers components, so the label vectors are identical.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def pair_means_from_aggregates(sum_per_pair, n_frames: int, dtype=np.float32) -> np.ndarray:
    """``sum / T`` rounded to float32 -- the dtype the reference's ``dist.mean(axis=0)`` yields (:342, :496)."""
    return (np.asarray(sum_per_pair, dtype=np.float64) / float(n_frames)).astype(dtype)


def residue_components(n_res: int, pairs: Sequence[Tuple[int, int]], pair_means, cut: float) -> List[int]:
    """Component label of every residue; an edge (i, j) exists iff ``float(mean_ij) < cut`` (strict, in double).
    Host-side convenience over arrays (vectorised: scipy's connected components + a renumbering in the order of each
    component's lowest residue); ``FeaturePlan.residue_components`` is the same on the device."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    means = np.asarray(pair_means, dtype=np.float64)
    if len(pairs) != len(means):
        raise ValueError("one mean per pair expected")
    edges = pairs[means < float(cut)]
    graph = coo_matrix((np.ones(len(edges), dtype=np.int8), (edges[:, 0], edges[:, 1])), shape=(n_res, n_res))
    _, raw = connected_components(graph, directed=False)
    # renumber: component ids in the order of their first (= lowest) member, as the reference's outer loop finds them
    _, first = np.unique(raw, return_index=True)
    rank = np.empty(len(first), dtype=np.int64)
    rank[np.argsort(first)] = np.arange(len(first))
    return rank[raw].tolist()


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
