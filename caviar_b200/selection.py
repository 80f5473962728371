"""Input side of the hot path as CAVIAR-1.0.py builds it: C-alpha selection, N/C-terminal trimming and the
all-pairs list whose order defines the output columns."""
from __future__ import annotations

from typing import Sequence

import numpy as np


def all_pairs(n_res: int, minsep: int = 5) -> np.ndarray:
    """(P, 2) int32 array of every (i, j) with j >= i + minsep, i-major -- the list CAVIAR-1.0.py:331,667 builds as
    ``[(i, j) for i in range(n_res) for j in range(i + minsep, n_res)]`` (default minsep 5, :52), vectorised so that
    the 495 510 pairs of a 1 000-residue map cost milliseconds instead of a Python loop."""
    n_res, minsep = int(n_res), int(minsep)
    if n_res <= 0:
        return np.zeros((0, 2), dtype=np.int32)
    minsep = max(minsep, -n_res)            # range(i + minsep, n) with a negative minsep starts below i, like the reference
    i, j = np.meshgrid(np.arange(n_res, dtype=np.int64), np.arange(n_res, dtype=np.int64), indexing="ij")
    keep = (j >= i + minsep) & (j >= 0)
    return np.stack([i[keep], j[keep]], axis=1).astype(np.int32)


def trim_ca_indices(ca_idx: Sequence[int], trim_n: int, trim_c: int):
    """Drop ``trim_n`` leading and ``trim_c`` trailing residues (CAVIAR-1.0.py:889-894); same SystemExit when
    nothing would be left."""
    total = len(ca_idx)
    lead, tail = max(0, int(trim_n)), max(0, int(trim_c))
    if lead + tail >= total:
        raise SystemExit(f"[ERR] trimN+trimC ({lead}+{tail}) >= n_res ({total})")
    return ca_idx[lead:total - tail]


def ca_atom_indices(topology) -> np.ndarray:
    """0-based indices of all atoms named CA, in topology order (mdtraj ``select('name CA')``, :322-323,:662)."""
    return np.array([k for k, name in enumerate(topology.atom_names) if name == "CA"], dtype=np.int64)
