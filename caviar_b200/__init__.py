"""caviar_b200 -- B200-native (sm_100a) implementation of CAVIAR's pair-feature hot path.

Public surface (mirrors the reference's entry points for this path only):

* ``compute_distances_parallel``  -- CAVIAR-1.0.py:110-122
* ``FeaturePlan``                 -- pairs/triplets + PBC + cut-offs -> distances, angles, flags, scores, aggregates
* ``caviar_b200.extractor``       -- CAVIAR_Distances-Extractor.py command line

Importing the package loads ``csrc/libcaviar_b200.so``; there is no CPU fallback.
"""
from ._lib import CaviarError, lib as _clib  # noqa: F401  (fails loudly when the library is missing)
from .distances import compute_distances_parallel, release_cache  # noqa: F401
from .components import cohens_d_from_aggregates, pair_means_from_aggregates, residue_components  # noqa: F401
from .selection import all_pairs, ca_atom_indices, trim_ca_indices  # noqa: F401
from .features import (Aggregates, FeaturePlan, FrameOutputs, normalize_box, unpack_flags,  # noqa: F401
                       DEFAULT_PAIR_CUTOFF, DEFAULT_HBOND_DIST_CUTOFF, DEFAULT_HBOND_ANGLE_CUTOFF)

__all__ = ["compute_distances_parallel", "release_cache", "FeaturePlan", "FrameOutputs", "Aggregates", "CaviarError",
           "normalize_box", "unpack_flags"]
