"""ctypes front end of the C oracle (``oracle/caviar_oracle.c``).  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``
may import this module; the product never does.  ``build()`` compiles the library with gcc (``oracle/Makefile``);
the ``.so`` is git-ignored and travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcaviar_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "caviar_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        res = subprocess.run(["make", "-C", HERE, "-B", "libcaviar_oracle.so"], stdout=subprocess.PIPE,
                             stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0:
            raise RuntimeError("building the C oracle failed:\n" + res.stdout)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
        L.co_pair_distances_f32.argtypes = [vp, i64, i32, vp, i64, vp, C.c_int]
        L.co_pbc_distances.argtypes = [vp, i64, i32, vp, i64, vp, C.c_int, vp, C.c_int]
        L.co_triplet_angles.argtypes = [vp, i64, i32, vp, i64, vp, C.c_int, vp, vp, C.c_int]
        L.co_fused_pass.argtypes = [vp, i64, i32, vp, i64, vp, i64, vp, C.c_int, dbl, dbl, dbl,
                                    vp, vp, vp, vp, vp, vp, vp, C.c_int]
        for f in (L.co_pair_distances_f32, L.co_pbc_distances, L.co_triplet_angles, L.co_fused_pass, L.co_version):
            f.restype = C.c_int
        _lib = L
    return _lib


def _threads(n_threads):
    return max(1, int(n_threads)) if n_threads else max((os.cpu_count() or 2) - 1, 1)


def _xyz(coords):
    x = np.ascontiguousarray(coords, dtype=np.float32)
    if x.ndim != 3 or x.shape[2] != 3:
        raise ValueError(f"coordinates must be (T, n, 3), got {x.shape}")
    return x


def _idx(rows, width, n_atoms):
    a = np.asarray(rows if rows is not None else np.zeros((0, width)), dtype=np.int64).reshape(-1, width)
    if a.size and (a.min() < -n_atoms or a.max() >= n_atoms):
        raise IndexError("index out of range")
    return np.ascontiguousarray(np.where(a < 0, a + n_atoms, a), dtype=np.int32)


def _box(box, n_frames):
    """None | (3,) | (T,3) | (3,3) | (T,3,3) -> (kind, float64 array or None); the numpy oracle's shape rules."""
    if box is None:
        return 0, None
    from .caviar_oracle import _as_box_matrices
    h = _as_box_matrices(box, n_frames)                       # (T,3,3), rows = lattice vectors
    off = h.copy()
    off[:, 0, 0] = off[:, 1, 1] = off[:, 2, 2] = 0.0
    if not off.any():                                         # diagonal: hand over edge lengths (box_kind 1)
        return 1, np.ascontiguousarray(np.stack([h[:, 0, 0], h[:, 1, 1], h[:, 2, 2]], axis=1))
    return 2, np.ascontiguousarray(h)


def _ptr(a):
    return None if a is None else a.ctypes.data


def pair_distances_f32(coords, pairs, n_threads=None) -> np.ndarray:
    """CAVIAR-1.0.py:110-122 on float32 coordinates, bit for bit (no box)."""
    x = _xyz(coords)
    idx = _idx(pairs, 2, x.shape[1])
    if len(idx) == 0:
        raise ValueError("need at least one array to concatenate")
    out = np.empty((x.shape[0], len(idx)), dtype=np.float32)
    lib().co_pair_distances_f32(_ptr(x), x.shape[0], x.shape[1], _ptr(idx), len(idx), _ptr(out), _threads(n_threads))
    return out


def pbc_pair_distances(coords, pairs, box=None, n_threads=None) -> np.ndarray:
    x = _xyz(coords)
    idx = _idx(pairs, 2, x.shape[1])
    kind, b = _box(box, x.shape[0])
    out = np.empty((x.shape[0], len(idx)), dtype=np.float64)
    lib().co_pbc_distances(_ptr(x), x.shape[0], x.shape[1], _ptr(idx), len(idx), _ptr(b), kind, _ptr(out), _threads(n_threads))
    return out


def triplet_angles(coords, triplets, box=None, n_threads=None, with_end_distances=False):
    x = _xyz(coords)
    idx = _idx(triplets, 3, x.shape[1])
    kind, b = _box(box, x.shape[0])
    ang = np.empty((x.shape[0], len(idx)), dtype=np.float64)
    end = np.empty((x.shape[0], len(idx)), dtype=np.float64) if with_end_distances else None
    lib().co_triplet_angles(_ptr(x), x.shape[0], x.shape[1], _ptr(idx), len(idx), _ptr(b), kind, _ptr(ang), _ptr(end),
                            _threads(n_threads))
    return (ang, end) if with_end_distances else ang


def fused_pass(coords, pairs, triplets, box=None, pair_cutoff=8.0, hbond_dist_cutoff=3.0, hbond_angle_cutoff=135.0,
               n_threads=None, per_frame=True) -> dict:
    """The whole path in one pass per frame on the host cores: float32 features, flags, frame scores, occupancy and
    fp64 moment sums.  ``per_frame=False`` keeps only the aggregates and scores (nothing of size T x P is written)."""
    x = _xyz(coords)
    T, n = x.shape[0], x.shape[1]
    p = _idx(pairs, 2, n)
    t = _idx(triplets, 3, n)
    P, A = len(p), len(t)
    kind, b = _box(box, T)
    res = {"score": np.zeros(T, dtype=np.int32), "occupancy": np.zeros(P + A, dtype=np.uint64),
           "sum": np.zeros(P + A), "sumsq": np.zeros(P + A)}
    if per_frame:
        res["dist"] = np.empty((T, P), dtype=np.float32)
        res["angle"] = np.empty((T, A), dtype=np.float32)
        res["flags"] = np.zeros((T, P + A), dtype=np.uint8)
    used = lib().co_fused_pass(_ptr(x), T, n, _ptr(p), P, _ptr(t), A, _ptr(b), kind, float(pair_cutoff),
                               float(hbond_dist_cutoff), float(hbond_angle_cutoff), _ptr(res.get("dist")),
                               _ptr(res.get("angle")), _ptr(res.get("flags")), _ptr(res["score"]),
                               _ptr(res["occupancy"]), _ptr(res["sum"]), _ptr(res["sumsq"]), _threads(n_threads))
    if used < 0:
        raise MemoryError("C oracle: out of memory")
    res["threads"] = used
    return res
