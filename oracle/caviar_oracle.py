"""CPU oracle for the CAVIAR pair-feature hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  The product
(``caviar_b200``) never imports it and has no CPU fallback.

What is pinned and what is not
------------------------------
PINNED against the reference's own code (run in the build container by
``tests/golden/make_golden.py``, vectors committed under ``tests/golden/``):

* ``pair_distances``            <- CAVIAR-1.0.py:110-122  (bit-exact, float32)
* ``all_pairs``                 <- CAVIAR-1.0.py:331,667
* ``trim_ca_indices``           <- CAVIAR-1.0.py:889-894
* ``pair_time_mean``            <- CAVIAR-1.0.py:342,496,674,783 (numpy mean)
* ``residue_components``        <- CAVIAR-1.0.py:183-199 (strict ``<``, double compare)
* ``cohens_d``                  <- CAVIAR-1.0.py:539-545
* ``parse_pair_token`` / ``parse_pairs_str`` / ``read_pairs``
                                <- CAVIAR_Distances-Extractor.py:43-86
* ``tic_csv_rows``              <- CAVIAR-1.0.py:141-147 (the file the Extractor reads)

PARITY UNPINNED (the reference delegates these to the external ``cpptraj``
binary, CAVIAR_Distances-Extractor.py:143,159, which is not vendored, not
version-pinned and absent here; or does not contain them at all):

* minimum-image PBC distances (orthorhombic / triclinic)
* atom-triplet angles
* per-frame contact / H-bond flags, per-pair occupancy, per-frame score

For those the functions below are OUR fp64 definition, following the
reference's conventions (strict ``<`` in double, float32 features, ``(T, P)``
row-major, column k <-> pair k) and cpptraj's published behaviour (wrap to the
unit cell in fractional space, then search the 27 neighbouring images).
"""
from __future__ import annotations

import re
import sys
from os import cpu_count
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

# ----------------------------------------------------------------------------
# a1 -- pair distances, no PBC                    (CAVIAR-1.0.py:110-122)
# ----------------------------------------------------------------------------

def _index_columns(block: Sequence[Tuple[int, int]]):
    # CAVIAR-1.0.py:116-117 -- platform int (int64) index vectors
    left = np.fromiter((ab[0] for ab in block), dtype=int)
    right = np.fromiter((ab[1] for ab in block), dtype=int)
    return left, right


def _block_norms(coords: np.ndarray, block) -> np.ndarray:
    # CAVIAR-1.0.py:118-119 -- two fancy-index gathers, a subtraction and
    # numpy's 2-norm along the xyz axis.
    left, right = _index_columns(block)
    return np.linalg.norm(coords[:, left, :] - coords[:, right, :], axis=2)


def pair_distances(coords, pairs, n_jobs=None, chunk_size=8000) -> np.ndarray:
    """(T, n, 3) coordinates and a list of (i, j) -> (T, P) distances.

    Same blocking as the reference (CAVIAR-1.0.py:111-113,121-122): pairs are
    cut into blocks of ``chunk_size`` columns, blocks are evaluated by a
    joblib thread pool, and the per-block matrices are joined on axis 1.
    dtype follows the input (float32 in, float32 out).  Empty ``pairs``
    raises ValueError from ``np.concatenate`` just like the reference.
    """
    from joblib import Parallel, delayed

    if n_jobs is None or n_jobs <= 0:
        n_jobs = max((cpu_count() or 2) - 1, 1)
    blocks = [pairs[s:s + chunk_size] for s in range(0, len(pairs), chunk_size)]
    parts = Parallel(n_jobs=n_jobs, prefer="threads")(
        delayed(_block_norms)(coords, blk) for blk in blocks)
    return np.concatenate(parts, axis=1)


def pair_distances_f32_steps(coords: np.ndarray, i: np.ndarray, j: np.ndarray) -> np.ndarray:
    """The exact float32 operation order the CUDA kernel implements.

    numpy's ``norm(axis=2)`` on a real float32 array evaluates
    ``sqrt(add.reduce(d*d, axis=2))``; for a length-3 axis that is
    ``fl(fl(dx*dx + dy*dy) + dz*dz)`` with every product and sum rounded to
    float32 and no fused multiply-add.  Checked
    bit-for-bit against ``pair_distances`` in tests/test_oracle.py.
    """
    c = np.asarray(coords, dtype=np.float32)
    d = c[:, i, :] - c[:, j, :]
    sx = d[..., 0] * d[..., 0]
    sy = d[..., 1] * d[..., 1]
    sz = d[..., 2] * d[..., 2]
    return np.sqrt((sx + sy) + sz)


# ----------------------------------------------------------------------------
# a2 / a3 -- pair list and C-alpha trimming
# ----------------------------------------------------------------------------

def all_pairs(n_res: int, minsep: int) -> List[Tuple[int, int]]:
    """i-major list of (i, j) with j >= i + minsep   (CAVIAR-1.0.py:331,667)."""
    out = []
    for i in range(n_res):
        for j in range(i + minsep, n_res):
            out.append((i, j))
    return out


def trim_ca_indices(ca_idx, trim_n, trim_c):
    """Drop ``trim_n`` leading / ``trim_c`` trailing entries (CAVIAR-1.0.py:889-894)."""
    total = len(ca_idx)
    lead = max(0, int(trim_n))
    tail = max(0, int(trim_c))
    if lead + tail >= total:
        raise SystemExit(f"[ERR] trimN+trimC ({lead}+{tail}) >= n_res ({total})")
    stop = None if tail == 0 else -tail
    return ca_idx[lead:stop]


# ----------------------------------------------------------------------------
# a4 / a5 / a6 -- per-pair aggregation and thresholding
# ----------------------------------------------------------------------------

def pair_time_mean(dist: np.ndarray) -> np.ndarray:
    """Column mean over frames, dtype of the input (CAVIAR-1.0.py:342,496,674,783)."""
    return dist.mean(axis=0)


def residue_components(n_res: int, pair_means: Dict[Tuple[int, int], float], cut: float) -> List[int]:
    """Connected components of the graph {(i,j): mean < cut}  (CAVIAR-1.0.py:183-199).

    The comparison is a strict ``<`` between two Python floats (doubles); the
    labels are assigned in order of the lowest residue index of each component,
    so any traversal order gives the same labelling as the reference's DFS.
    """
    parent = list(range(n_res))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for (i, j), m in pair_means.items():
        if m < cut:
            ri, rj = find(i), find(j)
            if ri != rj:
                parent[max(ri, rj)] = min(ri, rj)
    labels, next_id, seen = [-1] * n_res, 0, {}
    for r in range(n_res):
        root = find(r)
        if root not in seen:
            seen[root] = next_id
            next_id += 1
        labels[r] = seen[root]
    return labels


def cohens_d(a: np.ndarray, b: np.ndarray) -> float:
    """Effect size between two columns in fp64, ddof=1 (CAVIAR-1.0.py:539-545)."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    var_a = a.var(ddof=1) if a.size > 1 else 0.0
    var_b = b.var(ddof=1) if b.size > 1 else 0.0
    pooled = np.sqrt(((a.size - 1) * var_a + (b.size - 1) * var_b) / max(a.size + b.size - 2, 1))
    return (a.mean() - b.mean()) / (pooled + 1e-12)


# ----------------------------------------------------------------------------
# a7 -- Extractor pair-label parser       (CAVIAR_Distances-Extractor.py:43-86)
# ----------------------------------------------------------------------------

_LABEL_PAIR = re.compile(r"([A-Z]{3})(\d+)\s*-\s*([A-Z]{3})(\d+)", re.IGNORECASE)
_DIST_PREFIX = re.compile(r"dist\s*\d*\s*:\s*([A-Z]{3}\s*\d+)\s*-\s*([A-Z]{3}\s*\d+)", re.IGNORECASE)

PairLabel = Tuple[str, int, str, int]


def _as_label(match) -> PairLabel:
    return (match.group(1).upper(), int(match.group(2)), match.group(3).upper(), int(match.group(4)))


def parse_pair_token(token: str) -> List[PairLabel]:
    """First ``RES12-RES34`` found in the token, else the ``dist N: ...`` form (Extractor:47-61)."""
    token = token.strip()
    hit = _LABEL_PAIR.search(token)
    if hit:
        return [_as_label(hit)]
    alt = _DIST_PREFIX.search(token)
    if alt:
        joined = alt.group(1).replace(" ", "") + "-" + alt.group(2).replace(" ", "")
        hit = _LABEL_PAIR.search(joined)
        if hit:
            return [_as_label(hit)]
    return []


def parse_pairs_str(text: str) -> List[PairLabel]:
    """Comma-separated tokens (Extractor:63-67)."""
    found: List[PairLabel] = []
    for piece in text.split(','):
        found += parse_pair_token(piece)
    return found


def read_pairs(pairs: str, pairs_file: str) -> List[PairLabel]:
    """--pairs then --pairs-file lines, first occurrence kept (Extractor:69-86)."""
    found: List[PairLabel] = []
    if pairs:
        found += parse_pairs_str(pairs)
    if pairs_file:
        with open(pairs_file, 'r', encoding='utf-8', errors='ignore') as fh:
            for line in fh:
                found += parse_pair_token(line)
    ordered = list(dict.fromkeys(found))
    if not ordered:
        sys.exit("Nessuna coppia valida trovata (es. ASP25-GLU216).")
    return ordered


def tic_csv_rows(sel_pairs, tic_vec, res3_labels) -> List[str]:
    """Lines of the ranked-loading CSV the Extractor consumes (CAVIAR-1.0.py:141-147)."""
    order = np.argsort(np.abs(tic_vec))[::-1]
    lines = ["rank,pair,res1,res2,abs_loading"]
    for rank, idx in enumerate(order, 1):
        i, j = sel_pairs[idx]
        lines.append(f"{rank},{res3_labels[i]}-{res3_labels[j]},{res3_labels[i]},"
                     f"{res3_labels[j]},{abs(tic_vec[idx]):.6f}")
    return lines


# ----------------------------------------------------------------------------
# Extensions (PARITY UNPINNED -- our fp64 definition, see module docstring)
# ----------------------------------------------------------------------------

def box_matrix(lengths, angles_deg) -> np.ndarray:
    """(a,b,c,alpha,beta,gamma) -> 3x3 with ROWS = lattice vectors, lower triangular.

    Standard crystallographic setting used by Amber/cpptraj: a along x,
    b in the xy plane.  Accepts (..., 3) arrays and returns (..., 3, 3).
    """
    lengths = np.asarray(lengths, dtype=np.float64)
    ang = np.deg2rad(np.asarray(angles_deg, dtype=np.float64))
    a, b, c = lengths[..., 0], lengths[..., 1], lengths[..., 2]
    ca, cb, cg = np.cos(ang[..., 0]), np.cos(ang[..., 1]), np.cos(ang[..., 2])
    sg = np.sin(ang[..., 2])
    h = np.zeros(lengths.shape[:-1] + (3, 3), dtype=np.float64)
    h[..., 0, 0] = a
    h[..., 1, 0] = b * cg
    h[..., 1, 1] = b * sg
    h[..., 2, 0] = c * cb
    h[..., 2, 1] = c * (ca - cb * cg) / sg
    h[..., 2, 2] = np.sqrt(np.maximum(c * c - h[..., 2, 0] ** 2 - h[..., 2, 1] ** 2, 0.0))
    return h


def _as_box_matrices(box, n_frames: int):
    """None | (3,) | (T,3) ortho lengths | (3,3) | (T,3,3) rows=lattice vectors -> (T,3,3) or None."""
    if box is None:
        return None
    b = np.asarray(box, dtype=np.float64)
    if b.shape == (3,):
        b = np.broadcast_to(b, (n_frames, 3))
    if b.shape == (3, 3) and n_frames == 3 and b[0, 1] == 0 and b[0, 2] == 0 and b[1, 2] == 0:
        b = np.broadcast_to(b, (3, 3, 3))      # T == 3: a lower-triangular (3,3) is a cell, not three length rows
    if b.shape == (n_frames, 3):
        h = np.zeros((n_frames, 3, 3))
        h[:, 0, 0], h[:, 1, 1], h[:, 2, 2] = b[:, 0], b[:, 1], b[:, 2]
        return h
    if b.shape == (3, 3):
        b = np.broadcast_to(b, (n_frames, 3, 3))
    if b.shape != (n_frames, 3, 3):
        raise ValueError(f"box shape {b.shape} not understood for {n_frames} frames")
    return np.array(b)


_IMAGE_SHIFTS = np.array([(p, q, r) for p in (-1, 0, 1) for q in (-1, 0, 1) for r in (-1, 0, 1)],
                         dtype=np.float64)


def min_image(delta: np.ndarray, box) -> np.ndarray:
    """Minimum-image displacement vectors, fp64.

    ``delta`` is (T, K, 3) (r_i - r_j); ``box`` as in ``_as_box_matrices``.
    Rule: fractional coordinates of the displacement are reduced with
    round-half-even to [-0.5, 0.5], then all 27 neighbouring lattice
    translations are tried and the shortest vector kept (first minimum on
    ties).  For an orthorhombic box this equals the per-axis wrap.
    """
    delta = np.asarray(delta, dtype=np.float64)
    h = _as_box_matrices(box, delta.shape[0])
    if h is None:
        return delta
    h_inv = np.linalg.inv(h)                                   # (T,3,3)
    frac = np.einsum('tkc,tcd->tkd', delta, h_inv)             # row-vector convention: d = s @ H
    frac -= np.rint(frac)
    best = None
    best_len = None
    for shift in _IMAGE_SHIFTS:
        cand = np.einsum('tkc,tcd->tkd', frac + shift, h)
        length = np.einsum('tkc,tkc->tk', cand, cand)
        if best is None:
            best, best_len = cand, length
        else:
            take = length < best_len
            best = np.where(take[..., None], cand, best)
            best_len = np.where(take, length, best_len)
    return best


def min_image_gap(delta: np.ndarray, box) -> np.ndarray:
    """Length difference between the two SHORTEST images of every displacement, (T, K), fp64 (inf without a box).

    A displacement within a rounding error of a Voronoi face of the lattice has two images of (nearly) the same
    length: its distance is well defined, but WHICH image is "the minimum" -- and with it the direction of the
    vector, hence an angle built on it -- is not.  The parity tests use this to set aside image-ambiguous arms."""
    delta = np.asarray(delta, dtype=np.float64)
    h = _as_box_matrices(box, delta.shape[0])
    if h is None:
        return np.full(delta.shape[:2], np.inf)
    frac = np.einsum('tkc,tcd->tkd', delta, np.linalg.inv(h))
    frac -= np.rint(frac)
    lens = np.stack([np.sqrt(np.einsum('tkc,tkc->tk', c, c))
                     for c in (np.einsum('tkc,tcd->tkd', frac + s, h) for s in _IMAGE_SHIFTS)], axis=0)
    lens.sort(axis=0)
    return lens[1] - lens[0]


def triplet_arm_gaps(coords, triplets, box=None) -> np.ndarray:
    """min_image_gap of both arms of every triplet, the smaller of the two: (T, A)."""
    c = np.asarray(coords, dtype=np.float64)
    idx = np.asarray(triplets, dtype=np.int64).reshape(-1, 3)
    gu = min_image_gap(c[:, idx[:, 0], :] - c[:, idx[:, 1], :], box)
    gv = min_image_gap(c[:, idx[:, 2], :] - c[:, idx[:, 1], :], box)
    return np.minimum(gu, gv)


def pbc_pair_distances(coords, pairs, box=None) -> np.ndarray:
    """fp64 minimum-image distances, (T, P).  With ``box=None`` plain Euclidean."""
    c = np.asarray(coords, dtype=np.float64)
    idx = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    d = min_image(c[:, idx[:, 0], :] - c[:, idx[:, 1], :], box)
    return np.sqrt(np.einsum('tkc,tkc->tk', d, d))


def triplet_angles(coords, triplets, box=None) -> np.ndarray:
    """Angle at the MIDDLE atom b of each (a, b, c), degrees, fp64, (T, A).

    u = minimg(r_a - r_b), v = minimg(r_c - r_b),
    theta = acos(clamp(u.v / (|u||v|), -1, 1)).  A zero-length arm gives 0.
    """
    c = np.asarray(coords, dtype=np.float64)
    idx = np.asarray(triplets, dtype=np.int64).reshape(-1, 3)
    u = min_image(c[:, idx[:, 0], :] - c[:, idx[:, 1], :], box)
    v = min_image(c[:, idx[:, 2], :] - c[:, idx[:, 1], :], box)
    dot = np.einsum('tkc,tkc->tk', u, v)
    nu = np.sqrt(np.einsum('tkc,tkc->tk', u, u))
    nv = np.sqrt(np.einsum('tkc,tkc->tk', v, v))
    denom = nu * nv
    with np.errstate(invalid='ignore', divide='ignore'):
        cosang = np.where(denom > 0, dot / np.where(denom > 0, denom, 1.0), 1.0)
    return np.degrees(np.arccos(np.clip(cosang, -1.0, 1.0)))


def triplet_end_distances(coords, triplets, box=None) -> np.ndarray:
    """fp64 minimum-image distance between the END atoms a and c of each triplet (donor..acceptor)."""
    idx = np.asarray(triplets, dtype=np.int64).reshape(-1, 3)
    return pbc_pair_distances(coords, idx[:, [0, 2]], box)


def f32_at_or_above(cut: float) -> np.float32:
    """Smallest float32 t with float(t) >= cut.  For float32 d:  d < cut  <=>  d < t."""
    t = np.float32(cut)
    if float(t) < cut:
        t = np.nextafter(t, np.float32(np.inf))
    return t


def f32_at_or_below(cut: float) -> np.float32:
    """Largest float32 t with float(t) <= cut.  For float32 a:  a > cut  <=>  a > t."""
    t = np.float32(cut)
    if float(t) > cut:
        t = np.nextafter(t, np.float32(-np.inf))
    return t


def contact_flags(dist, cut: float) -> np.ndarray:
    """flag[t,p] = float(dist[t,p]) < cut -- strict, compared in double (CAVIAR-1.0.py:186 convention)."""
    return np.asarray(dist).astype(np.float64) < float(cut)


def hbond_flags(end_dist, angle_deg, dist_cut: float, angle_cut: float) -> np.ndarray:
    """flag = d(a,c) < dist_cut  and  angle(a-b-c) > angle_cut, both compared in double."""
    return (np.asarray(end_dist).astype(np.float64) < float(dist_cut)) & \
           (np.asarray(angle_deg).astype(np.float64) > float(angle_cut))


def occupancy(flags: np.ndarray) -> np.ndarray:
    """Per-feature count of set flags over frames, uint64 (P,)."""
    return flags.sum(axis=0, dtype=np.uint64)


def frame_scores(flags: np.ndarray) -> np.ndarray:
    """Per-frame count of set flags, int32 (T,)."""
    return flags.sum(axis=1, dtype=np.int64).astype(np.int32)


def weighted_frame_scores(flags: np.ndarray, weights) -> np.ndarray:
    """Per-frame sum of the weights of the set flags, float64 (T,) (weights = e.g. the abs_loading
    column of CAVIAR-1.0.py:141-147).  Weights are rounded to 32.32 fixed point and summed in int64: exact, and the
    same number whatever the summation order -- the definition the device follows bit for bit."""
    wq = np.rint(np.asarray(weights, dtype=np.float64) * 4294967296.0).astype(np.int64)
    return (flags.astype(np.int64) @ wq).astype(np.float64) / 4294967296.0


def pack_flags(flags: np.ndarray) -> np.ndarray:
    """(T, F) bool -> (T, ceil(F/32)) uint32; bit b of word w is feature 32*w + b."""
    t, f = flags.shape
    words = (f + 31) // 32
    padded = np.zeros((t, words * 32), dtype=np.uint64)
    padded[:, :f] = flags
    weights = (np.uint64(1) << np.arange(32, dtype=np.uint64))
    return (padded.reshape(t, words, 32) * weights).sum(axis=2).astype(np.uint32)


def unpack_flags(words: np.ndarray, n_features: int) -> np.ndarray:
    """Inverse of ``pack_flags``."""
    w = np.asarray(words, dtype=np.uint32)
    bits = (w[:, :, None] >> np.arange(32, dtype=np.uint32)) & np.uint32(1)
    return bits.reshape(w.shape[0], -1)[:, :n_features].astype(bool)


def moment_sums(values: np.ndarray):
    """fp64 (sum, sum of squares) per column of a float32 feature matrix."""
    v = np.asarray(values).astype(np.float64)
    return v.sum(axis=0), (v * v).sum(axis=0)


def near_cutoff(values64: np.ndarray, cut: float, tol: float) -> np.ndarray:
    """Boolean mask of entries within ``tol`` of ``cut`` (the flags allowed to differ)."""
    return np.abs(np.asarray(values64, dtype=np.float64) - float(cut)) <= tol
