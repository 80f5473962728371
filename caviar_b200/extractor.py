#!/usr/bin/env python3
"""Drop-in for ``CAVIAR_Distances-Extractor.py``: per-frame CA-CA distances of labelled residue pairs -> CSV.

Same command line (``--top --traj --pairs-file --pairs --numbering --offset --stride --out``), same pair-label
grammar, same CSV layout (one frame column + one column per pair titled ``ASP25-GLU216``), same ``[DONE]`` /
``[INFO]`` lines and ``sys.exit`` messages as the reference (CAVIAR_Distances-Extractor.py:96-210).  What changes
is the engine: instead of writing a cpptraj script and shelling out (:120-159), the trajectory is read in frame
blocks, the selected atoms go to the GPU and the sm_100a kernels evaluate every ``distance dNNNN selA selB``
series (minimum-image when the trajectory carries a box, as cpptraj does by default).

Additive options (a reference command line runs unchanged): ``--triplets-file``, ``--cutoff``, ``--hbond-cutoffs``,
``--noimage``, ``--npy-out``, ``--summary-out``, ``--decimals``, ``--frames-per-block``.

PARITY NOTE: cpptraj is not vendored or pinned by the reference, so its exact number formatting and frame column
name are not pinned either; we write ``#Frame`` and 4 fixed decimals (cpptraj's defaults).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import re
import sys
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

PairLabel = Tuple[str, int, str, int]


class PairGrammar:
    """The label grammar of the reference (Extractor:43-67): ``RES12-RES34`` anywhere in a token (3-letter residue
    names, ASCII hyphen, case-insensitive, first hit wins) or the ``dist 3: ALA 15 - GLY 26`` form."""

    plain = re.compile(r"([A-Z]{3})(\d+)\s*-\s*([A-Z]{3})(\d+)", re.I)
    dist_line = re.compile(r"dist\s*\d*\s*:\s*([A-Z]{3}\s*\d+)\s*-\s*([A-Z]{3}\s*\d+)", re.I)

    @classmethod
    def token(cls, text: str) -> Optional[PairLabel]:
        text = text.strip()
        m = cls.plain.search(text)
        if m is None:
            alt = cls.dist_line.search(text)
            if alt is not None:
                m = cls.plain.search(alt.group(1).replace(" ", "") + "-" + alt.group(2).replace(" ", ""))
        if m is None:
            return None
        return m.group(1).upper(), int(m.group(2)), m.group(3).upper(), int(m.group(4))

    @classmethod
    def many(cls, chunks: Iterable[str]) -> List[PairLabel]:
        return [lab for lab in (cls.token(c) for c in chunks) if lab is not None]


def read_pairs(pairs: str, pairs_file: str) -> List[PairLabel]:
    """``--pairs`` (comma separated) first, then one label per line of ``--pairs-file``; duplicates dropped keeping
    the first occurrence (= output column order); exits when nothing parses (Extractor:69-86)."""
    found: List[PairLabel] = []
    if pairs:
        found += PairGrammar.many(pairs.split(","))
    if pairs_file:
        with open(pairs_file, "r", encoding="utf-8", errors="ignore") as fh:
            found += PairGrammar.many(fh)
    unique = list(dict.fromkeys(found))
    if not unique:
        sys.exit("Nessuna coppia valida trovata (es. ASP25-GLU216).")
    return unique


def read_tic_weights(path: str, labels: Sequence[PairLabel]) -> np.ndarray:
    """Extension: the ``abs_loading`` CAVIAR writes next to every pair of its TIC file (``rank,pair,res1,res2,
    abs_loading``, CAVIAR-1.0.py:141-147) as per-pair weights, in the order of ``labels``; pairs the file does not
    weight get 0.  The first line of a label wins, like the column de-duplication."""
    by_label = {}
    with open(path, "r", encoding="utf-8", errors="ignore") as fh:
        for line in fh:
            hit = PairGrammar.many([line])
            cells = line.strip().split(",")
            if not hit or len(cells) < 2:
                continue
            try:
                w = float(cells[-1])
            except ValueError:
                continue
            by_label.setdefault(hit[0], w)
    return np.array([by_label.get(lab, 0.0) for lab in labels], dtype=np.float64)


def read_triplets(path: str) -> np.ndarray:
    """Extension: one ``a b c`` line of 1-based atom numbers per triplet (angle at b, distance a..c)."""
    rows = []
    with open(path) as fh:
        for line in fh:
            line = line.split("#")[0].replace(",", " ").split()
            if len(line) >= 3:
                rows.append([int(line[0]) - 1, int(line[1]) - 1, int(line[2]) - 1])
    return np.asarray(rows, dtype=np.int64).reshape(-1, 3)


def select_atoms(top, labels: Sequence[PairLabel], numbering: str, offset: int) -> np.ndarray:
    """Label -> CA atom index, as the cpptraj masks ``:{idx}@CA`` / ``resid {n}@CA`` would (Extractor:127-138).
    The residue NAME of a label only titles the column; it does not take part in the selection (:142)."""
    out = np.empty((len(labels), 2), dtype=np.int64)
    for k, (ra, na, rb, nb) in enumerate(labels):
        for side, num in enumerate((na, nb)):
            if numbering == "sequential":
                idx = num - offset
                if idx < 1:
                    sys.exit(f"[ERR] label-offset < 1 per {ra}{na}-{rb}{nb} (offset={offset})")
                res0 = idx - 1
                if res0 >= top.n_res:
                    sys.exit(f"[ERR] residuo {idx} oltre la topologia ({top.n_res} residui) per {ra}{na}-{rb}{nb}")
            else:
                if top.res_seq is None:
                    sys.exit("[ERR] --numbering resSeq richiede una topologia con numerazione PDB (pdb/psf/gro); "
                             "con un prmtop usare --numbering sequential e --offset")
                res0 = top.residue_by_resseq(num)
                if res0 < 0:
                    sys.exit(f"[ERR] resid {num} non trovato nella topologia per {ra}{na}-{rb}{nb}")
            atom = top.atom_of(res0, "CA")
            if atom < 0:
                sys.exit(f"[ERR] nessun atomo CA nel residuo {res0 + 1} per {ra}{na}-{rb}{nb}")
            out[k, side] = atom
    return out


def format_rows(values: np.ndarray, first_index: int, step: int, decimals: int, crlf: bool = True, copy: bool = True):
    """``index,v,v,...`` lines through the native formatter (caviar_format_csv).  Lines end in CR LF by default: the
    reference writes its rows through ``csv.writer`` (CAVIAR_Distances-Extractor.py:196-198), whose line terminator
    that is.  ``copy=False`` returns a memoryview of the formatter's buffer (what a file's ``write`` takes) instead of bytes."""
    from . import _lib
    values = np.ascontiguousarray(values, dtype=np.float32)
    rows, cols = values.shape
    # ordinary magnitudes (< 1e9) take at most 12 + decimals bytes per value; rows with wilder numbers make the
    # formatter answer CV_E_NOMEM and the call is repeated with the worst-case capacity (65 bytes per value)
    for per_value in (13 + int(decimals), 66):
        buf = np.empty(int(rows * (26 + per_value * cols)) + 16, dtype=np.uint8)   # not zero-filled, not copied
        n = _lib.lib.caviar_format_csv(values.ctypes.data, rows, cols, cols, first_index, step, decimals, int(bool(crlf)),
                                       buf.ctypes.data, buf.size)
        if n != _lib.CV_E_NOMEM:
            break
    if n < 0:
        _lib.check(int(n))
    return buf[:n].tobytes() if copy else memoryview(buf[:n])


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(description="Distanze CA-CA per frame su GPU B200 (drop-in di CAVIAR_Distances-Extractor.py).")
    ap.add_argument("--top", required=True, help="Topologia (parm7/prmtop/pdb/psf/gro)")
    ap.add_argument("--traj", required=True, help="Traiettoria MD (Amber NetCDF .nc, CHARMM/NAMD .dcd, GROMACS .trr, oppure .npy (T,N,3))")
    ap.add_argument("--pairs-file", help="File testo con coppie tipo ASP25-GLU216")
    ap.add_argument("--pairs", help='Coppie comma-separated, es: "ASP25-GLU216,ALA15-GLY26"')
    ap.add_argument("--numbering", default="sequential", choices=["sequential", "resSeq"],
                    help="sequential: indice Amber 1..N con --offset; resSeq: numerazione PDB (richiede PDB)")
    ap.add_argument("--offset", type=int, default=0, help="Offset etichette CAVIAR (solo con sequential)")
    ap.add_argument("--stride", type=int, default=1, help="Leggi 1 frame ogni N (default 1)")
    ap.add_argument("--out", default="distances_ca.csv", help="Output CSV")
    # ---- additive extensions ----
    ap.add_argument("--noimage", action="store_true", help="ignore the periodic box (cpptraj `noimage`)")
    ap.add_argument("--triplets-file", help="a b c atom numbers (1-based) per line: angle at b, H-bond flag on a..c")
    ap.add_argument("--cutoff", type=float, default=8.0, help="contact cut-off in Angstrom (default 8.0, CAVIAR-1.0.py:62)")
    ap.add_argument("--hbond-cutoffs", type=float, nargs=2, default=(3.0, 135.0), metavar=("DIST", "ANGLE"))
    ap.add_argument("--npy-out", help="also save the (T, P) float32 matrix (fast binary side output)")
    ap.add_argument("--summary-out", help="JSON with per-pair occupancy / mean / variance and per-frame scores")
    ap.add_argument("--weighted-score", action="store_true",
                    help="--pairs-file is CAVIAR's TIC csv: weight every contact by its abs_loading in the per-frame score "
                         "(written to --summary-out as frame_score_weighted)")
    ap.add_argument("--decimals", type=int, default=4)
    ap.add_argument("--frames-per-block", type=int, default=0)
    return ap


def main(argv: Optional[Sequence[str]] = None) -> int:
    args = build_parser().parse_args(argv)
    from . import FeaturePlan, trajio, unpack_flags

    labels = read_pairs(args.pairs or "", args.pairs_file or "")
    top = trajio.read_topology(args.top)
    traj = trajio.Trajectory(args.traj, n_atoms=top.n_atoms)
    if traj.n_atoms != top.n_atoms:
        sys.exit(f"[ERR] topologia ({top.n_atoms} atomi) e traiettoria ({traj.n_atoms} atomi) non corrispondono")
    pair_atoms = select_atoms(top, labels, args.numbering, args.offset)
    triplets = read_triplets(args.triplets_file) if args.triplets_file else np.zeros((0, 3), dtype=np.int64)
    colnames = [f"{ra}{na}-{rb}{nb}" for ra, na, rb, nb in labels]

    stride = args.stride if args.stride and args.stride > 1 else 1
    frame_ids = np.arange(0, traj.n_frames, stride)             # `trajin traj 1 last stride` (Extractor:122-125)
    n_sel = len(frame_ids)
    box_kind = None if args.noimage else traj.box_kind()        # cpptraj images by default when there is a box

    # compaction (CAVIAR-1.0.py:335 atom_slice): only the referenced atoms ever leave the host
    used = np.unique(np.concatenate([pair_atoms.ravel(), triplets.ravel()]))
    pairs_c = np.searchsorted(used, pair_atoms)
    trip_c = np.searchsorted(used, triplets)

    want = ["dist", "agg", "score"] + (["angle"] if len(triplets) else [])
    npy = np.lib.format.open_memmap(args.npy_out, mode="w+", dtype=np.float32,
                                    shape=(n_sel, len(labels))) if args.npy_out else None
    # ~32 MB of input + output per block: the page-locked staging buffers of the host path stay small (pinning costs
    # ~0.4 ms per MB, once) and compaction / H2D / kernels / D2H / CSV formatting of successive blocks stay cache-warm
    block = args.frames_per_block or max(64, min(8192, (32 << 20) // max(1, 12 * len(used) + 4 * len(labels))))
    occ = np.zeros(len(labels) + len(triplets), dtype=np.uint64)
    s1 = np.zeros(len(occ)); s2 = np.zeros(len(occ))
    scores, wscores = [], []
    weights = None
    if args.weighted_score:
        if not args.pairs_file:
            sys.exit("[ERR] --weighted-score legge i pesi (abs_loading) da --pairs-file")
        weights = np.concatenate([read_tic_weights(args.pairs_file, labels), np.zeros(len(triplets))])
    with FeaturePlan(len(used), pairs_c, trip_c if len(triplets) else None, box=box_kind, pair_cutoff=args.cutoff,
                     hbond_dist_cutoff=args.hbond_cutoffs[0], hbond_angle_cutoff=args.hbond_cutoffs[1]) as plan, \
            open(args.out, "wb") as fout:
        # header and rows end in CR LF: the reference writes both through csv.writer (Extractor:196-198)
        fout.write((",".join(["#Frame"] + colnames) + "\r\n").encode())
        # the atoms leave the mapped file already in the PLAN's order (per-tile blocks, CAVIAR-1.0.py:335 with the
        # plan's list): the fused kernel streams those blocks as they are, no permutation pass on the device
        order = plan.atom_order
        sel = used if order is None else used[order]
        for b0 in range(0, n_sel, block):
            ids = frame_ids[b0:b0 + block]
            lo, hi = int(ids[0]), int(ids[-1]) + 1
            xyz = traj.coords_of(sel, lo, hi, stride)           # native multi-threaded gather (+ byte order) from the mapped file
            cell = traj.cell(lo, hi, stride) if box_kind else None
            res = plan.features_host(xyz, cell, want=want, weights=weights, plan_order=order is not None)
            fout.write(format_rows(res["dist"], b0 + 1, 1, args.decimals, copy=False))       # cpptraj's 1-based set index
            if npy is not None:
                npy[b0:b0 + len(ids)] = res["dist"]
            occ += res["occupancy"]; s1 += res["sum"]; s2 += res["sumsq"]
            scores.append(res["score"])
            if weights is not None:
                wscores.append(res["score_weighted"])
    traj.close()
    if npy is not None:
        npy.flush()
    if args.summary_out:
        n = max(n_sel, 1)
        mean = s1 / n
        var = (s2 - n * mean * mean) / max(n - 1, 1)
        with open(args.summary_out, "w") as fh:
            json.dump({"frames": int(n_sel), "columns": colnames, "cutoff_A": args.cutoff,
                       "imaging": box_kind or "none", "occupancy": occ.tolist(), "mean": mean.tolist(),
                       "variance": var.tolist(),
                       "frame_score": np.concatenate(scores).tolist() if scores else [],
                       **({"pair_weights": weights[:len(labels)].tolist(),
                           "frame_score_weighted": np.concatenate(wscores).tolist() if wscores else []}
                          if weights is not None else {})}, fh)
    print(f"[DONE] Salvato CSV: {args.out}")
    print(f"[INFO] Colonne: {', '.join(colnames)}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
