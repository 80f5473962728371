#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Runs only in the build container (needs /root/reference).  The GPU box never
executes this; it only reads the committed ``*.npz`` / ``*.json`` files.

    python tests/golden/make_golden.py

Loader recipe: CAVIAR-1.0.py imports mdtraj, matplotlib,
pyemma and seaborn at module level; none of them is installed here and none is
touched by the hot-path functions, so empty stand-in modules are registered
before the file is executed.  The Extractor imports with the stdlib only.
"""
import importlib.util
import json
import os
import sys
import tempfile
import types

import numpy as np

REF = os.environ.get("CAVIAR_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    for name in ("mdtraj", "matplotlib", "matplotlib.pyplot", "pyemma", "seaborn"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].use = lambda *a, **k: None
    mods = {}
    for tag, fname in (("caviar", "CAVIAR-1.0.py"), ("extractor", "CAVIAR_Distances-Extractor.py")):
        spec = importlib.util.spec_from_file_location(f"caviar_ref_{tag}", os.path.join(REF, fname))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mods[tag] = mod
    return mods["caviar"], mods["extractor"]


def distance_cases(ref):
    """Inputs + reference outputs of compute_distances_parallel (CAVIAR-1.0.py:110-122)."""
    out = {}
    rng = np.random.default_rng(20261019)

    def walk(t, n, scale):
        base = np.cumsum(rng.normal(0, 0.38 * scale, size=(n, 3)), axis=0)
        return (base[None] + rng.normal(0, 0.05 * scale, size=(t, n, 3))).astype(np.float32)

    # A: nm-scale random walk, random pairs incl. repeated atoms and i>j order
    x = walk(40, 300, 1.0)
    pairs = [(int(a), int(b)) for a, b in rng.integers(0, 300, size=(200, 2))]
    out["A_x"], out["A_pairs"] = x, np.array(pairs, dtype=np.int64)
    out["A_d"] = ref.compute_distances_parallel(x, pairs, n_jobs=1)

    # B: the reference's own all-pairs list (minsep 5), Angstrom scale, several chunks + threads
    n = 48
    x = walk(25, n, 10.0)
    pairs = [(i, j) for i in range(n) for j in range(i + 5, n)]
    out["B_x"], out["B_pairs"] = x, np.array(pairs, dtype=np.int64)
    out["B_d"] = ref.compute_distances_parallel(x, pairs, n_jobs=3, chunk_size=97)

    # C: negative indices wrap (numpy semantics), identical atoms -> 0, single frame
    x = walk(1, 17, 10.0)
    pairs = [(-1, 0), (3, 3), (-17, 16), (5, -2)]
    out["C_x"], out["C_pairs"] = x, np.array(pairs, dtype=np.int64)
    out["C_d"] = ref.compute_distances_parallel(x, pairs, n_jobs=2)

    # D: large magnitudes / tiny separations (rounding-sensitive), ragged last chunk
    x = (walk(16, 64, 10.0) + np.float32(5000.0)).astype(np.float32)
    x[:, 1::2, :] = x[:, 0::2, :] + rng.normal(0, 1e-3, size=(16, 32, 3)).astype(np.float32)
    pairs = [(int(a), int(b)) for a, b in rng.integers(0, 64, size=(130, 2))]
    out["D_x"], out["D_pairs"] = x, np.array(pairs, dtype=np.int64)
    out["D_d"] = ref.compute_distances_parallel(x, pairs, n_jobs=4, chunk_size=64)

    # E: float64 coordinates give float64 output (callers cast afterwards, :337)
    x64 = walk(5, 20, 1.0).astype(np.float64) + 1e-9
    pairs = [(0, 19), (4, 7), (9, 2)]
    out["E_x"], out["E_pairs"] = x64, np.array(pairs, dtype=np.int64)
    out["E_d"] = ref.compute_distances_parallel(x64, pairs, n_jobs=1)
    assert out["E_d"].dtype == np.float64 and out["A_d"].dtype == np.float32
    return out


def aggregate_cases(ref, dist):
    """mean over frames (:342), thresholded residue graph (:183-199, :496-499), trim (:889-894)."""
    out = {}
    n = 48
    pairs = [(i, j) for i in range(n) for j in range(i + 5, n)]
    d = dist["B_d"].astype(np.float32, copy=False)
    mean32 = d.mean(axis=0)
    out["B_mean32"] = mean32
    pair_mean = {pairs[k]: float(mean32[k]) for k in range(len(pairs))}
    comps = {}
    for cut in (8.0, 12.0, 18.0, 30.0, float(np.float32(mean32[7]))):
        comps[repr(cut)] = ref.build_residue_components(n, pair_mean, cut)
    trims = {}
    idx = list(range(100, 130))
    for tn, tc in ((0, 0), (3, 0), (0, 4), (2, 5), (-1, 2)):
        trims[f"{tn},{tc}"] = list(ref._trim_ca_indices(idx, tn, tc))
    return out, {"components": comps, "trim": trims, "n_res": n, "minsep": 5}


PARSER_TOKENS = [
    "ASP25-GLU216", "ALA15 - GLY26", "dist 2: ALA15-GLY26", "dist: ser 7 - thr 19", "asp25-glu216",
    "1,ASP46-GLU237,ASP46,GLU237,0.412345", "A25-G216", "ASP25–GLU216", "NASP25-GLU216X",
    "", "   ", "rank,pair,res1,res2,abs_loading", "HIS1-HIS1", "LYS0012-ARG0003 trailing LEU5-LEU6",
    "dist 12 : GLY 3 - PRO 44", "ASP25- GLU216", "ASP-25-GLU216",
]
PARSER_STRINGS = [
    "ASP25-GLU216,ALA15-GLY26", "ASP25-GLU216, bogus ,ALA15-GLY26,ASP25-GLU216", "", "junk,,",
]


def parser_cases(ext):
    out = {"tokens": [], "strings": [], "files": []}
    for tok in PARSER_TOKENS:
        out["tokens"].append([tok, [list(p) for p in ext._parse_pair_token(tok)]])
    for s in PARSER_STRINGS:
        out["strings"].append([s, [list(p) for p in ext.parse_pairs_str(s)]])
    body = "\n".join(["# comment", "ASP25-GLU216", "dist 2: ALA15-GLY26", "ASP25-GLU216", "",
                      "3,SER30-THR41,SER30,THR41,0.100000", "ALA15 - GLY26"]) + "\n"
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as fh:
        fh.write(body)
        path = fh.name
    try:
        got = ext.read_pairs("LYS9-ARG10,ASP25-GLU216", path)
    finally:
        os.remove(path)
    out["files"].append({"pairs": "LYS9-ARG10,ASP25-GLU216", "body": body, "result": [list(p) for p in got]})
    return out


def tic_csv_case(ref):
    rng = np.random.default_rng(7)
    sel_pairs = [(0, 5), (2, 9), (1, 7), (3, 8)]
    vec = rng.normal(size=4)
    labels = {i: f"{['ASP','GLU','ALA','GLY','SER','THR','LYS','ARG','HIS','PRO'][i]}{i + 22}" for i in range(10)}
    with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as fh:
        path = fh.name
    try:
        ref.save_tic_csv(path, sel_pairs, vec, labels)
        with open(path) as fh:
            text = fh.read()
    finally:
        os.remove(path)
    return {"sel_pairs": sel_pairs, "tic_vec": vec.tolist(), "labels": {str(k): v for k, v in labels.items()},
            "csv": text}


def main():
    ref, ext = load_reference()
    dist = distance_cases(ref)
    agg_np, agg_js = aggregate_cases(ref, dist)
    np.savez_compressed(os.path.join(HERE, "distances.npz"), **dist, **agg_np)
    meta = {"aggregates": agg_js, "parser": parser_cases(ext), "tic_csv": tic_csv_case(ref),
            "numpy": np.__version__, "reference_files": ["CAVIAR-1.0.py", "CAVIAR_Distances-Extractor.py"]}
    with open(os.path.join(HERE, "reference_cases.json"), "w") as fh:
        json.dump(meta, fh, sort_keys=True)
    print("wrote", os.path.join(HERE, "distances.npz"), "and reference_cases.json")


if __name__ == "__main__":
    main()
