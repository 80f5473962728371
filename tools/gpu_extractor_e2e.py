#!/usr/bin/env python3
"""GPU box: wall clock of the Extractor drop-in on a synthetic Amber NetCDF + prmtop (page cache hot), by phase.
   python tools/gpu_extractor_e2e.py [n_res=3000] [frames=4000] [pairs=2000]"""
import os, sys, time, json, tempfile, cProfile, pstats, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from caviar_b200 import extractor, trajio

n_res = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
P = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
rng = np.random.default_rng(0)
res3 = ["ASP", "GLU", "ALA", "GLY", "SER", "THR", "LYS", "ARG", "HIS", "PRO"]
names, labels, first = [], [], []
for r in range(n_res):
    first.append(len(names)); labels.append(res3[r % 10]); names += ["N", "H", "CA", "HA", "C", "O", "CB"]
N = len(names)
L = np.array([90.0, 95.0, 85.0])
d = tempfile.mkdtemp(dir="/tmp")
top, nc = os.path.join(d, "s.prmtop"), os.path.join(d, "t.nc")
trajio.write_prmtop(top, names, labels, first)
base = (rng.random((N, 3)) * L).astype(np.float32)
xyz = np.empty((T, N, 3), dtype=np.float32)
for t0 in range(0, T, 256):
    xyz[t0:t0 + 256] = base[None] + rng.normal(0, 0.4, size=(min(256, T - t0), N, 3)).astype(np.float32)
trajio.write_netcdf(nc, xyz, L)
del xyz
pr = set()
while len(pr) < P:
    a, b = rng.integers(0, n_res, 2)
    if a != b: pr.add((int(a), int(b)))
pf = os.path.join(d, "pairs.txt")
open(pf, "w").write("\n".join(f"{labels[a]}{a + 1}-{labels[b]}{b + 1}" for a, b in pr) + "\n")
print(json.dumps({"atoms": N, "frames": T, "pairs": P, "file_GB": os.path.getsize(nc) / 1e9}))
for rep in range(2):
    prof = cProfile.Profile()
    t0 = time.perf_counter()
    prof.enable()
    with contextlib.redirect_stdout(io.StringIO()):          # [INFO] Colonne: ... lists every column
      extractor.main(["--top", top, "--traj", nc, "--pairs-file", pf, "--out", os.path.join(d, "o.csv"), "--summary-out", os.path.join(d, "s.json")])
    prof.disable()
    wall = time.perf_counter() - t0
    s = io.StringIO(); pstats.Stats(prof, stream=s).sort_stats("cumulative").print_stats(14)
    print(f"run {rep}: wall {wall:.3f} s = {T * P / wall:.3g} evals/s, {os.path.getsize(nc) / wall / 1e9:.2f} GB/s of trajectory")
    if rep == 1:
        print("\n".join(l for l in s.getvalue().splitlines() if "/" in l or "{" in l)[:2500])
