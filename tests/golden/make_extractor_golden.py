#!/usr/bin/env python3
"""Golden CSV files of the Extractor boundary, produced by RUNNING THE REFERENCE'S OWN main().

    python tests/golden/make_extractor_golden.py [out_dir]        (default: tests/golden/extractor)

The unmodified /root/reference/CAVIAR_Distances-Extractor.py is run as a subprocess with a stand-in ``cpptraj`` on
PATH (oracle/cpptraj_stub.py: accepts exactly the script text of Extractor:120-151, evaluates the distances with the
fp64 oracle, writes ``#Frame,dNNNN...`` rows with 4 decimals).  Everything after the subprocess -- delimiter
detection, header rename, the csv.writer that terminates every row with CR LF (Extractor:167-198), the stdout lines
(:203-204) -- is the reference's own code.  Committed per case: the CSV it wrote, the cpptraj script it generated and
its stdout.  The drop-in must reproduce the CSV byte for byte (tests/test_extractor_golden.py, tests/test_gpu_extractor.py).

The input system is small and written here WITHOUT caviar_b200: a hand-formatted Amber prmtop, a PDB with its own
residue numbering, and Amber NetCDF trajectories written by scipy with an UNLIMITED frame dimension (the record
layout real Amber files have: coordinates, cell_lengths and cell_angles of one frame interleaved) -- so they double
as foreign-written fixtures for caviar_b200.trajio.  Frames are chosen so that no distance sits within 1.5e-5 A of a
4-decimal rounding boundary: a float32 result within tolerance then prints the same digits as the fp64 oracle.
"""
import json
import os
import shutil
import stat
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("CAVIAR_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
from oracle import caviar_oracle as orc  # noqa: E402

RES3 = ["ASP", "GLU", "ALA", "GLY", "SER", "THR", "LYS", "ARG", "HIS", "PRO", "LEU", "ILE", "VAL", "PHE", "TYR", "TRP"]
BACKBONE = ["N", "CA", "C", "O"]
N_RES = 36
LABEL_OFFSET = 21                       # CAVIAR's residue_label_offset default (CAVIAR-1.0.py:55)
PDB_FIRST_RESSEQ = 101                  # the PDB numbers its residues 101.. with a gap after the 20th
CELL = (31.0, 33.5, 29.0)

# (name, argv after the script, topology kind) -- argv uses paths relative to the case directory
CASES = [
    ("seq_offset_file", ["--top", "sys.prmtop", "--traj", "traj_ortho.nc", "--pairs-file", "pairs_tic.csv",
                         "--offset", str(LABEL_OFFSET), "--out", "OUT"]),
    ("seq_stride_pairs", ["--top", "sys.prmtop", "--traj", "traj_ortho.nc", "--pairs", "ASP1-SER5,GLY4 - LEU27, ala3-his9,ASP1-SER5",
                          "--stride", "3", "--out", "OUT"]),
    ("resseq_pdb_triclinic", ["--top", "sys.pdb", "--traj", "traj_tric.nc", "--pairs", "ASP101-LYS127,GLU102-VAL133,SER105-PRO110",
                              "--numbering", "resSeq", "--stride", "2", "--out", "OUT"]),
    ("nobox", ["--top", "sys.prmtop", "--traj", "traj_nobox.nc", "--pairs-file", "pairs.txt", "--out", "OUT"]),
]


def residue_names():
    return [RES3[i % len(RES3)] for i in range(N_RES)]


def write_prmtop(path):
    """Minimal Amber prmtop by hand: POINTERS (NATOM first, NRES twelfth), ATOM_NAME 20a4, RESIDUE_LABEL 20a4,
    RESIDUE_POINTER 10I8 -- the sections `:{idx}@CA` masks need."""
    natom = N_RES * len(BACKBONE)
    ptr = [0] * 31
    ptr[0], ptr[11] = natom, N_RES

    def fmt_i(vals):
        return "\n".join("".join("%8d" % v for v in vals[i:i + 10]) for i in range(0, len(vals), 10))

    def fmt_a(vals):
        return "\n".join("".join("%-4s" % v for v in vals[i:i + 20]) for i in range(0, len(vals), 20))

    names = BACKBONE * N_RES
    first = [1 + len(BACKBONE) * r for r in range(N_RES)]
    with open(path, "w") as fh:
        fh.write("%VERSION  VERSION_STAMP = V0001.000  DATE = 10/20/26  12:00:00\n")
        fh.write("%FLAG TITLE\n%FORMAT(20a4)\nextractor golden fixture\n")
        fh.write("%FLAG POINTERS\n%FORMAT(10I8)\n" + fmt_i(ptr) + "\n")
        fh.write("%FLAG ATOM_NAME\n%FORMAT(20a4)\n" + fmt_a(names) + "\n")
        fh.write("%FLAG RESIDUE_LABEL\n%FORMAT(20a4)\n" + fmt_a(residue_names()) + "\n")
        fh.write("%FLAG RESIDUE_POINTER\n%FORMAT(10I8)\n" + fmt_i(first) + "\n")


def pdb_resseq(r):
    return PDB_FIRST_RESSEQ + r + (5 if r >= 20 else 0)


def write_pdb(path, xyz0):
    with open(path, "w") as fh:
        fh.write("REMARK extractor golden fixture\n")
        serial = 1
        for r, res in enumerate(residue_names()):
            for k, name in enumerate(BACKBONE):
                x, y, z = xyz0[len(BACKBONE) * r + k]
                fh.write("ATOM  %5d %-4s %3s A%4d    %8.3f%8.3f%8.3f  1.00  0.00\n" % (serial, " " + name if len(name) < 4 else name,
                                                                                      res, pdb_resseq(r), x, y, z))
                serial += 1
        fh.write("END\n")


def write_netcdf_unlimited(path, xyz, lengths=None, angles=None):
    """Amber NetCDF convention 1.0 as sander / cpptraj write it: NetCDF-3 64-bit offset, UNLIMITED frame dimension,
    record variables time / coordinates / cell_lengths / cell_angles (scipy's writer, not ours)."""
    from scipy.io import netcdf_file
    T, N, _ = xyz.shape
    nc = netcdf_file(path, "w", version=2)
    nc.title = "extractor golden fixture"
    nc.application = "AMBER"
    nc.program = "sander"
    nc.programVersion = "22.0"
    nc.Conventions = "AMBER"
    nc.ConventionVersion = "1.0"
    nc.createDimension("frame", None)
    nc.createDimension("spatial", 3)
    nc.createDimension("atom", N)
    nc.createDimension("label", 5)
    nc.createDimension("cell_spatial", 3)
    nc.createDimension("cell_angular", 3)
    sp = nc.createVariable("spatial", "c", ("spatial",))
    sp[:] = np.array(list("xyz"), dtype="S1")
    tm = nc.createVariable("time", "f", ("frame",))
    tm.units = "picosecond"
    co = nc.createVariable("coordinates", "f", ("frame", "atom", "spatial"))
    co.units = "angstrom"
    if lengths is not None:
        cl = nc.createVariable("cell_lengths", "d", ("frame", "cell_spatial"))
        cl.units = "angstrom"
        ca = nc.createVariable("cell_angles", "d", ("frame", "cell_angular"))
        ca.units = "degree"
    for t in range(T):
        tm[t] = 2.0 * (t + 1)
        co[t] = xyz[t]
        if lengths is not None:
            cl[t] = lengths[t]
            ca[t] = angles[t]
    nc.close()


def ca_index(res0):
    return len(BACKBONE) * res0 + BACKBONE.index("CA")


def build_system(rng):
    """Backbone random walk (3.8 A between consecutive CA), then drifted so that it straddles the cell faces."""
    ca = np.cumsum(3.8 * rng.normal(size=(N_RES, 3)) / np.sqrt(3.0), axis=0) + np.array(CELL) * 0.8
    offs = {"N": (-1.2, 0.4, 0.3), "CA": (0, 0, 0), "C": (1.3, 0.5, -0.2), "O": (1.9, 1.5, -0.3)}
    return np.concatenate([ca[r][None] + np.array([offs[a] for a in BACKBONE]) for r in range(N_RES)], axis=0)


def main():
    out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "extractor")
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.default_rng(20261020)
    base = build_system(rng)
    T = 18
    # every residue pair any case names, as CA atom pairs (margins are checked for all of them, with and without a box)
    all_pairs = np.array([(ca_index(i), ca_index(j)) for i, j in
                          [(0, 4), (3, 26), (2, 8), (0, 21), (1, 27), (4, 9), (24, 34), (3, 10), (5, 30), (7, 19)]])
    scale = 1.0 + 1e-3 * np.sin(0.37 * np.arange(T))
    len_ortho = np.array(CELL)[None, :] * scale[:, None]
    ang_ortho = np.full((T, 3), 90.0)
    len_tric = np.full((T, 3), 34.0) * scale[:, None]
    ang_tric = np.full((T, 3), 109.4712206)
    h_tric = np.stack([orc.box_matrix(len_tric[t], ang_tric[t]) for t in range(T)])

    # frames must be well conditioned for all three imaging settings at once
    frames = []
    tries = 0
    while len(frames) < T:
        tries += 1
        if tries > 500000:
            raise RuntimeError("could not find enough well-conditioned frames")
        t = len(frames)
        x = (base + rng.normal(0, 0.6, size=base.shape)).astype(np.float32)
        ok = True
        for box in (None, len_ortho[t][None], h_tric[t][None]):
            d = orc.pbc_pair_distances(x[None], all_pairs, box)
            if np.abs(d * 1e4 - np.floor(d * 1e4) - 0.5).min() < 0.15:
                ok = False
                break
        if ok:
            frames.append(x)
    xyz = np.stack(frames)

    write_prmtop(os.path.join(out_dir, "sys.prmtop"))
    write_pdb(os.path.join(out_dir, "sys.pdb"), xyz[0])
    write_netcdf_unlimited(os.path.join(out_dir, "traj_ortho.nc"), xyz, len_ortho, ang_ortho)
    write_netcdf_unlimited(os.path.join(out_dir, "traj_tric.nc"), xyz, len_tric, ang_tric)
    write_netcdf_unlimited(os.path.join(out_dir, "traj_nobox.nc"), xyz)
    names = residue_names()

    def label(r, off):
        return f"{names[r]}{r + 1 + off}"
    # CAVIAR's own TIC csv (save_tic_csv, CAVIAR-1.0.py:141-147) as the pairs file of case 1: labels carry the offset
    with open(os.path.join(out_dir, "pairs_tic.csv"), "w") as fh:
        fh.write("rank,pair,res1,res2,abs_loading\n")
        for k, (i, j) in enumerate([(0, 21), (1, 27), (4, 9), (24, 34), (0, 21)], start=1):
            a, b = label(i, LABEL_OFFSET), label(j, LABEL_OFFSET)
            fh.write(f"{k},{a}-{b},{a},{b},{0.9 / k:.6f}\n")
    with open(os.path.join(out_dir, "pairs.txt"), "w") as fh:
        fh.write("# free-form list\n" + f"{label(3, 0)}-{label(10, 0)}\n" + f"dist 2: {label(5, 0)} - {label(30, 0)}\n\n"
                 + f"{label(7, 0)}-{label(19, 0)}\n")

    # the stand-in cpptraj on PATH
    bin_dir = tempfile.mkdtemp(prefix="cpptraj_stub_")
    stub = os.path.join(bin_dir, "cpptraj")
    with open(stub, "w") as fh:
        fh.write(f"#!/bin/sh\nexec {sys.executable} {os.path.join(ROOT, 'oracle', 'cpptraj_stub.py')} \"$@\"\n")
    os.chmod(stub, os.stat(stub).st_mode | stat.S_IXUSR | stat.S_IXGRP | stat.S_IXOTH)
    env = dict(os.environ, PATH=bin_dir + os.pathsep + os.environ.get("PATH", ""))
    cases = []
    try:
        for name, argv in CASES:
            out_csv = f"{name}.csv"
            args = [out_csv if a == "OUT" else a for a in argv]
            env["CPPTRAJ_STUB_SCRIPT_COPY"] = os.path.join(out_dir, f"{name}.cpptraj.in")
            res = subprocess.run([sys.executable, os.path.join(REF, "CAVIAR_Distances-Extractor.py")] + args, cwd=out_dir, env=env,
                                 stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            if res.returncode != 0:
                raise RuntimeError(f"reference main() failed for {name}: {res.stdout}\n{res.stderr}")
            with open(os.path.join(out_dir, f"{name}.stdout"), "w") as fh:
                fh.write(res.stdout)
            cases.append({"name": name, "argv": args, "csv": out_csv, "script": f"{name}.cpptraj.in", "stdout": f"{name}.stdout"})
    finally:
        shutil.rmtree(bin_dir, ignore_errors=True)
    with open(os.path.join(out_dir, "cases.json"), "w") as fh:
        json.dump({"cases": cases, "frames": T, "n_res": N_RES, "label_offset": LABEL_OFFSET,
                   "reference": "CAVIAR_Distances-Extractor.py main(), unmodified, cpptraj = oracle/cpptraj_stub.py"}, fh, indent=1)
    print("wrote", out_dir, [c["name"] for c in cases])


if __name__ == "__main__":
    main()
