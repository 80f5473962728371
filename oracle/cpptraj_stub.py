#!/usr/bin/env python3
"""Stand-in for the external ``cpptraj`` binary -- TEST INFRASTRUCTURE ONLY (see oracle/README: nothing under
``caviar_b200/`` imports or executes this file; parity unpinned against the real cpptraj, which is not vendored).

The reference's Extractor writes a cpptraj input script and shells out (CAVIAR_Distances-Extractor.py:120-159).
This program accepts EXACTLY the script text those lines generate and nothing else -- so running the unmodified
reference ``main()`` against it pins the script grammar --

    parm "<top>"                                   Extractor:120
    trajin "<traj>" 1 last [<stride>]              Extractor:122-125
    distance dNNNN :<i>@CA :<j>@CA                 Extractor:133-134,143   (--numbering sequential)
    distance dNNNN resid <a>@CA resid <b>@CA       Extractor:137-138,143   (--numbering resSeq, as the reference emits it)
    run                                            Extractor:145
    writedata "<csv>" dNNNN ... delim comma        Extractor:151

evaluates the distances with the repository's fp64 oracle (imaging on whenever the trajectory has a box, as cpptraj's
``distance`` does by default [ext]) and writes what cpptraj's ``writedata ... delim comma`` is documented to write
[ext]: a ``#Frame`` column holding the 1-based set index, one column per data set, fixed 4 decimals, LF line ends.

It has its own minimal topology / trajectory readers (prmtop, PDB; Amber NetCDF through scipy) so that the golden
files it produces do not depend on ``caviar_b200.trajio``.
"""
import os
import re
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import caviar_oracle as orc  # noqa: E402

RE_PARM = re.compile(r'^parm "(.+)"$')
RE_TRAJIN = re.compile(r'^trajin "(.+)" 1 last(?: (\d+))?$')
RE_DIST_SEQ = re.compile(r"^distance (d\d{4}) :(\d+)@CA :(\d+)@CA$")
RE_DIST_RES = re.compile(r"^distance (d\d{4}) resid (\d+)@CA resid (\d+)@CA$")
RE_WRITE = re.compile(r'^writedata "(.+)" ((?:d\d{4} )+)delim comma$')


def die(msg):
    sys.stderr.write(f"cpptraj_stub: {msg}\n")
    sys.exit(1)


def read_topology(path):
    """-> (atom_names, first_atom_of_residue, resseq_or_None)."""
    with open(path, "r", errors="replace") as fh:
        text = fh.read()
    if "%FLAG POINTERS" in text:                                   # Amber prmtop: fixed-width Fortran sections
        sec, name = {}, None
        for line in text.splitlines():
            if line.startswith("%FLAG"):
                name = line.split()[1]
                sec[name] = []
            elif line.startswith("%FORMAT") or line.startswith("%VERSION") or line.startswith("%COMMENT"):
                continue
            elif name:
                sec[name].append(line)
        ptr = " ".join(sec["POINTERS"]).split()
        natom, nres = int(ptr[0]), int(ptr[11])
        names = []
        for line in sec["ATOM_NAME"]:
            names += [line[i:i + 4].strip() for i in range(0, len(line), 4)]
        first = [int(v) - 1 for v in " ".join(sec["RESIDUE_POINTER"]).split()]
        return names[:natom], first[:nres], None
    names, first, seqs, last = [], [], [], None                    # PDB
    for line in text.splitlines():
        if line.startswith(("ATOM", "HETATM")):
            key = (line[21], line[22:27])
            if key != last:
                last = key
                first.append(len(names))
                seqs.append(int(line[22:26]))
            names.append(line[12:16].strip())
        elif line.startswith("ENDMDL"):
            break
    if not names:
        die(f"cannot read topology {path}")
    return names, first, seqs


def ca_of_residue(names, first, r0):
    hi = first[r0 + 1] if r0 + 1 < len(first) else len(names)
    for a in range(first[r0], hi):
        if names[a] == "CA":
            return a
    die(f"no CA atom in residue {r0 + 1}")


def read_netcdf(path):
    from scipy.io import netcdf_file
    nc = netcdf_file(path, "r", mmap=False)
    xyz = np.array(nc.variables["coordinates"][:], dtype=np.float32)
    box = None
    if "cell_lengths" in nc.variables and "cell_angles" in nc.variables:
        lengths = np.array(nc.variables["cell_lengths"][:], dtype=np.float64)
        angles = np.array(nc.variables["cell_angles"][:], dtype=np.float64)
        if len(lengths) and np.all(lengths[0] > 0):
            if np.allclose(angles, 90.0, atol=1e-6):
                box = lengths
            else:
                box = np.stack([orc.box_matrix(lengths[t], angles[t]) for t in range(len(lengths))])
    nc.close()
    return xyz, box


def main(argv):
    if len(argv) != 3 or argv[1] != "-i":
        die("usage: cpptraj -i <script>")
    copy_to = os.environ.get("CPPTRAJ_STUB_SCRIPT_COPY")
    if copy_to:
        shutil.copyfile(argv[2], copy_to)
    with open(argv[2]) as fh:
        lines = fh.read().split("\n")
    if lines[-1] != "":
        die("script does not end with a newline")
    lines = lines[:-1]
    if len(lines) < 5:
        die("script too short")
    m = RE_PARM.match(lines[0]) or die(f"line 1: expected parm, got {lines[0]!r}")
    top_path = m.group(1)
    m = RE_TRAJIN.match(lines[1]) or die(f"line 2: expected trajin, got {lines[1]!r}")
    traj_path, stride = m.group(1), int(m.group(2) or 1)
    names, first, seqs = read_topology(top_path)
    sets, pairs = [], []
    k = 2
    while k < len(lines) and lines[k].startswith("distance "):
        ms, mr = RE_DIST_SEQ.match(lines[k]), RE_DIST_RES.match(lines[k])
        if ms:
            ra, rb = int(ms.group(2)) - 1, int(ms.group(3)) - 1
            if min(ra, rb) < 0 or max(ra, rb) >= len(first):
                die(f"line {k + 1}: residue out of range")
            sets.append(ms.group(1))
        elif mr:
            if seqs is None:
                die("resid masks need a PDB topology")
            try:
                ra, rb = seqs.index(int(mr.group(2))), seqs.index(int(mr.group(3)))
            except ValueError:
                die(f"line {k + 1}: resid not found")
            sets.append(mr.group(1))
        else:
            die(f"line {k + 1}: unrecognised distance command {lines[k]!r}")
        if sets[-1] != f"d{len(sets):04d}":
            die(f"line {k + 1}: data set names must run d0001, d0002, ...")
        pairs.append((ca_of_residue(names, first, ra), ca_of_residue(names, first, rb)))
        k += 1
    if not pairs or k + 2 != len(lines) or lines[k] != "run":
        die("expected: distance ..., run, writedata")
    m = RE_WRITE.match(lines[k + 1]) or die(f"last line: expected writedata, got {lines[k + 1]!r}")
    if m.group(2).split() != sets:
        die("writedata must list every data set in order")
    xyz, box = read_netcdf(traj_path)
    if xyz.shape[1] != len(names):
        die("topology and trajectory disagree on the atom count")
    sel = slice(0, None, stride)
    xyz = xyz[sel]
    box = None if box is None else box[sel]
    d = orc.pbc_pair_distances(xyz, np.asarray(pairs), box) if box is not None else \
        orc.pbc_pair_distances(xyz, np.asarray(pairs), None)
    with open(m.group(1), "w", newline="") as out:
        out.write(",".join(["#Frame"] + sets) + "\n")
        for t in range(d.shape[0]):
            out.write(",".join([str(t + 1)] + ["%.4f" % v for v in d[t]]) + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
