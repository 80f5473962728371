"""Trajectory / topology ingest for the Extractor drop-in.

The reference hands ``--top`` / ``--traj`` to the external cpptraj binary (CAVIAR_Distances-Extractor.py:120-125).
Here the two formats that workflow uses are read directly:

* Amber NetCDF trajectories (``.nc``; NetCDF-3, 64-bit offset): ``coordinates (frame, atom, 3)`` float32 Angstrom,
  ``cell_lengths`` / ``cell_angles (frame, 3)`` float64 -- through ``scipy.io.netcdf_file(mmap=True)``, so frame
  blocks are paged in on demand and never copied as a whole;
* Amber prmtop topologies (``%FLAG ATOM_NAME / RESIDUE_LABEL / RESIDUE_POINTER``), PDB files and CHARMM / NAMD PSF
  files, only as far as the ``:{idx}@CA`` / ``resid {n}@CA`` selections of the reference need (Extractor:127-138).

* CHARMM / NAMD DCD trajectories (``.dcd``): Fortran-record file, per frame an optional unit-cell record (6 doubles)
  and three float32 arrays X[N], Y[N], Z[N]; either byte order; memory-mapped, the selected atoms are gathered
  natively.  Files with fixed atoms (NAMNF > 0) or 4-D coordinates are refused.

* GROMACS TRR trajectories (``.trr``; XDR, big-endian, single precision, constant frame size): coordinates and box
  in nm are converted to Angstrom (x10) after the gather.  ``.gro`` files serve as topologies.

A ``.npy`` array ``(T, N, 3)`` is accepted as a trajectory too (no box), which is what the tests and benchmarks use
when no NetCDF file is at hand.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np


# ---------------------------------------------------------------------------------------------------------
# topology
# ---------------------------------------------------------------------------------------------------------
@dataclass
class Topology:
    atom_names: List[str]
    res_labels: List[str]           # per residue
    res_first_atom: List[int]       # 0-based index of each residue's first atom
    res_seq: Optional[List[int]]    # PDB resSeq per residue (None for prmtop)

    @property
    def n_atoms(self) -> int:
        return len(self.atom_names)

    @property
    def n_res(self) -> int:
        return len(self.res_labels)

    def atom_of(self, res_index0: int, atom_name: str = "CA") -> int:
        """0-based atom index of ``atom_name`` in residue ``res_index0`` (0-based); -1 when the residue has none."""
        lo = self.res_first_atom[res_index0]
        hi = self.res_first_atom[res_index0 + 1] if res_index0 + 1 < self.n_res else self.n_atoms
        for a in range(lo, hi):
            if self.atom_names[a] == atom_name:
                return a
        return -1

    def residue_by_resseq(self, resseq: int) -> int:
        if self.res_seq is None:
            raise ValueError("this topology has no PDB residue numbers (resSeq needs a PDB topology)")
        try:
            return self.res_seq.index(resseq)
        except ValueError:
            return -1


def _prmtop_sections(path: str) -> Dict[str, List[str]]:
    sections: Dict[str, List[str]] = {}
    name, fmt = None, None
    with open(path, "r", errors="ignore") as fh:
        for line in fh:
            if line.startswith("%FLAG"):
                name = line.split()[1]
                sections[name] = []
                fmt = None
            elif line.startswith("%FORMAT"):
                fmt = line.strip()[8:-1]
                sections.setdefault(name, []).append("%" + fmt)
            elif name is not None and not line.startswith("%"):
                sections[name].append(line.rstrip("\n"))
    return sections


def _fixed_width(lines: List[str]) -> List[str]:
    """Split the data lines of one prmtop section according to its Fortran format (20a4, 10I8, 5E16.8)."""
    fmt = lines[0][1:].lower()
    import re
    m = re.match(r"(\d+)([aie])(\d+)", fmt)
    width = int(m.group(3)) if m else 4
    out = []
    for ln in lines[1:]:
        out += [ln[i:i + width] for i in range(0, len(ln), width)]
    return out


def read_prmtop(path: str) -> Topology:
    sec = _prmtop_sections(path)
    for need in ("POINTERS", "ATOM_NAME", "RESIDUE_LABEL", "RESIDUE_POINTER"):
        if need not in sec:
            raise ValueError(f"{path}: not an Amber prmtop (missing %FLAG {need})")
    ptr = [int(x) for x in _fixed_width(sec["POINTERS"]) if x.strip()]
    natom, nres = ptr[0], ptr[11]
    names = [x.strip() for x in _fixed_width(sec["ATOM_NAME"])][:natom]
    labels = [x.strip() for x in _fixed_width(sec["RESIDUE_LABEL"])][:nres]
    first = [int(x) - 1 for x in _fixed_width(sec["RESIDUE_POINTER"]) if x.strip()][:nres]
    if len(names) != natom or len(labels) != nres or len(first) != nres:
        raise ValueError(f"{path}: truncated prmtop")
    return Topology(names, labels, first, None)


def read_pdb(path: str) -> Topology:
    names, labels, first, seqs = [], [], [], []
    last = None
    with open(path, "r", errors="ignore") as fh:
        for line in fh:
            if line.startswith("ENDMDL"):
                break                                               # a multi-MODEL file: the first model is the topology
            if not (line.startswith("ATOM") or line.startswith("HETATM")):
                continue
            key = (line[21], line[22:27])
            if key != last:
                last = key
                first.append(len(names))
                labels.append(line[17:20].strip())
                seqs.append(int(line[22:26]))
            names.append(line[12:16].strip())
    if not names:
        raise ValueError(f"{path}: no ATOM records")
    return Topology(names, labels, first, seqs)


def read_psf(path: str) -> Topology:
    """CHARMM / NAMD / X-PLOR PSF (what usually accompanies a DCD): the ``!NATOM`` section, whitespace separated
    ``atomid segid resid resname atomname ...``; a residue changes when (segid, resid) does."""
    names, labels, first, seqs = [], [], [], []
    with open(path, "r", errors="ignore") as fh:
        lines = iter(fh)
        n_atom = None
        for line in lines:
            if "!NATOM" in line:
                n_atom = int(line.split()[0])
                break
        if n_atom is None:
            raise ValueError(f"{path}: no !NATOM section (not a PSF file)")
        last = None
        for _ in range(n_atom):
            f = next(lines).split()
            if len(f) < 5:
                raise ValueError(f"{path}: short atom record")
            key = (f[1], f[2])
            if key != last:
                last = key
                first.append(len(names))
                labels.append(f[3][:3] if len(f[3]) > 3 else f[3])
                digits = "".join(ch for ch in f[2] if ch.isdigit() or ch == "-")     # resid may carry an insertion code
                seqs.append(int(digits) if digits not in ("", "-") else len(first))
            names.append(f[4])
    return Topology(names, labels, first, seqs)


def read_gro(path: str) -> Topology:
    """GROMACS .gro (fixed columns: resid[0:5] resname[5:10] atomname[10:15] atomnr[15:20] ...): first frame only."""
    names, labels, first, seqs = [], [], [], []
    with open(path, "r", errors="ignore") as fh:
        fh.readline()
        n_atom = int(fh.readline().split()[0])
        last = None
        for _ in range(n_atom):
            line = fh.readline()
            if len(line) < 20:
                raise ValueError(f"{path}: short atom line")
            key = (line[0:5], line[5:10])
            if key != last:
                last = key
                first.append(len(names))
                labels.append(line[5:10].strip()[:3])
                seqs.append(int(line[0:5]))
            names.append(line[10:15].strip())
    return Topology(names, labels, first, seqs)


def read_topology(path: str) -> Topology:
    ext = os.path.splitext(path)[1].lower()
    if ext in (".pdb", ".ent"):
        return read_pdb(path)
    if ext == ".psf":
        return read_psf(path)
    if ext == ".gro":
        return read_gro(path)
    return read_prmtop(path)


def write_prmtop(path: str, atom_names: List[str], res_labels: List[str], res_first_atom: List[int]) -> None:
    """Minimal prmtop with the four sections read above (for tests and synthetic fixtures)."""
    def block(flag, fmt, items, per, width, conv):
        out = [f"%FLAG {flag}", f"%FORMAT({fmt})"]
        for i in range(0, max(len(items), 1), per):
            out.append("".join(conv(x, width) for x in items[i:i + per]))
        return out
    ptr = [0] * 31
    ptr[0], ptr[11] = len(atom_names), len(res_labels)
    lines = ["%VERSION  VERSION_STAMP = V0001.000  DATE = 01/01/26  00:00:00"]
    lines += block("TITLE", "20a4", ["caviar_b200 synthetic"], 1, 80, lambda x, w: x.ljust(w))
    lines += block("POINTERS", "10I8", ptr, 10, 8, lambda x, w: str(x).rjust(w))
    lines += block("ATOM_NAME", "20a4", atom_names, 20, 4, lambda x, w: x.ljust(w)[:w])
    lines += block("RESIDUE_LABEL", "20a4", res_labels, 20, 4, lambda x, w: x.ljust(w)[:w])
    lines += block("RESIDUE_POINTER", "10I8", [i + 1 for i in res_first_atom], 10, 8, lambda x, w: str(x).rjust(w))
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")


# ---------------------------------------------------------------------------------------------------------
# cell
# ---------------------------------------------------------------------------------------------------------
def cell_to_matrix(lengths: np.ndarray, angles_deg: np.ndarray) -> np.ndarray:
    """(T,3) lengths + (T,3) angles [alpha, beta, gamma] -> (T,3,3), rows = lattice vectors, lower triangular
    (a along x, b in the xy plane): the Amber / cpptraj setting."""
    L = np.asarray(lengths, dtype=np.float64).reshape(-1, 3)
    ang = np.deg2rad(np.asarray(angles_deg, dtype=np.float64).reshape(-1, 3))
    ca, cb, cg, sg = np.cos(ang[:, 0]), np.cos(ang[:, 1]), np.cos(ang[:, 2]), np.sin(ang[:, 2])
    h = np.zeros((L.shape[0], 3, 3))
    h[:, 0, 0] = L[:, 0]
    h[:, 1, 0] = L[:, 1] * cg
    h[:, 1, 1] = L[:, 1] * sg
    h[:, 2, 0] = L[:, 2] * cb
    h[:, 2, 1] = L[:, 2] * (ca - cb * cg) / sg
    h[:, 2, 2] = np.sqrt(np.maximum(L[:, 2] ** 2 - h[:, 2, 0] ** 2 - h[:, 2, 1] ** 2, 0.0))
    h[np.abs(h) < 1e-12 * np.abs(h).max()] = 0.0
    return h


# ---------------------------------------------------------------------------------------------------------
# trajectories
# ---------------------------------------------------------------------------------------------------------
class Trajectory:
    """Random access to frame blocks: ``coords(lo, hi, stride)`` -> float32 (n, N, 3); ``cell(...)`` -> box or None."""

    def __init__(self, path: str, n_atoms: Optional[int] = None):
        """``n_atoms`` is only needed for formats that do not record it (Amber ASCII mdcrd): pass the topology's."""
        self.path = path
        ext = os.path.splitext(path)[1].lower()
        self._nc = None
        self._dcd = None
        self._trr = None
        self._crd = None
        self._xtc = None
        if ext == ".xtc":
            self._xtc = _XtcFile(path)
            self._xyz = None
            self._len = self._ang = None
            self.n_frames, self.n_atoms = self._xtc.n_frames, self._xtc.n_atoms
            return
        if ext in (".mdcrd", ".crd", ".trj"):
            if not n_atoms:
                sys.exit("[ERR] una traiettoria Amber ASCII (mdcrd) non contiene il numero di atomi: serve la topologia")
            self._crd = _MdcrdFile(path, int(n_atoms))
            self._xyz = None
            self._len = self._ang = None
            self.n_frames, self.n_atoms = self._crd.n_frames, self._crd.n_atoms
            return
        if ext == ".trr":
            self._trr = _TrrFile(path)
            self._xyz = None
            self._len = self._ang = None
            self.n_frames, self.n_atoms = self._trr.n_frames, self._trr.n_atoms
            return
        if ext == ".dcd":
            self._dcd = _DcdFile(path)
            self._xyz = None
            self._len = self._ang = None
            self.n_frames, self.n_atoms = self._dcd.n_frames, self._dcd.n_atoms
            return
        if ext == ".npy":
            self._xyz = np.load(path, mmap_mode="r")
            if self._xyz.ndim != 3 or self._xyz.shape[2] != 3:
                raise ValueError(f"{path}: expected a (T, N, 3) array")
            self._len = self._ang = None
        else:
            # everything else must be a NetCDF-3 file (Amber .nc / .ncdf): look at the magic bytes before handing the
            # file to scipy, whose errors on foreign formats are opaque
            with open(path, "rb") as fh:
                magic = fh.read(8)
            if magic[:3] != b"CDF" or magic[3:4] not in (b"\x01", b"\x02"):
                if magic[:8] == b"\x89HDF\r\n\x1a\n":
                    what = "NetCDF-4 / HDF5 (convert with `cpptraj` or `ncks -3` to NetCDF-3, the Amber convention)"
                else:
                    what = f"unknown format '{ext or path}'"
                sys.exit(f"[ERR] formato di traiettoria non supportato: {what}. Supportati: Amber NetCDF-3 (.nc), mdcrd, DCD, TRR, XTC, .npy")
            from scipy.io import netcdf_file
            self._nc = netcdf_file(path, "r", mmap=True)
            if "coordinates" not in self._nc.variables:
                raise ValueError(f"{path}: no 'coordinates' variable (not an Amber NetCDF trajectory)")
            self._xyz = self._nc.variables["coordinates"]
            self._len = self._nc.variables.get("cell_lengths")
            self._ang = self._nc.variables.get("cell_angles")
        self.n_frames, self.n_atoms = int(self._xyz.shape[0]), int(self._xyz.shape[1])

    @property
    def has_box(self) -> bool:
        if self._crd is not None:
            return self._crd.has_box
        if self._xtc is not None:
            return self.n_frames > 0 and bool(np.all(np.diag(self._xtc.box(0, 1)[0]) > 0))
        if self._trr is not None:
            return self._trr.has_box and self.n_frames > 0 and bool(np.all(np.diag(self._trr.box(0, 1)[0]) > 0))
        if self._dcd is not None:
            return self._dcd.has_cell and self.n_frames > 0 and bool(np.all(self._dcd.cell(0, 1)[0][0] > 0))
        if self._len is None or self._ang is None or self.n_frames == 0:
            return False
        return bool(np.all(np.asarray(self._len[0]) > 0))

    def box_kind(self) -> Optional[str]:
        if not self.has_box:
            return None
        if self._crd is not None:
            return "ortho"                                        # an mdcrd box line holds the three edge lengths only
        if self._trr is not None or self._xtc is not None:
            h = (self._trr or self._xtc).box(0, 1)[0]
            off = np.abs(h - np.diag(np.diag(h))).max()
            return "ortho" if off <= 1e-6 * np.abs(np.diag(h)).max() else "triclinic"
        ang = self._dcd.cell(0, 1)[1][0] if self._dcd is not None else np.asarray(self._ang[0], dtype=np.float64)
        return "ortho" if np.allclose(ang, 90.0, atol=1e-6) else "triclinic"

    def coords(self, lo: int, hi: int, stride: int = 1) -> np.ndarray:
        if self._crd is not None:
            return self._crd.coords(lo, hi, stride)
        if self._dcd is not None or self._trr is not None or self._xtc is not None:
            return self.coords_of(np.arange(self.n_atoms), lo, hi, stride)
        return np.ascontiguousarray(self._xyz[lo:hi:stride], dtype=np.float32)

    def coords_of(self, atoms: np.ndarray, lo: int, hi: int, stride: int = 1, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Frames ``lo:hi:stride`` of the selected atoms only, native float32 (n, len(atoms), 3): one multi-threaded
        native pass over the memory-mapped file (byte order conversion included) instead of numpy's byte-swapping
        copy of whole frames plus a fancy-index pass."""
        from . import _lib
        atoms = np.ascontiguousarray(atoms, dtype=np.int32)
        stride = max(1, int(stride))
        n = len(range(lo, hi, stride))
        if out is None:
            out = np.empty((n, len(atoms), 3), dtype=np.float32)
        if self._crd is not None:                                  # text, fixed-width fields: only the selected atoms are parsed
            sel = np.asarray(atoms)
            if n and len(sel) and (sel.min() < 0 or sel.max() >= self.n_atoms):
                raise IndexError(f"atom index out of range for {self.n_atoms} atoms")
            if out.dtype == np.float32 and out.flags.c_contiguous and out.shape == (n, len(sel), 3):
                self._crd.read(np.arange(lo, hi, stride), sel, out)
            else:
                out[...] = self._crd.coords(lo, hi, stride)[:, sel, :]
            return out
        if self._xtc is not None:                                  # decompressed frame by frame on the host threads, in Angstrom
            if n and len(atoms):
                self._xtc.read(np.arange(lo, hi, stride), atoms, out)
            return out
        if self._trr is not None:
            if n and len(atoms):
                r = self._trr
                _lib.check(_lib.lib.caviar_compact_frames_host(r.x_ptr(lo), n, r.frame_bytes * stride, r.n_atoms,
                                                               atoms.ctypes.data, len(atoms), 1 if r.big else 0, out.ctypes.data))
                out *= np.float32(10.0)                              # nm -> Angstrom
            return out
        if self._dcd is not None:
            if n and len(atoms):
                d = self._dcd
                _lib.check(_lib.lib.caviar_compact_frames_soa_host(
                    d.frame_ptr(lo), n, d.frame_bytes * stride, d.n_atoms, d.x_off, d.y_off, d.z_off,
                    atoms.ctypes.data, len(atoms), 1 if d.swap else 0, out.ctypes.data))
            return out
        arr = self._xyz if isinstance(self._xyz, np.ndarray) else getattr(self._xyz, "data", None)
        ok = (isinstance(arr, np.ndarray) and arr.ndim == 3 and arr.dtype.kind == "f" and arr.dtype.itemsize == 4
              and arr.strides[2] == 4 and arr.strides[1] == 12 and arr.strides[0] >= 12 * self.n_atoms)
        if not ok or n == 0 or len(atoms) == 0:                    # exotic layouts: numpy does it
            out[...] = np.asarray(self._xyz[lo:hi:stride], dtype=np.float32)[:, atoms, :]
            return out
        big = arr.dtype.byteorder == ">" or (arr.dtype.byteorder == "=" and sys.byteorder == "big")
        first = arr[lo:lo + 1]
        _lib.check(_lib.lib.caviar_compact_frames_host(first.ctypes.data, n, arr.strides[0] * stride, self.n_atoms,
                                                       atoms.ctypes.data, len(atoms), 1 if big else 0, out.ctypes.data))
        return out

    def cell(self, lo: int, hi: int, stride: int = 1):
        kind = self.box_kind()
        if kind is None:
            return None
        if self._crd is not None:
            return self._crd.box(lo, hi, stride)
        if self._trr is not None or self._xtc is not None:
            h = (self._trr or self._xtc).box(lo, hi, stride) * 10.0  # nm -> Angstrom; rows = box vectors (a along x)
            return np.stack([h[:, 0, 0], h[:, 1, 1], h[:, 2, 2]], axis=1).astype(np.float32) if kind == "ortho" \
                else h.astype(np.float32)
        if self._dcd is not None:
            L, A = self._dcd.cell(lo, hi, stride)
        else:
            L = np.asarray(self._len[lo:hi:stride], dtype=np.float64)
            A = np.asarray(self._ang[lo:hi:stride], dtype=np.float64)
        if kind == "ortho":
            return L.astype(np.float32)
        return cell_to_matrix(L, A).astype(np.float32)

    def close(self):
        if self._crd is not None:
            self._crd.close()
            self._crd = None
        if self._trr is not None:
            self._trr.close()
            self._trr = None
        if self._xtc is not None:
            self._xtc.close()
            self._xtc = None
        if self._dcd is not None:
            self._dcd.close()
            self._dcd = None
        if self._nc is not None:
            self._xyz = self._len = self._ang = None
            try:
                self._nc.close()
            except Exception:
                pass
            self._nc = None



class _XtcFile:
    """GROMACS XTC (XDR header + compressed integer coordinates per frame) over a read-only memory map.  The frames are
    indexed once (they have different sizes) and decompressed by the native library (``cv_xtc.cu``), in parallel, only
    the selected atoms coming back.  Precision is whatever the writer chose (0.001 nm by default): the positions in the
    file already ARE rounded, nothing is lost or added here."""

    def __init__(self, path: str):
        from . import _lib
        self._mm = np.memmap(path, dtype=np.uint8, mode="r")
        n_atoms = C.c_int32(0)
        n = _lib.lib.caviar_xtc_index(self._mm.ctypes.data, self._mm.size, None, 0, C.byref(n_atoms))
        if n < 0:
            why = (_lib.lib.caviar_last_error() or b"").decode("utf-8", "replace")
            sys.exit(f"[ERR] traiettoria XTC non leggibile ({path}): {why}")
        self.offsets = np.zeros(max(int(n), 1), dtype=np.int64)
        if n > 0:
            got = _lib.lib.caviar_xtc_index(self._mm.ctypes.data, self._mm.size, self.offsets.ctypes.data, int(n), C.byref(n_atoms))
            assert got == n
        self.n_frames, self.n_atoms = int(n), int(n_atoms.value)

    def read(self, frames: np.ndarray, atoms: np.ndarray, out: np.ndarray) -> None:
        from . import _lib
        off = np.ascontiguousarray(self.offsets[np.asarray(frames, dtype=np.int64)])
        _lib.check(_lib.lib.caviar_xtc_decode_host(self._mm.ctypes.data, self._mm.size, off.ctypes.data, len(off),
                                                   atoms.ctypes.data, len(atoms), 10.0, out.ctypes.data, None))

    def box(self, lo: int, hi: int, stride: int = 1) -> np.ndarray:
        """(n, 3, 3) float64, rows = box vectors, in nm (like ``_TrrFile.box``)."""
        idx = np.arange(lo, hi, max(1, stride))
        out = np.empty((len(idx), 3, 3), dtype=np.float64)
        for k, t in enumerate(idx):
            o = int(self.offsets[t]) + 16
            out[k] = self._mm[o:o + 36].view(">f4").reshape(3, 3)
        return out

    def close(self):
        self._mm = None


_XTC_MAGIC = [0] * 9 + [8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512, 645, 812, 1024, 1290,
                        1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007, 32768, 41285, 52015,
                        65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561, 832255, 1048576,
                        1321122, 1664510, 2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607, 10568983, 13316085,
                        16777216]


class _BitWriter:
    def __init__(self):
        self.acc, self.n, self.out = 0, 0, bytearray()

    def put(self, nbits: int, value: int):
        if nbits <= 0:
            return
        self.acc = (self.acc << nbits) | (int(value) & ((1 << nbits) - 1))
        self.n += nbits
        while self.n >= 8:
            self.n -= 8
            self.out.append((self.acc >> self.n) & 0xFF)
        self.acc &= (1 << self.n) - 1

    def put3(self, nbits: int, sizes, nums):
        """Three integers as one mixed-radix number, 8-bit groups, least significant group first."""
        v = (int(nums[0]) * int(sizes[1]) + int(nums[1])) * int(sizes[2]) + int(nums[2])
        while nbits > 8:
            self.put(8, v & 0xFF)
            v >>= 8
            nbits -= 8
        self.put(nbits, v)

    def bytes(self) -> bytes:
        tail = bytes([(self.acc << (8 - self.n)) & 0xFF]) if self.n else b""
        return bytes(self.out) + tail


def write_xtc(path: str, xyz_nm: np.ndarray, box_nm=None, precision: float = 1000.0, magic: int = 1995) -> np.ndarray:
    """Test / tooling writer: ``xyz_nm`` (T, N, 3) in nm, compressed the way GROMACS' xdrfile does it -- integers
    ``round(x * precision)``, one mixed-radix number per atom over the frame's bounding box, runs of atoms close to their
    predecessor as small offsets in the adaptive radix, the first two atoms of a run swapped -- or, for N <= 9, as plain
    floats.  Returns the positions a reader must reproduce (T, N, 3) in nm (= the rounded integers / precision)."""
    import struct
    xyz_nm = np.asarray(xyz_nm, dtype=np.float32)
    T, N = xyz_nm.shape[:2]
    want = np.empty_like(xyz_nm)
    first, last = 9, len(_XTC_MAGIC)
    with open(path, "wb") as fh:
        for t in range(T):
            box = np.zeros((3, 3), np.float32) if box_nm is None else np.asarray(box_nm, np.float32).reshape(-1, 3, 3)[min(t, len(np.asarray(box_nm).reshape(-1, 3, 3)) - 1)]
            fh.write(struct.pack(">iiif", magic, N, t, float(t)) + box.astype(">f4").tobytes() + struct.pack(">i", N))
            if N <= 9:
                fh.write(xyz_nm[t].astype(">f4").tobytes())
                want[t] = xyz_nm[t]
                continue
            ints = np.rint(xyz_nm[t].astype(np.float64) * precision).astype(np.int64)
            want[t] = (ints.astype(np.float32) * np.float32(1.0 / np.float32(precision)))
            lo, hi = ints.min(axis=0), ints.max(axis=0)
            size = (hi - lo + 1).astype(np.int64)
            if int(size.max()) > 0xFFFFFF:
                bits3, bitsize = [int(v).bit_length() for v in size], 0
            else:
                bits3, bitsize = None, int(int(size[0]) * int(size[1]) * int(size[2])).bit_length()
            diffs = np.abs(np.diff(ints, axis=0)).sum(axis=1)
            mindiff = int(diffs.min()) if len(diffs) else 0
            smallidx = first
            while smallidx < last - 1 and _XTC_MAGIC[smallidx] < mindiff:
                smallidx += 1
            maxidx = min(last - 1, smallidx + 8)
            minidx = maxidx - 8
            smaller = _XTC_MAGIC[max(first, smallidx - 1)] // 2
            smallnum = _XTC_MAGIC[smallidx] // 2
            larger = _XTC_MAGIC[maxidx] // 2
            head_smallidx = smallidx
            bw = _BitWriter()
            c = ints.copy()
            i, prevrun = 0, -1
            prev = np.zeros(3, np.int64)
            while i < N:
                is_small = 0
                this = c[i]
                if smallidx < maxidx and i >= 1 and np.all(np.abs(this - prev) < larger):
                    is_smaller = 1
                elif smallidx > minidx:
                    is_smaller = -1
                else:
                    is_smaller = 0
                if i + 1 < N and np.all(np.abs(c[i] - c[i + 1]) < smallnum):
                    c[[i, i + 1]] = c[[i + 1, i]]              # the second atom is coded first (water: O between its H)
                    this = c[i]
                    is_small = 1
                rel = this - lo
                if bitsize == 0:
                    for k in range(3):
                        bw.put(bits3[k], rel[k])
                else:
                    bw.put3(bitsize, size, rel)
                prev = this.copy()
                i += 1
                run = []
                if is_small == 0 and is_smaller == -1:
                    is_smaller = 0
                while is_small and len(run) < 8 * 3:
                    this = c[i]
                    if is_smaller == -1 and int(((this - prev) ** 2).sum()) >= smaller * smaller:
                        is_smaller = 0
                    run.extend((this - prev + smallnum).tolist())
                    prev = this.copy()
                    i += 1
                    is_small = 1 if i < N and np.all(np.abs(c[i] - prev) < smallnum) else 0
                if len(run) != prevrun or is_smaller != 0:
                    prevrun = len(run)
                    bw.put(1, 1)
                    bw.put(5, len(run) + is_smaller + 1)
                else:
                    bw.put(1, 0)
                sizes = [_XTC_MAGIC[smallidx]] * 3
                for k in range(0, len(run), 3):
                    bw.put3(smallidx, sizes, run[k:k + 3])
                if is_smaller != 0:
                    smallidx += is_smaller
                    if is_smaller < 0:
                        smallnum = smaller
                        smaller = _XTC_MAGIC[smallidx - 1] // 2
                    else:
                        smaller = smallnum
                        smallnum = _XTC_MAGIC[smallidx] // 2
            data = bw.bytes()
            fh.write(struct.pack(">f", precision) + struct.pack(">iii", *[int(v) for v in lo]) + struct.pack(">iii", *[int(v) for v in hi])
                     + struct.pack(">i", head_smallidx))
            fh.write(struct.pack(">q", len(data)) if magic == 2023 else struct.pack(">i", len(data)))
            fh.write(data + b"\0" * ((-len(data)) % 4))
    return want

class _MdcrdFile:
    """Amber ASCII trajectory (mdcrd): a title line, then per frame 3N coordinates in Fortran 10F8.3 (eight characters
    per value, ten per line, no separator guaranteed) and, for periodic runs, one more line with the three box lengths.
    The file records neither N nor whether it has box lines: N comes from the topology, the box is inferred from the
    file size (every frame has the same number of bytes).  Frames are parsed on demand from a read-only memory map."""

    def __init__(self, path: str, n_atoms: int):
        self._mm = np.memmap(path, dtype=np.uint8, mode="r")
        mm = self._mm
        self.n_atoms = n_atoms
        nl = int(np.flatnonzero(mm[:4096] == 10)[0]) if (mm[:4096] == 10).any() else -1
        if nl < 0:
            raise ValueError(f"{path}: no title line (not an Amber mdcrd file)")
        self.first = nl + 1
        n_val = 3 * n_atoms
        lines = (n_val + 9) // 10
        eol = 2 if nl > 0 and mm[nl - 1] == 13 else 1                 # CR LF files exist
        self.eol = eol
        self.coord_bytes = 8 * n_val + eol * lines
        body = mm.size - self.first
        box_bytes = 24 + eol
        if body % self.coord_bytes == 0 and body % (self.coord_bytes + box_bytes) != 0:
            self.has_box = False
        elif body % (self.coord_bytes + box_bytes) == 0:
            # both divide only for contrived sizes: then look whether the line after the first frame is a 3-value line
            self.has_box = body % self.coord_bytes != 0 or self._line_len(self.first + self.coord_bytes) == 24
        else:
            raise ValueError(f"{path}: file size does not match {n_atoms} atoms (with or without box lines)")
        self.frame_bytes = self.coord_bytes + (box_bytes if self.has_box else 0)
        self.n_frames = body // self.frame_bytes
        # positions of the value characters inside one frame's coordinate block (newlines skipped)
        k = np.arange(8 * n_val, dtype=np.int64)
        self._pos = k + (k // 80) * eol

    def _line_len(self, pos: int) -> int:
        seg = self._mm[pos:pos + 200]
        nl = np.flatnonzero(seg == 10)
        return int(nl[0]) - (self.eol - 1) if len(nl) else -1

    def _numbers(self, raw: np.ndarray) -> np.ndarray:
        return np.char.strip(raw.view("S8")).astype(np.float64)

    def read(self, frames: np.ndarray, atoms: Optional[np.ndarray], out: np.ndarray) -> None:
        """Frames ``frames`` (indices), atoms ``atoms`` (None: all) -> ``out`` (n, n_sel, 3) float32: the fixed-width
        fields of the selected atoms only, parsed by the native library on the host threads (``cv_mdcrd.cu``)."""
        from . import _lib
        frames = np.ascontiguousarray(frames, dtype=np.int64)
        if atoms is not None:
            atoms = np.ascontiguousarray(atoms, dtype=np.int32)
        if len(frames) == 0 or (atoms is not None and len(atoms) == 0):
            return
        try:
            _lib.check(_lib.lib.caviar_mdcrd_decode_host(self._mm.ctypes.data, self._mm.size, self.first, self.frame_bytes, self.eol,
                                                         self.n_atoms, frames.ctypes.data, len(frames),
                                                         None if atoms is None else atoms.ctypes.data,
                                                         0 if atoms is None else len(atoms), out.ctypes.data))
        except _lib.CaviarError as e:                              # what the numpy string conversion raised
            raise ValueError(str(e)) from None

    def coords(self, lo: int, hi: int, stride: int = 1) -> np.ndarray:
        idx = np.arange(lo, hi, max(1, stride))
        out = np.empty((len(idx), self.n_atoms, 3), dtype=np.float32)
        self.read(idx, None, out)
        return out

    def coords_numpy(self, lo: int, hi: int, stride: int = 1) -> np.ndarray:
        """The same through numpy's string -> float64 conversion (single-threaded, every atom): kept as the checker of
        the native parser in the tests."""
        idx = np.arange(lo, hi, max(1, stride))
        out = np.empty((len(idx), self.n_atoms, 3), dtype=np.float32)
        for k, t in enumerate(idx):
            o = self.first + int(t) * self.frame_bytes
            block = np.asarray(self._mm[o:o + self.coord_bytes])
            out[k] = self._numbers(np.ascontiguousarray(block[self._pos])).reshape(self.n_atoms, 3)
        return out

    def box(self, lo: int, hi: int, stride: int = 1) -> np.ndarray:
        idx = np.arange(lo, hi, max(1, stride), dtype=np.int64)
        if len(idx) == 0:
            return np.empty((0, 3), dtype=np.float32)
        pos = (self.first + idx * self.frame_bytes + self.coord_bytes)[:, None] + np.arange(24, dtype=np.int64)[None, :]
        return self._numbers(np.ascontiguousarray(self._mm[pos.ravel()])).reshape(len(idx), 3).astype(np.float32)   # all box lines at once

    def close(self):
        self._mm = None


class _TrrFile:
    """Layout of a GROMACS TRR file (XDR: big-endian) over a read-only memory map; single precision, every frame of
    the same size (the common case: positions, optionally velocities / forces, written at one interval)."""

    _INTS = ("ir", "e", "box", "vir", "pres", "top", "sym", "x", "v", "f", "natoms", "step", "nre")

    def __init__(self, path: str):
        self._mm = np.memmap(path, dtype=np.uint8, mode="r")
        self.big = sys.byteorder != "big"                        # file is big-endian: swap on little-endian hosts
        h = self._header(0)
        if h["x"] == 0:
            raise ValueError(f"{path}: the first TRR frame carries no coordinates")
        real = h["x"] // (3 * h["natoms"])
        if real != 4:
            raise ValueError(f"{path}: double-precision TRR files are not supported")
        self.n_atoms = h["natoms"]
        self.has_box = h["box"] == 36
        self.header_bytes = h["_header"]
        self.box_off = self.header_bytes
        self.x_off = self.header_bytes + h["box"] + h["vir"] + h["pres"]
        self.frame_bytes = self.x_off + h["x"] + h["v"] + h["f"]
        self.n_frames = self._mm.size // self.frame_bytes
        for t in (1, self.n_frames - 1):
            if 0 < t < self.n_frames:
                g = self._header(t * self.frame_bytes)
                if any(g[k] != h[k] for k in ("box", "vir", "pres", "x", "v", "f", "natoms")):
                    raise ValueError(f"{path}: TRR frames of different sizes are not supported")

    def _header(self, pos: int) -> dict:
        mm = self._mm
        if int(mm[pos:pos + 4].view(">i4")[0]) != 1993:
            raise ValueError("not a TRR file (magic number 1993 missing)")
        slen = int(mm[pos + 8:pos + 12].view(">i4")[0])
        p = pos + 12 + (slen + 3) // 4 * 4
        vals = mm[p:p + 52].view(">i4")
        h = {k: int(v) for k, v in zip(self._INTS, vals)}
        real = 8 if (h["box"] == 72 or (h["box"] == 0 and h["natoms"] and h["x"] // (3 * h["natoms"]) == 8)) else 4
        h["_header"] = (p + 52 + 2 * real) - pos
        return h

    def x_ptr(self, t: int) -> int:
        return self._mm.ctypes.data + t * self.frame_bytes + self.x_off

    def box(self, lo: int, hi: int, stride: int = 1) -> np.ndarray:
        idx = np.arange(lo, hi, max(1, stride))
        out = np.empty((len(idx), 3, 3), dtype=np.float64)
        for k, t in enumerate(idx):
            o = int(t) * self.frame_bytes + self.box_off
            out[k] = self._mm[o:o + 36].view(">f4").reshape(3, 3)
        return out

    def close(self):
        self._mm = None


def write_trr(path: str, xyz_nm: np.ndarray, box_nm=None, with_velocities: bool = False) -> None:
    """Minimal single-precision TRR writer (tests and fixtures): coordinates / box in nm."""
    xyz_nm = np.asarray(xyz_nm, dtype=np.float32)
    T, N, _ = xyz_nm.shape
    i4 = lambda *v: np.array(v, dtype=">i4").tobytes()
    with open(path, "wb") as fh:
        for t in range(T):
            box_b = 36 if box_nm is not None else 0
            v_b = 12 * N if with_velocities else 0
            fh.write(i4(1993, 13, 12) + b"GMX_trn_file")
            fh.write(i4(0, 0, box_b, 0, 0, 0, 0, 12 * N, v_b, 0, N, t, 0))
            fh.write(np.array([float(t), 0.0], dtype=">f4").tobytes())
            if box_nm is not None:
                b = np.asarray(box_nm, dtype=np.float64)
                b = b[t] if b.ndim == 3 else b
                fh.write(np.asarray(b, dtype=">f4").tobytes())
            fh.write(xyz_nm[t].astype(">f4").tobytes())
            if with_velocities:
                fh.write(np.zeros((N, 3), dtype=">f4").tobytes())


class _DcdFile:
    """Layout of a CHARMM / NAMD DCD file over a read-only memory map (no frame is copied here)."""

    def __init__(self, path: str):
        self._mm = np.memmap(path, dtype=np.uint8, mode="r")
        mm = self._mm
        if mm.size < 100:
            raise ValueError(f"{path}: too short for a DCD file")
        first = int(mm[:4].view("<i4")[0])
        if first == 84:
            self.swap, e = sys.byteorder != "little", "<"
        elif int(mm[:4].view(">i4")[0]) == 84:
            self.swap, e = sys.byteorder != "big", ">"
        else:
            raise ValueError(f"{path}: not a DCD file (first record marker is not 84)")
        if bytes(mm[4:8]) != b"CORD":
            raise ValueError(f"{path}: DCD header without the CORD tag")
        icntrl = mm[8:88].view(e + "i4")
        n_set, namnf, has_cell, four_d = int(icntrl[0]), int(icntrl[8]), int(icntrl[10]), int(icntrl[11])
        if namnf != 0 or four_d != 0:
            raise ValueError(f"{path}: DCD files with fixed atoms or 4-D coordinates are not supported")
        pos = 92
        title = int(mm[pos:pos + 4].view(e + "i4")[0])
        pos += 4 + title + 4
        if int(mm[pos:pos + 4].view(e + "i4")[0]) != 4:
            raise ValueError(f"{path}: malformed atom-count record")
        self.n_atoms = int(mm[pos + 4:pos + 8].view(e + "i4")[0])
        pos += 12
        self.has_cell = bool(has_cell)
        self.first_frame = pos
        cell_b = 56 if self.has_cell else 0                     # 4 + 6 doubles + 4
        arr_b = 4 * self.n_atoms + 8                            # marker + N floats + marker
        self.frame_bytes = cell_b + 3 * arr_b
        self.x_off, self.y_off, self.z_off = cell_b + 4, cell_b + arr_b + 4, cell_b + 2 * arr_b + 4
        avail = (mm.size - pos) // self.frame_bytes if self.frame_bytes else 0
        self.n_frames = min(n_set, avail) if n_set > 0 else avail      # some writers leave NSET at 0
        self._e = e

    def frame_ptr(self, t: int) -> int:
        return self._mm.ctypes.data + self.first_frame + t * self.frame_bytes

    def cell(self, lo: int, hi: int, stride: int = 1):
        """(lengths (n,3), angles_deg (n,3)) from the per-frame record A, gamma, B, beta, alpha, C (CHARMM stores the
        cosines of the angles, NAMD and others degrees)."""
        idx = np.arange(lo, hi, max(1, stride))
        raw = np.empty((len(idx), 6), dtype=np.float64)
        for k, t in enumerate(idx):
            o = self.first_frame + int(t) * self.frame_bytes + 4
            raw[k] = self._mm[o:o + 48].view(self._e + "f8")
        L = raw[:, [0, 2, 5]]
        ang = raw[:, [4, 3, 1]]
        if np.all(np.abs(ang) <= 1.0):
            ang = np.degrees(np.arccos(ang))
        return L, ang

    def close(self):
        self._mm = None


def write_dcd(path: str, xyz: np.ndarray, cell_lengths=None, cell_angles=None, big_endian: bool = False,
              cosines: bool = False) -> None:
    """Minimal CHARMM-style DCD writer (tests and fixtures): optional unit cell, either byte order."""
    xyz = np.asarray(xyz, dtype=np.float32)
    T, N, _ = xyz.shape
    e = ">" if big_endian else "<"
    i4 = lambda *v: np.array(v, dtype=e + "i4").tobytes()
    icntrl = np.zeros(20, dtype=e + "i4")
    icntrl[0], icntrl[1], icntrl[2], icntrl[19] = T, 1, 1, 24
    icntrl[10] = 1 if cell_lengths is not None else 0
    title = b"CAVIAR_B200 TEST FIXTURE".ljust(80)
    with open(path, "wb") as fh:
        fh.write(i4(84) + b"CORD" + icntrl.tobytes() + i4(84))
        fh.write(i4(84) + i4(1) + title + i4(84))
        fh.write(i4(4) + i4(N) + i4(4))
        if cell_lengths is not None:
            L = np.broadcast_to(np.asarray(cell_lengths, dtype=np.float64), (T, 3))
            A = np.broadcast_to(np.asarray(cell_angles if cell_angles is not None else [90.0] * 3, dtype=np.float64), (T, 3))
            if cosines:
                A = np.cos(np.radians(A))
        for t in range(T):
            if cell_lengths is not None:
                rec = np.array([L[t, 0], A[t, 2], L[t, 1], A[t, 1], A[t, 0], L[t, 2]], dtype=e + "f8")
                fh.write(i4(48) + rec.tobytes() + i4(48))
            for c in range(3):
                fh.write(i4(4 * N) + xyz[t, :, c].astype(e + "f4").tobytes() + i4(4 * N))


def write_netcdf(path: str, xyz: np.ndarray, cell_lengths=None, cell_angles=None) -> None:
    """Amber-convention NetCDF-3 (64-bit offset) trajectory; for tests and synthetic fixtures."""
    from scipy.io import netcdf_file
    xyz = np.asarray(xyz, dtype=np.float32)
    T, N, _ = xyz.shape
    nc = netcdf_file(path, "w", version=2)
    nc.Conventions = "AMBER"
    nc.ConventionVersion = "1.0"
    nc.program = "caviar_b200"
    nc.createDimension("frame", T)
    nc.createDimension("spatial", 3)
    nc.createDimension("atom", N)
    v = nc.createVariable("coordinates", "f", ("frame", "atom", "spatial"))
    v.units = "angstrom"
    v[:] = xyz
    if cell_lengths is not None:
        nc.createDimension("cell_spatial", 3)
        nc.createDimension("cell_angular", 3)
        cl = nc.createVariable("cell_lengths", "d", ("frame", "cell_spatial"))
        cl.units = "angstrom"
        cl[:] = np.broadcast_to(np.asarray(cell_lengths, dtype=np.float64), (T, 3))
        ca = nc.createVariable("cell_angles", "d", ("frame", "cell_angular"))
        ca.units = "degree"
        ca[:] = np.broadcast_to(np.asarray(cell_angles if cell_angles is not None else [90.0] * 3, dtype=np.float64), (T, 3))
    nc.close()
