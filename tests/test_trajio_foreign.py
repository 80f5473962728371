"""Trajectory readers against files assembled BYTE BY BYTE from the published format descriptions -- not by
caviar_b200.trajio's own writers -- so that a mis-read convention (record markers, unit-cell order, box rows, XDR
strings, interleaved records) cannot hide behind a writer that shares it.

  DCD   CHARMM / NAMD / X-PLOR binary trajectory: Fortran unformatted records (4-byte length before and after each);
        header 'CORD' + 20 int32 (NSET first, unit-cell flag 11th, CHARMM version 20th), title record (NTITLE x 80
        chars), atom-count record; per frame an optional 6-double cell record ordered A, gamma, B, beta, alpha, C
        (NAMD stores the COSINES of the angles, CHARMM >= 25 degrees), then X, Y, Z float32 records.
  TRR   GROMACS full-precision trajectory, XDR (big-endian): per frame magic 1993, version-string (XDR string: length
        word + bytes padded to 4), 13 int32 sizes (ir, e, box, vir, pres, top, sym, x, v, f, natoms, step, nre), time and
        lambda reals, then box (3x3 reals, rows = box vectors), virial, pressure, x, v, f.  nm.
  NetCDF  tests/golden/extractor/*.nc (written by scipy with an UNLIMITED frame dimension) are checked in
        tests/test_extractor_golden.py.
"""
import struct

import numpy as np
import pytest


@pytest.fixture(scope="module")
def trajio():
    import __graft_entry__ as g
    g.build()
    from caviar_b200 import trajio
    return trajio


def _coords(T, N):
    t, a, c = np.meshgrid(np.arange(T), np.arange(N), np.arange(3), indexing="ij")
    return (100.0 * t + a + 0.25 * c - 7.5).astype(np.float32)


def _fortran(e, payload: bytes) -> bytes:
    return struct.pack(e + "i", len(payload)) + payload + struct.pack(e + "i", len(payload))


def _dcd(e, xyz, cell=None, cosines=False, charmm_version=24, n_title=2, nset=None):
    T, N, _ = xyz.shape
    icntrl = [0] * 20
    icntrl[0] = T if nset is None else nset            # NSET
    icntrl[1], icntrl[2] = 1000, 500                   # ISTART, NSAVC
    icntrl[3] = T * 500                                # NSTEP
    icntrl[10] = 1 if cell is not None else 0
    icntrl[19] = charmm_version
    head = b"CORD" + struct.pack(e + "9i", *icntrl[:9]) + struct.pack(e + "f", 0.0488882) + struct.pack(e + "10i", *icntrl[10:])
    out = _fortran(e, head)
    out += _fortran(e, struct.pack(e + "i", n_title) + b"".join((b"REMARKS line %d" % k).ljust(80) for k in range(n_title)))
    out += _fortran(e, struct.pack(e + "i", N))
    for t in range(T):
        if cell is not None:
            (a, b, c), (al, be, ga) = cell[t]
            ang = [np.cos(np.radians(v)) for v in (al, be, ga)] if cosines else [al, be, ga]
            out += _fortran(e, struct.pack(e + "6d", a, ang[2], b, ang[1], ang[0], c))       # A, gamma, B, beta, alpha, C
        for k in range(3):
            out += _fortran(e, struct.pack(e + "%df" % N, *xyz[t, :, k].tolist()))
    return out


@pytest.mark.parametrize("e,cosines,version", [("<", True, 24), (">", False, 27), ("<", False, 0)])
def test_dcd_assembled_from_the_format_description(trajio, tmp_path, e, cosines, version):
    T, N = 5, 37
    xyz = _coords(T, N)
    cell = [((40.0 + t, 41.5, 39.25), (88.0, 93.5, 101.25 + 0.5 * t)) for t in range(T)] if version else None
    p = tmp_path / "foreign.dcd"
    p.write_bytes(_dcd(e, xyz, cell, cosines, version, n_title=3, nset=0 if version == 27 else None))
    tr = trajio.Trajectory(str(p))
    assert (tr.n_frames, tr.n_atoms) == (T, N)
    assert np.array_equal(tr.coords(0, T), xyz) and np.array_equal(tr.coords(1, T, 2), xyz[1:T:2])
    sel = np.array([36, 0, 5, 5, 12], dtype=np.int32)
    assert np.array_equal(tr.coords_of(sel, 0, T), xyz[:, sel, :])
    if cell is None:
        assert tr.box_kind() is None and tr.cell(0, T) is None
    else:
        assert tr.box_kind() == "triclinic"
        from oracle import caviar_oracle as orc
        want = np.stack([orc.box_matrix(L, A) for L, A in cell])
        np.testing.assert_allclose(tr.cell(0, T), want, rtol=0, atol=2e-5)
    tr.close()
    # orthorhombic cell given as cosines that are exactly 0
    p2 = tmp_path / "ortho.dcd"
    p2.write_bytes(_dcd(e, xyz, [((30.0, 31.0, 32.0), (90.0, 90.0, 90.0))] * T, True, 24))
    tr = trajio.Trajectory(str(p2))
    assert tr.box_kind() == "ortho" and np.allclose(tr.cell(0, T), [[30.0, 31.0, 32.0]] * T, atol=1e-5)
    tr.close()


def _trr(xyz_nm, box=None, with_vir_pres=False, with_v=False, with_f=False, version=b"GMX_trn_file"):
    T, N, _ = xyz_nm.shape
    out = b""
    for t in range(T):
        sizes = [0, 0, 36 if box is not None else 0, 36 if with_vir_pres else 0, 36 if with_vir_pres else 0, 0, 0,
                 12 * N, 12 * N if with_v else 0, 12 * N if with_f else 0, N, 10 * t, 0]
        pad = (-len(version)) % 4
        out += struct.pack(">ii", 1993, len(version) + 1)                      # magic, slen (counts the terminating NUL)
        out += struct.pack(">i", len(version)) + version + b"\0" * pad         # XDR string
        out += struct.pack(">13i", *sizes)
        out += struct.pack(">ff", 0.002 * t, 0.5)                              # t, lambda
        if box is not None:
            out += struct.pack(">9f", *np.asarray(box[t], dtype=np.float64).ravel().tolist())
        if with_vir_pres:
            out += struct.pack(">9f", *range(9)) + struct.pack(">9f", *range(9, 18))
        out += struct.pack(">%df" % (3 * N), *xyz_nm[t].ravel().tolist())
        if with_v:
            out += struct.pack(">%df" % (3 * N), *([0.125] * (3 * N)))
        if with_f:
            out += struct.pack(">%df" % (3 * N), *([-3.5] * (3 * N)))
    return out


@pytest.mark.parametrize("vp,v,f", [(False, False, False), (True, True, True), (True, False, True)])
def test_trr_assembled_from_the_format_description(trajio, tmp_path, vp, v, f):
    T, N = 4, 23
    xyz_nm = _coords(T, N) / np.float32(16.0)                                   # exact in float32
    # rows = box vectors a, b, c (GROMACS: a along x, b in the xy plane), nm
    box = [np.array([[4.0 + 0.125 * t, 0, 0], [-1.25, 3.5, 0], [0.75, -1.5, 3.0]]) for t in range(T)]
    p = tmp_path / "foreign.trr"
    p.write_bytes(_trr(xyz_nm, box, vp, v, f))
    tr = trajio.Trajectory(str(p))
    assert (tr.n_frames, tr.n_atoms) == (T, N)
    want = xyz_nm * np.float32(10.0)                                            # Angstrom
    assert np.array_equal(tr.coords(0, T), want)
    sel = np.array([22, 3, 3, 0], dtype=np.int32)
    assert np.array_equal(tr.coords_of(sel, 1, T, 2), want[1:T:2][:, sel, :])
    assert tr.box_kind() == "triclinic"
    np.testing.assert_allclose(tr.cell(0, T), np.stack(box) * 10.0, rtol=0, atol=1e-5)
    tr.close()
    p2 = tmp_path / "nobox.trr"
    p2.write_bytes(_trr(xyz_nm))
    tr = trajio.Trajectory(str(p2))
    assert tr.box_kind() is None and np.array_equal(tr.coords(0, T), want)
    tr.close()
    ortho = [np.diag([3.0, 4.0, 5.0])] * T
    p3 = tmp_path / "ortho.trr"
    p3.write_bytes(_trr(xyz_nm, ortho, with_v=True))
    tr = trajio.Trajectory(str(p3))
    assert tr.box_kind() == "ortho" and np.allclose(tr.cell(0, T), [[30.0, 40.0, 50.0]] * T)
    tr.close()


def _mdcrd(xyz, box=None, eol="\n"):
    out = "extractor test trajectory" + eol
    for t in range(xyz.shape[0]):
        vals = xyz[t].ravel()
        for k in range(0, len(vals), 10):
            out += "".join("%8.3f" % v for v in vals[k:k + 10]) + eol
        if box is not None:
            out += "".join("%8.3f" % v for v in box[t]) + eol
    return out.encode()


@pytest.mark.parametrize("n,with_box,eol", [(37, True, "\n"), (40, False, "\n"), (11, True, "\r\n"), (1, True, "\n")])
def test_mdcrd_assembled_from_the_format_description(trajio, tmp_path, n, with_box, eol):
    """Amber ASCII trajectory: title, 10F8.3 coordinate lines (values may touch: '-123.456-234.567'), optional box line;
    the atom count comes from the topology."""
    T = 6
    xyz = np.round(_coords(T, n) - 300.0, 3).astype(np.float32)          # negative, touching fields
    box = [(40.0 + t, 41.125, 39.5) for t in range(T)] if with_box else None
    p = tmp_path / "foreign.mdcrd"
    p.write_bytes(_mdcrd(xyz, box, eol))
    tr = trajio.Trajectory(str(p), n_atoms=n)
    assert (tr.n_frames, tr.n_atoms) == (T, n)
    np.testing.assert_allclose(tr.coords(0, T), xyz, rtol=0, atol=6e-5)
    np.testing.assert_allclose(tr.coords(1, T, 2), xyz[1:T:2], rtol=0, atol=6e-5)
    sel = np.array([n - 1, 0, 0], dtype=np.int32)
    np.testing.assert_allclose(tr.coords_of(sel, 0, T), xyz[:, sel, :], rtol=0, atol=6e-5)
    if with_box:
        assert tr.box_kind() == "ortho"
        np.testing.assert_allclose(tr.cell(0, T), np.array(box, dtype=np.float32), atol=1e-3)
    else:
        assert tr.box_kind() is None and tr.cell(0, T) is None
    tr.close()
    with pytest.raises(SystemExit):
        trajio.Trajectory(str(p))                                        # no atom count, no guess
    with pytest.raises(ValueError):
        trajio.Trajectory(str(p), n_atoms=n + 1)


# ---------------------------------------------------------------------------------------------------------
# XTC: GROMACS compressed trajectory (xdrfile's xdr3dfcoord).  Per frame, XDR: magic 1995 (2023: 64-bit byte count),
# natoms, step, time, box (3x3, nm), natoms; for natoms <= 9 plain floats; otherwise precision, minint[3], maxint[3],
# smallidx, byte count, and a most-significant-bit-first bit stream: per atom the three integers (x * precision - minint)
# as ONE number in the mixed radix (sizeint) cut into 8-bit groups, LEAST significant group first, the last group with the
# remaining bits; then a 1-bit "run changed" flag (+ 5 bits: run length * 3 + (smaller/same/larger radix) code) and,
# inside a run, one small number per atom: offsets (+ smallnum) from the previous atom in the radix magicints[smallidx],
# smallidx bits each; the first two atoms of a run are stored swapped.  The fixture below is assembled with Python's big
# integers and string formatting -- nothing of trajio's writer -- and holds a run that is written out by hand.
# ---------------------------------------------------------------------------------------------------------
def _bits(value, n):
    return format(int(value), "0{}b".format(n)) if n else ""


def _groups(value, nbits):
    out = ""
    while nbits > 8:
        out += _bits(value & 0xFF, 8)
        value >>= 8
        nbits -= 8
    return out + _bits(value, nbits)


def _xtc_frame(ints, precision, box, step, runs=()):
    """`ints` (N, 3) integer positions; `runs`: indices i such that atoms i, i+1, i+2 form one run of two small offsets
    (i+1 and i+2 relative to their predecessors), written the way the format asks: atom i+1 first (large), then atom i
    and atom i+2 as small numbers."""
    N = len(ints)
    lo, hi = ints.min(axis=0), ints.max(axis=0)
    size = [int(v) for v in (hi - lo + 1)]
    bitsize = (size[0] * size[1] * size[2]).bit_length()
    smallidx = 21                                   # radix 128 (magicints[21]), offsets within +-64, 21 bits per atom
    radix, smallnum = 128, 64
    stream, i, prevrun = "", 0, -1
    while i < N:
        if i in runs:
            first, second, third = ints[i + 1], ints[i], ints[i + 2]          # stored order: swapped pair, then the third
            rel = first - lo
            stream += _groups((int(rel[0]) * size[1] + int(rel[1])) * size[2] + int(rel[2]), bitsize)
            stream += "1" + _bits(6 + 1, 5) if prevrun != 6 else "0"          # run of 2 atoms = 6 numbers, same radix
            prevrun = 6
            for cur, prev in ((second, first), (third, second)):
                o = cur - prev + smallnum
                assert (o >= 0).all() and (o < radix).all()
                stream += _groups((int(o[0]) * radix + int(o[1])) * radix + int(o[2]), smallidx)
            i += 3
        else:
            rel = ints[i] - lo
            stream += _groups((int(rel[0]) * size[1] + int(rel[1])) * size[2] + int(rel[2]), bitsize)
            stream += "1" + _bits(0 + 1, 5) if prevrun != 0 else "0"
            prevrun = 0
            i += 1
    stream += "0" * ((-len(stream)) % 8)
    data = bytes(int(stream[k:k + 8], 2) for k in range(0, len(stream), 8))
    head = struct.pack(">iiif", 1995, N, step, 0.5 * step) + np.asarray(box, ">f4").tobytes() + struct.pack(">i", N)
    body = struct.pack(">f", precision) + struct.pack(">6i", *[int(v) for v in lo], *[int(v) for v in hi]) + struct.pack(">ii", smallidx, len(data))
    return head + body + data + b"\0" * ((-len(data)) % 4)


def test_xtc_assembled_from_the_format_description(trajio, tmp_path):
    rng = np.random.default_rng(3)
    N, T = 23, 3
    box = np.array([[3.0, 0, 0], [0.5, 3.1, 0], [0.25, -0.5, 2.9]], np.float32)            # triclinic, nm
    frames, blob = [], b""
    for t in range(T):
        ints = rng.integers(-1500, 4200, size=(N, 3))
        for r in (4, 11):                                  # three atoms within +-40 of each other: a hand-written run
            ints[r + 1] = ints[r] + rng.integers(-40, 40, size=3)
            ints[r + 2] = ints[r + 1] + rng.integers(-20, 20, size=3)
        frames.append(ints)
        blob += _xtc_frame(ints, 1000.0, box, t, runs=(4, 11))
    path = tmp_path / "hand.xtc"
    path.write_bytes(blob)
    tr = trajio.Trajectory(str(path))
    assert (tr.n_frames, tr.n_atoms, tr.box_kind()) == (T, N, "triclinic")
    want = np.stack(frames).astype(np.float32) * np.float32(1.0 / 1000.0) * np.float32(10.0)   # nm -> Angstrom
    assert np.array_equal(tr.coords(0, T), want)
    sel = np.array([12, 0, 5, 22], np.int32)
    assert np.array_equal(tr.coords_of(sel, 1, 3), want[1:3][:, sel])
    np.testing.assert_allclose(tr.cell(0, T), np.broadcast_to(box * 10.0, (T, 3, 3)), rtol=1e-6)
    # few atoms: plain big-endian floats
    x = rng.random((2, 5, 3)).astype(np.float32)
    small = b"".join(struct.pack(">iiif", 1995, 5, t, 0.0) + np.zeros(9, ">f4").tobytes() + struct.pack(">i", 5) + x[t].astype(">f4").tobytes()
                     for t in range(2))
    p2 = tmp_path / "small.xtc"
    p2.write_bytes(small)
    t2 = trajio.Trajectory(str(p2))
    assert t2.n_frames == 2 and not t2.has_box and np.array_equal(t2.coords(0, 2), x * np.float32(10.0))
    # damaged files are refused, not mis-read
    bad = tmp_path / "bad.xtc"
    bad.write_bytes(blob[:-8])
    with pytest.raises(SystemExit) as e:
        trajio.Trajectory(str(bad))
    assert "[ERR]" in str(e.value) and "XTC" in str(e.value)


def test_xtc_writer_reader_round_trip_with_runs_and_radix_changes(trajio, tmp_path):
    """trajio.write_xtc follows the compressor of the format (runs, swapped water atoms, adaptive radix); the native
    reader must give back exactly the rounded positions, whatever mixture the data produces."""
    rng = np.random.default_rng(0)
    T, nw = 4, 300
    centers = rng.random((T, nw, 1, 3)) * 5.0
    water = centers + rng.normal(0, 0.05, size=(T, nw, 3, 3))
    chain = np.cumsum(rng.normal(0, 0.08, size=(T, 150, 3)), axis=1) + 2.5          # a polymer: long runs of close atoms
    far = rng.random((T, 40, 3)) * 5.0
    xyz = np.concatenate([far[:, :10], water.reshape(T, -1, 3), chain, far[:, 10:]], axis=1).astype(np.float32)
    box = np.diag([5.0, 5.1, 5.2]).astype(np.float32)
    for magic, scale in ((1995, 1.0), (2023, 1.0), (1995, 9000.0)):                # last: ranges too wide for one number
        path = str(tmp_path / f"rt_{magic}_{int(scale)}.xtc")
        want = trajio.write_xtc(path, xyz * np.float32(scale), box, magic=magic)
        tr = trajio.Trajectory(path)
        assert (tr.n_frames, tr.n_atoms, tr.box_kind()) == (T, xyz.shape[1], "ortho")
        assert np.array_equal(tr.coords(0, T), want * np.float32(10.0))
        assert np.abs(want - xyz * np.float32(scale)).max() <= 0.0005 * max(1.0, scale / 64)
        tr.close()


def test_mdcrd_native_parser_equals_numpy_conversion_on_odd_fields(trajio, tmp_path):
    """cv_mdcrd.cu against numpy's string -> float64 conversion (the parser trajio used before), bit for bit: touching
    negative fields, Fortran's omitted leading zero, an explicit plus sign, left-justified and integer fields, NaN; CR LF
    lines; a selection in arbitrary order with repeats.  A field that is no number ('********') is a ValueError both ways."""
    rng = np.random.default_rng(3)
    n, T = 23, 5
    fields = ["%8.3f" % v for v in rng.uniform(-999.0, 9999.0, size=T * 3 * n)]
    odd = ["   -.123", "    .500", "  +1.500", "1.5     ", "     -12", "12345678", "     NaN", "-999.999", "9999.999", "  -0.000"]
    for k, f in enumerate(odd):
        fields[7 * k + 3] = f
    for eol in ("\n", "\r\n"):
        text = "odd fields" + eol
        for t in range(T):
            row = fields[t * 3 * n:(t + 1) * 3 * n]
            for k in range(0, len(row), 10):
                text += "".join(row[k:k + 10]) + eol
            text += "%8.3f%8.3f%8.3f" % (50.0, 60.0, 70.0) + eol
        p = tmp_path / ("odd%d.mdcrd" % len(eol))
        p.write_bytes(text.encode())
        tr = trajio.Trajectory(str(p), n_atoms=n)
        assert tr.n_frames == T and tr.box_kind() == "ortho"
        want = tr._crd.coords_numpy(0, T)
        got = tr.coords(0, T)
        assert got.tobytes() == want.tobytes()
        assert np.isnan(got).sum() == 1
        sel = np.array([n - 1, 0, 7, 7, 3], dtype=np.int64)
        assert tr.coords_of(sel, 1, T, 2).tobytes() == np.ascontiguousarray(want[1:T:2][:, sel, :]).tobytes()
        with pytest.raises(IndexError):
            tr.coords_of(np.array([n]), 0, T)
        tr.close()
    bad = text.replace(fields[5], "********", 1)
    pb = tmp_path / "bad.mdcrd"
    pb.write_bytes(bad.encode())
    tr = trajio.Trajectory(str(pb), n_atoms=n)
    with pytest.raises(ValueError):
        tr.coords(0, T)
    with pytest.raises(ValueError):
        tr._crd.coords_numpy(0, T)
    tr.close()


def test_mdcrd_native_parser_random_fields(trajio, tmp_path):
    """20 000 random eight-character fields in several fixed-point shapes (F8.1 ... F8.5, integers, left-justified,
    leading zero dropped): the native parser must return numpy's correctly rounded conversion, bit for bit."""
    rng = np.random.default_rng(17)
    n, T = 667, 10                                     # 3 * 667 = 2001 values per frame: a ragged last line
    fields = []
    for k in range(T * 3 * n):
        kind = k % 7
        if kind == 0:
            f = "%8.3f" % rng.uniform(-999.0, 9999.0)
        elif kind == 1:
            f = "%8.1f" % rng.uniform(-99999.0, 999999.0)
        elif kind == 2:
            f = "%8.5f" % rng.uniform(-9.0, 99.0)
        elif kind == 3:
            f = "%8d" % rng.integers(-999999, 9999999)
        elif kind == 4:
            f = "%-8.3f" % rng.uniform(-99.0, 999.0)
        elif kind == 5:
            f = ("%8.3f" % rng.uniform(-0.999, 0.999)).replace("0.", " .", 1).replace("- .", " -.", 1)
        else:
            f = "%8.2f" % rng.uniform(-9999.0, 99999.0)
        assert len(f) == 8, f
        fields.append(f)
    text = "random fields\n"
    for t in range(T):
        row = fields[t * 3 * n:(t + 1) * 3 * n]
        for k in range(0, len(row), 10):
            text += "".join(row[k:k + 10]) + "\n"
    p = tmp_path / "random.mdcrd"
    p.write_bytes(text.encode())
    tr = trajio.Trajectory(str(p), n_atoms=n)
    assert tr.n_frames == T and tr.box_kind() is None
    want = np.array([float(f) for f in fields], dtype=np.float64).astype(np.float32).reshape(T, n, 3)
    assert tr.coords(0, T).tobytes() == want.tobytes()
    assert tr._crd.coords_numpy(0, T).tobytes() == want.tobytes()
    sel = rng.integers(0, n, size=100)
    assert tr.coords_of(sel, 0, T, 3).tobytes() == np.ascontiguousarray(want[0:T:3][:, sel, :]).tobytes()
    tr.close()
