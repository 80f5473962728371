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
