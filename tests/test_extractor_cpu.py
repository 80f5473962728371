"""Host logic of the Extractor drop-in: label grammar pinned to the reference's own outputs, topology / trajectory
readers, atom selection, CSV formatting.  No GPU needed."""
import json
import os

import numpy as np
import pytest


@pytest.fixture(scope="module")
def ext():
    import __graft_entry__ as g
    g.build()
    from caviar_b200 import extractor
    return extractor


@pytest.fixture(scope="module")
def meta(golden_dir):
    with open(os.path.join(golden_dir, "reference_cases.json")) as fh:
        return json.load(fh)


def test_label_grammar_matches_reference_vectors(ext, meta, tmp_path):
    par = meta["parser"]                                   # produced by the reference's _parse_pair_token etc.
    for tok, want in par["tokens"]:
        got = ext.PairGrammar.token(tok)
        assert ([list(got)] if got else []) == want, tok
    for s, want in par["strings"]:
        assert [list(p) for p in ext.PairGrammar.many(s.split(","))] == want, s
    for case in par["files"]:
        f = tmp_path / "pairs.txt"
        f.write_text(case["body"])
        assert [list(p) for p in ext.read_pairs(case["pairs"], str(f))] == case["result"]
    with pytest.raises(SystemExit) as e:
        ext.read_pairs("nothing here", "")
    assert "Nessuna coppia valida" in str(e.value)


def _toy_topology(n_res=12):
    names, labels, first = [], [], []
    res3 = ["ASP", "GLU", "ALA", "GLY", "SER", "THR", "LYS", "ARG", "HIS", "PRO", "VAL", "LEU"]
    for r in range(n_res):
        first.append(len(names))
        labels.append(res3[r % len(res3)])
        names += ["N", "H", "CA", "HA", "C", "O"] if r % 3 else ["N", "CA", "C", "O", "CB"]
    return names, labels, first


def test_prmtop_roundtrip_and_selection(ext, tmp_path):
    from caviar_b200 import trajio
    names, labels, first = _toy_topology()
    path = str(tmp_path / "toy.prmtop")
    trajio.write_prmtop(path, names, labels, first)
    top = trajio.read_topology(path)
    assert top.n_atoms == len(names) and top.n_res == len(labels)
    assert top.atom_names == names and top.res_labels == labels and top.res_first_atom == first
    ca = [top.atom_of(r) for r in range(top.n_res)]
    assert all(names[a] == "CA" for a in ca) and ca[0] == 1 and ca[1] == first[1] + 2
    labs = [("ASP", 22, "GLU", 23), ("ALA", 24, "LEU", 33)]
    atoms = ext.select_atoms(top, labs, "sequential", 21)          # label - offset = 1-based residue index
    assert atoms.tolist() == [[ca[0], ca[1]], [ca[2], ca[11]]]
    with pytest.raises(SystemExit) as e:
        ext.select_atoms(top, [("ASP", 21, "GLU", 23)], "sequential", 21)
    assert "label-offset < 1" in str(e.value)
    with pytest.raises(SystemExit):
        ext.select_atoms(top, [("ASP", 22, "GLU", 99)], "sequential", 21)


def test_pdb_resseq_selection(ext, tmp_path):
    from caviar_b200 import trajio
    lines = []
    serial = 1
    for resseq, res in ((40, "ASP"), (41, "GLU"), (45, "ALA")):
        for nm in ("N", "CA", "C"):
            lines.append(f"ATOM  {serial:5d} {nm:^4s} {res} A{resseq:4d}    {1.0:8.3f}{2.0:8.3f}{3.0:8.3f}  1.00  0.00")
            serial += 1
    p = tmp_path / "toy.pdb"
    p.write_text("\n".join(lines) + "\nEND\n")
    top = trajio.read_topology(str(p))
    assert top.res_seq == [40, 41, 45] and top.n_atoms == 9
    atoms = ext.select_atoms(top, [("ASP", 40, "ALA", 45)], "resSeq", 0)
    assert atoms.tolist() == [[1, 7]]


def test_netcdf_roundtrip_and_cell(tmp_path):
    from caviar_b200 import trajio
    from oracle import caviar_oracle as orc
    rng = np.random.default_rng(0)
    xyz = rng.uniform(0, 50, size=(7, 11, 3)).astype(np.float32)
    p = str(tmp_path / "t.nc")
    trajio.write_netcdf(p, xyz, [50.0, 51.0, 52.0], [109.4712206] * 3)
    tr = trajio.Trajectory(p)
    assert (tr.n_frames, tr.n_atoms) == (7, 11) and tr.box_kind() == "triclinic"
    assert np.array_equal(tr.coords(0, 7), xyz) and np.array_equal(tr.coords(1, 7, 2), xyz[1:7:2])
    cell = tr.cell(0, 7)
    want = orc.box_matrix([50.0, 51.0, 52.0], [109.4712206] * 3)
    assert np.allclose(cell[3], want, atol=1e-4) and cell[0, 0, 1] == 0 and cell[0, 0, 2] == 0 and cell[0, 1, 2] == 0
    tr.close()
    trajio.write_netcdf(p, xyz, [50.0, 51.0, 52.0])
    tr = trajio.Trajectory(p)
    assert tr.box_kind() == "ortho" and np.allclose(tr.cell(0, 2), [[50, 51, 52]] * 2)
    tr.close()
    trajio.write_netcdf(p, xyz)
    tr = trajio.Trajectory(p)
    assert tr.box_kind() is None and tr.cell(0, 2) is None
    tr.close()


def test_csv_rows_match_printf(ext):
    """Rows as Python's csv.writer would write printf-formatted fields (Extractor:196-198): CR LF line ends."""
    import csv
    import io
    rng = np.random.default_rng(1)
    v = rng.uniform(0, 300, size=(257, 33)).astype(np.float32)
    v[0, :6] = [0.00005, 0.12345, 0.12355, 7.99995, 0.0, 123456.7]
    got = ext.format_rows(v, 1, 1, 4).decode()
    sio = io.StringIO(newline="")
    csv.writer(sio).writerows([[str(r + 1)] + ["%.4f" % float(x) for x in v[r]] for r in range(v.shape[0])])
    assert got == sio.getvalue() and got.count("\r\n") == v.shape[0]
    assert ext.format_rows(v[:2, :2], 5, 3, 2, crlf=False).decode() == "".join(
        f"{5 + 3 * r}," + ",".join("%.2f" % float(x) for x in v[r, :2]) + "\n" for r in range(2))


def test_csv_rows_of_any_magnitude(ext):
    """Values whose scaled magnitude leaves 64 bits, negative zeros, non-finite values: printf's text, no overrun."""
    v = np.array([[5e11, -5e11, 3e38, -3.4028235e38, 1.8e11, 1.9e11, -0.0, -1e-9, np.inf, -np.inf, np.nan, 1e15, 123.456]],
                 dtype=np.float32)
    for dec in (0, 4, 8, 12, 17):
        got = ext.format_rows(np.repeat(v, 3, axis=0), 7, 2, dec, crlf=False).decode().splitlines()
        for r, line in enumerate(got):
            want = f"{7 + 2 * r}," + ",".join("%.*f" % (dec, float(x)) for x in v[0])
            assert line == want, (dec, line, want)


def test_residue_components_match_reference_vectors(meta, golden_dir):
    """f4: thresholded mean-distance graph -> components, pinned to build_residue_components' own output."""
    from caviar_b200.components import pair_means_from_aggregates, residue_components
    gold = np.load(os.path.join(golden_dir, "distances.npz"))
    agg = meta["aggregates"]
    n, m = agg["n_res"], agg["minsep"]
    pairs = [(i, j) for i in range(n) for j in range(i + m, n)]
    d = gold["B_d"]                                                     # the reference's own (T, P) distances
    means = pair_means_from_aggregates(d.astype(np.float64).sum(0), d.shape[0])
    for cut_repr, want in agg["components"].items():
        cut = float(cut_repr)
        # the reference thresholds ITS float32-accumulated mean; skip cut-offs that sit within float32 rounding of one
        if np.any(np.abs(means.astype(np.float64) - cut) < 4e-6 * cut):
            continue
        assert residue_components(n, pairs, means, cut) == want


def test_cohens_d_from_running_sums_matches_reference_formula():
    from caviar_b200.components import cohens_d_from_aggregates
    from oracle import caviar_oracle as orc
    rng = np.random.default_rng(4)
    a = rng.normal(7.0, 1.5, size=(400, 6)).astype(np.float32)
    b = rng.normal(7.6, 0.9, size=(250, 6)).astype(np.float32)
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    got = cohens_d_from_aggregates(a64.sum(0), (a64 ** 2).sum(0), 400, b64.sum(0), (b64 ** 2).sum(0), 250)
    want = np.array([orc.cohens_d(a[:, k], b[:, k]) for k in range(6)])
    np.testing.assert_allclose(got, want, rtol=1e-9)


def test_all_pairs_and_trim_mirror_the_reference(meta):
    from caviar_b200.selection import all_pairs, ca_atom_indices, trim_ca_indices
    from oracle import caviar_oracle as orc
    for n, m in ((48, 5), (10, 1), (7, 0), (6, 9), (1, 5), (300, 5)):
        want = np.array(orc.all_pairs(n, m), dtype=np.int32).reshape(-1, 2)
        assert np.array_equal(all_pairs(n, m), want), (n, m)
    assert len(all_pairs(1000, 5)) == 495510
    idx = list(range(100, 130))
    for key, want in meta["aggregates"]["trim"].items():          # produced by the reference's _trim_ca_indices
        tn, tc = (int(v) for v in key.split(","))
        assert list(trim_ca_indices(idx, tn, tc)) == want
    with pytest.raises(SystemExit):
        trim_ca_indices(idx, 20, 10)

    class _Top:
        atom_names = ["N", "CA", "C", "O", "N", "H", "CA", "C"]
    assert ca_atom_indices(_Top()).tolist() == [1, 6]


def test_tic_weights_follow_the_label_order(ext, tmp_path):
    """abs_loading of CAVIAR's TIC csv (CAVIAR-1.0.py:141-147) as per-pair weights; unlisted pairs weigh 0."""
    from oracle import caviar_oracle as orc
    labels3 = {0: "ASP25", 1: "GLU216", 2: "ALA15", 3: "GLY26"}
    rows = orc.tic_csv_rows([(0, 1), (2, 3)], np.array([-0.75, 0.125]), labels3)
    f = tmp_path / "tic.csv"
    f.write_text("\n".join(rows) + "\n")
    labs = ext.read_pairs("SER1-THR2", str(f))
    assert labs == [("SER", 1, "THR", 2), ("ASP", 25, "GLU", 216), ("ALA", 15, "GLY", 26)]
    assert ext.read_tic_weights(str(f), labs).tolist() == [0.0, 0.75, 0.125]


@pytest.mark.parametrize("fmt", ["nc", "npy"])
def test_native_frame_compaction_equals_numpy(tmp_path, fmt):
    """Trajectory.coords_of: selected atoms of strided frames straight from the mapped file (big-endian NetCDF records
    or a native .npy), against numpy's byte-swapping copy + fancy index."""
    from caviar_b200 import trajio
    rng = np.random.default_rng(4)
    T, N = 23, 157
    xyz = rng.normal(0, 30, size=(T, N, 3)).astype(np.float32)
    path = str(tmp_path / f"t.{fmt}")
    if fmt == "nc":
        trajio.write_netcdf(path, xyz, np.array([50.0, 60.0, 70.0]))
    else:
        np.save(path, xyz)
    tr = trajio.Trajectory(path)
    atoms = np.array([156, 0, 3, 3, 77, 12])
    for lo, hi, stride in ((0, T, 1), (2, 21, 3), (5, 6, 1), (4, 4, 1)):
        got = tr.coords_of(atoms, lo, hi, stride)
        want = xyz[lo:hi:stride][:, atoms, :]
        assert got.dtype == np.float32 and got.flags.c_contiguous and got.shape == want.shape
        assert got.tobytes() == want.tobytes()
    assert tr.coords_of(np.zeros(0, dtype=np.int64), 0, T).shape == (T, 0, 3)
    with pytest.raises(Exception):
        tr.coords_of(np.array([N]), 0, T)
    tr.close()


@pytest.mark.parametrize("big,cell,cosines", [(False, None, False), (True, "ortho", False), (False, "tric", True), (True, "tric", False)])
def test_dcd_trajectories(tmp_path, big, cell, cosines):
    """CHARMM / NAMD DCD: header, either byte order, optional unit cell (degrees or cosines), native gather of the
    selected atoms out of the X / Y / Z records."""
    from caviar_b200 import trajio
    rng = np.random.default_rng(7)
    T, N = 19, 83
    xyz = rng.normal(0, 25, size=(T, N, 3)).astype(np.float32)
    L = rng.uniform(40, 60, size=(T, 3))
    ang = {"ortho": [90.0, 90.0, 90.0], "tric": [80.0, 95.0, 109.4712206]}.get(cell)
    path = str(tmp_path / "t.dcd")
    trajio.write_dcd(path, xyz, L if cell else None, ang, big_endian=big, cosines=cosines)
    tr = trajio.Trajectory(path)
    assert (tr.n_frames, tr.n_atoms) == (T, N)
    assert tr.box_kind() == {None: None, "ortho": "ortho", "tric": "triclinic"}[cell]
    atoms = np.array([82, 0, 5, 5, 41])
    for lo, hi, stride in ((0, T, 1), (3, 18, 4), (7, 8, 1)):
        got = tr.coords_of(atoms, lo, hi, stride)
        assert got.tobytes() == np.ascontiguousarray(xyz[lo:hi:stride][:, atoms, :]).tobytes()
    assert tr.coords(2, 5).tobytes() == xyz[2:5].tobytes()
    if cell == "ortho":
        np.testing.assert_allclose(tr.cell(0, T, 2), L[::2].astype(np.float32))
    elif cell == "tric":
        want = trajio.cell_to_matrix(L[1:9:3], np.broadcast_to(ang, (3, 3)))
        np.testing.assert_allclose(tr.cell(1, 9, 3), want.astype(np.float32), rtol=1e-6, atol=1e-5)
    tr.close()
    bad = tmp_path / "bad.dcd"
    bad.write_bytes(b"\\x00" * 200)
    with pytest.raises(ValueError):
        trajio.Trajectory(str(bad))


def test_psf_topology(ext, tmp_path):
    """CHARMM / NAMD PSF next to a DCD: residues from (segid, resid), CA selection as with prmtop / PDB."""
    from caviar_b200 import trajio
    rows, k = [], 0
    for seg, resids in (("PROA", [5, 6, 7]), ("PROB", [5, 9])):
        for r in resids:
            for name in ("N", "HN", "CA", "C", "O"):
                k += 1
                rows.append(f"{k:10d} {seg:8s} {r:<8d} {'ASP' if r % 2 else 'GLUP':8s} {name:8s} {'NH1':6s} {-0.47:10.6f} {14.007:13.4f} {0:11d}")
    psf = tmp_path / "sys.psf"
    psf.write_text("PSF EXT CMAP\n\n         1 !NTITLE\n* test\n\n" + f"{k:10d} !NATOM\n" + "\n".join(rows) + "\n\n         0 !NBOND: bonds\n")
    top = trajio.read_topology(str(psf))
    assert top.n_atoms == 25 and top.n_res == 5
    assert top.res_labels == ["ASP", "GLU", "ASP", "ASP", "ASP"] and top.res_seq == [5, 6, 7, 5, 9]
    assert [top.atom_of(r) for r in range(5)] == [2, 7, 12, 17, 22]
    atoms = ext.select_atoms(top, [("ASP", 1, "GLU", 2), ("ASP", 4, "ASP", 5)], "sequential", 0)
    assert atoms.tolist() == [[2, 7], [17, 22]]
    with pytest.raises(ValueError):
        trajio.read_psf(str(tmp_path / "sys.psf").replace("sys", "nope")) if False else trajio.read_psf(__file__)


@pytest.mark.parametrize("box,vel", [(None, False), ("ortho", True), ("tric", False)])
def test_trr_trajectories_and_gro_topology(ext, tmp_path, box, vel):
    """GROMACS TRR (XDR, nm) -> Angstrom, box rows = lattice vectors; .gro as topology."""
    from caviar_b200 import trajio
    rng = np.random.default_rng(3)
    T, N = 11, 37
    x_nm = rng.normal(0, 2.5, size=(T, N, 3)).astype(np.float32)
    h = None
    if box == "ortho":
        h = np.diag([4.0, 5.0, 6.0])
    elif box == "tric":
        h = np.array([[5.0, 0, 0], [1.0, 4.5, 0], [-0.5, 1.5, 4.0]])
    path = str(tmp_path / "t.trr")
    trajio.write_trr(path, x_nm, h, with_velocities=vel)
    tr = trajio.Trajectory(path)
    assert (tr.n_frames, tr.n_atoms) == (T, N)
    assert tr.box_kind() == {None: None, "ortho": "ortho", "tric": "triclinic"}[box]
    atoms = np.array([36, 0, 7, 7])
    got = tr.coords_of(atoms, 1, 10, 3)
    assert got.tobytes() == (x_nm[1:10:3][:, atoms, :] * np.float32(10.0)).tobytes()
    assert tr.coords(0, 2).shape == (2, N, 3)
    if box == "ortho":
        np.testing.assert_allclose(tr.cell(0, T), np.broadcast_to([40.0, 50.0, 60.0], (T, 3)))
    elif box == "tric":
        np.testing.assert_allclose(tr.cell(2, 4), np.broadcast_to(h * 10.0, (2, 3, 3)), rtol=1e-6)
    tr.close()
    gro = tmp_path / "s.gro"
    lines = ["test", f"{6:5d}"]
    for k, (rid, rn, an) in enumerate([(1, "ASP", "N"), (1, "ASP", "CA"), (1, "ASP", "C"), (2, "GLU", "N"), (2, "GLU", "CA"), (2, "GLU", "C")], 1):
        lines.append(f"{rid:5d}{rn:<5s}{an:>5s}{k:5d}{0.1 * k:8.3f}{0.0:8.3f}{0.0:8.3f}")
    gro.write_text("\n".join(lines) + "\n   4.00000   5.00000   6.00000\n")
    top = trajio.read_topology(str(gro))
    assert top.n_atoms == 6 and top.res_labels == ["ASP", "GLU"] and [top.atom_of(0), top.atom_of(1)] == [1, 4]
    assert ext.select_atoms(top, [("ASP", 1, "GLU", 2)], "resSeq", 0).tolist() == [[1, 4]]


def test_unsupported_inputs_fail_with_a_clear_message(ext, tmp_path):
    """Formats this reader does not take, and damaged files, exit with [ERR],
    not with a scipy traceback; resSeq numbering needs a topology that carries residue numbers."""
    from caviar_b200 import trajio
    for name, head in (("t4.nc", b"\x89HDF\r\n\x1a\n" + b"\0" * 60), ("t.foo", b"whatever")):
        p = tmp_path / name
        p.write_bytes(head)
        with pytest.raises(SystemExit) as e:
            trajio.Trajectory(str(p))
        assert "[ERR] formato di traiettoria non supportato" in str(e.value), name
    # XTC is read (tests/test_trajio_foreign.py); a damaged one is refused with a message, not mis-read
    p = tmp_path / "t.xtc"
    p.write_bytes(b"\x00\x00\x07\xcb" + b"\0" * 60)
    with pytest.raises(SystemExit) as e:
        trajio.Trajectory(str(p))
    assert "[ERR]" in str(e.value) and "XTC" in str(e.value)
    names, labels, first = ["N", "CA", "C"] * 4, ["ALA", "GLY", "SER", "THR"], [0, 3, 6, 9]
    top_path = tmp_path / "s.prmtop"
    trajio.write_prmtop(str(top_path), names, labels, first)
    top = trajio.read_topology(str(top_path))
    with pytest.raises(SystemExit) as e:
        ext.select_atoms(top, [("ALA", 1, "GLY", 2)], "resSeq", 0)
    assert "resSeq" in str(e.value) and "[ERR]" in str(e.value)
    # a multi-MODEL PDB is one topology, not the models concatenated
    pdb = tmp_path / "m.pdb"
    model = "".join("ATOM  %5d  CA  %3s A%4d    %8.3f%8.3f%8.3f\n" % (k + 1, labels[k], 10 + k, k, 0.0, 0.0) for k in range(4))
    pdb.write_text("MODEL        1\n" + model + "ENDMDL\nMODEL        2\n" + model + "ENDMDL\nEND\n")
    t = trajio.read_topology(str(pdb))
    assert t.n_atoms == 4 and t.n_res == 4 and t.res_seq == [10, 11, 12, 13]
