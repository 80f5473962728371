"""End-to-end run of the Extractor drop-in on a synthetic Amber NetCDF + prmtop, checked against the CPU oracle."""
import csv
import json

import numpy as np
import pytest

from oracle import caviar_oracle as orc

pytestmark = pytest.mark.gpu


def _system(tmp_path, box, fmt="nc"):
    from caviar_b200 import trajio
    rng = np.random.default_rng(9)
    n_res, T = 60, 53
    names, labels, first = [], [], []
    res3 = ["ASP", "GLU", "ALA", "GLY", "SER", "THR", "LYS", "ARG", "HIS", "PRO"]
    for r in range(n_res):
        first.append(len(names))
        labels.append(res3[r % 10])
        names += ["N", "H", "CA", "HA", "C", "O", "CB"]
    n_atoms = len(names)
    L = np.array([38.0, 40.0, 36.0])
    base = np.cumsum(rng.normal(0, 1.5, size=(n_atoms, 3)), axis=0) % L
    xyz = ((base[None] + rng.normal(0, 0.4, size=(T, n_atoms, 3))) % L).astype(np.float32)
    top = str(tmp_path / "sys.prmtop")
    trajio.write_prmtop(top, names, labels, first)
    nc = str(tmp_path / f"traj.{fmt}")
    write = trajio.write_netcdf if fmt == "nc" else trajio.write_dcd
    if box == "ortho":
        write(nc, xyz, L)
    elif box == "triclinic":
        write(nc, xyz, L, [109.4712206] * 3)
    else:
        write(nc, xyz)
    ca = np.array([f + 2 for f in first])
    return top, nc, xyz, L, ca, labels


@pytest.mark.parametrize("box,stride,fmt", [("ortho", 1, "nc"), ("triclinic", 2, "nc"), (None, 3, "nc"),
                                            ("triclinic", 1, "dcd"), (None, 2, "dcd")])
def test_extractor_cli_against_oracle(tmp_path, capsys, box, stride, fmt):
    from caviar_b200 import extractor, trajio
    top, nc, xyz, L, ca, labels = _system(tmp_path, box, fmt)
    offset = 21
    # the ranked-loading CSV CAVIAR-1.0.py writes (save_tic_csv, :141-147) is a valid --pairs-file
    sel = [(0, 9), (3, 40), (59, 12), (7, 8), (3, 40)]
    res_label = {i: f"{labels[i]}{i + 1 + offset}" for i in range(len(labels))}
    rows = orc.tic_csv_rows([(a, b) for a, b in sel[:4]], np.array([0.5, -0.4, 0.3, 0.2]), res_label)
    pf = tmp_path / "pooled_distances_TIC1.csv"
    pf.write_text("\n".join(rows) + "\n")
    out = tmp_path / "out.csv"
    npy = tmp_path / "out.npy"
    summ = tmp_path / "summary.json"
    rc = extractor.main(["--top", top, "--traj", nc, "--pairs-file", str(pf), "--pairs", f"{res_label[3]}-{res_label[40]}",
                         "--offset", str(offset), "--stride", str(stride), "--out", str(out),
                         "--npy-out", str(npy), "--summary-out", str(summ), "--weighted-score"])
    assert rc == 0
    printed = capsys.readouterr().out
    assert f"[DONE] Salvato CSV: {out}" in printed and "[INFO] Colonne:" in printed
    # expected column order: --pairs first, then the file rows, de-duplicated (Extractor:69-86)
    order = [(3, 40), (0, 9), (59, 12), (7, 8)]
    want_cols = [f"{res_label[a]}-{res_label[b]}" for a, b in order]
    with open(out) as fh:
        table = list(csv.reader(fh))
    assert table[0] == ["#Frame"] + want_cols
    frames = xyz[::stride]
    assert [int(r[0]) for r in table[1:]] == list(range(1, len(frames) + 1))
    got = np.array([[float(v) for v in r[1:]] for r in table[1:]])
    pairs = [(int(ca[a]), int(ca[b])) for a, b in order]
    if box == "ortho":
        cell = np.broadcast_to(L, (len(frames), 3))
    elif box == "triclinic":
        cell = trajio.cell_to_matrix(np.broadcast_to(L, (len(frames), 3)), np.full((len(frames), 3), 109.4712206))
        cell = cell.astype(np.float32).astype(np.float64)
    else:
        cell = None
    want = orc.pbc_pair_distances(frames, pairs, cell)
    assert np.abs(got - want).max() <= 1e-4 + 0.5e-4            # 1e-4 A tolerance + 4-decimal rounding
    d = np.load(npy)
    assert d.shape == want.shape and np.abs(d - want).max() <= 1e-4
    if box is None:                                             # no box: bit-exact reference arithmetic
        assert d.tobytes() == orc.pair_distances(frames, pairs, n_jobs=1).tobytes()
    s = json.load(open(summ))
    assert s["frames"] == len(frames) and s["columns"] == want_cols
    assert s["occupancy"] == (d.astype(np.float64) < 8.0).sum(0).tolist()
    np.testing.assert_allclose(s["mean"], d.astype(np.float64).mean(0), rtol=1e-12)
    assert s["frame_score"] == (d.astype(np.float64) < 8.0).sum(1).tolist()
    # --weighted-score: abs_loading of the TIC file per column ((3,40) is listed there with |-0.4|)
    assert s["pair_weights"] == [0.4, 0.5, 0.3, 0.2]
    assert s["frame_score_weighted"] == orc.weighted_frame_scores(d.astype(np.float64) < 8.0, s["pair_weights"]).tolist()


@pytest.mark.parametrize("box", [None, "ortho"])
def test_extractor_cli_reads_gromacs_xtc_and_gro(tmp_path, capsys, box):
    """GROMACS inputs end to end: .gro topology + compressed .xtc trajectory (nm, 0.001 nm precision).  The positions
    the reader must deliver are the rounded ones the file holds; on those the no-box distances are bit-exact."""
    from caviar_b200 import extractor, trajio
    rng = np.random.default_rng(2)
    n_res, T = 40, 31
    res3 = ["ASP", "GLU", "ALA", "GLY", "SER"]
    L = np.array([40.0, 35.0, 30.0])                                   # 4.0 / 3.5 / 3.0 nm: exact in float32
    n_atoms = 3 * n_res
    base = np.cumsum(rng.normal(0, 1.5, size=(n_atoms, 3)), axis=0) % L
    xyz = ((base[None] + rng.normal(0, 0.4, size=(T, n_atoms, 3))) % L).astype(np.float32)
    gro = tmp_path / "sys.gro"
    lines = ["toy", str(n_atoms)]
    for a in range(n_atoms):
        r = a // 3
        lines.append("%5d%-5s%5s%5d%8.3f%8.3f%8.3f" % (r + 1, res3[r % 5], ("N", "CA", "C")[a % 3], a + 1, *(xyz[0, a] / 10)))
    lines.append("%10.5f%10.5f%10.5f" % tuple(L / 10))
    gro.write_text("\n".join(lines) + "\n")
    xtc = str(tmp_path / "traj.xtc")
    kept = trajio.write_xtc(xtc, xyz / np.float32(10.0), np.diag(L / 10) if box else None) * np.float32(10.0)
    out, npy = tmp_path / "o.csv", tmp_path / "o.npy"
    labels = {i: f"{res3[i % 5]}{i + 1}" for i in range(n_res)}
    sel = [(0, 9), (3, 39), (20, 21)]
    rc = extractor.main(["--top", str(gro), "--traj", xtc, "--pairs", ",".join(f"{labels[a]}-{labels[b]}" for a, b in sel),
                         "--stride", "2", "--out", str(out), "--npy-out", str(npy)])
    assert rc == 0 and "[DONE] Salvato CSV" in capsys.readouterr().out
    frames = kept[::2]
    pairs = [(3 * a + 1, 3 * b + 1) for a, b in sel]
    d = np.load(npy)
    if box is None:
        assert d.tobytes() == orc.pair_distances(frames, pairs, n_jobs=1).tobytes()
    else:
        want = orc.pbc_pair_distances(frames, pairs, np.broadcast_to(L, (len(frames), 3)))
        assert np.abs(d - want).max() <= 1e-4


def test_extractor_error_paths(tmp_path):
    from caviar_b200 import extractor
    top, nc, xyz, L, ca, labels = _system(tmp_path, None)
    with pytest.raises(SystemExit) as e:
        extractor.main(["--top", top, "--traj", nc, "--pairs", "bogus"])
    assert "Nessuna coppia valida" in str(e.value)
    with pytest.raises(SystemExit) as e:
        extractor.main(["--top", top, "--traj", nc, "--pairs", "ASP5-GLU30", "--offset", "21"])
    assert "label-offset < 1" in str(e.value)


def test_drop_in_reproduces_the_reference_main_byte_for_byte(tmp_path, capsys, golden_dir, monkeypatch):
    """B2 / a10: same command lines as tests/golden/extractor/cases.json (CSV, stdout produced by the UNMODIFIED
    reference main() with oracle/cpptraj_stub.py as cpptraj): the drop-in's CSV must be identical byte for byte --
    header, '#Frame' set index, 4 decimals, CR LF line ends -- and so must the two stdout lines."""
    import os
    import shutil
    from caviar_b200 import extractor
    gdir = os.path.join(golden_dir, "extractor")
    with open(os.path.join(gdir, "cases.json")) as fh:
        cases = json.load(fh)["cases"]
    work = tmp_path / "w"
    shutil.copytree(gdir, work)
    monkeypatch.chdir(work)
    for c in cases:
        out = f"dropin_{c['csv']}"
        argv = [out if a == c["csv"] else a for a in c["argv"]]
        assert extractor.main(argv) == 0
        printed = capsys.readouterr().out
        assert open(work / out, "rb").read() == open(os.path.join(gdir, c["csv"]), "rb").read(), c["name"]
        assert printed.replace(out, c["csv"]) == open(os.path.join(gdir, c["stdout"])).read(), c["name"]


def test_extractor_reads_amber_ascii_mdcrd(tmp_path, capsys, golden_dir):
    """The formats the reference's help text names (Extractor:99): an Amber ASCII trajectory (10F8.3, box line per frame)
    written here by hand goes through the drop-in; distances against the oracle on the 3-decimal coordinates."""
    import os
    from scipy.io import netcdf_file
    from caviar_b200 import extractor
    gdir = os.path.join(golden_dir, "extractor")
    nc = netcdf_file(os.path.join(gdir, "traj_ortho.nc"), "r", mmap=False)
    xyz = np.round(np.array(nc.variables["coordinates"][:], dtype=np.float64), 3)
    box = np.round(np.array(nc.variables["cell_lengths"][:], dtype=np.float64), 3)
    nc.close()
    text = "converted from traj_ortho.nc\n"
    for t in range(xyz.shape[0]):
        vals = xyz[t].ravel()
        for k in range(0, len(vals), 10):
            text += "".join("%8.3f" % v for v in vals[k:k + 10]) + "\n"
        text += "".join("%8.3f" % v for v in box[t]) + "\n"
    crd = tmp_path / "traj.mdcrd"
    crd.write_text(text)
    out = tmp_path / "out.csv"
    assert extractor.main(["--top", os.path.join(gdir, "sys.prmtop"), "--traj", str(crd), "--pairs", "ASP1-SER5,GLY4-LEU27",
                           "--out", str(out)]) == 0
    capsys.readouterr()
    with open(out, newline="") as fh:
        table = list(csv.reader(fh))
    assert table[0] == ["#Frame", "ASP1-SER5", "GLY4-LEU27"] and len(table) == 1 + xyz.shape[0]
    got = np.array([[float(v) for v in r[1:]] for r in table[1:]])
    ca = lambda r: 4 * r + 1
    want = orc.pbc_pair_distances(xyz.astype(np.float32), [(ca(0), ca(4)), (ca(3), ca(26))], box.astype(np.float32))
    assert np.abs(got - want).max() <= 1e-4 + 0.5e-4
