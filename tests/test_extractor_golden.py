"""The Extractor boundary (B2 / a10) pinned by the reference's OWN main(): tests/golden/extractor/ holds the CSV files,
cpptraj scripts and stdout the unmodified CAVIAR_Distances-Extractor.py produced for four command lines, with
oracle/cpptraj_stub.py standing in for the cpptraj binary (tests/golden/make_extractor_golden.py).

CPU tests here: the goldens are reproducible from /root/reference when it is present, the stub accepts exactly the
script text the reference generates, and everything of the drop-in that needs no GPU -- topology / trajectory readers
on these foreign-written files, label -> atom selection, header, row formatting with CR LF -- reproduces the golden
bytes when fed the oracle's distances.  The GPU test (tests/test_gpu_extractor.py) runs the drop-in's main() itself.
"""
import json
import os
import shutil
import stat
import subprocess
import sys

import numpy as np
import pytest

from oracle import caviar_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/CAVIAR_Distances-Extractor.py"


@pytest.fixture(scope="module")
def gdir(golden_dir):
    return os.path.join(golden_dir, "extractor")


@pytest.fixture(scope="module")
def cases(gdir):
    with open(os.path.join(gdir, "cases.json")) as fh:
        return json.load(fh)["cases"]


def _argval(argv, flag, default=None):
    return argv[argv.index(flag) + 1] if flag in argv else default


def test_golden_csv_layout_is_the_reference_writers(gdir, cases):
    """csv.writer output (Extractor:196-198): every line CR LF terminated, '#Frame' kept, labels in the column titles."""
    assert len(cases) == 4
    for c in cases:
        raw = open(os.path.join(gdir, c["csv"]), "rb").read()
        lines = raw.split(b"\r\n")
        assert lines[-1] == b"" and b"\n" not in raw.replace(b"\r\n", b""), c["name"]
        header = lines[0].decode().split(",")
        assert header[0] == "#Frame"
        info = open(os.path.join(gdir, c["stdout"])).read().splitlines()
        assert info[0] == f"[DONE] Salvato CSV: {c['csv']}"
        assert info[1] == "[INFO] Colonne: " + ", ".join(header[1:])
        stride = int(_argval(c["argv"], "--stride", "1"))
        assert len(lines) - 2 == len(range(0, 18, stride))
        assert [ln.split(b",")[0] for ln in lines[1:-1]] == [str(k + 1).encode() for k in range(len(lines) - 2)]


@pytest.mark.skipif(not os.path.exists(REF), reason="the reference is only present in the build container")
def test_goldens_are_reproducible_from_the_reference(gdir, tmp_path):
    out = tmp_path / "regen"
    subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "make_extractor_golden.py"), str(out)], check=True,
                   stdout=subprocess.DEVNULL)
    names = sorted(os.listdir(gdir))
    assert names == sorted(os.listdir(out))
    for n in names:
        a, b = open(os.path.join(gdir, n), "rb").read(), open(out / n, "rb").read()
        assert a == b, n


def _run_stub(tmp_path, script_text, cwd):
    p = tmp_path / "in.cpptraj"
    p.write_text(script_text)
    return subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "cpptraj_stub.py"), "-i", str(p)], cwd=cwd,
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


def test_stub_accepts_exactly_the_reference_script_text(gdir, cases, tmp_path):
    work = tmp_path / "w"
    shutil.copytree(gdir, work)
    c = cases[0]
    text = open(os.path.join(gdir, c["script"])).read()
    res = _run_stub(tmp_path, text, str(work))
    assert res.returncode == 0, res.stderr
    got = open(work / "cpptraj_tmp_dist.csv").read().splitlines()
    gold = open(os.path.join(gdir, c["csv"]), newline="").read().split("\r\n")
    assert got[0].startswith("#Frame,d0001,") and got[1:] == gold[1:-1]      # the reference copied the rows verbatim
    for bad in (text.replace("1 last", "1 last stride 2"), text.replace(":1@CA", ":1@CB"), text.replace("run\n", ""),
                text.replace("delim comma", ""), text.replace("d0002 ", "", 1).replace("distance  ", "distance d0002 "),
                text + "quit\n", text.replace('parm "', 'parm ')):
        assert bad != text
        assert _run_stub(tmp_path, bad, str(work)).returncode != 0, bad


def test_drop_in_host_side_reproduces_the_golden_bytes(gdir, cases):
    """trajio + select_atoms + header + format_rows over the ORACLE's distances == the reference's CSV, byte for byte
    (what remains for the GPU test is that the kernels return those distances to within half a unit of the 4th decimal)."""
    import __graft_entry__ as g
    g.build()
    from caviar_b200 import extractor, trajio
    for c in cases:
        argv = c["argv"]
        top = trajio.read_topology(os.path.join(gdir, _argval(argv, "--top")))
        traj = trajio.Trajectory(os.path.join(gdir, _argval(argv, "--traj")))
        pf = _argval(argv, "--pairs-file")
        labels = extractor.read_pairs(_argval(argv, "--pairs", ""), os.path.join(gdir, pf) if pf else "")
        atoms = extractor.select_atoms(top, labels, _argval(argv, "--numbering", "sequential"), int(_argval(argv, "--offset", "0")))
        stride = int(_argval(argv, "--stride", "1"))
        xyz = traj.coords(0, traj.n_frames, stride)
        cell = traj.cell(0, traj.n_frames, stride)
        assert (cell is None) == ("nobox" in c["name"])
        if "tric" in c["name"]:
            assert traj.box_kind() == "triclinic" and cell.shape[1:] == (3, 3)
        # the unlimited-dimension record layout interleaves coordinates with time and cell: the native compaction must
        # stride over whole records
        sel = np.unique(atoms.ravel())
        assert np.array_equal(traj.coords_of(sel, 0, traj.n_frames, stride), xyz[:, sel, :])
        traj.close()
        d = orc.pbc_pair_distances(xyz, atoms, cell).astype(np.float32)
        head = (",".join(["#Frame"] + [f"{ra}{na}-{rb}{nb}" for ra, na, rb, nb in labels]) + "\r\n").encode()
        got = head + extractor.format_rows(d, 1, 1, 4)
        assert got == open(os.path.join(gdir, c["csv"]), "rb").read(), c["name"]
