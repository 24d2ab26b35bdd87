"""The host-side setup refuses what the reference refuses, with the reference's messages (INPUTS/parallelInputs.cpp:98-105,
:196-216, :712-716), and refuses loudly -- instead of computing something else -- what lies outside the covered hot path."""
import json
import os
import subprocess

import pytest

from chiml_b200 import inputs as I

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "chiml_b200", "chiml_plan")


def _base():
    return I.c1_te_vacuum(n=47, steps=10, pml_cells=8)


def _run(cfg, tmp_path):
    subprocess.run(["make", "-C", os.path.join(ROOT, "chiml_b200", "host"), os.path.join("..", "chiml_plan")], check=True, stdout=subprocess.DEVNULL)
    p = tmp_path / "in.json"
    p.write_text(json.dumps(cfg))
    return subprocess.run([TOOL, str(p), str(tmp_path / "out")], capture_output=True, text=True)


def test_courant_limit(tmp_path):
    cfg = _base()
    cfg["CompCell"]["courant"] = 1.5
    r = _run(cfg, tmp_path)
    assert r.returncode != 0 and "Time step is larger than the stable time step" in r.stderr


def test_pml_thicker_than_cell(tmp_path):
    cfg = _base()
    cfg["PML"]["thickness"] = [0.3, 0.3, 0.0]
    r = _run(cfg, tmp_path)
    assert r.returncode != 0 and "PML size is larger than the cell size" in r.stderr


def test_source_outside_cell(tmp_path):
    cfg = _base()
    cfg["SourceList"][0]["loc"] = [5.0, 0.0, 0.0]
    r = _run(cfg, tmp_path)
    assert r.returncode != 0 and "outside the FDTD Cell" in r.stderr


def test_detector_outside_cell(tmp_path):
    cfg = _base()
    cfg["DetectorList"][0]["loc"] = [0.0, 7.0, 0.0]
    r = _run(cfg, tmp_path)
    assert r.returncode != 0 and "detector is outside the FDTD cell" in r.stderr


@pytest.mark.parametrize("mutate,needle", [
    (lambda c: c.__setitem__("TFSF", [{"dummy": 1}]), "TFSF sources are outside the covered hot path"),
    (lambda c: c["CompCell"].__setitem__("cplxFields", True), "complex fields without periodic boundaries"),
    (lambda c: c["ObjectList"].append(I.block([0.1, 0.1, 0.0], [0, 0, 0], pols=[I.lorentz_pole(0.5, 0.1, 2.0, sigma_m=0.4, dip_or_m="unidirectional")])), "oriented magnetic"),
    (lambda c: c["ObjectList"].append(I.block([0.1, 0.1, 0.0], [0, 0, 0], pols=[I.lorentz_pole(0.5, 0.1, 2.0, dip_or_e="normal")])), "2-D grid"),
    (lambda c: c["ObjectList"].append(I.block([0.1, 0.1, 0.0], [0, 0, 0], pols=[I.lorentz_pole(0.5, 0.1, 2.0, dip_or_e="lat_tangent")])), "lat_tangent"),
])
def test_out_of_scope_inputs_fail_loudly(mutate, needle, tmp_path):
    cfg = _base()
    mutate(cfg)
    r = _run(cfg, tmp_path)
    assert r.returncode != 0 and needle in r.stderr, r.stderr


def test_valid_input_builds(tmp_path):
    r = _run(_base(), tmp_path)
    assert r.returncode == 0, r.stderr
    assert os.path.exists(tmp_path / "out.rank0.plan")


def test_periodic_slab_plans_carry_the_ring_wraps(tmp_path):
    """A periodic run on several slabs: every slab's PERIODIC records describe the x / z wraps of its owned rows only (ymax = ny = -1; the y
    direction is the ring of ghost-row pushes, chiml_b200/slab.py); Bloch-periodic (complex) runs on several slabs are refused."""
    import json
    import shutil
    import sys
    import util
    sys.path.insert(0, ROOT)
    from chiml_b200 import plan as P
    shutil.copy(os.path.join(util.GOLDEN, "pbc3d.json"), tmp_path / "pbc3d.json")
    r = subprocess.run([TOOL, "pbc3d.json", "out", "--ranks", "3"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    whole = util.load_plan("pbc3d")
    for rank in range(3):
        pl = P.read_plan(str(tmp_path / f"out.rank{rank}.plan"))
        assert sorted(pl.periodic) == sorted(whole.periodic)
        for comp, w in pl.periodic.items():
            ref = whole.periodic[comp]
            assert w[1] == -1 and w[4] == -1 and (w[0], w[2], w[3], w[5], w[6]) == (ref[0], ref[2], ref[3], ref[5], ref[6])
    cfg = json.load(open(os.path.join(util.GOLDEN, "cplx3d.json")))
    json.dump(cfg, open(tmp_path / "cplx3d.json", "w"))
    r = subprocess.run([TOOL, "cplx3d.json", "out2", "--ranks", "2"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0 and "several slabs" in r.stderr, r.stderr
