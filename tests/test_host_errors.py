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
    (lambda c: c["ObjectList"].append(I.block([0.1, 0.1, 0.0], [0, 0, 0], pols=[I.lorentz_pole(0.5, 0.1, 2.0, dip_or_e="normal")])), "surface-normal-relative"),
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


def test_periodic_run_on_several_slabs_is_refused(tmp_path):
    """Real-field periodic boundaries are covered on one slab; the reference's multi-rank periodic run (applyBCProcMid on every rank,
    SURVEY.md appendix B.5) is not reproduced, so the setup must say so instead of building slab plans."""
    import shutil
    import util
    shutil.copy(os.path.join(util.GOLDEN, "pbc3d.json"), tmp_path / "pbc3d.json")
    r = subprocess.run([TOOL, "pbc3d.json", "out", "--ranks", "2"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0 and "single-slab" in r.stderr, r.stderr
