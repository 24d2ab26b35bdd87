"""The host driver executable (chiml_b200/chiml, the counterpart of the reference's main.cpp): JSON in, detector / population
files out, everything between through the C ABI on the GPU.  Its files must equal the files the unmodified reference wrote for the
same input (tests/golden/out_expected, produced by oracle/_ref/chiml_ref): detector files character for character (18 significant
digits are printed, so this is bit equality), population files to the detector tolerance of 1e-9."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
pytestmark = pytest.mark.gpu

FILES = {"te_vacuum": {"out/te/dtc_field_0.dat": "dtc_field_0.dat"},
         "ml3d_two": {"out/m2/dtc_field_0.dat": "dtc_field_0.dat", "output_data/qe_0_level_3.dat": "qe_0_level_3.dat",
                      "output_data/qe_0_level_1.dat": "qe_0_level_1.dat"},
         # a 4 x 3 x 3 box written by a BIN detector (DTC/parallelDTC_BIN.cpp) and an SI-scaled TXT detector
         "vac3d_bin": {"out/vb/dtc_field_0.dat": "dtc_field_0.dat", "out/vb/dtc_field_1.dat": "dtc_field_1.dat"}}


@pytest.mark.parametrize("case", sorted(FILES))
def test_host_driver_writes_the_reference_files(case, tmp_path):
    exe = os.path.join(ROOT, "chiml_b200", "chiml")
    assert os.path.exists(exe), "build it with make -C chiml_b200/host"
    src = os.path.join(GOLDEN, case + ".json")
    if not os.path.exists(src):
        src = os.path.join(GOLDEN, "out_expected", case, case + ".json")
    shutil.copy(src, tmp_path / (case + ".json"))
    r = subprocess.run([exe, case + ".json"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    for produced, expected in FILES[case].items():
        got = open(tmp_path / produced, "rb").read()
        ref = open(os.path.join(GOLDEN, "out_expected", case, expected), "rb").read()
        if "level" in produced:
            a, b = np.loadtxt(tmp_path / produced), np.loadtxt(os.path.join(GOLDEN, "out_expected", case, expected))
            assert a.shape == b.shape
            assert np.abs(a - b).max() <= 1e-9 * max(np.abs(b).max(), 1e-300)
        else:
            assert got == ref, f"{produced} differs from the reference's file"
