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
         # periodic boundaries (CompCell.PBC, real fields): JSON -> wrap descriptions -> k_wrap between the half steps
         "pbc3d": {"out/p3/dtc_field_0.dat": "dtc_field_0.dat"}, "pbc_tm": {"out/ptm/dtc_field_0.dat": "dtc_field_0.dat"},
         # frequency detectors (field, SI power over three fields, map output) beside a flux box: DFT sets on the GPU, files by the host
         "freq3d": {"out/fq/ez_field_1.dat": "ez_field_1.dat", "out/fq/epow_field_2.dat": "epow_field_2.dat", "out/fq/box.dat": "box.dat",
                    "out/fq/map_field_3.dat.1.000000": "map_field_3.dat.1.000000", "out/fq/map_field_3.dat.2.000000": "map_field_3.dat.2.000000"},
         # Bloch-periodic runs (k-point != 0, complex fields): two real field sets on the device, coupled by k_wrap_bloch
         "cplx3d": {"out/c3/dtc_field_0.dat": "dtc_field_0.dat"}, "cplx_tm": {"out/ctm/dtc_field_0.dat": "dtc_field_0.dat"},
         # magnetic-dispersive and chiral media set up by this repository's own host code (B grids, upB_ / upLorB_ / upChiD_ / upChiB_, copy2PrevFields_)
         "mag3d": {"out/mg/dtc_field_0.dat": "dtc_field_0.dat"}, "mag3d_pml": {"out/mgp/dtc_field_0.dat": "dtc_field_0.dat"},
         "mag_tm": {"out/mtm/dtc_field_0.dat": "dtc_field_0.dat"},
         "chi3d": {"out/ch/dtc_field_0.dat": "dtc_field_0.dat"}, "chi3d_pml": {"out/chp/dtc_field_0.dat": "dtc_field_0.dat"},
         # dipoles oriented relative to the surface normal: the host's own findGradient / dipole grids
         "dipnorm3d": {"out/dn/dtc_field_0.dat": "dtc_field_0.dat"}, "dipnorm3d_pml": {"out/dnp/dtc_field_0.dat": "dtc_field_0.dat"},
         "vac3d_bin": {"out/vb/dtc_field_0.dat": "dtc_field_0.dat", "out/vb/dtc_field_1.dat": "dtc_field_1.dat"}}


def test_power_and_console_detectors(tmp_path):
    """H-power detector (|Hz|^2 with the squared SI factor, DTC/parallelDTC.hpp:87-91) written as TXT, and a console (COUT) detector
    whose lines must equal the ones the reference printed (tests/golden/make_out_expected.py)."""
    import sys
    sys.path.insert(0, GOLDEN)
    from make_out_expected import cout_lines
    exe = os.path.join(ROOT, "chiml_b200", "chiml")
    shutil.copy(os.path.join(GOLDEN, "out_expected", "te_hpow", "te_hpow.json"), tmp_path / "te_hpow.json")
    r = subprocess.run([exe, "te_hpow.json"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    got = open(tmp_path / "out" / "hp" / "pow_field_0.dat", "rb").read()
    assert got == open(os.path.join(GOLDEN, "out_expected", "te_hpow", "pow_field_0.dat"), "rb").read()
    assert b"e-" in got and got.count(b"\n") == 32                      # 60 steps, every second one, + t = 0, + header
    assert cout_lines(r.stdout) == open(os.path.join(GOLDEN, "out_expected", "te_hpow", "cout.txt")).read()


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


def _read_dft_file(path):
    """chiml_b200/host/main.cpp: "CHIMLDFT", int32 nsets, int32 nfreq, freq[nfreq], then per stored field int32 field, npts, nlines, every,
    uint64 len, re[len], im[len]."""
    raw = open(path, "rb").read()
    assert raw[:8] == b"CHIMLDFT"
    nsets, nfreq = np.frombuffer(raw, "<i4", 2, 8)
    pos = 16 + 8 * int(nfreq)
    sets = []
    for _ in range(int(nsets)):
        n = int(np.frombuffer(raw, "<u8", 1, pos + 16)[0])
        re = np.frombuffer(raw, "<f8", n, pos + 24)
        im = np.frombuffer(raw, "<f8", n, pos + 24 + 8 * n)
        sets.append((re, im))
        pos += 24 + 16 * n
    assert pos == len(raw)
    return sets


@pytest.mark.parametrize("case,regions", [("flux3d", ["out/f3/box", "out/f3/px", "out/f3/py", "out/f3/pz"]),
                                          ("te_flux", ["out/tef/box", "out/tef/lx", "out/tef/ly"]), ("tm_flux", None)])
def test_host_driver_flux_accumulators_equal_the_reference(case, regions, tmp_path):
    """JSON -> flux surfaces -> running DFT on the GPU -> files: every accumulator of every stored field of every flux region equals,
    bit for bit, fInReal_ / fInCplx_ of the reference's parallelStorageFreqDTCReal objects after the same run (<case>.expect.npz,
    arrays dft<slot>r / dft<slot>i in the order parallelFluxDTC::fieldIn walks them)."""
    import json
    exe = os.path.join(ROOT, "chiml_b200", "chiml")
    assert os.path.exists(exe), "build it with make -C chiml_b200/host"
    shutil.copy(os.path.join(GOLDEN, case + ".json"), tmp_path / (case + ".json"))
    if regions is None:
        regions = [f["name"] for f in json.load(open(os.path.join(GOLDEN, case + ".json")))["FluxList"]]
    r = subprocess.run([exe, case + ".json"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    expect = np.load(os.path.join(GOLDEN, case + ".expect.npz"))
    slot = 0
    for name in regions:
        for re, im in _read_dft_file(tmp_path / (name + ".dft")):
            ref_re, ref_im = expect[f"dft{slot}r"].ravel(), expect[f"dft{slot}i"].ravel()
            assert np.abs(ref_re).max() > 0, f"{case}: accumulator {slot} of the reference is all zero"
            assert np.array_equal(re, ref_re) and np.array_equal(im, ref_im), f"{case}: {name} stored field {slot} differs from the reference"
            slot += 1
        # ... and the flux spectrum the driver derives from them is the reference's <flux name>.dat, character for character
        got = open(tmp_path / (name + ".dat"), "rb").read()
        ref = open(os.path.join(GOLDEN, "out_expected", case, os.path.basename(name) + ".dat"), "rb").read()
        assert got == ref, f"{case}: {name}.dat differs from the reference's file"
    assert f"dft{slot}r" not in expect.files and slot > 0
