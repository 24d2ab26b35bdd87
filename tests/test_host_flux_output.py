"""Flux spectra files (the reference's parallelFluxDTC::getFlux, restated in chiml_b200/host/flux_out.cpp): fed with the
reference's own running-DFT accumulators (tests/golden/<case>.expect.npz, dft<k>r / dft<k>i), the host-side post-processing must
write, character for character, the <flux name>.dat files the unmodified reference wrote for the same input
(tests/golden/out_expected/<case>/, made by tests/golden/make_flux_expected.py).  No GPU involved: this pins the surface averaging,
the Poynting integrand, the Simpson integration and the formatting; tests/test_gpu_host_driver.py pins the whole chain."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def write_dft_files(case, workdir):
    """Accumulator files in the driver's format (chiml_b200/host/main.cpp) from the reference dump."""
    cfg = json.load(open(os.path.join(util.GOLDEN, case + ".json")))
    plan = util.load_plan(case)
    exp = util.load_expect(case)
    for g, fl in enumerate(cfg["FluxList"]):
        sets = [(k, d) for k, d in enumerate(plan.dfts) if d.group == g]
        path = os.path.join(workdir, fl["name"] + ".dft")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "wb") as f:
            freq = np.asarray(sets[0][1].freq, "<f8")
            f.write(b"CHIMLDFT" + struct.pack("<ii", len(sets), len(freq)) + freq.tobytes())
            for k, d in sets:
                re, im = exp[f"dft{k}r"].ravel(), exp[f"dft{k}i"].ravel()
                f.write(struct.pack("<iiii", d.field, d.npts, len(d.lines), d.every) + struct.pack("<Q", len(re)))
                f.write(np.asarray(re, "<f8").tobytes() + np.asarray(im, "<f8").tobytes())
    return [fl["name"] for fl in cfg["FluxList"]]


@pytest.mark.parametrize("case", ["tm_flux", "te_flux", "flux3d"])
def test_flux_files_from_reference_accumulators_equal_reference_files(case, tmp_path):
    subprocess.run(["make", "-C", os.path.join(ROOT, "chiml_b200", "host"), os.path.join("..", "chiml_flux")], check=True, stdout=subprocess.DEVNULL)
    names = write_dft_files(case, str(tmp_path))
    r = subprocess.run([os.path.join(ROOT, "chiml_b200", "chiml_flux"), os.path.join(util.GOLDEN, case + ".json")], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for name in names:
        got = open(tmp_path / (name + ".dat"), "rb").read()
        ref = open(os.path.join(util.GOLDEN, "out_expected", case, os.path.basename(name) + ".dat"), "rb").read()
        assert got == ref, f"{case}: {name}.dat differs from the reference's file"


@pytest.mark.parametrize("case,nranks", [("flux3d", 3), ("te_flux", 2), ("tm_flux", 4)])
def test_flux_files_from_slab_accumulator_parts(case, nranks, tmp_path):
    """Several slabs: every rank's driver writes the accumulators of ITS parts of the flux surfaces (<name>.dft.rank<r>);
    `chiml_flux --ranks N` puts the parts together (an accumulator is identified by region, surface, stored field and global grid
    point) and must write the same files as a single-rank run.  The parts are cut here from the reference's single-rank accumulators
    through the slab plans of the host-side setup -- that the engines produce exactly these parts is what tests/test_slab_gloo.py
    and tests/test_gpu_slabs.py check."""
    from chiml_b200 import plan as P
    host = os.path.join(ROOT, "chiml_b200", "host")
    subprocess.run(["make", "-C", host, os.path.join("..", "chiml_flux"), os.path.join("..", "chiml_plan")], check=True, stdout=subprocess.DEVNULL)
    cfg = json.load(open(os.path.join(util.GOLDEN, case + ".json")))
    whole = util.load_plan(case)
    exp = util.load_expect(case)
    ref = util.dft_point_map(whole, [exp[f"dft{k}r"].ravel() + 1j * exp[f"dft{k}i"].ravel() for k in range(len(whole.dfts))])
    subprocess.run([os.path.join(ROOT, "chiml_b200", "chiml_plan"), os.path.join(util.GOLDEN, case + ".json"), str(tmp_path / "p"), "--ranks", str(nranks)], check=True)
    for r in range(nranks):
        slab = P.read_plan(str(tmp_path / f"p.rank{r}.plan"))
        lnx, lny, lnz = slab.ln
        for g, fl in enumerate(cfg["FluxList"]):
            sets = [d for d in slab.dfts if d.group == g]
            if not sets:
                continue
            path = str(tmp_path / (fl["name"] + f".dft.rank{r}"))
            os.makedirs(os.path.dirname(path), exist_ok=True)
            with open(path, "wb") as f:
                freq = np.asarray(sets[0].freq, "<f8")
                f.write(b"CHIMLDFT" + struct.pack("<ii", len(sets), len(freq)) + freq.tobytes())
                for d in sets:
                    acc = np.zeros(d.acc_len, dtype=complex)
                    for li, (ind, o) in enumerate(d.lines):
                        if li > 0 and ind == 0 and o == 0:
                            continue
                        for i in range(d.npts):
                            row, x = divmod(int(ind) + i * d.stride, lnx)
                            y, z = divmod(row, lnz)
                            for k in range(d.nfreq):
                                acc[int(o) + k + d.nfreq * i] = ref[(d.group, d.field, x, y + slab.y_start, z, k)]
                    f.write(struct.pack("<iiii", d.field, d.npts, len(d.lines), d.every) + struct.pack("<Q", d.acc_len))
                    f.write(np.asarray(acc.real, "<f8").tobytes() + np.asarray(acc.imag, "<f8").tobytes())
    r = subprocess.run([os.path.join(ROOT, "chiml_b200", "chiml_flux"), os.path.join(util.GOLDEN, case + ".json"), "--ranks", str(nranks)],
                       cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for fl in cfg["FluxList"]:
        got = open(tmp_path / (fl["name"] + ".dat"), "rb").read()
        want = open(os.path.join(util.GOLDEN, "out_expected", case, os.path.basename(fl["name"]) + ".dat"), "rb").read()
        assert got == want, f"{case}: {fl['name']}.dat from {nranks} slabs differs from the reference's file"


def _flux_tool(tmp_path, json_path, case="flux3d", extra=()):
    subprocess.run(["make", "-C", os.path.join(ROOT, "chiml_b200", "host"), os.path.join("..", "chiml_flux")], check=True, stdout=subprocess.DEVNULL)
    write_dft_files(case, str(tmp_path))
    return subprocess.run([os.path.join(ROOT, "chiml_b200", "chiml_flux"), json_path, *extra], cwd=tmp_path, capture_output=True, text=True)


def test_saved_surface_fields_equal_the_reference_file(tmp_path):
    """"save": true on a 3-D flux box: <name>_fields.dat (parallelFluxDTC::saveFields, DTC/parallelFlux.hpp:616-659) byte for byte."""
    d = os.path.join(util.GOLDEN, "out_expected", "flux3d_save")
    r = _flux_tool(tmp_path, os.path.join(d, "flux3d_save.json"))
    assert r.returncode == 0, r.stderr
    assert open(tmp_path / "out/f3/box_fields.dat", "rb").read() == open(os.path.join(d, "box_fields.dat"), "rb").read()


def test_loaded_incident_fields_are_subtracted_like_the_reference(tmp_path):
    """"load": true: the surface fields start from minus the fields an earlier run saved (loadFields(-1.0), :664-722; here the
    reference's own save of the cell without the scatterer) and the spectrum file equals the reference's for that input."""
    import shutil
    d = os.path.join(util.GOLDEN, "out_expected", "flux3d_load")
    shutil.copy(os.path.join(d, "empty_fields.dat"), tmp_path)
    r = _flux_tool(tmp_path, os.path.join(d, "flux3d_load.json"))
    assert r.returncode == 0, r.stderr
    assert open(tmp_path / "out/f3/box.dat", "rb").read() == open(os.path.join(d, "box.dat"), "rb").read()
    # the other regions of the input load nothing and stay as in the plain run
    assert open(tmp_path / "out/f3/px.dat", "rb").read() == open(os.path.join(util.GOLDEN, "out_expected", "flux3d", "px.dat"), "rb").read()


def test_load_refuses_fields_of_another_region(tmp_path):
    import shutil
    d = os.path.join(util.GOLDEN, "out_expected", "flux3d_load")
    cfg = json.load(open(os.path.join(d, "flux3d_load.json")))
    cfg["FluxList"][1]["load"] = True
    cfg["FluxList"][1]["incd_fileds"] = "empty_fields.dat"          # a 3-frequency box file for the 4-frequency plane
    json.dump(cfg, open(tmp_path / "bad.json", "w"))
    shutil.copy(os.path.join(d, "empty_fields.dat"), tmp_path)
    r = _flux_tool(tmp_path, str(tmp_path / "bad.json"))
    assert r.returncode != 0 and "do not match size and frequency" in (r.stderr + r.stdout)


def test_frequency_detector_files_equal_the_reference_files(tmp_path):
    """dtc_class "freq" detectors (DTC/parallelDTC_FREQ.hpp): a field type over a box, an SI-scaled E-power type over three stored fields
    sampled every second step, and a map output (one file per frequency) -- from the reference's accumulators, every file must equal the
    reference's character for character; the flux box of the same input is written beside them."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "chiml_b200", "host"), os.path.join("..", "chiml_flux")], check=True, stdout=subprocess.DEVNULL)
    case = "freq3d"
    cfg = json.load(open(os.path.join(util.GOLDEN, case + ".json")))
    plan = util.load_plan(case)
    exp = util.load_expect(case)
    freq_names = [d["fname"] + f"_field_{i}.dat" for i, d in enumerate(cfg["DetectorList"]) if d["dtc_class"] == "freq"]
    names = [fl["name"] for fl in cfg["FluxList"]] + freq_names
    for g, name in enumerate(names):
        sets = [(k, d) for k, d in enumerate(plan.dfts) if d.group == g]
        path = os.path.join(tmp_path, name + ".dft")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "wb") as f:
            freq = np.asarray(sets[0][1].freq, "<f8")
            f.write(b"CHIMLDFT" + struct.pack("<ii", len(sets), len(freq)) + freq.tobytes())
            for k, d in sets:
                re, im = exp[f"dft{k}r"].ravel(), exp[f"dft{k}i"].ravel()
                f.write(struct.pack("<iiii", d.field, d.npts, len(d.lines), d.every) + struct.pack("<Q", len(re)))
                f.write(np.asarray(re, "<f8").tobytes() + np.asarray(im, "<f8").tobytes())
    r = subprocess.run([os.path.join(ROOT, "chiml_b200", "chiml_flux"), os.path.join(util.GOLDEN, case + ".json")], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    want = sorted(os.listdir(os.path.join(util.GOLDEN, "out_expected", case)))
    assert len(want) >= 5
    for n in want:
        got = open(tmp_path / "out" / "fq" / n, "rb").read()
        ref = open(os.path.join(util.GOLDEN, "out_expected", case, n), "rb").read()
        assert got == ref, f"{n} differs from the reference's file"


def test_incident_normalisation_of_a_tfsf_run_equals_the_reference_file(tmp_path):
    """getFlux with incident fields (DTC/parallelFlux.hpp:455-482): the incident-field series of the TFSF source, Fourier transformed and
    scaled by the area of the region's faces, fill the abs(incd) / real(incd) / imag(incd) columns.  Series and accumulators are the
    reference's own (tests/golden/out_expected/tfsf3d/incd.bin, tfsf3d.expect.npz); the file must equal the reference's."""
    d = os.path.join(util.GOLDEN, "out_expected", "tfsf3d")
    r = _flux_tool(tmp_path, os.path.join(util.GOLDEN, "tfsf3d.json"), case="tfsf3d", extra=("--incd", os.path.join(d, "incd.bin")))
    assert r.returncode == 0, r.stderr
    got = open(tmp_path / "out/t3/box.dat", "rb").read()
    assert got == open(os.path.join(d, "box.dat"), "rb").read()
    assert float(got.split(b"\n")[1].split()[1]) > 0.0          # the incident column is not the all-zero one of a run without TFSF
