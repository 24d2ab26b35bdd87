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
