"""The drop-in, compiled and run: `oracle/_ref/chiml_ref --gpu` is the UNMODIFIED reference (every translation unit compiled in place by
oracle/Makefile) whose step() has been replaced by the bindGpu() / step() stub of INTEGRATION.md (oracle/ref_driver.cpp: the
reference's constructor builds its lists, the lists go straight to the C ABI of libchiml_b200.so, the time loop runs on the GPU, and
the reference's own dtc->output / toFile, flux->getFlux and dtcPop->toFile write the files).  Reference interface replaced:
FDTD_MANAGER/parallelFDTDField.hpp:1228-1303 (step), main.cpp:54-118 (loop and output).

Checked against (a) the committed state of the reference's own CPU run (tests/golden/<case>.expect.npz), bit for bit, and (b) the files the
same binary writes when it steps on the CPU, byte for byte (detector files print 18 significant digits, flux spectra are computed by
the reference's getFlux from the accumulators the GPU filled)."""
import filecmp
import os
import shutil
import subprocess

import numpy as np
import pytest

import util
from chiml_b200 import plan as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "chiml_ref")
pytestmark = pytest.mark.gpu

CASES = ["te_vacuum", "tm_au", "aniso_slab3d", "lorentz3d", "kappa3d", "ml3d_two", "ml_te", "c4_small", "tm_flux", "flux3d", "pbc3d", "pbc_tm", "pbc_te", "tfsf_tm", "tfsf_te",
         "tfsf3d", "tfsf3d_slab", "freq3d", "mag3d", "mag3d_pml", "mag_tm", "chi3d", "chi3d_pml", "dipnorm3d", "dipnorm3d_pml", "pbc_ml3d"]


def _run(case, workdir, gpu):
    os.makedirs(workdir, exist_ok=True)
    shutil.copy(os.path.join(util.GOLDEN, case + ".json"), os.path.join(workdir, case + ".json"))
    cmd = [REF, case + ".json", "--dump", "state.dump", "--quiet"] + (["--gpu"] if gpu else [])
    r = subprocess.run(cmd, cwd=workdir, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return {nm: arr for (rank, nm), (ln, ys, arr) in P.read_dump(os.path.join(workdir, "state.dump")).items()}


def _files(root):
    out = []
    for d, _, names in os.walk(root):
        for n in names:
            if n.endswith(".dump") or n.endswith(".json"):
                continue
            out.append(os.path.relpath(os.path.join(d, n), root))
    return sorted(out)


@pytest.mark.parametrize("case", CASES)
def test_reference_with_the_engine_behind_step_reproduces_its_own_cpu_run(case, tmp_path):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/chiml_ref is built where /root/reference is present (make -C oracle ref)")
    expect = util.load_expect(case)
    got = _run(case, str(tmp_path / "gpu"), True)
    for name, ref in expect.items():
        assert name in got, f"{case}: the drop-in run did not dump {name}"
        if "pop" in name:
            assert np.abs(got[name] - ref).max() <= 1e-9 * max(np.abs(ref).max(), 1e-300), f"{case}/{name}"
        else:
            assert np.array_equal(got[name], ref), f"{case}/{name}: max |diff| {np.abs(got[name] - ref).max():.3e}"
    # the files: the same binary stepping on the CPU
    _run(case, str(tmp_path / "cpu"), False)
    files = _files(str(tmp_path / "cpu"))
    assert files, f"{case}: the reference wrote no output file"
    assert _files(str(tmp_path / "gpu")) == files
    for f in files:
        a, b = str(tmp_path / "gpu" / f), str(tmp_path / "cpu" / f)
        if "level" in f or "qe_" in f:
            x, y = np.loadtxt(a), np.loadtxt(b)      # population = a sum over emitters: tree order on the GPU, detector tolerance 1e-9
            assert x.shape == y.shape and np.abs(x - y).max() <= 1e-9 * max(np.abs(y).max(), 1e-300), f
        else:
            assert filecmp.cmp(a, b, shallow=False), f"{case}: {f} differs between the GPU-backed and the CPU run of the reference"
