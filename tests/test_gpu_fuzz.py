"""GPU parity on RANDOM inputs (tests/fuzz/gen_inputs.py): the host-side setup turns a random case into a plan, the CUDA engine and
the CPU oracle step it from a seeded random state, and every field, psi, pole, emitter and flux-accumulator array must agree bit for
bit.  Random geometry puts object faces, edges and corners, CPML corners and transition layers at arbitrary positions inside the
64 x 8 tiles, which is what the tile classifier's rectangle records and the flag-specialised kernels have to get right."""
import os
import subprocess
import sys

import numpy as np
import pytest

import util
from chiml_b200 import capi
from oracle_api import OracleSim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "fuzz"))
TOOL = os.path.join(ROOT, "chiml_b200", "chiml_plan")
pytestmark = pytest.mark.gpu


def run_case(cfg, tmp_path, steps, march=None):
    from chiml_b200 import inputs as I, plan as P
    I.write(cfg, str(tmp_path / "c.json"))
    subprocess.run(["make", "-C", os.path.join(ROOT, "chiml_b200", "host"), os.path.join("..", "chiml_plan")], check=True, stdout=subprocess.DEVNULL)
    subprocess.run([TOOL, str(tmp_path / "c.json"), str(tmp_path / "p")], check=True)
    plan = P.read_plan(str(tmp_path / "p.rank0.plan"))
    rng = np.random.default_rng(4321)
    gpu, cpu = capi.GpuSim(plan, march=march), OracleSim(plan)
    lnx, lny, lnz = plan.ln
    for f in plan.fields_present():
        a = rng.uniform(-1.0, 1.0, size=(lny, lnz, lnx))
        gpu.set_field(f, a)
        cpu.field(f)[...] = a
    gpu.step_n(steps)
    cpu.step_n(steps)
    for name in util.state_names(plan):
        g, c = util.state_array(gpu, name), util.state_array(cpu, name)
        assert np.array_equal(g, c), f"{name}: max |diff| {np.abs(g - c).max():.3e}"
    for comp, part in [(c.comp, c.part) for c in plan.cpml if c.has_psi]:
        assert np.array_equal(gpu.psi(comp, part), cpu.psi(comp, part)), f"psi comp {comp} part {part}"
    gpu.close()
    cpu.close()


# forced column lengths of the y-marching kernels (tests/test_gpu_parity.py MARCH): every seed runs with the automatic choice (single
# planes on grids this small) and with one forced length, so that object faces / CPML seams meet column seams at random places
FORCED = [2, 3, 7, 1 << 20]


@pytest.mark.parametrize("forced", [False, True], ids=["auto", "march"])
@pytest.mark.parametrize("seed", [1, 2, 3, 4, 6, 8, 11, 12, 15, 18, 19, 26, 33, 50])
def test_gpu_matches_oracle_on_random_media_cells(seed, forced, tmp_path, oracle_lib):
    import gen_inputs
    run_case(gen_inputs.rnd_case(seed, steps=12, pulses="random"), tmp_path, 12, FORCED[seed % 4] if forced else None)


@pytest.mark.parametrize("forced", [False, True], ids=["auto", "march"])
@pytest.mark.parametrize("seed", [1, 8, 9, 10, 14, 15])
def test_gpu_matches_oracle_on_random_emitter_blocks(seed, forced, tmp_path, oracle_lib):
    import gen_inputs
    run_case(gen_inputs.rnd_ml_case(seed), tmp_path, 12, FORCED[seed % 4] if forced else None)


@pytest.mark.parametrize("forced", [False, True], ids=["auto", "march"])
@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5, 10, 12, 13, 14, 16, 22])
def test_gpu_matches_oracle_on_random_magnetic_and_chiral_media(seed, forced, tmp_path, oracle_lib):
    """tests/fuzz/gen_inputs.rnd_mag_case through this repository's own host setup: magnetic and chiral objects whose faces, edges and corners fall
    anywhere inside the tiles, some reaching through the CPML (the H-side CPML then acts on B)."""
    import gen_inputs
    run_case(gen_inputs.rnd_mag_case(seed, steps=12), tmp_path, 12, FORCED[seed % 4] if forced else None)


@pytest.mark.parametrize("forced", [False, True], ids=["auto", "march"])
@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5, 6, 9, 11, 13])
def test_gpu_matches_oracle_on_random_periodic_inputs(seed, forced, tmp_path, oracle_lib):
    """tests/fuzz/gen_inputs.rnd_pbc_case on one slab: k_wrap between the half steps (and inside the one-launch 2-D kernel), from a random state."""
    import gen_inputs
    run_case(gen_inputs.rnd_pbc_case(seed), tmp_path, 12, FORCED[seed % 4] if forced else None)


@pytest.mark.parametrize("forced", [False, True], ids=["auto", "march"])
@pytest.mark.parametrize("seed", [0, 1, 2, 4, 5, 6, 9, 10, 11, 14])
def test_gpu_matches_oracle_on_random_surface_normal_dipoles(seed, forced, tmp_path, oracle_lib):
    """tests/fuzz/gen_inputs.rnd_dipnorm_case through the host setup's own dipole grids: k_ordip_poles<true> reads the dipole vector of every node."""
    import gen_inputs
    run_case(gen_inputs.rnd_dipnorm_case(seed, steps=12), tmp_path, 12, FORCED[seed % 4] if forced else None)
