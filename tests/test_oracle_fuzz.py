"""The restated CPU oracle against the UNMODIFIED reference on RANDOM inputs (tests/fuzz/gen_inputs.py): oracle/_ref/chiml_ref steps a
random case and dumps every state array (fields, D, pole and node-pole grids, CPML-free emitter states, flux accumulators); the oracle,
fed with the reference constructor's own lists, must reproduce them bit for bit.  Together with tests/test_oracle_vs_ref.py (the
committed fixtures) this is the pin of the checker the GPU parity tests rely on.  Needs the reference build; skipped without it."""
import os
import subprocess
import sys

import numpy as np
import pytest

import util
from oracle_api import OracleSim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "fuzz"))
REF = os.path.join(ROOT, "oracle", "_ref", "chiml_ref")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="needs the reference build oracle/_ref/chiml_ref")


def run_case(cfg, tmp_path, min_nonzero):
    from chiml_b200 import inputs as I, plan as P
    I.write(cfg, str(tmp_path / "c.json"))
    r = subprocess.run([REF, "c.json", "--plan", str(tmp_path / "ref"), "--dump", str(tmp_path / "ref.dump"), "--quiet", "--no-output"],
                       cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    plan = P.read_plan(str(tmp_path / "ref.rank0.plan"))
    expect = {nm: arr for (rank, nm), (ln, ys, arr) in P.read_dump(str(tmp_path / "ref.dump")).items() if rank == 0}
    sim = OracleSim(plan)
    sim.step_n(plan.n_steps)
    nonzero = 0
    for name, ref in expect.items():
        got = util.state_array(sim, name)
        assert got is not None, f"oracle has no array {name}"
        nonzero += int(np.abs(ref).max() > 0)
        assert np.array_equal(got, ref), f"{name}: max |diff| = {np.abs(got - ref).max():.3e}"
    sim.close()
    assert nonzero >= min_nonzero, "the random case does not exercise the state arrays"


@pytest.mark.parametrize("seed", [1, 2, 4, 6, 8, 12, 15, 18, 19, 26])
def test_oracle_matches_reference_on_random_media_cells(seed, tmp_path, oracle_lib):
    import gen_inputs
    run_case(gen_inputs.rnd_case(seed, steps=12, pulses="random"), tmp_path, 5)


@pytest.mark.parametrize("seed", [1, 8, 9, 10, 14, 15])
def test_oracle_matches_reference_on_random_emitter_blocks(seed, tmp_path, oracle_lib):
    import gen_inputs
    cfg = gen_inputs.rnd_ml_case(seed)
    cfg["CompCell"]["tLim"] = 12 * gen_inputs.DT - 0.5 * gen_inputs.DT
    run_case(cfg, tmp_path, 10)


@pytest.mark.parametrize("seed", [0, 1, 3, 4, 5, 10, 12, 14, 16, 22, 25, 26])      # (seeds whose source is not overwritten by D->E / B->H of an object around it)
def test_oracle_matches_reference_on_random_magnetic_and_chiral_media(seed, tmp_path, oracle_lib):
    """tests/fuzz/gen_inputs.rnd_mag_case: B / H / M cells, chiral cells with their eight-point stencils and prev-field copies, the H-side CPML on B."""
    import gen_inputs
    run_case(gen_inputs.rnd_mag_case(seed, steps=12), tmp_path, 5)


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 5, 6, 9, 11, 13])
def test_oracle_matches_reference_on_random_periodic_inputs(seed, tmp_path, oracle_lib):
    """tests/fuzz/gen_inputs.rnd_pbc_case: the wrap copies of applyBC1Proc with objects that span the periodic faces."""
    import gen_inputs
    run_case(gen_inputs.rnd_pbc_case(seed), tmp_path, 5)


@pytest.mark.parametrize("seed", [0, 1, 2, 4, 5, 6, 9, 10, 11, 14])
def test_oracle_matches_reference_on_random_surface_normal_dipoles(seed, tmp_path, oracle_lib):
    """tests/fuzz/gen_inputs.rnd_dipnorm_case: UpdateLorPolOrDip with position-dependent dipole grids."""
    import gen_inputs
    run_case(gen_inputs.rnd_dipnorm_case(seed, steps=12), tmp_path, 5)
