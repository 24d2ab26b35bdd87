"""Size-independent properties at sizes the CPU checker cannot reach: the engine at a C3/C5-like grid of
tens of millions of cells, checked without a reference run.

* linearity: without emitters the update is linear in the fields, and scaling by a power of two is exact in binary floating point,
  so doubling every source amplitude must double every field value BIT FOR BIT (any cell updated twice, skipped, or fed from a wrong
  neighbour breaks it only if it breaks linearity -- so the second property complements it);
* tiling independence: the same grid run with differently shaped work decompositions (one call of N steps vs N calls of one step;
  a grid whose x extent shifts every tile boundary) must agree exactly where they overlap in meaning: here, one call vs many calls.
"""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from chiml_b200 import capi, inputs as I, plan as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _plan(nx, ny, nz, steps, sheet):
    work = tempfile.mkdtemp(prefix="chiml_prop_")
    cfg = I.c5_aniso_ml(nx=nx - 1, ny=ny - 1, nz=nz - 1, steps=steps, sheet=sheet, out="prop_out/c5")
    I.write(cfg, os.path.join(work, "p.json"))
    subprocess.run([os.path.join(ROOT, "chiml_b200", "chiml_plan"), os.path.join(work, "p.json"), os.path.join(work, "p")], check=True)
    return P.read_plan(os.path.join(work, "p.rank0.plan"))


def test_linearity_is_exact_at_scale():
    plan = _plan(384, 96, 160, 24, sheet=False)          # 5.9 Mcell, oriented-dipole slab through the CPML on every face
    a, b = capi.GpuSim(plan), capi.GpuSim(plan)
    n = 24
    amp = a.src_amp(0, n)
    assert np.abs(amp).max() > 0
    a.step_n(n, amp)
    b.step_n(n, 2.0 * amp)
    for f in plan.fields_present():
        fa, fb = a.field(f), b.field(f)
        assert np.abs(fa).max() > 0, P.FIELD_NAMES[f]
        assert np.array_equal(2.0 * fa, fb), f"{P.FIELD_NAMES[f]}: doubling the source did not double the field exactly"
    for c in range(3):
        assert np.array_equal(2.0 * a.ordip_pole(c, 0), b.ordip_pole(c, 0))
    a.close(); b.close()


def test_one_call_equals_many_calls_with_emitters():
    plan = _plan(256, 96, 128, 16, sheet=True)           # with the emitter sheet: density matrices, P feedback, population detector
    a, b = capi.GpuSim(plan), capi.GpuSim(plan)
    a.step_n(16)
    for _ in range(16):
        b.step_n(1)
    for f in plan.fields_present():
        assert np.array_equal(a.field(f), b.field(f)), P.FIELD_NAMES[f]
    e = plan.emitters[0]
    assert e.nemit > 1000
    for w in range(5):
        assert np.array_equal(a.emitter_state(0, 0, w), b.emitter_state(0, 0, w))
    assert np.array_equal(a.population(0, 0), b.population(0, 0))
    rho = a.emitter_state(0, 0, 0)
    # trace and hermiticity of every density matrix are conserved by the propagator to rounding
    assert np.abs(rho[:, 0] + rho[:, 3] - 1.0).max() < 1e-12
    assert np.abs(rho[:, 1] - np.conj(rho[:, 2])).max() < 1e-15
    a.close(); b.close()
