"""GPU parity at sizes where the y-marching kernels form columns ON THEIR OWN (no forced column length): the grids below give the
automatic choice of chiml_gpu_commit 6..64-plane columns, like the benchmark grids, and are still small enough for the threaded CPU
oracle to step them in seconds.  Every state array -- fields, D, CPML psi, isotropic and oriented-dipole poles, density matrices and
their derivative histories, emitter polarisation -- must agree with the oracle BIT FOR BIT after 10 steps from a seeded random state
(the tolerance north_star asks for is rel. L2 <= 1e-10; the arithmetic order is reproducible, so the test demands equality).

Reference arithmetic: UTIL/FDTD_up_eq.cpp:10-35 (curl), PML/parallelPML.cpp:12-40 (CPML), UTIL/FDTD_up_eq.cpp:425-631,838-889 (poles,
D->E), ML/parallelQE.hpp:614-770 (emitters), restated by oracle/chiml_oracle.c and pinned against the reference build."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import util
from chiml_b200 import capi, inputs as I, plan as P
from oracle_api import OracleSim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu
NTHREADS = max(1, min(16, os.cpu_count() or 1))


def _plan(cfg):
    work = tempfile.mkdtemp(prefix="chiml_mid_")
    I.write(cfg, os.path.join(work, "p.json"))
    subprocess.run([os.path.join(ROOT, "chiml_b200", "chiml_plan"), os.path.join(work, "p.json"), os.path.join(work, "p")], check=True)
    return P.read_plan(os.path.join(work, "p.rank0.plan"))


def _compare(plan, march, steps=10, seed=99, persistent=None):
    rng = np.random.default_rng(seed)
    gpu, cpu = capi.GpuSim(plan, march=march, persistent=persistent), OracleSim(plan)
    lnx, lny, lnz = plan.ln
    for f in plan.fields_present():
        a = rng.uniform(-1.0, 1.0, size=(lny, lnz, lnx))
        gpu.set_field(f, a)
        cpu.field(f)[...] = a
    gpu.step_n(steps)
    cpu.step_n(steps, NTHREADS)
    bad = []
    for name in util.state_names(plan):
        g, c = util.state_array(gpu, name), util.state_array(cpu, name)
        if name in P.FIELD_NAMES:
            assert np.abs(c).max() > 0, f"{name}: the oracle array is identically zero, the case checks nothing"
        if not np.array_equal(g, c):
            bad.append(f"{name}: rel L2 {util.rel_l2(g, c):.3e}, {int((g != c).sum())} of {g.size} values differ")
    for comp, part in [(c.comp, c.part) for c in plan.cpml if c.has_psi]:
        if not np.array_equal(gpu.psi(comp, part), cpu.psi(comp, part)):
            bad.append(f"psi comp {comp} part {part}")
    stats = {k["name"]: k["launches"] for k in gpu.kernel_stats()}
    gpu.close(); cpu.close()
    assert not bad, "; ".join(bad)
    return stats


# (64, 32) are the column lengths of the benchmark grids (chiml_gpu_commit: 64 planes for k_fast, 32 for k_uniform)
@pytest.mark.parametrize("march", [None, (64, 32)], ids=["auto", "bench_columns"])
def test_c5_slab_with_emitter_sheet_384x192x160(march, oracle_lib):
    """BASELINE C5 in small: oriented-dipole Lorentz slab through the CPML on every face + two-level emitter sheet.  23 040 tiles per
    half step -> automatic columns of 19 planes."""
    plan = _plan(I.c5_aniso_ml(nx=383, ny=191, nz=159, steps=10, sheet=True, out="mid_out/c5"))
    assert plan.emitters and plan.emitters[0].nemit > 10000 and plan.n_ordip_poles == 1
    stats = _compare(plan, march)
    assert stats["k_fast<E>"] > 0 and stats["k_uniform<E>"] > 0 and stats["k_ordip_poles"] > 0 and stats["k_emit_density"] > 0


@pytest.mark.parametrize("march", [None, (64, 32)], ids=["auto", "bench_columns"])
def test_c4_gold_cubes_under_emitter_sheet_192cubed(march, oracle_lib):
    """BASELINE C4 in small: 2 x 2 six-pole Au cubes (isotropic poles fused into the UNIFORM D->E bodies, object edges and corners as
    rectangle records, GENERAL tiles none) under a two-level emitter sheet, plane source."""
    plan = _plan(I.c4_plasmonic_ml(n=191, steps=10, narray=2, cube=40, pitch=70, sheet=100, out="mid_out/c4"))
    assert plan.n_lor_poles == 6 and plan.emitters[0].nemit == 100 * 100
    stats = _compare(plan, march)
    assert stats["k_uniform<E>"] > 0 and stats["k_emit_density"] > 0


@pytest.mark.parametrize("march", [None, (64, 32)], ids=["auto", "bench_columns"])
def test_c2_drude_rod_2048sq(march, oracle_lib):
    """BASELINE C2 at full size (2-D TM, Drude nanorod, CPML, flux box with running DFT): eight one-row tiles per block, each warp
    marching its own column (automatic column length 6)."""
    plan = _plan(I.c2_tm_drude(n=2047, steps=10, nfreq=8, out="mid_out/c2"))
    stats = _compare(plan, march, persistent=False)
    assert stats["k_fast<E>"] > 0 and stats["k_dft"] > 0
    stats = _compare(plan, march, persistent=True)      # every step in one cooperative launch (csrc/chiml_persist.cuh), forced
    assert stats["k_steps_2d"] == 1 and stats["k_fast<E>"] == 0
    stats = _compare(plan, march)                       # the automatic choice for a grid of 4.2 M points: the launch-per-phase path
    assert stats["k_steps_2d"] == 0 and stats["k_fast<E>"] > 0


def test_c1_te_vacuum_1024sq_whole_columns(oracle_lib):
    """2-D TE (Ex, Ey, Hz): the other one-component curl variants, with every stack of equal tiles merged into one column."""
    plan = _plan(I.c1_te_vacuum(n=1023, steps=10, out="mid_out/c1"))
    _compare(plan, 1 << 20)
