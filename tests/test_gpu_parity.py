"""GPU parity: the CUDA path, driven through the C ABI with the reference constructor's own lists,
against (a) the restated oracle on the same inputs and (b) the committed output of the unmodified
reference.  FP64 field work: the bar is rel. L2 <= 1e-10; because the kernels perform
the reference's rounded operations in the reference's order the tests demand bit equality."""
import numpy as np
import pytest

import util
from chiml_b200 import capi
from oracle_api import OracleSim

pytestmark = pytest.mark.gpu

TOL_L2 = 1e-10
TOL_DTC = 1e-9

# Column lengths of the y-marching kernels (chiml_gpu_set_march).  The fixtures are too small to get columns from the automatic
# choice (it keeps >= 8 work items per SM), so without forcing it the register-carried neighbour planes of k_fast / k_uniform --
# the code that moves ~100 % of the bytes of the large configurations -- would run with one plane only.  2, 3 and 7 put the column
# seams at different planes; WHOLE merges every stack of equal tiles into one column (up to the full height of the grid).
WHOLE = 1 << 20
MARCH = [None, 2, 3, 7, WHOLE]
MARCH_IDS = ["auto", "ny2", "ny3", "ny7", "whole"]


@pytest.mark.parametrize("march", MARCH, ids=MARCH_IDS)
@pytest.mark.parametrize("case", util.CASES)
def test_gpu_matches_reference_fixture(case, march):
    plan = util.load_plan(case)
    expect = util.load_expect(case)
    sim = capi.GpuSim(plan, march=march)
    sim.step_n(plan.n_steps)
    for name, ref in expect.items():
        got = util.state_array(sim, name)
        assert got.shape == ref.shape, f"{case}/{name}: shape {got.shape} != {ref.shape}"
        if "pop" in name:
            # population detector = a sum over all emitters: the GPU reduces in tree order, so the bar is the detector
            # tolerance of BASELINE.md (max relative error <= 1e-9), not bit equality
            scale = np.abs(ref).max()
            assert np.abs(got - ref).max() <= TOL_DTC * max(scale, 1e-300), f"{case}/{name}: max rel {np.abs(got - ref).max() / scale:.3e}"
            continue
        assert util.rel_l2(got, ref) <= TOL_L2, f"{case}/{name}: rel L2 {util.rel_l2(got, ref):.3e}"
        assert np.array_equal(got, ref), f"{case}/{name}: not bit-identical, max |diff| {np.abs(got - ref).max():.3e}"
    assert sim.launch_count() > 0
    sim.close()


@pytest.mark.parametrize("march", MARCH, ids=MARCH_IDS)
@pytest.mark.parametrize("case", util.CASES)
def test_gpu_matches_oracle_from_random_state(case, march, oracle_lib):
    """Seeded random fill of every state array (so nothing is identically zero), then N steps."""
    plan = util.load_plan(case)
    rng = np.random.default_rng(1234)
    gpu, cpu = capi.GpuSim(plan, march=march), OracleSim(plan)
    lnx, lny, lnz = plan.ln
    for f in plan.fields_present():
        a = rng.uniform(-1.0, 1.0, size=(lny, lnz, lnx))
        gpu.set_field(f, a)
        cpu.field(f)[...] = a
    n = 25
    gpu.step_n(n)
    cpu.step_n(n)
    for f in plan.fields_present():
        g, c = gpu.field(f), cpu.field(f)
        assert np.array_equal(g, c), f"{case}/{util.P.FIELD_NAMES[f]}: rel L2 {util.rel_l2(g, c):.3e}"
    for comp, part in [(c.comp, c.part) for c in plan.cpml if c.has_psi]:
        assert np.array_equal(gpu.psi(comp, part), cpu.psi(comp, part)), f"{case}: psi comp {comp} part {part}"
    for q, e in enumerate(plan.emitters):
        for sy in range(e.nsys):
            for w in range(5):
                assert np.array_equal(gpu.emitter_state(q, sy, w), cpu.emitter_state(q, sy, w)), f"{case}: emitter set {q} system {sy} array {w}"
        for c in range(3):
            if c in plan.fields_present():
                assert np.array_equal(gpu.emitter_P(q, c), cpu.emitter_P(q, c)), f"{case}: emitter P {q}/{c}"
    gpu.close(); cpu.close()


@pytest.mark.parametrize("kernel", ["thread", "group"])
@pytest.mark.parametrize("case", ["ml3d_two", "ml3d_four", "ml_te", "ml_tm", "c4_small"])
def test_both_density_kernels_match_reference_fixture(case, kernel, monkeypatch):
    """The density matrices are propagated by one thread per emitter (N = 2, 3, 6) or by a group of lanes per emitter exchanging
    operands with warp shuffles (N = 4, 5), csrc/chiml_emitters.cuh; CHIML_B200_EMIT_KERNEL forces either for N <= 5.  Both must
    reproduce the reference bit for bit (two levels, three levels in TE, four levels with two level systems, TM)."""
    monkeypatch.setenv("CHIML_B200_EMIT_KERNEL", kernel)
    plan = util.load_plan(case)
    expect = util.load_expect(case)
    sim = capi.GpuSim(plan)
    sim.step_n(plan.n_steps)
    for name, ref in expect.items():
        got = util.state_array(sim, name)
        if "pop" in name:
            assert np.abs(got - ref).max() <= TOL_DTC * max(np.abs(ref).max(), 1e-300), f"{case}/{name}"
        else:
            assert np.array_equal(got, ref), f"{case}/{name} ({kernel} kernel): max |diff| {np.abs(got - ref).max():.3e}"
    sim.close()


TWO_D = [c for c in util.CASES if util.load_plan(c).ln[2] == 1]


@pytest.mark.parametrize("march", [None, 3], ids=["auto", "ny3"])
@pytest.mark.parametrize("case", TWO_D)
def test_2d_launch_per_phase_path_matches_reference_fixture(case, march):
    """2-D grids without emitters step through the persistent cooperative kernel by default (csrc/chiml_persist.cuh; every other test
    of a 2-D case runs it); this one selects the launch-per-phase path, and both must reproduce the reference's output -- including
    the detector series and the running-DFT accumulators, which the persistent kernel samples inside other phases."""
    plan = util.load_plan(case)
    expect = util.load_expect(case)
    a, b = capi.GpuSim(plan, march=march, persistent=False), capi.GpuSim(plan, march=march, persistent=True)
    for sim in (a, b):
        for n in (3, 4, plan.n_steps - 7):      # several calls: the sample counters and the P buffer parity carry over
            sim.step_n(n)
    for name, ref in expect.items():
        if "pop" in name:
            continue
        for tag, sim in (("launch path", a), ("persistent kernel", b)):
            got = util.state_array(sim, name)
            assert np.array_equal(got, ref), f"{case}/{name} ({tag}): {int((got != ref).sum())} of {got.size} values differ, max |diff| {np.abs(got - ref).max():.3e}"
    for d in range(len(plan.detectors)):
        da, db = a.detector(d), b.detector(d)
        assert da.shape[0] == plan.n_steps // plan.detectors[d].every + 1
        assert np.array_equal(da, db), f"{case}: detector {d}"
    sa = {k["name"]: k["launches"] for k in a.kernel_stats()}
    sb = {k["name"]: k["launches"] for k in b.kernel_stats()}
    assert sa["k_steps_2d"] == 0
    if not plan.emitters and not plan.tfsf and not plan.cplx and not plan.has_B:      # (emitters, TFSF surfaces, complex-field pairs and magnetic media take the launch path)
        assert sb["k_steps_2d"] == 3 and sb["k_fast<E>"] == 0, sb
    a.close(); b.close()


@pytest.mark.parametrize("case", ["te_vacuum", "vac3d"])
def test_detector_series_matches_oracle(case, oracle_lib):
    plan = util.load_plan(case)
    gpu, cpu = capi.GpuSim(plan), OracleSim(plan)
    det = plan.detectors[0]
    box = capi.local_box(plan, det.loc, det.sz)
    (x0, y0, z0), _ = box
    series = [cpu.field(det.field)[y0, z0, x0]]
    for _ in range(40):
        cpu.step_n(1)
        series.append(cpu.field(det.field)[y0, z0, x0])
    gpu.step_n(40)
    got = gpu.detector(0)[:, 0, 0, 0]
    assert got.shape[0] == 41
    ref = np.array(series)
    scale = np.abs(ref).max()
    assert scale > 0
    assert np.max(np.abs(got - ref)) <= 1e-9 * scale
    assert np.array_equal(got, ref)
    gpu.close(); cpu.close()


def test_detector_and_population_rings_wrap_and_grow(oracle_lib):
    """Sample rings (include/chiml_gpu.h: read, then consume): a host that drains after every call keeps the rings at their initial
    size over a run of 3 x the ring capacity (device memory does not grow, samples wrap around); a host that never consumes gets the
    same series from rings that were re-allocated -- at the start of a call, with the retained samples unwrapped."""
    plan = util.load_plan("ml_tm")
    total, chunk = 12500, 700
    a, b = capi.GpuSim(plan), capi.GpuSim(plan)
    bytes0 = a.device_bytes()
    det, pop = [], []
    done = 0
    while done < total:
        n = min(chunk, total - done)
        a.step_n(n); b.step_n(n)
        d = a.detector(0)
        det.append(d)
        a.consume_detector(0, sum(len(x) for x in det))
        p = a.population(0, 0)
        pop.append(p)
        a.consume_population(0, sum(len(x) for x in pop))
        done += n
    assert a.device_bytes() == bytes0, "a drained ring must not grow"
    assert b.device_bytes() > bytes0, "an undrained ring of 12 501 samples must have grown beyond its 4096 slots"
    det, pop = np.concatenate(det), np.concatenate(pop)
    assert det.shape[0] == total + 1 and pop.shape[0] == total
    assert np.array_equal(det, b.detector(0))
    assert np.array_equal(pop, b.population(0, 0))
    # consumed samples are gone, later ones are still addressable by absolute number
    out = np.empty((4,) + det.shape[1:])
    assert a.detector_range(0, 10, 4, out) == 0
    b.consume_detector(0, total - 2)
    assert b.detector_range(0, total - 2, 4, out) == 3 and np.array_equal(out[:3], det[total - 2:])
    # the first steps agree with the CPU oracle (the series is not merely self-consistent)
    cpu = OracleSim(plan)
    d0 = plan.detectors[0]
    (x0, y0, z0), _ = capi.local_box(plan, d0.loc, d0.sz)
    ref = [cpu.field(d0.field)[y0, z0, x0]]
    for _ in range(50):
        cpu.step_n(1)
        ref.append(cpu.field(d0.field)[y0, z0, x0])
    assert np.array_equal(det[:51, 0, 0, 0], np.array(ref))
    a.close(); b.close(); cpu.close()


@pytest.mark.parametrize("case", ["tfsf_te", "tfsf3d_obl", "tfsf3d_slab"])
def test_tfsf_surface_variants_match_oracle(case, oracle_lib):
    """Surface records the reference's fixtures cannot supply (directions with a negative component make its own incident line overflow):
    negative incident strides -- read from the far end, like the BLAS call they replace -- and an eps / mu line on every surface
    (addIncdFieldsEPChange, SOURCE/parallelTFSF.cpp:93-105).  Built from a fixture's records; GPU against the oracle, bit for bit."""
    import copy
    plan = copy.deepcopy(util.load_plan(case))
    rng = np.random.default_rng(7)
    for k, t in enumerate(plan.tfsf):
        if k % 2 == 0:
            t.stride_incd = -t.stride_incd if t.stride_incd else -1
            span = (max(t.n, 1) - 1) * abs(t.stride_incd)
            for pr in (t.pairs_D, t.pairs_U):
                if len(pr):
                    pr[:, 0] = np.minimum(pr[:, 0], t.incd_len - 1 - span)
        if t.ep_mu is None and k % 3 != 1:
            t.ep_mu = rng.uniform(1.0, 3.0, size=t.incd_len)
    gpu, cpu = capi.GpuSim(plan), OracleSim(plan)
    for n in (1, 2, 27):
        gpu.step_n(n)
        cpu.step_n(n)
    for f in plan.fields_present():
        g, c = gpu.field(f), cpu.field(f)
        assert np.abs(c).max() > 0
        assert np.array_equal(g, c), f"{case}/{util.P.FIELD_NAMES[f]}: rel L2 {util.rel_l2(g, c):.3e}"
    gpu.close(); cpu.close()


def test_tfsf_surface_inside_the_cpml_is_refused():
    import copy
    plan = copy.deepcopy(util.load_plan("tfsf_tm"))
    t = next(t for t in plan.tfsf if len(t.pairs_U))
    t.pairs_U[0, 1] = 2 + plan.ln[0] * 2          # a cell of the CPML corner
    with pytest.raises(capi.ChimlError, match="inside the CPML"):
        capi.GpuSim(plan)
