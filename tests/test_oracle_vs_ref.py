"""Pins the restated CPU oracle (oracle/chiml_oracle.c) against the UNMODIFIED reference: the
committed fixtures were produced by oracle/_ref/chiml_ref (the reference's own sources compiled in
place, tests/golden/make_golden.py); the oracle, fed with the reference constructor's own lists,
must reproduce every field and polarisation grid BIT FOR BIT."""
import numpy as np
import pytest

import util
from oracle_api import OracleSim


@pytest.mark.parametrize("case", util.CASES)
def test_oracle_matches_reference_bitwise(case, oracle_lib):
    plan = util.load_plan(case)
    expect = util.load_expect(case)
    sim = OracleSim(plan)
    sim.step_n(plan.n_steps)
    assert expect, "empty fixture"
    for name, ref in expect.items():
        got = util.state_array(sim, name)
        assert got is not None, f"{case}: oracle has no array {name}"
        assert np.abs(ref).max() > 0 or name[0] in "pPoq", f"{case}: reference array {name} is all zero - fixture does not exercise it"
        assert np.array_equal(got, ref), f"{case}/{name}: max |diff| = {np.abs(got - ref).max():.3e}"
    sim.close()


@pytest.mark.parametrize("case", ["lorentz3d", "aniso_slab3d", "tm_au"])
def test_threaded_oracle_is_identical(case, oracle_lib):
    plan = util.load_plan(case)
    a, b = OracleSim(plan), OracleSim(plan)
    a.step_n(20, nthreads=1)
    b.step_n(20, nthreads=4)
    for name in util.state_names(plan):
        assert np.array_equal(util.state_array(a, name), util.state_array(b, name)), name
    a.close(); b.close()


def test_fixture_covers_every_list_kind():
    kinds = set()
    for case in util.CASES:
        plan = util.load_plan(case)
        kinds |= {k for (k, c), v in plan.lists.items() if len(v)}
    assert kinds == {0, 1, 2, 3, 4, 5}
