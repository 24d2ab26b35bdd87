"""Error behaviour of the C ABI: every misuse returns a status and a message (no exception crosses the boundary, nothing is
computed "approximately"): wrong call order, lists that leave the grid, inputs outside the covered hot path, slabs that were never
bound, DFT sets stepped without twiddles."""
import ctypes as C

import numpy as np
import pytest

import util
from chiml_b200 import capi, plan as P

pytestmark = pytest.mark.gpu


def _ctx(plan, **over):
    L = capi.lib()
    g = capi.GridDesc()
    g.mode = plan.mode
    g.ln[:] = plan.ln
    g.d[:] = plan.d
    g.dt = plan.dt
    g.has_D, g.pml_on_D, g.n_objects, g.rank, g.nranks = plan.has_D, plan.pml_on_D, plan.n_objects, 0, 1
    for k, v in over.items():
        setattr(g, k, v)
    h = C.c_void_p()
    rc = L.chiml_gpu_create(C.byref(g), 0, C.byref(h))
    return L, h, rc


def _msg(L, h):
    return L.chiml_gpu_last_error(h).decode()


def test_call_order_and_arguments():
    plan = util.load_plan("vac3d")
    L, h, rc = _ctx(plan)
    assert rc == 0
    assert L.chiml_gpu_step_n(h, 1, None) == 4 and "before commit" in _msg(L, h)            # ERR_STATE
    runs = np.ascontiguousarray(plan.get_list(P.LIST_U, 0)[:1]).copy()
    runs["ind"][0] = plan.ncell - 1
    runs["n"][0] = 50
    assert L.chiml_gpu_set_update_list(h, P.LIST_U, 0, runs.ctypes.data_as(C.c_void_p), 1) == 1 and "outside the grid" in _msg(L, h)
    assert L.chiml_gpu_set_update_list(h, 7, 0, None, 0) == 1
    assert L.chiml_gpu_set_update_list(h, P.LIST_LORD, 3, runs.ctypes.data_as(C.c_void_p), 1) in (1, 3)   # magnetic lists: unsupported
    loc, sz = (C.c_int32 * 3)(0, 0, 0), (C.c_int32 * 3)(1, 1, 10 ** 6)
    assert L.chiml_gpu_add_source(h, 2, loc, sz, None) == 1 and "outside the local grid" in _msg(L, h)
    assert L.chiml_gpu_commit(h) == 0
    assert L.chiml_gpu_commit(h) == 4 and "twice" in _msg(L, h)
    assert L.chiml_gpu_add_source(h, 2, loc, (C.c_int32 * 3)(1, 1, 1), None) == 4
    assert L.chiml_gpu_step_n(h, -1, None) == 1
    L.chiml_gpu_destroy(h)


def test_bad_grid_descriptions():
    plan = util.load_plan("vac3d")
    for over in ({"mode": 7}, {"n_objects": 0}, {"mode": P.MODE_TE}):           # TE with ln[2] > 1
        L, h, rc = _ctx(plan, **over)
        assert rc == 1 and L.chiml_gpu_last_error(None)
    g_ok = _ctx(plan)
    assert g_ok[2] == 0
    g_ok[0].chiml_gpu_destroy(g_ok[1])


def test_unbound_slab_refuses_to_step():
    plan = util.load_plan("vac3d")
    L, h, rc = _ctx(plan, nranks=2)
    assert rc == 0 and L.chiml_gpu_commit(h) == 0
    assert L.chiml_gpu_step_n(h, 1, None) != 0          # two slabs declared, halo never bound
    assert L.chiml_gpu_halo_bind(h, None, 0, None, 0) == 1 and "neighbours" in _msg(L, h)
    L.chiml_gpu_destroy(h)


def test_dft_sets_need_twiddles():
    plan = util.load_plan("tm_flux")
    sim = capi.GpuSim(plan)
    amp = np.ascontiguousarray(sim.src_amp(0, 1))
    rc = capi.lib().chiml_gpu_step_n(sim.h, 1, amp.ctypes.data_as(C.c_void_p))
    assert rc == 1 and "chiml_gpu_step_n_dft" in _msg(capi.lib(), sim.h)
    sim.close()


def test_emitter_limits():
    plan = util.load_plan("ml3d_two")
    L, h, rc = _ctx(plan)
    keep = []
    d = capi.emitter_desc(plan.emitters[0], keep)
    d.nlevel = 9
    assert L.chiml_gpu_add_emitters(h, C.byref(d), None) == 3 and "nlevel" in _msg(L, h)   # ERR_UNSUPPORTED
    d = capi.emitter_desc(plan.emitters[0], keep)
    d.box_lo[0] = plan.ln[0]
    assert L.chiml_gpu_add_emitters(h, C.byref(d), None) == 1
    L.chiml_gpu_destroy(h)
