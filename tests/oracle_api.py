"""ctypes binding of the restated CPU oracle (oracle/chiml_oracle.c).  TEST INFRASTRUCTURE: imported
only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from chiml_b200 import plan as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_ref", "liboracle.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "chiml_ref")


class GridDesc(C.Structure):
    _fields_ = [("mode", C.c_int32), ("ln", C.c_int32 * 3), ("d", C.c_double * 3), ("dt", C.c_double),
                ("has_D", C.c_int32), ("pml_on_D", C.c_int32), ("n_objects", C.c_int32), ("rank", C.c_int32),
                ("nranks", C.c_int32)]


def grid_desc(plan: P.Plan) -> GridDesc:
    g = GridDesc()
    g.mode = plan.mode
    g.ln[:] = plan.ln
    g.d[:] = plan.d
    g.dt = plan.dt
    g.has_D, g.pml_on_D, g.n_objects, g.rank, g.nranks = plan.has_D, plan.pml_on_D, plan.n_objects, plan.rank, plan.nranks
    return g


class EmitterDesc(C.Structure):
    """include/chiml_gpu.h ChimlEmitterDesc"""
    _fields_ = [("nlevel", C.c_int32), ("nsys", C.c_int32), ("nemit", C.c_int32), ("box_lo", C.c_int32 * 3), ("box_n", C.c_int32 * 3),
                ("dt", C.c_double), ("inv_hbar", C.c_double), ("na", C.c_double),
                ("h0", C.c_void_p), ("weight", C.c_void_p), ("mu", C.c_void_p), ("gam_ptr", C.c_void_p), ("gam_col", C.c_void_p),
                ("gam_val", C.c_void_p), ("loc", C.c_void_p), ("eps", C.c_void_p), ("npop", C.c_int32), ("pop_level", C.c_void_p),
                ("pop_every", C.c_int32), ("npoints", C.c_int32), ("object", C.c_int32)]


def emitter_desc(e: P.PlanEmitter, keep: list) -> EmitterDesc:
    """Builds the C struct from a plan record; the numpy arrays it points to are appended to `keep`."""
    d = EmitterDesc()
    d.nlevel, d.nsys, d.nemit = e.nlevel, e.nsys, e.nemit
    d.box_lo[:] = e.box_lo
    d.box_n[:] = e.box_n
    d.dt, d.inv_hbar, d.na = e.dt, e.inv_hbar, e.na
    arrs = {"h0": np.ascontiguousarray(e.h0, np.complex128), "weight": np.ascontiguousarray(e.weight, np.float64),
            "mu": np.ascontiguousarray(e.mu, np.complex128), "gam_ptr": np.ascontiguousarray(e.gam_ptr, np.int32),
            "gam_col": np.ascontiguousarray(e.gam_col, np.int32), "gam_val": np.ascontiguousarray(e.gam_val, np.float64),
            "loc": np.ascontiguousarray(e.loc, np.int32), "eps": np.ascontiguousarray(e.eps, np.float64),
            "pop_level": np.ascontiguousarray(e.pop_level, np.int32)}
    for k, a in arrs.items():
        keep.append(a)
        setattr(d, k, a.ctypes.data if a.size else None)
    d.npop, d.pop_every, d.npoints = e.npop, e.pop_every, e.npoints
    d.object = e.object
    return d


_lib = None


def build_oracle() -> None:
    subprocess.run(["make", "-C", ORACLE_DIR, "oracle"], check=True, stdout=subprocess.DEVNULL)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build_oracle()
        L = C.CDLL(LIB_PATH)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(GridDesc)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_set_update_list.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        L.oracle_set_object.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_set_cpml.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.oracle_add_source.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.oracle_commit.argtypes = [C.c_void_p]
        L.oracle_step_n.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.oracle_step_phase.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_set_periodic.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_set_object_chiral.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_set_prev_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.oracle_set_dip_grid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.oracle_chi_pole.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.oracle_chi_pole.restype = C.POINTER(C.c_double)
        L.oracle_prev_field.argtypes = [C.c_void_p, C.c_int]
        L.oracle_prev_field.restype = C.POINTER(C.c_double)
        L.oracle_set_magnetic.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oracle_set_object_magnetic.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_mag_pole.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.oracle_mag_pole.restype = C.POINTER(C.c_double)
        L.oracle_pair_step_n.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_add_tfsf_surface.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_step_n_tfsf.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        L.oracle_add_dft.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t]
        L.oracle_step_n_dft.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_step_phase_dft.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_dft.restype = C.POINTER(C.c_double)
        L.oracle_dft.argtypes = [C.c_void_p, C.c_int, C.c_int]
        for fn in ("oracle_field", "oracle_psi"):
            getattr(L, fn).restype = C.POINTER(C.c_double)
        L.oracle_field.argtypes = [C.c_void_p, C.c_int]
        L.oracle_psi.argtypes = [C.c_void_p, C.c_int, C.c_int]
        for fn in ("oracle_pole", "oracle_ordip_pole"):
            getattr(L, fn).restype = C.POINTER(C.c_double)
            getattr(L, fn).argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.oracle_n_poles.argtypes = [C.c_void_p]
        L.oracle_add_emitters.argtypes = [C.c_void_p, C.POINTER(EmitterDesc)]
        L.oracle_emitter_state.restype = C.POINTER(C.c_double)
        L.oracle_emitter_state.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oracle_emitter_P.restype = C.POINTER(C.c_double)
        L.oracle_emitter_P.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oracle_population.restype = C.c_size_t
        L.oracle_population.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p) if a.size else None


class OracleSim:
    """The oracle configured from a plan (the same inputs the C ABI gets)."""

    def __init__(self, plan: P.Plan, _part: str = "re"):
        L = lib()
        self.plan = plan
        self._keep = []
        self.imag = None
        self._part = _part
        g = grid_desc(plan)
        self.h = L.oracle_create(C.byref(g))
        if not self.h:
            raise MemoryError("oracle_create failed")
        for (kind, comp), runs in plan.lists.items():
            runs = np.ascontiguousarray(runs)
            self._chk(L.oracle_set_update_list(self.h, kind, comp, _ptr(runs), len(runs)))
        for o in plan.objects:
            a, x, gm, dp = (np.ascontiguousarray(v, dtype=np.float64) for v in (o.alpha, o.xi, o.gamma, o.dip))
            self._chk(L.oracle_set_object(self.h, o.obj, o.npoles, _ptr(a), _ptr(x), _ptr(gm), o.use_or_dip, _ptr(dp)))
        if plan.has_B:
            self._chk(L.oracle_set_magnetic(self.h, plan.has_B, plan.pml_on_B))
            for obj, (a, x, gm) in sorted(plan.mag_objects.items()):
                a, x, gm = (np.ascontiguousarray(v, dtype=np.float64) for v in (a, x, gm))
                self._chk(L.oracle_set_object_magnetic(self.h, obj, len(a), _ptr(a), _ptr(x), _ptr(gm)))
        for (comp, pole), g in sorted(plan.dip_grids.items()):
            g = np.ascontiguousarray(g, dtype=np.float64)
            self._chk(L.oracle_set_dip_grid(self.h, comp, pole, _ptr(g)))
        for obj, arrs in sorted(plan.chi_objects.items()):
            a, x, gm, gp = (np.ascontiguousarray(v, dtype=np.float64) for v in arrs)
            self._chk(L.oracle_set_object_chiral(self.h, obj, len(a), _ptr(a), _ptr(x), _ptr(gm), _ptr(gp)))
        if plan.prev_copy is not None:
            rows = np.ascontiguousarray(plan.prev_copy, dtype=np.int32)
            self._chk(L.oracle_set_prev_copy(self.h, _ptr(rows), len(rows)))
        for c in plan.cpml:
            psi, grid = np.ascontiguousarray(c.psi), np.ascontiguousarray(c.grid)
            self._chk(L.oracle_set_cpml(self.h, c.comp, c.part, c.has_psi, _ptr(psi), len(psi), _ptr(grid), len(grid)))
        for s in plan.sources:
            loc = (C.c_int32 * 3)(*s.loc)
            sz = (C.c_int32 * 3)(*s.sz)
            self._chk(L.oracle_add_source(self.h, s.field, loc, sz))
        for d in plan.dfts:
            lines = np.ascontiguousarray(d.lines, dtype=np.int32)
            self._chk(L.oracle_add_dft(self.h, d.field, d.group, d.every, d.nfreq, d.npts, d.stride, _ptr(lines), len(lines), d.acc_len))
        for e in plan.emitters:
            d = emitter_desc(e, self._keep)
            self._chk(L.oracle_add_emitters(self.h, C.byref(d)))
        for t in plan.tfsf:
            from chiml_b200 import capi
            d = capi.tfsf_surface(t, self._keep)
            self._chk(L.oracle_add_tfsf_surface(self.h, C.byref(d)))
        if not plan.cplx:
            for comp, w in sorted(plan.periodic.items()):
                self._chk(L.oracle_set_periodic(self.h, comp, (C.c_int32 * 7)(*w)))
        self._chk(L.oracle_commit(self.h))
        self.steps_done = 0
        if plan.cplx and _part == "re":
            # complex fields: a second simulation over the same lists holds the imaginary parts; the pair is stepped through this one
            self.imag = OracleSim(plan, _part="im")

    @staticmethod
    def _chk(rc):
        if rc != 0:
            raise RuntimeError(f"oracle call failed with status {rc}")

    def src_amp(self, start: int, n: int) -> np.ndarray:
        ns = len(self.plan.sources)
        amp = np.zeros((n, max(ns, 1)), dtype=np.float64)
        for q, s in enumerate(self.plan.sources):
            seg = s.amp[start:start + n]
            amp[:len(seg), q] = seg
        return amp

    def src_amp_im(self, start: int, n: int) -> np.ndarray:
        ns = len(self.plan.sources)
        amp = np.zeros((n, max(ns, 1)), dtype=np.float64)
        for q, s in enumerate(self.plan.sources):
            seg = s.amp_im[start:start + n]
            amp[:len(seg), q] = seg
        return amp

    def step_n(self, n: int, nthreads: int = 1, amp: np.ndarray | None = None) -> None:
        if amp is None:
            amp = self.src_amp(self.steps_done, n)
        amp = np.ascontiguousarray(amp, dtype=np.float64)
        if self.plan.cplx:
            assert self.imag is not None, "the imaginary part of a complex-field pair is stepped through the real part"
            amp_im = np.ascontiguousarray(self.src_amp_im(self.steps_done, n))
            wraps = (C.c_int32 * 42)()
            has = (C.c_int * 6)()
            for comp, w in self.plan.periodic.items():
                wraps[7 * comp:7 * comp + 7] = w
                has[comp] = 1
            k = (C.c_double * 3)(*self.plan.k_point)
            self._chk(lib().oracle_pair_step_n(self.h, self.imag.h, n, _ptr(amp), _ptr(amp_im), wraps, has, k))
            self.steps_done += n
            self.imag.steps_done += n
            return
        if self.plan.tfsf:
            from chiml_b200 import capi
            tw = np.ascontiguousarray(P.dft_twiddles(self.plan, self.steps_done, n)) if self.plan.dfts else None
            rows = capi.tfsf_rows(self.plan, self.steps_done, n)
            self._chk(lib().oracle_step_n_tfsf(self.h, n, _ptr(amp), _ptr(tw) if tw is not None else None, _ptr(rows), rows.shape[1], nthreads))
        elif self.plan.dfts:
            tw = np.ascontiguousarray(P.dft_twiddles(self.plan, self.steps_done, n))
            self._chk(lib().oracle_step_n_dft(self.h, n, _ptr(amp), _ptr(tw), nthreads))
        else:
            self._chk(lib().oracle_step_n(self.h, n, _ptr(amp), nthreads))
        self.steps_done += n

    def step_phase(self, phase: int, amp: np.ndarray) -> None:
        """One phase of one step (chiml_b200/slab.py); amp = the amplitudes of this step, shape (1, n_sources)."""
        amp = np.ascontiguousarray(amp, dtype=np.float64)
        if self.plan.dfts:
            if phase == 3:
                self._tw = np.ascontiguousarray(P.dft_twiddles(self.plan, self.steps_done, 1))
            self._chk(lib().oracle_step_phase_dft(self.h, phase, _ptr(amp), _ptr(self._tw) if phase == 3 else None))
        else:
            self._chk(lib().oracle_step_phase(self.h, phase, _ptr(amp)))
        if phase == 3:
            self.steps_done += 1

    def _view(self, p) -> np.ndarray | None:
        if not p:
            return None
        lnx, lny, lnz = self.plan.ln
        return np.ctypeslib.as_array(p, shape=(lny, lnz, lnx))

    def field(self, f: int):
        return self._view(lib().oracle_field(self.h, f))

    def pole(self, comp: int, pole: int, prev: int = 0):
        return self._view(lib().oracle_pole(self.h, comp, pole, prev))

    def chi_pole(self, comp: int, pole: int, prev: int = 0):
        return self._view(lib().oracle_chi_pole(self.h, comp, pole, prev))

    def prev_field(self, comp: int):
        return self._view(lib().oracle_prev_field(self.h, comp))

    def mag_pole(self, comp: int, pole: int, prev: int = 0):
        return self._view(lib().oracle_mag_pole(self.h, comp, pole, prev))

    def ordip_pole(self, comp: int, pole: int, prev: int = 0):
        return self._view(lib().oracle_ordip_pole(self.h, comp, pole, prev))

    def psi(self, comp: int, part: int):
        return self._view(lib().oracle_psi(self.h, comp, part))

    def emitter_state(self, slot: int, sys: int, which: int) -> np.ndarray:
        """(nemit, N*N) complex view of rho (which=0) or a derivative history (1..4) of level system `sys`."""
        e = self.plan.emitters[slot]
        p = lib().oracle_emitter_state(self.h, slot, which)
        a = np.ctypeslib.as_array(p, shape=(e.nsys, e.nemit, e.nlevel * e.nlevel, 2))
        return a[sys, :, :, 0] + 1j * a[sys, :, :, 1]

    def emitter_P(self, slot: int, comp: int) -> np.ndarray:
        e = self.plan.emitters[slot]
        p = lib().oracle_emitter_P(self.h, slot, comp)
        return np.ctypeslib.as_array(p, shape=(e.box_n[1] + 2, e.pz, e.box_n[0] + 2))

    def population(self, slot: int, det: int) -> np.ndarray:
        n = lib().oracle_population(self.h, slot, det, None, 0)
        out = np.zeros((n, 2))
        lib().oracle_population(self.h, slot, det, _ptr(out) if n else None, n)
        return out[:, 0] + 1j * out[:, 1]

    def dft(self, slot: int) -> np.ndarray:
        d = self.plan.dfts[slot]
        re = np.ctypeslib.as_array(lib().oracle_dft(self.h, slot, 0), shape=(max(d.acc_len, 1),))[:d.acc_len]
        im = np.ctypeslib.as_array(lib().oracle_dft(self.h, slot, 1), shape=(max(d.acc_len, 1),))[:d.acc_len]
        return re + 1j * im

    def n_poles(self) -> int:
        return lib().oracle_n_poles(self.h)

    def close(self):
        if self.imag is not None:
            self.imag.close()
            self.imag = None
        if self.h:
            lib().oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
