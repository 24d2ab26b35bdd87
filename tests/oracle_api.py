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
        for fn in ("oracle_field", "oracle_psi"):
            getattr(L, fn).restype = C.POINTER(C.c_double)
        L.oracle_field.argtypes = [C.c_void_p, C.c_int]
        L.oracle_psi.argtypes = [C.c_void_p, C.c_int, C.c_int]
        for fn in ("oracle_pole", "oracle_ordip_pole"):
            getattr(L, fn).restype = C.POINTER(C.c_double)
            getattr(L, fn).argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.oracle_n_poles.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p) if a.size else None


class OracleSim:
    """The oracle configured from a plan (the same inputs the C ABI gets)."""

    def __init__(self, plan: P.Plan):
        L = lib()
        self.plan = plan
        self._keep = []
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
        for c in plan.cpml:
            psi, grid = np.ascontiguousarray(c.psi), np.ascontiguousarray(c.grid)
            self._chk(L.oracle_set_cpml(self.h, c.comp, c.part, c.has_psi, _ptr(psi), len(psi), _ptr(grid), len(grid)))
        for s in plan.sources:
            loc = (C.c_int32 * 3)(*s.loc)
            sz = (C.c_int32 * 3)(*s.sz)
            self._chk(L.oracle_add_source(self.h, s.field, loc, sz))
        self._chk(L.oracle_commit(self.h))
        self.steps_done = 0

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

    def step_n(self, n: int, nthreads: int = 1, amp: np.ndarray | None = None) -> None:
        if amp is None:
            amp = self.src_amp(self.steps_done, n)
        amp = np.ascontiguousarray(amp, dtype=np.float64)
        self._chk(lib().oracle_step_n(self.h, n, _ptr(amp), nthreads))
        self.steps_done += n

    def _view(self, p) -> np.ndarray | None:
        if not p:
            return None
        lnx, lny, lnz = self.plan.ln
        return np.ctypeslib.as_array(p, shape=(lny, lnz, lnx))

    def field(self, f: int):
        return self._view(lib().oracle_field(self.h, f))

    def pole(self, comp: int, pole: int, prev: int = 0):
        return self._view(lib().oracle_pole(self.h, comp, pole, prev))

    def ordip_pole(self, comp: int, pole: int, prev: int = 0):
        return self._view(lib().oracle_ordip_pole(self.h, comp, pole, prev))

    def psi(self, comp: int, part: int):
        return self._view(lib().oracle_psi(self.h, comp, part))

    def n_poles(self) -> int:
        return lib().oracle_n_poles(self.h)

    def close(self):
        if self.h:
            lib().oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
