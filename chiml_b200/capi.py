"""ctypes binding of the C ABI (include/chiml_gpu.h) of the B200 engine.

PyTorch is not involved in the product path: the shared library owns its device memory and
streams.  This module fails loudly when the CUDA library is missing or no GPU is visible -- there
is no CPU fallback (the CPU oracle under oracle/ is test infrastructure and is never imported here).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import plan as P

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libchiml_b200.so")

STATUS = {0: "OK", 1: "ERR_ARG", 2: "ERR_CUDA", 3: "ERR_UNSUPPORTED", 4: "ERR_STATE", 5: "ERR_NO_DEVICE"}

EXPORTED_SYMBOLS = [
    "chiml_gpu_device_count", "chiml_gpu_create", "chiml_gpu_destroy", "chiml_gpu_last_error",
    "chiml_gpu_set_update_list", "chiml_gpu_set_object", "chiml_gpu_set_cpml", "chiml_gpu_add_source",
    "chiml_gpu_add_detector", "chiml_gpu_commit", "chiml_gpu_step_n", "chiml_gpu_sync", "chiml_gpu_step_n_timed",
    "chiml_gpu_launch_count", "chiml_gpu_upload_field", "chiml_gpu_download_field", "chiml_gpu_download_pole",
    "chiml_gpu_upload_pole", "chiml_gpu_download_ordip_pole", "chiml_gpu_download_psi", "chiml_gpu_read_detector",
    "chiml_gpu_device_bytes", "chiml_gpu_set_kernel_timing", "chiml_gpu_n_kernel_kinds", "chiml_gpu_kernel_stat",
    "chiml_gpu_reset_kernel_stats", "chiml_gpu_read_detector_range", "chiml_gpu_add_emitters",
    "chiml_gpu_download_emitter_state", "chiml_gpu_download_emitter_pol", "chiml_gpu_read_population",
    "chiml_gpu_halo_export", "chiml_gpu_halo_bind", "chiml_gpu_add_dft", "chiml_gpu_step_n_dft", "chiml_gpu_download_dft",
    "chiml_gpu_set_march", "chiml_gpu_set_ordip_pole_count", "chiml_gpu_reserve_steps", "chiml_gpu_consume_detector",
    "chiml_gpu_consume_population", "chiml_gpu_set_persistent", "chiml_gpu_set_periodic", "chiml_gpu_add_tfsf_surface",
    "chiml_gpu_step_n_tfsf", "chiml_gpu_bind_imag", "chiml_gpu_step_n_cplx",
    "chiml_gpu_set_magnetic", "chiml_gpu_set_object_magnetic", "chiml_gpu_download_mag_pole",
    "chiml_gpu_set_dip_grid", "chiml_gpu_set_object_chiral", "chiml_gpu_set_prev_copy", "chiml_gpu_download_chi_pole", "chiml_gpu_download_prev_field",
]


class ChimlError(RuntimeError):
    pass


class GridDesc(C.Structure):
    _fields_ = [("mode", C.c_int32), ("ln", C.c_int32 * 3), ("d", C.c_double * 3), ("dt", C.c_double),
                ("has_D", C.c_int32), ("pml_on_D", C.c_int32), ("n_objects", C.c_int32), ("rank", C.c_int32),
                ("nranks", C.c_int32)]


class EmitterDesc(C.Structure):
    """include/chiml_gpu.h ChimlEmitterDesc"""
    _fields_ = [("nlevel", C.c_int32), ("nsys", C.c_int32), ("nemit", C.c_int32), ("box_lo", C.c_int32 * 3), ("box_n", C.c_int32 * 3),
                ("dt", C.c_double), ("inv_hbar", C.c_double), ("na", C.c_double),
                ("h0", C.c_void_p), ("weight", C.c_void_p), ("mu", C.c_void_p), ("gam_ptr", C.c_void_p), ("gam_col", C.c_void_p),
                ("gam_val", C.c_void_p), ("loc", C.c_void_p), ("eps", C.c_void_p), ("npop", C.c_int32), ("pop_level", C.c_void_p),
                ("pop_every", C.c_int32), ("npoints", C.c_int32), ("object", C.c_int32)]


class TfsfSurface(C.Structure):
    """include/chiml_gpu.h ChimlTfsfSurface"""
    _fields_ = [("comp", C.c_int32), ("incd_offset", C.c_int32), ("incd_len", C.c_int32), ("n", C.c_int32), ("stride_incd", C.c_int32),
                ("stride_main", C.c_int32), ("npairs_D", C.c_int32), ("npairs_U", C.c_int32), ("prefactor", C.c_double),
                ("pairs_D", C.c_void_p), ("pairs_U", C.c_void_p), ("ep_mu", C.c_void_p)]


def tfsf_surface(t: "P.PlanTfsfSurface", keep: list) -> TfsfSurface:
    d = TfsfSurface()
    d.comp, d.incd_offset, d.incd_len, d.n, d.stride_incd, d.stride_main = t.comp, t.incd_offset, t.incd_len, t.n, t.stride_incd, t.stride_main
    d.npairs_D, d.npairs_U, d.prefactor = len(t.pairs_D), len(t.pairs_U), t.prefactor
    pd, pu = np.ascontiguousarray(t.pairs_D, np.int32), np.ascontiguousarray(t.pairs_U, np.int32)
    em = np.ascontiguousarray(t.ep_mu, np.float64) if t.ep_mu is not None else None
    keep += [pd, pu, em]
    d.pairs_D = pd.ctypes.data if pd.size else None
    d.pairs_U = pu.ctypes.data if pu.size else None
    d.ep_mu = em.ctypes.data if em is not None else None
    return d


def tfsf_rows(plan: "P.Plan", start: int, n: int) -> np.ndarray:
    """Rows start .. start+n-1 of the incident-line table (recorded from the reference's own 1-D line, plan record TFSFLINE)."""
    if plan.tfsf_lines is None or start + n > len(plan.tfsf_lines):
        raise ValueError("the plan holds TFSF surfaces but no incident-line table for these steps")
    return np.ascontiguousarray(plan.tfsf_lines[start:start + n], dtype=np.float64)


def emitter_desc(e: "P.PlanEmitter", keep: list) -> EmitterDesc:
    """C struct of one plan emitter record; the numpy arrays it points to are appended to `keep`."""
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


class KernelStat(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("launches", C.c_int64), ("timed_launches", C.c_int64), ("ms_total", C.c_double),
                ("alg_bytes_per_launch", C.c_double), ("alg_bytes_per_step", C.c_double)]


class _Tolerant:
    """Attribute access on the CDLL that yields a dummy for symbols an OLDER build of the library lacks (development aid for A/B
    timing of two builds through CHIML_B200_LIB); calling such a symbol raises."""

    class _Missing:
        def __init__(self, name):
            self.name = name

        def __call__(self, *a):
            raise ChimlError(f"{self.name} is not exported by this build of the library")

    def __init__(self, dll):
        object.__setattr__(self, "_dll", dll)

    def __getattr__(self, name):
        try:
            return getattr(self._dll, name)
        except AttributeError:
            return _Tolerant._Missing(name)


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load chiml_b200/libchiml_b200.so (built in-tree by __graft_entry__.build / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ChimlError(f"{LIB_PATH} is missing: build it with `make -C chiml_b200/csrc` (nvcc, sm_100a). "
                         "There is no CPU fallback.")
    L = _Tolerant(C.CDLL(os.environ.get("CHIML_B200_LIB", LIB_PATH)))
    vp, i, sz = C.c_void_p, C.c_int, C.c_size_t
    L.chiml_gpu_device_count.restype = i
    L.chiml_gpu_create.argtypes = [C.POINTER(GridDesc), i, C.POINTER(vp)]
    L.chiml_gpu_destroy.argtypes = [vp]
    L.chiml_gpu_destroy.restype = None
    L.chiml_gpu_last_error.argtypes = [vp]
    L.chiml_gpu_last_error.restype = C.c_char_p
    L.chiml_gpu_set_update_list.argtypes = [vp, i, i, vp, sz]
    L.chiml_gpu_set_object.argtypes = [vp, i, i, vp, vp, vp, i, vp]
    L.chiml_gpu_set_cpml.argtypes = [vp, i, i, i, vp, sz, vp, sz]
    L.chiml_gpu_add_source.argtypes = [vp, i, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(i)]
    L.chiml_gpu_add_detector.argtypes = [vp, i, C.POINTER(C.c_int32), C.POINTER(C.c_int32), i, C.POINTER(i)]
    L.chiml_gpu_commit.argtypes = [vp]
    L.chiml_gpu_set_march.argtypes = [vp, i, i]
    L.chiml_gpu_step_n.argtypes = [vp, i, vp]
    L.chiml_gpu_sync.argtypes = [vp]
    L.chiml_gpu_step_n_timed.argtypes = [vp, i, vp, vp, C.POINTER(C.c_float)]
    L.chiml_gpu_set_ordip_pole_count.argtypes = [vp, i]
    L.chiml_gpu_set_persistent.argtypes = [vp, i]
    L.chiml_gpu_reserve_steps.argtypes = [vp, C.c_longlong]
    L.chiml_gpu_consume_detector.argtypes = [vp, i, sz]
    L.chiml_gpu_consume_population.argtypes = [vp, i, sz]
    L.chiml_gpu_launch_count.argtypes = [vp]
    L.chiml_gpu_launch_count.restype = C.c_int64
    L.chiml_gpu_upload_field.argtypes = [vp, i, vp]
    L.chiml_gpu_download_field.argtypes = [vp, i, vp]
    L.chiml_gpu_download_pole.argtypes = [vp, i, i, i, vp]
    L.chiml_gpu_upload_pole.argtypes = [vp, i, i, i, vp]
    L.chiml_gpu_download_ordip_pole.argtypes = [vp, i, i, i, vp]
    L.chiml_gpu_download_psi.argtypes = [vp, i, i, vp]
    L.chiml_gpu_read_detector.argtypes = [vp, i, vp, sz, C.POINTER(sz)]
    L.chiml_gpu_read_detector_range.argtypes = [vp, i, sz, sz, vp, C.POINTER(sz)]
    L.chiml_gpu_halo_export.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.chiml_gpu_halo_bind.argtypes = [vp, C.c_char_p, sz, C.c_char_p, sz]
    L.chiml_gpu_set_periodic.argtypes = [vp, i, vp]
    L.chiml_gpu_set_object_chiral.argtypes = [vp, i, i, vp, vp, vp, vp]
    L.chiml_gpu_set_prev_copy.argtypes = [vp, vp, sz]
    L.chiml_gpu_set_dip_grid.argtypes = [vp, i, i, vp]
    L.chiml_gpu_download_chi_pole.argtypes = [vp, i, i, i, vp]
    L.chiml_gpu_download_prev_field.argtypes = [vp, i, vp]
    L.chiml_gpu_set_magnetic.argtypes = [vp, i, i]
    L.chiml_gpu_set_object_magnetic.argtypes = [vp, i, i, vp, vp, vp]
    L.chiml_gpu_download_mag_pole.argtypes = [vp, i, i, i, vp]
    L.chiml_gpu_bind_imag.argtypes = [vp, vp, vp]
    L.chiml_gpu_step_n_cplx.argtypes = [vp, i, vp, vp]
    L.chiml_gpu_add_tfsf_surface.argtypes = [vp, C.POINTER(TfsfSurface)]
    L.chiml_gpu_step_n_tfsf.argtypes = [vp, i, vp, vp, vp, sz]
    L.chiml_gpu_add_dft.argtypes = [vp, i, i, i, i, i, i, vp, sz, sz, C.POINTER(i)]
    L.chiml_gpu_step_n_dft.argtypes = [vp, i, vp, vp]
    L.chiml_gpu_download_dft.argtypes = [vp, i, vp, vp]
    L.chiml_gpu_add_emitters.argtypes = [vp, C.POINTER(EmitterDesc), C.POINTER(i)]
    L.chiml_gpu_download_emitter_state.argtypes = [vp, i, i, i, vp]
    L.chiml_gpu_download_emitter_pol.argtypes = [vp, i, i, vp]
    L.chiml_gpu_read_population.argtypes = [vp, i, i, vp, sz, C.POINTER(sz)]
    L.chiml_gpu_device_bytes.argtypes = [vp]
    L.chiml_gpu_device_bytes.restype = sz
    L.chiml_gpu_set_kernel_timing.argtypes = [vp, i]
    L.chiml_gpu_n_kernel_kinds.restype = i
    L.chiml_gpu_kernel_stat.argtypes = [vp, i, C.POINTER(KernelStat)]
    L.chiml_gpu_reset_kernel_stats.argtypes = [vp]
    _lib = L
    return L


def device_count() -> int:
    return lib().chiml_gpu_device_count()


def _ptr(a: Optional[np.ndarray]):
    if a is None or a.size == 0:
        return None
    return a.ctypes.data_as(C.c_void_p)


class GpuSim:
    """One y-slab of the propagator on one GPU, configured from a Plan (the reference's own lists)."""

    def __init__(self, plan: P.Plan, device: int = 0, detectors: bool = True, march=None, persistent: Optional[bool] = None, _part: str = "re"):
        """march: None (automatic column length), an int, or (fast, uniform) planes per column of the y-marching kernels.
        A plan with complex fields (plan.cplx) builds a pair: this object holds the real parts, self.imag the imaginary parts."""
        L = lib()
        self.plan = plan
        self.imag = None
        g = GridDesc()
        g.mode = plan.mode
        g.ln[:] = plan.ln
        g.d[:] = plan.d
        g.dt = plan.dt
        g.has_D, g.pml_on_D, g.n_objects, g.rank, g.nranks = plan.has_D, plan.pml_on_D, plan.n_objects, plan.rank, plan.nranks
        h = C.c_void_p()
        rc = L.chiml_gpu_create(C.byref(g), device, C.byref(h))
        if rc != 0:
            raise ChimlError(f"chiml_gpu_create: {STATUS.get(rc, rc)}: {L.chiml_gpu_last_error(None).decode()}")
        self.h = h
        self.steps_done = 0
        self.det_slots = []
        try:
            if plan.has_B:
                self._chk(L.chiml_gpu_set_magnetic(self.h, plan.has_B, plan.pml_on_B))
            for (kind, comp), runs in plan.lists.items():
                runs = np.ascontiguousarray(runs, dtype=P.RUN_DTYPE)
                self._chk(L.chiml_gpu_set_update_list(self.h, kind, comp, _ptr(runs), len(runs)))
            for o in plan.objects:
                a, x, gm, dp = (np.ascontiguousarray(v, dtype=np.float64) for v in (o.alpha, o.xi, o.gamma, o.dip))
                self._chk(L.chiml_gpu_set_object(self.h, o.obj, o.npoles, _ptr(a), _ptr(x), _ptr(gm), o.use_or_dip, _ptr(dp)))
            keep = []                                    # the engine reads the grids at commit: they must outlive this loop
            for (comp, pole), g in sorted(plan.dip_grids.items()):
                g = np.ascontiguousarray(g, dtype=np.float64)
                keep.append(g)
                self._chk(L.chiml_gpu_set_dip_grid(self.h, comp, pole, _ptr(g)))
            self._dip_keep = keep
            for obj, (a, x, gm) in sorted(plan.mag_objects.items()):
                a, x, gm = (np.ascontiguousarray(v, dtype=np.float64) for v in (a, x, gm))
                self._chk(L.chiml_gpu_set_object_magnetic(self.h, obj, len(a), _ptr(a), _ptr(x), _ptr(gm)))
            for obj, arrs in sorted(plan.chi_objects.items()):
                a, x, gm, gp = (np.ascontiguousarray(v, dtype=np.float64) for v in arrs)
                self._chk(L.chiml_gpu_set_object_chiral(self.h, obj, len(a), _ptr(a), _ptr(x), _ptr(gm), _ptr(gp)))
            if plan.prev_copy is not None:
                rows = np.ascontiguousarray(plan.prev_copy, dtype=np.int32)
                self._chk(L.chiml_gpu_set_prev_copy(self.h, _ptr(rows), len(rows)))
            for c in plan.cpml:
                psi = np.ascontiguousarray(c.psi, dtype=P.PSI_DTYPE)
                grid = np.ascontiguousarray(c.grid, dtype=P.GRIDP_DTYPE)
                self._chk(L.chiml_gpu_set_cpml(self.h, c.comp, c.part, c.has_psi, _ptr(psi), len(psi), _ptr(grid), len(grid)))
            for s in plan.sources:
                slot = C.c_int()
                self._chk(L.chiml_gpu_add_source(self.h, s.field, (C.c_int32 * 3)(*s.loc), (C.c_int32 * 3)(*s.sz), C.byref(slot)))
            keep = []
            for e in plan.emitters:
                slot = C.c_int()
                d = emitter_desc(e, keep)
                self._chk(L.chiml_gpu_add_emitters(self.h, C.byref(d), C.byref(slot)))
            for d in plan.dfts:
                lines = np.ascontiguousarray(d.lines, dtype=np.int32)
                slot = C.c_int()
                self._chk(L.chiml_gpu_add_dft(self.h, d.field, d.group, d.every, d.nfreq, d.npts, d.stride, _ptr(lines), len(lines), d.acc_len, C.byref(slot)))
            if detectors:
                for d in plan.detectors:
                    box = local_box(plan, d.loc, d.sz)
                    if box is None:
                        self.det_slots.append(-1)
                        continue
                    slot = C.c_int()
                    self._chk(L.chiml_gpu_add_detector(self.h, d.field, (C.c_int32 * 3)(*box[0]), (C.c_int32 * 3)(*box[1]), d.every, C.byref(slot)))
                    self.det_slots.append(slot.value)
            for comp, w in sorted(plan.periodic.items()):
                self._chk(L.chiml_gpu_set_periodic(self.h, comp, (C.c_int32 * 7)(*w)))
            for t in plan.tfsf:
                d = tfsf_surface(t, keep)
                self._chk(L.chiml_gpu_add_tfsf_surface(self.h, C.byref(d)))
            if plan.nranks > 1:
                self._chk(L.chiml_gpu_set_ordip_pole_count(self.h, plan.n_ordip_poles))
            if persistent is not None:
                self._chk(L.chiml_gpu_set_persistent(self.h, 1 if persistent else 0))
            if march is not None:
                mf, mu = (march, march) if isinstance(march, int) else march
                self._chk(L.chiml_gpu_set_march(self.h, mf, mu))
            self._chk(L.chiml_gpu_commit(self.h))
            if plan.cplx and _part == "re":
                self.imag = GpuSim(plan, device=device, detectors=detectors, march=march, persistent=persistent, _part="im")
                self._chk(L.chiml_gpu_bind_imag(self.h, self.imag.h, (C.c_double * 3)(*plan.k_point)))
        except Exception:
            self.close()
            raise

    def _chk(self, rc: int) -> None:
        if rc != 0:
            msg = lib().chiml_gpu_last_error(self.h).decode()
            raise ChimlError(f"{STATUS.get(rc, rc)}: {msg}")

    # ---- stepping -----------------------------------------------------------------------------------
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

    def step_n(self, n: int, amp: Optional[np.ndarray] = None) -> None:
        if amp is None:
            amp = self.src_amp(self.steps_done, n)
        amp = np.ascontiguousarray(amp, dtype=np.float64)
        if self.plan.cplx:
            amp_im = np.ascontiguousarray(self.src_amp_im(self.steps_done, n))
            ns = len(self.plan.sources)
            self._chk(lib().chiml_gpu_step_n_cplx(self.h, n, _ptr(amp) if ns else None, _ptr(amp_im) if ns else None))
            self.steps_done += n
            self.imag.steps_done += n
            return
        if self.plan.tfsf:
            tw = np.ascontiguousarray(P.dft_twiddles(self.plan, self.steps_done, n)) if self.plan.dfts else None
            rows = tfsf_rows(self.plan, self.steps_done, n)
            self._chk(lib().chiml_gpu_step_n_tfsf(self.h, n, _ptr(amp) if len(self.plan.sources) else None, _ptr(tw), _ptr(rows), rows.shape[1]))
        elif self.plan.dfts:
            tw = np.ascontiguousarray(P.dft_twiddles(self.plan, self.steps_done, n))
            self._chk(lib().chiml_gpu_step_n_dft(self.h, n, _ptr(amp) if len(self.plan.sources) else None, _ptr(tw)))
        else:
            self._chk(lib().chiml_gpu_step_n(self.h, n, _ptr(amp) if len(self.plan.sources) else None))
        self.steps_done += n

    def step_n_timed(self, n: int, amp: Optional[np.ndarray] = None) -> float:
        if amp is None:
            amp = self.src_amp(self.steps_done, n)
        amp = np.ascontiguousarray(amp, dtype=np.float64)
        ms = C.c_float()
        tw = np.ascontiguousarray(P.dft_twiddles(self.plan, self.steps_done, n)) if self.plan.dfts else None
        self._chk(lib().chiml_gpu_step_n_timed(self.h, n, _ptr(amp) if len(self.plan.sources) else None, _ptr(tw), C.byref(ms)))
        self.steps_done += n
        return float(ms.value)

    def sync(self) -> None:
        self._chk(lib().chiml_gpu_sync(self.h))

    # ---- state ----------------------------------------------------------------------------------------
    def _shape(self):
        lnx, lny, lnz = self.plan.ln
        return (lny, lnz, lnx)

    def field(self, f: int) -> np.ndarray:
        out = np.empty(self._shape(), dtype=np.float64)
        self._chk(lib().chiml_gpu_download_field(self.h, f, _ptr(out)))
        return out

    def set_field(self, f: int, a: np.ndarray) -> None:
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == self._shape()
        self._chk(lib().chiml_gpu_upload_field(self.h, f, _ptr(a)))

    def pole(self, comp: int, pole: int, prev: int = 0) -> np.ndarray:
        out = np.empty(self._shape(), dtype=np.float64)
        self._chk(lib().chiml_gpu_download_pole(self.h, comp, pole, prev, _ptr(out)))
        return out

    def set_pole(self, comp: int, pole: int, prev: int, a: np.ndarray) -> None:
        a = np.ascontiguousarray(a, dtype=np.float64)
        self._chk(lib().chiml_gpu_upload_pole(self.h, comp, pole, prev, _ptr(a)))

    def chi_pole(self, comp: int, pole: int, prev: int = 0) -> np.ndarray:
        out = np.empty(self._shape(), dtype=np.float64)
        self._chk(lib().chiml_gpu_download_chi_pole(self.h, comp, pole, prev, _ptr(out)))
        return out

    def prev_field(self, comp: int) -> np.ndarray:
        out = np.empty(self._shape(), dtype=np.float64)
        self._chk(lib().chiml_gpu_download_prev_field(self.h, comp, _ptr(out)))
        return out

    def mag_pole(self, comp: int, pole: int, prev: int = 0) -> np.ndarray:
        out = np.empty(self._shape(), dtype=np.float64)
        self._chk(lib().chiml_gpu_download_mag_pole(self.h, comp, pole, prev, _ptr(out)))
        return out

    def ordip_pole(self, comp: int, pole: int, prev: int = 0) -> np.ndarray:
        out = np.empty(self._shape(), dtype=np.float64)
        self._chk(lib().chiml_gpu_download_ordip_pole(self.h, comp, pole, prev, _ptr(out)))
        return out

    def psi(self, comp: int, part: int) -> np.ndarray:
        out = np.empty(self._shape(), dtype=np.float64)
        self._chk(lib().chiml_gpu_download_psi(self.h, comp, part, _ptr(out)))
        return out

    def detector(self, index: int) -> np.ndarray:
        """All samples of plan.detectors[index] so far: array (n_samples, sy, sz, sx)."""
        slot = self.det_slots[index]
        if slot < 0:
            return np.zeros((0, 0, 0, 0))
        n = C.c_size_t()
        self._chk(lib().chiml_gpu_read_detector(self.h, slot, None, 0, C.byref(n)))
        box = local_box(self.plan, self.plan.detectors[index].loc, self.plan.detectors[index].sz)
        sx, sy, sz = box[1]
        out = np.empty((n.value, sy, sz, sx), dtype=np.float64)
        self._chk(lib().chiml_gpu_read_detector(self.h, slot, _ptr(out), n.value, C.byref(n)))
        return out

    # ---- y-slabs -----------------------------------------------------------------------------------
    def halo_export(self) -> bytes:
        n = C.c_size_t()
        self._chk(lib().chiml_gpu_halo_export(self.h, None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        self._chk(lib().chiml_gpu_halo_export(self.h, buf, n.value, C.byref(n)))
        return buf.raw[:n.value]

    def halo_bind(self, dist, rank: int, world: int) -> None:
        """Exchanges the export blobs through torch.distributed (any backend) and binds the neighbours."""
        blobs = [None] * world
        dist.all_gather_object(blobs, self.halo_export())
        ring = bool(self.plan.periodic) and world > 1          # a periodic run closes the slabs into a ring (chiml_b200/slab.py)
        lower = blobs[(rank - 1) % world] if (rank > 0 or ring) else None
        upper = blobs[(rank + 1) % world] if (rank < world - 1 or ring) else None
        self._chk(lib().chiml_gpu_halo_bind(self.h, lower, len(lower) if lower else 0, upper, len(upper) if upper else 0))

    def detector_range(self, index: int, first: int, n: int, out: np.ndarray) -> int:
        """Copies samples [first, first+n) of plan.detectors[index] into `out` (host buffer); returns how many existed."""
        slot = self.det_slots[index]
        if slot < 0:
            return 0
        m = C.c_size_t()
        self._chk(lib().chiml_gpu_read_detector_range(self.h, slot, first, n, _ptr(out), C.byref(m)))
        return int(m.value)

    def consume_detector(self, index: int, upto: int) -> None:
        if self.det_slots[index] >= 0:
            self._chk(lib().chiml_gpu_consume_detector(self.h, self.det_slots[index], upto))

    def consume_population(self, slot: int, upto: int) -> None:
        self._chk(lib().chiml_gpu_consume_population(self.h, slot, upto))

    def reserve_steps(self, n: int) -> None:
        self._chk(lib().chiml_gpu_reserve_steps(self.h, n))

    def emitter_state(self, slot: int, sys: int, which: int) -> np.ndarray:
        """(nemit, N*N) complex: rho (which=0) or d rho/dt at n, n-1, n-2, n-3 (which=1..4) of level system `sys`."""
        e = self.plan.emitters[slot]
        out = np.empty((e.nemit, e.nlevel * e.nlevel, 2), dtype=np.float64)
        if e.nemit:          # a slab that holds only the rim of the object has a set without emitters
            self._chk(lib().chiml_gpu_download_emitter_state(self.h, slot, sys, which, _ptr(out)))
        return out[..., 0] + 1j * out[..., 1]

    def emitter_P(self, slot: int, comp: int) -> np.ndarray:
        e = self.plan.emitters[slot]
        out = np.empty((e.box_n[1] + 2, e.pz, e.box_n[0] + 2), dtype=np.float64)
        self._chk(lib().chiml_gpu_download_emitter_pol(self.h, slot, comp, _ptr(out)))
        return out

    def population(self, slot: int, det: int) -> np.ndarray:
        n = C.c_size_t()
        self._chk(lib().chiml_gpu_read_population(self.h, slot, det, None, 0, C.byref(n)))
        out = np.zeros((n.value, 2), dtype=np.float64)
        if n.value:
            self._chk(lib().chiml_gpu_read_population(self.h, slot, det, _ptr(out), n.value, C.byref(n)))
        return out[:, 0] + 1j * out[:, 1]

    def dft(self, slot: int) -> np.ndarray:
        """Complex accumulator of DFT set `slot` (fInReal_ + i fInCplx_)."""
        d = self.plan.dfts[slot]
        re, im = np.empty(d.acc_len), np.empty(d.acc_len)
        self._chk(lib().chiml_gpu_download_dft(self.h, slot, _ptr(re), _ptr(im)))
        return re + 1j * im

    def set_kernel_timing(self, on: bool) -> None:
        self._chk(lib().chiml_gpu_set_kernel_timing(self.h, 1 if on else 0))

    def reset_kernel_stats(self) -> None:
        self._chk(lib().chiml_gpu_reset_kernel_stats(self.h))

    def kernel_stats(self):
        """[{name, launches, timed_launches, ms_total, alg_bytes_per_launch}] for every kernel of the step loop."""
        out = []
        for k in range(lib().chiml_gpu_n_kernel_kinds()):
            ks = KernelStat()
            self._chk(lib().chiml_gpu_kernel_stat(self.h, k, C.byref(ks)))
            out.append({"name": ks.name.decode(), "launches": int(ks.launches), "timed_launches": int(ks.timed_launches),
                        "ms_total": float(ks.ms_total), "alg_bytes_per_launch": float(ks.alg_bytes_per_launch),
                        "alg_bytes_per_step": float(ks.alg_bytes_per_step)})
        return out

    def launch_count(self) -> int:
        return int(lib().chiml_gpu_launch_count(self.h))

    def device_bytes(self) -> int:
        return int(lib().chiml_gpu_device_bytes(self.h))

    def close(self) -> None:
        # (the imaginary part of a pair runs on the real part's stream: it goes first)
        if getattr(self, "imag", None) is not None:
            self.imag.close()
            self.imag = None
        if getattr(self, "h", None):
            lib().chiml_gpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def local_box(plan: P.Plan, gloc, gsz):
    """Intersection of a box given in global grid coordinates with this rank's slab, in local
    ghost-inclusive coordinates (the +1 ghost offset of parallelGrid, y shifted by y_start)."""
    x0, y0, z0 = gloc
    sx, sy, sz = gsz
    ly0 = y0 - plan.y_start + 1
    ly1 = ly0 + sy
    lo, hi = max(ly0, 1), min(ly1, plan.ln[1] - 1)
    if hi <= lo:
        return None
    zl = z0 + 1 if plan.ln[2] > 1 else 0
    return (x0 + 1, lo, zl), (sx, hi - lo, sz if plan.ln[2] > 1 else 1)
