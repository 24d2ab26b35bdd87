"""Reader / writer for plan files (include/chiml_plan.h): the flattened per-rank output of the
propagator constructor -- update lists, pole constants, CPML lists, sources, detectors -- i.e.
exactly what the C ABI in include/chiml_gpu.h consumes.

The record layouts mirror the reference's own PODs:
  ChimlRun        == std::pair<std::array<int,6>, std::array<double,4>>  (reference UTIL/typedefs.hpp:14)
  ChimlPsiParams  == updatePsiParams   (reference PML/parallelPML.hpp:18-26)
  ChimlGridParams == updateGridParams  (reference PML/parallelPML.hpp:30-38)
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

PLAN_VERSION = 1

MODE_TE, MODE_TM, MODE_3D = 0, 1, 2
LIST_U, LIST_D, LIST_LORD, LIST_ORDIPD, LIST_ORDIPP, LIST_CHID = 0, 1, 2, 3, 4, 5
FIELD_NAMES = ["Ex", "Ey", "Ez", "Hx", "Hy", "Hz", "Dx", "Dy", "Dz", "Bx", "By", "Bz"]

RUN_DTYPE = np.dtype([("n", "<i4"), ("ind", "<i4"), ("ind_i", "<i4"), ("ind_j", "<i4"), ("ind_k", "<i4"),
                      ("obj", "<i4"), ("pf", "<f8", (4,))])
PSI_DTYPE = np.dtype([("transSz", "<i4"), ("stride", "<i4"), ("ind", "<i4"), ("indOff", "<i4"),
                      ("b", "<f8"), ("c", "<f8")])
GRIDP_DTYPE = np.dtype([("nAx", "<i4"), ("stride", "<i4"), ("ind", "<i4"), ("indOff", "<i4"),
                        ("Db", "<f8"), ("DbField", "<f8")])
assert RUN_DTYPE.itemsize == 56 and PSI_DTYPE.itemsize == 32 and GRIDP_DTYPE.itemsize == 32

_GRID_FMT = "<i3i3dd5i4x" + "i3iiiiid"   # ChimlGridDesc (natural alignment, 72 bytes) + packed ChimlPlanGrid tail
_GRID_SIZE = struct.calcsize(_GRID_FMT)


@dataclass
class PlanObject:
    obj: int
    npoles: int
    use_or_dip: int
    ml: int
    eps_inf: float
    mu_inf: float
    alpha: np.ndarray
    xi: np.ndarray
    gamma: np.ndarray
    dip: np.ndarray  # (npoles, 3)


@dataclass
class PlanCpml:
    comp: int
    part: int
    has_psi: int
    psi: np.ndarray
    grid: np.ndarray


@dataclass
class PlanSource:
    field: int
    loc: Tuple[int, int, int]
    sz: Tuple[int, int, int]
    amp: np.ndarray
    amp_im: Optional[np.ndarray] = None     # complex fields: dt * Im(sum pulse(t_k)) (record SRCIMAG)


@dataclass
class PlanDetector:
    detector: int
    field: int
    loc: Tuple[int, int, int]     # global coordinates of the stored (grown) box
    sz: Tuple[int, int, int]
    offset: Tuple[int, int, int]
    every: int
    type: int
    conv: float
    t_conv: float


@dataclass
class PlanEmitter:
    """One parallelQE object restricted to this slab (include/chiml_gpu.h ChimlEmitterDesc)."""
    object: int
    nlevel: int
    nsys: int
    nemit: int
    box_lo: Tuple[int, int, int]
    box_n: Tuple[int, int, int]
    pz: int
    npop: int
    pop_every: int
    npoints: int
    dt: float
    inv_hbar: float
    na: float
    h0: np.ndarray        # (nsys, N*N) complex
    weight: np.ndarray    # (nsys,)
    mu: np.ndarray        # (3, N*N) complex
    gam_ptr: np.ndarray   # (N*N+1,) int32
    gam_col: np.ndarray
    gam_val: np.ndarray
    loc: np.ndarray       # (nemit, 3) int32
    eps: np.ndarray       # (n1+2, pz, n0+2)
    pop_level: np.ndarray


@dataclass
class PlanDft:
    """One stored field of a flux object (include/chiml_gpu.h chiml_gpu_add_dft)."""
    field: int
    group: int
    every: int
    nfreq: int
    npts: int
    stride: int
    acc_len: int
    freq: np.ndarray      # (nfreq,) the group's freqList_ (angular, FDTD units)
    lines: np.ndarray     # (nlines, 2) int32: grid index, accumulator index


@dataclass
class PlanTfsfSurface:
    """One TFSF surface (include/chiml_gpu.h ChimlTfsfSurface)."""
    comp: int
    incd_offset: int
    incd_len: int
    n: int
    stride_incd: int
    stride_main: int
    prefactor: float
    pairs_D: np.ndarray   # (npairs, 2) int32: incident index, main-grid index
    pairs_U: np.ndarray
    ep_mu: Optional[np.ndarray]


_EMIT_FMT = "<4i3i3i4i2i3d"
_EMIT_SIZE = struct.calcsize(_EMIT_FMT)
assert _EMIT_SIZE == 88


@dataclass
class Plan:
    mode: int = MODE_3D
    ln: Tuple[int, int, int] = (0, 0, 0)
    d: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    dt: float = 0.0
    has_D: int = 0
    pml_on_D: int = 0
    n_objects: int = 1
    rank: int = 0
    nranks: int = 1
    y_start: int = 0
    n_global: Tuple[int, int, int] = (0, 0, 0)
    n_steps: int = 0
    n_lor_poles: int = 0
    n_ordip_poles: int = 0
    t_max: float = 0.0
    lists: Dict[Tuple[int, int], np.ndarray] = field(default_factory=dict)   # (kind, comp) -> RUN_DTYPE array
    objects: List[PlanObject] = field(default_factory=list)
    cpml: List[PlanCpml] = field(default_factory=list)
    sources: List[PlanSource] = field(default_factory=list)
    detectors: List[PlanDetector] = field(default_factory=list)
    emitters: List[PlanEmitter] = field(default_factory=list)
    dfts: List[PlanDft] = field(default_factory=list)
    has_B: int = 0                              # magnetic-dispersive media (record MAGNETIC): B grids exist
    pml_on_B: int = 0
    n_mag_poles: int = 0
    mag_objects: Dict[int, Tuple[np.ndarray, np.ndarray, np.ndarray]] = field(default_factory=dict)   # obj -> (magAlpha, magXi, magGamma)
    chi_objects: Dict[int, Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]] = field(default_factory=dict)   # obj -> (chiAlpha, chiXi, chiGamma, chiGammaPrev)
    dip_grids: Dict[Tuple[int, int], np.ndarray] = field(default_factory=dict)    # (comp, pole) -> dipP_[comp][pole], the whole ghost-inclusive grid (REL_TO_NORM orientations)
    prev_copy: Optional[np.ndarray] = None      # (nrows, 4) int32: copy2PrevFields_ rows {length, x, y, z}
    cplx: bool = False                          # complex fields (record COMPLEX): real and imaginary parts are two field sets over the same lists
    k_point: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    tfsf: List[PlanTfsfSurface] = field(default_factory=list)
    tfsf_lines: Optional[np.ndarray] = None     # (n_steps, per_step): the incident-line table of chiml_gpu_step_n_tfsf
    periodic: Dict[int, Tuple[int, ...]] = field(default_factory=dict)      # comp -> (nx, ny, nz, xmax, ymax, zmin, zmax) (ChimlWrap)

    @property
    def n_chi_poles(self) -> int:
        return max([len(a[0]) for a in self.chi_objects.values()] + [0])

    @property
    def ncell(self) -> int:
        return self.ln[0] * self.ln[1] * self.ln[2]

    def fields_present(self) -> List[int]:
        """E/H(/D) components that exist in this mode (parallelFDTDField.hpp:391-443)."""
        if self.mode == MODE_TE:
            f = [0, 1, 5]
        elif self.mode == MODE_TM:
            f = [2, 3, 4]
        else:
            f = [0, 1, 2, 3, 4, 5]
        if self.has_D:
            f += [6 + c for c in f if c < 3]
        if self.has_B:
            f += [9 + (c - 3) for c in f if 3 <= c < 6]
        return f

    def get_list(self, kind: int, comp: int) -> np.ndarray:
        return self.lists.get((kind, comp), np.zeros(0, dtype=RUN_DTYPE))


def read_plan(path: str) -> Plan:
    data = open(path, "rb").read()
    pos = 0
    plan = Plan()
    first = True
    while pos < len(data):
        tag = data[pos:pos + 8].decode("ascii").strip()
        (nbytes,) = struct.unpack_from("<Q", data, pos + 8)
        payload = data[pos + 16:pos + 16 + nbytes]
        pos += 16 + nbytes
        if first:
            if tag != "CHIMLPLN":
                raise ValueError(f"{path}: not a plan file")
            (ver,) = struct.unpack_from("<i", payload, 0)
            if ver != PLAN_VERSION:
                raise ValueError(f"{path}: plan version {ver} != {PLAN_VERSION}")
            first = False
            continue
        if tag == "GRID":
            v = struct.unpack_from(_GRID_FMT, payload, 0)
            plan.mode = v[0]; plan.ln = tuple(v[1:4]); plan.d = tuple(v[4:7]); plan.dt = v[7]
            plan.has_D, plan.pml_on_D, plan.n_objects, plan.rank, plan.nranks = v[8:13]
            plan.y_start = v[13]; plan.n_global = tuple(v[14:17]); plan.n_steps = v[17]
            plan.n_lor_poles = v[18]; plan.n_ordip_poles = v[19]; plan.t_max = v[21]
        elif tag == "UPLIST":
            kind, comp, n = struct.unpack_from("<iiQ", payload, 0)
            plan.lists[(kind, comp)] = np.frombuffer(payload, dtype=RUN_DTYPE, count=n, offset=16).copy()
        elif tag == "OBJECT":
            obj, npoles, use_or_dip, ml, eps_inf, mu_inf = struct.unpack_from("<iiiidd", payload, 0)
            off = 32
            arrs = []
            for _ in range(3):
                arrs.append(np.frombuffer(payload, dtype="<f8", count=npoles, offset=off).copy()); off += 8 * npoles
            dip = np.frombuffer(payload, dtype="<f8", count=3 * npoles, offset=off).copy().reshape(npoles, 3)
            plan.objects.append(PlanObject(obj, npoles, use_or_dip, ml, eps_inf, mu_inf, arrs[0], arrs[1], arrs[2], dip))
        elif tag == "CPML":
            comp, part, has_psi, _pad, npsi, ngrid = struct.unpack_from("<iiiiQQ", payload, 0)
            psi = np.frombuffer(payload, dtype=PSI_DTYPE, count=npsi, offset=32).copy()
            grid = np.frombuffer(payload, dtype=GRIDP_DTYPE, count=ngrid, offset=32 + 32 * npsi).copy()
            plan.cpml.append(PlanCpml(comp, part, has_psi, psi, grid))
        elif tag == "SOURCE":
            v = struct.unpack_from("<i3i3ii", payload, 0)
            amp = np.frombuffer(payload, dtype="<f8", count=v[7], offset=32).copy()
            plan.sources.append(PlanSource(v[0], tuple(v[1:4]), tuple(v[4:7]), amp))
        elif tag == "DETECTOR":
            v = struct.unpack_from("<ii3i3i3iiiidd", payload, 0)
            plan.detectors.append(PlanDetector(v[0], v[1], tuple(v[2:5]), tuple(v[5:8]), tuple(v[8:11]), v[11], v[12], v[14], v[15]))
        elif tag == "DFT":
            fld, group, every, nfreq, npts, stride, nlines, acc_len = struct.unpack_from("<6iQQ", payload, 0)
            freq = np.frombuffer(payload, dtype="<f8", count=nfreq, offset=40).copy()
            lines = np.frombuffer(payload, dtype="<i4", count=2 * nlines, offset=40 + 8 * nfreq).copy().reshape(nlines, 2)
            plan.dfts.append(PlanDft(fld, group, every, nfreq, npts, stride, acc_len, freq, lines))
        elif tag == "MAGNETIC":
            plan.has_B, plan.pml_on_B, plan.n_mag_poles, _ = struct.unpack_from("<4i", payload, 0)
        elif tag == "OBJMAG":
            obj, np_ = struct.unpack_from("<ii", payload, 0)
            arrs = [np.frombuffer(payload, dtype="<f8", count=np_, offset=8 + 8 * np_ * k).copy() for k in range(3)]
            plan.mag_objects[obj] = (arrs[0], arrs[1], arrs[2])
        elif tag == "OBJCHI":
            obj, np_ = struct.unpack_from("<ii", payload, 0)
            arrs = [np.frombuffer(payload, dtype="<f8", count=np_, offset=8 + 8 * np_ * k).copy() for k in range(4)]
            plan.chi_objects[obj] = tuple(arrs)
        elif tag == "DIPGRID":
            comp, pole = struct.unpack_from("<ii", payload, 0)
            plan.dip_grids[(comp, pole)] = np.frombuffer(payload, dtype="<f8", offset=8).copy()
        elif tag == "PREVCOPY":
            (nr,) = struct.unpack_from("<Q", payload, 0)
            plan.prev_copy = np.frombuffer(payload, dtype="<i4", count=4 * nr, offset=8).copy().reshape(nr, 4)
        elif tag == "COMPLEX":
            v = struct.unpack_from("<ii3d", payload, 0)
            plan.cplx = bool(v[0]); plan.k_point = tuple(v[2:5])
        elif tag == "SRCIMAG":
            (ns,) = struct.unpack_from("<i", payload, 0)
            plan.sources[-1].amp_im = np.frombuffer(payload, dtype="<f8", count=ns, offset=4).copy()
        elif tag == "TFSFSURF":
            v = struct.unpack_from("<10id", payload, 0)
            off = 48
            pd = np.frombuffer(payload, dtype="<i4", count=2 * v[6], offset=off).copy().reshape(v[6], 2); off += 8 * v[6]
            pu = np.frombuffer(payload, dtype="<i4", count=2 * v[7], offset=off).copy().reshape(v[7], 2); off += 8 * v[7]
            em = np.frombuffer(payload, dtype="<f8", count=v[2], offset=off).copy() if v[8] else None
            plan.tfsf.append(PlanTfsfSurface(v[0], v[1], v[2], v[3], v[4], v[5], v[10], pd, pu, em))
        elif tag == "TFSFLINE":
            ns, per = struct.unpack_from("<ii", payload, 0)
            plan.tfsf_lines = np.frombuffer(payload, dtype="<f8", count=ns * per, offset=8).copy().reshape(ns, per)
        elif tag == "PERIODIC":
            v = struct.unpack_from("<8i", payload, 0)
            plan.periodic[v[0]] = tuple(v[1:8])
        elif tag == "EMITTER":
            v = struct.unpack_from(_EMIT_FMT, payload, 0)
            obj, N, nsys, nemit = v[0:4]
            lo, bn = tuple(v[4:7]), tuple(v[7:10])
            nnz, npop, pop_every, npoints, pz = v[10], v[11], v[12], v[13], v[14]
            dt, inv_hbar, na = v[16], v[17], v[18]
            off = _EMIT_SIZE
            n2 = N * N

            def take(dtype, count):
                nonlocal off
                a = np.frombuffer(payload, dtype=dtype, count=count, offset=off).copy()
                off += a.nbytes
                return a
            h0 = take("<c16", nsys * n2).reshape(nsys, n2)
            weight = take("<f8", nsys)
            mu = take("<c16", 3 * n2).reshape(3, n2)
            gptr = take("<i4", n2 + 1)
            gcol = take("<i4", nnz)
            gval = take("<f8", nnz)
            loc = take("<i4", 3 * nemit).reshape(nemit, 3)
            eps = take("<f8", (bn[0] + 2) * (bn[1] + 2) * pz).reshape(bn[1] + 2, pz, bn[0] + 2)
            pl = take("<i4", npop)
            plan.emitters.append(PlanEmitter(obj, N, nsys, nemit, lo, bn, pz, npop, pop_every, npoints, dt, inv_hbar, na, h0, weight, mu,
                                             gptr, gcol, gval, loc, eps, pl))
    return plan


def dft_twiddles(plan: "Plan", start: int, n: int) -> np.ndarray:
    """exp(-i freq t) for steps start .. start+n-1 (t = time after the step, as parallelFluxDTC::fieldIn gets tcur_,
    DTC/parallelFlux.hpp:298): shape (n, sum of nfreq over groups, 2), evaluated like the reference (std::exp of a complex)."""
    groups = {}
    for d in plan.dfts:
        groups[d.group] = d.freq
    if not groups:
        return np.zeros((n, 0, 2))
    freq = np.concatenate([groups[g] for g in sorted(groups)])
    out = np.empty((n, len(freq), 2))
    t = 0.0
    for k in range(start + n):          # tcur_ is accumulated step by step (tcur_ += dt_)
        t += plan.dt
        if k >= start:
            # std::exp(cplx(0, y)) = (cos y, sin y) from libm: use the same library functions, not numpy's own complex exp
            for j, fr in enumerate(freq):
                y = -1.0 * t * float(fr)
                out[k - start, j, 0], out[k - start, j, 1] = math.cos(y), math.sin(y)
    return out


def _rec(tag: str, payload: bytes) -> bytes:
    return tag.ljust(8).encode("ascii") + struct.pack("<Q", len(payload)) + payload


def write_plan(path: str, plan: Plan) -> None:
    out = [_rec("CHIMLPLN", struct.pack("<i", PLAN_VERSION))]
    out.append(_rec("GRID", struct.pack(_GRID_FMT, plan.mode, *plan.ln, *plan.d, plan.dt, plan.has_D, plan.pml_on_D,
                                        plan.n_objects, plan.rank, plan.nranks, plan.y_start, *plan.n_global,
                                        plan.n_steps, plan.n_lor_poles, plan.n_ordip_poles, 0, plan.t_max)))
    for (kind, comp), runs in sorted(plan.lists.items()):
        runs = np.ascontiguousarray(runs, dtype=RUN_DTYPE)
        out.append(_rec("UPLIST", struct.pack("<iiQ", kind, comp, len(runs)) + runs.tobytes()))
    for o in plan.objects:
        out.append(_rec("OBJECT", struct.pack("<iiiidd", o.obj, o.npoles, o.use_or_dip, o.ml, o.eps_inf, o.mu_inf)
                        + np.asarray(o.alpha, "<f8").tobytes() + np.asarray(o.xi, "<f8").tobytes()
                        + np.asarray(o.gamma, "<f8").tobytes() + np.asarray(o.dip, "<f8").tobytes()))
    for c in plan.cpml:
        out.append(_rec("CPML", struct.pack("<iiiiQQ", c.comp, c.part, c.has_psi, 0, len(c.psi), len(c.grid))
                        + np.ascontiguousarray(c.psi, PSI_DTYPE).tobytes() + np.ascontiguousarray(c.grid, GRIDP_DTYPE).tobytes()))
    if plan.has_B:
        out.append(_rec("MAGNETIC", struct.pack("<4i", plan.has_B, plan.pml_on_B, plan.n_mag_poles, 0)))
        for obj, (a, x, g) in sorted(plan.mag_objects.items()):
            out.append(_rec("OBJMAG", struct.pack("<ii", obj, len(a)) + np.asarray(a, "<f8").tobytes() + np.asarray(x, "<f8").tobytes() + np.asarray(g, "<f8").tobytes()))
    for obj, arrs in sorted(plan.chi_objects.items()):
        out.append(_rec("OBJCHI", struct.pack("<ii", obj, len(arrs[0])) + b"".join(np.asarray(a, "<f8").tobytes() for a in arrs)))
    for (comp, pole), g in sorted(plan.dip_grids.items()):
        out.append(_rec("DIPGRID", struct.pack("<ii", comp, pole) + np.ascontiguousarray(g, "<f8").tobytes()))
    if plan.prev_copy is not None:
        out.append(_rec("PREVCOPY", struct.pack("<Q", len(plan.prev_copy)) + np.ascontiguousarray(plan.prev_copy, "<i4").tobytes()))
    if plan.cplx:
        out.append(_rec("COMPLEX", struct.pack("<ii3d", 1, 0, *plan.k_point)))
    for s in plan.sources:
        out.append(_rec("SOURCE", struct.pack("<i3i3ii", s.field, *s.loc, *s.sz, len(s.amp)) + np.asarray(s.amp, "<f8").tobytes()))
        if s.amp_im is not None:
            out.append(_rec("SRCIMAG", struct.pack("<i", len(s.amp_im)) + np.asarray(s.amp_im, "<f8").tobytes()))
    for d in plan.detectors:
        out.append(_rec("DETECTOR", struct.pack("<ii3i3i3iiiidd", d.detector, d.field, *d.loc, *d.sz, *d.offset, d.every, d.type, 0,
                                                d.conv, d.t_conv)))
    for e in plan.emitters:
        out.append(_rec("EMITTER", struct.pack(_EMIT_FMT, e.object, e.nlevel, e.nsys, e.nemit, *e.box_lo, *e.box_n, len(e.gam_col), e.npop,
                                               e.pop_every, e.npoints, e.pz, 0, e.dt, e.inv_hbar, e.na)
                        + np.ascontiguousarray(e.h0, "<c16").tobytes() + np.ascontiguousarray(e.weight, "<f8").tobytes()
                        + np.ascontiguousarray(e.mu, "<c16").tobytes() + np.ascontiguousarray(e.gam_ptr, "<i4").tobytes()
                        + np.ascontiguousarray(e.gam_col, "<i4").tobytes() + np.ascontiguousarray(e.gam_val, "<f8").tobytes()
                        + np.ascontiguousarray(e.loc, "<i4").tobytes() + np.ascontiguousarray(e.eps, "<f8").tobytes()
                        + np.ascontiguousarray(e.pop_level, "<i4").tobytes()))
    for d in plan.dfts:
        out.append(_rec("DFT", struct.pack("<6iQQ", d.field, d.group, d.every, d.nfreq, d.npts, d.stride, len(d.lines), d.acc_len)
                        + np.ascontiguousarray(d.freq, "<f8").tobytes() + np.ascontiguousarray(d.lines, "<i4").tobytes()))
    for t in plan.tfsf:
        out.append(_rec("TFSFSURF", struct.pack("<10id", t.comp, t.incd_offset, t.incd_len, t.n, t.stride_incd, t.stride_main, len(t.pairs_D), len(t.pairs_U),
                                                1 if t.ep_mu is not None else 0, 0, t.prefactor)
                        + np.ascontiguousarray(t.pairs_D, "<i4").tobytes() + np.ascontiguousarray(t.pairs_U, "<i4").tobytes()
                        + (np.ascontiguousarray(t.ep_mu, "<f8").tobytes() if t.ep_mu is not None else b"")))
    if plan.tfsf_lines is not None:
        out.append(_rec("TFSFLINE", struct.pack("<ii", *plan.tfsf_lines.shape) + np.ascontiguousarray(plan.tfsf_lines, "<f8").tobytes()))
    for comp, w in sorted(plan.periodic.items()):
        out.append(_rec("PERIODIC", struct.pack("<8i", comp, *w)))
    with open(path, "wb") as f:
        f.write(b"".join(out))


def read_dump(path: str) -> Dict[Tuple[int, str], Tuple[Tuple[int, int, int], int, np.ndarray]]:
    """Field dump written by oracle/ref_driver.cpp --dump: {(rank, name): (ln, y_start, array[y][z][x])}."""
    data = open(path, "rb").read()
    if data[:8] != b"CHIMLDMP":
        raise ValueError(f"{path}: not a field dump")
    pos = 12
    out = {}
    while pos < len(data):
        rank = struct.unpack_from("<i", data, pos)[0]
        name = data[pos + 4:pos + 20].split(b"\0")[0].decode()
        lnx, lny, lnz, ys = struct.unpack_from("<4i", data, pos + 20)
        n = lnx * lny * lnz
        arr = np.frombuffer(data, dtype="<f8", count=n, offset=pos + 36).copy().reshape(lny, lnz, lnx)
        out[(rank, name)] = ((lnx, lny, lnz), ys, arr)
        pos += 36 + 8 * n
    return out
