"""Builders for chiML JSON inputs (the input contract, reference py_utilities/write_json.py:5-594 and
INPUTS/parallelInputs.cpp:12-840) and the synthetic configurations of BASELINE.md (C1..C5), scalable
so that the same geometry can be run by the CPU reference at small size and by the GPU at full size.

Units follow the reference: lengths in units of `a`, time in a/c, c = eps0 = mu0 = 1.  The source
entries are written by hand with a "loc" key because write_json.write_normal_src emits loc_x/loc_y
while the parser reads "loc" (write_json.py:483 vs parallelInputs.cpp:198).
"""
from __future__ import annotations

import json
import math
from typing import Dict, List, Optional, Sequence

EV_TO_HZ = 1.0 / 4.135666e-15      # parallelInputs.cpp:1181-1184 ev2FDTD
SPEED_OF_LIGHT = 299792458.0       # UTIL/ml_consts.hpp


def comp_cell(size: Sequence[float], res: int, t_lim: float, pol: str, courant: float = 0.5, a: float = 1e-7,
              pbc: bool = False) -> Dict:
    return {"InputMaps_x": [], "InputMaps_y": [], "InputMaps_z": [], "procs": 1, "size": list(size), "res": res,
            "courant": courant, "tLim": t_lim, "PBC": pbc, "pol": pol, "E_max": 1.0, "a": a, "cplxFields": False}


def pml(thickness: Sequence[float], a_max: float = 0.25, ma: float = 1.0, m: float = 3.0, sig_opt_rat: float = 1.0,
        kappa_max: float = 1.0) -> Dict:
    return {"aMax": a_max, "ma": ma, "m": m, "thickness": list(thickness), "sigOptRat": sig_opt_rat, "kappaMax": kappa_max}


def gaussian_pulse(fcen: float, fwidth: float, cutoff: float = 5.0, intensity: float = 1.0, t_0: Optional[float] = None) -> Dict:
    if t_0 is None:
        t_0 = cutoff / fwidth   # BASELINE.md / SURVEY.md A.6: set t_0 explicitly
    return {"profile": "gaussian", "Field_Intensity": intensity, "fcen": fcen, "fwidth": fwidth, "t_0": t_0, "cutoff": cutoff}


def normal_source(pol: str, loc: Sequence[float], size: Sequence[float], pulses: List[Dict]) -> Dict:
    return {"PulseList": pulses, "size": list(size), "pol": pol, "loc": list(loc)}


def lorentz_pole(sigma_p: float, gamma: float, omega: float, dip_or_e: str = "isotropic", dir_dip_e: Sequence[float] = (0.0, 0.0, 0.0),
                 sigma_m: float = 0.0, tau: float = 0.0, pol_ang_e: Optional[float] = None, az_ang_e: Optional[float] = None, tan_iso: bool = False,
                 dip_or_m: str = "isotropic") -> Dict:
    """sigma_m > 0 makes the pole a magnetic one as well (magnetisation M driven by H, OBJECTS/Obj.cpp setUpConsts); tau != 0 a chiral one
    (cross terms between E and H, chiAlpha / chiXi / chiGamma / chiGammaPrev, OBJECTS/Obj.cpp:345-353).  dip_or_e "normal" / "tangent" /
    "rel_norm" orient the dipole relative to the object's surface normal (polar / azimuthal angles in degrees, INPUTS/parallelInputs.cpp:1329-1357;
    tan_iso splits the pole into a lateral and a longitudinal tangent one)."""
    d = {"dipOrE": dip_or_e, "dipOrM": dip_or_m, "sigma_p": sigma_p, "sigma_m": sigma_m, "tau": tau, "gamma": gamma,
         "omega": omega, "dirDipE": list(dir_dip_e), "dirDipM": [0.0, 0.0, 0.0]}
    if pol_ang_e is not None:
        d["polAngRelNormE"] = pol_ang_e
    if az_ang_e is not None:
        d["azAngRelNormE"] = az_ang_e
    if tan_iso:
        d["tanIso"] = True
    return d


def ev_to_fdtd(ev: float, a: float = 1e-7) -> float:
    return ev * EV_TO_HZ * a / SPEED_OF_LIGHT


def drude_pole(omega_p_ev: float, gamma_ev: float, a: float = 1e-7) -> Dict:
    """A Drude pole expressed the way getMetal does (parallelInputs.cpp:1187-1221): a Lorentz pole whose
    resonance is ev2FDTD(1e-20); JSON 'gamma'/'omega' are multiplied by pi / 2 pi by the parser (:1300-1303)."""
    wp = ev_to_fdtd(omega_p_ev, a)
    omg = ev_to_fdtd(1.0e-20, a)
    return lorentz_pole(sigma_p=(wp / omg) ** 2, gamma=ev_to_fdtd(gamma_ev, a), omega=omg)


def block(size: Sequence[float], loc: Sequence[float], material: str = "custom", eps: float = 1.0, pols: Optional[List[Dict]] = None) -> Dict:
    return {"shape": "block", "material": material, "loc": list(loc), "size": list(size), "orPhi": 0.0, "orTheta": 90.0,
            "unit_vectors": [], "pols": pols or [], "eps": eps, "mu": 1.0, "tellegen": 0.0, "Basis_Set": []}


def sphere(radius: float, loc: Sequence[float], material: str = "custom", eps: float = 1.0, pols: Optional[List[Dict]] = None) -> Dict:
    return {"shape": "sphere", "material": material, "loc": list(loc), "radius": radius, "unit_vectors": [], "pols": pols or [],
            "eps": eps, "mu": 1.0, "tellegen": 0.0, "Basis_Set": []}


def ml_block(size: Sequence[float], loc: Sequence[float], mol_den: float, e_levels_ev: Sequence[float], dipole_debye: float,
             relax_rate: float, dephasing_rate: float, dtc_levs: Sequence[int] = (), pop_fname_base: str = "output_data/qe_") -> Dict:
    """Two-level emitter block (write_json.write_ml_block): basis (l,m) = (0,0),(1,0), delta-function levels,
    couplings [0, mu, mu, 0], one relaxation 1 -> 0."""
    o = block(size, loc)
    o["Basis_Set"] = [{"l": 0, "m": 0}, {"l": 1, "m": 0}]
    o["mol_den"] = mol_den
    o["Energy_Levels"] = [{"distribution": "delta_fxn", "E_cen": [e], "weights": [1.0], "nstates": 1, "levs_described": 1} for e in e_levels_ev]
    o["couplings"] = [0.0, dipole_debye, dipole_debye, 0.0]
    o["gam"] = []
    o["RelaxationOperators"] = [{"state_i": 1, "state_f": 0, "rate": relax_rate, "del_omg": 0.0, "radiative": False, "dephasing_rate": dephasing_rate}]
    o["dtc_levs"] = list(dtc_levs)
    o["levDTC_timeInt"] = 1
    o["dtc_pop_fname_base"] = pop_fname_base
    o["output_pol"] = False
    return o


def ml_object(size: Sequence[float], loc: Sequence[float], mol_den: float, basis: Sequence[Sequence[int]], levels: List[Dict],
              couplings_debye: Sequence[float], relaxations: List[Dict], eps: float = 1.0, dtc_levs: Sequence[int] = (),
              pop_fname_base: str = "output_data/qe_", pop_every: int = 1) -> Dict:
    """General emitter block: `basis` = [(l, m), ...]; `levels` = Energy_Levels entries ({"E_cen": [...eV], "weights": [...],
    "levs_described": k}); `couplings_debye` = N*N transition dipoles; `relaxations` = [{"state_i", "state_f", "rate", "dephasing_rate"}]
    (rates in s^-1) -- INPUTS/parallelInputs.cpp:427-666."""
    o = block(size, loc, eps=eps)
    o["Basis_Set"] = [{"l": int(l), "m": int(m)} for l, m in basis]
    o["mol_den"] = mol_den
    o["Energy_Levels"] = [{"distribution": "delta_fxn", "E_cen": list(lv["E_cen"]), "weights": list(lv.get("weights", [1.0] * len(lv["E_cen"]))),
                           "nstates": 1, "levs_described": int(lv.get("levs_described", 1))} for lv in levels]
    o["couplings"] = list(couplings_debye)
    o["gam"] = []
    o["RelaxationOperators"] = [{"state_i": r["state_i"], "state_f": r["state_f"], "rate": r["rate"], "del_omg": 0.0, "radiative": False,
                                 "dephasing_rate": r.get("dephasing_rate", 0.0)} for r in relaxations]
    o["dtc_levs"] = list(dtc_levs)
    o["levDTC_timeInt"] = pop_every
    o["dtc_pop_fname_base"] = pop_fname_base
    o["output_pol"] = False
    return o


def detector(loc: Sequence[float], size: Sequence[float], typ: str, fname: str, dtc_class: str = "txt", time_int: float = 0.0, si: bool = False) -> Dict:
    return {"loc": list(loc), "size": list(size), "SI": si, "dtc_class": dtc_class, "fname": fname, "type": typ, "txt_dat_type": "real",
            "txt_format_type": "none", "Time_Interval": time_int, "timeIntegrateMap": False, "t_start": 0.0, "t_end": 1e8}


def freq_detector(loc: Sequence[float], size: Sequence[float], typ: str, fname: str, fcen: float, fwidth: float, nfreq: int, time_int: float = 0.0,
                  si: bool = False, output_map: bool = False) -> Dict:
    """A frequency-domain detector (dtc_class "freq", parsed at INPUTS/parallelInputs.cpp:671-751): running DFT of the fields over a box."""
    d = detector(loc, size, typ, fname, dtc_class="freq", time_int=time_int, si=si)
    d.update({"fcen": fcen, "fwidth": fwidth, "nfreq": nfreq, "output_map": output_map})
    return d


def flux(name: str, loc: Sequence[float], size: Sequence[float], fcen: float, fwidth: float, nfreq: int, weight: float = 1.0) -> Dict:
    return {"name": name, "save": False, "load": False, "loc": list(loc), "size": list(size), "SI": False, "fcen": fcen, "fwidth": fwidth,
            "nfreq": nfreq, "weight": weight, "cross_sec": False}


def tfsf(size: Sequence[float], loc: Sequence[float], pulses: List[Dict], m: Sequence[int] = (1, 0, 0), psi: float = 90.0, pml_thick: int = 10,
         pml_a_max: float = 0.0, pml_m: float = 3.0, pml_ma: float = 1.0) -> Dict:
    """A total-field / scattered-field box (write_json.write_tfsf_m_def; parsed at INPUTS/parallelInputs.cpp:219-408): plane wave along
    the integer direction m, polarisation angle psi (degrees)."""
    return {"loc": list(loc), "size": list(size), "m": list(m), "psi": psi, "PulseList": pulses, "circPol": "Ex", "ellpiticalKRat": 1.0,
            "pmlThick": pml_thick, "pmlAMax": pml_a_max, "pmlM": pml_m, "pmlMa": pml_ma}


def config(cell: Dict, pml_: Dict, sources: List[Dict], objects: List[Dict], detectors: List[Dict], fluxes: Optional[List[Dict]] = None) -> Dict:
    return {"CompCell": cell, "PML": pml_, "SourceList": sources, "TFSF": [], "ObjectList": objects, "DetectorList": detectors,
            "FluxList": fluxes or []}


def default_dt(res: int, courant: float = 0.5) -> float:
    """dt = courant / sqrt(sum 1/d^2) with d = 1/res in all three directions, 2-D included (parallelInputs.cpp:26-27)."""
    d = 1.0 / res
    return courant / math.sqrt(3.0 / (d * d))


def write(cfg: Dict, path: str) -> None:
    with open(path, "w") as f:
        json.dump(cfg, f, indent=1)


# ------------------------------------------------------------------------------------------------
# BASELINE.md configurations.  `n` is the number of cells per side (the grid has n+1 points); every
# geometric length is given in cells and converted with res so that scaled-down copies keep the mix.
# ------------------------------------------------------------------------------------------------
def c1_te_vacuum(n: int = 511, steps: int = 4000, res: int = 100, pml_cells: int = 20, out: str = "output_data/c1") -> Dict:
    """C1: 2-D TE vacuum, 1-cell Hz Gaussian dipole at the centre, CPML, 1-cell Hz txt detector 64 cells off-centre."""
    dt = default_dt(res)
    off = min(64, n // 4) / res
    return config(comp_cell([n / res, n / res, 0.0], res, steps * dt - 0.5 * dt, "Hz"),
                  pml([pml_cells / res, pml_cells / res, 0.0]),
                  [normal_source("Hz", [0.0, 0.0, 0.0], [0.0, 0.0, 0.0], [gaussian_pulse(1.5, 1.0)])],
                  [],
                  [detector([off, 0.0, 0.0], [0.0, 0.0, 0.0], "Hz", out + "/dtc", time_int=dt * 1.0000001)])


def c2_tm_drude(n: int = 2047, steps: int = 8000, res: int = 100, pml_cells: int = 20, rod: Sequence[int] = (400, 60), material: str = "drude",
                nfreq: int = 64, out: str = "output_data/c2") -> Dict:
    """C2: 2-D TM, Drude (or built-in Au) nanorod, Ez line source, 4-edge flux box."""
    dt = default_dt(res)
    rx, ry = rod[0] / res, rod[1] / res
    if material == "drude":
        obj = block([rx, ry, 0.0], [0.0, 0.0, 0.0], eps=1.0, pols=[drude_pole(9.03, 0.053)])
    else:
        obj = block([rx, ry, 0.0], [0.0, 0.0, 0.0], material=material)
    half = n / res / 2.0
    src_y = -half + (pml_cells + 10) / res
    span = (n - 2 * pml_cells - 20) / res
    bx, by = rx / 2 + 20 / res, ry / 2 + 20 / res
    fl = [flux(out + "/flux_left", [-bx, 0.0, 0.0], [0.0, 2 * by, 0.0], 1.5, 1.0, nfreq, -1.0),
          flux(out + "/flux_right", [bx, 0.0, 0.0], [0.0, 2 * by, 0.0], 1.5, 1.0, nfreq, 1.0),
          flux(out + "/flux_bot", [0.0, -by, 0.0], [2 * bx, 0.0, 0.0], 1.5, 1.0, nfreq, -1.0),
          flux(out + "/flux_top", [0.0, by, 0.0], [2 * bx, 0.0, 0.0], 1.5, 1.0, nfreq, 1.0)] if nfreq > 0 else []
    return config(comp_cell([n / res, n / res, 0.0], res, steps * dt - 0.5 * dt, "Ez"),
                  pml([pml_cells / res, pml_cells / res, 0.0]),
                  [normal_source("Ez", [0.0, src_y, 0.0], [span, 0.0, 0.0], [gaussian_pulse(1.5, 1.0)])],
                  [obj],
                  [detector([0.0, by + 5 / res, 0.0], [0.0, 0.0, 0.0], "Ez", out + "/dtc", time_int=dt * 1.0000001)],
                  fl)


def c3_aniso_slab(n: int = 511, steps: int = 1000, res: int = 100, pml_cells: int = 20, slab_cells: int = 40, out: str = "output_data/c3",
                  nz: Optional[int] = None, ny: Optional[int] = None) -> Dict:
    """C3: 3-D anisotropic (oriented-dipole Lorentz) dielectric slab through the PML, Ex dipole in the slab, 3 point detectors."""
    dt = default_dt(res)
    ny = n if ny is None else ny
    nz = n if nz is None else nz
    s2 = 1.0 / math.sqrt(2.0)
    pole = lorentz_pole(sigma_p=1.5, gamma=0.05, omega=2.5, dip_or_e="unidirectional", dir_dip_e=[s2, s2, 0.0])
    slab = block([(n + 2) / res, (ny + 2) / res, slab_cells / res], [0.0, 0.0, 0.0], eps=2.25, pols=[pole])
    q = min(40, n // 6) / res
    dets = [detector([q, 0.0, 0.0], [0.0, 0.0, 0.0], "Ex", out + "/dtc_a", time_int=dt * 1.0000001),
            detector([0.0, q, 0.0], [0.0, 0.0, 0.0], "Ey", out + "/dtc_b", time_int=dt * 1.0000001),
            detector([q, q, 0.0], [0.0, 0.0, 0.0], "Hz", out + "/dtc_c", time_int=dt * 1.0000001)]
    return config(comp_cell([n / res, ny / res, nz / res], res, steps * dt - 0.5 * dt, "Ex"),
                  pml([pml_cells / res] * 3),
                  [normal_source("Ex", [0.0, 0.0, 0.0], [0.0, 0.0, 0.0], [gaussian_pulse(1.5, 1.0)])],
                  [slab], dets)


def c4_plasmonic_ml(n: int = 767, steps: int = 500, res: int = 100, pml_cells: int = 20, cube: int = 40, pitch: int = 70, narray: int = 10,
                    sheet: int = 1000, metal: str = "Au", out: str = "output_data/c4", ny: Optional[int] = None, nz: Optional[int] = None,
                    sheet_gap: int = 10, src_margin: int = 10) -> Dict:
    """C4: narray x narray metal cubes under a one-node-thick two-level emitter sheet, Ez plane source."""
    dt = default_dt(res)
    ny = n if ny is None else ny
    nz = n if nz is None else nz
    objs = []
    x0 = -(narray - 1) * pitch / 2.0
    for iy in range(narray):
        for ix in range(narray):
            objs.append(block([cube / res] * 3, [(x0 + ix * pitch) / res, (x0 + iy * pitch) / res, 0.0], material=metal))
    # one node thick: exactly on a grid node (nodes sit at (k - cells/2) * d, half-integers for odd cell counts)
    zs = (math.floor(nz / 2.0 + cube / 2 + sheet_gap) - nz / 2.0) / res
    objs.append(ml_block([(sheet - 1) / res, (sheet - 1) / res, 0.0], [0.0, 0.0, zs], mol_den=1e25, e_levels_ev=[0.0, 2.0], dipole_debye=10.0,
                         relax_rate=1e12, dephasing_rate=1e13, dtc_levs=[3], pop_fname_base=out + "/qe_"))
    half_z = nz / res / 2.0
    mx_, my_ = min(10, (n - 2 * pml_cells) // 4), min(10, (ny - 2 * pml_cells) // 4)      # source margin to the CPML, in cells
    span_x, span_y = (n - 2 * pml_cells - 2 * mx_) / res, (ny - 2 * pml_cells - 2 * my_) / res
    src_z = half_z - (pml_cells + src_margin) / res        # keep it outside the emitter object's 3-cell buffer: a source inside a D-cell is overwritten
    return config(comp_cell([n / res, ny / res, nz / res], res, steps * dt - 0.5 * dt, "Ez"),
                  pml([pml_cells / res] * 3),
                  [normal_source("Ex", [0.0, 0.0, src_z], [span_x, span_y, 0.0], [gaussian_pulse(1.5 * 2.0 / 1.86, 1.0)])],
                  objs,
                  [detector([0.0, 0.0, zs], [0.0, 0.0, 0.0], "Ex", out + "/dtc", time_int=dt * 1.0000001)])


def c5_aniso_ml(nx: int = 2047, ny: int = 255, nz: int = 1023, steps: int = 200, res: int = 100, pml_cells: int = 20, slab_cells: int = 40,
                sheet: bool = True, sheet_margin: int = 20, out: str = "output_data/c5") -> Dict:
    """C5: the C3 anisotropic (oriented-dipole Lorentz) slab through the PML plus the C4 two-level emitter sheet (one node
    thick, 10 cells above the slab, inside the non-PML interior), Ex dipole in the slab, point detectors.  nx, ny, nz are
    cells per side of the WHOLE grid (points = cells + 1); the y-slab decomposition cuts ny."""
    dt = default_dt(res)
    s2 = 1.0 / math.sqrt(2.0)
    pole = lorentz_pole(sigma_p=1.5, gamma=0.05, omega=2.5, dip_or_e="unidirectional", dir_dip_e=[s2, s2, 0.0])
    objs = [block([(nx + 2) / res, (ny + 2) / res, slab_cells / res], [0.0, 0.0, 0.0], eps=2.25, pols=[pole])]
    # the sheet is one node thick: put it exactly on a grid node (nodes sit at (k - cells/2) * d, half-integers for odd cell counts)
    zs = (math.floor(nz / 2.0 + slab_cells / 2 + 10) - nz / 2.0) / res
    if sheet:
        mx, my = nx - 2 * pml_cells - 2 * sheet_margin, ny - 2 * pml_cells - 2 * sheet_margin
        if mx < 1 or my < 1:
            raise ValueError("c5_aniso_ml: grid too small for the emitter sheet")
        objs.append(ml_block([mx / res, my / res, 0.0], [0.0, 0.0, zs], mol_den=1e25, e_levels_ev=[0.0, 2.0], dipole_debye=10.0,
                             relax_rate=1e12, dephasing_rate=1e13, dtc_levs=[3], pop_fname_base=out + "/qe_"))
    q = min(40, nx // 6) / res
    qy = min(40, ny // 6) / res
    dets = [detector([q, 0.0, 0.0], [0.0, 0.0, 0.0], "Ex", out + "/dtc_a", time_int=dt * 1.0000001),
            detector([0.0, qy, zs], [0.0, 0.0, 0.0], "Ey", out + "/dtc_b", time_int=dt * 1.0000001),
            detector([q, qy, 0.0], [0.0, 0.0, 0.0], "Hz", out + "/dtc_c", time_int=dt * 1.0000001)]
    return config(comp_cell([nx / res, ny / res, nz / res], res, steps * dt - 0.5 * dt, "Ex"),
                  pml([pml_cells / res] * 3),
                  [normal_source("Ex", [0.0, 0.0, 0.0], [0.0, 0.0, 0.0], [gaussian_pulse(1.5, 1.0)])],
                  objs, dets)
