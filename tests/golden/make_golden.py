"""Regenerates the committed golden fixtures from the UNMODIFIED reference (oracle/_ref/chiml_ref,
built in place from /root/reference by oracle/Makefile).  Run here (the container that has
/root/reference); the fixtures travel to the GPU box.

For every case:   <case>.json          the chiML input (chiml_b200.inputs builders)
                  <case>.rank0.plan    the reference constructor's own lists (include/chiml_plan.h)
                  <case>.expect.npz    every public field / pole grid of the reference after n_steps
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from chiml_b200 import inputs as I  # noqa: E402
from chiml_b200 import plan as P    # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "chiml_ref")
RES = 100
DT = I.default_dt(RES)


def _short_pulse(cfg):
    for s in cfg["SourceList"]:
        for p in s["PulseList"]:
            p["t_0"] = 0.25
            p["cutoff"] = 2.5
    return cfg


def cases():
    c = {}
    c["te_vacuum"] = I.c1_te_vacuum(n=47, steps=150, pml_cells=8, out="out/te")
    c["tm_drude"] = I.c2_tm_drude(n=63, steps=100, pml_cells=8, rod=(20, 6), nfreq=0, out="out/tm")
    c["tm_au"] = I.c2_tm_drude(n=63, steps=100, pml_cells=8, rod=(20, 6), material="Au", nfreq=0, out="out/tmau")
    c["vac3d"] = I.config(I.comp_cell([21 / RES, 17 / RES, 23 / RES], RES, 60 * DT - 0.5 * DT, "Ex"), I.pml([5 / RES] * 3),
                          [I.normal_source("Ez", [0, 0, 0], [0, 0, 0], [I.gaussian_pulse(1.5, 1.0)])], [],
                          [I.detector([0.03, 0, 0], [0, 0, 0], "Ez", "out/v3/dtc", time_int=DT * 1.0000001)])
    c["lorentz3d"] = I.config(I.comp_cell([23 / RES, 19 / RES, 21 / RES], RES, 60 * DT - 0.5 * DT, "Ex"), I.pml([5 / RES] * 3),
                              [I.normal_source("Ez", [0, 0, 0.04], [0.05, 0.04, 0], [I.gaussian_pulse(1.5, 1.0)])],
                              [I.block([0.08, 0.06, 0.05], [0.01, 0, -0.02], eps=2.0, pols=[I.lorentz_pole(1.2, 0.1, 2.0), I.lorentz_pole(0.5, 0.05, 3.0)]),
                               I.sphere(0.04, [-0.03, 0.02, 0.03], eps=1.5, pols=[I.lorentz_pole(0.7, 0.2, 1.0)])],
                              [I.detector([0.03, 0, 0], [0, 0, 0], "Ez", "out/l3/dtc", time_int=DT * 1.0000001)])
    c["aniso_slab3d"] = I.c3_aniso_slab(n=23, steps=50, pml_cells=5, slab_cells=6, out="out/c3")
    c["kappa3d"] = I.config(I.comp_cell([19 / RES, 21 / RES, 17 / RES], RES, 50 * DT - 0.5 * DT, "Ex"),
                            I.pml([5 / RES, 6 / RES, 4 / RES], a_max=0.2, ma=2.0, m=3.5, sig_opt_rat=0.9, kappa_max=3.0),
                            [I.normal_source("Ey", [0, 0, 0], [0, 0, 0], [I.gaussian_pulse(1.5, 1.0)])],
                            [I.block([0.06, 0.3, 0.05], [0.0, 0.0, 0.0], eps=3.0)],
                            [I.detector([0.03, 0, 0], [0, 0, 0], "Ey", "out/k3/dtc", time_int=DT * 1.0000001)])
    c = {k: _short_pulse(v) for k, v in c.items()}
    # ---- Maxwell-Liouville emitter cases (sources in vacuum: a source inside a D-cell is overwritten by D->E) ----
    relax1 = [{"state_i": 1, "state_f": 0, "rate": 1e12, "dephasing_rate": 1e13}]
    two = lambda sz, loc, lv: I.ml_object(sz, loc, 1e25, [(0, 0), (1, 0)], [{"E_cen": [0.0]}, {"E_cen": [2.0]}], [0, 10.0, 10.0, 0], relax1,  # noqa: E731
                                          eps=1.5, dtc_levs=lv)
    pulse = lambda f, a: I.gaussian_pulse(f, 1.0, intensity=a, t_0=0.25, cutoff=2.5)  # noqa: E731
    c["ml3d_two"] = I.config(I.comp_cell([23 / RES, 19 / RES, 21 / RES], RES, 60 * DT - 0.5 * DT, "Ex"), I.pml([5 / RES] * 3),
                             [I.normal_source("Ez", [-0.05, -0.04, -0.04], [0, 0, 0], [pulse(1.5, 3e13)])],
                             [I.block([0.08, 0.06, 0.05], [0.05, 0, -0.03], eps=2.0), two([0.05, 0.04, 0.02], [-0.005, 0.005, 0.015], [3, 1])],
                             [I.detector([0.03, 0, 0], [0, 0, 0], "Ez", "out/m2/dtc", time_int=DT * 1.0000001)])
    four = I.ml_object([0.05, 0.04, 0.02], [-0.005, 0.005, 0.015], 1e25, [(0, 0), (1, -1), (1, 0), (1, 1)],
                       [{"E_cen": [0.0]}, {"E_cen": [1.9, 2.1], "weights": [0.6, 0.4], "levs_described": 3}],
                       [0, 10, 8, 6, 10, 0, 0, 0, 8, 0, 0, 0, 6, 0, 0, 0],
                       [{"state_i": 1, "state_f": 0, "rate": 1e12, "dephasing_rate": 1e13}, {"state_i": 2, "state_f": 0, "rate": 2e12, "dephasing_rate": 0.5e13},
                        {"state_i": 3, "state_f": 0, "rate": 1.5e12}], eps=1.2, dtc_levs=[5, 0], pop_every=2)
    c["ml3d_four"] = I.config(I.comp_cell([23 / RES, 19 / RES, 21 / RES], RES, 60 * DT - 0.5 * DT, "Ex"), I.pml([5 / RES] * 3),
                              [I.normal_source("Ez", [-0.05, -0.04, -0.04], [0, 0, 0], [pulse(1.5, 3e13)]),
                               I.normal_source("Ex", [0.04, -0.04, 0.04], [0, 0, 0], [pulse(1.2, 2e13)])],
                              [I.block([0.08, 0.06, 0.05], [0.05, 0, -0.03], eps=2.0), four],
                              [I.detector([0.03, 0, 0], [0, 0, 0], "Ez", "out/m4/dtc", time_int=DT * 1.0000001)])
    c["ml_tm"] = I.config(I.comp_cell([47 / RES, 39 / RES, 0], RES, 80 * DT - 0.5 * DT, "Ez"), I.pml([8 / RES, 8 / RES, 0]),
                          [I.normal_source("Ez", [-0.1, -0.08, 0], [0, 0, 0], [pulse(1.5, 3e13)])],
                          [two([0.08, 0.06, 0.0], [0.02, 0.01, 0.0], [3])],
                          [I.detector([0.03, 0, 0], [0, 0, 0], "Ez", "out/mtm/dtc", time_int=DT * 1.0000001)])
    pxy = I.ml_object([0.08, 0.06, 0.0], [0.02, 0.01, 0.0], 1e25, [(0, 0), (1, -1), (1, 1)], [{"E_cen": [0.0]}, {"E_cen": [2.0], "levs_described": 2}],
                      [0, 10, 10, 10, 0, 0, 10, 0, 0], [{"state_i": 1, "state_f": 0, "rate": 1e12, "dephasing_rate": 1e13},
                                                         {"state_i": 2, "state_f": 0, "rate": 1e12}], eps=1.5, dtc_levs=[4])
    c["ml_te"] = I.config(I.comp_cell([47 / RES, 39 / RES, 0], RES, 80 * DT - 0.5 * DT, "Hz"), I.pml([8 / RES, 8 / RES, 0]),
                          [I.normal_source("Ex", [-0.1, -0.08, 0], [0, 0, 0], [pulse(1.5, 3e13)])],
                          [pxy],
                          [I.detector([0.03, 0, 0], [0, 0, 0], "Ex", "out/mte/dtc", time_int=DT * 1.0000001)])
    # (two emitter objects in one input cannot be run by the reference: every parallelQE is handed the energy levels of ALL
    # emitter objects, parallelFDTDField.cpp:424, and its constructor then reads h0 out of bounds, ML/parallelQE.hpp:221-229 --
    # tests/test_gpu_slabs.py covers several emitter sets per slab against the oracle instead)
    # ---- oriented-dipole objects of finite y extent with different pole counts: with several slabs some slabs hold no node cell,
    # some hold the one-pole film, some the two-pole block; every slab still exchanges the whole grid's two node P_y rows ----
    s2_ = 1.0 / np.sqrt(2.0)
    filmA = I.block([0.3, 0.07, 0.06], [0.0, -0.045, 0.0], eps=2.25,
                    pols=[I.lorentz_pole(1.5, 0.05, 2.5, dip_or_e="unidirectional", dir_dip_e=[s2_, s2_, 0.0])])
    blockB = I.block([0.10, 0.06, 0.08], [0.01, 0.045, 0.0], eps=1.8,
                     pols=[I.lorentz_pole(0.9, 0.1, 2.0, dip_or_e="unidirectional", dir_dip_e=[0.0, 0.6, 0.8]),
                           I.lorentz_pole(0.4, 0.02, 3.1, dip_or_e="unidirectional", dir_dip_e=[0.0, 0.6, 0.8])])
    c["aniso_mixed3d"] = _short_pulse(I.config(
        I.comp_cell([23 / RES, 23 / RES, 21 / RES], RES, 50 * DT - 0.5 * DT, "Ex"), I.pml([5 / RES] * 3),
        [I.normal_source("Ey", [0.0, 0.0, 0.0], [0, 0, 0], [I.gaussian_pulse(1.5, 1.0)])],
        [filmA, blockB],
        [I.detector([0.03, 0, 0], [0, 0, 0], "Ey", "out/am/dtc", time_int=DT * 1.0000001)]))
    # ---- C4 in miniature: built-in 6-pole Au cubes under a two-level emitter sheet, Ex plane source, CPML on every face ----
    c4 = I.c4_plasmonic_ml(n=27, ny=25, nz=37, steps=60, pml_cells=5, cube=6, pitch=10, narray=2, sheet=12, out="out/c4s", sheet_gap=3, src_margin=2)
    for s_ in c4["SourceList"]:
        for p_ in s_["PulseList"]:
            p_["Field_Intensity"] = 3e13
    c["c4_small"] = _short_pulse(c4)
    # ---- running DFT on the four edges of a flux box around a Drude rod (DTC/parallelFlux.hpp, parallelStorageFreqDTC.cpp:21-30) ----
    c["tm_flux"] = _short_pulse(I.c2_tm_drude(n=63, steps=100, pml_cells=8, rod=(20, 6), nfreq=5, out="out/tmflux"))
    # ---- every surface kind of a flux region: 3-D box (six faces), single X / Y / Z planes, unequal sampling intervals ----
    c["flux3d"] = _short_pulse(I.config(
        I.comp_cell([21 / RES, 17 / RES, 23 / RES], RES, 50 * DT - 0.5 * DT, "Ex"), I.pml([5 / RES] * 3),
        [I.normal_source("Ez", [-0.03, 0, 0], [0, 0, 0], [I.gaussian_pulse(1.5, 1.0)])],
        [I.block([0.04, 0.04, 0.04], [0.02, 0.0, 0.0], eps=2.0, pols=[I.lorentz_pole(1.2, 0.1, 2.0)])],
        [I.detector([0.03, 0, 0], [0, 0, 0], "Ez", "out/f3/dtc", time_int=DT * 1.0000001)],
        [I.flux("out/f3/box", [0.02, 0.0, 0.0], [0.06, 0.06, 0.08], 1.5, 1.0, 3),
         dict(I.flux("out/f3/px", [0.05, 0.0, 0.01], [0.0, 0.06, 0.08], 1.5, 1.0, 4), Time_Interval=2.0 * DT),
         I.flux("out/f3/py", [0.0, 0.04, 0.0], [0.1, 0.0, 0.06], 1.2, 0.6, 3),
         dict(I.flux("out/f3/pz", [0.01, 0.0, -0.05], [0.08, 0.06, 0.0], 1.5, 1.0, 5), Time_Interval=3.0 * DT)]))
    te = I.c1_te_vacuum(n=47, steps=80, pml_cells=8, out="out/tef")
    te["FluxList"] = [I.flux("out/tef/box", [0.0, 0.0, 0.0], [0.2, 0.14, 0.0], 1.5, 1.0, 4),
                      I.flux("out/tef/lx", [0.1, 0.0, 0.0], [0.0, 0.2, 0.0], 1.5, 1.0, 3),
                      dict(I.flux("out/tef/ly", [0.0, -0.08, 0.0], [0.22, 0.0, 0.0], 1.5, 1.0, 3), Time_Interval=2.0 * DT)]
    c["te_flux"] = _short_pulse(te)
    # ---- periodic boundaries, real fields (CompCell.PBC, k-point 0: applyBC1Proc, UTIL/FDTD_up_eq.cpp:1058-1116) ----
    # 3-D: a Lorentz film spanning the periodic x / y faces with CPML in z only (the usual array set-up), an eps block against the +x
    # face, a plane source; 120 steps let the wave go round the cell.  2-D: TM Drude rod and TE vacuum with CPML in y only.
    c["pbc3d"] = _short_pulse(I.config(
        I.comp_cell([21 / RES, 17 / RES, 25 / RES], RES, 120 * DT - 0.5 * DT, "Ex", pbc=True), I.pml([0.0, 0.0, 6 / RES]),
        [I.normal_source("Ex", [0.0, 0.0, 0.07], [0.21, 0.17, 0.0], [I.gaussian_pulse(1.5, 1.0)]),
         I.normal_source("Ez", [0.09, -0.07, -0.02], [0, 0, 0], [I.gaussian_pulse(1.2, 1.0)])],
        [I.block([0.3, 0.3, 0.04], [0.0, 0.0, -0.02], eps=2.0, pols=[I.lorentz_pole(1.2, 0.1, 2.0), I.lorentz_pole(0.5, 0.05, 3.0)]),
         I.block([0.06, 0.05, 0.05], [0.09, 0.02, 0.04], eps=3.0)],
        [I.detector([0.03, 0, 0], [0, 0, 0], "Ex", "out/p3/dtc", time_int=DT * 1.0000001)]))
    c["pbc3d_all"] = _short_pulse(I.config(
        I.comp_cell([15 / RES, 19 / RES, 13 / RES], RES, 90 * DT - 0.5 * DT, "Ex", pbc=True), I.pml([0.0, 0.0, 0.0]),
        [I.normal_source("Ey", [0.06, 0.08, -0.05], [0, 0, 0], [I.gaussian_pulse(1.5, 1.0)])],
        [I.sphere(0.05, [-0.06, -0.08, 0.05], eps=2.5, pols=[I.lorentz_pole(0.7, 0.2, 1.0)])],
        [I.detector([0.03, 0, 0], [0, 0, 0], "Ey", "out/p3a/dtc", time_int=DT * 1.0000001)]))
    tm = I.c2_tm_drude(n=63, steps=160, pml_cells=8, rod=(20, 6), nfreq=0, out="out/ptm")
    tm["CompCell"]["PBC"] = True
    tm["PML"]["thickness"] = [0.0, 8 / RES, 0.0]
    c["pbc_tm"] = _short_pulse(tm)
    te = I.c1_te_vacuum(n=47, steps=160, pml_cells=8, out="out/pte")
    te["CompCell"]["PBC"] = True
    te["PML"]["thickness"] = [8 / RES, 0.0, 0.0]
    te["ObjectList"] = [I.block([0.1, 1.0, 0.0], [0.08, 0.0, 0.0], eps=2.2, pols=[I.lorentz_pole(0.8, 0.1, 2.0)])]
    c["pbc_te"] = _short_pulse(te)
    # ---- C4 run periodic, in miniature: a two-level emitter sheet across the whole periodic cell (its polarisation box reaches into the x / y ghost
    # layers) above a Lorentz film, CPML in z only ----
    c["pbc_ml3d"] = I.config(I.comp_cell([21 / RES, 17 / RES, 25 / RES], RES, 70 * DT - 0.5 * DT, "Ex", pbc=True), I.pml([0.0, 0.0, 6 / RES]),
                             [I.normal_source("Ex", [0.0, 0.0, 0.07], [0.21, 0.17, 0.0], [pulse(1.5, 3e13)]),
                              I.normal_source("Ez", [-0.05, -0.04, -0.04], [0, 0, 0], [pulse(1.2, 2e13)])],
                             [I.block([0.3, 0.3, 0.04], [0.0, 0.0, -0.03], eps=2.0, pols=[I.lorentz_pole(1.2, 0.1, 2.0)]),
                              two([0.5, 0.5, 0.02], [0.0, 0.0, 0.025], [3, 1])],
                             [I.detector([0.03, 0, 0], [0, 0, 0], "Ez", "out/pml/dtc", time_int=DT * 1.0000001)])
    # ---- TFSF plane-wave sources (SOURCE/parallelTFSF.hpp): the surface corrections run on the device, the 1-D incident line is the
    # reference's own (its per-step values are recorded into the plan, record TFSFLINE).  2-D TM along +y over a Drude rod; 2-D TE
    # along (1, 2) (incident strides 1 and 2) over a Lorentz block; 3-D along +z over a Lorentz sphere with a flux box around it
    # (incident-field normalisation of getFlux); 3-D along +z through a Lorentz slab that cuts the box (D targets and
    # addIncdFieldsEPChange); 3-D along (1, 0, 1).  (Directions with a negative component make the reference's own incident line
    # overflow to inf / NaN within a few steps here, and m = (1, 0, 0) on a 2-D TM grid leaves its y faces without Hx corrections:
    # such inputs are no fixtures.) ----
    tp = lambda f: [I.gaussian_pulse(f, 1.0, t_0=0.25, cutoff=2.5)]  # noqa: E731
    tm = I.c2_tm_drude(n=63, steps=120, pml_cells=8, rod=(20, 6), nfreq=0, out="out/ttm")
    tm["SourceList"] = []
    tm["TFSF"] = [I.tfsf([0.3, 0.3, 0.0], [0.0, 0.0, 0.0], tp(1.5), m=(0, 1, 0))]
    c["tfsf_tm"] = tm
    te = I.c1_te_vacuum(n=55, steps=120, pml_cells=8, out="out/tte")
    te["SourceList"] = []
    te["ObjectList"] = [I.block([0.08, 0.06, 0.0], [0.02, 0.01, 0.0], eps=2.2, pols=[I.lorentz_pole(0.8, 0.1, 2.0)])]
    te["TFSF"] = [I.tfsf([0.26, 0.22, 0.0], [0.0, 0.0, 0.0], tp(1.5), m=(1, 2, 0), psi=0.0)]
    c["tfsf_te"] = te
    cell3 = lambda steps: I.comp_cell([25 / RES, 23 / RES, 27 / RES], RES, steps * DT - 0.5 * DT, "Ex")  # noqa: E731
    ball = I.sphere(0.03, [0.0, 0.01, 0.0], eps=2.0, pols=[I.lorentz_pole(1.2, 0.1, 2.0)])
    c["tfsf3d"] = I.config(cell3(70), I.pml([5 / RES] * 3), [], [ball],
                           [I.detector([0.03, 0, 0], [0, 0, 0], "Ex", "out/t3/dtc", time_int=DT * 1.0000001)],
                           [I.flux("out/t3/box", [0.0, 0.0, 0.0], [0.06, 0.06, 0.06], 1.5, 1.0, 3)])
    c["tfsf3d"]["TFSF"] = [I.tfsf([0.1, 0.1, 0.12], [0.0, 0.0, 0.0], tp(1.5), m=(0, 0, 1), psi=90.0)]
    c["tfsf3d_slab"] = I.config(cell3(70), I.pml([5 / RES] * 3), [I.normal_source("Ez", [0.02, -0.03, -0.03], [0, 0, 0], tp(1.2))],
                                [I.block([1.0, 1.0, 0.04], [0.0, 0.0, 0.02], eps=2.2, pols=[I.lorentz_pole(0.8, 0.1, 2.0)])],
                                [I.detector([0.03, 0, 0], [0, 0, 0], "Ey", "out/t3s/dtc", time_int=DT * 1.0000001)])
    c["tfsf3d_slab"]["TFSF"] = [I.tfsf([0.1, 0.1, 0.12], [0.0, 0.0, 0.0], tp(1.5), m=(0, 0, 1), psi=60.0)]
    c["tfsf3d_obl"] = I.config(cell3(60), I.pml([5 / RES] * 3), [], [ball],
                               [I.detector([0.03, 0, 0], [0, 0, 0], "Ex", "out/t3o/dtc", time_int=DT * 1.0000001)])
    c["tfsf3d_obl"]["TFSF"] = [I.tfsf([0.1, 0.1, 0.12], [0.0, 0.0, 0.0], tp(1.5), m=(1, 0, 1), psi=90.0)]
    # ---- frequency detectors (DTC/parallelDTC_FREQ.hpp): a field type over a box, an SI-scaled E-power type (three stored fields) every
    # second step, a map output over a plane; beside a flux box, so that the twiddle groups of both kinds are in one run ----
    c["freq3d"] = _short_pulse(I.config(
        I.comp_cell([23 / RES, 19 / RES, 21 / RES], RES, 60 * DT - 0.5 * DT, "Ex"), I.pml([5 / RES] * 3),
        [I.normal_source("Ez", [0, 0, 0.04], [0.05, 0.04, 0], [I.gaussian_pulse(1.5, 1.0)])],
        [I.block([0.08, 0.06, 0.05], [0.01, 0, -0.02], eps=2.0, pols=[I.lorentz_pole(1.2, 0.1, 2.0)])],
        [I.detector([0.03, 0, 0], [0, 0, 0], "Ez", "out/fq/dtc", time_int=DT * 1.0000001),
         I.freq_detector([0.03, 0, 0], [0.02, 0.01, 0.03], "Ez", "out/fq/ez", 1.5, 1.0, 4, time_int=DT * 1.0000001),
         I.freq_detector([0.0, 0.02, 0.0], [0.02, 0.0, 0.01], "E_pow", "out/fq/epow", 1.5, 1.0, 4, time_int=2 * DT * 1.0000001, si=True),
         I.freq_detector([0.0, 0.0, 0.02], [0.02, 0.02, 0.0], "Hy", "out/fq/map", 1.5, 1.0, 2, time_int=DT * 1.0000001, output_map=True)],
        [I.flux("out/fq/box", [0.0, 0.0, 0.0], [0.06, 0.06, 0.06], 1.5, 1.0, 3)]))
    # ---- Bloch-periodic runs: a k-point switches the reference to complex fields (parallelFDTDFieldCplx); the pulse is complex too.  3-D with a
    # Lorentz film and CPML in z, k along x; 3-D periodic on all faces with a k-point in every direction (edges, corners); 2-D TM Drude rod,
    # 2-D TE Lorentz block ----
    def _bloch(cfg, k):
        cfg = _short_pulse(cfg)
        cfg["CompCell"]["PBC"] = True
        cfg["CompCell"]["k-point"] = list(k)
        return cfg
    c["cplx3d"] = _bloch(I.config(
        I.comp_cell([21 / RES, 17 / RES, 25 / RES], RES, 100 * DT - 0.5 * DT, "Ex", pbc=True), I.pml([0.0, 0.0, 6 / RES]),
        [I.normal_source("Ex", [0.0, 0.0, 0.07], [0.21, 0.17, 0.0], [I.gaussian_pulse(1.5, 1.0)]),
         I.normal_source("Ez", [0.09, -0.07, -0.02], [0, 0, 0], [I.gaussian_pulse(1.2, 1.0)])],
        [I.block([0.3, 0.3, 0.04], [0.0, 0.0, -0.02], eps=2.0, pols=[I.lorentz_pole(1.2, 0.1, 2.0), I.lorentz_pole(0.5, 0.05, 3.0)]),
         I.block([0.06, 0.05, 0.05], [0.09, 0.02, 0.04], eps=3.0)],
        [I.detector([0.03, 0, 0], [0, 0, 0], "Ex", "out/c3/dtc", time_int=DT * 1.0000001)]), [0.7, 0.0, 0.0])
    c["cplx3d_all"] = _bloch(I.config(
        I.comp_cell([15 / RES, 19 / RES, 13 / RES], RES, 90 * DT - 0.5 * DT, "Ex", pbc=True), I.pml([0.0, 0.0, 0.0]),
        [I.normal_source("Ey", [0.06, 0.08, -0.05], [0, 0, 0], [I.gaussian_pulse(1.5, 1.0)])],
        [I.sphere(0.05, [-0.06, -0.08, 0.05], eps=2.5, pols=[I.lorentz_pole(0.7, 0.2, 1.0)])],
        [I.detector([0.03, 0, 0], [0, 0, 0], "Ey", "out/c3a/dtc", time_int=DT * 1.0000001)]), [0.4, -0.3, 0.6])
    tm = I.c2_tm_drude(n=63, steps=160, pml_cells=8, rod=(20, 6), nfreq=0, out="out/ctm")
    tm["PML"]["thickness"] = [0.0, 8 / RES, 0.0]
    c["cplx_tm"] = _bloch(tm, [0.5, 0.0, 0.0])
    te = I.c1_te_vacuum(n=47, steps=160, pml_cells=8, out="out/cte")
    te["PML"]["thickness"] = [8 / RES, 0.0, 0.0]
    te["ObjectList"] = [I.block([0.1, 1.0, 0.0], [0.08, 0.0, 0.0], eps=2.2, pols=[I.lorentz_pole(0.8, 0.1, 2.0)])]
    c["cplx_te"] = _bloch(te, [0.0, -0.6, 0.0])
    # ---- magnetic-dispersive media (B / H / M mirror of D / E / P: updateMagH, updateB, B2H): a block with mu = 1.5 and poles that are electric
    # and magnetic beside a purely electric sphere, inside the CPML-free interior; the same block reaching through the CPML (magMatInPML_: the
    # H-side CPML acts on B, every CPML cell becomes a B cell); 2-D TM (Hx, Hy magnetic) and TE (Hz magnetic) ----
    magp = lambda sp, g, w, sm: I.lorentz_pole(sp, g, w, sigma_m=sm)  # noqa: E731
    mblock = dict(I.block([0.08, 0.06, 0.05], [0.01, 0, -0.02], eps=2.0, pols=[magp(1.2, 0.1, 2.0, 0.6), magp(0.5, 0.05, 3.0, 0.3)]), mu=1.5)
    c["mag3d"] = _short_pulse(I.config(
        I.comp_cell([23 / RES, 19 / RES, 21 / RES], RES, 60 * DT - 0.5 * DT, "Ex"), I.pml([5 / RES] * 3),
        [I.normal_source("Ez", [0, 0, 0.04], [0.05, 0.04, 0], [I.gaussian_pulse(1.5, 1.0)]),
         I.normal_source("Hy", [0.02, 0.0, -0.02], [0, 0, 0], [I.gaussian_pulse(1.2, 1.0)])],
        [mblock, I.sphere(0.03, [-0.04, 0.02, 0.04], eps=1.5, pols=[I.lorentz_pole(0.7, 0.2, 1.0)])],
        [I.detector([0.03, 0, 0], [0, 0, 0], "Hy", "out/mg/dtc", time_int=DT * 1.0000001)]))
    c["mag3d_pml"] = _short_pulse(I.config(
        I.comp_cell([21 / RES, 19 / RES, 17 / RES], RES, 50 * DT - 0.5 * DT, "Ex"), I.pml([5 / RES, 4 / RES, 5 / RES]),
        [I.normal_source("Ey", [0.0, 0.0, 0.0], [0, 0, 0], [I.gaussian_pulse(1.5, 1.0)])],
        [dict(I.block([0.5, 0.06, 0.05], [0.0, 0.01, 0.0], eps=1.0, pols=[magp(0.0, 0.1, 2.0, 0.8)]), mu=2.0)],
        [I.detector([0.03, 0, 0], [0, 0, 0], "Hz", "out/mgp/dtc", time_int=DT * 1.0000001)]))
    tm = I.c2_tm_drude(n=63, steps=100, pml_cells=8, rod=(20, 6), nfreq=0, out="out/mtm")
    tm["ObjectList"] = [dict(I.block([0.2, 0.06, 0.0], [0.0, 0.05, 0.0], eps=2.0, pols=[magp(0.9, 0.1, 2.0, 0.7)]), mu=1.3)]
    c["mag_tm"] = _short_pulse(tm)
    te = I.c1_te_vacuum(n=47, steps=100, pml_cells=8, out="out/mte")
    te["ObjectList"] = [dict(I.block([0.1, 0.1, 0.0], [0.1, -0.03, 0.0], eps=1.0, pols=[magp(0.0, 0.05, 1.8, 0.9)]), mu=1.8)]      # (off the Hz source: B2H overwrites a source inside)
    c["mag_te"] = _short_pulse(te)
    # ---- chiral media (the "chi" of chiML): a block whose poles are electric, magnetic and chiral (tau != 0: P gains a term driven by the 8-point
    # average of H and of the previous H, M one driven by E) beside an achiral Lorentz sphere; and the same kind of block reaching through the CPML ----
    chip = lambda sp, g, w, sm, tau: I.lorentz_pole(sp, g, w, sigma_m=sm, tau=tau)  # noqa: E731
    c["chi3d"] = _short_pulse(I.config(
        I.comp_cell([23 / RES, 19 / RES, 21 / RES], RES, 60 * DT - 0.5 * DT, "Ex"), I.pml([5 / RES] * 3),
        [I.normal_source("Ez", [0, 0, 0.04], [0.05, 0.04, 0], [I.gaussian_pulse(1.5, 1.0)]),
         I.normal_source("Hy", [0.06, 0.0, -0.04], [0, 0, 0], [I.gaussian_pulse(1.2, 1.0)])],
        [dict(I.block([0.08, 0.06, 0.05], [0.01, 0, -0.02], eps=2.0, pols=[chip(1.2, 0.1, 2.0, 0.6, 0.3), chip(0.5, 0.05, 3.0, 0.3, -0.2)]), mu=1.5),
         I.sphere(0.03, [-0.04, 0.02, 0.04], eps=1.5, pols=[I.lorentz_pole(0.7, 0.2, 1.0)])],
        [I.detector([0.03, 0, 0], [0, 0, 0], "Hy", "out/ch/dtc", time_int=DT * 1.0000001)]))
    c["chi3d_pml"] = _short_pulse(I.config(
        I.comp_cell([21 / RES, 19 / RES, 17 / RES], RES, 50 * DT - 0.5 * DT, "Ex"), I.pml([5 / RES, 4 / RES, 5 / RES]),
        [I.normal_source("Ey", [0.0, 0.04, 0.0], [0, 0, 0], [I.gaussian_pulse(1.5, 1.0)])],
        [dict(I.block([0.5, 0.06, 0.05], [0.0, -0.02, 0.0], eps=1.8, pols=[chip(0.9, 0.1, 2.0, 0.8, 0.4)]), mu=1.2)],
        [I.detector([0.03, 0, 0], [0, 0, 0], "Hz", "out/chp/dtc", time_int=DT * 1.0000001)]))
    # ---- dipoles oriented relative to the surface normal (REL_TO_NORM: setupDipMoments evaluates the object's surface gradient at every node,
    # parallelFDTDField.hpp:960-1048): a sphere with a "normal" and a "tangent" pole, a block with a pole at 30 / 60 degrees next to an isotropic
    # oriented pole, a bar with tangent-isotropic poles reaching through the CPML (2-D grids: the reference's own constructor asserts on oriented dipoles) ----
    reln = lambda sp, g, w, how, **kw: I.lorentz_pole(sp, g, w, dip_or_e=how, **kw)  # noqa: E731
    c["dipnorm3d"] = _short_pulse(I.config(
        I.comp_cell([23 / RES, 21 / RES, 21 / RES], RES, 50 * DT - 0.5 * DT, "Ex"), I.pml([5 / RES] * 3),
        [I.normal_source("Ey", [0.0, 0.0, 0.0], [0, 0, 0], [I.gaussian_pulse(1.5, 1.0)])],
        [I.sphere(0.045, [-0.035, 0.01, 0.02], eps=1.6, pols=[reln(0.9, 0.1, 2.0, "normal"), reln(0.6, 0.05, 2.6, "tangent")]),
         I.block([0.06, 0.07, 0.05], [0.05, -0.02, -0.02], eps=2.0,
                 pols=[reln(0.8, 0.08, 1.8, "rel_norm", pol_ang_e=30.0, az_ang_e=60.0), I.lorentz_pole(0.3, 0.02, 3.1, dip_or_e="unidirectional", dir_dip_e=[0.6, 0.0, 0.8])])],
        [I.detector([0.03, 0, 0], [0, 0, 0], "Ey", "out/dn/dtc", time_int=DT * 1.0000001)]))
    c["dipnorm3d_pml"] = _short_pulse(I.config(
        I.comp_cell([21 / RES, 19 / RES, 19 / RES], RES, 40 * DT - 0.5 * DT, "Ex"), I.pml([5 / RES, 4 / RES, 5 / RES]),
        [I.normal_source("Ez", [0.0, 0.03, 0.0], [0, 0, 0], [I.gaussian_pulse(1.5, 1.0)])],
        [I.block([0.5, 0.06, 0.07], [0.0, -0.02, 0.0], eps=1.8, pols=[reln(0.9, 0.1, 2.0, "tangent", tan_iso=True, dip_or_m="tangent")])],
        [I.detector([0.03, 0, 0], [0, 0, 0], "Ez", "out/dnp/dtc", time_int=DT * 1.0000001)]))
    return c


def main():
    if not os.path.exists(REF):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref", "-j8"], check=True)
    work = os.path.join(HERE, "_work")
    os.makedirs(work, exist_ok=True)
    only = set(sys.argv[1:])            # python make_golden.py [case ...]: regenerate only the named cases
    for name, cfg in cases().items():
        if only and name not in only:
            continue
        jpath = os.path.join(HERE, name + ".json")
        I.write(cfg, jpath)
        I.write(cfg, os.path.join(work, name + ".json"))   # the reference prefixes "stripped_" to the name as given: run on a cwd-relative copy
        subprocess.run([REF, name + ".json", "--dump", os.path.join(work, name + ".dump"), "--plan", os.path.join(HERE, name),
                        "--quiet", "--no-output"], check=True, cwd=work, stdout=subprocess.DEVNULL)
        dump = P.read_dump(os.path.join(work, name + ".dump"))
        arrays = {nm: arr for (rank, nm), (ln, ys, arr) in dump.items() if rank == 0}
        np.savez_compressed(os.path.join(HERE, name + ".expect.npz"), **arrays)
        print(name, {k: v.shape for k, v in list(arrays.items())[:1]}, len(arrays), "arrays",
              os.path.getsize(os.path.join(HERE, name + ".expect.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
