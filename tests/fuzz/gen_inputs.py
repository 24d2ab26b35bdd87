"""Random chiML inputs for the host-setup fuzz test (tests/test_host_plan_fuzz.py): 2-D TE / TM and 3-D cells with random CPML
parameters, blocks and spheres of random dielectric / Lorentz / Drude / oriented-dipole (3-D only) media that may reach into the
CPML, a random source box, detector box and flux regions (boxes, planes, lines, unequal sampling intervals)."""
import random

from chiml_b200 import inputs as I

RES = 100
DT = I.default_dt(RES)


def rnd_pulse(r):
    """One pulse of a random profile, with time constants of a few time steps so that a short run samples its shape
    (UTIL/PulseFxn.hpp; parameters as parallelInputs.cpp:560-640 reads them)."""
    prof = r.choice(["gaussian", "BH", "rectangle", "continuous", "ramped_cont", "ricker"])
    p = {"profile": prof, "Field_Intensity": r.choice([1.0, 3e13]), "fcen": r.uniform(0.8, 40.0)}
    if prof == "gaussian":
        p.update(fwidth=r.uniform(20.0, 200.0), cutoff=r.uniform(1.5, 4.0), t_0=r.uniform(0.0, 0.02))
    elif prof == "BH":
        p.update(fwidth=r.uniform(20.0, 200.0), tau=r.uniform(0.01, 0.05), t_0=r.uniform(0.0, 0.02))   # fwidth is read even when tau is given
    elif prof == "rectangle":
        p.update(tau=r.uniform(0.005, 0.03), t_0=r.uniform(0.0, 0.02), n=r.choice([10, 30]))
    elif prof == "ramped_cont":
        p.update(ramp_val=r.uniform(10.0, 200.0))
    elif prof == "ricker":
        p.update(fwidth=r.uniform(20.0, 200.0), cutoff=r.uniform(0.005, 0.03))
    return p


def rnd_case(seed, steps=10, pulses="gaussian"):
    r=random.Random(seed)
    mode=r.choice(["3d","te","tm"])
    if mode=="3d":
        n=[r.randint(17,27) for _ in range(3)]; pol="Ex"
    else:
        n=[r.randint(31,55), r.randint(31,55), 0]; pol="Hz" if mode=="te" else "Ez"
    size=[k/RES for k in n]
    pmlc=[r.randint(3,6) for _ in range(3)]
    if mode!="3d": pmlc[2]=0
    pml=I.pml([c/RES for c in pmlc], a_max=r.choice([0.25,0.1]), ma=r.choice([1.0,2.0]), m=r.choice([3.0,3.5]), kappa_max=r.choice([1.0,2.5]))
    objs=[]
    for _ in range(r.randint(0,3)):
        loc=[r.uniform(-0.3,0.3)*size[k] for k in range(3)]
        if mode!="3d": loc[2]=0.0
        pols=[]
        kind=r.choice(["eps","lor","uni","drude"]) if mode=="3d" else r.choice(["eps","lor","drude"])
        if kind=="lor": pols=[I.lorentz_pole(r.uniform(0.3,1.5), r.uniform(0.02,0.2), r.uniform(1,3)) for _ in range(r.randint(1,2))]
        if kind=="uni" :
            v=[r.uniform(-1,1) for _ in range(3)]
            if mode=="tm": v=[0,0,1.0]
            if mode=="te": v[2]=0.0
            nn=sum(x*x for x in v)**0.5 or 1.0
            pols=[I.lorentz_pole(r.uniform(0.3,1.5), r.uniform(0.02,0.2), r.uniform(1,3), dip_or_e="unidirectional", dir_dip_e=[x/nn for x in v])]
        if kind=="drude": pols=[I.drude_pole(r.uniform(5,9), r.uniform(0.05,0.2))]
        eps=r.choice([1.0,2.0,2.25,4.0])
        if r.random()<0.5:
            sz=[r.uniform(0.1,0.6)*size[k] for k in range(3)]
            if mode!="3d": sz[2]=0.0
            objs.append(I.block(sz, loc, eps=eps, pols=pols))
        else:
            objs.append(I.sphere(r.uniform(0.08,0.25)*min(s for s in size if s>0), loc, eps=eps, pols=pols))
    srcpol={"3d":r.choice(["Ex","Ey","Ez","Hx","Hz"]),"te":r.choice(["Hz","Ex","Ey"]),"tm":r.choice(["Ez","Hx","Hy"])}[mode]
    sloc=[r.uniform(-0.2,0.2)*size[k] for k in range(3)]; ssz=[r.choice([0.0, r.uniform(0,0.3)*size[k]]) for k in range(3)]
    if mode!="3d": sloc[2]=0.0; ssz[2]=0.0
    if pulses == "gaussian":
        srcs=[I.normal_source(srcpol, sloc, ssz, [I.gaussian_pulse(1.5,1.0,t_0=0.25,cutoff=2.5)])]
    else:
        srcs=[I.normal_source(srcpol, sloc, ssz, [rnd_pulse(r) for _ in range(r.randint(1, 2))])]
    dets=[]
    dpol={"3d":r.choice(["Ex","Ey","Ez","Hx","Hy","Hz"]),"te":r.choice(["Hz","Ex","Ey"]),"tm":r.choice(["Ez","Hx","Hy"])}[mode]
    dloc=[r.uniform(-0.2,0.2)*size[k] for k in range(3)]; dsz=[r.choice([0.0, r.uniform(0,0.2)*size[k]]) for k in range(3)]
    if mode!="3d": dloc[2]=0.0; dsz[2]=0.0
    dets.append(I.detector(dloc, dsz, dpol, f"out/fz{seed}/d", time_int=DT*r.choice([1.0000001,2.0000001])))
    fluxes=[]
    for k in range(r.randint(0,2)):
        floc=[r.uniform(-0.1,0.1)*size[j] for j in range(3)]
        fsz=[r.uniform(0.2,0.5)*size[j] for j in range(3)]
        if r.random()<0.5: fsz[r.randrange(3 if mode=="3d" else 2)]=0.0
        if mode!="3d": floc[2]=0.0; fsz[2]=0.0
        fl=I.flux(f"out/fz{seed}/f{k}", floc, fsz, 1.5, 1.0, r.randint(3,5))
        if r.random()<0.5: fl["Time_Interval"]=2.0*DT
        fluxes.append(fl)
    return I.config(I.comp_cell(size, RES, steps*DT-0.5*DT, pol), pml, srcs, objs, dets, fluxes)



def rnd_ml_case(seed):
    """A random Maxwell-Liouville emitter block (two, three or four levels; one or two level systems with weights; random energies,
    couplings, relaxation and dephasing rates, density, background eps, population detectors) in a small 3-D, TE or TM cell."""
    r = random.Random(1000 + seed)
    kind = r.choice(["two3d", "four3d", "twotm", "threete"])
    e1 = r.uniform(1.5, 2.5)
    cen = [e1] if r.random() < 0.5 else [e1 - r.uniform(0.05, 0.2), e1 + r.uniform(0.05, 0.2)]
    lev1 = {"E_cen": cen}
    if len(cen) == 2:
        w = r.uniform(0.2, 0.8)
        lev1["weights"] = [w, 1.0 - w]
    c = r.uniform(2.0, 15.0)
    rate = lambda: r.choice([1e12, 2e12, 5e11])      # noqa: E731
    deph = lambda: r.choice([1e13, 5e12])            # noqa: E731
    if kind in ("two3d", "twotm"):
        basis, levels, coup = [(0, 0), (1, 0)], [{"E_cen": [0.0]}, lev1], [0, c, c, 0]
        relax = [{"state_i": 1, "state_f": 0, "rate": rate(), "dephasing_rate": deph()}]
        nlev = 2
    elif kind == "threete":
        lev1["levs_described"] = 2
        basis, levels, coup = [(0, 0), (1, -1), (1, 1)], [{"E_cen": [0.0]}, lev1], [0, c, c, c, 0, 0, c, 0, 0]
        relax = [{"state_i": 1, "state_f": 0, "rate": rate(), "dephasing_rate": deph()}, {"state_i": 2, "state_f": 0, "rate": rate()}]
        nlev = 3
    else:
        lev1["levs_described"] = 3
        c2, c3 = r.uniform(2.0, 15.0), r.uniform(2.0, 15.0)
        basis, levels = [(0, 0), (1, -1), (1, 0), (1, 1)], [{"E_cen": [0.0]}, lev1]
        coup = [0, c, c2, c3, c, 0, 0, 0, c2, 0, 0, 0, c3, 0, 0, 0]
        relax = [{"state_i": 1, "state_f": 0, "rate": rate(), "dephasing_rate": deph()}, {"state_i": 2, "state_f": 0, "rate": rate(), "dephasing_rate": deph()},
                 {"state_i": 3, "state_f": 0, "rate": rate()}]
        nlev = 4
    dtc_levs = sorted(r.sample(range(nlev * nlev), r.randint(1, 2)))
    eps = r.choice([1.0, 1.2, 1.5])
    den = r.choice([1e24, 1e25, 3e25])
    if kind in ("two3d", "four3d"):
        n = [r.randint(21, 25), r.randint(17, 21), r.randint(19, 23)]
        size = [k / RES for k in n]
        pml = I.pml([5 / RES] * 3)
        sz = [r.randint(3, 6) / RES, r.randint(3, 5) / RES, r.randint(2, 4) / RES]
        loc = [r.randint(-2, 2) / RES + 0.005 * (k % 2 == 0) for k in range(3)]
        obj = I.ml_object(sz, loc, den, basis, levels, coup, relax, eps=eps, dtc_levs=dtc_levs, pop_every=r.choice([1, 2]))
        src = I.normal_source("Ez", [-0.05, -0.04, -0.04], [0, 0, 0], [I.gaussian_pulse(1.5, 1.0, intensity=3e13, t_0=0.25, cutoff=2.5)])
        det = I.detector([0.03, 0, 0], [0, 0, 0], "Ez", f"out/ml{seed}/d", time_int=DT * 1.0000001)
        return I.config(I.comp_cell(size, RES, 4 * DT - 0.5 * DT, "Ex"), pml, [src], [obj], [det])
    n = [r.randint(41, 49), r.randint(35, 41), 0]
    size = [k / RES for k in n]
    pml = I.pml([8 / RES, 8 / RES, 0])
    sz = [r.randint(5, 9) / RES, r.randint(4, 7) / RES, 0.0]
    loc = [r.randint(-3, 3) / RES, r.randint(-3, 3) / RES, 0.0]
    obj = I.ml_object(sz, loc, den, basis, levels, coup, relax, eps=eps, dtc_levs=dtc_levs, pop_every=r.choice([1, 2]))
    pol, spol = ("Ez", "Ez") if kind == "twotm" else ("Hz", "Ex")
    src = I.normal_source(spol, [-0.1, -0.08, 0], [0, 0, 0], [I.gaussian_pulse(1.5, 1.0, intensity=3e13, t_0=0.25, cutoff=2.5)])
    det = I.detector([0.03, 0, 0], [0, 0, 0], spol, f"out/ml{seed}/d", time_int=DT * 1.0000001)
    return I.config(I.comp_cell(size, RES, 4 * DT - 0.5 * DT, pol), pml, [src], [obj], [det])


def rnd_case_rotated(seed, steps=10, pulses="gaussian"):
    """rnd_case with its objects turned: blocks get random orientation angles (orPhi, and orTheta in 3-D), spheres become cylinders
    of random radius, length and orientation (OBJECTS/Obj.cpp: RealSpace2ObjectSpace, block / cylinder isObj)."""
    cfg = rnd_case(seed, steps=steps, pulses=pulses)
    r = random.Random(5000 + seed)
    three_d = cfg["CompCell"]["size"][2] != 0
    for o in cfg["ObjectList"]:
        o["orPhi"] = r.uniform(-90.0, 90.0)
        o["orTheta"] = r.uniform(20.0, 160.0) if three_d else 90.0
        if o["shape"] == "sphere":
            rad = o.pop("radius")
            o.update(shape="cylinder", radius=rad * r.uniform(0.6, 1.0), length=rad * r.uniform(1.0, 3.0))
    return cfg


def rnd_mag_case(seed, steps=10):
    """Magnetic-dispersive and chiral media at random places: objects with mu_inf > 1, magnetic poles (sigma_m), chiral poles (tau, 3-D only)
    beside plain dielectric / Lorentz ones, some of them long enough to reach through the CPML (B grids, the H-side CPML on B, upB_ / upLorB_ /
    upChiD_ / upChiB_, copy2PrevFields_)."""
    r=random.Random(7000+seed)
    mode=r.choice(["3d","3d","te","tm"])
    if mode=="3d":
        n=[r.randint(17,25) for _ in range(3)]; pol="Ex"
    else:
        n=[r.randint(31,49), r.randint(31,49), 0]; pol="Hz" if mode=="te" else "Ez"
    size=[k/RES for k in n]
    pmlc=[r.randint(3,5) for _ in range(3)]
    if mode!="3d": pmlc[2]=0
    pml=I.pml([c/RES for c in pmlc], a_max=r.choice([0.25,0.1]), ma=r.choice([1.0,2.0]), m=r.choice([3.0,3.5]), kappa_max=r.choice([1.0,2.5]))
    objs=[]
    for k in range(r.randint(1,3)):
        loc=[r.uniform(-0.25,0.25)*size[j] for j in range(3)]
        if mode!="3d": loc[2]=0.0
        kind=r.choice(["mag","mag","chi","mu","lor"]) if mode=="3d" else r.choice(["mag","mag","mu","lor"])
        if k==0 and kind in ("mu","lor"): kind="mag"
        pols=[]; mu=1.0
        if kind=="mag":
            pols=[I.lorentz_pole(r.choice([0.0, r.uniform(0.3,1.2)]), r.uniform(0.02,0.2), r.uniform(1,3), sigma_m=r.uniform(0.2,1.0)) for _ in range(r.randint(1,2))]
            mu=r.choice([1.0,1.5,2.0])
        if kind=="chi":
            pols=[I.lorentz_pole(r.choice([0.0, r.uniform(0.3,1.2)]), r.uniform(0.02,0.2), r.uniform(1,3), sigma_m=r.choice([0.0, r.uniform(0.2,1.0)]), tau=r.uniform(-0.5,0.5) or 0.1)
                  for _ in range(r.randint(1,2))]
            mu=r.choice([1.0,1.5])
        if kind=="mu": mu=r.choice([1.5,2.0,3.0])
        if kind=="lor": pols=[I.lorentz_pole(r.uniform(0.3,1.5), r.uniform(0.02,0.2), r.uniform(1,3))]
        eps=r.choice([1.0,2.0,2.25])
        if r.random()<0.6:
            sz=[r.uniform(0.15,0.5)*size[j] for j in range(3)]
            if r.random()<0.35: sz[r.randrange(3 if mode=="3d" else 2)]=3.0       # through the CPML on both sides
            if mode!="3d": sz[2]=0.0
            o=I.block(sz, loc, eps=eps, pols=pols)
        else:
            o=I.sphere(r.uniform(0.1,0.25)*min(s for s in size if s>0), loc, eps=eps, pols=pols)
        objs.append(dict(o, mu=mu))
    srcpol={"3d":r.choice(["Ex","Ey","Ez","Hx","Hz"]),"te":r.choice(["Hz","Ex","Ey"]),"tm":r.choice(["Ez","Hx","Hy"])}[mode]
    sloc=[r.uniform(-0.2,0.2)*size[k] for k in range(3)]; ssz=[r.choice([0.0, r.uniform(0,0.3)*size[k]]) for k in range(3)]
    if mode!="3d": sloc[2]=0.0; ssz[2]=0.0
    srcs=[I.normal_source(srcpol, sloc, ssz, [I.gaussian_pulse(1.5,1.0,t_0=0.25,cutoff=2.5)])]
    dpol={"3d":r.choice(["Ex","Ey","Ez","Hx","Hy","Hz"]),"te":r.choice(["Hz","Ex","Ey"]),"tm":r.choice(["Ez","Hx","Hy"])}[mode]
    dloc=[r.uniform(-0.2,0.2)*size[k] for k in range(3)]
    if mode!="3d": dloc[2]=0.0
    dets=[I.detector(dloc, [0.0,0.0,0.0], dpol, f"out/fm{seed}/d", time_int=DT*1.0000001)]
    return I.config(I.comp_cell(size, RES, steps*DT-0.5*DT, pol), pml, srcs, objs, dets, [])


def rnd_pbc_case(seed, steps=14):
    """Periodic boundaries (CompCell.PBC, real fields) with random media: 3-D grids periodic in x / y with CPML in z (or none at all), 2-D grids with
    CPML along one axis; blocks that span the periodic faces (crossing the seam between the last and the first y-slab of a multi-slab run), spheres
    and films with Lorentz / Drude poles."""
    r=random.Random(9000+seed)
    mode=r.choice(["3d","3d","te","tm"])
    if mode=="3d":
        n=[r.randint(15,23), r.randint(17,25), r.randint(15,25)]; pol="Ex"
        pmlc=[0,0,r.choice([0,4,6])]
    else:
        n=[r.randint(31,49), r.randint(33,49), 0]; pol="Hz" if mode=="te" else "Ez"
        pmlc=r.choice([[0,6,0],[6,0,0],[0,0,0]])
    size=[k/RES for k in n]
    pml=I.pml([c/RES for c in pmlc], a_max=r.choice([0.25,0.1]), ma=r.choice([1.0,2.0]), m=r.choice([3.0,3.5]), kappa_max=r.choice([1.0,2.5]))
    objs=[]
    for _ in range(r.randint(1,3)):
        loc=[r.uniform(-0.4,0.4)*size[k] for k in range(3)]
        if mode!="3d": loc[2]=0.0
        kind=r.choice(["eps","lor","drude","lor"])
        pols=[]
        if kind=="lor": pols=[I.lorentz_pole(r.uniform(0.3,1.5), r.uniform(0.02,0.2), r.uniform(1,3)) for _ in range(r.randint(1,2))]
        if kind=="drude": pols=[I.drude_pole(r.uniform(5,9), r.uniform(0.05,0.2))]
        eps=r.choice([1.0,2.0,2.25,4.0])
        if r.random()<0.65:
            sz=[r.uniform(0.1,0.5)*size[k] for k in range(3)]
            for k in range(2):
                if r.random()<0.4 and pmlc[k]==0: sz[k]=3.0*size[k]          # spans the periodic faces of that axis
            if mode!="3d": sz[2]=0.0
            objs.append(I.block(sz, loc, eps=eps, pols=pols))
        else:
            objs.append(I.sphere(r.uniform(0.08,0.25)*min(s for s in size if s>0), loc, eps=eps, pols=pols))
    srcpol={"3d":r.choice(["Ex","Ey","Ez","Hx","Hz"]),"te":r.choice(["Hz","Ex","Ey"]),"tm":r.choice(["Ez","Hx","Hy"])}[mode]
    sloc=[r.uniform(-0.3,0.3)*size[k] for k in range(3)]; ssz=[r.choice([0.0, r.uniform(0,0.3)*size[k]]) for k in range(3)]
    if mode!="3d": sloc[2]=0.0; ssz[2]=0.0
    srcs=[I.normal_source(srcpol, sloc, ssz, [I.gaussian_pulse(1.5,1.0,t_0=0.25,cutoff=2.5)])]
    dpol={"3d":r.choice(["Ex","Ey","Ez","Hx","Hy","Hz"]),"te":r.choice(["Hz","Ex","Ey"]),"tm":r.choice(["Ez","Hx","Hy"])}[mode]
    dloc=[r.uniform(-0.3,0.3)*size[k] for k in range(3)]
    if mode!="3d": dloc[2]=0.0
    dets=[I.detector(dloc, [0.0,0.0,0.0], dpol, f"out/fp{seed}/d", time_int=DT*1.0000001)]
    return I.config(I.comp_cell(size, RES, steps*DT-0.5*DT, pol, pbc=True), pml, srcs, objs, dets, [])


def rnd_dipnorm_case(seed, steps=10):
    """Oriented dipoles relative to the surface normal (REL_TO_NORM) at random places: spheres and (unrotated) blocks with "normal", "tangent" and
    polar / azimuthal-angle poles, optionally tangent-isotropic pairs, beside unidirectional and isotropic-oriented poles."""
    r=random.Random(11000+seed)
    n=[r.randint(17,25) for _ in range(3)]
    size=[k/RES for k in n]
    pmlc=[r.randint(3,5) for _ in range(3)]
    pml=I.pml([c/RES for c in pmlc])
    objs=[]
    for k in range(r.randint(1,3)):
        loc=[r.uniform(-0.25,0.25)*size[j] for j in range(3)]
        pols=[]
        for _ in range(r.randint(1,2)):
            how=r.choice(["normal","tangent","rel_norm","unidirectional","taniso"]) if k==0 or r.random()<0.7 else "plain"
            sp, g, w = r.uniform(0.3,1.5), r.uniform(0.02,0.2), r.uniform(1,3)
            if how=="plain": pols.append(I.lorentz_pole(sp,g,w)); break
            if how=="unidirectional":
                v=[r.uniform(-1,1) for _ in range(3)]; nn=sum(x*x for x in v)**0.5 or 1.0
                pols.append(I.lorentz_pole(sp,g,w,dip_or_e="unidirectional",dir_dip_e=[x/nn for x in v]))
            elif how=="rel_norm":
                pols.append(I.lorentz_pole(sp,g,w,dip_or_e="rel_norm",pol_ang_e=r.choice([20.0,30.0,60.0,75.0]),az_ang_e=r.choice([10.0,45.0,60.0,80.0])))
            elif how=="taniso":
                pols.append(I.lorentz_pole(sp,g,w,dip_or_e="tangent",tan_iso=True,dip_or_m="tangent"))
            else:
                pols.append(I.lorentz_pole(sp,g,w,dip_or_e=how))
        eps=r.choice([1.0,2.0,2.25])
        if r.random()<0.5:
            sz=[r.uniform(0.15,0.5)*size[j] for j in range(3)]
            if r.random()<0.3: sz[r.randrange(3)]=3.0
            objs.append(I.block(sz, loc, eps=eps, pols=pols))
        else:
            objs.append(I.sphere(r.uniform(0.1,0.25)*min(size), loc, eps=eps, pols=pols))
    srcpol=r.choice(["Ex","Ey","Ez"])
    sloc=[r.uniform(-0.2,0.2)*size[k] for k in range(3)]
    srcs=[I.normal_source(srcpol, sloc, [0.0,0.0,0.0], [I.gaussian_pulse(1.5,1.0,t_0=0.25,cutoff=2.5)])]
    dets=[I.detector([r.uniform(-0.2,0.2)*size[k] for k in range(3)], [0.0,0.0,0.0], r.choice(["Ex","Ey","Ez"]), f"out/fd{seed}/d", time_int=DT*1.0000001)]
    return I.config(I.comp_cell(size, RES, steps*DT-0.5*DT, "Ex"), pml, srcs, objs, dets, [])
