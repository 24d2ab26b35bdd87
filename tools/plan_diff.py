"""Compare two plan files record by record (development tool): host-built vs reference-built."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from chiml_b200 import plan as P


def diff(a: P.Plan, b: P.Plan, verbose=True):
    bad = []
    for k in ("mode", "ln", "d", "dt", "has_D", "pml_on_D", "n_objects", "rank", "nranks", "y_start", "n_global", "n_steps",
              "n_lor_poles", "n_ordip_poles", "t_max"):
        if getattr(a, k) != getattr(b, k):
            bad.append(f"grid.{k}: {getattr(a, k)} != {getattr(b, k)}")
    if a.cplx != b.cplx or tuple(a.k_point) != tuple(b.k_point):
        bad.append(f"complex: {a.cplx} {a.k_point} != {b.cplx} {b.k_point}")
    for q, (sa, sb) in enumerate(zip(a.sources, b.sources)):
        if (sa.amp_im is None) != (sb.amp_im is None) or (sa.amp_im is not None and np.asarray(sa.amp_im).tobytes() != np.asarray(sb.amp_im).tobytes()):
            bad.append(f"source {q}: imaginary amplitudes differ")
    for k in ("has_B", "pml_on_B", "n_mag_poles"):
        if getattr(a, k) != getattr(b, k):
            bad.append(f"magnetic.{k}: {getattr(a, k)} != {getattr(b, k)}")
    for nm in ("mag_objects", "chi_objects"):
        da, db = getattr(a, nm), getattr(b, nm)
        if sorted(da) != sorted(db):
            bad.append(f"{nm}: objects {sorted(da)} != {sorted(db)}")
        for o in sorted(set(da) & set(db)):
            for i, (u, v) in enumerate(zip(da[o], db[o])):
                if np.asarray(u).tobytes() != np.asarray(v).tobytes():
                    bad.append(f"{nm}[{o}][{i}]: {u} != {v}")
    if (a.prev_copy is None) != (b.prev_copy is None) or (a.prev_copy is not None and np.asarray(a.prev_copy).tobytes() != np.asarray(b.prev_copy).tobytes()):
        bad.append(f"prev_copy: {None if a.prev_copy is None else a.prev_copy.shape} != {None if b.prev_copy is None else b.prev_copy.shape}")
    if sorted(a.dip_grids) != sorted(b.dip_grids):
        bad.append(f"dip grids: {sorted(a.dip_grids)} != {sorted(b.dip_grids)}")
    else:
        # compared at the cells of the oriented-dipole node list: the only ones the update reads (the reference's grids hold whatever its
        # constructor found at the other cells, out-of-range pole indices of other objects included)
        runs = b.get_list(P.LIST_ORDIPP, 0)
        npol = {o.obj: o.npoles for o in b.objects}
        for (comp, pole), gb in sorted(b.dip_grids.items()):
            ga = a.dip_grids[(comp, pole)]
            for r in runs:
                if pole >= npol.get(int(r["obj"]), 0):
                    continue
                sl = slice(int(r["ind"]), int(r["ind"]) + int(r["n"]))
                if np.asarray(ga[sl]).tobytes() != np.asarray(gb[sl]).tobytes():
                    i = next(i for i in range(int(r["n"])) if ga[sl][i] != gb[sl][i])
                    bad.append(f"dip grid {(comp, pole)}: cell {int(r['ind']) + i}: {ga[sl][i]!r} != {gb[sl][i]!r}")
                    break
    if a.periodic != b.periodic:
        bad.append(f"periodic: {a.periodic} != {b.periodic}")
    for key in sorted(set(a.lists) | set(b.lists)):
        la, lb = a.get_list(*key), b.get_list(*key)
        if len(la) != len(lb):
            bad.append(f"list{key}: len {len(la)} != {len(lb)}")
            n = min(len(la), len(lb))
            idx = [i for i in range(n) if la[i] != lb[i]][:1]
            if idx:
                bad.append(f"   first diff at {idx[0]}: {la[idx[0]]} vs {lb[idx[0]]}")
            elif n < len(la): bad.append(f"   extra in A: {la[n]}")
            elif n < len(lb): bad.append(f"   extra in B: {lb[n]}")
        elif la.tobytes() != lb.tobytes():
            i = next(i for i in range(len(la)) if la[i] != lb[i])
            bad.append(f"list{key}: first diff at {i}: {la[i]} vs {lb[i]}")
    if len(a.objects) != len(b.objects):
        bad.append(f"objects: {len(a.objects)} != {len(b.objects)}")
    for oa, ob in zip(a.objects, b.objects):
        for k in ("obj", "npoles", "use_or_dip", "ml", "eps_inf", "mu_inf"):
            if getattr(oa, k) != getattr(ob, k):
                bad.append(f"object {oa.obj}.{k}: {getattr(oa, k)} != {getattr(ob, k)}")
        for k in ("alpha", "xi", "gamma", "dip"):
            if np.asarray(getattr(oa, k)).tobytes() != np.asarray(getattr(ob, k)).tobytes():
                bad.append(f"object {oa.obj}.{k}: {getattr(oa, k)} != {getattr(ob, k)}")
    ca = {(c.comp, c.part): c for c in a.cpml}
    cb = {(c.comp, c.part): c for c in b.cpml}
    for key in sorted(set(ca) | set(cb)):
        if key not in ca or key not in cb:
            bad.append(f"cpml{key}: present only in {'A' if key in ca else 'B'}")
            continue
        x, y = ca[key], cb[key]
        if x.has_psi != y.has_psi:
            bad.append(f"cpml{key}.has_psi {x.has_psi} != {y.has_psi}")
        for nm in ("psi", "grid"):
            u, v = getattr(x, nm), getattr(y, nm)
            if len(u) != len(v):
                bad.append(f"cpml{key}.{nm}: len {len(u)} != {len(v)}")
            elif u.tobytes() != v.tobytes():
                i = next(i for i in range(len(u)) if u[i] != v[i])
                bad.append(f"cpml{key}.{nm}: first diff at {i}/{len(u)}: {u[i]} vs {v[i]}")
    if len(a.sources) != len(b.sources):
        bad.append(f"sources: {len(a.sources)} != {len(b.sources)}")
    for i, (sa, sb) in enumerate(zip(a.sources, b.sources)):
        if (sa.field, sa.loc, sa.sz) != (sb.field, sb.loc, sb.sz):
            bad.append(f"source {i}: {(sa.field, sa.loc, sa.sz)} != {(sb.field, sb.loc, sb.sz)}")
        if sa.amp.tobytes() != sb.amp.tobytes():
            n = min(len(sa.amp), len(sb.amp))
            j = next((j for j in range(n) if sa.amp[j] != sb.amp[j]), n)
            bad.append(f"source {i}: amp len {len(sa.amp)} vs {len(sb.amp)}, first diff at {j}: "
                       f"{sa.amp[j] if j < len(sa.amp) else None!r} vs {sb.amp[j] if j < len(sb.amp) else None!r}")
    if len(a.detectors) != len(b.detectors):
        bad.append(f"detectors: {len(a.detectors)} != {len(b.detectors)}")
    for i, (da, db) in enumerate(zip(a.detectors, b.detectors)):
        if da != db:
            bad.append(f"detector {i}: {da} != {db}")
    if len(a.emitters) != len(b.emitters):
        bad.append(f"emitters: {len(a.emitters)} != {len(b.emitters)}")
    for i, (ea, eb) in enumerate(zip(a.emitters, b.emitters)):
        for k in ("object", "nlevel", "nsys", "nemit", "box_lo", "box_n", "pz", "npop", "pop_every", "npoints", "dt", "inv_hbar", "na"):
            if getattr(ea, k) != getattr(eb, k):
                bad.append(f"emitter {i}.{k}: {getattr(ea, k)!r} != {getattr(eb, k)!r}")
        for k in ("h0", "weight", "mu", "gam_ptr", "gam_col", "gam_val", "loc", "eps", "pop_level"):
            u, v = np.asarray(getattr(ea, k)), np.asarray(getattr(eb, k))
            if u.shape != v.shape:
                bad.append(f"emitter {i}.{k}: shape {u.shape} != {v.shape}")
            elif u.tobytes() != v.tobytes():
                j = np.flatnonzero(u.ravel() != v.ravel())
                bad.append(f"emitter {i}.{k}: {len(j)} entries differ, first at {j[:1]}: {u.ravel()[j[:1]]} vs {v.ravel()[j[:1]]}")
    if len(a.dfts) != len(b.dfts):
        bad.append(f"dft sets: {len(a.dfts)} != {len(b.dfts)}")
    for i, (da, db) in enumerate(zip(a.dfts, b.dfts)):
        for k in ("field", "group", "every", "nfreq", "npts", "stride", "acc_len"):
            if getattr(da, k) != getattr(db, k):
                bad.append(f"dft {i}.{k}: {getattr(da, k)} != {getattr(db, k)}")
        if da.freq.tobytes() != db.freq.tobytes() or da.lines.tobytes() != db.lines.tobytes():
            bad.append(f"dft {i}: frequency list or line list differs")
    return bad


if __name__ == "__main__":
    bad = diff(P.read_plan(sys.argv[1]), P.read_plan(sys.argv[2]))
    print("\n".join(bad) if bad else "plans identical")
    sys.exit(1 if bad else 0)
