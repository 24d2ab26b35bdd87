"""Shared helpers of the parity tests."""
import glob
import os

import numpy as np

from chiml_b200 import plan as P

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-len(".rank0.plan")] for p in glob.glob(os.path.join(GOLDEN, "*.rank0.plan")))


def load_plan(case: str) -> P.Plan:
    return P.read_plan(os.path.join(GOLDEN, case + ".rank0.plan"))


def load_expect(case: str):
    with np.load(os.path.join(GOLDEN, case + ".expect.npz")) as z:
        return {k: z[k] for k in z.files}


def state_array(sim, name: str):
    """Fetch the array the reference dump calls `name` (oracle/ref_driver.cpp grabGrid names) from an
    OracleSim or a GpuSim (both expose field / pole / ordip_pole)."""
    if name.endswith("_im"):
        return state_array(sim.imag, name[:-3])          # complex fields: the imaginary part is the second simulation of the pair
    if name in P.FIELD_NAMES:
        return sim.field(P.FIELD_NAMES.index(name))
    if name.startswith("dft"):
        a = sim.dft(int(name[3:-1]))
        return (a.imag if name.endswith("i") else a.real).reshape(1, 1, -1).copy()
    if name.startswith("q"):
        # emitter arrays of the reference dump: q<slot>s<sys>w<which> (states), q<slot>P<c> (P boxes), q<slot>pop<det>
        import re
        m = re.fullmatch(r"q(\d+)s(\d+)w(\d)", name)
        if m:
            a = sim.emitter_state(int(m.group(1)), int(m.group(2)), int(m.group(3)))
            out = np.empty((a.shape[0], 1, 2 * a.shape[1]))
            out[:, 0, 0::2], out[:, 0, 1::2] = a.real, a.imag
            return out
        m = re.fullmatch(r"q(\d+)P([xyz])", name)
        if m:
            return np.asarray(sim.emitter_P(int(m.group(1)), "xyz".index(m.group(2))))
        m = re.fullmatch(r"q(\d+)pop(\d+)", name)
        if m:
            a = sim.population(int(m.group(1)), int(m.group(2)))
            out = np.empty((a.shape[0], 1, 2))
            out[:, 0, 0], out[:, 0, 1] = a.real, a.imag
            return out
        raise KeyError(name)
    # chiral media: vE / vH = prevE_ / prevH_; cP / cvP = lorChiHP_ / its previous value; cM / cvM = lorChiEM_ / previous
    if name[:2] in ("vE", "vH") and len(name) == 3:
        return sim.prev_field((0 if name[1] == "E" else 3) + "xyz".index(name[2]))
    if name.startswith("cvP") or name.startswith("cvM"):
        return sim.chi_pole((0 if name[2] == "P" else 3) + "xyz".index(name[3]), int(name[4:]), 1)
    if name.startswith("cP") or name.startswith("cM"):
        return sim.chi_pole((0 if name[1] == "P" else 3) + "xyz".index(name[2]), int(name[3:]), 0)
    if name.startswith("pM"):
        return sim.mag_pole("xyz".index(name[2]), int(name[3:]), 1)
    if name.startswith("M"):
        return sim.mag_pole("xyz".index(name[1]), int(name[2:]), 0)
    if name.startswith("poP"):
        return sim.ordip_pole("xyz".index(name[3]), int(name[4:]), 1)
    if name.startswith("oP"):
        return sim.ordip_pole("xyz".index(name[2]), int(name[3:]), 0)
    if name.startswith("pP"):
        return sim.pole("xyz".index(name[2]), int(name[3:]), 1)
    if name.startswith("P"):
        return sim.pole("xyz".index(name[1]), int(name[2:]), 0)
    raise KeyError(name)


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    den = float(np.linalg.norm(b.ravel()))
    num = float(np.linalg.norm((a - b).ravel()))
    return num / den if den > 0 else num


def state_names(plan: P.Plan):
    names = [P.FIELD_NAMES[f] for f in plan.fields_present()]
    for c in range(3):
        if (6 + c) in plan.fields_present():
            for p in range(plan.n_lor_poles):
                names += [f"P{'xyz'[c]}{p}", f"pP{'xyz'[c]}{p}"]
            for p in range(plan.n_ordip_poles):
                names += [f"oP{'xyz'[c]}{p}", f"poP{'xyz'[c]}{p}"]
        if (9 + c) in plan.fields_present():
            for p in range(plan.n_mag_poles):
                names += [f"M{'xyz'[c]}{p}", f"pM{'xyz'[c]}{p}"]
    if plan.chi_objects:
        names += [f"v{f}{c}" for f in "EH" for c in "xyz"]
        for c in "xyz":
            for p in range(plan.n_chi_poles):
                names += [f"cP{c}{p}", f"cvP{c}{p}", f"cM{c}{p}", f"cvM{c}{p}"]
    for k in range(len(plan.dfts)):
        names += [f"dft{k}r", f"dft{k}i"]
    for q, e in enumerate(plan.emitters):
        names += [f"q{q}s{s}w{w}" for s in range(e.nsys) for w in range(5)]
        names += [f"q{q}P{'xyz'[c]}" for c in range(3) if c in plan.fields_present()]
    return names


def dft_point_map(plan: P.Plan, accs):
    """Running-DFT accumulators keyed by what they ARE rather than by where a rank stores them: (flux region, field, global grid
    point, frequency) -> complex value.  `accs[k]` = complex accumulator array of plan.dfts[k].  An accumulator is a pure function
    of its key (a field sample times the region's twiddle, summed over the sampled steps), so maps of different slab
    decompositions of the same run must agree entry for entry; entries reached through two stored fields (box edges) must agree
    with each other."""
    lnx, lny, lnz = plan.ln
    out = {}
    for d, acc in zip(plan.dfts, accs):
        for li, (ind, o) in enumerate(d.lines):
            if li > 0 and ind == 0 and o == 0:
                continue                      # unfilled tail entries of fInGridInds_
            for i in range(d.npts):
                g = int(ind) + i * d.stride
                row, x = divmod(g, lnx)
                y, z = divmod(row, lnz)
                for f in range(d.nfreq):
                    key = (d.group, d.field, x, y + plan.y_start, z, f)
                    v = complex(acc[int(o) + f + d.nfreq * i])
                    if key in out and out[key] != v:
                        raise AssertionError(f"accumulator {key} stored twice with different values")
                    out[key] = v
    return out


def assert_slabs_tile_whole(whole, slabs):
    """y-slab decomposition of the host-side setup: the update-list cells of the slab plans, mapped back to global rows, are exactly
    the cells of the single-rank plan (same prefactors), every emitter is owned by exactly one slab, and the running-DFT lines of the
    slabs cover the single-rank sets' points exactly once."""
    lnx, lny, lnz = whole.ln
    assert sum(s.ln[1] - 2 for s in slabs) == lny - 2
    for key, runs in whole.lists.items():
        def cells(plan, runs, ys):
            m = {}
            for r in runs:
                row, x0 = divmod(int(r["ind"]), plan.ln[0])
                y, z = divmod(row, plan.ln[2])
                for i in range(int(r["n"])):
                    m[(x0 + i, y + ys, z)] = (float(r["pf"][1]), float(r["pf"][2]), float(r["pf"][3]))
            return m
        ref = cells(whole, runs, 0)
        got = {}
        for s in slabs:
            part = cells(s, s.get_list(*key), s.y_start)
            assert not (set(part) & set(got)), f"list {key}: a cell is owned by two slabs"
            got.update(part)
        assert got == ref, f"list {key}: slabs do not tile the single-rank list"
    if whole.emitters:
        assert sum(e.nemit for s in slabs for e in s.emitters) == sum(e.nemit for e in whole.emitters)

    def dft_points(plan):
        pts = {}
        for d in plan.dfts:
            for li, (ind, o) in enumerate(d.lines):
                if li > 0 and ind == 0 and o == 0:
                    continue
                for i in range(d.npts):
                    row, x = divmod(int(ind) + i * d.stride, plan.ln[0])
                    y, z = divmod(row, plan.ln[2])
                    key = (d.group, d.field, x, y + plan.y_start, z)
                    pts[key] = pts.get(key, 0) + 1
        return pts
    ref = dft_points(whole)
    got = {}
    for s in slabs:
        for k, v in dft_points(s).items():
            got[k] = got.get(k, 0) + v
    assert got == ref, "the slabs' running-DFT lines do not cover the single-rank sets exactly"
