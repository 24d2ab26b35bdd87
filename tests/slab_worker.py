"""World_size-N worker of the slab tests (launched with torch.distributed.run, backend gloo): every rank builds the plan of
its y-slab with the host-side setup, steps the CPU checker through chiml_b200.slab.step_slab, and rank 0 compares the gathered
owned rows with the single-rank output of the reference (tests/golden/<case>.expect.npz) bit for bit."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from chiml_b200 import plan as P, slab  # noqa: E402
from oracle_api import OracleSim  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    all_ok = True
    for case in sys.argv[1:]:          # several cases per launch: the start-up of the process group costs more than a case
        all_ok = run_case(case, rank, world) and all_ok
    dist.destroy_process_group()
    sys.exit(0 if all_ok else 1)


def run_case(label, rank, world):
    case = label
    work = tempfile.mkdtemp(prefix=f"slab_{case.replace(':', '_')}_r{rank}_")
    tool = os.path.join(ROOT, "chiml_b200", "chiml_plan")
    fuzz = case.startswith("fuzz:")
    if fuzz:
        # fuzz:<seed>[:ml] -- a random input (tests/fuzz/gen_inputs.py); the expected arrays come from the single-rank oracle
        sys.path.insert(0, os.path.join(ROOT, "tests", "fuzz"))
        import gen_inputs
        from chiml_b200 import inputs as I
        parts = case.split(":")
        # fuzz:<seed>:pbc -- a periodic run with random media (the slabs form a ring); fuzz:<seed>:mag -- magnetic-dispersive media
        if len(parts) > 2 and parts[2] == "pbc":
            cfg = gen_inputs.rnd_pbc_case(int(parts[1]))
        elif len(parts) > 2 and parts[2] == "mag":
            cfg = gen_inputs.rnd_mag_case(int(parts[1]), steps=12)
        else:
            cfg = gen_inputs.rnd_ml_case(int(parts[1])) if len(parts) > 2 else gen_inputs.rnd_case(int(parts[1]), steps=12, pulses="random")
            if len(parts) > 2:
                cfg["CompCell"]["tLim"] = 12 * gen_inputs.DT - 0.5 * gen_inputs.DT
        src = os.path.join(work, "c.json")
        I.write(cfg, src)
        case = "c"
    else:
        src = os.path.join(util.GOLDEN, case + ".json")
    subprocess.run([tool, src, os.path.join(work, case), "--ranks", str(world), "--only", str(rank)], check=True)
    plan = P.read_plan(os.path.join(work, f"{case}.rank{rank}.plan"))
    if fuzz:
        subprocess.run([tool, src, os.path.join(work, "whole")], check=True)
        whole = P.read_plan(os.path.join(work, "whole.rank0.plan"))
    else:
        whole = util.load_plan(case)
    sim = OracleSim(plan)

    def send(to, arr):
        dist.send(torch.from_numpy(arr), dst=to)

    def recv(frm, out):
        t = torch.from_numpy(out)
        dist.recv(t, src=frm)

    for k in range(whole.n_steps):
        slab.step_slab(sim, plan, sim.src_amp(k, 1), send, recv)

    # gather owned rows on rank 0
    ny = plan.ln[1] - 2
    names = [n for n in util.state_names(whole) if not n.startswith("q") and not n.startswith("dft")]
    mine = {n: np.ascontiguousarray(util.state_array(sim, n)[1:ny + 1]) for n in names}
    mine["__dft__"] = util.dft_point_map(plan, [sim.dft(k) for k in range(len(plan.dfts))])
    emit = []
    for q, e in enumerate(plan.emitters):
        coords = np.stack([e.box_lo[0] + e.loc[:, 0], e.box_lo[1] + e.loc[:, 1] + plan.y_start, e.box_lo[2] + e.loc[:, 2]], axis=1) if e.nemit else np.zeros((0, 3), int)
        emit.append((e.object, coords, [[sim.emitter_state(q, sy, w).copy() for w in range(5)] for sy in range(e.nsys)]))
    gathered = [None] * world
    dist.gather_object((plan.y_start, mine, emit), gathered if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        if fuzz:
            one = OracleSim(whole)
            one.step_n(whole.n_steps)
            expect = {n: util.state_array(one, n) for n in util.state_names(whole)}
        else:
            expect = util.load_expect(case)
        gathered.sort(key=lambda t: t[0])
        for n in names:
            got = np.concatenate([g[1][n] for g in gathered], axis=0)
            ref = expect[n][1:-1]
            if whole.periodic and n == "Ey":
                # the wrap row of Ey (global row ln_y, the image of row 1): nothing in the step reads it, the slab ring does not carry it
                got, ref = got[:-1], ref[:-1]
            if not np.array_equal(got, ref):
                ok = False
                print(f"MISMATCH {case}/{n}: max |diff| {np.abs(got - ref).max():.3e} of {np.abs(ref).max():.3e}")
        if whole.dfts:
            # running-DFT accumulators of the flux regions: the slabs' parts, keyed by (region, field, global point, frequency),
            # are together exactly the single-rank reference's accumulators
            ref = util.dft_point_map(whole, [expect[f"dft{k}r"].ravel() + 1j * expect[f"dft{k}i"].ravel() for k in range(len(whole.dfts))])
            got = {}
            for g in gathered:
                got.update(g[1]["__dft__"])
            if set(got) != set(ref):
                ok = False
                print(f"MISMATCH {case}: the slabs hold {len(got)} DFT accumulators, the single-rank run {len(ref)}")
            else:
                nbad = sum(1 for key in ref if ref[key] != got[key])
                if nbad:
                    ok = False
                    print(f"MISMATCH {case}: {nbad} of {len(ref)} DFT accumulators differ from the reference")
        for q, e in enumerate(whole.emitters):
            gcoord = np.stack([e.box_lo[0] + e.loc[:, 0], e.box_lo[1] + e.loc[:, 1], e.box_lo[2] + e.loc[:, 2]], axis=1)
            index = {tuple(c): i for i, c in enumerate(gcoord)}
            seen = 0
            for _, _, em in gathered:
                for obj, coords, states in em:
                    if obj != e.object:
                        continue
                    idx = np.array([index[tuple(c)] for c in coords], dtype=int)
                    seen += len(idx)
                    for sy in range(e.nsys):
                        for w in range(5):
                            r = expect[f"q{q}s{sy}w{w}"][:, 0, :]
                            ref = (r[:, 0::2] + 1j * r[:, 1::2])[idx]
                            if not np.array_equal(states[sy][w], ref):
                                ok = False
                                print(f"MISMATCH {case}/q{q}s{sy}w{w}: max |diff| {np.abs(states[sy][w] - ref).max():.3e}")
            if seen != e.nemit:
                ok = False
                print(f"MISMATCH {case}: {seen} emitters over the slabs, {e.nemit} in the single-rank run")
        print(f"{label}: " + ("SLAB_OK" if ok else "SLAB_FAIL"), flush=True)
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, src=0)
    sim.close()
    return int(flag.item()) == 1


if __name__ == "__main__":
    main()
