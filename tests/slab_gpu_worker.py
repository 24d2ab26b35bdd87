"""One process per GPU (torch.distributed.run): every rank builds its y-slab plan, binds the CUDA engine's native halo
(peer-to-peer stores over NVLink), steps, and rank 0 compares the gathered owned rows with the single-rank output of the
reference (tests/golden/<case>.expect.npz) bit for bit.  usage: slab_gpu_worker.py <case> [<case> ...]"""
import os
import subprocess
import sys
import tempfile

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from chiml_b200 import capi, plan as P  # noqa: E402


def add_second_species(plan):
    """A second emitter species in the SAME region: every emitter set of the plan is duplicated with another object index, density
    and dipole strength.  The two sets of a slab then have equal boxes, so across a slab boundary only ChimlEmitterDesc::object
    tells them apart (chiml_gpu_halo_bind).  The reference cannot run two emitter objects in one input (tests/golden/make_golden.py),
    so the expected result comes from the single-rank oracle."""
    import dataclasses
    plan.emitters += [dataclasses.replace(e, object=e.object + 100, na=0.6 * e.na, mu=0.7 * e.mu) for e in plan.emitters]


def oracle_expect(whole, names):
    """State arrays of the single-rank oracle after whole.n_steps steps, keyed like a reference dump."""
    from oracle_api import OracleSim
    cpu = OracleSim(whole)
    cpu.step_n(whole.n_steps)
    out = {n: np.array(util.state_array(cpu, n)) for n in names}
    for q, e in enumerate(whole.emitters):
        for d in range(e.npop):
            a = cpu.population(q, d)
            out[f"q{q}pop{d}"] = np.stack([a.real, a.imag], axis=1)[:, None, :]
    cpu.close()
    return out


def run_case(case, rank, world, local, log=print):
    """Returns True when the gathered slabs equal the single-rank result bit for bit (rank 0 decides; other ranks return True).
    `log` receives the per-case verdict and any mismatch lines (bench.py sends them to stderr)."""
    case, pair = (case[:-5], True) if case.endswith("+pair") else (case, False)
    work = tempfile.mkdtemp(prefix=f"slabgpu_{case.replace(':', '_')}_r{rank}_")
    tool = os.path.join(ROOT, "chiml_b200", "chiml_plan")
    fuzz = case.startswith("fuzz:")
    if fuzz:
        # fuzz:<seed>:pbc / fuzz:<seed>:mag -- a random periodic / magnetic input (tests/fuzz/gen_inputs.py); expected arrays from the single-rank oracle
        sys.path.insert(0, os.path.join(ROOT, "tests", "fuzz"))
        import gen_inputs
        from chiml_b200 import inputs as I
        _, seed, kind = case.split(":")
        cfg = gen_inputs.rnd_pbc_case(int(seed)) if kind == "pbc" else gen_inputs.rnd_mag_case(int(seed), steps=12)
        src = os.path.join(work, "c.json")
        I.write(cfg, src)
        subprocess.run([tool, src, os.path.join(work, "whole")], check=True)
        whole = P.read_plan(os.path.join(work, "whole.rank0.plan"))
        label, case = case, "c"
    else:
        src = os.path.join(util.GOLDEN, case + ".json")
        whole = util.load_plan(case)
        label = case
    subprocess.run([tool, src, os.path.join(work, case), "--ranks", str(world), "--only", str(rank)], check=True)
    plan = P.read_plan(os.path.join(work, f"{case}.rank{rank}.plan"))
    if pair:
        add_second_species(plan)
        add_second_species(whole)
    sim = capi.GpuSim(plan, device=local)
    sim.halo_bind(dist, rank, world)
    # several calls of uneven length: the flags count steps across calls
    done = 0
    for n in (1, 2, whole.n_steps - 3):
        sim.step_n(n)
        done += n
    sim.sync()
    ny = plan.ln[1] - 2
    names = [n for n in util.state_names(whole) if not n.startswith("q") and not n.startswith("dft")]
    mine = {n: np.ascontiguousarray(util.state_array(sim, n)[1:ny + 1]) for n in names}
    # running-DFT accumulators of this slab's parts of the flux surfaces, keyed by (region, field, global point, frequency)
    mine["__dft__"] = util.dft_point_map(plan, [sim.dft(k) for k in range(len(plan.dfts))])
    emit = []
    for q, e in enumerate(plan.emitters):
        coords = np.stack([e.box_lo[0] + e.loc[:, 0], e.box_lo[1] + e.loc[:, 1] + plan.y_start, e.box_lo[2] + e.loc[:, 2]], axis=1) if e.nemit else np.zeros((0, 3), int)
        emit.append((e.object, coords, [[sim.emitter_state(q, sy, w).copy() for w in range(5)] for sy in range(e.nsys)],
                     [sim.population(q, d) for d in range(e.npop)]))
    launches = sim.launch_count()
    sim.close()
    gathered = [None] * world
    dist.gather_object((plan.y_start, mine, emit), gathered if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        expect = oracle_expect(whole, util.state_names(whole)) if (pair or fuzz) else util.load_expect(case)
        gathered.sort(key=lambda t: t[0])
        for n in names:
            got = np.concatenate([g[1][n] for g in gathered], axis=0)
            ref = expect[n][1:-1]
            if whole.periodic and n == "Ey":
                # the wrap row of Ey (global row ln_y, the image of row 1): nothing in the step reads it, the slab ring does not carry it
                got, ref = got[:-1], ref[:-1]
            if not np.array_equal(got, ref):
                ok = False
                log(f"MISMATCH {case}/{n}: max |diff| {np.abs(got - ref).max():.3e} of {np.abs(ref).max():.3e}")
        if whole.dfts:
            ref = util.dft_point_map(whole, [expect[f"dft{k}r"].ravel() + 1j * expect[f"dft{k}i"].ravel() for k in range(len(whole.dfts))])
            got = {}
            for g in gathered:
                for key, v in g[1]["__dft__"].items():
                    if key in got and got[key] != v:
                        ok = False
                        log(f"MISMATCH {case}: accumulator {key} differs between slabs")
                    got[key] = v
            if set(got) != set(ref):
                ok = False
                log(f"MISMATCH {case}: the slabs hold {len(got)} DFT accumulators, the single-rank run {len(ref)}")
            else:
                nbad = sum(1 for key in ref if ref[key] != got[key])
                if nbad:
                    ok = False
                    log(f"MISMATCH {case}: {nbad} of {len(ref)} DFT accumulators differ from the reference")
        for q, e in enumerate(whole.emitters):
            gcoord = np.stack([e.box_lo[0] + e.loc[:, 0], e.box_lo[1] + e.loc[:, 1], e.box_lo[2] + e.loc[:, 2]], axis=1)
            index = {tuple(c): i for i, c in enumerate(gcoord)}
            seen = 0
            pops = None
            for _, _, em in gathered:
                for obj, coords, states, pop in em:
                    if obj != e.object:
                        continue
                    idx = np.array([index[tuple(c)] for c in coords], dtype=int)
                    seen += len(idx)
                    pops = pop if pops is None else [a + b for a, b in zip(pops, pop)]   # the host adds the slabs (QEPopDtc::toFile)
                    for sy in range(e.nsys):
                        for w in range(5):
                            r = expect[f"q{q}s{sy}w{w}"][:, 0, :]
                            ref = (r[:, 0::2] + 1j * r[:, 1::2])[idx]
                            if not np.array_equal(states[sy][w], ref):
                                ok = False
                                log(f"MISMATCH {case}/q{q}s{sy}w{w}: max |diff| {np.abs(states[sy][w] - ref).max():.3e}")
            if seen != e.nemit:
                ok = False
                log(f"MISMATCH {case}: {seen} emitters over the slabs, {e.nemit} in the single-rank run")
            for d in range(e.npop):
                r = expect[f"q{q}pop{d}"][:, 0, :]
                ref = r[:, 0] + 1j * r[:, 1]
                if pops is None or len(pops[d]) != len(ref) or np.abs(pops[d] - ref).max() > 1e-9 * max(np.abs(ref).max(), 1e-300):
                    ok = False
                    log(f"MISMATCH {case}/q{q}pop{d}")
        log(f"{label}{'+pair' if pair else ''}: {'SLAB_GPU_OK' if ok else 'SLAB_GPU_FAIL'} ({world} slabs, {launches} launches on rank 0)")
    return ok


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # more ranks than GPUs: several slabs share a device (their contexts are time-sliced; the peer stores go through CUDA IPC all the same)
    local = int(os.environ.get("LOCAL_RANK", "0")) % max(1, capi.device_count())
    ok = True
    for case in sys.argv[1:]:
        ok = run_case(case, rank, world, local) and ok
        dist.barrier()
    import torch
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
