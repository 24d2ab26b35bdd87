"""Development aid: steps a periodic fixture on N slabs one step at a time and reports the first step / cell where the gathered fields leave
the single-rank oracle.  usage: torchrun --nproc-per-node N tools/ring_debug.py <case> [steps_per_call]"""
import os, subprocess, sys, tempfile
import numpy as np
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from chiml_b200 import capi, plan as P
from oracle_api import OracleSim

case = sys.argv[1]
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
work = tempfile.mkdtemp()
subprocess.run([os.path.join(ROOT, "chiml_b200", "chiml_plan"), os.path.join(util.GOLDEN, case + ".json"), os.path.join(work, case), "--ranks", str(world), "--only", str(rank)], check=True)
plan = P.read_plan(os.path.join(work, f"{case}.rank{rank}.plan"))
whole = util.load_plan(case)
sim = capi.GpuSim(plan, device=int(os.environ.get("LOCAL_RANK", 0)) % capi.device_count())
sim.halo_bind(dist, rank, world)
cpu = OracleSim(whole) if rank == 0 else None
ny = plan.ln[1] - 2
names = ["Ex", "Ey", "Ez", "Hx", "Hy", "Hz"]
done = 0
bad = False
while done < whole.n_steps and not bad:
    n = min(chunk, whole.n_steps - done)
    sim.step_n(n); sim.sync(); done += n
    mine = {nm: np.ascontiguousarray(util.state_array(sim, nm)[1:ny + 1]) for nm in names}
    g = [None] * world
    dist.gather_object((plan.y_start, mine), g if rank == 0 else None, dst=0)
    if rank == 0:
        cpu.step_n(n)
        g.sort(key=lambda t: t[0])
        for nm in names:
            got = np.concatenate([x[1][nm] for x in g], axis=0)
            ref = util.state_array(cpu, nm)[1:-1]
            if nm == "Ey": got, ref = got[:-1], ref[:-1]
            if not np.array_equal(got, ref):
                idx = np.argwhere(got != ref)
                ys = sorted(set(int(i[0]) + 1 for i in idx))
                print(f"step {done}: {nm} differs in {len(idx)} cells, global rows {ys[:12]}, z {sorted(set(int(i[1]) for i in idx))[:8]}, x {sorted(set(int(i[2]) for i in idx))[:8]}, max {np.abs(got-ref).max():.2e} of {np.abs(ref).max():.2e}; slab starts {[x[0] for x in g]}")
                bad = True
    flag = [bad]
    dist.broadcast_object_list(flag, src=0)
    bad = flag[0]
if rank == 0 and not bad: print("RING_DEBUG_OK")
sim.close()
dist.destroy_process_group()
