"""Which CPML face orientation costs what: 3-D vacuum grids with the CPML on ONE axis only (and on all three), per-kernel
algorithmic GB/s of the UNIFORM kernels (development tool).  usage: face_probe.py [n=384] [steps=10]"""
import json
import os
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chiml_b200 import capi, inputs as I, plan as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
res = 100
dt = I.default_dt(res)
work = tempfile.mkdtemp(prefix="face_probe_")
for name, th in (("x", [20, 0, 0]), ("y", [0, 20, 0]), ("z", [0, 0, 20]), ("xyz", [20, 20, 20])):
    cfg = I.config(I.comp_cell([(n - 1) / res] * 3, res, (steps + 8) * dt - 0.5 * dt, "Ex"), I.pml([t / res for t in th]),
                   [I.normal_source("Ez", [0, 0, 0], [0, 0, 0], [I.gaussian_pulse(1.5, 1.0)])], [], [])
    I.write(cfg, os.path.join(work, name + ".json"))
    subprocess.run([os.path.join(ROOT, "chiml_b200", "chiml_plan"), os.path.join(work, name + ".json"), os.path.join(work, name)], check=True)
    plan = P.read_plan(os.path.join(work, name + ".rank0.plan"))
    sim = capi.GpuSim(plan, detectors=False)
    sim.step_n(4)
    sim.sync()
    sim.reset_kernel_stats()
    sim.set_kernel_timing(True)
    ms = sim.step_n_timed(steps)
    out = {"faces": name, "ms_per_step": round(ms / steps, 4)}
    for s in sim.kernel_stats():
        if s["timed_launches"] and s["alg_bytes_per_step"]:
            out[s["name"]] = f'{s["alg_bytes_per_step"] / 1e9:.3f} GB {s["alg_bytes_per_step"] / (s["ms_total"] / steps * 1e-3) / 1e9:.0f} GB/s'
    print(json.dumps(out))
    sim.close()
