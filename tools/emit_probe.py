"""Density-matrix kernels on a large emitter sheet: one thread per emitter against a group of lanes per emitter (csrc/chiml_emitters.cuh),
for two- and four-level emitters (development tool).  usage: emit_probe.py [sheet=1000] [steps=10]"""
import json
import os
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chiml_b200 import capi, inputs as I, plan as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sheet = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
res = 100
dt = I.default_dt(res)
n, nz = sheet + 60, 47
work = tempfile.mkdtemp(prefix="emit_probe_")
relax1 = [{"state_i": 1, "state_f": 0, "rate": 1e12, "dephasing_rate": 1e13}]
objs = {
    "N2": I.ml_object([(sheet - 1) / res, (sheet - 1) / res, 0.0], [0, 0, 0.0], 1e25, [(0, 0), (1, 0)], [{"E_cen": [0.0]}, {"E_cen": [2.0]}],
                      [0, 10.0, 10.0, 0], relax1, eps=1.0, dtc_levs=[3]),
    "N4": I.ml_object([(sheet - 1) / res, (sheet - 1) / res, 0.0], [0, 0, 0.0], 1e25, [(0, 0), (1, -1), (1, 0), (1, 1)],
                      [{"E_cen": [0.0]}, {"E_cen": [1.9, 2.1], "weights": [0.6, 0.4], "levs_described": 3}],
                      [0, 10, 8, 6, 10, 0, 0, 0, 8, 0, 0, 0, 6, 0, 0, 0],
                      [{"state_i": 1, "state_f": 0, "rate": 1e12, "dephasing_rate": 1e13}, {"state_i": 2, "state_f": 0, "rate": 2e12, "dephasing_rate": 0.5e13},
                       {"state_i": 3, "state_f": 0, "rate": 1.5e12}], eps=1.0, dtc_levs=[5]),
}
for name, obj in objs.items():
    cfg = I.config(I.comp_cell([(n - 1) / res, (n - 1) / res, (nz - 1) / res], res, (steps + 8) * dt - 0.5 * dt, "Ex"), I.pml([10 / res] * 3),
                   [I.normal_source("Ex", [0, 0, 0.1], [0, 0, 0], [I.gaussian_pulse(1.5, 1.0)])], [obj], [])
    I.write(cfg, os.path.join(work, name + ".json"))
    subprocess.run([os.path.join(ROOT, "chiml_b200", "chiml_plan"), os.path.join(work, name + ".json"), os.path.join(work, name)], check=True)
    plan = P.read_plan(os.path.join(work, name + ".rank0.plan"))
    for kernel in ("thread", "group"):
        os.environ["CHIML_B200_EMIT_KERNEL"] = kernel
        sim = capi.GpuSim(plan, detectors=False)
        sim.step_n(4)
        sim.sync()
        sim.reset_kernel_stats()
        sim.set_kernel_timing(True)
        sim.step_n_timed(steps)
        st = {s["name"]: s for s in sim.kernel_stats()}["k_emit_density"]
        ms = st["ms_total"] / steps
        print(json.dumps({"emitters": plan.emitters[0].nemit, "levels": plan.emitters[0].nlevel, "level_systems": plan.emitters[0].nsys, "kernel": kernel,
                          "ms": round(ms, 4), "alg_GBps": round(st["alg_bytes_per_step"] / (ms * 1e-3) / 1e9)}))
        sim.close()
