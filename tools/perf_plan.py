"""Quick device-side throughput check of a plan file (development tool, not the bench contract)."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from chiml_b200 import capi, census, plan as P

path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
plan = P.read_plan(path)
cs = census.census(plan)
sim = capi.GpuSim(plan, detectors=False)
rng = np.random.default_rng(1234)
lnx, lny, lnz = plan.ln
for f in plan.fields_present():
    sim.set_field(f, rng.uniform(-1, 1, size=(lny, lnz, lnx)))
sim.step_n(5, amp=np.zeros((5, max(1, len(plan.sources)))))
sim.sync()
times = []
for rep in range(5):
    ms = sim.step_n_timed(steps, amp=np.zeros((steps, max(1, len(plan.sources)))))
    times.append(ms / steps)
t = float(np.median(times)) * 1e-3
print(json.dumps({"plan": os.path.basename(path), "ms_per_step": t * 1e3, "Mcell_per_s": cs.cells / t / 1e6,
                  "alg_GBps": cs.bytes_per_step / t / 1e9, "census": cs.as_dict(), "device_GB": sim.device_bytes() / 1e9}))
