tools/ab.sh ${TAG:-r2e} chiml_b200/libchiml_b200.so
for w in ${WL:-c3 c4}; do
  CHIML_B200_DEBUG_TILES=1 timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG:-r2e}_$w.json 2> gpurun_out/${TAG:-r2e}_$w.err
  grep "chiml tiles" gpurun_out/${TAG:-r2e}_$w.err | head -14
  python - gpurun_out/${TAG:-r2e}_$w.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["config"]["workload"][:40], d["ms_per_step"], d["value"], d["roofline"]["whole_step"])
for k in d["roofline"]["kernels"]:
    if k["avg_ms"] > 0.01: print("   ", k["name"], round(k["avg_ms"],3), k["launches_per_step"], k["alg_GB_per_launch"], k["alg_GBps"])
PY
done
