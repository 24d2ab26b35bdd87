#!/bin/bash
# A/B of one library under two environments on ONE box: tools/ab_env.sh <tag> <workload> "<ENV=1 ...>" : runs bench.py --workload twice with and
# twice without the environment assignment and prints ms per step and the per-kernel times.
tag=$1; wl=$2; envs=$3
mkdir -p gpurun_out
for rep in 1 2; do
for mode in off on; do
    if [ $mode = on ]; then pre="env $envs"; else pre=""; fi
    $pre timeout 600 python bench.py --workload $wl --steps 12 --warmup 4 --no-cpu-baseline > gpurun_out/${tag}_${wl}_${mode}_$rep.json 2> gpurun_out/${tag}_${wl}_${mode}_$rep.err
    python - "$mode" gpurun_out/${tag}_${wl}_${mode}_$rep.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    ks = {k["name"]: k["ms_per_step"] for k in d["roofline"]["kernels"]}
    print(f'{sys.argv[1]:>6s} {d["ms_per_step"]:7.3f} ms  frac {d["roofline"]["whole_step"]["frac"]:.3f}  ' + "  ".join(f'{n}={ks[n]:.3f}' for n in ("k_fast<E>", "k_uniform<E>", "k_fast<H>", "k_uniform<H>", "k_ordip_poles", "k_emit_density") if n in ks))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
done
