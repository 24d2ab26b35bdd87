#!/bin/bash
# A/B of library builds on ONE box (box-to-box spread is a few per cent): tools/ab.sh <tag> lib1.so lib2.so ...
# runs the default bench with each library (CHIML_B200_LIB) and prints ms per step and the per-kernel averages.
tag=$1; shift
mkdir -p gpurun_out
for rep in 1 2; do
for lib in "$@"; do
    n=$(basename $lib .so)
    CHIML_B200_LIB=$lib timeout 600 python bench.py --steps 12 --warmup 4 --no-cpu-baseline > gpurun_out/${tag}_${n}_$rep.json 2> gpurun_out/${tag}_${n}_$rep.err
    python - "$n" gpurun_out/${tag}_${n}_$rep.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    ks = {k["name"]: k["ms_per_step"] for k in d["roofline"]["kernels"]}
    print(f'{sys.argv[1]:>16s} {d["ms_per_step"]:7.3f} ms  ' + "  ".join(f'{n}={ks[n]:.3f}' for n in ("k_fast<E>", "k_uniform<E>", "k_general<E>", "k_fast<H>", "k_uniform<H>", "k_general<H>", "k_ordip_poles") if n in ks))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
done
