#!/bin/bash
# A/B of library builds on the 2-D record workloads (C1, C2): tools/ab2d.sh <tag> lib1.so lib2.so ...
tag=$1; shift
for w in c1 c2; do
for lib in "$@"; do
    n=$(basename $lib .so)
    CHIML_B200_LIB=$lib timeout 300 python bench.py --workload $w --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/${tag}_${w}_${n}.json 2> gpurun_out/${tag}_${w}_${n}.err
    python - "$w $n" gpurun_out/${tag}_${w}_${n}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(f'{sys.argv[1]:>24s} {d["ms_per_step"]:8.4f} ms {d["value"]:9.0f} Mcell/s  ' + "  ".join(f'{k["name"]}={k["avg_ms"]:.4f}' for k in d["roofline"]["kernels"] if k["avg_ms"] > 0.002))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
done
