python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
for lib in chiml_b200/libchiml_b200.so chiml_b200/variants/lib_prev.so; do
  CHIML_B200_LIB=$lib timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', round(d['ms_per_step'],3), {k['name']:round(k['avg_ms'],3) for k in d['roofline']['kernels'] if k['avg_ms']>0.1})"
done
