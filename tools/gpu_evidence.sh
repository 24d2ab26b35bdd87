#!/bin/bash
# One gpurun call that gathers the round's measured evidence: bench lines of every BASELINE configuration (with the CPU reference beside
# them), the ncu launch list of the default bench, one --set full capture of the field kernels and one of the emitter kernels on the
# 1e6-emitter C4 workload.   usage: tools/gpu_evidence.sh <tag>   (outputs under gpurun_out/<tag>_*)
tag=${1:-r3}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
nproc >> gpurun_out/${tag}_gpu.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err; echo "bench rc=$?"
for w in c1 c2 c3 c4; do
  timeout 900 python bench.py --workload $w --steps 12 --warmup 4 > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err; echo "bench $w rc=$?"
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k_fast|k_uniform|k_general|k_ordip|k_source|k_detector|k_emit' -s 40 -c 40 \
    --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 3 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_fast|k_uniform|k_ordip' -s 12 -c 6 \
    -o gpurun_out/${tag}_prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --nx 1024 --ny-per-gpu 128 --nz 512 > gpurun_out/${tag}_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_emit' -s 6 -c 3 \
    -o gpurun_out/${tag}_emit python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_emit.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_uniform' -s 4 -c 2 \
    -o gpurun_out/${tag}_c3 python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_c3.log 2>&1
# the reports stay on the box (tens of MB each): their per-kernel digests and the raw metric pages come back
for r in prof emit c3; do
  python tools/ncu_summary.py gpurun_out/${tag}_$r.ncu-rep > gpurun_out/${tag}_ncu_${r}_summary.txt 2>&1
  ncu -i gpurun_out/${tag}_$r.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_${r}_raw.csv 2>/dev/null
  rm -f gpurun_out/${tag}_$r.ncu-rep
done
ls -la gpurun_out | grep ${tag}_
