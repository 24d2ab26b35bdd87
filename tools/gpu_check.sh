#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list of the same command and one full capture.
# usage: tools/gpu_check.sh <tag>      (outputs under gpurun_out/<tag>_*)
tag=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
nproc >> gpurun_out/${tag}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/${tag}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k_fast|k_uniform|k_general|k_ordip|k_source|k_detector|k_emit' -s 40 -c 40 \
    --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 3 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_fast|k_uniform|k_general|k_ordip' -s 16 -c 8 \
    -o gpurun_out/${tag}_prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --nx 1024 --ny-per-gpu 128 --nz 512 > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
