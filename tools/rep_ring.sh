# repeated periodic-ring runs on 4 / 3 / 2 slabs (whole-column and automatic marching): any MISMATCH line is a race or a protocol error
cases=${CASES:-"pbc3d pbc3d_all"}
reps=${REPS:-6}
for n in 4 3 2; do for m in 1048576 0; do
  fails=0
  for i in $(seq 1 $reps); do
    if [ $m = 0 ]; then unset CHIML_B200_MARCH_NY; else export CHIML_B200_MARCH_NY=$m; fi
    out=$(timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port $((29800+i)) tests/slab_gpu_worker.py $cases 2>&1 | grep -c "MISMATCH\|rror")
    if [ "$out" != "0" ]; then fails=$((fails+1)); fi
  done
  echo "n=$n march=$m failing runs: $fails of $reps"
done; done
