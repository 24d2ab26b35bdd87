for n in 4 3 2; do for m in 1048576 0; do
  fails=0
  for i in 1 2 3 4 5 6; do
    if [ $m = 0 ]; then unset CHIML_B200_MARCH_NY; else export CHIML_B200_MARCH_NY=$m; fi
    out=$(timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port $((29800+i)) tests/slab_gpu_worker.py pbc3d pbc3d_all 2>&1 | grep -c "MISMATCH")
    if [ "$out" != "0" ]; then fails=$((fails+1)); fi
  done
  echo "n=$n march=$m failing runs: $fails of 6"
done; done
