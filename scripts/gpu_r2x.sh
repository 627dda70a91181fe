mkdir -p gpurun_out
for N in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r2x_bench$N.json 2> gpurun_out/r2x_bench$N.err
  echo "N=$N rc=$?"; tail -n 1 gpurun_out/r2x_bench$N.err
done
