mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 2> gpurun_out/r2i_bench2.err | tee gpurun_out/r2i_bench2.json | cut -c1-3200
grep -v "lucille\] info" gpurun_out/r2i_bench2.err | tail -8
