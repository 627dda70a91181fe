mkdir -p gpurun_out
python bench.py --steps 6 --warmup 3 2> gpurun_out/r2h_bench1.err | tee gpurun_out/r2h_bench1.json | cut -c1-3000
tail -5 gpurun_out/r2h_bench1.err
