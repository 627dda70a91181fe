mkdir -p gpurun_out
./lucille_b200/variants/ubench_ffma2.bin | tee gpurun_out/r2f_ubench_ffma2.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2f_bench.err | tee gpurun_out/r2f_bench.json
B200_POOL32=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2f_bench_generic.err | tee gpurun_out/r2f_bench_generic.json
