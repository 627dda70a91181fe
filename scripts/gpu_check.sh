mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/gpu_diag.py 65536 2>&1 | grep -v "lucille\]" | tee gpurun_out/diag.txt
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench1.err | tee gpurun_out/bench1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --points 65536 > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
