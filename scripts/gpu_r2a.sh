# round 2, call A: parity of the packed-fp32 (FFMA2) pooled kernels, bench, ray-order experiment, one ncu capture of each pooled kernel
mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2a_bench.err | tee gpurun_out/r2a_bench.json
python scripts/exp_sort.py 2>&1 | tee gpurun_out/r2a_sort.txt
B200_STREAMED=0 ncu --set full --clock-control none --import-source on -k regex:occluded_pool -s 2 -c 1 -o gpurun_out/r2a_prof_occ -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu_occ.log 2>&1
tail -2 gpurun_out/r2a_ncu_occ.log
