mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r3g_tests.txt
python bench.py > gpurun_out/r3g_bench.json 2> gpurun_out/r3g_bench.err; tail -n 2 gpurun_out/r3g_bench.err
bash scripts/gpu_r2p.sh > gpurun_out/r3g_prof.log 2>&1
tail -n 14 gpurun_out/r3g_prof.log
