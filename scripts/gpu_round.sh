# What the driver runs at round end, plus the profiler passes whose summaries go to profiles/.
# The ncu passes run with B200_STREAMED=0: ncu serialises the copy stream behind the kernel, so the streamed host-buffer path
# (one persistent launch polling the upload cursor) would only time out into its launch-per-piece fallback.
mkdir -p gpurun_out
set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --impl reference --steps 3 --warmup 3 2> gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json
python bench.py --steps 10 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json
B200_STREAMED=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
B200_STREAMED=0 ncu --set full --clock-control none --import-source on -k regex:occluded_pool -s 2 -c 1 -o gpurun_out/prof_bench python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
