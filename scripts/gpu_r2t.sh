mkdir -p gpurun_out
( time python bench.py > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err ) 2> gpurun_out/r2t_bench.time
tail -n 3 gpurun_out/r2t_bench.time; tail -n 5 gpurun_out/r2t_bench.err; head -c 600 gpurun_out/r2t_bench.json; echo
( time python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2t_ref.json 2> gpurun_out/r2t_ref.err ) 2> gpurun_out/r2t_ref.time
tail -n 3 gpurun_out/r2t_ref.time; head -c 400 gpurun_out/r2t_ref.json; echo
bash scripts/gpu_r2p.sh > gpurun_out/r2t_prof.log 2>&1
tail -n 12 gpurun_out/r2t_prof.log
