mkdir -p gpurun_out
compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitize_small.py > gpurun_out/r2w_memcheck.txt 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|hybrid|shade_trace|light_samples|done" gpurun_out/r2w_memcheck.txt | tail -12
compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/sanitize_small.py > gpurun_out/r2w_racecheck.txt 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|done" gpurun_out/r2w_racecheck.txt | tail -3
