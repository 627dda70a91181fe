mkdir -p gpurun_out
NTRIS=10000000 SEED=c5 POINTS=65536 python scripts/exp_sort.py 2>&1 | grep -E "^order|done" | tee gpurun_out/r3s_sort_c5.txt
