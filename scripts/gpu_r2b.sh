# round 2, call B: row-2-as-LDG.128 layout, quad-fetch knob x ray order, full-size parity tests, point API
mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/exp_sort.py 2>&1 | tee gpurun_out/r2b_sort.txt
B200_QUADFETCH=1 python scripts/exp_sort.py 2>&1 | tee gpurun_out/r2b_sort_quad.txt
B200_QUADFETCH=1 B200_REFILL=8 python scripts/exp_sort.py 2>&1 | tee gpurun_out/r2b_sort_quad8.txt
