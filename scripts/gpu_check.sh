mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/gpu_diag.py ${1:-65536} 2>&1 | grep -v "lucille\]" | tee gpurun_out/diag.txt
