mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2s_tests.txt
python scripts/hybrid_rate.py 262144 exact inexact 2>&1 | grep -v "lucille\]" | tee gpurun_out/r2s_hybrid.txt
