mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r3c_tests.txt
python scripts/hybrid_rate.py 65536 2>&1 | grep -v "lucille\]" | tee gpurun_out/r3c_hybrid.txt
