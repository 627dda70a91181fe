mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2z_tests.txt
python scripts/hybrid_rate.py 65536 2>&1 | grep -v "lucille\]" | tee gpurun_out/r2z_hybrid.txt
python scripts/transport_rates.py 2>&1 | grep -E "whitted|dirt map 4x4 f64|hit mask" | tee gpurun_out/r2z_transports.txt
