mkdir -p gpurun_out
python scripts/transport_rates.py 2>&1 | grep -v "lucille\]" | tee gpurun_out/r3y_transports.txt
python scripts/hybrid_rate.py 65536 2>&1 | grep -v "lucille\]" | tee gpurun_out/r3y_hybrid.txt
python scripts/c5_rate.py 2>&1 | grep -E "C5 soup" | tee gpurun_out/r3y_c5.txt
