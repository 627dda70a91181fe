mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r3u_tests.txt
python scripts/hybrid_rate.py 65536 exact 2>&1 | grep -E "rays:|closest"
python scripts/transport_rates.py 2>&1 | grep -E "f64|whitted|hit mask|gather" | tee gpurun_out/r3u_transports.txt
