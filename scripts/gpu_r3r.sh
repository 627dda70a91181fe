mkdir -p gpurun_out
python scripts/hybrid_rate.py 65536 exact inexact 2>&1 | grep -E "rays:|closest" | tee gpurun_out/r3r_hyb.txt
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "hybrid or city or rib_scene" 2>&1 | tail -1
python scripts/transport_rates.py 2>&1 | grep -E "whitted|dirt map 4x4 f64|hit mask|AO 8x8 f64  " | tee -a gpurun_out/r3r_hyb.txt
