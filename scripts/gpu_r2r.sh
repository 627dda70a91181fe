mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "hybrid or occlusion_matches or axis_parallel or edge_case" 2>&1 | tail -15 | tee gpurun_out/r2r_tests.txt
python scripts/hybrid_rate.py 65536 2>&1 | grep -v "lucille\]" | tee gpurun_out/r2r_hybrid.txt
