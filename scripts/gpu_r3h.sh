mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "city or ragged or hybrid" 2>&1 | tail -8 | tee gpurun_out/r3h_tests.txt
