mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2v_tests.txt
for v in default d128x5 d128x4 d128x6 d256x3; do
  echo "== variant $v"
  if [ $v = default ]; then unset B200_LIB; else export B200_LIB=$PWD/lucille_b200/variants/lib_$v.so; fi
  python scripts/gpu_diag.py 65536 2>&1 | grep -E "closest f64|anyhit  f64|closest f32|C2 closest"
done 2>&1 | tee gpurun_out/r2v_closest64.txt
unset B200_LIB
python scripts/transport_rates.py 2>&1 | grep -E "whitted|gather|dirt map 4x4 f64" | tee gpurun_out/r2v_transports.txt
