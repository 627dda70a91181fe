mkdir -p gpurun_out
for v in default hp; do
  echo "== variant $v"
  if [ $v = default ]; then unset B200_LIB; else export B200_LIB=$PWD/lucille_b200/variants/lib_$v.so; fi
  python scripts/hybrid_rate.py 65536 exact inexact 2>&1 | grep -E "rays:"
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hybrid" 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r3q_hp.txt
bash scripts/gpu_final.sh
