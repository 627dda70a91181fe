mkdir -p gpurun_out
for v in default hc6 hc4 hc7; do
  echo "== variant $v"
  if [ $v = default ]; then unset B200_LIB; else export B200_LIB=$PWD/lucille_b200/variants/lib_$v.so; fi
  python scripts/hybrid_rate.py 65536 exact 2>&1 | grep -E "closest hit"
done 2>&1 | tee gpurun_out/r3a_hc.txt
