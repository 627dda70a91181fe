mkdir -p gpurun_out
for v in p0 p1 p0 p1; do
  echo "== variant $v"
  B200_LIB=$PWD/lucille_b200/variants/lib_$v.so ORDERS=batch python scripts/exp_sort.py 2>&1 | grep order
done | tee gpurun_out/r2e_ab.txt
