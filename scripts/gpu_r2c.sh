# round 2, call C: A/B of library variants on one box (scripts/build_variants.sh)
mkdir -p gpurun_out
for v in v0 v1 v2 v1 v0 v2; do
  echo "== variant $v"
  B200_LIB=$PWD/lucille_b200/variants/lib_$v.so ORDERS=batch,random python scripts/exp_sort.py 2>&1 | grep order
done | tee gpurun_out/r2c_ab.txt
