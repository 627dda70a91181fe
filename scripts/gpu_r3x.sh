mkdir -p gpurun_out
for v in default q6 q10 q12 q16; do
  echo "== variant $v"
  if [ $v = default ]; then unset B200_LIB; else export B200_LIB=$PWD/lucille_b200/variants/lib_$v.so; fi
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-frames 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C3 value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
done 2>&1 | tee gpurun_out/r3x_refill.txt
